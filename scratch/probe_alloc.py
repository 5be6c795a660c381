import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from qpsim_b200 import capi
for rep in range(4):
    t0 = time.perf_counter()
    ctx = capi.Context(ny=512, nx=512, ne=128, nw=320, ncell=183808, flags=capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI, dx=1.0, dE=1.0)
    t1 = time.perf_counter()
    ctx.close()
    t2 = time.perf_counter()
    print(f"rep {rep}: create {t1-t0:.4f} s, destroy {t2-t1:.4f} s", flush=True)
