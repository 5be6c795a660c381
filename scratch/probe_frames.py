"""Timing of qpb_get_frames (64 MiB snapshot into a fresh numpy array) for the host-copy settings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from qpsim_b200 import capi
ny = nx = 256; ne = 128
mask = np.ones((ny, nx), dtype=bool); n = ny * nx
with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=0, dx=1.0, dE=1.0) as ctx:
    ctx.upload_geometry(mask); ctx.set_state(np.random.default_rng(0).random((ne, n)))
    for rep in range(6):
        t0 = time.perf_counter(); f = ctx.get_frames(); t1 = time.perf_counter()
        keep = f  # keep one alive so that the allocator cannot hand the same pages back
        if rep >= 2: print(f"get_frames {1e3*(t1-t0):.2f} ms  ({f.nbytes/1e6/(t1-t0)/1e3:.1f} GB/s)")
