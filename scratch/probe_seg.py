"""x / y sweep launch times and the diffusion step on a segmented shape (full rectangle), for A/B runs with QPB_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scratch"))
import numpy as np
import qpsim_b200 as Q
from qpsim_b200 import capi
import cases
import _libswitch  # noqa: F401
for ny, nx, ne, dt, fmax in ((2048, 2048, 16, 0.2, 3.0), (1024, 1024, 32, 0.05, 10.0)):
    mask = np.ones((ny, nx), bool)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, fmax, ne)
    D = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
    edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    n = ny * nx
    st = np.exp(np.random.default_rng(3).standard_normal((ne, n)) * 0.3) * 1e-4
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(D); ctx.prepare_diffusion(0, dt); ctx.set_state(st)
        ctx.advance(2, dt)
        ctx.advance(3, dt); tot = ctx.diag()["last_advance_ms"] / 3
        ctx.enable_timers(True); ctx.reset_timers(); ctx.advance(3, dt); ctx.enable_timers(False)
        tx, nxl = ctx.timer(0); ty, nyl = ctx.timer(1)
    print(f"{ny}x{nx}x{ne}: step {tot:.3f} ms  x {1e3*tx/nxl:.1f} us x{nxl//3}/step  y {1e3*ty/nyl:.1f} us x{nyl//3}/step  other {tot-(tx+ty)/3:.3f} ms", flush=True)
