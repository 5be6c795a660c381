"""Per-kernel event times of the direct spectral CN solve on the 2048 x 2048 grid at nb bins (timer 0: the two
transforms, timer 1: the Thomas pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = bench.c3_workload(); tabs = bench.build_tables(w, Q, want_state=False)
mask = w["mask"]; ny, nx = mask.shape
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
sel = np.linspace(0, w["num_energy_bins"] - 1, nb).astype(int)
with capi.Context(ny=ny, nx=nx, ne=nb, nw=0, ncell=ny * nx, flags=capi.F_DIFFUSION, dx=1.0, dE=tabs["dE"]) as ctx:
    ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(tabs["D"][sel]); ctx.prepare_diffusion(0, w["dt"])
    ctx.set_state_separable(tabs["weights"][sel], w["initial_field"][mask], None)
    ctx.advance(2, w["dt"])
    ctx.enable_timers(True); ctx.reset_timers()
    ctx.advance(3, w["dt"])
    t0, n0 = ctx.timer(0); t1, n1 = ctx.timer(1)
    gb = 8.0 * ny * nx * nb / 1e9
    print(f"2048^2 x {nb} bins: transforms {t0 / n0:.2f} ms per launch ({2 * gb / (t0 / n0 * 1e-3) / 1e3:.2f} TB/s on 16 B per cell), "
          f"Thomas pass {t1 / n1:.2f} ms ({4 * gb / (t1 / n1 * 1e-3) / 1e3:.2f} TB/s on 32 B per cell), step {ctx.diag()['last_advance_ms'] / 3:.2f} ms")
