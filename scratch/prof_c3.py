"""Short driver for ncu at the C3 shapes (BASELINE configs[2]) without the whole bench: the direct spectral CN solve of
16 of the 256 bins on the 2048 x 2048 grid, and the collision half step at 256 bins on a block of cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi

w = bench.c3_workload()
tabs = bench.build_tables(w, Q, want_state=False)
mask = w["mask"]; ny, nx = mask.shape
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
nb = 16
sel = np.linspace(0, w["num_energy_bins"] - 1, nb).astype(int)
with capi.Context(ny=ny, nx=nx, ne=nb, nw=0, ncell=ny * nx, flags=capi.F_DIFFUSION, dx=1.0, dE=tabs["dE"]) as ctx:
    ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(tabs["D"][sel]); ctx.prepare_diffusion(0, w["dt"])
    ctx.set_state(tabs["weights"][sel][:, None] * w["initial_field"][mask][None, :])
    ctx.advance(2, w["dt"])
    d = ctx.diag()
    print("diffusion 2048^2 x 16 bins: ms/step", d["last_advance_ms"] / 2, "sweep_path", d["sweep_path"])
ne = w["num_energy_bins"]; nw = tabs["omega"].size; ncell = 148 * 16 * 8
rng = np.random.default_rng(1)
state = tabs["weights"][:, None] * (1e-4 * np.exp(0.5 * rng.standard_normal((1, ncell))))
with capi.Context(ny=1, nx=ncell, ne=ne, nw=nw, ncell=ncell, flags=capi.F_SCATTERING | capi.F_RECOMBINATION,
                  dx=1.0, dE=tabs["dE"]) as ctx:
    ctx.upload_geometry(np.ones((1, ncell), np.uint8))
    ctx.upload_collision(tabs["Kr"][None], tabs["Ks"][None], tabs["rho"][None], None, tabs["idx_diff"], tabs["idx_sum"], tabs["sign"])
    ctx.set_state_uniform_phonons(state, tabs["phonon_bins"])
    ctx.enable_timers(True)
    for _ in range(3):
        ctx.collide(0.1)
    ms, nl = ctx.timer(2)
    print(f"collision NE=256, {ncell} cells: {ms / nl:.3f} ms per call, {21.0 * ne * ne * ncell / (ms / nl * 1e-3) / 1e12:.2f} TFLOP/s")
