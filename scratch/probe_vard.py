"""Per-direction sweep times of the variable-D (non-uniform gap) diffusion against uniform D on the same grid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
side, ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 128
w = bench.c2_workload(ny=side, nx=side, ne=ne)
mask = w["mask"]; ny, nx = mask.shape; n = int(mask.sum())
E, dE = Q.build_energy_grid(w["energy_gap"], 1.0, 5.0, ne)
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
yy, xx = np.mgrid[:ny, :nx]
gap = (cases.GAP * (1.0 - 0.1 * xx / nx))[mask]
Dv = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - np.minimum(gap[None, :] / E[:, None], 1.0) ** 2))
Du = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
rho = Q.density_of_states(E, cases.GAP, cases.GAMMA); wts = rho / (rho.sum() * dE)
state = wts[:, None] * w["initial_field"][mask][None, :]
for name, flags, D in (("uniform D", capi.F_DIFFUSION, Du), ("variable D", capi.F_DIFFUSION | capi.F_VARIABLE_D, Dv)):
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=flags, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(D); ctx.prepare_diffusion(0, w["dt"])
        ctx.set_state(state)
        ctx.advance(2, w["dt"])
        ctx.enable_timers(True); ctx.reset_timers()
        d0 = ctx.diag()
        ctx.advance(2, w["dt"])
        d1 = ctx.diag()
        tx, nxl = ctx.timer(0); ty, nyl = ctx.timer(1)
        print(f"{side}x{side}x{ne} {name}: x {tx / max(1, nxl) * 1e3:.1f} us x {nxl}, y {ty / max(1, nyl) * 1e3:.1f} us x {nyl}, "
              f"sweeps/step {(d1['sweeps'] - d0['sweeps']) / 2:.0f}, path {d1['sweep_path']}, step {d1['last_advance_ms'] / 2:.2f} ms", flush=True)
