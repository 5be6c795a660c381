"""Diffusion step of the C2 workload (or a side x side x ne variant): bin-resident cluster solve against the launched
pipelined sweeps (QPB_NO_RESIDENT=1), time per step and the difference of the two results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
side, ne = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = bench.c2_workload(ny=side, nx=side, ne=ne)
mask = w["mask"]; ny, nx = mask.shape; n = int(mask.sum())
E, dE = Q.build_energy_grid(w["energy_gap"], 1.0, 5.0, ne)
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
Du = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
rho = Q.density_of_states(E, cases.GAP, cases.GAMMA); wts = rho / (rho.sum() * dE)
state = wts[:, None] * w["initial_field"][mask][None, :]
out = {}
for name, env in (("resident", "0"), ("launched sweeps", "1")):
    os.environ["QPB_NO_RESIDENT"] = env
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(Du); ctx.prepare_diffusion(0, w["dt"])
        ctx.set_state(state)
        ctx.advance(2, w["dt"])
        d0 = ctx.diag()
        ctx.advance(steps, w["dt"])
        d1 = ctx.diag()
        out[name] = ctx.get_state(want_phonons=False)[0]
        print(f"{side}x{side}x{ne} {name}: path {d1['sweep_path']}, {d1['last_advance_ms'] / steps * 1e3:.1f} us per step, "
              f"bin-sweeps/step {(d1['bin_sweeps'] - d0['bin_sweeps']) / steps:.0f}, launches/step "
              f"{(d1['kernel_launches'] - d0['kernel_launches']) / steps:.1f}", flush=True)
a, b = out["resident"], out["launched sweeps"]
print("max |resident - launched| / max|u| per bin:", float(np.max(np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1))))
