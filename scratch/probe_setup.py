"""Host time of qpb_prepare_diffusion at the C2 shape with and without the bin-resident plan, and of one 20-step drop-in call."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
w = bench.c2_workload()
mask = w["mask"]; ny, nx = mask.shape; n = int(mask.sum()); ne = w["num_energy_bins"]
E, dE = Q.build_energy_grid(w["energy_gap"], 1.0, 5.0, ne)
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
Du = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
for env in ("0", "1", "0", "1"):
    os.environ["QPB_NO_RESIDENT"] = env
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(Du)
        t0 = time.perf_counter(); ctx.prepare_diffusion(0, w["dt"]); t1 = time.perf_counter()
        print(f"QPB_NO_RESIDENT={env}: prepare_diffusion {1e3 * (t1 - t0):.2f} ms, path {ctx.diag()['sweep_path']}", flush=True)
kw = bench.solver_kwargs(w, Q, 20, edges, bcs)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for env in ("0", "1", "0", "1"):
        os.environ["QPB_NO_RESIDENT"] = env
        Q.run_2d_crank_nicolson(**{**kw, "total_time": w["dt"] * 3, "store_every": 3})
        t0 = time.perf_counter(); Q.run_2d_crank_nicolson(**kw); t1 = time.perf_counter()
        print(f"QPB_NO_RESIDENT={env}: 20-step call {1e3 * (t1 - t0):.1f} ms", flush=True)
