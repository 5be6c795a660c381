"""Host time of the per-rank setup of the sharded C3 run (rank 0 of 8), call by call, on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi, multigpu
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = [time.perf_counter()]
def lap(what):
    T.append(time.perf_counter()); print(f"{what:40s} {1e3 * (T[-1] - T[-2]):8.1f} ms", flush=True)
w = bench.c3_workload(); lap("workload (mask, field)")
mask = w["mask"]; ny, nx = mask.shape
edges = Q.extract_edge_segments(mask); lap("extract_edge_segments")
bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"]); lap("compile_boundaries")
tabs = bench.build_tables(w, Q, want_state=False); lap("build_tables")
n, ne, nw = tabs["n"], w["num_energy_bins"], tabs["omega"].size
plan = multigpu.ShardPlan(ne, n, world, 0, interleave=True)
c0, c1 = plan.cells(); bins = plan.bins(); nloc = c1 - c0
import torch
torch.cuda.set_device(0); torch.zeros(1, device="cuda"); lap("torch cuda init")
ctx_d = capi.Context(ny=ny, nx=nx, ne=bins.size, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=w["dx"], dE=tabs["dE"]); lap("create diffusion ctx")
ctx_d.upload_geometry(mask, bcx, bcy, src); lap("  upload_geometry")
ctx_d.upload_diffusion(tabs["D"][bins]); lap("  upload_diffusion")
ctx_d.prepare_diffusion(0, w["dt"]); lap("  prepare_diffusion")
fl = capi.F_PAULI | capi.F_SCATTERING | capi.F_RECOMBINATION
ctx_c = capi.Context(ny=1, nx=nloc, ne=ne, nw=nw, ncell=nloc, ngap=1, flags=fl, dx=w["dx"], dE=tabs["dE"]); lap("create collision ctx")
ctx_c.upload_geometry(np.ones((1, nloc), dtype=np.uint8)); lap("  upload_geometry (strip)")
ctx_c.upload_collision(tabs["Kr"][None], tabs["Ks"][None], tabs["rho"][None], None, tabs["idx_diff"], tabs["idx_sum"], tabs["sign"]); lap("  upload_collision")
ctx_c.set_state_separable(tabs["weights"], w["initial_field"][mask][c0:c1], tabs["phonon_bins"]); lap("  set_state_separable")
ctx_c.close(); ctx_d.close(); lap("close")
print("total", 1e3 * (T[-1] - T[0]), "ms")
