"""Short driver for ncu: C2 workload, a few device-resident steps (no CPU baseline, no e2e)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = bench.c2_workload(); tabs = bench.build_tables(w, Q)
mask = w["mask"]; ny, nx = mask.shape; n, ne, nw = tabs["n"], w["num_energy_bins"], tabs["omega"].size
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
flags = capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI
with capi.Context(ny=ny, nx=nx, ne=ne, nw=nw, ncell=n, flags=flags, dx=w["dx"], dE=tabs["dE"]) as ctx:
    ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(tabs["D"]); ctx.prepare_diffusion(0, w["dt"])
    ctx.upload_collision(tabs["Kr"][None], tabs["Ks"][None], tabs["rho"][None], None, tabs["idx_diff"], tabs["idx_sum"], tabs["sign"])
    ctx.set_state(tabs["state"], tabs["phonons"])
    ctx.advance(steps, w["dt"], t_start=0.0, want_pauli=True, gen_mode=capi.GEN_PULSE, rate=w["pulse_rate"], pulse_start=0.0, pulse_duration=5.0)
    d = ctx.diag()
    print("ms/step", d["last_advance_ms"] / steps, "sweeps", d["sweeps"] / steps, "launches", d["kernel_launches"])
