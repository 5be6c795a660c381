"""Kernel timers on the C2 workload (device resident): collision ms/launch, x/y sweep ms/launch, step time.
Environment switches of the library (QPB_COLL_TJ, ...) are read at launch time, so one process can compare variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
variants = [v for v in sys.argv[1:]] or [""]
import _libswitch  # noqa: F401  (QPB_LIB=... selects another build of the library)
w = bench.c2_workload(); tabs = bench.build_tables(w, Q)
mask = w["mask"]; ny, nx = mask.shape; n, ne, nw = tabs["n"], w["num_energy_bins"], tabs["omega"].size
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
flags = capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI
ref = None
for var in variants:
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=nw, ncell=n, flags=flags, dx=w["dx"], dE=tabs["dE"]) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(tabs["D"]); ctx.prepare_diffusion(0, w["dt"])
        ctx.upload_collision(tabs["Kr"][None], tabs["Ks"][None], tabs["rho"][None], None, tabs["idx_diff"], tabs["idx_sum"], tabs["sign"])
        ctx.set_state(tabs["state"], tabs["phonons"])
        kw = dict(t_start=0.0, want_pauli=True, gen_mode=capi.GEN_PULSE, rate=w["pulse_rate"], pulse_start=0.0, pulse_duration=5.0)
        ctx.advance(3, w["dt"], **kw)
        ctx.enable_timers(True); ctx.reset_timers()
        ctx.advance(4, w["dt"], **kw)
        ctx.enable_timers(False)
        tx, nxl = ctx.timer(0); ty, nyl = ctx.timer(1); tc, ncl = ctx.timer(2)
        ctx.advance(10, w["dt"], **kw)
        d = ctx.diag()
        st, ph = ctx.get_state()
        if ref is None: ref = (st.copy(), ph.copy())
        es = np.max(np.abs(st - ref[0])) / np.max(np.abs(ref[0])); ep = np.max(np.abs(ph - ref[1]) / np.maximum(np.abs(ref[1]), 1e-300))
        print(f"[{var}] step {d['last_advance_ms']/10:.3f} ms  collide {tc/max(ncl,1):.4f} ms x{ncl}  sweep_x {1e3*tx/max(nxl,1):.1f} us x{nxl}  sweep_y {1e3*ty/max(nyl,1):.1f} us x{nyl}  diff-vs-first: n {es:.1e} n_ph {ep:.1e}", flush=True)
    for kv in var.split(","):
        if "=" in kv: os.environ.pop(kv.split("=")[0], None)
