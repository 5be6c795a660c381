"""Time-dependent custom generation at the C2 shape: device program (QPB_GEN_PROGRAM) against host evaluation + one
NE x N upload per step (QPB_NO_GEN_PROGRAM=1), through the drop-in, 10 steps."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
w = bench.c2_workload()
mask = w["mask"]
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
kw = bench.solver_kwargs(w, Q, 10, edges, bcs)
kw["external_generation"] = Q.ExternalGenerationSpec(
    mode="custom", custom_body="params['a'] * np.exp(-t / 2.0) * np.where(E < 400.0, 1.0, 0.25) * (0.5 + y * x)",
    custom_params={"a": 3e-8})
kw["store_energy_frames"] = False
out = {}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name, env in (("device program", "0"), ("host evaluation + upload per step", "1")):
        os.environ["QPB_NO_GEN_PROGRAM"] = env
        Q.run_2d_crank_nicolson(**{**kw, "total_time": w["dt"] * 2})
        t0 = time.perf_counter()
        res = Q.run_2d_crank_nicolson(**kw)
        dt = time.perf_counter() - t0
        info = dict(Q.solver.last_run_info)
        out[name] = np.array(res[2])
        print(f"{name}: {dt:.3f} s for 10 steps, uploads {info['generation_uploads']}, on device {info['generation_on_device']}", flush=True)
a, b = out.values()
print("mass histories agree to", float(np.max(np.abs(a - b) / np.abs(b))))
