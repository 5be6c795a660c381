import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, warnings
import bench, cases
import qpsim_b200 as Q
w = bench.c2_workload(); mask = w["mask"]
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
K = 20
gen = Q.ExternalGenerationSpec(mode="pulse", pulse_rate=w["pulse_rate"], pulse_start=0.0, pulse_duration=5.0)
kw = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=w["initial_field"], diffusion_coefficient=w["diffusion_coefficient"],
          dt=w["dt"], total_time=w["dt"] * K, dx=w["dx"], store_every=K, energy_gap=w["energy_gap"], energy_min_factor=1.0,
          energy_max_factor=w["energy_max_factor"], num_energy_bins=128, enable_diffusion=True, enable_recombination=True,
          enable_scattering=True, dynes_gamma=w["dynes_gamma"], tau_0=w["tau_0"], T_c=w["T_c"], bath_temperature=w["bath_temperature"],
          external_generation=gen)
warnings.simplefilter("ignore")
from qpsim_b200 import capi
acc = {}
def wrap(name):
    f = getattr(capi.Context, name)
    def g(self, *a, **k):
        t0 = time.perf_counter(); r = f(self, *a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(capi.Context, name, g)
for nm in ("__init__", "close", "upload_geometry", "upload_diffusion", "prepare_diffusion", "upload_collision", "set_state", "set_state_separable", "set_state_uniform_phonons", "get_state", "get_frames", "get_integrated", "advance", "pauli"):
    wrap(nm)
for rep in range(4):
    acc.clear()
    t0 = time.perf_counter(); Q.run_2d_crank_nicolson(**kw); print("run", rep, round(time.perf_counter() - t0, 4), {k: round(v, 4) for k, v in acc.items()})
pr = cProfile.Profile(); pr.enable(); Q.run_2d_crank_nicolson(**kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
