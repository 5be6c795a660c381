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
for rep in range(2):
    t0 = time.perf_counter(); Q.run_2d_crank_nicolson(**kw); print("run", rep, time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); Q.run_2d_crank_nicolson(**kw); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
