"""BASELINE configs[4] member: 512 x 512 x 128 MKID mask through the public API; 3 members on one GPU."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, cases
import qpsim_b200 as Q
w = bench.c2_workload(ny=512, nx=512)
import cases as C
mask = C.meander_mask(512, 512, pad=16, slot=8, pitch=32, gap_len=64)
field = C.gaussian_field(mask, cx=0.4, cy=0.5, sigma=0.05, base=1e-4, amp=2e-4)
edges = Q.extract_edge_segments(mask); bcs = C.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
K = 20
base = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=field, diffusion_coefficient=C.D0, dt=0.5,
            total_time=0.5 * K, dx=1.0, store_every=K, energy_gap=C.GAP, energy_min_factor=1.0, energy_max_factor=5.0,
            num_energy_bins=128, enable_diffusion=True, enable_recombination=True, enable_scattering=True,
            dynes_gamma=C.GAMMA, tau_0=C.TAU, T_c=C.TC, bath_temperature=C.TBATH,
            external_generation=Q.ExternalGenerationSpec(mode="constant", rate=1e-8), store_energy_frames=False)
members = Q.parameter_grid(base, bath_temperature=[0.05, 0.175, 0.3])
warnings.simplefilter("ignore")
Q.run_2d_crank_nicolson(**{**members[0], "total_time": 1.0, "store_every": 2})
res = Q.run_ensemble(members)   # untimed pass (module load, allocator cache, iteration counts)
t0 = time.perf_counter()
res = Q.run_ensemble(members)
dt = time.perf_counter() - t0
n = int(mask.sum())
print(f"C5 members: {len(members)} runs of {n} cells x 128 bins x {K} steps in {dt:.3f} s -> {dt/len(members)*1e3:.1f} ms per run, "
      f"{n*128*K*len(members)/dt/1e9:.3f} G updates/s end to end; masses {[round(r[1][-1], 6) for r in res]}")
import cProfile, pstats
from qpsim_b200 import capi
acc = {}
def wrap(name):
    f = getattr(capi.Context, name)
    def g(self, *a, **k):
        t0 = time.perf_counter(); r = f(self, *a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(capi.Context, name, g)
for nm in ("__init__", "close", "upload_geometry", "upload_diffusion", "prepare_diffusion", "upload_collision", "set_state", "set_state_uniform_phonons", "get_state", "get_frames", "get_integrated", "advance", "pauli"):
    wrap(nm)
pr = cProfile.Profile(); pr.enable(); Q.run_2d_crank_nicolson(**members[1]); pr.disable()
print({k: round(v, 4) for k, v in acc.items()}, Q.solver.last_run_info)
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
