"""BASELINE configs[4]: parameter-sweep ensemble of 64 independent 512 x 512 x 128-bin MKID runs dealt over the ranks of
a torchrun job (one process per GPU, replicas only - no data-path collective).  Parameters from a seeded grid over
T_bath in [0.05, 0.3], tau_0 in [100, 800], D0 in [2, 10], generation rate in [0, 1e-7].
  torchrun --nproc-per-node 8 scratch/run_c5.py        prints one JSON line from rank 0"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases as C
import qpsim_b200 as Q
from qpsim_b200 import multigpu
import torch, torch.distributed as dist

rank, world, local = multigpu.init_process_group("nccl")
mask = C.meander_mask(512, 512, pad=16, slot=8, pitch=32, gap_len=64)
field = C.gaussian_field(mask, cx=0.4, cy=0.5, sigma=0.05, base=1e-4, amp=2e-4)
edges = Q.extract_edge_segments(mask); bcs = C.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
K = int(os.environ.get("C5_STEPS", "20")); NM = int(os.environ.get("C5_MEMBERS", "64"))
base = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=field, dt=0.5, total_time=0.5 * K, dx=1.0,
            store_every=K, energy_gap=C.GAP, energy_min_factor=1.0, energy_max_factor=5.0, num_energy_bins=128,
            enable_diffusion=True, enable_recombination=True, enable_scattering=True, dynes_gamma=C.GAMMA, T_c=C.TC,
            store_energy_frames=False)
rng = np.random.default_rng(20260105)
members = [dict(base, bath_temperature=float(rng.uniform(0.05, 0.3)), tau_0=float(rng.uniform(100, 800)),
                diffusion_coefficient=float(rng.uniform(2, 10)),
                external_generation=Q.ExternalGenerationSpec(mode="constant", rate=float(rng.uniform(0, 1e-7))))
           for _ in range(NM)]
warnings.simplefilter("ignore")
Q.run_2d_crank_nicolson(**{**members[rank], "total_time": 1.0, "store_every": 2, "device": local})   # warm-up
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
res = Q.run_ensemble(members, device=local)
dist.barrier()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
n = int(mask.sum())
if rank == 0:
    masses = [float(r[1][-1]) for r in res]
    print(json.dumps({"workload": f"C5 ensemble: {NM} x (512x512 meander, {n} cells x 128 bins), {K} steps each, end to end "
                      "through run_ensemble / run_2d_crank_nicolson (setup, uploads, steps, integrated frames back)",
                      "n_gpus": world, "seconds": float(dt.item()), "members_per_second": NM / float(dt.item()),
                      "updates_per_s": n * 128 * K * NM / float(dt.item()), "all_finite": bool(np.all(np.isfinite(masses))),
                      "mass_min_max": [min(masses), max(masses)]}), flush=True)
dist.destroy_process_group()
