"""Run under torchrun by tests/test_gpu_multi.py (one rank per GPU): the drop-in with devices=[0..P-1] against the
single-context result and the reference fixture, on both exchange paths (fused into the collision kernel over
cudaIpc peer memory / NCCL all-to-all).  Prints one line 'SHARDED-OK ...' from rank 0 when everything agrees."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import helpers  # noqa: E402
import qpsim_b200 as Q  # noqa: E402
from qpsim_b200 import multigpu  # noqa: E402


def main():
    import torch.distributed as dist

    rank, world, local = multigpu.init_process_group("nccl")
    devices = list(range(world))
    report = []
    runs = [("c2_meander_40x40x12", next(c for c in cases.golden_cases() if c["name"] == "c2_meander_40x40x12")),
            ("annulus_32x5_mixed", next(c for c in cases.golden_cases() if c["name"] == "annulus_32x5_mixed")),
            ("nonuniform_10x14x8", next(c for c in cases.golden_cases() if c["name"] == "nonuniform_10x14x8")),
            ("c2_meander_96x80x24", cases.meander_c2(ny=96, nx=80, ne=24, steps=3)),
            # rows of 128 cells: every rank's diffusion context solves its bins bin-resident (k_pr_resident) on the state
            # the other ranks' collision kernels have stored into
            ("c2_meander_64x128x10", cases.meander_c2(ny=64, nx=128, ne=10, steps=3))]
    runs += [(c["name"], c) for c in cases.custom_mode_cases() if c["name"] in ("custom_gen_static", "custom_gen_timedep")]
    for name, case in runs:
        single = helpers.run_dropin(case, device=local) if rank == 0 else None
        for fused in (True, False):
            os.environ["QPB_NO_FUSED_EXCHANGE"] = "0" if fused else "1"
            got = helpers.run_dropin(case, devices=devices)
            info = dict(Q.solver.last_run_info)
            if rank != 0:
                continue
            assert info["world"] == world and info["exchanges"] > 0 or not case["enable_diffusion"]
            if name == "c2_meander_64x128x10":
                assert info["sweep_path"] == 5, info
            if fused and case["enable_recombination"] and "gap_values" not in case:
                assert info["fused_exchange"], info
            e_single = helpers.rel_err(got["state"], single["state"])
            assert e_single <= 1e-12, (name, fused, e_single)
            np.testing.assert_allclose(got["mass"], single["mass"], rtol=1e-12)
            if "phonons" in single:
                assert helpers.rel_err(got["phonons"], single["phonons"]) <= 1e-12
            e_gold = None
            try:
                want = helpers.load_golden(name)
                e_gold = helpers.rel_err(got["state"], want["state"])
                assert e_gold <= 1e-9, (name, fused, e_gold)
            except FileNotFoundError:
                pass
            report.append(f"{name}:{'fused' if fused else 'a2a'}:{e_single:.1e}:{e_gold if e_gold is None else format(e_gold, '.1e')}")
    dist.barrier()
    if rank == 0:
        print(f"SHARDED-OK world={world} " + " ".join(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
