"""The sharded multi-GPU loop on real ranks (needs >= 2 B200 on the box; run under `gpurun --gpus 2|4`):
run_2d_crank_nicolson(devices=[...]) - bins <-> cells with the exchange fused into the collision kernel over cudaIpc
peer memory, and the NCCL all-to-all path - against the single-context result (1e-12) and the reference fixtures
(1e-9).  World sizes 2 and 3 (uneven cell and bin cuts) when the GPUs are there; plus the spawn mode of the drop-in."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import helpers
import qpsim_b200 as Q

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    return Q.capi.load_library().qpb_device_count()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_dropin_on_real_ranks(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(HERE, "mp_sharded_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"SHARDED-OK world={world}" in res.stdout, res.stdout[-3000:]
    print([ln for ln in res.stdout.splitlines() if ln.startswith("SHARDED-OK")][0])


def test_dropin_spawns_its_ranks_from_a_plain_process():
    """devices=[0, 1] from an ordinary (non-torchrun) caller: the workers are spawned for the call."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    case = cases.meander_c2(ny=40, nx=40, ne=12, steps=2)
    want = helpers.load_golden(case["name"])
    got = helpers.run_dropin(case, devices=[0, 1])
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-9)
    helpers.assert_close(got["phonons"], want["phonons"], "n_ph", rtol=helpers.RTOL_PHONON)


def test_devices_with_one_entry_is_the_single_gpu_path():
    case = cases.meander_c2(ny=40, nx=40, ne=12, steps=2)
    a = helpers.run_dropin(case, devices=[0])
    b = helpers.run_dropin(case)
    assert np.array_equal(a["state"], b["state"])
