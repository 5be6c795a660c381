"""Validate the EXPECTATIONS of tests/test_gpu_behaviour.py by running them against the unmodified reference
(CPU, build container only)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from refimport import load_reference
qpsim = load_reference()
import qpsim.solver as S
import qpsim_b200 as Q
def ref(*a, **k):
    for extra in ("device", "diffusion_tolerance", "store_energy_frames"): k.pop(extra, None)
    return S.run_2d_crank_nicolson(*a, **k)
Q.run_2d_crank_nicolson = ref
import pytest
sys.exit(pytest.main(["-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_behaviour.py"), "-p", "no:cacheprovider"]))
