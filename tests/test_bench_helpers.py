"""Host-side pieces of bench.py that decide what the bench line claims: the Crank-Nicolson equation check used for the
full-size diffusion parity sample, and the bounded CPU baseline.  CPU only."""
import os
import sys

import numpy as np

import cases
import qpsim_b200 as Q
from oracle import qp_oracle as O

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_cn_equation_check_accepts_the_direct_solve_and_rejects_a_perturbed_one():
    case = cases.meander_c2(ny=40, nx=48, ne=3, steps=1, bc="mixed")
    mask = case["mask"]
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "mixed", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    D, dt = 4.3, 0.5
    u0 = np.where(mask, cases.lognormal_field(mask, 5), 0.0)
    op = O.DiffusionCN(mask, edges, bcs, 1.0, np.full((1, int(mask.sum())), D), dt, False)
    st = u0[mask][None, :].copy()
    op.step(st)
    u1 = np.zeros_like(u0)
    u1[mask] = st[0]
    ok = bench.cn_equations_error(mask, bcx, bcy, src, 0.5 * dt * D, dt * D, u0, u1)
    assert ok["componentwise"] < 1e-14 and ok["max_norm"] < 1e-14
    u1[mask] *= 1.0 + 1e-9
    bad = bench.cn_equations_error(mask, bcx, bcy, src, 0.5 * dt * D, dt * D, u0, u1)
    assert bad["componentwise"] > 1e-11


def test_cpu_baseline_is_a_labelled_extrapolation_of_a_fixed_sample(monkeypatch):
    monkeypatch.setattr(bench, "CPU_CELLS", 64)
    w = bench.c2_workload(ny=64, nx=64, ne=16)
    a = bench.cpu_baseline(w)
    assert a["extrapolated"] is True and a["kind"] == "port" and a["cores"] >= 1
    assert a["value"] > 0 and a["single_core_value"] > 0 and "fixed 64-cell sample" in a["sample"]
    assert a["single_core_value"] <= a["value"] * 1.5
