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
    # what was actually timed rides along: the reference arm reports it as its step time
    assert 0 < a["sample_step_s"] <= a["sample_wall_s"]
    assert a["est_ms_per_step"] > 1e3 * a["sample_step_s"] * 0.5   # the sample is a small part of the workload


def test_reference_arm_line_reports_the_measured_sample_step(monkeypatch, capsys):
    """`bench.py --impl reference`: ms_per_step is the measured time of a pass over the bounded sample, the time scaled
    to the whole workload is est_ms_per_step, and value = updates_per_sampled_step / step time."""
    import argparse
    import json

    monkeypatch.setattr(bench, "CPU_CELLS", 64)
    monkeypatch.setattr(bench, "c3_workload", lambda: dict(bench.c2_workload(ny=64, nx=64, ne=16), name="miniature"))
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(argparse.Namespace(workload="c3", gpus=1, steps=1, warmup=0))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["extrapolated"] is True and line["steps_timed"] == 1
    assert line["e2e"]["value"] == line["value"] and line["cpu_baseline"]["value"] == line["value"]
    per_step = line["config"]["updates_per_sampled_step"]
    assert abs(per_step / (line["ms_per_step"] * 1e-3) - line["value"]) <= 1e-9 * line["value"]
    n, ne = line["config"]["cells"], line["config"]["energy_bins"]
    assert abs(n * ne / (line["est_ms_per_step"] * 1e-3) - line["value"]) <= 1e-9 * line["value"]


def test_c3_roofline_traffic_is_scaled_from_the_committed_slice_captures():
    """roofline.traffic of the C3 lines: the committed ncu captures were taken on a slice (37 888 cells / 16 bins); the
    bytes are scaled to the launch of the rank and the line says where they come from."""
    full = bench.load_profile_traffic("c3", 4194304, 256)
    half = bench.load_profile_traffic("c3", 2097152, 128)
    assert full["collide"] > 0 and full["sweep"] > 0
    assert abs(full["collide"] - 2 * half["collide"]) <= 1e-9 * full["collide"]
    assert abs(full["sweep"] - 2 * half["sweep"]) <= 1e-9 * full["sweep"]
    assert "scaled" in full["collide_source"] and "not this run" in full["sweep_source"]
    # 64 NE bytes per cell and call is the algorithmic figure; the capture sits below it (L2 hits)
    assert 0.5 * 64 * 256 * 4194304 < full["collide"] < 1.2 * 64 * 256 * 4194304
    c2 = bench.load_profile_traffic("c2")
    assert c2["collide"] > 0 and c2["sweep"] > 0 and "collide_source" not in c2
