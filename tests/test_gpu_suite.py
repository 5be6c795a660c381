"""CUDA drop-in against recordings of the UNMODIFIED reference at the sizes and on the inputs the reference itself
tests (north_star: "a bit-for-tolerance match to qpsim.solver on all data/test_cases"):

* all 28 runs of ``generate_test_suite()`` (qpsim/test_cases.py:1133-1178): 10 strip, 9 rectangle, 4 polygon donut,
  3 recombination, 2 scattering cases (tests/golden/suite_cases.npz);
* BASELINE configs[1] at full size, 256 x 256 meander x 128 bins, two steps (tests/golden/c2_full_256x256x128.npz);
* BASELINE configs[0], 128 cells x 64 bins, 20 steps; the input of the reference's tests/test_mkid_crosscheck.py.

Bar: element-wise 1e-9 (helpers.rel_err) on n(x,y,E) and 1e-9 on the total number.
"""
import os
import sys

import numpy as np
import pytest

import cases
import helpers

pytestmark = pytest.mark.gpu

SUITE_IDS = helpers.suite_case_ids()


@pytest.mark.parametrize("k", range(len(SUITE_IDS)), ids=SUITE_IDS)
def test_dropin_matches_reference_validation_suite(k):
    kw, want = helpers.load_suite_case(k)
    got = helpers.run_suite_case_dropin(kw, want["keep"])
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-9)
    helpers.assert_close(got["state"], want["state"], "field")
    # total number to 1e-9 of the number of quasiparticles present (the antisymmetric eigenmodes of the suite sum to
    # zero up to rounding: a tolerance relative to that sum itself would compare noise)
    ne = want["state"].shape[1]
    dE = 1.0 if ne == 1 or kw["energy_gap"] <= 0 else (kw["energy_max_factor"] - kw["energy_min_factor"]) * kw["energy_gap"] / ne
    present = float(np.max(np.sum(np.abs(want["state"]), axis=(1, 2)))) * dE * kw["dx"] ** 2
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL, atol=helpers.RTOL * present)


def test_c2_full_size_two_steps_match_reference():
    """256 x 256 x 128 exactly as bench.py runs it; the reference's answer after two full steps."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    want = helpers.load_golden("c2_full_256x256x128")
    case = bench.c2_case(steps=2)
    got = helpers.run_dropin(case)
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    bins, cells, ph_cells = want["bins"], want["cells"], want["ph_cells"]
    helpers.assert_close(got["state"][:, bins, :], want["state_bins"], "n(E,cell), all cells of 4 bins")
    helpers.assert_close(got["state"][:, :, cells], want["state_cells"], "n(E,cell), all bins of 2048 cells")
    dE = (case["energy_max_factor"] - case["energy_min_factor"]) * case["energy_gap"] / case["num_energy_bins"]
    helpers.assert_close((got["state"].sum(axis=1) * dE)[:, None, :], want["integrated"][:, None, :], "integrated field")
    helpers.assert_close(got["phonons"][:, :, ph_cells], want["phonons_cells"], "n_ph", rtol=helpers.RTOL_PHONON)


@pytest.mark.parametrize("name", ["c1_strip_128x64_20steps", "mkid_crosscheck_48x12"])
def test_strip_runs_match_reference(name):
    if name.startswith("c1"):
        case = cases.strip_c1(steps=20, nx=128, ne=64)
    else:
        case = cases.strip_c1(steps=12, nx=48, ne=12)
        case.update(tau_0=400.0, store_every=1)
    want = helpers.load_golden(name)
    got = helpers.run_dropin(case)
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-12)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    helpers.assert_close(got["phonons"], want["phonons"], "n_ph", rtol=helpers.RTOL_PHONON)


CUSTOM = cases.custom_mode_cases()


@pytest.mark.parametrize("case", CUSTOM, ids=[c["name"] for c in CUSTOM])
def test_user_expression_options_match_reference(case):
    """Custom generation bodies (resident when time independent, evaluated on the device otherwise), initial-condition specs
    and gap expressions through the drop-in, against the unmodified reference (solver.py:918-962, 1094-1124,
    1186-1196; reference tests: tests/test_initial_condition_split.py:144-173, tests/test_regressions.py:435-499)."""
    import qpsim_b200 as Q

    gold = helpers.load_golden("custom_modes")
    p = case["name"] + "/"
    got = helpers.run_dropin(case)
    np.testing.assert_allclose(got["times"], gold[p + "times"], rtol=0, atol=1e-12)
    helpers.assert_close(got["state"], gold[p + "state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], gold[p + "mass"], rtol=helpers.RTOL)
    helpers.assert_close(got["phonons"], gold[p + "phonons"], "n_ph", rtol=helpers.RTOL_PHONON)
    info = Q.solver.last_run_info
    if case["name"] == "custom_gen_static":
        assert info["generation_uploads"] == 1          # evaluated and uploaded once, resident afterwards
    if case["name"] == "custom_gen_timedep":
        # the body runs on the device at the time of every step: no host evaluation, no NE x N upload
        assert info["generation_uploads"] == 0 and info["steps_done"] == 6 and info["generation_on_device"]


def test_time_dependent_custom_generation_on_the_device_and_on_the_host_agree(monkeypatch):
    """The same time-dependent body translated into a device program (QPB_GEN_PROGRAM, whole batches of steps) and
    evaluated on the host with one upload per step (QPB_NO_GEN_PROGRAM=1): same run to rounding of the libm calls."""
    import qpsim_b200 as Q
    case = [c for c in CUSTOM if c["name"] == "custom_gen_timedep"][0]
    a = helpers.run_dropin(case)
    assert Q.solver.last_run_info["generation_on_device"]
    monkeypatch.setenv("QPB_NO_GEN_PROGRAM", "1")
    b = helpers.run_dropin(case)
    assert not Q.solver.last_run_info["generation_on_device"] and Q.solver.last_run_info["generation_uploads"] == 6
    helpers.assert_close(a["state"], b["state"], "device program vs host evaluation", rtol=1e-12)


GEN_BODIES = [
    "params['a'] * np.exp(-t / 0.5) * np.where(E < 400.0, 1.0, 0.25) * (0.5 + y * x)",
    "1e-8 * (1 + math.sin(6.0 * t) ** 2) * max(E, 250.0) / 250.0 * abs(x - 0.5)",
    "(2e-8 if E < 300 else 5e-9) * (t < 0.4 or x > 0.7) + 1e-9 * (0.2 < y <= 0.6)",
    "np.clip(1e-8 * np.power(E / 200.0, -1.5) * np.heaviside(0.5 - t, 0.5), 1e-10, 4e-9) + 1e-9 * (int(10 * x) % 3) + 0 * t",
    "1e-8 * np.minimum(np.maximum(x, 0.3), y + 0.1) * np.tanh(t) * np.sqrt(E) / (1.0 + np.log10(E)) + 1e-9 * (7 // 2) * float(t > 0)",
    "np.full_like(x, 3e-9) * (not (t > 1.0)) * np.ones_like(y) + pow(x, 2) * 1e-9 * bool(E > 0) * np.cos(t) ** 2",
]


@pytest.mark.parametrize("body", GEN_BODIES)
def test_generation_program_reproduces_the_host_evaluator(body):
    """qpb_eval_generation_program: the device's value of g(E, x, y, t) for every bin and cell against the package's
    host evaluator (itself pinned bit for bit to the reference's, tests/test_userexpr.py)."""
    import qpsim_b200 as Q
    from qpsim_b200 import capi, userexpr
    mask = cases.annulus_mask(24)
    n = int(mask.sum())
    E, dE = Q.build_energy_grid(180.0, 1.0, 4.0, 9)
    spec = Q.ExternalGenerationSpec(mode="custom", custom_body=body, custom_params={"a": 3e-8})
    gen = userexpr.CustomGeneration(spec, E, mask)
    assert gen.program is not None
    with capi.Context(ny=mask.shape[0], nx=mask.shape[1], ne=E.size, nw=0, ncell=n, flags=0, dx=1.0, dE=dE) as ctx:
        ctx.upload_generation_program(gen.program, gen.E, gen.x, gen.y)
        for t in (0.0, 0.3, 1.7):
            got, want = ctx.eval_generation_program(t), gen(t)
            np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-30)


@pytest.mark.parametrize("body,msg", [("1e-8 * (0.5 - t) * x", "negative values"), ("1e-8 * np.sqrt(0.7 - t) + 0 * x", "non-finite")])
def test_generation_program_values_are_checked_like_the_reference(body, msg):
    """solver.py:954-962: a custom body that turns negative or non-finite during the run raises ValueError with the
    reference's messages - also when the values never reach the host."""
    import qpsim_b200 as Q
    case = dict([c for c in CUSTOM if c["name"] == "custom_gen_timedep"][0])
    case["generation"] = dict(mode="custom", custom_body=body, custom_params={})
    with pytest.raises(ValueError, match=msg):
        helpers.run_dropin(case)


def test_unsafe_custom_generation_is_rejected():
    """tests/test_regressions.py:501-525 of the reference."""
    import qpsim_b200 as Q

    mask = np.ones((1, 2), dtype=bool)
    edges = Q.extract_edge_segments(mask)
    bcs = {e.edge_id: Q.BoundaryCondition(kind="reflective") for e in edges}
    gen = Q.ExternalGenerationSpec(mode="custom", custom_body="__import__('os').system('echo unsafe')")
    with pytest.raises(ValueError):
        Q.run_2d_crank_nicolson(mask=mask, edges=edges, edge_conditions=bcs, initial_field=np.zeros((1, 2)),
                                diffusion_coefficient=6.0, dt=0.1, total_time=0.1, dx=1.0, energy_gap=180.0,
                                energy_min_factor=1.0, energy_max_factor=3.0, num_energy_bins=8,
                                enable_diffusion=False, external_generation=gen)


def test_nonuniform_gap_trap_96x96x16_matches_reference():
    """Non-uniform gap at a size where it matters (solver.py:235-321, 1145-1164, 834-875): 96 x 96 mask with all five
    wall kinds, 16 bins, a gap that steps across the device and dips cell by cell inside a trap (44 distinct values):
    per-cell D(E, x) with harmonic-mean faces in the diffusion, one collision table per gap value."""
    case = cases.golden_cases_large()[0]
    want = helpers.load_golden(case["name"])
    got = helpers.run_dropin(case)
    keep = want["keep"]
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-12)
    helpers.assert_close(got["state"][keep], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    helpers.assert_close(got["phonons"][keep][:, :, want["ph_cells"]], want["phonons"], "n_ph", rtol=helpers.RTOL_PHONON)


@pytest.mark.parametrize("name", ["custom_gen_static", "custom_gen_timedep", "custom_gen_scalar_body"])
def test_custom_generation_in_the_sharded_loop(name, monkeypatch):
    """The sharded loop behind devices=[...] (multigpu._run_spmd; here as a world of one spawned rank,
    QPB_FORCE_SHARDED=1) applies custom generation per rank on its own cells: the device program for time-dependent
    bodies, one evaluated-and-uploaded slice for time-independent ones.  Against the unmodified reference."""
    import qpsim_b200 as Q
    case = [c for c in CUSTOM if c["name"] == name][0]
    gold = helpers.load_golden("custom_modes")
    monkeypatch.setenv("QPB_FORCE_SHARDED", "1")
    got = helpers.run_dropin(case, devices=[0])
    info = dict(Q.solver.last_run_info)
    assert info.get("world") == 1, info
    p = name + "/"
    helpers.assert_close(got["state"], gold[p + "state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], gold[p + "mass"], rtol=helpers.RTOL)
    if name == "custom_gen_timedep":
        assert info["generation_on_device"] and info["generation_uploads"] == 0
    else:
        assert info["generation_uploads"] == 1


def test_sharded_custom_generation_raises_the_reference_errors(monkeypatch):
    import qpsim_b200 as Q
    case = dict([c for c in CUSTOM if c["name"] == "custom_gen_timedep"][0])
    case["generation"] = dict(mode="custom", custom_body="1e-8 * (0.5 - t) * x", custom_params={})
    monkeypatch.setenv("QPB_FORCE_SHARDED", "1")
    with pytest.raises(ValueError, match="negative values"):
        helpers.run_dropin(case, devices=[0])
