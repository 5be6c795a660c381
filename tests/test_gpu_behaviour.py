"""Behaviour the reference's own tests pin at the run_2d_crank_nicolson boundary, restated for the drop-in
(reference files: tests/test_regressions.py, tests/test_physics_safety.py, qpsim/validation.py thresholds).
Everything here goes through the C ABI on a B200:  python -m pytest tests -m gpu
"""
import warnings

import numpy as np
import pytest

import cases
import qpsim_b200 as Q
from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu

GAP = cases.GAP


def _geom(ny, nx, kind="reflective"):
    mask = np.ones((ny, nx), dtype=bool)
    edges = Q.extract_edge_segments(mask)
    return mask, edges, {e.edge_id: Q.BoundaryCondition(kind=kind) for e in edges}


# ---- tests/test_regressions.py:232-301 ----------------------------------------------------------------------
def test_uniform_field_between_reflective_walls_does_not_move():
    mask, edges, bcs = _geom(2, 2)
    _, frames, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.full((2, 2), 3.0), 1.0, 0.2, 1.0, 1.0, store_every=1)
    assert len(frames) == 6
    for f in frames:
        assert np.allclose(f[mask], 3.0, atol=1e-12)
    assert np.allclose(mass, 12.0, atol=1e-12)


def test_last_step_is_shortened_to_hit_total_time():
    mask, edges, bcs = _geom(2, 2)
    times, frames, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((2, 2)), 1.0, 0.3, 1.0, 1.0, store_every=1)
    assert abs(times[-1] - 1.0) < 1e-12 and len(times) == 5
    # remainder step against the oracle (two prepared operators: dt and the remainder)
    mask, edges, bcs = _geom(5, 7, "absorbing")
    field = cases.gaussian_field(mask, base=0.2, amp=1.0, sigma=0.2)
    t, fr, m, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, field, 2.0, 0.3, 1.0, 0.7)
    ref = O.run(mask, edges, bcs, field, 2.0, 0.3, 1.0, 0.7)
    assert np.allclose(t, ref.times, atol=1e-12)
    assert np.max(np.abs(fr[-1][mask] - ref.state_frames[-1][0])) <= 1e-9 * np.max(np.abs(ref.state_frames[-1]))


def test_progress_callback_sees_every_stored_frame_and_may_raise():
    mask, edges, bcs = _geom(2, 2)
    seen = []

    def cb(t, frame):
        seen.append((t, frame.copy()))
        raise RuntimeError("callbacks must not break the run")   # solver.py:1375-1379 swallows this

    times, frames, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((2, 2)), 1.0, 0.1, 0.3, 1.0, progress_callback=cb)
    assert len(seen) == len(times) == 4 and seen[0][0] == 0.0 and abs(seen[-1][0] - times[-1]) < 1e-12
    assert np.allclose(seen[-1][1], frames[-1])


def test_store_every_nonpositive_means_one_and_frames_are_nan_outside_mask():
    mask = cases.annulus_mask(12)
    edges = Q.extract_edge_segments(mask)
    bcs = {e.edge_id: Q.BoundaryCondition(kind="reflective") for e in edges}
    times, frames, mass, limits, ef, E = Q.run_2d_crank_nicolson(
        mask, edges, bcs, np.where(mask, 1.0, 0.0), 6.0, 0.5, 1.5, 1.0, store_every=0, energy_gap=GAP,
        energy_max_factor=3.0, num_energy_bins=4)
    assert len(times) == 4 and len(ef) == 4 and len(ef[0]) == 4 and E.shape == (4,)
    assert np.all(np.isnan(frames[-1][~mask])) and np.all(np.isfinite(frames[-1][mask]))
    assert np.all(np.isnan(ef[-1][2][~mask]))
    assert limits[0] <= limits[1]
    assert abs(mass[-1] - mass[0]) <= 1e-10 * mass[0]


# ---- argument checks: solver.py:1057-1077, 1129 ---------------------------------------------------------------
def test_argument_errors_match_the_reference():
    mask, edges, bcs = _geom(2, 3)
    f = np.ones((2, 3))
    with pytest.raises(ValueError, match="dt and total_time must be positive"):
        Q.run_2d_crank_nicolson(mask, edges, bcs, f, 1.0, 0.0, 1.0, 1.0)
    with pytest.raises(ValueError, match="Diffusion coefficient must be positive"):
        Q.run_2d_crank_nicolson(mask, edges, bcs, f, 0.0, 0.1, 1.0, 1.0)
    with pytest.raises(ValueError, match="Initial field shape must match mask shape"):
        Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((3, 2)), 1.0, 0.1, 1.0, 1.0)
    with pytest.raises(ValueError, match="no interior points"):
        Q.run_2d_crank_nicolson(np.zeros((2, 3), bool), [], {}, f, 1.0, 0.1, 1.0, 1.0)
    with pytest.raises(ValueError):
        Q.run_2d_crank_nicolson(mask, edges, bcs, f, 1.0, 0.1, 1.0, 1.0, energy_gap=GAP, collision_solver="euler")
    with pytest.raises(ValueError, match="All edges must be assigned boundary conditions") as ei:
        Q.run_2d_crank_nicolson(mask, edges, {}, f, 1.0, 0.1, 1.0, 1.0)
    assert type(ei.value).__name__ == "BoundaryAssignmentError"
    # diffusion disabled: no boundary conditions needed (tests/test_regressions.py:782-801)
    t, fr, *_ = Q.run_2d_crank_nicolson(mask, [], {}, f, 1.0, 0.1, 0.2, 1.0, energy_gap=GAP, num_energy_bins=3,
                                        energy_max_factor=2.0, enable_diffusion=False)
    assert len(t) == 3


# ---- tests/test_physics_safety.py:57-106 ------------------------------------------------------------------------
def _pauli_kwargs(enforce):
    mask, edges, bcs = _geom(1, 1)
    return dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=np.array([[2.0]]), diffusion_coefficient=6.0,
                dt=0.1, total_time=0.2, dx=1.0, energy_gap=180.0, energy_min_factor=1.5, energy_max_factor=1.5,
                num_energy_bins=1, enable_diffusion=False, enable_recombination=False, enable_scattering=False,
                enforce_pauli=enforce, pauli_error_threshold=1.0)


def test_pauli_violation_raises_with_the_reference_message():
    with pytest.raises(ValueError, match="Pauli occupation exceeded limit"):
        Q.run_2d_crank_nicolson(**_pauli_kwargs(True))


def test_pauli_violation_warns_once_when_not_enforced():
    with pytest.warns(UserWarning, match="Pauli occupation exceeded limit") as rec:
        Q.run_2d_crank_nicolson(**_pauli_kwargs(False))
    assert len([w for w in rec if "Pauli" in str(w.message)]) == 1


# ---- qpsim/validation.py thresholds (:116, :172, :208, :251) -----------------------------------------------------
def test_thermal_state_with_frozen_thermal_phonons_is_stationary():
    nx, ne = 16, 24
    mask, edges, bcs = _geom(1, nx)
    E, dE = Q.build_energy_grid(GAP, 1.0, 4.0, ne)
    n_eq = Q.thermal_qp_weights(E, GAP, 0.1, 0.18)
    amp = float(np.sum(n_eq) * dE)
    *_, ef, _ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.full((1, nx), amp), 6.0, 0.1, 0.5, 1.0, store_every=1,
                                         energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=ne,
                                         energy_weights=n_eq, enable_recombination=True, enable_scattering=True,
                                         dynes_gamma=0.18, tau_s=440.0, tau_r=440.0, T_c=1.2, bath_temperature=0.1,
                                         freeze_phonon_dynamics=True)
    s0 = np.array([f[0] for f in ef[0]])
    s1 = np.array([f[0] for f in ef[-1]])
    assert np.max(np.abs(s1 - s0)) / np.max(np.abs(s0)) <= 1e-6


def test_pure_diffusion_conserves_number_on_a_reflective_line():
    nx = 64
    mask, edges, bcs = _geom(1, nx)
    x = (np.arange(nx) + 0.5) / nx
    _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, (1.0 + 0.4 * np.cos(2 * np.pi * x))[None, :], 6.0, 0.1, 1.0, 1.0)
    assert abs(mass[-1] - mass[0]) / abs(mass[0]) <= 1e-10


def test_pure_scattering_conserves_number_and_recombination_only_removes():
    mask, edges, bcs = _geom(1, 4)
    E, _ = Q.build_energy_grid(GAP, 1.0, 4.0, 24)
    w = np.exp(-((E - 2.6 * GAP) / (0.6 * GAP)) ** 2)
    kw = dict(energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=24, energy_weights=w,
              enable_diffusion=False, dynes_gamma=0.18, tau_s=440.0, tau_r=440.0, T_c=1.2, bath_temperature=0.1,
              freeze_phonon_dynamics=True, store_every=1)
    _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.full((1, 4), 2e-4), 6.0, 0.1, 0.5, 1.0,
                                             enable_scattering=True, enable_recombination=False, **kw)
    assert abs(mass[-1] - mass[0]) / abs(mass[0]) <= 5e-5   # the reference itself drifts by 2.3e-5 on this input
    mask, edges, bcs = _geom(1, 1)
    _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.array([[1e-3]]), 6.0, 0.1, 1.0, 1.0, energy_gap=GAP,
                                             energy_min_factor=1.5, energy_max_factor=1.5, num_energy_bins=1,
                                             enable_diffusion=False, enable_recombination=True, dynes_gamma=0.0,
                                             tau_r=440.0, T_c=1.2, bath_temperature=0.0, freeze_phonon_dynamics=True)
    assert all(mass[i + 1] <= mass[i] + 1e-15 for i in range(len(mass) - 1)) and mass[-1] < mass[0]


# ---- external generation: tests/test_regressions.py:435-499 -------------------------------------------------------
def test_generation_modes():
    mask, edges, bcs = _geom(3, 3)
    kw = dict(energy_gap=180.0, energy_max_factor=5.0, store_every=1)
    _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.full((3, 3), 0.1), 6.0, 1.0, 5.0, 1.0, num_energy_bins=8,
                                             external_generation=Q.ExternalGenerationSpec(mode="constant", rate=0.01), **kw)
    E, dE = Q.build_energy_grid(180.0, 1.0, 5.0, 8)
    # reflective walls, no collisions: every step adds rate*dt in every bin and cell (solver.py:1464)
    assert np.allclose(np.diff(mass), 0.01 * 1.0 * 8 * dE * 9, rtol=1e-10)
    mask, edges, bcs = _geom(2, 2)
    pulse = Q.ExternalGenerationSpec(mode="pulse", pulse_rate=1.0, pulse_start=0.0, pulse_duration=2.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.zeros((2, 2)), 6.0, 1.0, 4.0, 1.0, num_energy_bins=5,
                                                 enable_diffusion=False, external_generation=pulse, enforce_pauli=False, **kw)
    assert mass[2] > mass[0] and abs(mass[3] - mass[2]) < 1e-10 and abs(mass[4] - mass[2]) < 1e-10
    a = Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((2, 2)), 6.0, 1.0, 3.0, 1.0, num_energy_bins=5,
                                external_generation=Q.ExternalGenerationSpec(mode="none"), **kw)[2]
    b = Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((2, 2)), 6.0, 1.0, 3.0, 1.0, num_energy_bins=5, **kw)[2]
    assert np.allclose(a, b, rtol=0, atol=1e-12)


def test_pair_breaking_creates_quasiparticles_from_an_empty_state():
    """tests/test_regressions.py:591-620: hot phonons + empty quasiparticle state -> density grows."""
    mask, edges, bcs = _geom(1, 3)
    _, _, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.zeros((1, 3)), 6.0, 0.5, 2.0, 1.0, energy_gap=GAP,
                                             energy_max_factor=3.0, num_energy_bins=10, enable_diffusion=False,
                                             enable_recombination=True, tau_r=100.0, T_c=1.2, bath_temperature=1.0,
                                             dynes_gamma=0.18)
    assert mass[0] == 0.0 and mass[-1] > 0.0


def test_ensemble_runs_members_through_the_single_gpu_path():
    """BASELINE configs[4] in miniature: a parameter sweep over bath temperature and tau_0 on one mask; every member
    equals a direct call of run_2d_crank_nicolson with the same arguments."""
    import cases
    import helpers

    case = cases.meander_c2(ny=24, nx=32, ne=8, steps=2)
    edges = Q.extract_edge_segments(case["mask"])
    bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
    gen = Q.ExternalGenerationSpec(**case["generation"])
    base = cases.solver_kwargs(case, edges, bcs, gen, Q.physics)
    members = Q.parameter_grid(base, bath_temperature=[0.05, 0.3], tau_0=[100.0, 800.0])
    res = Q.run_ensemble(members, reduce=lambda out: (out[0], out[2], np.array(out[4][-1])))
    assert len(res) == 4
    for kw, (times, mass, last) in zip(members, res):
        t2, _, m2, _, ef, _ = Q.run_2d_crank_nicolson(**kw)
        assert times == t2 and mass == m2
        np.testing.assert_array_equal(last, np.array(ef[-1]))
    assert len({tuple(r[1]) for r in res}) == 4   # the parameters matter


# ---- robustness of the sweep iteration where the reference's direct solve has no such concept ------------------
@pytest.mark.parametrize("alpha", [40.0, 2500.0])
def test_stiff_steps_on_a_slotted_mask_converge_like_the_direct_solve(alpha):
    """dt D / dx^2 in the tens to thousands (fine mesh, long step) on a non-commuting geometry: the residual test is
    floored at what fp64 resolves and the cyclic shift set grows with log(hi/lo), so the iteration ends where the
    reference's SuperLU solve simply returns."""
    case = cases.meander_c2(ny=48, nx=48, ne=4, steps=2)
    case.update(enable_recombination=False, enable_scattering=False, generation=None, dynes_gamma=0.0,
                dt=2.0 * alpha / cases.D0, total_time=4.0 * alpha / cases.D0, store_every=1)
    import helpers

    got = helpers.run_dropin(case)
    want = helpers.run_oracle(case)
    assert helpers.norm_err(got["state"], want["state"]) <= 1e-9
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-9)


def test_negative_robin_coefficient_is_solved_or_refused_loudly():
    mask, edges, _ = _geom(6, 9)
    for beta, ok in ((-0.05, True), (-5.0, False)):
        bcs = {e.edge_id: Q.BoundaryCondition(kind="robin", value=beta, aux_value=0.1) for e in edges}
        field = cases.gaussian_field(mask, sigma=0.3)
        args = (mask, edges, bcs, field, 1.0, 0.4, 0.8, 1.0)
        if not ok:
            with pytest.raises(Q.capi.QpbError, match="negative Robin"):
                Q.run_2d_crank_nicolson(*args)
            continue
        _, frames, mass, *_ = Q.run_2d_crank_nicolson(*args)
        res = O.run(mask, edges, bcs, field, 1.0, 0.4, 0.8, 1.0)
        np.testing.assert_allclose(frames[-1][mask], res.state_frames[-1][0], rtol=1e-9)


@pytest.mark.parametrize("name", ["c2_meander_40x40x12", "nonuniform_10x14x8", "annulus_32x5_mixed"])
def test_krylov_solve_reproduces_the_reference_fixtures(name, monkeypatch):
    """The preconditioned BiCGStab path (stiff steps) forced on at ordinary step lengths: same fixtures, same bar -
    uniform D on a slotted mask, per-cell D (non-uniform gap), all five boundary kinds with a remainder step."""
    import helpers

    monkeypatch.setenv("QPB_KRYLOV_RATIO", "0")
    case = next(c for c in cases.golden_cases() if c["name"] == name)
    want = helpers.load_golden(name)
    got = helpers.run_dropin(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-9)
