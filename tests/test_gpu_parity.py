"""CUDA path (through the C ABI / the Python drop-in) against the reference's golden fixtures and the oracle.

Bar (BASELINE.json north_star): max relative error 1e-9 on n(x,y,E), total quasiparticle number to the same
tolerance.  All tests here need a B200:  python -m pytest tests -m gpu
"""
import numpy as np
import pytest

import cases
import helpers
import qpsim_b200 as Q
from oracle import qp_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = cases.golden_cases()


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_dropin_matches_reference_fixture(case):
    want = helpers.load_golden(case["name"])
    got = helpers.run_dropin(case)
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-12)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    np.testing.assert_allclose(got["limits"], want["limits"], rtol=1e-8)
    if "phonons" in want:
        helpers.assert_close(got["phonons"], want["phonons"], "n_ph(omega,cell)", rtol=helpers.RTOL_PHONON)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True)])
def test_collision_helper_matches_reference_pixels(tag, flags):
    """apply_collision_step_fischer_catelani_uniform on the fixture's random states (solver.py:794-831)."""
    z = helpers.load_golden("tables_and_pixels")
    rec, sc = flags
    n, ph = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
    dE = float(z[f"{tag}_dE"])
    args = (z[f"{tag}_Kr"], z[f"{tag}_Ks"], z[f"{tag}_rho"], z[f"{tag}_idx_diff"], z[f"{tag}_idx_sum"], z[f"{tag}_sign"])
    Q.apply_collision_step_fischer_catelani_uniform(n, ph, *args, dE, 0.3, enable_recombination=rec,
                                                    enable_scattering=sc)
    if rec and sc:
        want_n, want_ph = z[f"{tag}_n_out"], z[f"{tag}_ph_out"]
    else:
        want_n, want_ph = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
        O.collide(want_n, want_ph, *args, dE, 0.3, recomb=rec, scat=sc)
    helpers.assert_close(n.T, want_n.T, "n", rtol=1e-10)
    helpers.assert_close(ph.T, want_ph.T, "n_ph", rtol=helpers.RTOL_PHONON)


@pytest.mark.parametrize("generic", [False, True])
def test_collision_accuracy_against_extended_precision(generic, monkeypatch):
    """One collision call, element by element, both kernels.  Truth = the reference's formulas in 80-bit
    arithmetic.  The float64 reference is off from it by e_ref (its (exp(x)-1)/b and the emission-absorption
    cancellation lose digits); the CUDA result must be as accurate: |cuda - truth| <= 4 max(e_ref) per energy /
    phonon bin family, and the quasiparticle update must agree with the reference to 1e-10 outright."""
    monkeypatch.setenv("QPB_FORCE_GENERIC", "1" if generic else "0")
    z = helpers.load_golden("tables_and_pixels")
    for tag in "abc":
        dE = float(z[f"{tag}_dE"])
        args = (z[f"{tag}_Kr"], z[f"{tag}_Ks"], z[f"{tag}_rho"], z[f"{tag}_idx_diff"], z[f"{tag}_idx_sum"], z[f"{tag}_sign"])
        n, ph = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
        Q.apply_collision_step_fischer_catelani_uniform(n, ph, *args, dE, 0.3, enable_recombination=True,
                                                        enable_scattering=True)
        helpers.assert_close(n.T, z[f"{tag}_n_out"].T, "n", rtol=1e-10)
        for c in range(n.shape[1]):
            tn, tp = helpers.collide_pixel_extended(z[f"{tag}_n_in"][:, c], z[f"{tag}_ph_in"][:, c], *args, dE, 0.3)
            tn, tp = tn.astype(float), tp.astype(float)
            for got, ref, truth in ((n[:, c], z[f"{tag}_n_out"][:, c], tn), (ph[:, c], z[f"{tag}_ph_out"][:, c], tp)):
                scale = np.maximum(np.abs(truth), 1e-300)
                e_ref = np.abs(ref - truth) / scale
                e_gpu = np.abs(got - truth) / scale
                assert np.max(e_gpu) <= 4.0 * np.max(e_ref) + 1e-13, (tag, c, np.max(e_gpu), np.max(e_ref))


@pytest.mark.parametrize("path,ne,n,ngaps", [("generic", 12, 23, 2), ("grouped", 12, 23, 2), ("grouped", 40, 300, 7),
                                             ("grouped", 136, 150, 3), ("generic", 40, 60, 60)])
def test_collision_kernels_nonuniform_tables(path, ne, n, ngaps, monkeypatch):
    """Per-pixel tables (solver.py:834-875).  "generic": per-cell table lookups with shared-memory atomics (forced, or
    chosen because every cell has its own gap).  "grouped": the structured kernel with the cells regrouped so that a
    CTA's cells share one gap table (ragged groups, padding lanes, 32- and 16-cell CTAs)."""
    if path == "generic" and ngaps < n:
        monkeypatch.setenv("QPB_FORCE_GENERIC", "1")
    rng = np.random.default_rng(5)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, 4.0, ne)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    gaps = cases.GAP * (1.0 - 0.07 * rng.integers(0, ngaps, n) / max(1, ngaps - 1)) if ngaps < n else \
        cases.GAP * (1.0 - 0.07 * np.arange(n) / n)
    rho_all = np.stack([Q.density_of_states(E, g, 0.18) for g in gaps])
    Kr_all = np.stack([Q.recombination_kernel_base(E, g, 440.0, 1.2) for g in gaps])
    Ks_all = np.stack([Q.scattering_kernel_base(E, g, 440.0, 1.2) for g in gaps])
    state = rho_all.T * rng.uniform(0, 0.5, (ne, n))
    ph = Q.thermal_phonon_occupation(om, 0.3)[:, None] * rng.uniform(0.5, 2, (om.size, n))
    s_ref, p_ref = state.copy(), ph.copy()
    O.collide(s_ref, p_ref, Kr_all, Ks_all, rho_all, idd, ids, sg, dE, 0.35, recomb=True, scat=True)
    Q.apply_collision_step_fischer_catelani_nonuniform(state, ph, Kr_all, Ks_all, rho_all, idd, ids, sg, dE, 0.35,
                                                       enable_recombination=True, enable_scattering=True)
    helpers.assert_close(state.T, s_ref.T, "n", rtol=1e-10)
    helpers.assert_close(ph.T, p_ref.T, "n_ph", rtol=helpers.RTOL_PHONON)


@pytest.mark.parametrize("ne,n", [(24, 700), (64, 1500), (56, 333)])
def test_collision_cta_shapes_on_small_energy_grids_agree(ne, n, monkeypatch):
    """Energy grids of at most 64 bins run 256-thread CTAs, two per SM; QPB_COLL_NT=512 keeps the wide CTA the larger
    grids use.  Same tiles, another split of the diagonals between the warps: both against the oracle (solver.py:703-791)
    and against each other (the pieces of a diagonal meet through shared-memory atomics, so the last bits may differ)."""
    rng = np.random.default_rng(11)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, 3.0, ne)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    rho = Q.density_of_states(E, cases.GAP, 0.18)
    Kr = Q.recombination_kernel_base(E, cases.GAP, 440.0, 1.2)
    Ks = Q.scattering_kernel_base(E, cases.GAP, 440.0, 1.2)
    state0 = rho[:, None] * rng.uniform(0, 0.4, (ne, n))
    ph0 = Q.thermal_phonon_occupation(om, 0.3)[:, None] * rng.uniform(0.5, 2, (om.size, n))
    s_ref, p_ref = state0.copy(), ph0.copy()
    O.collide(s_ref, p_ref, Kr, Ks, rho, idd, ids, sg, dE, 0.4, recomb=True, scat=True)
    got = {}
    for nt in ("", "512"):
        if nt:
            monkeypatch.setenv("QPB_COLL_NT", nt)
        else:
            monkeypatch.delenv("QPB_COLL_NT", raising=False)
        s, p = state0.copy(), ph0.copy()
        Q.apply_collision_step_fischer_catelani_uniform(s, p, Kr, Ks, rho, idd, ids, sg, dE, 0.4,
                                                        enable_recombination=True, enable_scattering=True)
        helpers.assert_close(s.T, s_ref.T, f"n (QPB_COLL_NT={nt or 'default'})", rtol=1e-10)
        helpers.assert_close(p.T, p_ref.T, f"n_ph (QPB_COLL_NT={nt or 'default'})", rtol=helpers.RTOL_PHONON)
        got[nt] = (s, p)
    helpers.assert_close(got[""][0].T, got["512"][0].T, "n, narrow against wide CTA", rtol=1e-13)
    helpers.assert_close(got[""][1].T, got["512"][1].T, "n_ph, narrow against wide CTA", rtol=1e-9)


def test_reflective_uniform_field_is_stationary():
    """tests/test_regressions.py:232-252 of the reference: uniform field, reflective walls, mass = 12."""
    mask = np.ones((3, 4), dtype=bool)
    edges = Q.extract_edge_segments(mask)
    bcs = {e.edge_id: Q.BoundaryCondition(kind="reflective") for e in edges}
    times, frames, mass, *_ = Q.run_2d_crank_nicolson(mask, edges, bcs, np.ones((3, 4)), 1.0, 0.1, 1.0, 1.0)
    assert np.allclose(frames[-1], 1.0, atol=1e-12)
    assert all(abs(m - 12.0) < 1e-10 for m in mass)


def test_mass_conserved_and_medium_meander():
    """All-reflective masked diffusion conserves total number to the parity tolerance; larger than the fixtures."""
    case = cases.meander_c2(ny=96, nx=96, ne=16, steps=4, bc="reflective", pad=8, pitch=16, gap_len=24)
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    got = helpers.run_dropin(case)
    want = helpers.run_oracle(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    assert abs(got["mass"][-1] - got["mass"][0]) <= 1e-9 * got["mass"][0]
    assert got["frames_nan_outside"]


@pytest.mark.parametrize("fused", [False, True], ids=["all_to_all", "fused_exchange"])
def test_sharded_driver_on_one_gpu_matches_single_context(fused):
    """multigpu.DeviceStages + ShardedStepper with world = 1 (two contexts, block scatter/gather, row permutation,
    shared stream) against qpb_advance on one context and against the oracle.  fused: the collision kernel stores
    into / loads from the diffusion context's state itself (qpb_set_exchange; here the only "peer" is this rank)."""
    from qpsim_b200 import capi
    from qpsim_b200.multigpu import DeviceStages, ShardedProblem, ShardedStepper, ShardPlan
    import torch

    case = cases.meander_c2(ny=40, nx=48, ne=12, steps=3)
    mask = case["mask"]
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, case["dx"])
    E, dE = Q.build_energy_grid(case["energy_gap"], 1.0, case["energy_max_factor"], case["num_energy_bins"])
    rho = Q.density_of_states(E, case["energy_gap"], case["dynes_gamma"])
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    n = int(mask.sum())
    state = (rho / (rho.sum() * dE))[:, None] * case["initial_field"][mask][None, :]
    phon = Q.thermal_phonon_occupation(om, case["bath_temperature"])[:, None] * np.ones((1, n))
    D = case["diffusion_coefficient"] * np.sqrt(np.maximum(0.0, 1.0 - (case["energy_gap"] / E) ** 2))
    Kr = Q.recombination_kernel_base(E, case["energy_gap"], case["tau_0"], case["T_c"])
    Ks = Q.scattering_kernel_base(E, case["energy_gap"], case["tau_0"], case["T_c"])
    prob = ShardedProblem(mask=mask, bcx=bcx, bcy=bcy, src=src, dx=case["dx"], dE=dE, D=D, variable_D=False,
                          rho=rho[None], Kr=Kr[None], Ks=Ks[None], gap_id=None, idx_diff=idd, idx_sum=ids, sign=sg,
                          nw=om.size, state=state, phonons=phon)
    plan = ShardPlan(E.size, n, 1, 0, interleave=True)
    stages = DeviceStages(plan, prob, 0, case["dt"])
    if fused:
        assert stages.enable_fused_exchange(prob)
    g = case["generation"]
    with torch.cuda.stream(stages.stream):
        st = ShardedStepper(plan, stages, diffusion=True, collisions=True)
        assert st.fused == fused
        t, recs = 0.0, []
        for _ in range(3):
            rate = g["pulse_rate"] if g["pulse_start"] <= t < g["pulse_start"] + g["pulse_duration"] else None
            recs.append(st.step(case["dt"], 0, rate, want_pauli=True))
            t += case["dt"]
        got = st.gather_state()
        merged = st.merge_pauli(recs)
    stages.close()
    want = helpers.run_oracle(case)
    helpers.assert_close(got, want["state"][-1], "sharded n(E,cell)")
    single = helpers.run_dropin(case)
    helpers.assert_close(got, single["state"][-1], "sharded vs single context", rtol=1e-12)
    assert all(m[2] == -1 and 0.0 < m[0] < 1.0 for m in merged)


@pytest.mark.parametrize("shape", [(48, 64), (70, 96), (260, 48), (33, 512), (512, 34)])
@pytest.mark.parametrize("bc", ["reflective", "mixed"])
def test_pipelined_sweeps_match_oracle(shape, bc):
    """Shapes that route the x and/or y sweeps through the persistent TMA-pipelined kernels (qpb_sweep_pipe.cu):
    rows that are multiples of 16 cells, ragged last tiles, columns longer than one TMA box, narrow strips."""
    ny, nx = shape
    case = cases.meander_c2(ny=ny, nx=nx, ne=5, steps=2, bc=bc, pad=3, pitch=9, gap_len=max(6, nx // 5))
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # the mixed boundary sources drive the occupation up; not the point here
        got = helpers.run_dropin(case, enforce_pauli=False)
    want = helpers.run_oracle(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    info = Q.solver.last_run_info
    assert info["direct_mode"] == 0 and info["pr_iterations"] > 0


@pytest.mark.parametrize("shape", [(40, 1040), (1050, 48), (530, 544)])
@pytest.mark.parametrize("bc", ["reflective", "mixed"])
def test_segmented_sweeps_match_oracle(shape, bc):
    """Lines longer than 512 cells: the pipelined kernels solve them in overlapping segments (halo = carry reach of
    the factor tables).  Rows only, columns only, and both directions segmented; masked geometry with slots."""
    ny, nx = shape
    case = cases.meander_c2(ny=ny, nx=nx, ne=3, steps=2, bc=bc, pad=3, pitch=max(9, ny // 12),
                            gap_len=max(6, nx // 5))
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = helpers.run_dropin(case, enforce_pauli=False)
    info = dict(Q.solver.last_run_info)
    want = helpers.run_oracle(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)
    assert info["sweep_path"] == 3, info


def test_pipelined_and_legacy_sweeps_agree(monkeypatch):
    """Same run with QPB_NO_PIPE=1 (chunked table kernels of qpb_sweep_fast.cu): both solve the same linear
    system to the same residual tolerance."""
    case = cases.meander_c2(ny=64, nx=80, ne=6, steps=3)
    case["enable_recombination"] = case["enable_scattering"] = False
    a = helpers.run_dropin(case)
    monkeypatch.setenv("QPB_NO_PIPE", "1")
    b = helpers.run_dropin(case)
    helpers.assert_close(a["state"], b["state"], "pipe vs legacy", rtol=1e-10)


@pytest.mark.parametrize("shape", [(256, 256, "mixed"), (96, 160, "mixed"), (200, 96, "short_absorbing"), (48, 256, "mixed")],
                         ids=["256x256", "96x160", "200x96_ragged_rows", "48x256"])
def test_bin_resident_cluster_solve_matches_the_launched_sweeps_and_the_oracle(shape, monkeypatch):
    """Masks of up to 256 x 256 cells with rows that are a multiple of 16 long are solved bin-resident
    (qpb_resident.cu: a thread-block cluster keeps u, b and the correction of one bin in shared memory for the whole
    Peaceman-Rachford iteration, sweep_path 5).  Same linear system, same stop test: the result must agree with the
    launched sweeps (QPB_NO_RESIDENT=1) far inside the tolerance and with the oracle's SuperLU solve at 1e-9.  Shapes:
    8 CTAs x 32 rows, 6 x 16, 7 x 32 with a last CTA that is half empty, 3 x 16; all five wall kinds."""
    ny, nx, bc = shape
    case = cases.meander_c2(ny=ny, nx=nx, ne=5, steps=3)
    case["bc"] = bc
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    got = helpers.run_dropin(case, enforce_pauli=False)
    assert Q.solver.last_run_info["sweep_path"] == 5, Q.solver.last_run_info
    monkeypatch.setenv("QPB_NO_RESIDENT", "1")
    ref = helpers.run_dropin(case, enforce_pauli=False)
    assert Q.solver.last_run_info["sweep_path"] in (1, 2), Q.solver.last_run_info
    helpers.assert_close(got["state"], ref["state"], "resident vs launched sweeps", rtol=1e-11)
    want = helpers.run_oracle(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=helpers.RTOL)


def test_bin_resident_cluster_solve_is_reproducible_run_to_run():
    """The CTAs of a cluster hand carries, halo rows and verdicts to each other through shared memory (st.async +
    mbarrier) with single-buffered slots: a missing ordering would show as a result that changes from run to run.  Eight
    runs of the same solve (many bins per cluster in flight, different arrival orders) must agree bit for bit."""
    case = cases.meander_c2(ny=256, nx=256, ne=40, steps=2)
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    first = None
    for _ in range(8):
        got = helpers.run_dropin(case, enforce_pauli=False)["state"]
        assert Q.solver.last_run_info["sweep_path"] == 5
        if first is None:
            first = got
        else:
            assert np.array_equal(first, got)


@pytest.mark.parametrize("resident", ["1", "0"], ids=["launched_sweeps", "bin_resident"])
def test_a_sweep_iteration_that_hits_its_cap_falls_back_to_the_krylov_solve(resident, monkeypatch):
    """The reference's SuperLU solve always returns; the sweep iteration has a cap (512 iterations, QPB_MAXIT here to
    reach it on purpose).  A solve that hits the cap restarts from b with the preconditioned BiCGStab solve - from the
    launched sweeps and from the bin-resident solve (which has written b out for exactly this case) - and the run still
    matches the oracle."""
    case = cases.meander_c2(ny=64, nx=128, ne=4, steps=2)
    case["enable_recombination"] = case["enable_scattering"] = False
    case["generation"] = None
    monkeypatch.setenv("QPB_MAXIT", "3")
    monkeypatch.setenv("QPB_NO_RESIDENT", resident)
    got = helpers.run_dropin(case, enforce_pauli=False)
    want = helpers.run_oracle(case)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)")
    monkeypatch.setenv("QPB_NO_KRYLOV", "1")
    with pytest.raises(Exception, match="did not reach tolerance"):
        helpers.run_dropin(case, enforce_pauli=False)


@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True)])
def test_frozen_uniform_phonons_use_packed_kernels_and_match(flags, monkeypatch):
    """freeze_phonon_dynamics with the same occupations in every cell runs the fused 4-product kernel
    (qpb_collide_uniform.cuh); it must agree with the general structured kernel and with the oracle."""
    rec, sc = flags
    rng = np.random.default_rng(17)
    ne, n = 40, 75
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, 4.0, ne)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    rho = Q.density_of_states(E, cases.GAP, 0.18)
    Kr = Q.recombination_kernel_base(E, cases.GAP, 300.0, 1.2)
    Ks = Q.scattering_kernel_base(E, cases.GAP, 500.0, 1.2)
    state0 = rho[:, None] * rng.uniform(0, 0.6, (ne, n))
    ph0 = Q.thermal_phonon_occupation(om, 0.35)[:, None] * np.ones((1, n))
    outs = []
    for no_uniform in ("0", "1"):
        monkeypatch.setenv("QPB_NO_UNIFORM", no_uniform)
        s, p = state0.copy(), ph0.copy()
        Q.apply_collision_step_fischer_catelani_uniform(s, p, Kr, Ks, rho, idd, ids, sg, dE, 0.4, enable_recombination=rec,
                                                        enable_scattering=sc, update_phonons=False)
        assert np.array_equal(p, ph0)
        outs.append(s)
    s_ref, p_ref = state0.copy(), ph0.copy()
    O.collide(s_ref, p_ref, Kr, Ks, rho, idd, ids, sg, dE, 0.4, recomb=rec, scat=sc, update_phonons=False)
    helpers.assert_close(outs[0].T, s_ref.T, "packed kernels vs oracle", rtol=1e-11)
    helpers.assert_close(outs[0].T, outs[1].T, "packed vs structured kernel", rtol=1e-11)


@pytest.mark.parametrize("flags", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("shape", [(72, 300), (128, 257)])
def test_frozen_uniform_phonons_tensor_core_gemm(flags, shape, monkeypatch):
    """From 64 energy bins on, the frozen cell-independent products run as one FP64 tensor-core GEMM
    (qpb_collide_gemm.cuh: mma.m8n8k4.f64, 64 x 128 CTA tiles).  Bin and cell counts that are not multiples of the
    tile sizes; compared with the oracle, the fused GEMV kernel and the general structured kernel."""
    rec, sc = flags
    ne, n = shape
    rng = np.random.default_rng(23)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, 6.0, ne)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    rho = Q.density_of_states(E, cases.GAP, 0.18)
    Kr = Q.recombination_kernel_base(E, cases.GAP, 300.0, 1.2)
    Ks = Q.scattering_kernel_base(E, cases.GAP, 500.0, 1.2)
    state0 = rho[:, None] * rng.uniform(0, 0.6, (ne, n))
    ph0 = Q.thermal_phonon_occupation(om, 0.35)[:, None] * np.ones((1, n))
    outs = {}
    for tag, env in (("gemm", {}), ("gemv", {"QPB_NO_GEMM": "1"}), ("struct", {"QPB_NO_UNIFORM": "1"})):
        for k in ("QPB_NO_GEMM", "QPB_NO_UNIFORM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s, p = state0.copy(), ph0.copy()
        Q.apply_collision_step_fischer_catelani_uniform(s, p, Kr, Ks, rho, idd, ids, sg, dE, 0.4, enable_recombination=rec,
                                                        enable_scattering=sc, update_phonons=False)
        assert np.array_equal(p, ph0)
        outs[tag] = s
    s_ref, p_ref = state0.copy(), ph0.copy()
    O.collide(s_ref, p_ref, Kr, Ks, rho, idd, ids, sg, dE, 0.4, recomb=rec, scat=sc, update_phonons=False)
    helpers.assert_close(outs["gemm"].T, s_ref.T, "tensor-core GEMM vs oracle", rtol=1e-11)
    helpers.assert_close(outs["gemm"].T, outs["gemv"].T, "GEMM vs fused GEMV", rtol=1e-11)
    helpers.assert_close(outs["gemm"].T, outs["struct"].T, "GEMM vs structured kernel", rtol=1e-11)
    assert not np.array_equal(outs["gemm"], state0)


@pytest.mark.parametrize("tag", ["small", "gemm", "wide"])
@pytest.mark.parametrize("no_gemm", [False, True], ids=["auto", "gemv"])
def test_euler_fixed_bath_forms_match_reference_fixture(tag, no_gemm, monkeypatch):
    """apply_scattering_step / apply_recombination_step (solver.py:551-605) against the reference's own outputs
    (tests/golden/euler_steps.npz): 24 bins (fused GEMV), 72 and 130 bins (tensor-core GEMM; GEMV when forced)."""
    if no_gemm:
        monkeypatch.setenv("QPB_NO_GEMM", "1")
    g = helpers.load_golden("euler_steps")
    dE, dt = float(g[f"{tag}_dE"]), float(g[f"{tag}_dt"])
    Ks, Kr, rho, gth = g[f"{tag}_Ks"], g[f"{tag}_Kr"], g[f"{tag}_rho"], g[f"{tag}_G_therm"]
    s = g[f"{tag}_state"].copy()
    Q.apply_scattering_step(s, Ks, rho, dE, dt)
    helpers.assert_close(s, g[f"{tag}_after_scattering"], "scattering step", rtol=1e-12)
    s = g[f"{tag}_state"].copy()
    Q.apply_recombination_step(s, Kr, gth, dE, dt)
    helpers.assert_close(s, g[f"{tag}_after_recombination"], "recombination step", rtol=1e-12)
    s = g[f"{tag}_state"].copy()
    for _ in range(3):
        Q.apply_scattering_step(s, Ks, rho, dE, dt)
        Q.apply_recombination_step(s, Kr, gth, dE, dt)
    helpers.assert_close(s, g[f"{tag}_after_3_pairs"], "three scattering + recombination pairs", rtol=1e-11)


def test_separable_state_upload_is_bit_exact():
    """qpb_set_state_separable forms state[i] = spatial * weights[i] (solver.py:1281-1283) and the bath phonon state on
    the device: same bits as uploading the host products."""
    from qpsim_b200 import capi

    rng = np.random.default_rng(5)
    mask = cases.meander_mask(40, 48, pad=4, pitch=12, gap_len=12)
    n, ne, nw = int(mask.sum()), 7, 19
    weights, spatial, bins = rng.random(ne), rng.random(n) * 1e-3, rng.random(nw)
    out = []
    for separable in (False, True):
        with capi.Context(ny=40, nx=48, ne=ne, nw=nw, ncell=n, flags=capi.F_SCATTERING, dx=1.0, dE=2.0) as ctx:
            ctx.upload_geometry(mask)
            if separable:
                ctx.set_state_separable(weights, spatial, bins)
            else:
                ctx.set_state(spatial[None, :] * weights[:, None], bins[:, None] * np.ones((1, n)))
            out.append(ctx.get_state())
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[1][0], spatial[None, :] * weights[:, None])


@pytest.mark.parametrize("staged", ["0", "1"])
def test_large_downloads_staged_and_plain_agree(staged, monkeypatch):
    """qpb_get_frames / qpb_get_state above 16 MiB go through the pinned two-chunk pipeline (uneven last chunk)."""
    from qpsim_b200 import capi

    monkeypatch.setenv("QPB_STAGED_D2H", staged)
    rng = np.random.default_rng(6)
    ny, nx, ne = 250, 272, 33          # 17.1 MiB of frames: two full chunks and a short one
    mask = np.ones((ny, nx), dtype=bool)
    mask[:3] = False
    n = int(mask.sum())
    state = rng.random((ne, n))
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=0, dx=1.0, dE=1.0) as ctx:
        ctx.upload_geometry(mask)
        ctx.set_state(state)
        frames = ctx.get_frames()
        back = ctx.get_state(want_phonons=False)[0]
    assert np.array_equal(back, state)
    assert np.all(np.isnan(frames[:, ~mask]))
    assert np.array_equal(frames[:, mask], state)


def test_snapshot_download_overlaps_stepping_and_matches_get_frames():
    """qpb_frames_snapshot / qpb_frames_download: the frames of the snapshot instant, also when the context keeps
    stepping (and overwriting the state and its work arrays) while a helper thread downloads them."""
    import threading

    from qpsim_b200 import capi

    case = cases.meander_c2(ny=96, nx=96, ne=40, steps=2)       # 2.9 MiB of frames, plain copy path
    big = cases.meander_c2(ny=256, nx=272, ne=40, steps=2)      # 22 MiB: pinned two-chunk pipeline
    for c_ in (case, big):
        mask = c_["mask"]
        ny, nx = mask.shape
        n, ne = int(mask.sum()), c_["num_energy_bins"]
        edges = Q.extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, c_["bc"], Q.BoundaryCondition)
        bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
        E, dE = Q.build_energy_grid(cases.GAP, 1.0, 5.0, ne)
        D = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
        rng = np.random.default_rng(11)
        state = rng.random((ne, n)) * 1e-4
        with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
            ctx.upload_geometry(mask, bcx, bcy, src)
            ctx.upload_diffusion(D)
            ctx.prepare_diffusion(0, 0.5)
            ctx.set_state(state)
            want = ctx.get_frames()
            ctx.frames_snapshot()
            box = {}
            th = threading.Thread(target=lambda: box.update(frames=ctx.frames_download()))
            th.start()
            ctx.advance(3, 0.5)                      # overwrites the state while the download runs
            th.join()
            after = ctx.get_frames()
        assert np.array_equal(box["frames"], want, equal_nan=True)
        assert not np.array_equal(after, want, equal_nan=True)
        assert np.array_equal(want[:, mask], state) and np.all(np.isnan(want[:, ~mask]))


def _walls_by_normal(edges, kinds):
    out = {}
    for e in edges:
        kind, val, aux = kinds[e.normal]
        out[e.edge_id] = Q.BoundaryCondition(kind=kind, value=val, aux_value=aux)
    return out


@pytest.mark.parametrize("shape", [(48, 64), (33, 128), (70, 256), (20, 1024)])
@pytest.mark.parametrize("walls", ["reflective", "mixed_y"])
def test_spectral_direct_solve_matches_oracle(shape, walls, monkeypatch):
    """Full rectangles whose left / right walls are reflective (or carry a flux) take the direct solve: cosine
    transform along x, one tridiagonal system along y per mode (qpb_spectral.cu).  Any wall kind on top / bottom,
    odd row counts, sources.  Same run with the sweep iteration (QPB_NO_SPECTRAL=1): both within the bar."""
    ny, nx = shape
    mask = np.ones((ny, nx), dtype=bool)
    edges = Q.extract_edge_segments(mask)
    if walls == "reflective":
        kinds = {k: ("reflective", None, None) for k in ("up", "down", "left", "right")}
    else:
        kinds = {"up": ("absorbing", None, None), "down": ("robin", 0.7, 0.2), "left": ("reflective", None, None),
                 "right": ("neumann", 0.05, None)}
    bcs = _walls_by_normal(edges, kinds)
    field = cases.lognormal_field(mask, seed=3, scale=1e-4)
    kw = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=field, diffusion_coefficient=cases.D0, dt=0.4,
              total_time=1.0, dx=1.0, store_every=1, energy_gap=cases.GAP, energy_min_factor=1.0,
              energy_max_factor=3.0, num_energy_bins=5, enable_diffusion=True, dynes_gamma=cases.GAMMA,
              enforce_pauli=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, mass, _, ef, _ = Q.run_2d_crank_nicolson(**kw)
        info = dict(Q.solver.last_run_info)
        # the same solve with the right-hand side built by its own pass (k_build_rhs) instead of inside the Thomas pass
        monkeypatch.setenv("QPB_NO_SPECTRAL_FUSED", "1")
        _, _, _, _, ef_unfused, _ = Q.run_2d_crank_nicolson(**kw)
        assert Q.solver.last_run_info["sweep_path"] == 4
        monkeypatch.delenv("QPB_NO_SPECTRAL_FUSED")
        monkeypatch.setenv("QPB_NO_SPECTRAL", "1")
        _, _, mass_pr, _, ef_pr, _ = Q.run_2d_crank_nicolson(**kw)
        info_pr = dict(Q.solver.last_run_info)
    assert info["sweep_path"] == 4 and info_pr["sweep_path"] != 4
    helpers.assert_close(np.array([[f[mask] for f in t] for t in ef]), np.array([[f[mask] for f in t] for t in ef_unfused]),
                         "fused vs separate right-hand side", rtol=1e-10)   # rounding order only (b formed in real / mode space)
    res = O.run(mask, edges, bcs, field, cases.D0, 0.4, 1.0, 1.0, store_every=1, gap=cases.GAP, fmin=1.0, fmax=3.0,
                ne=5, gamma=cases.GAMMA)
    want = np.array(res.state_frames)
    got = np.array([[f[mask] for f in t] for t in ef])
    got_pr = np.array([[f[mask] for f in t] for t in ef_pr])
    helpers.assert_close(got, want, "spectral n(E,cell)")
    helpers.assert_close(got_pr, want, "sweep-iteration n(E,cell)")
    np.testing.assert_allclose(mass, res.mass, rtol=helpers.RTOL)


def test_spectral_path_is_not_taken_where_it_does_not_apply():
    """A wall term on the left, a wall kind that changes along the top wall, a hole, a row length that is not a power of
    two: all stay with the sweep iteration (and still match, see the other tests)."""
    def path(mask, kinds=None, custom=None):
        edges = Q.extract_edge_segments(mask)
        bcs = custom(edges) if custom else _walls_by_normal(edges, kinds)
        Q.run_2d_crank_nicolson(mask=mask, edges=edges, edge_conditions=bcs, initial_field=np.where(mask, 1e-4, 0.0),
                                diffusion_coefficient=6.0, dt=0.5, total_time=0.5, dx=1.0, energy_gap=180.0,
                                energy_max_factor=3.0, num_energy_bins=3)
        return Q.solver.last_run_info["sweep_path"]

    refl = {k: ("reflective", None, None) for k in ("up", "down", "left", "right")}
    full = np.ones((32, 64), dtype=bool)
    assert path(full, refl) == 4
    assert path(full, dict(refl, left=("absorbing", None, None))) != 4
    assert path(np.ones((32, 80), dtype=bool), refl) != 4
    holed = full.copy()
    holed[10:12, 20:30] = False
    assert path(holed, custom=lambda edges: {e.edge_id: Q.BoundaryCondition(kind="reflective") for e in edges}) != 4
