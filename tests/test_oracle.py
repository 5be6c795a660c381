"""The CPU oracle against the golden fixtures produced by the unmodified reference (and the live reference when
/root/reference exists).  Runs everywhere without a GPU."""
import numpy as np
import pytest

import cases
import helpers
import qpsim_b200 as Q
from oracle import qp_oracle as O
from refimport import load_reference

GOLDEN = cases.golden_cases()


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_oracle_matches_reference_fixture(case):
    want = helpers.load_golden(case["name"])
    got = helpers.run_oracle(case)
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-12)
    helpers.assert_close(got["state"], want["state"], "n(E,cell)", rtol=1e-11)
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-11)
    if "phonons" in want and "phonons" in got:
        helpers.assert_close(got["phonons"], want["phonons"][1:] if got["phonons"].shape[0] + 1 == want["phonons"].shape[0]
                             else want["phonons"], "n_ph(omega,cell)", rtol=1e-11)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_oracle_tables_and_pixel_update(tag):
    z = helpers.load_golden("tables_and_pixels")
    ne, fmin, fmax, gamma = z[f"{tag}_params"]
    E, dE = O.energy_grid(cases.GAP, fmin, fmax, int(ne))
    assert np.array_equal(E, z[f"{tag}_E"]) and dE == float(z[f"{tag}_dE"])
    assert np.array_equal(O.dos(E, cases.GAP, gamma), z[f"{tag}_rho"])
    assert np.array_equal(O.kr0(E, cases.GAP, cases.TAU, cases.TC), z[f"{tag}_Kr"])
    assert np.array_equal(O.ks0(E, cases.GAP, cases.TAU, cases.TC), z[f"{tag}_Ks"])
    om, idd, ids, sg = O.phonon_map(E)
    assert np.array_equal(om, z[f"{tag}_omega"]) and np.array_equal(idd, z[f"{tag}_idx_diff"])
    assert np.array_equal(ids, z[f"{tag}_idx_sum"]) and np.array_equal(sg, z[f"{tag}_sign"])
    assert np.array_equal(O.bose(om, 0.25), z[f"{tag}_nph_thermal"])
    assert np.array_equal(O.thermal_weights(E, cases.GAP, 0.25, gamma), z[f"{tag}_qp_weights"])
    n, ph = z[f"{tag}_n_in"].copy(), z[f"{tag}_ph_in"].copy()
    # statement-by-statement pixel form: bit exact
    for c in range(n.shape[1]):
        a, b = O.collide_pixel(n[:, c], ph[:, c], z[f"{tag}_Kr"], z[f"{tag}_Ks"], z[f"{tag}_rho"], idd, ids, sg, dE, 0.3,
                               recomb=True, scat=True)
        assert np.array_equal(a, z[f"{tag}_n_out"][:, c]) and np.array_equal(b, z[f"{tag}_ph_out"][:, c])
    # batched form: same to rounding
    O.collide(n, ph, z[f"{tag}_Kr"], z[f"{tag}_Ks"], z[f"{tag}_rho"], idd, ids, sg, dE, 0.3, recomb=True, scat=True)
    helpers.assert_close(n.T, z[f"{tag}_n_out"].T, "batched n", rtol=1e-13)
    helpers.assert_close(ph.T, z[f"{tag}_ph_out"].T, "batched n_ph", rtol=1e-13)


@pytest.mark.skipif(load_reference() is None, reason="reference tree not present on this box")
def test_oracle_against_live_reference():
    import qpsim.solver as S

    rng = np.random.default_rng(3)
    E, dE = S.build_energy_grid(cases.GAP, 1.0, 3.0, 20)
    om, idd, ids, sg = S._build_phonon_frequency_map(E)
    rho = S._dynes_density_of_states(E, cases.GAP, 0.18)
    Kr = S.recombination_kernel_base(E, cases.GAP, 300.0, 1.2)
    Ks = S.scattering_kernel_base(E, cases.GAP, 500.0, 1.2)
    state = rho[:, None] * rng.uniform(0, 0.6, (20, 11))
    ph = S.thermal_phonon_occupation(om, 0.3)[:, None] * rng.uniform(0.5, 2, (om.size, 11))
    for rec, sc in ((True, True), (True, False), (False, True)):
        s1, p1, s2, p2 = state.copy(), ph.copy(), state.copy(), ph.copy()
        S.apply_collision_step_fischer_catelani_uniform(s1, p1, Kr, Ks, rho, idd, ids, sg, dE, 0.4,
                                                        enable_recombination=rec, enable_scattering=sc)
        O.collide(s2, p2, Kr, Ks, rho, idd, ids, sg, dE, 0.4, recomb=rec, scat=sc)
        helpers.assert_close(s2.T, s1.T, "n", rtol=1e-13)
        helpers.assert_close(p2.T, p1.T, "n_ph", rtol=1e-13)


@pytest.mark.parametrize("tag", ["small", "gemm", "wide"])
def test_oracle_euler_forms_match_reference_fixture(tag):
    """Fixed-bath Euler forms and the bath-dressed kernels against tests/golden/euler_steps.npz (generated from the
    unmodified reference by tests/golden/make_golden_euler.py)."""
    g = helpers.load_golden("euler_steps")
    ne, ncell, fmax, tbath = g[f"{tag}_params"]
    E, dE = O.energy_grid(cases.GAP, 1.0, float(fmax), int(ne))
    np.testing.assert_array_equal(E, g[f"{tag}_E"])
    Kr = O.kr_dressed(E, cases.GAP, cases.TAU, cases.TC, float(tbath))
    Ks = O.ks_dressed(E, cases.GAP, cases.TAU, cases.TC, float(tbath))
    np.testing.assert_allclose(Kr, g[f"{tag}_Kr"], rtol=1e-14)
    np.testing.assert_allclose(Ks, g[f"{tag}_Ks"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(Q.physics.recombination_kernel(E, cases.GAP, cases.TAU, cases.TC, float(tbath)), g[f"{tag}_Kr"], rtol=1e-14)
    np.testing.assert_allclose(Q.physics.scattering_kernel(E, cases.GAP, cases.TAU, cases.TC, float(tbath)), g[f"{tag}_Ks"], rtol=1e-14)
    n_eq = O.thermal_weights(E, cases.GAP, float(tbath), cases.GAMMA)
    np.testing.assert_allclose(Q.physics.thermal_generation(n_eq, g[f"{tag}_Kr"], dE), g[f"{tag}_G_therm"], rtol=1e-14)
    dt = float(g[f"{tag}_dt"])
    s = g[f"{tag}_state"].copy()
    O.euler_scattering_step(s, g[f"{tag}_Ks"], g[f"{tag}_rho"], dE, dt)
    np.testing.assert_allclose(s, g[f"{tag}_after_scattering"], rtol=1e-13, atol=0)
    s = g[f"{tag}_state"].copy()
    O.euler_recombination_step(s, g[f"{tag}_Kr"], g[f"{tag}_G_therm"], dE, dt)
    np.testing.assert_allclose(s, g[f"{tag}_after_recombination"], rtol=1e-13, atol=0)
    s = g[f"{tag}_state"].copy()
    for _ in range(3):
        O.euler_scattering_step(s, g[f"{tag}_Ks"], g[f"{tag}_rho"], dE, dt)
        O.euler_recombination_step(s, g[f"{tag}_Kr"], g[f"{tag}_G_therm"], dE, dt)
    np.testing.assert_allclose(s, g[f"{tag}_after_3_pairs"], rtol=1e-12, atol=0)


SUITE_IDS = helpers.suite_case_ids()


@pytest.mark.parametrize("k", range(len(SUITE_IDS)), ids=SUITE_IDS)
def test_oracle_matches_reference_validation_suite(k):
    """All 28 runs of the reference's generate_test_suite() (qpsim/test_cases.py:1133-1178), recorded from the
    unmodified reference by tests/golden/make_golden_suite.py."""
    kw, want = helpers.load_suite_case(k)
    got = helpers.run_suite_case_oracle(kw, want["keep"])
    np.testing.assert_allclose(got["times"], want["times"], rtol=0, atol=1e-9)
    helpers.assert_close(got["state"], want["state"], "field", rtol=1e-11)
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-11, atol=1e-300)


def test_oracle_matches_reference_on_the_nonuniform_trap():
    case = cases.golden_cases_large()[0]
    want = helpers.load_golden(case["name"])
    got = helpers.run_oracle(case)
    helpers.assert_close(got["state"][want["keep"]], want["state"], "n(E,cell)", rtol=1e-11)
    np.testing.assert_allclose(got["mass"], want["mass"], rtol=1e-11)
