"""The package's own user-expression layer (userexpr.py) against what the UNMODIFIED reference's evaluators made of
the same bodies (tests/golden/custom_modes.npz, written by tests/golden/make_golden_custom.py), and - when the
reference tree is present - against its safe_eval accept/reject decisions.  CPU only."""
import numpy as np
import pytest

import cases
import helpers
import qpsim_b200 as Q
from qpsim_b200 import userexpr as U
from refimport import load_reference

CASES = {c["name"]: c for c in cases.custom_mode_cases()}


@pytest.fixture(scope="module")
def gold():
    return helpers.load_golden("custom_modes")


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c["generation"]])
def test_custom_generation_matches_reference_evaluator(name, gold):
    case = CASES[name]
    spec = Q.ExternalGenerationSpec(**case["generation"])
    gen = U.CustomGeneration(spec, gold[name + "/E"], case["mask"])
    assert gen.time_dependent == ("timedep" in name)
    for k, t in enumerate((0.0, 0.75)):
        np.testing.assert_array_equal(gen(t), gold[name + f"/gext_{k}"])


def test_initial_states_match_reference_builders(gold):
    for name in ("ic_full_custom", "ic_phonon_full_custom"):
        case = CASES[name]
        spec = Q.InitialConditionSpec(**case["ic_spec"])
        qp0 = U.initial_qp_state(case["mask"], gold[name + "/E"], spec)
        if name + "/qp0" in gold:
            np.testing.assert_array_equal(qp0, gold[name + "/qp0"])
        else:
            assert qp0 is None
        ph0, factors = U.initial_phonon_state(case["mask"], gold[name + "/omega"], spec, case["bath_temperature"])
        np.testing.assert_array_equal(ph0, gold[name + "/ph0"])
        if factors is not None:
            np.testing.assert_array_equal(np.outer(*factors), ph0)


def test_gap_expression_matches_reference(gold):
    case = CASES["gap_expression_step"]
    E = gold["gap_expression_step/E"]
    pre = U.precompute_from_gap_expression(case["gap_expression"], case["mask"], E, case["energy_gap"],
                                           case["diffusion_coefficient"])
    np.testing.assert_array_equal(pre["gap_values"], gold["gap_expression_step/gap_values"])
    assert not bool(pre["is_uniform"]) and pre["D_array"].shape == (E.size, int(case["mask"].sum()))
    with pytest.raises(ValueError, match="strictly positive"):
        U.gap_values_from_expression("x - 0.5", case["mask"], 180.0)


BODIES = [
    "return 1.0", "np.exp(-x) * E", "params.get('a', 2.0) * t", "max(E, 200.0) + abs(x)", "x if t < 1 else y",
    "np.where(x > 0.5, 1.0, 0.0)[0:2]", "math.sin(t) ** 2", "np.clip(x, a_min=0.2, a_max=0.4)", "x.size + len(y)",
    "__import__('os').system('echo unsafe')", "np.linalg.norm(x)", "x.__class__", "(lambda: 1)()", "open('f')",
    "np.load('f')", "x.sum()", "params.pop('a')", "[q for q in x]", "np['exp'](x)", "E; x", "math.exp.__self__",
    "globals()", "params.get(*x)", "np.exp(**params)", "x @ y", "",
]


@pytest.mark.parametrize("body", BODIES)
def test_expression_whitelist(body):
    """Accepted bodies evaluate to what plain Python gives; rejected ones raise ValueError before anything runs."""
    names = ("E", "x", "y", "t", "params")
    vals = dict(E=250.0, x=np.array([0.1, 0.6, 0.9]), y=np.array([0.2, 0.3, 0.4]), t=0.5, params={"a": 3.0})
    ref = load_reference()
    verdict = None
    if ref is not None:
        from qpsim.safe_eval import compile_safe_expression
        try:
            fn = compile_safe_expression(body, variable_names=names)
            verdict = ("ok", fn)
        except ValueError:
            verdict = ("reject", None)
    try:
        ex = U.Expression(body, names)
    except ValueError:
        assert verdict is None or verdict[0] == "reject", body
        assert any(tok in body for tok in ("__", "linalg", "lambda", "open", "load", ".sum", "pop", " for ", "np[",
                                           ";", "globals", "*x", "**", "@")), body
        return
    assert verdict is None or verdict[0] == "ok", body
    got = ex(**vals)
    if verdict is not None:
        np.testing.assert_array_equal(np.asarray(got, dtype=float), np.asarray(verdict[1](**vals), dtype=float))
    with pytest.raises(ValueError, match="Missing variables"):
        ex(E=1.0)
