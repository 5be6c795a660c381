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


# ---- translation into the device's postfix program (include/qpb.h: qpb_gen_op) ------------------------------------------
def _run_program(prog, E, x, y, t):
    """The program's meaning, one value at a time (what k_generation_program does per thread)."""
    import math
    st = []
    for op, val in prog:
        name = U.OPS[op]
        if name == "CONST": st.append(val)
        elif name in ("E", "X", "Y", "T"): st.append({"E": E, "X": x, "Y": y, "T": t}[name])
        elif name == "SELECT":
            b, a, c = st.pop(), st.pop(), st.pop(); st.append(a if c != 0.0 else b)
        elif name in ("NEG", "NOT", "TRUTH") or op >= U.OP["ABS"]:
            a = st.pop()
            f = {"NEG": lambda v: -v, "NOT": lambda v: 1.0 if v == 0.0 else 0.0, "TRUTH": lambda v: 1.0 if v != 0.0 else 0.0,
                 "ABS": abs, "SQRT": np.sqrt, "EXP": np.exp, "LOG": np.log, "LOG10": np.log10, "SIN": np.sin, "COS": np.cos,
                 "TAN": np.tan, "ASIN": np.arcsin, "ACOS": np.arccos, "ATAN": np.arctan, "SINH": np.sinh, "COSH": np.cosh,
                 "TANH": np.tanh, "FLOOR": np.floor, "CEIL": np.ceil, "TRUNC": np.trunc}[name]
            st.append(float(f(a)))
        else:
            b, a = st.pop(), st.pop()
            f = {"ADD": lambda: a + b, "SUB": lambda: a - b, "MUL": lambda: a * b, "DIV": lambda: a / b,
                 "POW": lambda: float(np.power(a, b)), "MOD": lambda: float(np.mod(a, b)), "FLOORDIV": lambda: float(np.floor(a / b)),
                 "LT": lambda: float(a < b), "LE": lambda: float(a <= b), "GT": lambda: float(a > b), "GE": lambda: float(a >= b),
                 "EQ": lambda: float(a == b), "NE": lambda: float(a != b), "AND": lambda: b if a != 0.0 else a,
                 "OR": lambda: a if a != 0.0 else b, "MIN": lambda: b if b < a else a, "MAX": lambda: b if b > a else a,
                 "NPMIN": lambda: float(np.minimum(a, b)), "NPMAX": lambda: float(np.maximum(a, b)),
                 "HEAVISIDE": lambda: float(np.heaviside(a, b))}[name]
            st.append(f())
    assert len(st) == 1
    return st[0]


TRANSLATABLE = [
    "params['a'] * np.exp(-t / 0.5) * np.where(E < 400.0, 1.0, 0.25) * (0.5 + y * x)",
    "1e-8 * (1 + math.sin(6.0 * t) ** 2) * max(E, 250.0, 100 * t) / 250.0 * abs(x - 0.5)",
    "(2e-8 if E < 300 else 5e-9) * (t < 0.4 or x > 0.7) + 1e-9 * (0.2 < y <= 0.6 < 1)",
    "np.clip(1e-8 * np.power(E / 200.0, -1.5) * np.heaviside(0.5 - t, 0.5), 1e-10, 4e-9) + 1e-9 * (int(10 * x) % 3)",
    "1e-8 * np.minimum(np.maximum(x, 0.3), y + 0.1) * np.tanh(t) * np.sqrt(E) / (1.0 + np.log10(E)) + (7 // 2) * float(t > 0)",
    "np.full_like(x, 3e-9) * (not (t > 1.0)) * np.ones_like(y) + pow(x, 2) * bool(E > 0) * np.cos(t) ** 2 - np.zeros_like(x)",
    "params.get('b', 2.5) * t + params.get('a') * np.pi + math.tau * -x + +y + (t and x) + min(x, y)",
]
NOT_TRANSLATABLE = ["x.size * t + len(y)", "np.where(x > 0.5, 1.0, 0.0)[0:2] * t", "np.clip(x, a_min=0.2, a_max=0.4) * t",
                    "np.arange(3)[0] * t", "params.get('missing') * t", "params['s'] * t", "(x, t)[0]"]


@pytest.mark.parametrize("body", TRANSLATABLE)
def test_translated_program_means_what_the_body_means(body):
    params = {"a": 3e-8, "s": "text"}
    expr = U.Expression(body, ("E", "x", "y", "t", "params"))
    prog = U.compile_program(expr, params)
    assert prog is not None and len(prog) <= U.MAX_OPS
    rng = np.random.default_rng(3)
    for E, x, y, t in zip(rng.uniform(180, 700, 40), rng.uniform(0, 1, 40), rng.uniform(0, 1, 40), rng.uniform(0, 2, 40)):
        want = float(expr(E=float(E), x=float(x), y=float(y), t=float(t), params=params))
        got = _run_program(prog, float(E), float(x), float(y), float(t))
        assert got == pytest.approx(want, rel=1e-15, abs=0.0), (body, E, x, y, t)


@pytest.mark.parametrize("body", NOT_TRANSLATABLE)
def test_bodies_without_a_per_value_meaning_stay_on_the_host(body):
    expr = U.Expression(body, ("E", "x", "y", "t", "params"))
    assert U.compile_program(expr, {"a": 3e-8, "s": "text"}) is None


def test_only_time_dependent_bodies_are_translated():
    mask = np.ones((4, 5), dtype=bool)
    E = np.linspace(180.0, 400.0, 3)
    static = U.CustomGeneration(Q.ExternalGenerationSpec(mode="custom", custom_body="1e-8 * x"), E, mask)
    timed = U.CustomGeneration(Q.ExternalGenerationSpec(mode="custom", custom_body="1e-8 * x * t"), E, mask)
    assert static.program is None and not static.time_dependent
    assert timed.program is not None and timed.time_dependent
