"""User expressions of the drop-in: custom generation bodies g(E, x, y, t, params), full custom initial states
F(x, y, E, params), split phonon initial conditions and gap expressions gap(x, y).

The reference evaluates these with a restricted ``eval`` (``qpsim/safe_eval.py``; call sites
``qpsim/solver.py:918-962``, ``qpsim/initial_conditions.py:341-372, 456-507, 542-632``).  This module is the
package's own implementation of the same user-facing language, so that ``run_2d_crank_nicolson`` is standalone:
one expression (optionally prefixed by ``return``) over the named variables, ``np.<whitelisted>``,
``math.<whitelisted>``, ``params.get`` / ``params[...]`` and a few builtins; anything else is a ``ValueError``.
Everything here is host-side setup.  What reaches the device is an array: a time-independent generation body is
evaluated once and stays resident (``qpb_upload_generation``), SURVEY.md section 8(f) rank 3.
"""
from __future__ import annotations

import ast
import math
from typing import Any

import numpy as np

_BUILTINS = {"abs": abs, "min": min, "max": max, "pow": pow, "len": len, "float": float, "int": int, "bool": bool}
_NP_CALLS = frozenset(
    "abs sqrt exp log log10 sin cos tan arcsin arccos arctan sinh cosh tanh where maximum minimum clip power "
    "heaviside arange zeros_like ones_like full_like".split())
_NP_VALUES = frozenset("pi e inf nan float64 float32 int64 int32 bool_".split())
_MATH_CALLS = frozenset("sqrt exp log log10 sin cos tan asin acos atan sinh cosh tanh floor ceil".split())
_MATH_VALUES = frozenset("pi e tau inf nan".split())
_VALUE_ATTRS = frozenset(("size", "shape"))
_NODE_TYPES = (ast.Expression, ast.BoolOp, ast.BinOp, ast.UnaryOp, ast.IfExp, ast.Compare, ast.Call, ast.Name,
               ast.Load, ast.Constant, ast.Attribute, ast.Subscript, ast.Slice, ast.Tuple, ast.List, ast.Dict,
               ast.keyword, ast.operator, ast.unaryop, ast.boolop, ast.cmpop)


def _reject(msg: str):
    raise ValueError(msg)


def _check_attribute(node: ast.Attribute, variables: frozenset, *, called: bool) -> None:
    if node.attr.startswith("__"):
        _reject("Dunder attribute access is not allowed in custom expressions.")
    if not isinstance(node.value, ast.Name):
        _reject("Nested attribute access is not allowed in custom expressions.")
    base, attr = node.value.id, node.attr
    if base == "np":
        ok = attr in _NP_CALLS if called else attr in (_NP_CALLS | _NP_VALUES)
        what = "numpy function" if called else "numpy attribute"
    elif base == "math":
        ok = attr in _MATH_CALLS if called else attr in (_MATH_CALLS | _MATH_VALUES)
        what = "math function" if called else "math attribute"
    elif base == "params":
        ok, what = attr == "get", "params attribute"
    elif base in variables and not called:
        ok, what = attr in _VALUE_ATTRS, "attribute"
    else:
        _reject("Method calls are not allowed in custom expressions." if called
                else f"Unsupported attribute base in custom expression: {base!r}.")
    if not ok:
        _reject(f"Unsupported {what} in custom expression: {base}.{attr}.")


def _validate(tree: ast.AST, variables: frozenset) -> None:
    names_ok = variables | set(_BUILTINS) | {"np", "math"}
    callee_ids = set()
    for node in ast.walk(tree):
        if not isinstance(node, _NODE_TYPES):
            _reject(f"Unsupported syntax in custom expression: {type(node).__name__}.")
        if isinstance(node, ast.Call):
            if any(k.arg is None for k in node.keywords):
                _reject("Starred keyword arguments are not allowed in custom expressions.")
            f = node.func
            if isinstance(f, ast.Name):
                if f.id not in _BUILTINS:
                    _reject(f"Unsupported function in custom expression: {f.id!r}.")
            elif isinstance(f, ast.Attribute):
                _check_attribute(f, variables, called=True)
                callee_ids.add(id(f))
            else:
                _reject("Unsupported call target in custom expressions.")
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            if node.id.startswith("__"):
                _reject("Dunder names are not allowed in custom expressions.")
            if node.id not in names_ok:
                _reject(f"Unsupported name in custom expression: {node.id!r}.")
        elif isinstance(node, ast.Attribute) and id(node) not in callee_ids:
            _check_attribute(node, variables, called=False)
        elif isinstance(node, ast.Subscript):
            if isinstance(node.value, ast.Name) and node.value.id in ("np", "math"):
                _reject("Subscript access on modules is not allowed in custom expressions.")


class Expression:
    """A validated user expression over ``variables``; call it with those variables as keywords."""

    def __init__(self, source: str, variables) -> None:
        text = str(source or "").strip() or "0.0"
        if "\n" not in text and text.startswith("return "):
            text = text[7:].strip()
        self.source = text
        self.variables = tuple(variables)
        try:
            tree = ast.parse(text, mode="eval")
        except SyntaxError as exc:
            raise ValueError(
                "Custom expressions must be a single expression (optionally prefixed by 'return ')."
            ) from exc
        _validate(tree, frozenset(self.variables))
        self.names_used = frozenset(n.id for n in ast.walk(tree) if isinstance(n, ast.Name))
        self._code = compile(tree, "<custom-expression>", "eval")

    def uses(self, name: str) -> bool:
        return name in self.names_used

    def __call__(self, **values: Any) -> Any:
        missing = [v for v in self.variables if v not in values]
        if missing:
            raise ValueError(f"Missing variables for custom expression evaluation: {', '.join(missing)}.")
        scope = {"__builtins__": {}, "np": np, "math": math}
        scope.update(_BUILTINS)
        scope.update(values)
        return eval(self._code, scope, {})


# ---------------------------------------------------------------------------------------------------------------
# translation of a body into a postfix program for the device (include/qpb.h: qpb_gen_op)
# ---------------------------------------------------------------------------------------------------------------
OPS = ("CONST E X Y T ADD SUB MUL DIV POW MOD FLOORDIV NEG NOT TRUTH LT LE GT GE EQ NE AND OR SELECT MIN MAX NPMIN NPMAX "
       "HEAVISIDE ABS SQRT EXP LOG LOG10 SIN COS TAN ASIN ACOS ATAN SINH COSH TANH FLOOR CEIL TRUNC").split()
OP = {name: k for k, name in enumerate(OPS)}
MAX_OPS, MAX_STACK = 512, 24
_UNARY_NP = {"abs": "ABS", "sqrt": "SQRT", "exp": "EXP", "log": "LOG", "log10": "LOG10", "sin": "SIN", "cos": "COS",
             "tan": "TAN", "arcsin": "ASIN", "arccos": "ACOS", "arctan": "ATAN", "sinh": "SINH", "cosh": "COSH",
             "tanh": "TANH"}
_UNARY_MATH = {"sqrt": "SQRT", "exp": "EXP", "log": "LOG", "log10": "LOG10", "sin": "SIN", "cos": "COS", "tan": "TAN",
               "asin": "ASIN", "acos": "ACOS", "atan": "ATAN", "sinh": "SINH", "cosh": "COSH", "tanh": "TANH",
               "floor": "FLOOR", "ceil": "CEIL"}
_BINOPS = {ast.Add: "ADD", ast.Sub: "SUB", ast.Mult: "MUL", ast.Div: "DIV", ast.Pow: "POW", ast.Mod: "MOD",
           ast.FloorDiv: "FLOORDIV"}
_CMPOPS = {ast.Lt: "LT", ast.LtE: "LE", ast.Gt: "GT", ast.GtE: "GE", ast.Eq: "EQ", ast.NotEq: "NE"}
_VALUES = {("np", "pi"): math.pi, ("np", "e"): math.e, ("np", "inf"): math.inf, ("np", "nan"): math.nan,
           ("math", "pi"): math.pi, ("math", "e"): math.e, ("math", "tau"): math.tau, ("math", "inf"): math.inf,
           ("math", "nan"): math.nan}


class _NotTranslatable(Exception):
    pass


def compile_program(expr: "Expression", params: dict, variables=("E", "x", "y", "t")):
    """Translate a validated body into a postfix program ``[(op, value), ...]`` with the value semantics of one scalar
    evaluation per (E, x, y, t) - what the reference's per-value fallback loop computes - or return None when the body
    uses something that has no per-value meaning on the device (``len``, ``.size``, subscripts, ``np.arange``, keyword
    arguments, non-numeric params): such bodies keep the host evaluation.  ``params`` are folded into constants."""
    prog = []

    def const(v):
        if isinstance(v, (bool, np.bool_)):
            v = 1.0 if v else 0.0
        if not isinstance(v, (int, float, np.integer, np.floating)):
            raise _NotTranslatable(f"constant {v!r}")
        prog.append((OP["CONST"], float(v)))

    def param_value(node):
        # params.get("key"[, default]) and params["key"], keys and defaults constant
        if isinstance(node, ast.Subscript):
            key = node.slice
            if not (isinstance(key, ast.Constant) and key.value in params):
                raise _NotTranslatable("params subscript")
            return params[key.value]
        args = node.args
        if node.keywords or not (1 <= len(args) <= 2) or not isinstance(args[0], ast.Constant):
            raise _NotTranslatable("params.get")
        if args[0].value in params:
            return params[args[0].value]
        if len(args) == 1:
            raise _NotTranslatable("params.get without default")   # None
        return ast.literal_eval(args[1]) if not isinstance(args[1], ast.Constant) else args[1].value

    def fold(args, op):   # left fold of a two-argument operator over a call's arguments
        if len(args) < 2:
            raise _NotTranslatable("min/max of one argument")
        emit(args[0])
        for a in args[1:]:
            emit(a)
            prog.append((OP[op], 0.0))

    def emit(node):
        if isinstance(node, ast.Expression):
            return emit(node.body)
        if isinstance(node, ast.Constant):
            return const(node.value)
        if isinstance(node, ast.Name):
            if node.id in variables:
                prog.append((OP[node.id.upper()], 0.0))
                return
            raise _NotTranslatable(f"name {node.id}")
        if isinstance(node, ast.Attribute):
            key = (getattr(node.value, "id", None), node.attr)
            if key in _VALUES:
                return const(_VALUES[key])
            raise _NotTranslatable(f"attribute {key}")
        if isinstance(node, ast.Subscript):
            if isinstance(node.value, ast.Name) and node.value.id == "params":
                return const(param_value(node))
            raise _NotTranslatable("subscript")
        if isinstance(node, ast.UnaryOp):
            emit(node.operand)
            if isinstance(node.op, ast.USub):
                prog.append((OP["NEG"], 0.0))
            elif isinstance(node.op, ast.Not):
                prog.append((OP["NOT"], 0.0))
            elif not isinstance(node.op, ast.UAdd):
                raise _NotTranslatable("unary operator")
            return
        if isinstance(node, ast.BinOp):
            if type(node.op) not in _BINOPS:
                raise _NotTranslatable("binary operator")
            emit(node.left)
            emit(node.right)
            prog.append((OP[_BINOPS[type(node.op)]], 0.0))
            return
        if isinstance(node, ast.BoolOp):
            emit(node.values[0])
            for v in node.values[1:]:
                emit(v)
                prog.append((OP["AND" if isinstance(node.op, ast.And) else "OR"], 0.0))
            return
        if isinstance(node, ast.Compare):
            if any(type(o) not in _CMPOPS for o in node.ops):
                raise _NotTranslatable("comparison operator")
            # a < b < c  ->  (a < b) and (b < c); operands have no side effects, so b may be evaluated twice
            operands = [node.left] + list(node.comparators)
            for k, o in enumerate(node.ops):
                emit(operands[k])
                emit(operands[k + 1])
                prog.append((OP[_CMPOPS[type(o)]], 0.0))
                if k:
                    prog.append((OP["AND"], 0.0))
            return
        if isinstance(node, ast.IfExp):
            emit(node.test)
            emit(node.body)
            emit(node.orelse)
            prog.append((OP["SELECT"], 0.0))
            return
        if isinstance(node, ast.Call):
            if node.keywords:
                raise _NotTranslatable("keyword arguments")
            f, args = node.func, node.args
            if isinstance(f, ast.Name):
                if f.id == "abs" and len(args) == 1:
                    emit(args[0]); prog.append((OP["ABS"], 0.0)); return
                if f.id in ("min", "max"):
                    return fold(args, f.id.upper())
                if f.id == "pow" and len(args) == 2:
                    emit(args[0]); emit(args[1]); prog.append((OP["POW"], 0.0)); return
                if f.id == "float" and len(args) == 1:
                    return emit(args[0])
                if f.id == "int" and len(args) == 1:
                    emit(args[0]); prog.append((OP["TRUNC"], 0.0)); return
                if f.id == "bool" and len(args) == 1:
                    emit(args[0]); prog.append((OP["TRUTH"], 0.0)); return
                raise _NotTranslatable(f"builtin {f.id}")
            base, name = f.value.id, f.attr
            if base == "params":
                return const(param_value(node))
            table = _UNARY_NP if base == "np" else _UNARY_MATH
            if name in table and len(args) == 1:
                emit(args[0]); prog.append((OP[table[name]], 0.0)); return
            if base == "np":
                if name == "where" and len(args) == 3:
                    emit(args[0]); emit(args[1]); emit(args[2]); prog.append((OP["SELECT"], 0.0)); return
                if name in ("maximum", "minimum") and len(args) == 2:
                    emit(args[0]); emit(args[1]); prog.append((OP["NPMAX" if name == "maximum" else "NPMIN"], 0.0)); return
                if name == "clip" and len(args) == 3:
                    emit(args[0]); emit(args[1]); prog.append((OP["NPMAX"], 0.0))
                    emit(args[2]); prog.append((OP["NPMIN"], 0.0)); return
                if name == "power" and len(args) == 2:
                    emit(args[0]); emit(args[1]); prog.append((OP["POW"], 0.0)); return
                if name == "heaviside" and len(args) == 2:
                    emit(args[0]); emit(args[1]); prog.append((OP["HEAVISIDE"], 0.0)); return
                if name in ("zeros_like", "ones_like") and len(args) == 1:
                    return const(0.0 if name == "zeros_like" else 1.0)
                if name == "full_like" and len(args) == 2:
                    return emit(args[1])
            raise _NotTranslatable(f"call {base}.{name}")
        raise _NotTranslatable(type(node).__name__)

    try:
        emit(ast.parse(expr.source, mode="eval"))
    except (_NotTranslatable, ValueError, SyntaxError):
        return None
    # stack discipline (the library checks it again)
    depth = peak = 0
    for op, _ in prog:
        name = OPS[op]
        pops = 0 if op <= OP["T"] else 3 if name == "SELECT" else 1 if (op >= OP["ABS"] or name in ("NEG", "NOT", "TRUTH")) else 2
        depth += 1 - pops
        peak = max(peak, depth)
    if depth != 1 or peak > MAX_STACK or len(prog) > MAX_OPS:
        return None
    return prog


# ---------------------------------------------------------------------------------------------------------------
# grids
# ---------------------------------------------------------------------------------------------------------------
def cell_coordinates(mask: np.ndarray):
    """Normalised cell-centre coordinates of the mask cells in compressed (row-major) order."""
    ny, nx = mask.shape
    rows, cols = np.nonzero(mask)
    return (cols + 0.5) / max(1, nx), (rows + 0.5) / max(1, ny)


# ---------------------------------------------------------------------------------------------------------------
# custom generation (solver.py:918-962)
# ---------------------------------------------------------------------------------------------------------------
class CustomGeneration:
    """g_ext(E, x, y, t, params) on the (NE, N) grid.  ``time_dependent`` is False when the body never names ``t``:
    the array is then evaluated once and can stay on the device for the whole run."""

    def __init__(self, spec, E_bins: np.ndarray, mask: np.ndarray) -> None:
        self.expr = Expression(spec.custom_body.strip() or "0.0", ("E", "x", "y", "t", "params"))
        self.params = dict(spec.custom_params or {})
        self.E = np.asarray(E_bins, dtype=float)
        self.x, self.y = cell_coordinates(np.asarray(mask, dtype=bool))
        self.time_dependent = self.expr.uses("t")
        # a time-dependent body that has a per-value meaning runs on the device (qpb_upload_generation_program)
        self.program = compile_program(self.expr, self.params) if self.time_dependent else None

    def __call__(self, t: float) -> np.ndarray:
        ne, n = self.E.size, self.x.size
        out = np.empty((ne, n), dtype=float)
        try:
            for i in range(ne):    # one bin at a time with a scalar E, like the reference: `max(E, c)` style bodies work
                val = np.asarray(self.expr(E=self.E[i], x=self.x, y=self.y, t=t, params=self.params), dtype=float)
                if val.ndim == 0:
                    out[i] = float(val)
                elif val.size == n:
                    out[i] = val.ravel()
                else:
                    raise ValueError(
                        "Vectorized custom generation must return a scalar or "
                        f"exactly {n} values per energy bin; got {val.size}."
                    )
        except Exception:
            for i in range(ne):
                for c in range(n):
                    out[i, c] = float(self.expr(E=float(self.E[i]), x=float(self.x[c]), y=float(self.y[c]), t=t,
                                                params=self.params))
        if not np.all(np.isfinite(out)):
            raise ValueError("External generation mode 'custom' produced non-finite values.")
        if np.any(out < 0):
            raise ValueError("External generation mode 'custom' produced negative values. "
                             "Generation rates must be non-negative.")
        return out


# ---------------------------------------------------------------------------------------------------------------
# gap expression -> per-cell gap and D(E, x)   (initial_conditions.py:240-289, precompute.py:171-228)
# ---------------------------------------------------------------------------------------------------------------
def _cells_from_xy_expression(expr: Expression, mask: np.ndarray, params: dict):
    x, y = cell_coordinates(mask)
    if x.size == 0:
        return np.empty(0)
    try:
        arr = np.asarray(expr(x=x, y=y, params=params), dtype=float)
        if arr.ndim == 0:
            return np.full(x.size, float(arr))
        if arr.size == x.size:
            return arr.reshape(x.size)
    except Exception:
        pass
    return np.array([float(expr(x=float(a), y=float(b), params=params)) for a, b in zip(x, y)], dtype=float)


def gap_values_from_expression(expression: str, mask: np.ndarray, default_gap: float) -> np.ndarray:
    mask = np.asarray(mask, dtype=bool)
    n = int(mask.sum())
    if not expression.strip():
        vals = np.full(n, float(default_gap))
    else:
        vals = _cells_from_xy_expression(Expression(expression, ("x", "y", "params")), mask, {})
    vals = np.asarray(vals, dtype=float).reshape(-1)
    if vals.size != n:
        raise ValueError(f"Gap expression returned {vals.size} values; expected {n} interior pixels.")
    if not np.all(np.isfinite(vals)):
        raise ValueError("Gap expression produced non-finite values.")
    if np.any(vals <= 0.0):
        raise ValueError("Gap expression must produce strictly positive values.")
    return vals


def precompute_from_gap_expression(expression: str, mask: np.ndarray, E_bins: np.ndarray, default_gap: float,
                                   diffusion_coefficient: float) -> dict:
    """The part of ``precompute_arrays(..., include_collision_kernels=False)`` the solver reads: gap_values,
    is_uniform and D(E, x) = D0 sqrt(max(0, 1 - min(gap(x)/E, 1)^2))."""
    gaps = gap_values_from_expression(expression, mask, default_gap)
    ratio = np.minimum(gaps[None, :] / np.asarray(E_bins, dtype=float)[:, None], 1.0)
    D = diffusion_coefficient * np.sqrt(np.maximum(0.0, 1.0 - ratio ** 2))
    return {"E_bins": np.asarray(E_bins, dtype=float), "gap_values": gaps,
            "is_uniform": np.array(np.unique(gaps).size == 1), "D_array": D}


# ---------------------------------------------------------------------------------------------------------------
# initial conditions from an InitialConditionSpec (duck-typed: any object with the reference's field names)
# ---------------------------------------------------------------------------------------------------------------
_DEFAULT_FULL_BODY = "return np.exp(-((x-0.5)**2 + (y-0.5)**2) / 0.02) * np.exp(-E / 500.0)"
_KB_IC = 86.173303   # ueV/K, the constant initial_conditions.py uses for its Bose-Einstein profile


def _truthy(v) -> bool:
    if isinstance(v, str):
        return v.strip().lower() in ("1", "true", "yes", "on")
    return bool(v)


def _to_energy_cell_array(arr: np.ndarray, ne: int, mask: np.ndarray, label: str) -> np.ndarray:
    ny, nx = mask.shape
    n = int(mask.sum())
    arr = np.asarray(arr, dtype=float)
    if arr.ndim == 0:
        return np.full((ne, n), float(arr))
    by_shape = (
        ((ne, n), lambda a: a),
        ((n, ne), lambda a: a.T),
        ((ne, ny, nx), lambda a: a[:, mask]),
        ((ny, nx, ne), lambda a: np.moveaxis(a, 2, 0)[:, mask]),
        ((ny, nx), lambda a: np.repeat(a[mask][None, :], ne, axis=0)),
        ((ne,), lambda a: np.repeat(a.reshape(ne, 1), n, axis=1)),
        ((n,), lambda a: np.repeat(a.reshape(1, n), ne, axis=0)),
    )
    for shape, conv in by_shape:
        if arr.shape == shape:
            return np.array(conv(arr), dtype=float)
    if arr.size == ne * n:
        return arr.reshape(ne, n).astype(float)
    raise ValueError(
        f"{label} expression returned shape {arr.shape}; expected scalar, "
        f"(N_E,), (N_x*N_y,), (N_E, N_x*N_y), or full-grid shapes tied to mask {mask.shape}."
    )


def full_custom_state(mask: np.ndarray, bins: np.ndarray, body: str, params: dict, label: str) -> np.ndarray:
    """F(x, y, E, params) on (bins, cells): broadcast evaluation first, scalar loop when the body needs it."""
    mask = np.asarray(mask, dtype=bool)
    bins = np.asarray(bins, dtype=float)
    if bins.size <= 0:
        raise ValueError("Energy bins must be non-empty for full custom profile evaluation.")
    expr = Expression(body.strip(), ("x", "y", "E", "params"))
    x, y = cell_coordinates(mask)
    try:
        raw = np.asarray(expr(x=x[None, :], y=y[None, :], E=bins[:, None], params=params), dtype=float)
    except Exception:
        raw = np.array([[float(expr(x=float(a), y=float(b), E=float(e), params=params)) for a, b in zip(x, y)]
                        for e in bins], dtype=float).reshape(bins.size, x.size)
    state = _to_energy_cell_array(raw, bins.size, mask, label)
    if not np.all(np.isfinite(state)):
        raise ValueError(f"{label} expression produced non-finite values.")
    if np.any(state < 0):
        raise ValueError(f"{label} expression must be non-negative.")
    return state


def initial_qp_state(mask, E_bins, spec):
    """Non-separable quasiparticle initial state (NE, N) or None (initial_conditions.py:494-507)."""
    if not _truthy(getattr(spec, "qp_full_custom_enabled", False)):
        return None
    body = str(getattr(spec, "qp_full_custom_body", "") or _DEFAULT_FULL_BODY)
    return full_custom_state(mask, E_bins, body, dict(getattr(spec, "qp_full_custom_params", None) or {}),
                             "Full quasiparticle profile")


def spatial_profile(mask: np.ndarray, kind: str, params: dict, body: str, body_params: dict) -> np.ndarray:
    """Spatial profile on the mask cells (compressed): gaussian / uniform / point / custom
    (initial_conditions.py:210-268)."""
    mask = np.asarray(mask, dtype=bool)
    ny, nx = mask.shape
    x, y = cell_coordinates(mask)
    mode = str(kind or "").strip().lower()
    if mode == "gaussian":
        sigma = max(1e-6, float(params.get("sigma", 0.12)))
        rr = (x - float(params.get("x0", 0.5))) ** 2 + (y - float(params.get("y0", 0.5))) ** 2
        vals = float(params.get("amplitude", 1.0)) * np.exp(-rr / (2.0 * sigma * sigma))
    elif mode == "uniform":
        vals = np.full(x.size, float(params.get("value", 1.0)))
    elif mode == "point":
        col = int(np.clip(round(float(params.get("x0", 0.5)) * (nx - 1)), 0, nx - 1))
        row = int(np.clip(round(float(params.get("y0", 0.5)) * (ny - 1)), 0, ny - 1))
        rows, cols = np.nonzero(mask)
        vals = np.zeros(x.size)
        if x.size:
            hit = np.nonzero((rows == row) & (cols == col))[0]
            k = int(hit[0]) if hit.size else int(np.argmin((rows - row) ** 2 + (cols - col) ** 2))
            vals[k] = float(params.get("value", 1.0))
    elif mode == "custom":
        vals = _cells_from_xy_expression(Expression(body, ("x", "y", "params")), mask, body_params)
    else:
        raise ValueError(f"Unsupported spatial initial-condition kind: '{kind}'.")
    if not np.all(np.isfinite(vals)):
        raise ValueError("Spatial initial-condition profile produced non-finite values.")
    return np.asarray(vals, dtype=float)


def phonon_energy_profile(omega: np.ndarray, spec, bath_temperature: float) -> np.ndarray:
    """Occupation per phonon bin from the split spec: bose_einstein / uniform / custom
    (initial_conditions.py:542-598)."""
    omega = np.asarray(omega, dtype=float).reshape(-1)
    if omega.size == 0:
        raise ValueError("omega_bins must be non-empty.")
    if not np.all(np.isfinite(omega)):
        raise ValueError("omega_bins must contain finite values.")
    if np.any(omega < 0):
        raise ValueError("omega_bins must be non-negative.")
    mode = str(getattr(spec, "phonon_energy_kind", "") or "bose_einstein").strip().lower()
    params = dict(getattr(spec, "phonon_energy_params", None) or {})
    if mode in ("bose_einstein", "be", "thermal"):
        temp = float(params.get("temperature", bath_temperature))
        if temp <= 0.0:
            vals = np.zeros_like(omega)
        else:
            den = np.expm1(np.clip(np.maximum(omega, 0.0) / (_KB_IC * temp), 0.0, 700.0))
            vals = np.divide(1.0, den, out=np.zeros_like(omega), where=den > 0.0)
    elif mode == "uniform":
        v = float(params.get("value", 1.0))
        if v < 0:
            raise ValueError("Uniform phonon energy profile value must be non-negative.")
        vals = np.full_like(omega, v)
    elif mode == "custom":
        body = str(getattr(spec, "phonon_energy_custom_body", "") or "").strip() or "return np.ones_like(E)"
        expr = Expression(body, ("E", "params"))
        cparams = dict(getattr(spec, "phonon_energy_custom_params", None) or {})
        try:
            vals = np.asarray(expr(E=omega, params=cparams), dtype=float)
        except Exception:
            vals = np.array([float(expr(E=float(e), params=cparams)) for e in omega], dtype=float)
        vals = vals.reshape(-1)
        if vals.size == 1:
            vals = np.full_like(omega, float(vals[0]))
        if vals.size != omega.size:
            raise ValueError(
                f"Custom phonon energy profile must return {omega.size} values or a scalar; got {vals.size}.")
    else:
        raise ValueError(f"Unsupported phonon energy initial-condition kind '{mode}'. "
                         "Supported: bose_einstein, uniform, custom.")
    if not np.all(np.isfinite(vals)):
        raise ValueError("Phonon energy profile produced non-finite values.")
    if np.any(vals < 0):
        raise ValueError("Phonon energy profile must be non-negative.")
    return vals


def initial_phonon_state(mask, omega_bins, spec, bath_temperature: float):
    """Phonon initial state from the spec.  Returns ``(state, factors)``: ``factors = (per_bin, per_cell)`` when the
    state is the outer product of a per-bin and a per-cell profile (the device then forms it itself), else None and
    ``state`` is the full (Nw, N) array (initial_conditions.py:601-632)."""
    mask = np.asarray(mask, dtype=bool)
    omega = np.asarray(omega_bins, dtype=float)
    if _truthy(getattr(spec, "phonon_full_custom_enabled", False)):
        body = str(getattr(spec, "phonon_full_custom_body", "") or _DEFAULT_FULL_BODY)
        return full_custom_state(mask, omega, body, dict(getattr(spec, "phonon_full_custom_params", None) or {}),
                                 "Full phonon profile"), None
    kind = str(getattr(spec, "phonon_spatial_kind", "") or "").strip().lower()
    if kind:
        sp = spatial_profile(mask, kind, dict(getattr(spec, "phonon_spatial_params", None) or {}),
                             str(getattr(spec, "phonon_spatial_custom_body", "") or "return 1.0"),
                             dict(getattr(spec, "phonon_spatial_custom_params", None) or {}))
    else:
        sp = spatial_profile(mask, "uniform", {"value": 1.0}, "return 1.0", {})
    en = phonon_energy_profile(omega, spec, bath_temperature)
    state = en.reshape(-1, 1) * sp.reshape(1, -1)
    if not np.all(np.isfinite(state)):
        raise ValueError("Phonon initial state produced non-finite values.")
    if np.any(state < 0):
        raise ValueError("Phonon initial state must be non-negative.")
    return state, (en, sp)
