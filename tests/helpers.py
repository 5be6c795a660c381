"""Shared runners: the same case description through the oracle, the CUDA drop-in, or a golden fixture."""
from __future__ import annotations

import os

import numpy as np

import cases
import qpsim_b200 as Q
from oracle import qp_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerance of BASELINE.json's north_star: max relative error 1e-9 (fp64) on n(x,y,E)
RTOL = 1e-9
# Phonon occupations are an internal state, not the north-star quantity.  The reference updates them with
# (exp(x)-1)/b (solver.py:697), which amplifies a one-ulp difference between libm exp implementations by 1/|x|;
# numpy's own exp changes by an ulp between SIMD builds, so the reference's n_ph is only reproducible to ~1e-7
# per call in bins where the source term dominates, and the difference feeds back over the steps of a run
# (see test_collision_accuracy_against_extended_precision for the per-call statement).
RTOL_PHONON = 1e-4


def collide_pixel_extended(n, nph, Kr, Ks, rho, idx_diff, idx_sum, sign, dE, dt):
    """solver.py:703-791 for one cell evaluated in np.longdouble (x87 80-bit): the value the reference's
    formulas define before float64 rounding.  Used to compare ACCURACY: the float64 reference and the CUDA
    kernel are both roundings of this."""
    L = np.longdouble
    n, nph, rho = n.astype(L), nph.astype(L), rho.astype(L)
    Kr, Ks = Kr.astype(L), Ks.astype(L)
    dE, dt = L(dE), L(dt)
    w = np.maximum(1 - n / np.maximum(rho, L(1e-30)), 0)
    p = rho * w
    nS, nD = nph[idx_sum], nph[idx_diff]
    Np = np.where(sign > 0, 1 + nD, nD)
    np.fill_diagonal(Np, 0)
    Ke = Ks * Np
    gain = dE * rho * w * (Ke.T @ n) + 2 * dE * p * ((Kr * nS) @ p)
    loss = dE * ((Ke * rho[None, :]) @ w) + 2 * dE * ((Kr * (1 + nS)) @ n)
    mu = np.maximum(loss, 0)
    P = np.maximum(gain + (mu - loss) * n, 0)
    decay = np.exp(-mu * dt)
    coeff = np.where(mu < 1e-14, dt, (1 - decay) / np.where(mu < 1e-14, 1, mu))
    n_new = np.maximum(decay * n + coeff * P, 0)
    a = np.zeros_like(nph)
    b = np.zeros_like(nph)
    S = dE * (n[:, None] * Ks * p[None, :])
    np.add.at(a, idx_diff[sign > 0], S[sign > 0])
    np.add.at(b, idx_diff[sign > 0], S[sign > 0])
    np.add.at(b, idx_diff[sign < 0], -S[sign < 0])
    R = dE * (n[:, None] * Kr * n[None, :])
    Bk = dE * (p[:, None] * Kr * p[None, :])
    np.add.at(a, idx_sum.ravel(), R.ravel())
    np.add.at(b, idx_sum.ravel(), (R - Bk).ravel())
    x = np.clip(b * dt, -80, 80)
    ex = np.exp(x)
    cf = np.where(np.abs(b) < 1e-14, dt, (ex - 1) / np.where(np.abs(b) < 1e-14, 1, b))
    return n_new, np.maximum(ex * nph + cf * a, 0)


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def gen_callable(case):
    g = case.get("generation")
    if not g:
        return None
    mode = g["mode"]
    if mode == "constant":
        return lambda t: g["rate"]
    if mode == "pulse":
        return lambda t: (g["pulse_rate"] if g["pulse_start"] <= t < g["pulse_start"] + g["pulse_duration"] else 0.0)
    raise ValueError(mode)


def run_oracle(case):
    mask = case["mask"]
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
    kw = {}
    if case.get("weights") == "thermal":
        kw["energy_weights"] = cases.thermal_weights_for(case, Q.physics)
    pre = cases.precomputed_for(case, Q.physics)
    if pre is not None:
        kw["D_array"] = pre["D_array"]
        kw["gap_values"] = pre["gap_values"]
    res = O.run(mask, edges, bcs, case["initial_field"], case["diffusion_coefficient"], case["dt"],
                case["total_time"], case["dx"], store_every=case["store_every"], gap=case["energy_gap"],
                fmin=case["energy_min_factor"], fmax=case["energy_max_factor"], ne=case["num_energy_bins"],
                diffusion=case["enable_diffusion"], recomb=case["enable_recombination"],
                scat=case["enable_scattering"], gamma=case["dynes_gamma"], tau_s=case["tau_0"], tau_r=case["tau_0"],
                Tc=case["T_c"], T_bath=case["bath_temperature"], gext=gen_callable(case),
                freeze_phonons=case.get("freeze_phonon_dynamics", False), **kw)
    out = {"times": np.array(res.times), "mass": np.array(res.mass), "state": np.array(res.state_frames)}
    if res.phonon_frames:
        out["phonons"] = np.array(res.phonon_frames)
    return out


def run_dropin(case, **extra):
    mask = case["mask"]
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
    gen = Q.ExternalGenerationSpec(**case["generation"]) if case.get("generation") else None
    kw = cases.solver_kwargs(case, edges, bcs, gen, Q.physics, Q.InitialConditionSpec)
    kw.update(extra)
    hist = {}
    times, frames, mass, limits, eframes, E = Q.run_2d_crank_nicolson(phonon_history_out=hist, **kw)
    out = {"times": np.array(times), "mass": np.array(mass), "limits": np.array(limits)}
    if eframes is not None and any(t is None for t in eframes):
        pass   # a rank other than 0 of a sharded run: integrated frames and masses only
    elif eframes is not None:
        out["state"] = np.array([[f[mask] for f in t] for t in eframes])
        out["phonons"] = np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]])
        out["frames_nan_outside"] = bool(np.all(np.isnan(frames[-1][~mask]))) if (~mask).any() else True
    else:
        out["state"] = np.array([f[mask] for f in frames])[:, None, :]
    return out


def rel_err(a, b, floor=1e-6):
    """per stored time and energy bin, cells with |ref| above ``floor`` times the bin's
    maximum are judged by their own relative error |a-b|/|ref|, all other cells by |a-b| / max|ref| of the bin.
    Returns the largest of these numbers (so a single bound covers both halves of the rule)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.max(np.abs(b), axis=-1, keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    diff = np.abs(a - b)
    big = np.abs(b) > floor * scale
    err = np.where(big, diff / np.where(big, np.abs(b), 1.0), diff / scale)
    return float(np.max(err)) if err.size else 0.0


def norm_err(a, b):
    """max |a-b| / max|ref| per stored time and bin (norm-wise; used where the element-wise form has no meaning,
    e.g. phonon histories compared at a looser bound)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.max(np.abs(b), axis=-1, keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def assert_close(got, want, what, rtol=RTOL):
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = rel_err(got, want)
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol:.1e}"
    return err


# ---------------------------------------------------------------------------------------------------------------
# the reference's own validation suite (qpsim/test_cases.py:1133-1178) as recorded by tests/golden/make_golden_suite.py
# ---------------------------------------------------------------------------------------------------------------
SUITE_BC_KINDS = ("reflective", "neumann", "dirichlet", "absorbing", "robin")
SUITE_SCALARS = ("diffusion_coefficient", "dt", "total_time", "dx", "store_every", "energy_gap", "energy_min_factor",
                 "energy_max_factor", "num_energy_bins", "enable_diffusion", "enable_recombination",
                 "enable_scattering", "dynes_gamma", "tau_0", "T_c", "bath_temperature")
_SUITE_INT = ("store_every", "num_energy_bins")
_SUITE_BOOL = ("enable_diffusion", "enable_recombination", "enable_scattering")


def suite_case_ids():
    with np.load(os.path.join(GOLDEN_DIR, "suite_cases.npz")) as z:
        return [str(s) for s in z["case_ids"]]


def load_suite_case(k: int):
    """(kwargs for run_2d_crank_nicolson, expected outputs) of the k-th recorded call of generate_test_suite()."""
    with np.load(os.path.join(GOLDEN_DIR, "suite_cases.npz")) as z:
        p = f"{k:02d}_"
        mask = z[p + "mask"].astype(bool)
        edges = Q.extract_edge_segments(mask)
        assert [e.edge_id for e in edges] == [str(s) for s in z[p + "edge_ids"]], "edge ids differ from the reference's"
        bcs = {}
        for e, (kind, val, aux) in zip(edges, z[p + "bc"]):
            bcs[e.edge_id] = Q.BoundaryCondition(kind=SUITE_BC_KINDS[int(kind)], value=None if np.isnan(val) else float(val),
                                                 aux_value=None if np.isnan(aux) else float(aux))
        kw = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=z[p + "initial_field"])
        for name, v in zip(SUITE_SCALARS, z[p + "scalars"]):
            kw[name] = int(v) if name in _SUITE_INT else bool(v) if name in _SUITE_BOOL else float(v)
        if p + "energy_weights" in z.files:
            kw["energy_weights"] = z[p + "energy_weights"]
        want = {"times": z[p + "times"], "mass": z[p + "mass"], "keep": z[p + "keep"], "state": z[p + "state"]}
    return kw, want


def run_suite_case_dropin(kw, keep):
    times, frames, mass, limits, eframes, E = Q.run_2d_crank_nicolson(**kw)
    mask = kw["mask"]
    if eframes is not None:
        state = np.array([[f[mask] for f in eframes[t]] for t in keep])
    else:
        state = np.array([frames[t][mask] for t in keep])[:, None, :]
    return {"times": np.array(times), "mass": np.array(mass), "state": state}


def run_suite_case_oracle(kw, keep):
    res = O.run(kw["mask"], kw["edges"], kw["edge_conditions"], kw["initial_field"], kw["diffusion_coefficient"],
                kw["dt"], kw["total_time"], kw["dx"], store_every=kw["store_every"], gap=kw["energy_gap"],
                fmin=kw["energy_min_factor"], fmax=kw["energy_max_factor"], ne=kw["num_energy_bins"],
                energy_weights=kw.get("energy_weights"), diffusion=kw["enable_diffusion"],
                recomb=kw["enable_recombination"], scat=kw["enable_scattering"], gamma=kw["dynes_gamma"],
                tau_s=kw["tau_0"], tau_r=kw["tau_0"], Tc=kw["T_c"], T_bath=kw["bath_temperature"])
    return {"times": np.array(res.times), "mass": np.array(res.mass),
            "state": np.array([res.state_frames[t] for t in keep])}
