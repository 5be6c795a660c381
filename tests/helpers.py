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
    kw = cases.solver_kwargs(case, edges, bcs, gen, Q.physics)
    kw.update(extra)
    hist = {}
    times, frames, mass, limits, eframes, E = Q.run_2d_crank_nicolson(phonon_history_out=hist, **kw)
    out = {"times": np.array(times), "mass": np.array(mass), "limits": np.array(limits)}
    if eframes is not None:
        out["state"] = np.array([[f[mask] for f in t] for t in eframes])
        out["phonons"] = np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]])
        out["frames_nan_outside"] = bool(np.all(np.isnan(frames[-1][~mask]))) if (~mask).any() else True
    else:
        out["state"] = np.array([f[mask] for f in frames])[:, None, :]
    return out


def rel_err(a, b):
    """max |a-b| / max|b| per stored time and energy bin (cells with tiny values are judged against the bin's
    scale), the form SURVEY.md section 8(c) prescribes for the 1e-9 bar."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.max(np.abs(b), axis=-1, keepdims=True)
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale))


def assert_close(got, want, what, rtol=RTOL):
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = rel_err(got, want)
    assert err <= rtol, f"{what}: max relative error {err:.3e} > {rtol:.1e}"
    return err
