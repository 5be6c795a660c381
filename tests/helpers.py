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
# in bins where the source term dominates (see test_phonon_update_within_reference_libm_band).
RTOL_PHONON = 2e-6


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
