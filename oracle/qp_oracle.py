"""CPU oracle for the qpsim time-stepping hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the algorithm in the reference's
``qpsim/solver.py`` (read-only mount ``/root/reference``).  It exists so that
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs have something to check the CUDA path against on a
box where the reference itself is not present.  The product package never
imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` (run in the build
container, where the reference is importable) executes the unmodified reference
and stores its outputs under ``tests/golden/*.npz``; ``tests/test_oracle.py``
checks this restatement against those fixtures on every CPU run, and against
the live reference when ``/root/reference`` exists.

Every function cites the reference lines it restates (``qpsim/solver.py`` unless
another file is named).  The restatement is vectorised over cells where the
reference loops in Python, but keeps the reference's arithmetic (same clamps,
same branch thresholds, same operation grouping where rounding could matter).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np
from scipy import sparse
from scipy.sparse import linalg as spla

KB_UEV_PER_K = 86.17333262145  # solver.py:347


def _exp_numpy(x):
    return np.exp(x)


def _exp_correctly_rounded(x):
    """fl(1 + expm1(x)) for small |x|: what a correctly rounded libm exp returns.  numpy's SIMD exp (AVX512
    builds) is off by one ulp for ~5-10 % of arguments; the reference's (exp(x)-1)/b and (1-exp(-mu dt))/mu
    (solver.py:661, 697) amplify that ulp by 1/|x|, so the reference itself is only defined up to this choice.
    Tests use the two variants to measure that band."""
    x = np.asarray(x, dtype=float)
    return np.where(np.abs(x) < 0.25, 1.0 + np.expm1(x), np.exp(x))


EXP = _exp_numpy  # swap with _exp_correctly_rounded to evaluate the libm-dependent band


# --------------------------------------------------------------------------
# A6: grids, density of states, kernels  (solver.py:61-84, 324-342, 350-370,
#     429-490, 668-683)
# --------------------------------------------------------------------------
def energy_grid(gap: float, fmin: float, fmax: float, ne: int):
    """solver.py:61-84 build_energy_grid."""
    if gap <= 0:
        raise ValueError("gap must be positive.")
    if ne <= 0:
        raise ValueError("num_energy_bins must be >= 1.")
    lo, hi = fmin * gap, fmax * gap
    if ne == 1:
        return np.array([0.5 * (lo + hi)], dtype=float), 1.0
    if hi <= lo:
        raise ValueError("energy_max_factor must be > energy_min_factor for num_energy_bins > 1.")
    dE = (hi - lo) / float(ne)
    return lo + (np.arange(ne, dtype=float) + 0.5) * dE, dE


def dos(E: np.ndarray, gap: float, gamma: float) -> np.ndarray:
    """solver.py:324-342 BCS / Dynes density of states."""
    E = np.asarray(E, dtype=float)
    if gamma <= 0:
        out = np.zeros_like(E)
        ok = E > gap
        out[ok] = E[ok] / np.sqrt(E[ok] ** 2 - gap ** 2)
        return out
    z = E - 1j * gamma
    with np.errstate(invalid="ignore"):
        val = np.real(z / np.sqrt(z ** 2 - gap ** 2))
    return np.maximum(val, 0.0)


def bose(omega: np.ndarray, T: float) -> np.ndarray:
    """solver.py:350-370 thermal_phonon_occupation."""
    omega = np.asarray(omega, dtype=float)
    if T <= 0:
        return np.zeros_like(omega)
    kT = KB_UEV_PER_K * float(T)
    x = np.minimum(omega / max(kT, 1e-30), 500.0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        occ = 1.0 / (np.exp(x) - 1.0)
    occ[~np.isfinite(occ)] = 0.0
    return np.maximum(occ, 0.0)


def thermal_weights(E: np.ndarray, gap: float, T: float, gamma: float = 0.0) -> np.ndarray:
    """solver.py:429-460 thermal_qp_weights."""
    rho = dos(E, gap, gamma)
    if T <= 0:
        return np.zeros_like(rho)
    kT = KB_UEV_PER_K * T
    return rho * (1.0 / (np.exp(np.minimum(E / kT, 500.0)) + 1.0))


def kr0(E: np.ndarray, gap: float, tau: float, Tc: float) -> np.ndarray:
    """solver.py:463-474 recombination_kernel_base."""
    kTc = KB_UEV_PER_K * Tc
    s = E[:, None] + E[None, :]
    p = E[:, None] * E[None, :]
    coh = 1.0 + gap ** 2 / np.maximum(p, 1e-30)
    return (1.0 / tau) * (s / kTc) ** 2 / kTc * coh


def ks0(E: np.ndarray, gap: float, tau: float, Tc: float) -> np.ndarray:
    """solver.py:477-490 scattering_kernel_base."""
    kTc = KB_UEV_PER_K * Tc
    d = E[:, None] - E[None, :]
    p = E[:, None] * E[None, :]
    coh = np.maximum(1.0 - gap ** 2 / np.maximum(p, 1e-30), 0.0)
    out = (1.0 / tau) * (d ** 2) / kTc ** 3 * coh
    np.fill_diagonal(out, 0.0)
    return out


def phonon_map(E: np.ndarray):
    """solver.py:668-683 _build_phonon_frequency_map."""
    E = np.asarray(E, dtype=float)
    ne = E.size
    dabs = np.abs(E[:, None] - E[None, :])
    ssum = E[:, None] + E[None, :]
    vals = np.concatenate([dabs.ravel(), ssum.ravel()])
    omega, inv = np.unique(np.round(vals, 12), return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    idx_diff = inv[: ne * ne].reshape(ne, ne)
    idx_sum = inv[ne * ne:].reshape(ne, ne)
    sign = np.sign(E[:, None] - E[None, :]).astype(np.int8)
    return omega, idx_diff, idx_sum, sign


# --------------------------------------------------------------------------
# A2/A3: Laplacian with boundary faces  (solver.py:112-212, 235-321)
# --------------------------------------------------------------------------
_DIRS = (("up", -1, 0), ("down", 1, 0), ("left", 0, -1), ("right", 0, 1))  # solver.py:25-30


def _bc_tuple(bc) -> tuple[str, float, float]:
    kind = bc.kind.strip().lower()
    val = float(bc.value or 0.0)
    aux = float(getattr(bc, "aux_value", None) or 0.0)
    return kind, val, aux


def face_lookup(edges, edge_conditions) -> dict:
    """solver.py:37-50 _build_face_bc_lookup."""
    out = {}
    for e in edges:
        bc = edge_conditions.get(e.edge_id)
        if bc is None:
            continue
        for f in e.faces:
            out[(int(f.row), int(f.col), f.direction)] = _bc_tuple(bc)
    return out


def laplacian(mask: np.ndarray, edges, edge_conditions, dx: float, D_cell: np.ndarray | None = None):
    """solver.py:152-212 (uniform) and :235-321 (variable D, harmonic-mean faces).

    Returns (L csr, source[N]).  With ``D_cell`` the operator already contains D
    (the reference's ``L_D``) and ``source`` is scaled by the cell's D.
    """
    mask = np.asarray(mask, dtype=bool)
    ny, nx = mask.shape
    idx = -np.ones(mask.shape, dtype=np.int64)
    coords = np.argwhere(mask)  # solver.py:53-58: row-major order
    n = len(coords)
    if n == 0:
        raise ValueError("Geometry mask has no interior points.")
    idx[mask] = np.arange(n)
    faces = face_lookup(edges, edge_conditions)
    inv_dx = 1.0 / dx
    inv_dx2 = inv_dx * inv_dx
    rows, cols, vals = [], [], []
    src = np.zeros(n)
    for p, (r, c) in enumerate(coords):
        Dp = 1.0 if D_cell is None else float(D_cell[p])
        for name, dr, dc in _DIRS:
            rr, cc = r + dr, c + dc
            if 0 <= rr < ny and 0 <= cc < nx and mask[rr, cc]:
                q = int(idx[rr, cc])
                if D_cell is None:
                    w = inv_dx2
                else:
                    Dq = float(D_cell[q])
                    w = 2.0 * Dp * Dq / max(Dp + Dq, 1e-30) * inv_dx2  # solver.py:283
                rows += [p, p]
                cols += [p, q]
                vals += [-w, w]
                continue
            bc = faces.get((int(r), int(c), name))
            if bc is None:
                raise ValueError(f"Missing boundary condition for face at cell ({r}, {c}) direction '{name}'.")
            kind, val, aux = bc
            if kind == "reflective":
                pass
            elif kind == "absorbing":  # solver.py:125-129
                rows.append(p); cols.append(p); vals.append(-2.0 * Dp * inv_dx2)
            elif kind == "dirichlet":  # solver.py:130-136
                rows.append(p); cols.append(p); vals.append(-2.0 * Dp * inv_dx2)
                src[p] += 2.0 * Dp * val * inv_dx2
            elif kind == "neumann":  # solver.py:137-140
                src[p] += Dp * val * inv_dx
            elif kind == "robin":  # solver.py:141-148
                rows.append(p); cols.append(p); vals.append(-Dp * val * inv_dx)
                src[p] += Dp * aux * inv_dx
            else:
                raise ValueError(f"Unsupported boundary kind: {kind}")
    L = sparse.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    return L, src


class DiffusionCN:
    """solver.py:221-232, 1143-1174, 1428-1452: per-bin unsplit Crank-Nicolson with SuperLU."""

    def __init__(self, mask, edges, edge_conditions, dx, D_array: np.ndarray, dt: float, variable: bool):
        self.ne = D_array.shape[0]
        self.dt = dt
        self.ops = []
        n = int(np.sum(mask))
        eye = sparse.eye(n, format="csc")
        if not variable:
            L, src = laplacian(mask, edges, edge_conditions, dx)
            for i in range(self.ne):
                Di = float(D_array[i, 0])
                a = 0.5 * dt * Di
                self.ops.append(((eye + a * L).tocsr(), spla.splu((eye - a * L).tocsc()), dt * Di * src))
        else:
            for i in range(self.ne):
                L, src = laplacian(mask, edges, edge_conditions, dx, D_array[i])
                a = 0.5 * dt
                self.ops.append(((eye + a * L).tocsr(), spla.splu((eye - a * L).tocsc()), dt * src))

    def step(self, state: np.ndarray) -> None:
        for i in range(self.ne):
            B, lu, s = self.ops[i]
            state[i] = lu.solve(B @ state[i] + s)


# --------------------------------------------------------------------------
# A5: local coupled quasiparticle + phonon collision update
#     (solver.py:640-665, 686-700, 703-791), vectorised over cells.
# --------------------------------------------------------------------------
def relax_update(n, gain, loss, dt):
    """solver.py:640-665 _apply_time_relaxation_update."""
    mu = np.maximum(loss, 0.0)
    P = np.maximum(gain + (mu - loss) * n, 0.0)
    decay = EXP(-mu * dt)
    small = mu < 1e-14
    with np.errstate(divide="ignore", invalid="ignore"):
        coeff = np.where(small, dt, (1.0 - decay) / np.where(small, 1.0, mu))
    return np.maximum(decay * n + coeff * P, 0.0)


def affine_growth(y, a, b, dt):
    """solver.py:686-700 _solve_affine_growth."""
    x = np.clip(b * dt, -80.0, 80.0)
    ex = EXP(x)
    small = np.abs(b) < 1e-14
    with np.errstate(divide="ignore", invalid="ignore"):
        coeff = np.where(small, dt, (ex - 1.0) / np.where(small, 1.0, b))
    return np.maximum(ex * y + coeff * a, 0.0)


def collide_pixel(n, nph, Kr, Ks, rho, idx_diff, idx_sum, sign, dE, dt, *, recomb: bool, scat: bool):
    """solver.py:703-791 for ONE cell, statement by statement (used to pin the batched form)."""
    rho_s = np.maximum(rho, 1e-30)
    w = np.maximum(1.0 - n / rho_s, 0.0)
    gain = np.zeros_like(n)
    loss = np.zeros_like(n)
    nS = nph[idx_sum]
    nD = nph[idx_diff]
    if scat and Ks is not None:
        Np = np.where(sign > 0, 1.0 + nD, nD)
        np.fill_diagonal(Np, 0.0)
        Ke = Ks * Np
        gain += dE * rho * w * (Ke.T @ n)
        loss += dE * ((Ke * rho[None, :]) @ w)
    p = rho * w
    if recomb and Kr is not None:
        loss += 2.0 * dE * ((Kr * (1.0 + nS)) @ n)
        gain += 2.0 * dE * p * ((Kr * nS) @ p)
    n_new = relax_update(n, gain, loss, dt)
    if not (scat or recomb):
        return n_new, nph
    a = np.zeros_like(nph)
    b = np.zeros_like(nph)
    nw = nph.size
    if scat and Ks is not None:
        S = dE * (n[:, None] * Ks * (rho[None, :] * w[None, :]))
        em = sign > 0
        ab = sign < 0
        if np.any(em):
            t = np.bincount(idx_diff[em].ravel(), weights=S[em].ravel(), minlength=nw)
            a += t
            b += t
        if np.any(ab):
            b -= np.bincount(idx_diff[ab].ravel(), weights=S[ab].ravel(), minlength=nw)
    if recomb and Kr is not None:
        R = dE * (n[:, None] * Kr * n[None, :])
        t = np.bincount(idx_sum.ravel(), weights=R.ravel(), minlength=nw)
        a += t
        b += t
        Bk = dE * (p[:, None] * Kr * p[None, :])
        b -= np.bincount(idx_sum.ravel(), weights=Bk.ravel(), minlength=nw)
    return n_new, affine_growth(nph, a, b, dt)


def collide(state, phonons, Kr, Ks, rho, idx_diff, idx_sum, sign, dE, dt, *,
            recomb: bool, scat: bool, update_phonons: bool = True, chunk: int = 256) -> None:
    """solver.py:794-831 / :834-875 over all cells at once, in place.

    ``state`` is (NE, N), ``phonons`` (Nw, N).  ``Kr``/``Ks``/``rho`` are either
    shared ((NE,NE)/(NE,)) or per cell ((N,NE,NE)/(N,NE)) as in the reference's
    nonuniform driver.
    """
    ne, ncell = state.shape
    nw = phonons.shape[0]
    if not (recomb or scat):
        state[:] = np.maximum(state, 0.0)  # relax_update with zero gain/loss clamps at 0
        return
    # scatter matrices for the bincount sums (built once per call; cell independent)
    em = (sign > 0).ravel()
    ab = (sign < 0).ravel()
    pair = np.arange(ne * ne)
    Md_em = sparse.csr_matrix((np.ones(em.sum()), (idx_diff.ravel()[em], pair[em])), shape=(nw, ne * ne))
    Md_ab = sparse.csr_matrix((np.ones(ab.sum()), (idx_diff.ravel()[ab], pair[ab])), shape=(nw, ne * ne))
    Ms = sparse.csr_matrix((np.ones(ne * ne), (idx_sum.ravel(), pair)), shape=(nw, ne * ne))
    offdiag = 1.0 - np.eye(ne)
    per_cell = rho.ndim == 2
    for c0 in range(0, ncell, chunk):
        c1 = min(ncell, c0 + chunk)
        n = state[:, c0:c1].T.copy()           # (C, NE)
        nph = phonons[:, c0:c1].T.copy()       # (C, Nw)
        r = rho[c0:c1] if per_cell else rho[None, :]
        w = np.maximum(1.0 - n / np.maximum(r, 1e-30), 0.0)
        p = r * w
        gain = np.zeros_like(n)
        loss = np.zeros_like(n)
        nS = nph[:, idx_sum]                   # (C, NE, NE)
        nD = nph[:, idx_diff]
        a = np.zeros_like(nph)
        b = np.zeros_like(nph)
        if scat and Ks is not None:
            K = Ks[c0:c1] if per_cell else Ks[None]
            Np = np.where(sign[None] > 0, 1.0 + nD, nD) * offdiag[None]
            Ke = K * Np
            gain += dE * r * w * np.einsum("cji,cj->ci", Ke, n)
            loss += dE * np.einsum("cij,cj->ci", Ke * (r[:, None, :] if per_cell else r[None]), w)
            S = (dE * (n[:, :, None] * K * p[:, None, :])).reshape(c1 - c0, -1)
            t = (Md_em @ S.T).T
            a += t
            b += t
            b -= (Md_ab @ S.T).T
        if recomb and Kr is not None:
            K = Kr[c0:c1] if per_cell else Kr[None]
            loss += 2.0 * dE * np.einsum("cij,cj->ci", K * (1.0 + nS), n)
            gain += 2.0 * dE * p * np.einsum("cij,cj->ci", K * nS, p)
            R = (dE * (n[:, :, None] * K * n[:, None, :])).reshape(c1 - c0, -1)
            t = (Ms @ R.T).T
            a += t
            b += t
            Bk = (dE * (p[:, :, None] * K * p[:, None, :])).reshape(c1 - c0, -1)
            b -= (Ms @ Bk.T).T
        state[:, c0:c1] = relax_update(n, gain, loss, dt).T
        if update_phonons:
            phonons[:, c0:c1] = affine_growth(nph, a, b, dt).T


# --------------------------------------------------------------------------
# Fixed-bath forward-Euler collision forms (solver.py:551-605) with the bath-dressed kernels
# (solver.py:493-548).  state is (NE, N), modified in place like the reference's helpers.
# --------------------------------------------------------------------------
def kr_dressed(E: np.ndarray, gap: float, tau: float, Tc: float, T_bath: float) -> np.ndarray:
    """solver.py:493-516 recombination_kernel."""
    kT = KB_UEV_PER_K * T_bath
    es = E[:, None] + E[None, :]
    n_p = 1.0 / (np.exp(np.minimum(es / kT, 500.0)) - 1.0) + 1.0 if kT > 0 else np.ones_like(es)
    return kr0(E, gap, tau, Tc) * n_p


def ks_dressed(E: np.ndarray, gap: float, tau: float, Tc: float, T_bath: float) -> np.ndarray:
    """solver.py:519-548 scattering_kernel."""
    kT = KB_UEV_PER_K * T_bath
    ed = E[:, None] - E[None, :]
    if kT > 0:
        with np.errstate(divide="ignore", invalid="ignore"):
            nbe = 1.0 / (np.exp(np.minimum(np.abs(ed) / kT, 500.0)) - 1.0)
        n_p = np.where(ed > 0, 1.0 + nbe, nbe)
    else:
        n_p = np.where(ed > 0, 1.0, 0.0)
    np.fill_diagonal(n_p, 0.0)
    return ks0(E, gap, tau, Tc) * n_p


def euler_scattering_step(state: np.ndarray, K_s: np.ndarray, rho_bins: np.ndarray, dE: float, dt: float) -> None:
    """solver.py:551-581 apply_scattering_step."""
    rho = rho_bins[:, None]
    one_minus_f = np.maximum(1.0 - state / np.maximum(rho, 1e-30), 0.0)
    scat_in = dE * rho * one_minus_f * (K_s.T @ state)
    scat_out = state * dE * ((K_s * rho_bins[None, :]) @ one_minus_f)
    state += dt * (scat_in - scat_out)
    np.maximum(state, 0.0, out=state)


def euler_recombination_step(state: np.ndarray, K_r: np.ndarray, G_therm: np.ndarray, dE: float, dt: float) -> None:
    """solver.py:584-605 apply_recombination_step."""
    recomb_rate = 2.0 * state * dE * (K_r @ state)
    state += dt * (G_therm[:, None] - recomb_rate)
    np.maximum(state, 0.0, out=state)


# --------------------------------------------------------------------------
# A8: Pauli diagnostics  (solver.py:967-996)
# --------------------------------------------------------------------------
def pauli_stats(state, rho_state, floor=1e-18):
    ok = rho_state > 1e-30
    forb = (~ok) & (state > floor)
    forb_idx = None
    if np.any(forb):
        k = np.unravel_index(int(np.argmax(forb)), forb.shape)
        forb_idx = (int(k[0]), int(k[1]))
    f = np.divide(state, np.maximum(rho_state, 1e-30), out=np.zeros_like(state), where=ok)
    k = np.unravel_index(int(np.argmax(f)), f.shape)
    return float(f[k]), (int(k[0]), int(k[1])), forb_idx


# --------------------------------------------------------------------------
# A1: the time loop  (solver.py:1085-1089, 1454-1494; scalar mode 1517-1571)
# --------------------------------------------------------------------------
@dataclass
class OracleResult:
    times: list
    state_frames: list          # [(NE, N)] at every stored time (compressed cell order)
    phonon_frames: list         # [(Nw, N)]
    mass: list
    E: np.ndarray | None
    dE: float
    extra: dict = field(default_factory=dict)


def step_plan(dt: float, total_time: float):
    """solver.py:1085-1089."""
    full = int(np.floor(total_time / dt + 1e-12))
    rem = float(total_time - full * dt)
    if rem < 1e-12:
        rem = 0.0
    return full, rem, full + (1 if rem > 0.0 else 0)


def run(mask, edges, edge_conditions, initial_field, D0, dt, total_time, dx, *, store_every=1,
        gap=0.0, fmin=1.0, fmax=10.0, ne=50, energy_weights=None, diffusion=True, recomb=False,
        scat=False, gamma=0.0, tau_s=440.0, tau_r=440.0, Tc=1.2, T_bath=0.1,
        gext: Callable[[float], Any] | None = None, D_array=None, gap_values=None,
        freeze_phonons=False, qp_state0=None, phonon_state0=None) -> OracleResult:
    """Energy-resolved loop of run_2d_crank_nicolson (solver.py:1092-1515) on compressed cells.

    ``gext(t)`` returns None, a scalar or an (NE,N) array: the reference's
    evaluate_external_generation result at time t (solver.py:878-964).
    ``D_array``/``gap_values`` given together select the nonuniform-gap branch
    (solver.py:1145-1164, 1203-1232).  ``gap == 0`` runs the legacy scalar loop.
    """
    mask = np.asarray(mask, dtype=bool)
    n = int(mask.sum())
    full, rem, total = step_plan(dt, total_time)
    if store_every <= 0:
        store_every = 1
    if gap <= 0.0:
        u = np.asarray(initial_field, dtype=float)[mask].astype(float)
        ops = {}
        if diffusion:
            Darr = np.full((1, n), float(D0))
            ops[dt] = DiffusionCN(mask, edges, edge_conditions, dx, Darr, dt, False)
            if rem > 0.0:
                ops[rem] = DiffusionCN(mask, edges, edge_conditions, dx, Darr, rem, False)
        res = OracleResult([0.0], [u[None].copy()], [], [float(u.sum() * dx * dx)], None, 1.0)
        t = 0.0
        cur = u[None].copy()
        for s in range(1, total + 1):
            h = dt if s <= full else rem
            if diffusion:
                ops[h].step(cur)
            t += h
            if s % store_every == 0 or s == total:
                res.times.append(float(t))
                res.state_frames.append(cur.copy())
                res.mass.append(float(cur.sum() * dx * dx))
        return res

    E, dE = energy_grid(gap, fmin, fmax, ne)
    variable = D_array is not None and gap_values is not None and len(np.unique(gap_values)) > 1
    if D_array is None:
        Db = D0 * np.sqrt(np.maximum(0.0, 1.0 - (gap / E) ** 2))  # solver.py:1135
        D_array = Db[:, None] * np.ones((1, n))
    ops = {}
    if diffusion:
        ops[dt] = DiffusionCN(mask, edges, edge_conditions, dx, D_array, dt, variable)
        if rem > 0.0:
            ops[rem] = DiffusionCN(mask, edges, edge_conditions, dx, D_array, rem, variable)
    omega, idx_diff, idx_sum, sign = phonon_map(E)
    phon = bose(omega, T_bath)[:, None] * np.ones((1, n)) if phonon_state0 is None else np.array(phonon_state0, dtype=float)
    if variable:
        ug = np.unique(gap_values)
        table = {float(g): (dos(E, float(g), gamma),
                            kr0(E, float(g), tau_r, Tc) if recomb else None,
                            ks0(E, float(g), tau_s, Tc) if scat else None) for g in ug}
        rho = np.stack([table[float(g)][0] for g in gap_values])
        Kr = np.stack([table[float(g)][1] for g in gap_values]) if recomb else None
        Ks = np.stack([table[float(g)][2] for g in gap_values]) if scat else None
        rho_state = rho.T
    else:
        rho = dos(E, gap, gamma)
        Kr = kr0(E, gap, tau_r, Tc) if recomb else None
        Ks = ks0(E, gap, tau_s, Tc) if scat else None
        rho_state = rho[:, None] * np.ones((1, n))
    if qp_state0 is not None:
        state = np.array(qp_state0, dtype=float)
    else:
        sv = np.asarray(initial_field, dtype=float)[mask].astype(float)
        if energy_weights is not None:  # solver.py:1254-1270
            w = np.asarray(energy_weights, dtype=float)
            tot = np.sum(w) * dE
            w = w / tot if tot > 0 else np.ones(ne) / (ne * dE)
        else:  # solver.py:1271-1278
            r0 = dos(E, gap, gamma)
            tot = np.sum(r0) * dE
            w = r0 / tot if tot > 0 else np.ones(ne) / (ne * dE)
        state = sv[None, :] * w[:, None]
    coll = bool(recomb or scat)

    def _collide(h):
        if h <= 0.0 or not coll:
            return
        collide(state, phon, Kr, Ks, rho, idx_diff, idx_sum, sign, dE, h,
                recomb=recomb, scat=scat, update_phonons=not freeze_phonons)

    integ = state.sum(axis=0) * dE
    res = OracleResult([0.0], [state.copy()], [phon.copy()], [float(integ.sum() * dx * dx)], E, dE)
    res.extra["pauli"] = [pauli_stats(state, rho_state)]
    res.extra["omega"] = omega
    t = 0.0
    for s in range(1, total + 1):
        h = dt if s <= full else rem
        if gext is not None:
            g = gext(t)
            if g is not None:
                state += h * g
        if coll and diffusion:  # solver.py:1469-1475
            _collide(0.5 * h)
            ops[h].step(state)
            _collide(0.5 * h)
        else:
            _collide(h)
            if diffusion:
                ops[h].step(state)
        res.extra["pauli"].append(pauli_stats(state, rho_state))
        t += h
        if s % store_every == 0 or s == total:
            integ = state.sum(axis=0) * dE
            res.times.append(float(t))
            res.state_frames.append(state.copy())
            res.phonon_frames.append(phon.copy())
            res.mass.append(float(integ.sum() * dx * dx))
    return res
