"""Host-side tables of the collision step, built once per run and uploaded.

Restates the closed-form builders the reference keeps on the host:
``build_energy_grid`` (qpsim/solver.py:61-84), BCS / Dynes density of states (:324-342), Bose occupation
(:350-370), thermal quasiparticle weights (:429-460), base recombination / scattering kernels (:463-490) and
the phonon frequency map (:668-683).  Pure numpy, O(NE^2), no per-cell work.
"""
from __future__ import annotations

import numpy as np

KB_UEV_PER_K = 86.17333262145  # Boltzmann constant in micro-eV / K (solver.py:347)


def build_energy_grid(gap: float, energy_min_factor: float, energy_max_factor: float, num_energy_bins: int):
    if gap <= 0:
        raise ValueError("gap must be positive.")
    if num_energy_bins <= 0:
        raise ValueError("num_energy_bins must be >= 1.")
    e_lo = energy_min_factor * gap
    e_hi = energy_max_factor * gap
    if num_energy_bins == 1:
        return np.array([0.5 * (e_lo + e_hi)], dtype=float), 1.0
    if e_hi <= e_lo:
        raise ValueError("energy_max_factor must be > energy_min_factor for num_energy_bins > 1.")
    width = (e_hi - e_lo) / float(num_energy_bins)
    centres = e_lo + (np.arange(num_energy_bins, dtype=float) + 0.5) * width
    return centres, width


def density_of_states(E: np.ndarray, gap: float, gamma: float = 0.0) -> np.ndarray:
    E = np.asarray(E, dtype=float)
    if gamma <= 0:
        rho = np.zeros_like(E)
        above = E > gap
        rho[above] = E[above] / np.sqrt(E[above] ** 2 - gap ** 2)
        return rho
    z = E - 1j * gamma
    with np.errstate(invalid="ignore"):
        rho = np.real(z / np.sqrt(z ** 2 - gap ** 2))
    return np.maximum(rho, 0.0)


def thermal_phonon_occupation(omega_bins: np.ndarray, temperature: float) -> np.ndarray:
    omega = np.asarray(omega_bins, dtype=float)
    if omega.ndim != 1:
        raise ValueError("omega_bins must be a 1D array.")
    if np.any(~np.isfinite(omega)):
        raise ValueError("omega_bins must contain only finite values.")
    if np.any(omega < 0):
        raise ValueError("omega_bins must be non-negative.")
    if temperature <= 0:
        return np.zeros_like(omega)
    kT = KB_UEV_PER_K * float(temperature)
    x = np.minimum(omega / max(kT, 1e-30), 500.0)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        occ = 1.0 / (np.exp(x) - 1.0)
    occ[~np.isfinite(occ)] = 0.0
    return np.maximum(occ, 0.0)


def thermal_qp_weights(E: np.ndarray, gap: float, temperature: float, dynes_gamma: float = 0.0) -> np.ndarray:
    rho = density_of_states(E, gap, dynes_gamma)
    if temperature <= 0:
        return np.zeros_like(rho)
    kT = KB_UEV_PER_K * temperature
    return rho * (1.0 / (np.exp(np.minimum(E / kT, 500.0)) + 1.0))


def recombination_kernel_base(E: np.ndarray, gap: float, tau_0: float, T_c: float) -> np.ndarray:
    kTc = KB_UEV_PER_K * T_c
    e_sum = E[:, None] + E[None, :]
    e_prod = E[:, None] * E[None, :]
    coherence = 1.0 + gap ** 2 / np.maximum(e_prod, 1e-30)
    return (1.0 / tau_0) * (e_sum / kTc) ** 2 / kTc * coherence


def scattering_kernel_base(E: np.ndarray, gap: float, tau_0: float, T_c: float) -> np.ndarray:
    kTc = KB_UEV_PER_K * T_c
    e_diff = E[:, None] - E[None, :]
    e_prod = E[:, None] * E[None, :]
    coherence = np.maximum(1.0 - gap ** 2 / np.maximum(e_prod, 1e-30), 0.0)
    out = (1.0 / tau_0) * (e_diff ** 2) / kTc ** 3 * coherence
    np.fill_diagonal(out, 0.0)
    return out


def recombination_kernel(E: np.ndarray, gap: float, tau_0: float, T_c: float, bath_temperature: float) -> np.ndarray:
    """Bath-dressed K^r (solver.py:493-516): base kernel x (1 + n_BE(E_i + E_j)) at the phonon bath temperature."""
    kT = KB_UEV_PER_K * bath_temperature
    e_sum = E[:, None] + E[None, :]
    if kT > 0:
        n_p = 1.0 / (np.exp(np.minimum(e_sum / kT, 500.0)) - 1.0) + 1.0
    else:
        n_p = np.ones_like(e_sum, dtype=float)
    return recombination_kernel_base(E, gap, tau_0, T_c) * n_p


def scattering_kernel(E: np.ndarray, gap: float, tau_0: float, T_c: float, bath_temperature: float) -> np.ndarray:
    """Bath-dressed K^s (solver.py:519-548): emission 1 + n_BE, absorption n_BE, zero diagonal."""
    kT = KB_UEV_PER_K * bath_temperature
    e_diff = E[:, None] - E[None, :]
    if kT > 0:
        arg = np.minimum(np.abs(e_diff) / kT, 500.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            n_be = 1.0 / (np.exp(arg) - 1.0)
        n_p = np.where(e_diff > 0, 1.0 + n_be, n_be)
    else:
        n_p = np.where(e_diff > 0, 1.0, 0.0)
    np.fill_diagonal(n_p, 0.0)
    return scattering_kernel_base(E, gap, tau_0, T_c) * n_p


def thermal_generation(n_eq: np.ndarray, K_r: np.ndarray, dE: float) -> np.ndarray:
    """G_therm = 2 n_eq dE (K_r n_eq): the generation that balances recombination at n_eq (precompute.py:240)."""
    return 2.0 * n_eq * dE * (K_r @ n_eq)


def phonon_frequency_map(E: np.ndarray):
    """omega grid = unique(round(|Ei-Ej| U Ei+Ej, 12)); index maps into it; sign(Ei-Ej)."""
    E = np.asarray(E, dtype=float)
    if E.ndim != 1:
        raise ValueError("E_bins must be a 1D array.")
    ne = E.size
    values = np.concatenate([np.abs(E[:, None] - E[None, :]).ravel(), (E[:, None] + E[None, :]).ravel()])
    omega, inverse = np.unique(np.round(values, 12), return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    idx_diff = inverse[: ne * ne].reshape(ne, ne)
    idx_sum = inverse[ne * ne:].reshape(ne, ne)
    sign = np.sign(E[:, None] - E[None, :]).astype(np.int8)
    return omega, idx_diff, idx_sum, sign


def integration_widths_from_centers(centers: np.ndarray, *, fallback_width: float = 1.0) -> np.ndarray:
    """Quadrature weight of every bin of a strictly increasing grid of bin centres: the distance between the midpoints
    to its two neighbours, the outermost bins extended symmetrically about their centre (role of solver.py:87-109 for
    the phonon history integrals; a single bin gets ``fallback_width``)."""
    c = np.asarray(centers, dtype=float).ravel()
    if c.size == 0:
        raise ValueError("centers must be non-empty.")
    if c.size == 1:
        return np.full(1, float(fallback_width))
    if not np.all(np.isfinite(c)):
        raise ValueError("centers must contain finite values.")
    gaps = np.diff(c)
    if not np.all(gaps > 0):
        raise ValueError("centers must be strictly increasing.")
    mid = c[:-1] + 0.5 * gaps
    walls = np.concatenate(([c[0] - 0.5 * gaps[0]], mid, [c[-1] + 0.5 * gaps[-1]]))
    widths = walls[1:] - walls[:-1]
    if not np.all(widths > 0):
        raise ValueError("Derived non-positive integration width from centers.")
    return widths
