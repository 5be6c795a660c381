"""Seeded workloads shared by the golden-fixture generator, the oracle tests and the GPU parity tests.

Physics constants follow SURVEY.md section 8(d): Al, gap 180 ueV, D0 = 6 um^2/ns, tau = 440 ns, Tc = 1.2 K,
T_bath = 0.1 K, Dynes gamma = 0.18 ueV, dx = 1 um.  Every case is a dict of plain keyword arguments in the
vocabulary of ``run_2d_crank_nicolson`` plus a "bc" recipe, so the same description can be handed to the
reference, to the oracle and to the CUDA drop-in.
"""
from __future__ import annotations

import numpy as np

GAP = 180.0
D0 = 6.0
TAU = 440.0
TC = 1.2
TBATH = 0.1
GAMMA = 0.18

BC_KINDS = ("reflective", "absorbing", "dirichlet", "neumann", "robin")


def annulus_mask(n: int) -> np.ndarray:
    y, x = np.mgrid[:n, :n] + 0.5
    r = np.hypot(x - n / 2, y - n / 2)
    return (r < 0.42 * n) & (r > 0.19 * n)


def meander_mask(ny: int, nx: int, pad: int = 8, slot: int = 4, pitch: int = 16, gap_len: int = 32) -> np.ndarray:
    """MKID-like meander: padded rectangle with horizontal slots alternating sides."""
    m = np.zeros((ny, nx), dtype=bool)
    m[pad:-pad, pad:-pad] = True
    for i, r in enumerate(range(pad + pitch, ny - pad - slot, pitch)):
        if i % 2 == 0:
            m[r:r + slot, pad:nx - pad - gap_len] = False
        else:
            m[r:r + slot, pad + gap_len:nx - pad] = False
    return m


def make_bcs(edges, recipe: str, bc_cls):
    """recipe: 'reflective' | 'absorbing' | 'mixed' (cycles the five kinds) | 'short_absorbing'."""
    out = {}
    if recipe == "mixed":
        for i, e in enumerate(edges):
            kind = BC_KINDS[i % 5]
            if kind == "robin":
                out[e.edge_id] = bc_cls(kind=kind, value=0.7, aux_value=0.2)
            elif kind in ("dirichlet", "neumann"):
                out[e.edge_id] = bc_cls(kind=kind, value=0.3)
            else:
                out[e.edge_id] = bc_cls(kind=kind)
        return out
    if recipe == "short_absorbing":
        lengths = sorted((len(e.faces), e.edge_id) for e in edges)
        short = {lengths[0][1], lengths[1][1]}
        for e in edges:
            out[e.edge_id] = bc_cls(kind="absorbing" if e.edge_id in short else "reflective")
        return out
    for e in edges:
        out[e.edge_id] = bc_cls(kind=recipe)
    return out


def gaussian_field(mask: np.ndarray, cx=0.4, cy=0.5, sigma=0.08, base=1e-4, amp=2e-4) -> np.ndarray:
    ny, nx = mask.shape
    y, x = np.mgrid[:ny, :nx]
    f = base + amp * np.exp(-(((x + 0.5) / nx - cx) ** 2 + ((y + 0.5) / ny - cy) ** 2) / (2 * sigma ** 2))
    return np.where(mask, f, 0.0)


def lognormal_field(mask: np.ndarray, seed: int, scale=1e-4) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.where(mask, scale * rng.lognormal(size=mask.shape), 0.0)


def strip_c1(steps=20, nx=128, ne=64):
    """BASELINE config 1: 1-D strip, diffusion + scattering + recombination + constant generation."""
    mask = np.ones((1, nx), dtype=bool)
    spatial = 1e-4 + 2e-4 * np.exp(-(((np.arange(nx) + 0.5) / nx - 0.3) ** 2) / (2.0 * 0.06 ** 2))
    return dict(
        name=f"c1_strip_{nx}x{ne}", mask=mask, bc="reflective", initial_field=spatial.reshape(1, nx),
        diffusion_coefficient=D0, dt=0.1, total_time=0.1 * steps, dx=1.0, store_every=max(1, steps // 4),
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=3.0, num_energy_bins=ne,
        weights="thermal", enable_diffusion=True, enable_recombination=True, enable_scattering=True,
        dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=TBATH,
        generation=dict(mode="constant", rate=2e-8),
    )


def meander_c2(ny=48, nx=48, ne=24, steps=3, dt=0.5, fmax=5.0, bc="short_absorbing", pad=4, pitch=12, gap_len=12):
    """BASELINE config 2 (scaled by arguments): masked CN diffusion + collisions + pulse generation."""
    mask = meander_mask(ny, nx, pad=pad, slot=3 if ny < 128 else 4, pitch=pitch, gap_len=gap_len)
    return dict(
        name=f"c2_meander_{ny}x{nx}x{ne}", mask=mask, bc=bc, initial_field=gaussian_field(mask, sigma=0.05 if ny >= 128 else 0.12),
        diffusion_coefficient=D0, dt=dt, total_time=dt * steps, dx=1.0, store_every=max(1, steps),
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=fmax, num_energy_bins=ne,
        weights=None, enable_diffusion=True, enable_recombination=True, enable_scattering=True,
        dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=TBATH,
        generation=dict(mode="pulse", pulse_rate=3e-8, pulse_start=0.0, pulse_duration=1.0),
    )


def rect_case(ny=20, nx=28, ne=6, steps=4, dt=0.5, bc="mixed", collisions=False, seed=7):
    mask = np.ones((ny, nx), dtype=bool)
    return dict(
        name=f"rect_{ny}x{nx}x{ne}_{bc}", mask=mask, bc=bc, initial_field=lognormal_field(mask, seed),
        diffusion_coefficient=D0, dt=dt, total_time=dt * steps, dx=1.0, store_every=2,
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=ne,
        weights=None, enable_diffusion=True, enable_recombination=collisions, enable_scattering=collisions,
        dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=TBATH, generation=None,
    )


def annulus_case(n=32, ne=5, steps=3, dt=0.8, bc="mixed"):
    mask = annulus_mask(n)
    return dict(
        name=f"annulus_{n}x{ne}_{bc}", mask=mask, bc=bc, initial_field=gaussian_field(mask, cx=0.25, cy=0.5, sigma=0.1),
        diffusion_coefficient=D0, dt=dt, total_time=dt * steps + 0.3, dx=1.0, store_every=1,
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=ne,
        weights=None, enable_diffusion=True, enable_recombination=False, enable_scattering=False,
        dynes_gamma=0.0, tau_0=TAU, T_c=TC, bath_temperature=TBATH, generation=None,
    )


def scalar_case(kind="rect", bc="mixed"):
    mask = np.ones((12, 18), dtype=bool) if kind == "rect" else annulus_mask(24)
    return dict(
        name=f"scalar_{kind}_{bc}", mask=mask, bc=bc, initial_field=gaussian_field(mask, base=0.1, amp=1.0, sigma=0.15),
        diffusion_coefficient=2.5, dt=0.4, total_time=2.1, dx=0.8, store_every=2,
        energy_gap=0.0, energy_min_factor=1.0, energy_max_factor=10.0, num_energy_bins=50,
        weights=None, enable_diffusion=True, enable_recombination=False, enable_scattering=False,
        dynes_gamma=0.0, tau_0=TAU, T_c=TC, bath_temperature=TBATH, generation=None,
    )


def collision_only_case(ne=24, n=37, frozen=False, fmax=4.0, only=None):
    """Pure collision run on a strip with diffusion off (single-pixel style groups of the reference suite)."""
    mask = np.ones((1, n), dtype=bool)
    rng = np.random.default_rng(11)
    return dict(
        name=f"coll_{ne}x{n}_{'frozen' if frozen else 'dyn'}_{only or 'both'}", mask=mask, bc="reflective",
        initial_field=(5e-3 * rng.uniform(0.2, 1.0, size=(1, n))),
        diffusion_coefficient=D0, dt=0.25, total_time=1.0, dx=1.0, store_every=2,
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=fmax, num_energy_bins=ne,
        weights=None, enable_diffusion=False, enable_recombination=only in (None, "recomb"),
        enable_scattering=only in (None, "scat"), dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=0.25,
        generation=None, freeze_phonon_dynamics=frozen,
    )


def nonuniform_gap_case(ny=10, nx=14, ne=8, steps=3, smooth=False):
    """Non-uniform gap: per-cell D(E,x) with harmonic-mean faces and per-gap kernel tables.  smooth=True adds a gap
    that varies from cell to cell inside a disc (a trap: every line has its own coefficients, 40 distinct gap values)."""
    mask = np.ones((ny, nx), dtype=bool)
    mask[0, :3] = False
    mask[-2:, -4:] = False
    n = int(mask.sum())
    yy, xx = np.mgrid[:ny, :nx]
    gap_field = np.where(xx < nx // 2, GAP, 0.9 * GAP) + np.where(yy > ny // 2, 4.0, 0.0)
    if smooth:
        r2 = ((xx - 0.6 * nx) ** 2 + (yy - 0.4 * ny) ** 2) / (0.2 * min(ny, nx)) ** 2
        gap_field = gap_field - np.round(20.0 * np.exp(-r2), 0) * 0.5      # quantised: 40 levels
    gap_values = gap_field[mask]
    case = dict(
        name=f"nonuniform_{ny}x{nx}x{ne}" + ("_trap" if smooth else ""), mask=mask, bc="mixed", initial_field=gaussian_field(mask, sigma=0.2),
        diffusion_coefficient=D0, dt=0.3, total_time=0.3 * steps, dx=1.0, store_every=1,
        energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=ne,
        weights=None, enable_diffusion=True, enable_recombination=True, enable_scattering=True,
        dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=TBATH, generation=None,
        gap_values=gap_values,
    )
    return case


def custom_mode_cases():
    """User-expression options of the drop-in boundary (custom generation bodies, initial-condition specs, gap
    expressions; reference call sites solver.py:918-962, 1094-1124, 1186-1196).  Recorded from the unmodified
    reference by tests/golden/make_golden_custom.py into tests/golden/custom_modes.npz."""
    mask = np.ones((12, 16), dtype=bool)
    mask[4:7, 5:9] = False
    mask[0, :2] = False
    base = dict(
        mask=mask, bc="mixed", initial_field=gaussian_field(mask, sigma=0.2), diffusion_coefficient=D0, dt=0.25,
        total_time=1.35, dx=1.0, store_every=2, energy_gap=GAP, energy_min_factor=1.0, energy_max_factor=4.0,
        num_energy_bins=10, weights=None, enable_diffusion=True, enable_recombination=True, enable_scattering=True,
        dynes_gamma=GAMMA, tau_0=TAU, T_c=TC, bath_temperature=TBATH, generation=None,
    )
    out = []
    out.append(dict(base, name="custom_gen_static", generation=dict(
        mode="custom", custom_body="return params.get('a', 1e-8) * (1.0 + x) * np.exp(-(E - 180.0) / 200.0) * (y < 0.8)",
        custom_params={"a": 3e-8})))
    out.append(dict(base, name="custom_gen_timedep", store_every=3, generation=dict(
        mode="custom", custom_body="params['a'] * np.exp(-t / 0.5) * np.where(E < 400.0, 1.0, 0.25) * (0.5 + y * x)",
        custom_params={"a": 5e-8})))
    out.append(dict(base, name="custom_gen_scalar_body", enable_diffusion=False, generation=dict(
        mode="custom", custom_body="return 2e-8 if E < 300 else 0.0", custom_params={})))
    out.append(dict(base, name="ic_full_custom", ic_spec=dict(
        qp_full_custom_enabled=True,
        qp_full_custom_body="return 1e-4 * (1 + np.exp(-((x-0.3)**2 + (y-0.6)**2) / 0.05)) * np.exp(-(E-180.0) / 90.0)",
        phonon_spatial_kind="gaussian", phonon_spatial_params={"amplitude": 1.5, "x0": 0.6, "y0": 0.4, "sigma": 0.3},
        phonon_energy_kind="custom", phonon_energy_custom_body="return params.get('s', 1.0) * np.exp(-E / 60.0)",
        phonon_energy_custom_params={"s": 0.02})))
    out.append(dict(base, name="ic_phonon_full_custom", enable_diffusion=False, ic_spec=dict(
        phonon_full_custom_enabled=True,
        phonon_full_custom_body="return 1e-3 * np.exp(-E / 100.0) * (1.0 + 0.5 * np.sin(3.0 * x) * y)",
        phonon_energy_kind="bose_einstein", phonon_energy_params={"temperature": 0.3})))
    for c in out:
        c["bc"] = "reflective"      # no boundary sources: the generation bodies / initial states carry the run
    out.append(dict(base, name="gap_expression_step", gap_expression="np.where(x > 0.5, 162.0, 180.0) + 3.0 * (y > 0.7)"))
    return out


def golden_cases():
    """The cases stored under tests/golden (small enough for the reference's Python loops)."""
    return [
        strip_c1(steps=8, nx=48, ne=16),
        strip_c1(steps=4, nx=128, ne=64),
        meander_c2(ny=40, nx=40, ne=12, steps=2),
        rect_case(bc="mixed"),
        rect_case(ny=16, nx=16, ne=4, bc="reflective", collisions=True, steps=3),
        annulus_case(),
        scalar_case("rect", "mixed"),
        scalar_case("annulus", "absorbing"),
        collision_only_case(frozen=False),
        collision_only_case(frozen=True),
        collision_only_case(only="scat", ne=16, n=9),
        collision_only_case(only="recomb", ne=16, n=9, fmax=5.0),
        nonuniform_gap_case(),
    ]


def golden_cases_large():
    """Larger recorded cases (tests/golden/make_golden.py large): kept out of golden_cases() so that the CPU suite
    stays short; the GPU parity tests and one oracle test use them."""
    return [nonuniform_gap_case(ny=96, nx=96, ne=16, steps=3, smooth=True)]


def thermal_weights_for(case, physics_mod):
    E, dE = physics_mod.build_energy_grid(case["energy_gap"], case["energy_min_factor"], case["energy_max_factor"],
                                          case["num_energy_bins"])
    w = physics_mod.thermal_qp_weights(E, case["energy_gap"], case["bath_temperature"], case["dynes_gamma"])
    return w / (np.sum(w) * dE)


def precomputed_for(case, physics_mod):
    """The `precomputed` dict of a non-uniform-gap case (what qpsim.precompute.precompute_arrays returns:
    D_array = D0*sqrt(max(0, 1-min(gap/E,1)^2)), gap_values, is_uniform; qpsim/precompute.py:211-228)."""
    if "gap_values" not in case:
        return None
    E, _ = physics_mod.build_energy_grid(case["energy_gap"], case["energy_min_factor"], case["energy_max_factor"],
                                         case["num_energy_bins"])
    gv = np.asarray(case["gap_values"], dtype=float)
    D = np.empty((E.size, gv.size))
    for i in range(E.size):
        ratio = np.minimum(gv / E[i], 1.0)
        D[i] = case["diffusion_coefficient"] * np.sqrt(np.maximum(0.0, 1.0 - ratio ** 2))
    return {"D_array": D, "gap_values": gv, "is_uniform": np.array(len(np.unique(gv)) == 1), "E_bins": E}


def solver_kwargs(case, edges, bcs, gen_spec, physics_mod, ic_spec_cls=None):
    """Keyword arguments for run_2d_crank_nicolson (reference or drop-in)."""
    kw = dict(
        mask=case["mask"], edges=edges, edge_conditions=bcs, initial_field=case["initial_field"],
        diffusion_coefficient=case["diffusion_coefficient"], dt=case["dt"], total_time=case["total_time"],
        dx=case["dx"], store_every=case["store_every"], energy_gap=case["energy_gap"],
        energy_min_factor=case["energy_min_factor"], energy_max_factor=case["energy_max_factor"],
        num_energy_bins=case["num_energy_bins"], enable_diffusion=case["enable_diffusion"],
        enable_recombination=case["enable_recombination"], enable_scattering=case["enable_scattering"],
        dynes_gamma=case["dynes_gamma"], tau_0=case["tau_0"], T_c=case["T_c"],
        bath_temperature=case["bath_temperature"], external_generation=gen_spec,
        freeze_phonon_dynamics=case.get("freeze_phonon_dynamics", False),
    )
    if case.get("weights") == "thermal":
        kw["energy_weights"] = thermal_weights_for(case, physics_mod)
    pre = precomputed_for(case, physics_mod)
    if pre is not None:
        kw["precomputed"] = pre
    if case.get("ic_spec") is not None:
        kw["initial_condition_spec"] = ic_spec_cls(**case["ic_spec"])
    if case.get("gap_expression"):
        kw["gap_expression"] = case["gap_expression"]
    return kw
