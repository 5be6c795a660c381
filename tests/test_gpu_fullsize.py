"""Parity at BASELINE.json's full shapes, through the C ABI.

The oracle cannot run whole steps at these sizes in test time, so every stage is checked the way SURVEY 8(c)
prescribes: collisions on seeded samples of cells against the oracle's per-cell update (cells are independent),
diffusion on sampled bins against the oracle's SuperLU Crank-Nicolson solve (bins are independent) or, on the
full rectangular grids, against the Crank-Nicolson equations themselves evaluated with a numpy stencil, plus the
size-independent properties (number conservation under reflective walls, positivity).
"""
import numpy as np
import pytest

import cases
import helpers
import qpsim_b200 as Q
from oracle import qp_oracle as O
from qpsim_b200 import capi

pytestmark = pytest.mark.gpu


def _tables(ne, fmax, tbath=cases.TBATH):
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, fmax, ne)
    rho = Q.density_of_states(E, cases.GAP, cases.GAMMA)
    Kr = Q.recombination_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    Ks = Q.scattering_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    nph = Q.thermal_phonon_occupation(om, tbath)
    D = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
    return dict(E=E, dE=dE, rho=rho, Kr=Kr, Ks=Ks, om=om, idd=idd, ids=ids, sg=sg, nph=nph, D=D)


@pytest.mark.parametrize("side", [256, 512], ids=["C2_256x256x128", "C5_512x512x128"])
def test_mkid_mask_full_size_stages(side):
    """BASELINE configs[1] (and the per-run shape of configs[4]): meander mask, 128 bins, dynamic phonons.
    One collision half step and one Crank-Nicolson solve at full size; sampled cells / bins against the oracle."""
    scale = side // 256
    mask = cases.meander_mask(side, side, pad=8 * scale, slot=4 * scale, pitch=16 * scale, gap_len=32 * scale)
    ny, nx = mask.shape
    n = int(mask.sum())
    t = _tables(128, 5.0)
    ne, nw = 128, t["om"].size
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    field = cases.gaussian_field(mask, cx=0.4, cy=0.5, sigma=0.05, base=1e-4, amp=2e-4)[mask]
    rng = np.random.default_rng(20260102)
    state = (t["rho"] / (t["rho"].sum() * t["dE"]))[:, None] * (field * np.exp(0.2 * rng.standard_normal(n)))[None, :]
    phon = t["nph"][:, None] * (1.0 + 0.1 * rng.random((1, n)))
    dt = 0.5
    flags = capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=nw, ncell=n, flags=flags, dx=1.0, dE=t["dE"]) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src)
        ctx.upload_diffusion(t["D"])
        ctx.prepare_diffusion(0, dt)
        ctx.upload_collision(t["Kr"][None], t["Ks"][None], t["rho"][None], None, t["idd"], t["ids"], t["sg"])
        ctx.set_state(state, phon)
        ctx.collide(0.5 * dt)
        s1, p1 = ctx.get_state()
        ctx.diffuse(0)
        s2, _ = ctx.get_state(want_phonons=False)
        info = ctx.diag()
    # collisions: 96 sampled cells
    pick = np.sort(rng.choice(n, size=96, replace=False))
    so, po = state[:, pick].copy(), phon[:, pick].copy()
    O.collide(so, po, t["Kr"], t["Ks"], t["rho"], t["idd"], t["ids"], t["sg"], t["dE"], 0.5 * dt, recomb=True, scat=True)
    helpers.assert_close(s1[:, pick].T, so.T, "collision half step n(E) on sampled cells")
    helpers.assert_close(p1[:, pick].T, po.T, "collision half step n_ph on sampled cells", rtol=helpers.RTOL_PHONON)
    # diffusion: 3 sampled bins through the oracle's SuperLU solve of the unsplit system
    bins = [1, 64, 127]
    op = O.DiffusionCN(mask, edges, bcs, 1.0, t["D"][bins][:, None] * np.ones((1, n)), dt, False)
    want = s1[bins].copy()
    op.step(want)
    helpers.assert_close(s2[bins], want, "Crank-Nicolson solve of sampled bins")
    assert info["sweep_path"] == (5 if side == 256 else 2) and info["direct_mode"] == 0, info
    assert np.all(np.isfinite(s2))   # Crank-Nicolson itself is not positivity preserving (nor is the reference's)


def _cn_residual(u_old, u_new, a):
    """max-norm of (I - aL) u_new - (I + aL) u_old on a full rectangle with reflective walls (solver.py:152-232)."""
    def lap(u):
        out = np.zeros_like(u)
        out[:, 1:, :] += u[:, :-1, :] - u[:, 1:, :]
        out[:, :-1, :] += u[:, 1:, :] - u[:, :-1, :]
        out[:, :, 1:] += u[:, :, :-1] - u[:, :, 1:]
        out[:, :, :-1] += u[:, :, 1:] - u[:, :, :-1]
        return out
    a = a[:, None, None]
    return np.max(np.abs((u_new - a * lap(u_new)) - (u_old + a * lap(u_old))), axis=(1, 2))


@pytest.mark.parametrize("solver", ["spectral", "sweeps"])
@pytest.mark.parametrize("shape", [(2048, 2048, 256, 3.0, 0.2), (1024, 1024, 512, 10.0, 0.05)],
                         ids=["C3_2048x2048", "C4_1024x1024"])
def test_large_grid_diffusion_solves_the_cn_system(shape, solver, monkeypatch):
    """BASELINE configs[2] and [3]: the full grids, the four bins of the named energy grid with the largest and
    smallest diffusion coefficients (bins are independent solves).  The result must satisfy the reference's unsplit
    Crank-Nicolson equations to the solver tolerance and conserve the number exactly (reflective walls) - on the
    direct spectral solve these grids take by default and on the segmented pipelined sweeps (QPB_NO_SPECTRAL=1)."""
    monkeypatch.setenv("QPB_NO_SPECTRAL", "0" if solver == "spectral" else "1")
    ny, nx, ne_full, fmax, dt = shape
    t = _tables(ne_full, fmax)
    bins = [1, 2, ne_full - 2, ne_full - 1]
    D = t["D"][bins]
    mask = np.ones((ny, nx), dtype=bool)
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    u0 = np.stack([cases.lognormal_field(mask, seed=20260103 + k, scale=1e-4)[mask] for k in range(len(bins))])
    with capi.Context(ny=ny, nx=nx, ne=len(bins), nw=0, ncell=ny * nx, flags=capi.F_DIFFUSION, dx=1.0, dE=t["dE"]) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src)
        ctx.upload_diffusion(D)
        ctx.prepare_diffusion(0, dt)
        ctx.set_state(u0)
        ctx.advance(1, dt)
        u1, _ = ctx.get_state(want_phonons=False)
        info = ctx.diag()
    res = _cn_residual(u0.reshape(-1, ny, nx), u1.reshape(-1, ny, nx), 0.5 * dt * D)
    scale = np.max(np.abs(u1), axis=1)
    assert np.all(res <= 4e-12 * scale), (res / scale)
    np.testing.assert_allclose(u1.sum(axis=1), u0.sum(axis=1), rtol=1e-12)
    assert info["sweep_path"] == (4 if solver == "spectral" else 3), info


@pytest.mark.parametrize("cfg", [(256, 3.0, 8192, 192), (512, 10.0, 2048, 64)], ids=["C3_NE256", "C4b_NE512"])
def test_large_energy_grid_dynamic_collisions_sampled(cfg):
    """BASELINE configs[2] and [3] (variant b): the coupled quasiparticle/phonon update at 256 and 512 bins on a
    block of cells with seeded lognormal occupations; a sample of cells against the oracle."""
    ne, fmax, ncell, nsample = cfg
    t = _tables(ne, fmax)
    nw = t["om"].size
    rng = np.random.default_rng(20260104)
    wts = O.thermal_weights(t["E"], cases.GAP, 0.3, cases.GAMMA)
    state = wts[:, None] * (1e-4 * np.exp(0.5 * rng.standard_normal((1, ncell)))) * np.exp(0.1 * rng.standard_normal((ne, ncell)))
    phon = t["nph"][:, None] * (1.0 + 0.2 * rng.random((nw, ncell)))
    with capi.Context(ny=1, nx=ncell, ne=ne, nw=nw, ncell=ncell, flags=capi.F_SCATTERING | capi.F_RECOMBINATION,
                      dx=1.0, dE=t["dE"]) as ctx:
        ctx.upload_geometry(np.ones((1, ncell), np.uint8))
        ctx.upload_collision(t["Kr"][None], t["Ks"][None], t["rho"][None], None, t["idd"], t["ids"], t["sg"])
        ctx.set_state(state, phon)
        ctx.collide(0.1)
        s1, p1 = ctx.get_state()
    pick = np.sort(rng.choice(ncell, size=nsample, replace=False))
    so, po = state[:, pick].copy(), phon[:, pick].copy()
    O.collide(so, po, t["Kr"], t["Ks"], t["rho"], t["idd"], t["ids"], t["sg"], t["dE"], 0.1, recomb=True, scat=True,
              chunk=16)
    helpers.assert_close(s1[:, pick].T, so.T, "n(E) on sampled cells")
    helpers.assert_close(p1[:, pick].T, po.T, "n_ph on sampled cells", rtol=helpers.RTOL_PHONON)
    assert np.all(s1 >= 0.0) and np.all(p1 >= 0.0)


def test_c4a_frozen_uniform_tensor_core_gemm_sampled():
    """BASELINE configs[3] variant a: 512 bins, frozen spatially uniform thermal phonons -> the FP64 tensor-core GEMM
    path; sampled cells against the oracle, and the unsampled rest against the general structured kernel."""
    ne, ncell, nsample = 512, 4096 + 37, 64
    t = _tables(ne, 10.0, tbath=0.25)
    nw = t["om"].size
    rng = np.random.default_rng(20260105)
    wts = O.thermal_weights(t["E"], cases.GAP, 0.3, cases.GAMMA)
    state = wts[:, None] * (1e-4 * np.exp(0.5 * rng.standard_normal((1, ncell)))) * np.exp(0.1 * rng.standard_normal((ne, ncell)))
    flags = capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_FREEZE_PHONONS
    outs = []
    import os
    for no_uniform in ("0", "1"):
        os.environ["QPB_NO_UNIFORM"] = no_uniform
        try:
            with capi.Context(ny=1, nx=ncell, ne=ne, nw=nw, ncell=ncell, flags=flags, dx=1.0, dE=t["dE"]) as ctx:
                ctx.upload_geometry(np.ones((1, ncell), np.uint8))
                ctx.upload_collision(t["Kr"][None], t["Ks"][None], t["rho"][None], None, t["idd"], t["ids"], t["sg"])
                ctx.set_state_uniform_phonons(state, t["nph"])
                ctx.collide(0.05)
                outs.append(ctx.get_state(want_phonons=False)[0])
        finally:
            del os.environ["QPB_NO_UNIFORM"]
    helpers.assert_close(outs[0].T, outs[1].T, "tensor-core GEMM vs structured kernel, all cells", rtol=1e-11)
    pick = np.sort(rng.choice(ncell, size=nsample, replace=False))
    so = state[:, pick].copy()
    po = t["nph"][:, None] * np.ones((1, nsample))
    O.collide(so, po, t["Kr"], t["Ks"], t["rho"], t["idd"], t["ids"], t["sg"], t["dE"], 0.05, recomb=True, scat=True,
              update_phonons=False, chunk=8)
    helpers.assert_close(outs[0][:, pick].T, so.T, "tensor-core GEMM vs oracle on sampled cells")


@pytest.mark.parametrize("shape", [(256, 256), (128, 1024)], ids=["whole_lines_exact_fit", "segmented_rows"])
def test_sweep_scheduling_switches_do_not_change_the_result(shape, monkeypatch):
    """The L1 prefetch of the factor tables, programmatic dependent launch, the reversed bin walk of the y sweep and
    the exact-fit tile variants only reorder loads and launches: a masked Crank-Nicolson solve must come out bit for
    bit the same with each of them switched off (the switches are read at launch time)."""
    ny, nx = shape
    ne = 12
    monkeypatch.setenv("QPB_NO_RESIDENT", "1")   # the switches belong to the launched sweeps
    mask = cases.meander_mask(ny, nx, pad=8, slot=4, pitch=16, gap_len=32)
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    n = int(mask.sum())
    t = _tables(ne, 5.0)
    u0 = np.stack([cases.gaussian_field(mask, cx=0.3 + 0.03 * k, cy=0.5, sigma=0.07)[mask] for k in range(ne)])

    def solve():
        with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=t["dE"]) as ctx:
            ctx.upload_geometry(mask, bcx, bcy, src)
            ctx.upload_diffusion(t["D"])
            ctx.prepare_diffusion(0, 0.5)
            ctx.set_state(u0)
            ctx.advance(3, 0.5)
            out, _ = ctx.get_state(want_phonons=False)
            info = ctx.diag()
        assert info["sweep_path"] in (2, 3), info
        return out, info["sweeps"]

    want, sweeps = solve()
    assert np.all(np.isfinite(want)) and sweeps > 6
    for switch in ("QPB_PIPE_PREFETCH", "QPB_PIPE_PDL", "QPB_PIPE_REV"):
        monkeypatch.setenv(switch, "0")
        got, sw = solve()
        monkeypatch.delenv(switch)
        assert sw == sweeps, switch
        assert np.array_equal(got, want), switch
    monkeypatch.setenv("QPB_PIPE_NOFULL", "1")
    got, sw = solve()
    assert sw == sweeps and np.array_equal(got, want)
