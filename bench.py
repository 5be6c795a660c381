#!/usr/bin/env python
"""Benchmark of the qpsim time-stepping hot path on B200 (contract: the task brief / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                        the reference algorithm on the host cores (oracle port)

Metric (BASELINE.json): cell * energy-bin updates per second per time step = N_cells * NE * steps / time.

Default workload: BASELINE configs[2], the configuration every north-star target is quoted on - 2048 x 2048 full grid
x 256 bins, CN diffusion + scattering + recombination with dynamic phonons.  It fits one
B200 (N = 1) and is cut across the ranks at N > 1 (strong scaling): diffusion sharded by energy bin, collisions by
cell, the layout exchange fused into the collision kernel over NVLink peer memory.  At N = 1 the same JSON line
carries a second block "c2": BASELINE configs[1] (256 x 256 MKID-like meander x 128 bins, the single-B200 config).
`--workload c2` runs the round-1 line instead (C2, weak scaling over --gpus).

Every line holds: value (device-resident, CUDA events on the stream the kernels run on), e2e (one
run_2d_crank_nicolson call with host buffers: setup, uploads, steps, downloads inside the timed region; bytes counted
at the copy calls), roofline / roofline_sweeps / roofline_collision (per-kernel event times of a short serialised
pass), time_shares, a sampled parity block (collision update of seeded cells against the oracle, CN equations of two
bins), clocks and - on rank 0 at N = 1 - cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cell*energy-bin updates/s per timestep"
UNIT = "updates/s"

_json_fd = None


def protect_stdout():
    """Libraries (NCCL's version banner, torchrun notices) print to stdout; the contract is ONE JSON line there.
    Everything written to fd 1 from here on goes to stderr; emit() writes the result to the real stdout."""
    global _json_fd
    if _json_fd is None:
        sys.stdout.flush()
        _json_fd = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _json_fd is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_json_fd, data)


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
def c2_workload(tile_y: int = 1, tile_x: int = 1, ny: int = 256, nx: int = 256, ne: int = 128):
    """BASELINE configs[1]; tile_y x tile_x copies of the 256 x 256 mask side by side for the weak-scaling runs
    (the copies touch, so the tiled mask is one connected domain with longer rows and columns)."""
    import cases

    base = cases.meander_mask(ny, nx, pad=8, slot=4, pitch=16, gap_len=32)
    mask = np.tile(base, (tile_y, tile_x))
    if tile_x > 1:   # connect neighbouring copies through their padding columns at mid height of every copy
        for ty in range(tile_y):
            r = ty * ny + ny // 2
            mask[r - 4:r + 4, :] |= True
            mask[r - 4:r + 4, :8] = False
            mask[r - 4:r + 4, -8:] = False
    if tile_y > 1:
        for tx in range(tile_x):
            c = tx * nx + nx // 2
            mask[:, c - 4:c + 4] |= True
            mask[:8, c - 4:c + 4] = False
            mask[-8:, c - 4:c + 4] = False
    field = cases.gaussian_field(mask, cx=0.4 / tile_x, cy=0.5 / tile_y, sigma=0.05, base=1e-4, amp=2e-4)
    return dict(
        name=f"C2 meander {ny * tile_y}x{nx * tile_x} x {ne} bins", mask=mask, bc="short_absorbing", initial_field=field,
        diffusion_coefficient=cases.D0, dt=0.5, dx=1.0, energy_gap=cases.GAP, energy_min_factor=1.0,
        energy_max_factor=5.0, num_energy_bins=ne, dynes_gamma=cases.GAMMA, tau_0=cases.TAU, T_c=cases.TC,
        bath_temperature=cases.TBATH, pulse_rate=3e-8, pulse_start=0.0, pulse_duration=5.0,
    )


def c2_case(steps: int = 2):
    """c2_workload() in the vocabulary of tests/cases.py (what the reference / oracle / drop-in runners take): the
    full-size parity fixture tests/golden/c2_full_256x256x128.npz is this case run by the unmodified reference."""
    w = c2_workload()
    return dict(
        name="c2_full_256x256x128", mask=w["mask"], bc=w["bc"], initial_field=w["initial_field"],
        diffusion_coefficient=w["diffusion_coefficient"], dt=w["dt"], total_time=w["dt"] * steps, dx=w["dx"],
        store_every=steps, energy_gap=w["energy_gap"], energy_min_factor=w["energy_min_factor"],
        energy_max_factor=w["energy_max_factor"], num_energy_bins=w["num_energy_bins"], weights=None,
        enable_diffusion=True, enable_recombination=True, enable_scattering=True, dynes_gamma=w["dynes_gamma"],
        tau_0=w["tau_0"], T_c=w["T_c"], bath_temperature=w["bath_temperature"],
        generation=dict(mode="pulse", pulse_rate=w["pulse_rate"], pulse_start=w["pulse_start"],
                        pulse_duration=w["pulse_duration"]),
    )


def c3_workload():
    """BASELINE configs[2]: 2048 x 2048 full mask, reflective walls, 256 bins on [gap, 3 gap],
    dt = 0.2 ns, seeded lognormal field x thermal weights, dynamic phonons.  Strong scaling: the grid is fixed and cut
    across the ranks.  QPB_C3_SIDE / QPB_C3_BINS shrink it for smoke runs of this script."""
    import cases

    ny = nx = int(os.environ.get("QPB_C3_SIDE", "2048"))
    ne = int(os.environ.get("QPB_C3_BINS", "256"))
    mask = np.ones((ny, nx), dtype=bool)
    field = cases.lognormal_field(mask, seed=20260103, scale=1e-4)
    return dict(
        name=f"C3 full grid {ny}x{nx} x {ne} bins", mask=mask, bc="reflective", initial_field=field,
        diffusion_coefficient=cases.D0, dt=0.2, dx=1.0, energy_gap=cases.GAP, energy_min_factor=1.0,
        energy_max_factor=3.0, num_energy_bins=ne, dynes_gamma=cases.GAMMA,
        tau_0=cases.TAU, T_c=cases.TC, bath_temperature=cases.TBATH, pulse_rate=None, weights="thermal",
    )


def thermal_weights(w, E, dE, rho):
    # normalised thermal quasiparticle weights rho(E) f(E, 0.3 K) (solver.py:429-460), k_B in ueV/K (solver.py:347)
    f = 1.0 / (np.exp(np.minimum(E / (86.17333262145 * 0.3), 500.0)) + 1.0)
    return rho * f / (np.sum(rho * f) * dE)


def build_tables(w, Q=None, cells=None, want_state=True):
    """Host-side setup shared by the device-resident run and the CPU baseline.  cells = (c0, c1): only that slice of
    the state is built and the phonon state stays in its per-bin form (large grids)."""
    if Q is None:
        import qpsim_b200 as Q
    mask = w["mask"]
    n = int(mask.sum())
    E, dE = Q.build_energy_grid(w["energy_gap"], w["energy_min_factor"], w["energy_max_factor"], w["num_energy_bins"])
    rho = Q.density_of_states(E, w["energy_gap"], w["dynes_gamma"])
    Kr = Q.recombination_kernel_base(E, w["energy_gap"], w["tau_0"], w["T_c"])
    Ks = Q.scattering_kernel_base(E, w["energy_gap"], w["tau_0"], w["T_c"])
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    nph = Q.thermal_phonon_occupation(om, w["bath_temperature"])
    wts = thermal_weights(w, E, dE, rho) if w.get("weights") == "thermal" else rho / (np.sum(rho) * dE)
    spatial = w["initial_field"][mask]
    if cells is not None:
        spatial = spatial[cells[0]:cells[1]]
    state = wts[:, None] * spatial[None, :] if want_state else None
    phon = None if (cells is not None or not want_state) else nph[:, None] * np.ones((1, n))
    D = w["diffusion_coefficient"] * np.sqrt(np.maximum(0.0, 1.0 - (w["energy_gap"] / E) ** 2))
    return dict(E=E, dE=dE, rho=rho, Kr=Kr, Ks=Ks, omega=om, idx_diff=idd, idx_sum=ids, sign=sg, state=state,
                phonons=phon, phonon_bins=nph, D=D, n=n, weights=wts)


def solver_kwargs(w, Q, steps, edges=None, bcs=None):
    """Arguments of the public entry point for K steps of workload w."""
    import cases

    mask = w["mask"]
    if edges is None:
        edges = Q.extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
    gen = None
    if w.get("pulse_rate") is not None:
        gen = Q.ExternalGenerationSpec(mode="pulse", pulse_rate=w["pulse_rate"], pulse_start=w["pulse_start"],
                                       pulse_duration=w["pulse_duration"])
    kw = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=w["initial_field"],
              diffusion_coefficient=w["diffusion_coefficient"], dt=w["dt"], total_time=w["dt"] * steps, dx=w["dx"],
              store_every=steps, energy_gap=w["energy_gap"], energy_min_factor=w["energy_min_factor"],
              energy_max_factor=w["energy_max_factor"], num_energy_bins=w["num_energy_bins"], enable_diffusion=True,
              enable_recombination=True, enable_scattering=True, dynes_gamma=w["dynes_gamma"], tau_0=w["tau_0"],
              T_c=w["T_c"], bath_temperature=w["bath_temperature"], external_generation=gen)
    if w.get("weights") == "thermal":
        E, dE = Q.build_energy_grid(w["energy_gap"], w["energy_min_factor"], w["energy_max_factor"],
                                    w["num_energy_bins"])
        kw["energy_weights"] = thermal_weights(w, E, dE, Q.density_of_states(E, w["energy_gap"], w["dynes_gamma"]))
    return kw


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        load = sorted(sm)[len(sm) // 2:] if sm else []   # upper half ~ samples under load
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference algorithm, bounded FIXED sample, warm worker pool)
# ----------------------------------------------------------------------------------------------------------
CPU_CELLS = 4096          # fixed collision sample: the same cells in every arm and every run
CPU_DIFF_SIDE = 512       # SuperLU sample grid when the workload's own grid is too large to factor in a bench run


def _cpu_warm(_):
    from oracle import qp_oracle  # noqa: F401  (import cost of scipy.sparse stays out of the timed region)
    return os.getpid()


def _cpu_collide_chunk(args):
    from oracle import qp_oracle as O

    st, ph, t, dt = args
    t0 = time.perf_counter()
    for _ in range(2):   # the two half steps of a time step (solver.py:1469-1472)
        O.collide(st, ph, t["Kr"], t["Ks"], t["rho"], t["idx_diff"], t["idx_sum"], t["sign"], t["dE"], 0.5 * dt,
                  recomb=True, scat=True, chunk=64)
    return time.perf_counter() - t0


def _cpu_diffuse_bins(args):
    from oracle import qp_oracle as O
    import qpsim_b200 as Q
    import cases

    mask, bc, dx, D, dt, field, reps = args
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, bc, Q.BoundaryCondition)
    op = O.DiffusionCN(mask, edges, bcs, dx, D[:, None] * np.ones((1, int(mask.sum()))), dt, False)  # setup, untimed
    st = np.repeat(field[None, :], D.size, axis=0)
    op.step(st)
    t0 = time.perf_counter()
    for _ in range(reps):
        op.step(st)
    return (time.perf_counter() - t0) / reps


def cpu_baseline(w, tabs=None):
    """The reference algorithm (oracle port: batched-over-cells collision step twice per time step, SuperLU
    Crank-Nicolson solve per bin; factorisation is setup and untimed, as in the reference) on a bounded fixed sample,
    scaled linearly to the workload - an EXTRAPOLATION, labelled so.  Reported for all host cores (independent cells /
    bins dealt to a warm process pool) and for one core (the stock reference runs on one)."""
    from concurrent.futures import ProcessPoolExecutor

    import qpsim_b200 as Q

    if tabs is None:
        tabs = build_tables(w, Q, want_state=False)
    cores = os.cpu_count() or 1
    mask = w["mask"]
    n, ne = tabs["n"], w["num_energy_bins"]
    rng = np.random.default_rng(20260102)
    pick = np.sort(rng.choice(n, size=min(CPU_CELLS, n), replace=False))
    spatial = w["initial_field"][mask][pick]
    st = tabs["weights"][:, None] * spatial[None, :]
    ph = tabs["phonon_bins"][:, None] * np.ones((1, pick.size))
    small = {k: tabs[k] for k in ("Kr", "Ks", "rho", "idx_diff", "idx_sum", "sign", "dE")}
    per = -(-pick.size // cores)
    jobs = [(st[:, k:k + per].copy(), ph[:, k:k + per].copy(), small, w["dt"]) for k in range(0, pick.size, per)]
    # diffusion sample: the workload's own geometry when a factorisation is cheap, else a CPU_DIFF_SIDE^2 corner of it
    ny, nx = mask.shape
    if ny * nx > 1024 * 1024:
        sub = mask[:CPU_DIFF_SIDE, :CPU_DIFF_SIDE].copy()
        sub_field = w["initial_field"][:CPU_DIFF_SIDE, :CPU_DIFF_SIDE][sub]
        diff_note = (f"{CPU_DIFF_SIDE}^2 corner of the grid, scaled by cells (a LOWER bound of the CPU time: "
                     "SuperLU fill grows faster than the cell count)")
    else:
        sub, sub_field, diff_note = mask, w["initial_field"][mask], "the workload's own mask"
    nbins = min(ne, cores)
    bsel = np.linspace(0, ne - 1, nbins).astype(int)
    djobs = [(sub, w["bc"], w["dx"], tabs["D"][[b]], w["dt"], sub_field, 3) for b in bsel]
    with ProcessPoolExecutor(max_workers=cores) as ex:
        list(ex.map(_cpu_warm, range(4 * cores)))            # workers started, scipy imported
        t_coll_wall, t_coll_1core = np.inf, np.inf
        for _ in range(3):                                   # best of three: the host cores of a GPU box are shared
            t0 = time.perf_counter()
            per_job = list(ex.map(_cpu_collide_chunk, jobs))
            t_coll_wall = min(t_coll_wall, time.perf_counter() - t0)   # wall time with every core busy
            t_coll_1core = min(t_coll_1core, float(np.sum(per_job)))   # CPU seconds of the sample
        t0 = time.perf_counter()
        per_bin = list(ex.map(_cpu_diffuse_bins, djobs))
        t_diff_wall = time.perf_counter() - t0               # factorisations included (untimed inside per_bin)
    scale = n / pick.size
    t_bin = float(np.mean(per_bin)) * (n / int(sub.sum()))   # seconds per bin and step at full size
    step_all = t_coll_wall * scale + t_bin * ne / min(cores, ne)
    step_one = t_coll_1core * scale + t_bin * ne
    return {
        "value": n * ne / step_all, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
        "single_core_value": n * ne / step_one,
        "sample": (f"oracle port of qpsim.solver: 2 collision half steps on a fixed {pick.size}-cell sample of {n} "
                   f"({t_coll_wall:.2f} s wall on {cores} warm worker processes, {t_coll_1core:.1f} CPU s) + SuperLU CN "
                   f"solve of {nbins} of {ne} bins on {diff_note} ({np.mean(per_bin) * 1e3:.1f} ms per bin and step at "
                   "sample size, factorisation untimed); scaled linearly in cells and bins"),
        "est_ms_per_step": step_all * 1e3, "est_ms_per_step_single_core": step_one * 1e3,
        # what was actually timed: one pass over the collision sample (best of three) and the CN solves of the bin sample
        "sample_step_s": t_coll_wall + float(np.sum(per_bin)) / min(cores, nbins),
        "sample_wall_s": 3 * t_coll_wall + t_diff_wall,
    }


# ----------------------------------------------------------------------------------------------------------
# sampled parity inside the bench run
# ----------------------------------------------------------------------------------------------------------
def cn_equations_error(mask, bcx, bcy, src, a, dtD, u0, u1):
    """How well u1 satisfies the reference's Crank-Nicolson system (I - aL) u1 = (I + aL) u0 + dt D s on the masked
    grid (solver.py:221-232, 1440-1452), evaluated on the host: componentwise max |A u1 - rhs| / (|A||u1| + |rhs|)
    and the max-norm form.  ||A^-1||_inf <= 1, so the max-norm residual bounds the error against the direct solve."""
    m = np.asarray(mask, dtype=bool)

    def G(u, absolute=False):
        out = np.abs(bcx + bcy) * np.abs(u) if absolute else (bcx + bcy) * u
        for axis, sh in ((0, 1), (0, -1), (1, 1), (1, -1)):
            nb = np.roll(u, sh, axis=axis)
            link = m & np.roll(m, sh, axis=axis)
            if axis == 0:
                link[0 if sh == 1 else -1, :] = False
            else:
                link[:, 0 if sh == 1 else -1] = False
            out = out + np.where(link, (np.abs(u) + np.abs(nb)) if absolute else (u - nb), 0.0)
        return np.where(m, out, 0.0)

    lhs = u1 + a * G(u1)
    rhs = np.where(m, u0 - a * G(u0) + dtD * src, 0.0)
    w = np.abs(u1) + a * G(u1, absolute=True) + np.abs(rhs)
    r = np.abs(lhs - rhs)
    return {"componentwise": float(np.max(np.where(m & (w > 0), r / np.where(w > 0, w, 1.0), 0.0))),
            "max_norm": float(np.max(r) / max(np.max(np.abs(u1)), 1e-300))}


def collision_sample_error(tabs, dt, n0, p0, n1, p1):
    """One collision half step of the sampled cells through the oracle against what the device produced."""
    from oracle import qp_oracle as O

    rn, rp = n0.copy(), p0.copy()
    O.collide(rn, rp, tabs["Kr"], tabs["Ks"], tabs["rho"], tabs["idx_diff"], tabs["idx_sum"], tabs["sign"], tabs["dE"],
              0.5 * dt, recomb=True, scat=True, chunk=64)
    big = np.abs(rn) > 1e-300
    en = float(np.max(np.where(big, np.abs(n1 - rn) / np.where(big, np.abs(rn), 1.0), 0.0)))
    bigp = np.abs(rp) > 1e-300
    ep = float(np.max(np.where(bigp, np.abs(p1 - rp) / np.where(bigp, np.abs(rp), 1.0), 0.0)))
    return {"cells": int(n0.shape[1]), "n_max_rel": en, "n_ph_max_rel": ep}


# ----------------------------------------------------------------------------------------------------------
# peaks
# ----------------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def load_profile_traffic(tag, n_coll_cells=None, nbins=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures under
    profiles/ (documentation of an earlier run of the same kernels, not a measurement of this one).  The C3 kernels were
    captured on a slice of the workload (collision: 37 888 cells of the 256-bin grid; spectral passes: 16 bins of the
    2048^2 grid): their bytes are scaled to the cells / bins of this rank's launches, and the line says so."""
    out = {}
    for name in ("r2_ncu_kernels.json", "r1_ncu_kernels.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            k = json.load(f)
        if tag == "c3":
            sw = [v["dram_bytes"] / v["slice"]["bins"] * (nbins or 0) for n, v in k.items()
                  if n.split("[")[-1] in ("c3_k_dct_forward]", "c3_k_dct_inverse]", "c3_k_thomas_fused]") and "slice" in v]
            co = [v["dram_bytes"] / v["slice"]["cells"] * (n_coll_cells or 0) for n, v in k.items()
                  if n.endswith("[c3_k_collide_struct]") and "slice" in v]
            if len(sw) == 3 and nbins and "sweep" not in out:
                out["sweep"] = float(np.mean(sw))
                out["sweep_source"] = ("committed ncu captures of the three spectral passes on 16 bins of this grid "
                                       "(profiles/r2_ncu_kernels.json), scaled to the bins of this rank; not this run")
            if co and n_coll_cells and "collide" not in out:
                out["collide"] = float(co[0])
                out["collide_source"] = ("committed ncu capture of this kernel on 37 888 cells of this energy grid "
                                         "(profiles/r2_ncu_kernels.json), scaled to the cells of this rank; not this run")
            continue
        sw = [v["dram_bytes"] for n, v in k.items() if n.endswith("[c2_k_pr_resident]")]
        if not sw:
            sw = [v["dram_bytes"] for n, v in k.items() if n.endswith("[k_sweep_x_pipe]") or n.endswith("[k_sweep_y_pipe]")]
        co = [v["dram_bytes"] for n, v in k.items() if n.endswith("k_collide_struct]") and "[c3_" not in n]
        if sw and "sweep" not in out:
            out["sweep"] = float(np.mean(sw))
        if co and "collide" not in out:
            out["collide"] = float(np.mean(co))
    return out


def rooflines(n_sweep_cells, ncd, ne, n_coll_cells, tx, ty, nsweep_launches, bin_sweeps, tc, ncoll, peaks, peak_src,
              fp64_peak, traffic, sweep_path=2):
    """HBM roofline of the sweeps (algorithmic 16 B per mask cell, bin and directional sweep) and FP64 roofline of the
    collision kernel (21 NE^2 flop per cell and call with dynamic phonons), from per-kernel event times."""
    sweep_ms = tx + ty
    rs = {"bound": "hbm", "achieved": 16.0 * n_sweep_cells * bin_sweeps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0,
          "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": traffic.get("sweep"),
          "traffic_source": traffic.get("sweep_source", "committed ncu capture (profiles/), not this run")
          if traffic.get("sweep") else None,
          "kernel": ("k_dct_forward + k_thomas_frozen + k_dct_inverse (direct spectral solve: three passes per bin)"
                     if sweep_path == 4 else
                     "k_pr_resident (bin-resident cluster solve: right-hand side, all line sweeps of the iteration and the "
                     "stop test of a bin in the shared memory of one 8-CTA cluster, one launch per solve)"
                     if sweep_path == 5 else "k_sweep_x_pipe + k_sweep_y_pipe (tridiagonal line sweeps)"),
          "launches": int(nsweep_launches),
          "ms_per_launch": sweep_ms / max(1, nsweep_launches), "peak_source": peak_src}
    rs["frac"] = rs["achieved"] / rs["peak"]
    # what the kernels move per dense cell, bin and pass: sweeps 24 B (x: u, b in, u* out; y: u*, u in, u out); the
    # spectral passes 16 B each (read and write the bin once)
    if sweep_path == 5:
        # the resident solve reads u and writes u and b once per bin and solve, whatever the number of line sweeps
        moved = 24.0 * ncd * ne * nsweep_launches
        rs["implementation"] = {"bytes_per_dense_cell_bin_solve": 24, "dense_cells": int(ncd),
                                "achieved": moved / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0, "unit": "GB/s",
                                "note": "the line sweeps run out of shared memory: the algorithmic 16 B per cell, bin and "
                                        "sweep of the contract are not HBM traffic any more; 'achieved' above is the "
                                        "rate at which the launched sweeps would have had to stream them"}
    else:
        impl = 16 if sweep_path == 4 else 24
        rs["implementation"] = {"bytes_per_dense_cell_bin_sweep": impl, "dense_cells": int(ncd),
                                "achieved": impl * ncd * bin_sweeps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0,
                                "unit": "GB/s"}
    rs["implementation"]["frac"] = rs["implementation"]["achieved"] / rs["peak"]
    flops = 21.0 * ne * ne * n_coll_cells * ncoll
    rc = {"bound": "fp64", "achieved": flops / (tc * 1e-3) / 1e12 if tc > 0 else 0.0, "peak": fp64_peak,
          "unit": "TFLOP/s", "traffic": traffic.get("collide"),
          "traffic_source": traffic.get("collide_source", "committed ncu capture (profiles/), not this run")
          if traffic.get("collide") else None,
          "kernel": "k_collide_struct", "launches": int(ncoll), "ms_per_launch": tc / max(1, ncoll),
          "peak_source": "measured in this run: DFMA loop on all SMs (qpb_measure_fp64)"}
    rc["frac"] = rc["achieved"] / rc["peak"] if fp64_peak else None
    return rs, rc


# ----------------------------------------------------------------------------------------------------------
# C2 on one GPU: single context through qpb_advance (the block "c2" of the default line, or --workload c2 at N = 1)
# ----------------------------------------------------------------------------------------------------------
def measure_c2_single(K, W, dev, with_cpu=True):
    import qpsim_b200 as Q
    import cases
    from qpsim_b200 import capi

    w = c2_workload()
    tabs = build_tables(w, Q)
    mask = w["mask"]
    ny, nx = mask.shape
    n, ne, nw = tabs["n"], w["num_energy_bins"], tabs["omega"].size
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
    flags = capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI
    gen = dict(gen_mode=capi.GEN_PULSE, rate=w["pulse_rate"], pulse_start=w["pulse_start"],
               pulse_duration=w["pulse_duration"])
    peaks, peak_src = load_peaks()
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=nw, ncell=n, flags=flags, dx=w["dx"], dE=tabs["dE"], device=dev) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src)
        ctx.upload_diffusion(tabs["D"])
        ctx.prepare_diffusion(0, w["dt"])
        ctx.upload_collision(tabs["Kr"][None], tabs["Ks"][None], tabs["rho"][None], None, tabs["idx_diff"],
                             tabs["idx_sum"], tabs["sign"])
        ctx.set_state(tabs["state"], tabs["phonons"])
        # ---- device-resident timing: inputs already in HBM, CUDA events on the library's stream ----
        ctx.advance(W, w["dt"], t_start=0.0, want_pauli=True, **gen)
        l0 = ctx.diag()["kernel_launches"]
        sampler = ClockSampler(dev)
        sampler.start()
        ctx.advance(K, w["dt"], t_start=W * w["dt"], want_pauli=True, **gen)
        d = ctx.diag()
        clocks = sampler.stop()
        ms_total = d["last_advance_ms"]
        launches = d["kernel_launches"] - l0
        # ---- per-kernel device times (serialised by events: the programmatic-dependent-launch overlap of consecutive
        # sweeps is lost here, so the parts sum to slightly more than the step) ----
        d0 = ctx.diag()
        ctx.enable_timers(True)
        ctx.reset_timers()
        ks = max(2, min(K, 4))
        ctx.advance(ks, w["dt"], t_start=(W + K) * w["dt"], want_pauli=True, **gen)
        ctx.enable_timers(False)
        d1 = ctx.diag()
        tx, nxl = ctx.timer(0)
        ty, nyl = ctx.timer(1)
        tc, ncl = ctx.timer(2)
        bin_sweeps = d1["bin_sweeps"] - d0["bin_sweeps"]
        sweeps_per_step = (d1["sweeps"] - d0["sweeps"]) / ks
        integ = ctx.get_integrated()
        assert np.all(np.isfinite(integ))
    value = n * ne * K / (ms_total * 1e-3)
    fp64_peak = capi.measure_fp64_tflops(dev)
    copy_gbs = capi.measure_copy_gbs(dev, 1 << 30)
    roof_sweep, roof_coll = rooflines(n, ny * nx, ne, n, tx, ty, nxl + nyl, bin_sweeps, tc, ncl, peaks, peak_src,
                                      fp64_peak, load_profile_traffic("c2"), sweep_path=d1["sweep_path"])
    dominant = roof_coll if tc >= tx + ty else roof_sweep
    shares = {"collision_ms_per_step": tc / ks, "sweeps_ms_per_step": (tx + ty) / ks,
              "other_ms_per_step": (ms_total / K) - (tc + tx + ty) / ks,
              "note": "kernel times from a serialised pass (an event pair per launch); the timed step overlaps the "
                      "prologue of each sweep with its predecessor's tail, so 'other' can come out slightly negative"}
    # ---- end to end through the public API: host buffers in, host results out ----
    kw = solver_kwargs(w, Q, K, edges, bcs)
    kw["device"] = dev
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):   # untimed warm-up calls (module load, allocator cache, host page faults)
            Q.run_2d_crank_nicolson(**{**kw, "total_time": w["dt"] * 3, "store_every": 3})
        capi.transfer_stats.update(h2d=0, d2h=0)
        t0 = time.perf_counter()
        Q.run_2d_crank_nicolson(**kw)
        t_e2e = time.perf_counter() - t0
    e2e = {"value": n * ne * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": capi.transfer_stats["h2d"] / K,
           "d2h_bytes_per_step": capi.transfer_stats["d2h"] / K, "seconds": t_e2e,
           "note": "one run_2d_crank_nicolson call (context creation, geometry compile, uploads, "
                   f"{K} steps, download of the t=0 and final energy frames), host numpy buffers in and out; bytes "
                   "counted at the copy calls of the binding"}
    block = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "cells": n, "energy_bins": ne, "phonon_bins": int(nw), "dt_ns": w["dt"],
                   "sweep_path": d1["sweep_path"],
                   "processes": "masked CN diffusion (exact, PR-sweep iteration, componentwise stop test) + "
                                "scattering + recombination, dynamic phonons, pulse generation, Pauli check every step",
                   "l2": "state (64 MiB) + phonons (190 MiB) + work arrays exceed the 126 MB L2",
                   "sweeps_per_step": sweeps_per_step,
                   "parity": "tests/test_gpu_suite.py::test_c2_full_size_two_steps_match_reference (this workload, "
                             "2 steps of the unmodified reference, element-wise 1e-9)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": dominant, "roofline_sweeps": roof_sweep, "roofline_collision": roof_coll,
        "time_shares": shares,
        "peaks": {"hbm_gbs": peaks["hbm_gbs"], "copy_gbs_this_run": copy_gbs, "fp64_tflops_this_run": fp64_peak},
    }
    if with_cpu:
        block["cpu_baseline"] = cpu_baseline(w, tabs)
    return block


# ----------------------------------------------------------------------------------------------------------
# sharded driver: C3 at every N (strong scaling), or C2 tiled (weak scaling, --workload c2 at N > 1)
# ----------------------------------------------------------------------------------------------------------
def run_sharded(args, strong: bool):
    import warnings

    import torch
    import torch.distributed as dist

    import cases
    import qpsim_b200 as Q
    from qpsim_b200 import capi, multigpu

    rank, world, local = multigpu.init_process_group("nccl")
    if strong:
        w = c3_workload()
    else:
        ty_, tx_ = multigpu.weak_tiling(world)
        w = c2_workload(tile_y=ty_, tile_x=tx_)
    mask = w["mask"]
    ny, nx = mask.shape
    n, ne = int(mask.sum()), w["num_energy_bins"]
    plan = multigpu.ShardPlan(ne, n, world, rank, interleave=True)
    c0, c1 = plan.cells()
    nloc = c1 - c0
    tabs = build_tables(w, Q, cells=(c0, c1))
    nw = int(tabs["omega"].size)
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
    prob = multigpu.ShardedProblem(
        mask=mask, bcx=bcx, bcy=bcy, src=src, dx=w["dx"], dE=tabs["dE"], D=tabs["D"], variable_D=False,
        rho=tabs["rho"][None], Kr=tabs["Kr"][None], Ks=tabs["Ks"][None], gap_id=None, idx_diff=tabs["idx_diff"],
        idx_sum=tabs["idx_sum"], sign=tabs["sign"], nw=nw, state=None, phonons=None, state_local=tabs["state"],
        phonon_bins=tabs["phonon_bins"])
    K, W = args.steps, args.warmup
    dt = w["dt"]
    peaks, peak_src = load_peaks()
    stages = multigpu.DeviceStages(plan, prob, local, dt)
    fused = stages.enable_fused_exchange(prob)
    gen_on = w.get("pulse_rate") is not None

    def rate_at(t):
        if not gen_on:
            return None
        return w["pulse_rate"] if w["pulse_start"] <= t < w["pulse_start"] + w["pulse_duration"] else None

    with torch.cuda.stream(stages.stream):
        stepper = multigpu.ShardedStepper(plan, stages, diffusion=True, collisions=True)
        t = 0.0
        for _ in range(W):
            stepper.step(dt, 0, rate_at(t), want_pauli=True)
            t += dt
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        l0 = stages.launches()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stages.stream)
        for k in range(K):
            stepper.step(dt, 0, rate_at(t), pauli_slot=k)   # occupancy record per step on the device
            t += dt
        e1.record(stages.stream)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if rank == 0 else None
        launches = stages.launches() - l0
        merged = stepper.merge_pauli(stages.pauli_fetch(K))
        ms_total = float(ms.item())
        # ---- per-kernel event times on this rank (serialised pass of ks steps), max over ranks ----
        ks = 2
        dd0 = stages.ctx_d.diag()
        for c in (stages.ctx_c, stages.ctx_d):
            c.enable_timers(True)
            c.reset_timers()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        es0.record(stages.stream)
        for k in range(ks):
            stepper.step(dt, 0, rate_at(t), pauli_slot=k)
            t += dt
        es1.record(stages.stream)
        torch.cuda.synchronize()
        for c in (stages.ctx_c, stages.ctx_d):
            c.enable_timers(False)
        dd1 = stages.ctx_d.diag()
        tx, nxl = stages.ctx_d.timer(0)
        ty, nyl = stages.ctx_d.timer(1)
        tc, ncl = stages.ctx_c.timer(2)
        bin_sweeps = dd1["bin_sweeps"] - dd0["bin_sweeps"]
        sweeps_per_step = (dd1["sweeps"] - dd0["sweeps"]) / ks
        ser_ms = es0.elapsed_time(es1)
        parts = torch.tensor([tc / ks, (tx + ty) / ks, ser_ms / ks], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(parts, op=dist.ReduceOp.MAX)
        # ---- sampled parity on this rank's data (rank 0 reports) ----
        parity = None
        stages.load_state(prob)
        if rank == 0:
            dev = stages.coll_state.device
            rng = np.random.default_rng(20260104)
            pick = np.sort(rng.choice(nloc, size=min(1024, nloc), replace=False))
            pick_t = torch.as_tensor(pick, device=dev)
            pptr, _ = stages.ctx_c.device_ptr(1)
            ph_view = torch.as_tensor(multigpu._DevArray(pptr, (nw, nloc)), device=dev)
            n0 = stages.coll_state.index_select(1, pick_t).cpu().numpy()
            p0 = ph_view.index_select(1, pick_t).cpu().numpy()
            stages.collide(0.5 * dt)
            torch.cuda.synchronize()
            n1 = stages.coll_state.index_select(1, pick_t).cpu().numpy()
            p1 = ph_view.index_select(1, pick_t).cpu().numpy()
            parity = {"collision": collision_sample_error(tabs, dt, n0, p0, n1, p1)}
            # two bins of this rank (lowest / highest D) through one more CN solve of the field the run left there
            sptr, _ = stages.ctx_d.device_ptr(0)
            nb = plan.nbins()
            dview = torch.as_tensor(multigpu._DevArray(sptr, (nb, ny, nx)), device=dev)
            rows = sorted({0, nb - 1})
            u0 = [dview[r].cpu().numpy().copy() for r in rows]
        # every rank solves once more (its own bins: no exchange involved)
        stages.ctx_d.diffuse(0)
        torch.cuda.synchronize()
        if rank == 0:
            errs = []
            for r, before in zip(rows, u0):
                after = dview[r].cpu().numpy()
                D_i = float(tabs["D"][plan.bins()[r]])
                errs.append(cn_equations_error(mask, bcx, bcy, src, 0.5 * dt * D_i / w["dx"] ** 2, dt * D_i, before,
                                               after))
            parity["diffusion_cn_equations"] = {
                "bins": [int(plan.bins()[r]) for r in rows],
                "componentwise_residual": max(e["componentwise"] for e in errs),
                "max_norm_residual": max(e["max_norm"] for e in errs),
                "note": "host evaluation of (I - aL)u' = (I + aL)u + dt D s on the full grid; the max-norm residual "
                        "bounds the error against the reference's direct solve (||A^-1||_inf <= 1)"}
        dist.barrier()
    stages.close()
    # ---- end to end through the public API: host buffers in, host results out, setup included ----
    kw = solver_kwargs(w, Q, K, edges, bcs)
    kw.update(devices=list(range(world)) if world > 1 else [local], store_energy_frames=not strong and world == 1)
    capi.transfer_stats.update(h2d=0, d2h=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dist.barrier()
        t0 = time.perf_counter()
        Q.run_2d_crank_nicolson(**kw)
        dist.barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    xfer = torch.tensor([capi.transfer_stats["h2d"], capi.transfer_stats["d2h"]], dtype=torch.float64,
                        device=f"cuda:{local}")
    dist.all_reduce(xfer)
    fp64_peak = capi.measure_fp64_tflops(local) if rank == 0 else None
    line = None
    if rank == 0:
        coll_ms, sweep_ms, ser_step = (float(v) for v in parts.tolist())
        roof_sweep, roof_coll = rooflines(n, ny * nx, ne, nloc, tx, ty, nxl + nyl, bin_sweeps, tc, ncl, peaks,
                                          peak_src, fp64_peak,
                                          load_profile_traffic("c3", nloc, plan.nbins()) if strong
                                          else load_profile_traffic("c2"),
                                          sweep_path=dd1["sweep_path"])
        roof_coll["scope"] = f"rank 0: {nloc} of {n} cells, all bins"
        roof_sweep["scope"] = f"rank 0: {plan.nbins()} of {ne} bins, all cells"
        dominant = roof_coll if coll_ms >= sweep_ms else roof_sweep
        step_ms = ms_total / K
        shares = {"collision_ms_per_step": coll_ms, "sweeps_ms_per_step": sweep_ms,
                  "exchange_barriers_other_ms_per_step": ser_step - coll_ms - sweep_ms,
                  "serialised_step_ms": ser_step, "timed_step_ms": step_ms,
                  "note": "max over ranks of per-kernel event times from a serialised pass of 2 steps; "
                          "exchange_barriers_other = that pass's step minus its collision and sweep kernels: peer "
                          "stores/loads are inside the collision kernel, so this is the stream-ordered barriers, "
                          "right-hand side / Pauli / generation kernels and rank skew"}
        e2e_note = ("one collective run_2d_crank_nicolson(devices=[0..N-1]) call: per-rank context creation, geometry "
                    "compile, table and state uploads, steps, integrated frames and masses back on every rank"
                    if world > 1 else "one run_2d_crank_nicolson call: context creation, geometry compile, uploads, "
                    "steps, integrated frames and masses back")
        if not kw["store_energy_frames"]:
            e2e_note += (f"; the per-energy frames of the two stored times ({8 * ne * n / 2**30:.1f} GiB each at this "
                         "size) are not requested (store_energy_frames=False)")
        line = {
            "metric": METRIC, "value": n * ne * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "cells": n, "energy_bins": ne, "phonon_bins": nw, "dt_ns": dt,
                       "processes": "CN diffusion (exact, PR-sweep iteration, componentwise stop test) + scattering + "
                                    "recombination, dynamic phonons, Pauli check every step",
                       "parallelism": (f"bins/{world} (diffusion) <-> cells/{world} (collisions), "
                                       + ("one GPU: both layouts on the device" if world == 1
                                          else "exchange fused into the collision kernel over NVLink peer memory, "
                                          "3 stream-ordered barriers per step" if fused
                                          else "NCCL all-to-all x2 per step")),
                       "exchange_bytes_per_gpu": plan.exchange_bytes(), "sweeps_per_step": sweeps_per_step,
                       "l2": "per-GPU state + phonons + work arrays exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": n * ne * K / float(t_e2e.item()), "unit": UNIT, "seconds": float(t_e2e.item()),
                    "h2d_bytes_per_step": float(xfer[0].item()) / K, "d2h_bytes_per_step": float(xfer[1].item()) / K,
                    "breakdown_rank0_s": dict(Q.solver.last_run_info.get("seconds", {})) if world > 1 else None,
                    "note": e2e_note + "; bytes counted at the copy calls of the binding, summed over ranks"},
            "roofline": dominant, "roofline_sweeps": roof_sweep, "roofline_collision": roof_coll,
            "time_shares": shares, "parity": parity, "max_occupation": max(r[0] for r in merged),
            "peaks": {"hbm_gbs": peaks["hbm_gbs"], "fp64_tflops_this_run": fp64_peak},
        }
    dist.barrier()
    dist.destroy_process_group()
    return line, w, local


def run_ours(args):
    world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    strong = args.workload == "c3"
    if not strong and world == 1 and args.gpus <= 1:
        block = measure_c2_single(args.steps, args.warmup, int(os.environ.get("LOCAL_RANK", "0")))
        block.update(scaling="weak", vs_baseline=None)
        emit(block)
        return
    line, w, local = run_sharded(args, strong)
    if line is None:
        return
    if strong and world == 1:
        # same process, same box: the reference algorithm on the host cores and the single-B200 configuration
        line["cpu_baseline"] = cpu_baseline(w)
        line["c2"] = measure_c2_single(args.steps, args.warmup, local, with_cpu=True)
    emit(line)


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference itself is pure
    Python and cannot travel to the GPU box), on the configuration of our own arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c3":
        w = c3_workload()
        scaling = "strong"
    else:
        from qpsim_b200.multigpu import weak_tiling

        ty, tx = weak_tiling(max(1, args.gpus))
        w = c2_workload(tile_y=ty, tile_x=tx)
        scaling = "weak"
    reps = max(1, min(args.steps, 3))
    runs = [cpu_baseline(w) for _ in range(reps)]
    v = float(np.mean([r["value"] for r in runs]))
    last = runs[-1]
    n, ne = int(w["mask"].sum()), w["num_energy_bins"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup,
        # a "step" of this arm is one pass over the bounded sample (cpu_baseline.sample): ms_per_step is its measured
        # time; `value` is the sample's time per cell*bin update inverted, est_ms_per_step that time scaled to the
        # whole workload (linear in cells and bins) - an extrapolation, never run
        "ms_per_step": float(np.mean([r["sample_step_s"] for r in runs])) * 1e3, "steps_timed": reps,
        "extrapolated": True, "est_ms_per_step": n * ne / v * 1e3,
        "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "cells": n, "energy_bins": ne,
                   "step": "one pass over a bounded sample of the workload (see cpu_baseline.sample)",
                   # value = this / (ms_per_step / 1e3): the cell*bin updates one sampled pass stands for
                   "updates_per_sampled_step": v * float(np.mean([r["sample_step_s"] for r in runs]))},
        "cpu_baseline": {"kind": last["kind"], "cores": last["cores"], "sample": last["sample"], "value": v,
                         "unit": UNIT, "extrapolated": True, "single_core_value": last["single_core_value"],
                         "spread": [float(r["value"]) for r in runs]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"],
                    help="c3 (default): BASELINE configs[2], 2048^2 x 256 bins, cut across the GPUs (strong scaling; "
                         "at N = 1 the line also carries the C2 block); c2: configs[1], weak scaling over --gpus")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
