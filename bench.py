#!/usr/bin/env python
"""Benchmark of the qpsim time-stepping hot path on B200 (contract: see the task brief / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                        the reference algorithm on the host cores (oracle port)

Metric (BASELINE.json): cell * energy-bin updates per second per time step = N_cells * NE * steps / time.
Workload at N = 1: BASELINE configs[1] — 256 x 256 MKID-like meander mask x 128 energy bins, masked CN diffusion
+ scattering + recombination with dynamic phonons, pulse generation.  At N > 1 the mask is
tiled N times along x so the work per GPU is fixed (weak scaling): diffusion is sharded by energy bin, collisions by
cell, with an NCCL all-to-all between the two layouts.
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
# workload
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
    across the ranks (bench.py --workload c3; not the default line)."""
    import cases

    ny = nx = int(os.environ.get("QPB_C3_SIDE", "2048"))
    mask = np.ones((ny, nx), dtype=bool)
    field = cases.lognormal_field(mask, seed=20260103, scale=1e-4)
    return dict(
        name=f"C3 full grid {ny}x{nx} x 256 bins", mask=mask, bc="reflective", initial_field=field,
        diffusion_coefficient=cases.D0, dt=0.2, dx=1.0, energy_gap=cases.GAP, energy_min_factor=1.0,
        energy_max_factor=3.0, num_energy_bins=int(os.environ.get("QPB_C3_BINS", "256")), dynes_gamma=cases.GAMMA,
        tau_0=cases.TAU, T_c=cases.TC, bath_temperature=cases.TBATH, pulse_rate=None, weights="thermal",
    )


def build_tables(w, Q=None, cells=None):
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
    if w.get("weights") == "thermal":
        # normalised thermal quasiparticle weights rho(E) f(E, 0.3 K) (solver.py:429-460), k_B in ueV/K (solver.py:347)
        f = 1.0 / (np.exp(np.minimum(E / (86.17333262145 * 0.3), 500.0)) + 1.0)
        wts = rho * f / (np.sum(rho * f) * dE)
    else:
        wts = rho / (np.sum(rho) * dE)
    spatial = w["initial_field"][mask]
    if cells is not None:
        spatial = spatial[cells[0]:cells[1]]
    state = wts[:, None] * spatial[None, :]
    phon = None if cells is not None else nph[:, None] * np.ones((1, n))
    D = w["diffusion_coefficient"] * np.sqrt(np.maximum(0.0, 1.0 - (w["energy_gap"] / E) ** 2))
    return dict(E=E, dE=dE, rho=rho, Kr=Kr, Ks=Ks, omega=om, idx_diff=idd, idx_sum=ids, sign=sg, state=state,
                phonons=phon, phonon_bins=nph, D=D, n=n)


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
# CPU baseline (oracle port of the reference algorithm, bounded sample, all host cores)
# ----------------------------------------------------------------------------------------------------------
def _cpu_collide_chunk(args):
    from oracle import qp_oracle as O

    st, ph, t, dt = args
    t0 = time.perf_counter()
    O.collide(st, ph, t["Kr"], t["Ks"], t["rho"], t["idx_diff"], t["idx_sum"], t["sign"], t["dE"], 0.5 * dt,
              recomb=True, scat=True, chunk=64)
    O.collide(st, ph, t["Kr"], t["Ks"], t["rho"], t["idx_diff"], t["idx_sum"], t["sign"], t["dE"], 0.5 * dt,
              recomb=True, scat=True, chunk=64)
    return time.perf_counter() - t0


def _cpu_diffuse_bins(args):
    from oracle import qp_oracle as O
    import qpsim_b200 as Q
    import cases

    mask, dx, D, dt, field, reps = args
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "short_absorbing", Q.BoundaryCondition)
    op = O.DiffusionCN(mask, edges, bcs, dx, D[:, None] * np.ones((1, int(mask.sum()))), dt, False)  # setup, untimed
    st = np.repeat(field[None, :], D.size, axis=0)
    t0 = time.perf_counter()
    for _ in range(reps):
        op.step(st)
    return (time.perf_counter() - t0) / reps


def cpu_baseline(w, tabs, cells_sample=2048, bins_sample=None, steps=1):
    """Time the reference algorithm (oracle port: batched-over-cells collision step twice per time step, and the
    SuperLU Crank-Nicolson solve per bin; operator factorisation is setup and untimed, as in the reference) on a
    bounded sample, every host core busy, and scale linearly to the full workload."""
    from concurrent.futures import ProcessPoolExecutor

    cores = os.cpu_count() or 1
    n, ne = tabs["n"], w["num_energy_bins"]
    bins_sample = bins_sample or min(ne, cores)
    rng = np.random.default_rng(20260102)
    pick = np.sort(rng.choice(n, size=min(cells_sample, n), replace=False))
    chunks = np.array_split(pick, cores)
    small = {k: tabs[k] for k in ("Kr", "Ks", "rho", "idx_diff", "idx_sum", "sign", "dE")}
    jobs = [(tabs["state"][:, c].copy(), tabs["phonons"][:, c].copy(), small, w["dt"]) for c in chunks if c.size]
    bsel = np.linspace(0, ne - 1, bins_sample).astype(int)
    djobs = [(w["mask"], w["dx"], tabs["D"][[b]], w["dt"], w["initial_field"][w["mask"]], 3) for b in bsel]
    with ProcessPoolExecutor(max_workers=cores) as ex:
        t0 = time.perf_counter()
        list(ex.map(_cpu_collide_chunk, jobs))
        t_coll = time.perf_counter() - t0            # wall time with all cores busy
        t1 = time.perf_counter()
        per_bin = list(ex.map(_cpu_diffuse_bins, djobs))
        _ = time.perf_counter() - t1
    t_coll_full = t_coll * (n / pick.size)
    t_diff_full = float(np.mean(per_bin)) * ne / min(cores, ne)   # bins are independent: spread over the cores
    t_step = t_coll_full + t_diff_full
    return {
        "value": n * ne / t_step, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"oracle port of qpsim.solver: 2 collision half-steps on {pick.size} of {n} cells "
                   f"({t_coll:.2f} s wall on {cores} procs) + SuperLU CN solve of {bins_sample} of {ne} bins "
                   f"({np.mean(per_bin) * 1e3:.1f} ms/bin/step, factorisation untimed); scaled linearly"),
        "est_ms_per_step": t_step * 1e3,
    }


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def load_profile_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures of this
    workload (profiles/r1_ncu_kernels.json; bytes), or nothing when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_kernels.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        k = json.load(f)
    out = {}
    # the captures of the bench workload itself (the file also holds the 512-bin GEMM and the segmented 2048^2 sweeps)
    sw = [v["dram_bytes"] for n, v in k.items() if n.endswith("[k_sweep_x_pipe]") or n.endswith("[k_sweep_y_pipe]")]
    if sw:
        out["sweep"] = float(np.mean(sw))
    co = [v["dram_bytes"] for n, v in k.items() if n.endswith("[k_collide_struct]")]
    if co:
        out["collide"] = float(np.mean(co))
    return out


def run_single_gpu(args):
    import qpsim_b200 as Q
    import cases
    from qpsim_b200 import capi

    dev = int(os.environ.get("LOCAL_RANK", "0"))
    w = c2_workload()
    tabs = build_tables(w, Q)
    mask = w["mask"]
    ny, nx = mask.shape
    n, ne, nw = tabs["n"], w["num_energy_bins"], tabs["omega"].size
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
    flags = capi.F_DIFFUSION | capi.F_SCATTERING | capi.F_RECOMBINATION | capi.F_PAULI
    K, W = args.steps, args.warmup
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
        sweeps_per_step = None
        # ---- per-kernel device times (serialised by events; used for shares and the roofline) ----
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
    # algorithmic work: 16 B per cell*bin per directional sweep; 21*NE^2 flop per cell per
    # collision call with dynamic phonons
    ncd = ny * nx
    sweep_bytes = 16.0 * n * bin_sweeps          # bins that converged early are skipped by the kernel
    sweep_ms = tx + ty
    coll_flops = 21.0 * ne * ne * n * ncl
    prof = load_profile_traffic()
    roof_sweep = {"bound": "hbm", "achieved": sweep_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0,
                  "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": prof.get("sweep"),
                  "kernel": "k_sweep (x+y tridiagonal sweeps)", "launches": int(nxl + nyl),
                  "ms_per_launch": sweep_ms / max(1, nxl + nyl), "peak_source": peak_src}
    roof_sweep["frac"] = roof_sweep["achieved"] / roof_sweep["peak"]
    # what the kernels actually move: 24 B per DENSE grid cell, bin and sweep (x: u and b in, u* out; y: u* and u in, u
    # out; cells outside the mask are part of the dense lines) -- the distance between this figure and `achieved` is
    # the mask fill (cells / dense cells) times 16/24, not idle memory pipes
    roof_sweep["implementation"] = {
        "bytes_per_dense_cell_bin_sweep": 24, "dense_cells": int(ncd),
        "achieved": 24.0 * ncd * bin_sweeps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0, "unit": "GB/s"}
    roof_sweep["implementation"]["frac"] = roof_sweep["implementation"]["achieved"] / roof_sweep["peak"]
    roof_coll = {"bound": "fp64", "achieved": coll_flops / (tc * 1e-3) / 1e12 if tc > 0 else 0.0, "peak": fp64_peak,
                 "unit": "TFLOP/s", "traffic": prof.get("collide"), "kernel": "k_collide_struct", "launches": int(ncl),
                 "ms_per_launch": tc / max(1, ncl),
                 "peak_source": "measured in this run: DFMA loop on all SMs (qpb_measure_fp64)"}
    roof_coll["frac"] = roof_coll["achieved"] / roof_coll["peak"]
    dominant = roof_coll if tc >= sweep_ms else roof_sweep
    shares = {"collision_ms_per_step": tc / ks, "sweeps_ms_per_step": sweep_ms / ks,
              "other_ms_per_step": max(0.0, (ms_total / K) - (tc + sweep_ms) / ks)}
    # ---- end to end through the public API: host buffers in, host results out ----
    gen_spec = Q.ExternalGenerationSpec(mode="pulse", pulse_rate=w["pulse_rate"], pulse_start=w["pulse_start"],
                                        pulse_duration=w["pulse_duration"])
    kw = dict(mask=mask, edges=edges, edge_conditions=bcs, initial_field=w["initial_field"],
              diffusion_coefficient=w["diffusion_coefficient"], dt=w["dt"], total_time=w["dt"] * K, dx=w["dx"],
              store_every=K, energy_gap=w["energy_gap"], energy_min_factor=w["energy_min_factor"],
              energy_max_factor=w["energy_max_factor"], num_energy_bins=ne, enable_diffusion=True,
              enable_recombination=True, enable_scattering=True, dynes_gamma=w["dynes_gamma"], tau_0=w["tau_0"],
              T_c=w["T_c"], bath_temperature=w["bath_temperature"], external_generation=gen_spec, device=dev)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(2):   # untimed warm-up calls (module load, allocator cache, host page faults)
            Q.run_2d_crank_nicolson(**{**kw, "total_time": w["dt"] * 3, "store_every": 3})
        t0 = time.perf_counter()
        times, frames, mass, _, eframes, _ = Q.run_2d_crank_nicolson(**kw)
        t_e2e = time.perf_counter() - t0
    # uploads of one call: mask + three boundary arrays, the two base kernels, and the factors of the default initial
    # state (energy weights, spatial field, bath phonon occupations: the (NE, N) product is formed on the device)
    h2d = 8 * (ne + n + nw) + ncd * (1 + 3 * 8) + 8 * 2 * ne * ne
    d2h = 8 * (2 * ne * ncd + n)   # t=0 and final energy frames (dense, NaN padded) + integrated field
    e2e = {"value": n * ne * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "seconds": t_e2e, "note": "one run_2d_crank_nicolson call (context creation, geometry compile, uploads, "
           f"{K} steps, download of the t=0 and final energy frames), host numpy buffers in and out"}
    cpu = cpu_baseline(w, tabs)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "cells": n, "energy_bins": ne, "phonon_bins": int(nw), "dt_ns": w["dt"],
                   "processes": "masked CN diffusion (exact, PR-sweep iteration) + scattering + recombination, "
                                "dynamic phonons, pulse generation, Pauli check every step",
                   "l2": "state (64 MiB) + phonons (190 MiB) + work arrays exceed the 126 MB L2",
                   "sweeps_per_step": sweeps_per_step},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": dominant, "roofline_sweeps": roof_sweep, "roofline_collision": roof_coll,
        "time_shares": shares, "cpu_baseline": cpu,
        "peaks": {"hbm_gbs": peaks["hbm_gbs"], "copy_gbs_this_run": copy_gbs, "fp64_tflops_this_run": fp64_peak},
    }
    emit(line)


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference itself is pure
    Python and cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import qpsim_b200 as Q

    from qpsim_b200.multigpu import weak_tiling

    ty, tx = weak_tiling(max(1, args.gpus))
    w = c2_workload(tile_y=ty, tile_x=tx)
    tabs = build_tables(w, Q)
    vals = []
    last = None
    for _ in range(max(1, min(args.steps, 3))):
        last = cpu_baseline(w, tabs, cells_sample=1024)
        vals.append(last["value"])
    v = float(np.mean(vals))
    n, ne = tabs["n"], w["num_energy_bins"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": n * ne / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "cells": n, "energy_bins": ne},
        "cpu_baseline": {"kind": last["kind"], "cores": last["cores"], "sample": last["sample"], "value": v,
                         "unit": UNIT},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 (default): BASELINE configs[1], weak scaling over --gpus; c3: configs[2], 2048^2 x 256 "
                         "bins cut across the GPUs (strong scaling, sharded driver at every N)")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "c3":
        from qpsim_b200 import multigpu

        multigpu.bench_main(args, c3_workload, build_tables, ClockSampler, METRIC, UNIT, emit)
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from qpsim_b200 import multigpu

        multigpu.bench_main(args, c2_workload, build_tables, ClockSampler, METRIC, UNIT, emit)
        return
    run_single_gpu(args)


if __name__ == "__main__":
    main()
