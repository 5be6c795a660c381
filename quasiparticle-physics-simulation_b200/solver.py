"""Drop-in replacement of ``qpsim.solver.run_2d_crank_nicolson`` running the time loop on a B200.

Signature, argument checks, error types/messages and the returned 6-tuple follow the reference
(``qpsim/solver.py:999-1036`` signature, ``:1057-1077`` checks, ``:1085-1089`` step plan, ``:1296-1344`` Pauli
policy, ``:1367-1379`` / ``:1479-1494`` stored outputs, ``:1517-1587`` scalar mode).  Only the host-side setup
that the reference also does once per run happens here; every time step executes in ``libqpb.so`` through the
C ABI of ``include/qpb.h``.  Nothing in this module computes a time step on the CPU.
"""
from __future__ import annotations

import os
import threading

import warnings
from typing import Any, Callable

import numpy as np

from . import capi, physics, userexpr
from .geometry import compile_boundaries
from .models import ExternalGenerationSpec, check_collision_solver


def reconstruct_field(mask: np.ndarray, values: np.ndarray) -> np.ndarray:
    """NaN-padded 2-D frame from a compressed cell vector (solver.py:215-218)."""
    out = np.full(mask.shape, np.nan, dtype=float)
    out[mask] = values
    return out


def _step_plan(dt: float, total_time: float):
    full = int(np.floor(total_time / dt + 1e-12))
    rem = float(total_time - full * dt)
    if rem < 1e-12:
        rem = 0.0
    return full, rem, full + (1 if rem > 0.0 else 0)


class _Generation:
    """External generation (solver.py:878-964) as qpb_advance arguments.  ``constant`` / ``pulse`` are evaluated on the
    device.  ``custom`` bodies are evaluated on the host by :mod:`userexpr`; a body that never names ``t`` is
    evaluated once, uploaded with the first batch and stays resident (later batches: QPB_GEN_RESIDENT, any number of
    steps); a time-dependent body is translated into a postfix program that the device evaluates at the time of every
    step (QPB_GEN_PROGRAM, any number of steps per batch).  Only a body without a per-value meaning (``len(x)``,
    subscripts, ``np.arange`` ...) is still re-evaluated on the host and uploaded for every step, as the reference does."""

    def __init__(self, spec, E_bins, mask):
        self.mode = "none" if spec is None else spec.mode.strip().lower()
        self.spec = spec
        self.custom = None
        self.uploaded = False
        self.uploads = 0
        if self.mode == "constant" and not np.isfinite(float(spec.rate)):
            raise ValueError("External generation mode 'constant' produced non-finite values.")
        if self.mode == "pulse" and not np.isfinite(float(spec.pulse_rate)):
            raise ValueError("External generation mode 'pulse' produced non-finite values.")
        if self.mode == "custom":
            self.custom = userexpr.CustomGeneration(spec, E_bins, mask)

    @property
    def on_device(self) -> bool:
        return (self.custom is not None and self.custom.program is not None
                and os.environ.get("QPB_NO_GEN_PROGRAM", "0") != "1")

    def upload_program(self, ctx) -> None:
        if self.on_device:
            ctx.upload_generation_program(self.custom.program, self.custom.E, self.custom.x, self.custom.y)

    @property
    def one_step_batches(self) -> bool:
        return self.custom is not None and self.custom.time_dependent and not self.on_device

    def advance_args(self, t: float) -> dict:
        sp = self.spec
        if self.mode == "constant":
            return dict(gen_mode=capi.GEN_CONSTANT, rate=float(sp.rate))
        if self.mode == "pulse":
            return dict(gen_mode=capi.GEN_PULSE, rate=float(sp.pulse_rate), pulse_start=float(sp.pulse_start),
                        pulse_duration=float(sp.pulse_duration))
        if self.mode == "custom":
            if self.on_device:
                return dict(gen_mode=capi.GEN_PROGRAM)
            if self.uploaded and not self.custom.time_dependent:
                return dict(gen_mode=capi.GEN_RESIDENT)
            self.uploaded = True
            self.uploads += 1
            return dict(gen_mode=capi.GEN_ARRAY, gen_array=self.custom(t))
        return dict(gen_mode=capi.GEN_NONE)


class _PauliPolicy:
    """solver.py:1296-1344: forbidden-state and occupation checks after every step."""

    def __init__(self, E_bins, mask, warn_thr, err_thr, enforce):
        # the (row, col) of a cell is only needed inside a message: found on demand, not tabulated for every cell
        self.E, self.mask = E_bins, np.asarray(mask, dtype=bool)
        self._flat = None
        self.warn_thr, self.err_thr, self.enforce = warn_thr, err_thr, enforce
        self.warned = False
        self.n = int(self.mask.sum())

    def _pixel(self, px: int):
        if self._flat is None:
            self._flat = np.flatnonzero(self.mask.ravel())
        return divmod(int(self._flat[px]), self.mask.shape[1])

    def check(self, rec, step_idx: int, time_ns: float) -> None:
        max_occ, max_index, forbidden = rec
        if forbidden >= 0:
            ie, px = divmod(int(forbidden), self.n)
            row, col = self._pixel(px)
            msg = (
                f"Detected non-zero quasiparticle density in forbidden state "
                f"(rho≈0): step={step_idx}, t={time_ns:.6g} ns, "
                f"E={self.E[ie]:.6g} μeV, pixel=({int(row)},{int(col)})."
            )
            if self.enforce:
                raise ValueError(msg)
            if not self.warned:
                warnings.warn(msg, stacklevel=3)
                self.warned = True
        ie, px = divmod(int(max_index), self.n)
        if self.err_thr is not None and max_occ > self.err_thr:
            row, col = self._pixel(px)
            msg = (
                f"Pauli occupation exceeded limit: f={max_occ:.6g} > {self.err_thr:.6g} "
                f"at step={step_idx}, t={time_ns:.6g} ns, "
                f"E={self.E[ie]:.6g} μeV, pixel=({int(row)},{int(col)})."
            )
            if self.enforce:
                raise ValueError(msg)
            if not self.warned:
                warnings.warn(msg, stacklevel=3)
                self.warned = True
        if self.warn_thr is not None and max_occ > self.warn_thr and not self.warned:
            row, col = self._pixel(px)
            warnings.warn(
                "High occupation detected (Pauli blocking regime): "
                f"max f={max_occ:.6g} at step={step_idx}, t={time_ns:.6g} ns, "
                f"E={self.E[ie]:.6g} μeV, pixel=({int(row)},{int(col)}).",
                stacklevel=3,
            )
            self.warned = True


def _fixed_phonon_history(mask, times, bath_temperature):
    """Scalar-mode phonon scaffold (solver.py:373-426 with phonon_energy_bins=None)."""
    n = int(np.sum(mask))
    base = reconstruct_field(mask, np.full(n, float(bath_temperature), dtype=float))
    meta = {
        "mode": "fixed_temperature",
        "phonon_temperature_K": float(bath_temperature),
        "field_units": "K",
        "energy_frame_units": "occupation",
        "omega_bins_match_qp_energy_bins": False,
    }
    return [base.copy() for _ in range(len(times))], None, None, meta


last_run_info: dict[str, Any] = {}


def run_2d_crank_nicolson(
    mask: np.ndarray,
    edges: list,
    edge_conditions: dict,
    initial_field: np.ndarray,
    diffusion_coefficient: float,
    dt: float,
    total_time: float,
    dx: float,
    store_every: int = 1,
    energy_gap: float = 0.0,
    energy_min_factor: float = 1.0,
    energy_max_factor: float = 10.0,
    num_energy_bins: int = 50,
    energy_weights: np.ndarray | None = None,
    enable_diffusion: bool = True,
    enable_recombination: bool = False,
    enable_scattering: bool = False,
    dynes_gamma: float = 0.0,
    collision_solver: str = "fischer_catelani_local",
    tau_0: float = 440.0,
    tau_s: float | None = None,
    tau_r: float | None = None,
    T_c: float = 1.2,
    bath_temperature: float = 0.1,
    external_generation: ExternalGenerationSpec | None = None,
    initial_condition_spec: Any | None = None,
    gap_expression: str = "",
    precomputed: dict | None = None,
    pauli_warn_threshold: float | None = 0.5,
    pauli_error_threshold: float | None = 1.0,
    enforce_pauli: bool = True,
    pauli_density_floor: float = 1e-18,
    freeze_phonon_dynamics: bool = False,
    phonon_history_out: dict[str, Any] | None = None,
    progress_callback: Callable[[float, np.ndarray], None] | None = None,
    *,
    device: int = 0,
    devices: list[int] | None = None,
    diffusion_tolerance: float = 0.0,
    store_energy_frames: bool = True,
):
    """Same contract as the reference function; the keyword-only extras select the GPU(s), the residual tolerance of
    the Crank-Nicolson solve and (for benchmarks) allow skipping the NE full-frame copies.

    ``devices=[g0, g1, ...]`` (energy-resolved mode) runs the loop sharded over several GPUs of one box: diffusion by
    energy bin, collisions by cell (:mod:`multigpu`).  Under ``torchrun`` every rank makes the
    same call and rank r drives ``devices[r]``; from a plain process the ranks are spawned for the duration of the
    call.  Rank 0 (the caller, when spawned) gets the full result; the other ranks get the same ``times``, ``frames``,
    ``mass`` and ``None`` in place of the per-energy frames."""
    if dt <= 0 or total_time <= 0:
        raise ValueError("dt and total_time must be positive.")
    if enable_diffusion and diffusion_coefficient <= 0:
        raise ValueError("Diffusion coefficient must be positive.")
    if store_every <= 0:
        store_every = 1
    # QPB_FORCE_SHARDED=1 (tests): a one-entry device list also takes the sharded loop, as a world of one rank
    force_sharded = devices is not None and os.environ.get("QPB_FORCE_SHARDED", "0") == "1"
    if devices is not None and len(devices) == 1 and not force_sharded:
        device = int(devices[0])
    mask = np.asarray(mask)
    if initial_field.shape != mask.shape:
        raise ValueError("Initial field shape must match mask shape.")
    mask_b = mask.astype(bool)
    n = int(np.sum(mask_b))
    if n == 0:
        raise ValueError("Geometry mask has no interior points.")
    if phonon_history_out is not None:
        phonon_history_out.clear()
    tau_s_eff = float(tau_s if tau_s is not None else tau_0)
    tau_r_eff = float(tau_r if tau_r is not None else tau_0)
    if enable_scattering and tau_s_eff <= 0:
        raise ValueError("tau_s must be positive when scattering is enabled.")
    if enable_recombination and tau_r_eff <= 0:
        raise ValueError("tau_r must be positive when recombination is enabled.")
    if external_generation is not None:
        external_generation.validate()
    bcx = bcy = src = None
    if enable_diffusion:
        bcx, bcy, src = compile_boundaries(mask_b, edges, edge_conditions, dx)
    ny, nx = mask_b.shape
    full_steps, remainder_dt, total_steps = _step_plan(dt, total_time)
    info = last_run_info
    info.clear()

    # ------------------------------------------------------------------ legacy scalar mode (energy_gap == 0)
    if not energy_gap > 0.0:
        values = np.asarray(initial_field)[mask_b].astype(float)
        flags = capi.F_SCALAR | (capi.F_DIFFUSION if enable_diffusion else 0)
        times = [0.0]
        frames = [reconstruct_field(mask_b, values)]
        mass = [float(np.sum(values) * dx * dx)]
        _callback(progress_callback, 0.0, frames[0])
        with capi.Context(ny=ny, nx=nx, ne=1, nw=0, ncell=n, flags=flags, dx=dx, dE=1.0, device=device,
                          diff_tol=diffusion_tolerance) as ctx:
            ctx.upload_geometry(mask_b, bcx, bcy, src)
            if enable_diffusion:
                ctx.upload_diffusion(np.array([float(diffusion_coefficient)]))
                ctx.prepare_diffusion(0, dt)
                if remainder_dt > 0.0:
                    ctx.prepare_diffusion(1, remainder_dt)
            ctx.set_state(values[None, :])
            current_time = 0.0
            step = 0
            while step < total_steps:
                nxt = min(((step // store_every) + 1) * store_every, total_steps)
                if nxt > full_steps and step < full_steps:
                    nxt = full_steps
                is_final = step >= full_steps
                h = remainder_dt if is_final else dt
                count = nxt - step
                ctx.advance(count, h, slot=1 if is_final else 0, t_start=current_time)
                for _ in range(count):
                    current_time += h
                step = nxt
                if step % store_every == 0 or step == total_steps:
                    cur = ctx.get_integrated()
                    times.append(float(current_time))
                    frame = reconstruct_field(mask_b, cur)
                    frames.append(frame)
                    mass.append(float(np.sum(cur) * dx * dx))
                    _callback(progress_callback, float(current_time), frame)
            info.update(ctx.diag())
        limits = _color_limits(frames)
        if phonon_history_out is not None:
            pf, pef, pb, meta = _fixed_phonon_history(mask_b, times, bath_temperature)
            phonon_history_out.update({"phonon_frames": pf, "phonon_energy_frames": pef,
                                       "phonon_energy_bins": pb, "phonon_metadata": meta})
        return times, frames, mass, limits, None, None

    # ------------------------------------------------------------------ energy-resolved mode
    gap = energy_gap
    E_bins, dE = physics.build_energy_grid(gap, energy_min_factor, energy_max_factor, num_energy_bins)
    ne = int(num_energy_bins)
    custom_qp_state = None
    if initial_condition_spec is not None:
        custom_qp_state = userexpr.initial_qp_state(mask_b, E_bins, initial_condition_spec)
    if precomputed is None and gap_expression.strip():
        # what precompute_arrays(..., include_collision_kernels=False) hands the solver (solver.py:1105-1124)
        precomputed = userexpr.precompute_from_gap_expression(gap_expression, mask_b, E_bins, energy_gap,
                                                              diffusion_coefficient)
    has_pre = precomputed is not None
    nonuniform_gap = has_pre and not bool(precomputed.get("is_uniform", True))
    check_collision_solver(collision_solver)
    if has_pre:
        D_array = np.asarray(precomputed["D_array"], dtype=float)
    else:
        # the reference replicates D(E) over the cells (solver.py:1141-1143) and then only reads column 0
        # (solver.py:1167); the device takes the NE values
        D_array = diffusion_coefficient * np.sqrt(np.maximum(0.0, 1.0 - (gap / E_bins) ** 2))
    collisions = bool(enable_recombination or enable_scattering)

    omega_bins, idx_diff, idx_sum, diff_sign = physics.phonon_frequency_map(E_bins)
    n_ph_eq = physics.thermal_phonon_occupation(omega_bins, bath_temperature)
    # default initial phonons: the bath occupation in every cell (solver.py:1183-1185).  The (Nw, N) host array is
    # only materialised when somebody needs it; the device broadcasts the Nw values itself.
    phonon_state = None
    if initial_condition_spec is not None:
        phonon_state, _ = userexpr.initial_phonon_state(mask_b, omega_bins, initial_condition_spec, bath_temperature)
    nw = int(omega_bins.size)

    # density of states and base kernels: one table per distinct gap value (solver.py:1203-1238 builds the same
    # cache, then replicates it per pixel; the device indexes the cache through gap_id instead)
    gap_id = None
    if nonuniform_gap:
        gap_values = precomputed.get("gap_values") if has_pre else None
        if gap_values is None:
            gap_values = np.full(n, gap, dtype=float)
        gap_values = np.asarray(gap_values, dtype=float).reshape(-1)
        uniq, gap_id = np.unique(gap_values, return_inverse=True)
        gap_id = np.asarray(gap_id).reshape(-1).astype(np.int32)
        rho_tab = np.stack([physics.density_of_states(E_bins, float(g), dynes_gamma) for g in uniq])
        Kr_tab = (np.stack([physics.recombination_kernel_base(E_bins, float(g), tau_r_eff, T_c) for g in uniq])
                  if enable_recombination else None)
        Ks_tab = (np.stack([physics.scattering_kernel_base(E_bins, float(g), tau_s_eff, T_c) for g in uniq])
                  if enable_scattering else None)
    else:
        rho_tab = physics.density_of_states(E_bins, gap, dynes_gamma)[None, :]
        Kr_tab = physics.recombination_kernel_base(E_bins, gap, tau_r_eff, T_c)[None] if enable_recombination else None
        Ks_tab = physics.scattering_kernel_base(E_bins, gap, tau_s_eff, T_c)[None] if enable_scattering else None
    ngap = rho_tab.shape[0]

    if custom_qp_state is not None:
        state = np.asarray(custom_qp_state, dtype=float)
        if state.shape != (ne, n):
            raise ValueError(
                f"Full custom quasiparticle profile must have shape ({ne}, {n}); got {state.shape}."
            )
        if not np.all(np.isfinite(state)):
            raise ValueError("Full custom quasiparticle profile produced non-finite values.")
        if np.any(state < 0):
            raise ValueError("Full custom quasiparticle profile must be non-negative.")
    else:
        spatial = np.asarray(initial_field)[mask_b].astype(float)
        if energy_weights is not None:
            raw = np.asarray(energy_weights, dtype=float)
            if raw.ndim != 1:
                raise ValueError("energy_weights must be a 1D array.")
            if raw.shape[0] != ne:
                raise ValueError(f"energy_weights must have length {ne}, got {raw.shape[0]}.")
            if not np.all(np.isfinite(raw)):
                raise ValueError("energy_weights must contain only finite values.")
            if np.any(raw < 0):
                raise ValueError("energy_weights must be non-negative.")
            total = np.sum(raw) * dE
            weights = raw / total if total > 0 else np.ones(ne, dtype=float) / (ne * dE)
        else:
            rho0 = physics.density_of_states(E_bins, gap, dynes_gamma)
            total = np.sum(rho0) * dE
            weights = rho0 / total if total > 0 else np.ones(ne, dtype=float) / (ne * dE)
        # state[i] = spatial * weights[i] (solver.py:1281-1283) is formed on the device from its two factors; the
        # (NE, N) host array is only materialised when an explicit phonon state has to travel with it
        state = None

    policy = _PauliPolicy(E_bins, mask_b, pauli_warn_threshold, pauli_error_threshold, enforce_pauli)

    flags = capi.F_PAULI
    if enable_diffusion:
        flags |= capi.F_DIFFUSION
        if nonuniform_gap:
            flags |= capi.F_VARIABLE_D
    if enable_scattering:
        flags |= capi.F_SCATTERING
    if enable_recombination:
        flags |= capi.F_RECOMBINATION
    if freeze_phonon_dynamics:
        flags |= capi.F_FREEZE_PHONONS

    want_ph_hist = phonon_history_out is not None
    if devices is not None and (len(devices) > 1 or force_sharded):
        from . import multigpu

        setup = dict(
            mask=mask_b, bcx=bcx, bcy=bcy, src=src, dx=dx, dE=dE, ne=ne, nw=nw, n=n, E_bins=E_bins,
            omega_bins=omega_bins, D=D_array if nonuniform_gap else (D_array[:, 0] if D_array.ndim == 2 else D_array),
            variable_D=bool(nonuniform_gap and enable_diffusion), rho=rho_tab, Kr=Kr_tab, Ks=Ks_tab, gap_id=gap_id,
            idx_diff=idx_diff, idx_sum=idx_sum, sign=diff_sign, state=state,
            weights=None if state is not None else weights, spatial=None if state is not None else spatial,
            phonon_state=phonon_state, phonon_bins=n_ph_eq, diffusion=bool(enable_diffusion),
            scattering=bool(enable_scattering), recombination=bool(enable_recombination),
            freeze_phonons=bool(freeze_phonon_dynamics), pauli_floor=pauli_density_floor,
            diff_tol=diffusion_tolerance, dt=dt, full_steps=full_steps, remainder_dt=remainder_dt,
            total_steps=total_steps, store_every=store_every, generation=external_generation,
            store_energy_frames=store_energy_frames, want_phonon_history=want_ph_hist, devices=list(devices),
            policy=policy,
        )
        out = multigpu.run_dropin(setup, progress_callback)
        info.update(out.pop("info", {}))
        if phonon_history_out is not None and out.get("phonon_history") is not None:
            phonon_history_out.clear()
            phonon_history_out.update(out["phonon_history"])
        frames = out["frames"]
        return out["times"], frames, out["mass"], _color_limits(frames), out["energy_frames"], E_bins
    ph_frames: list = []
    ph_energy_frames: list = []
    ph_widths = physics.integration_widths_from_centers(omega_bins, fallback_width=dE) if want_ph_hist else None

    def snapshot_phonons(ph):
        ph_energy_frames.append([reconstruct_field(mask_b, ph[i]) for i in range(ph.shape[0])])
        ph_frames.append(reconstruct_field(mask_b, np.sum(ph * ph_widths[:, None], axis=0)))

    with capi.Context(ny=ny, nx=nx, ne=ne, nw=nw if collisions else 0, ncell=n, ngap=ngap, flags=flags, dx=dx,
                      dE=dE, device=device, diff_tol=diffusion_tolerance, pauli_floor=pauli_density_floor) as ctx:
        ctx.upload_geometry(mask_b, bcx, bcy, src)
        if enable_diffusion:
            if nonuniform_gap:
                ctx.upload_diffusion(D_array)
            else:
                ctx.upload_diffusion(D_array[:, 0] if D_array.ndim == 2 else D_array)
            ctx.prepare_diffusion(0, dt)
            if remainder_dt > 0.0:
                ctx.prepare_diffusion(1, remainder_dt)
        ctx.upload_collision(Kr_tab, Ks_tab, rho_tab, gap_id,
                             idx_diff if collisions else None, idx_sum if collisions else None,
                             diff_sign if collisions else None)
        if state is None and (phonon_state is None or not collisions):
            ctx.set_state_separable(weights, spatial, n_ph_eq if collisions else None)
        else:
            if state is None:
                state = spatial[None, :] * weights[:, None]
            if collisions and phonon_state is None:
                ctx.set_state_uniform_phonons(state, n_ph_eq)
            else:
                ctx.set_state(state, phonon_state if collisions else None)
        policy.check(ctx.pauli(), 0, 0.0)

        if want_ph_hist:
            if phonon_state is None:
                phonon_state = n_ph_eq[:, None] * np.ones((1, n), dtype=float)
            snapshot_phonons(phonon_state)
        # sum_i state[i] * dE (solver.py:1356), summed on the device like every later frame
        integrated = ctx.get_integrated() if state is None else np.sum(state, axis=0) * dE
        times = [0.0]
        frames = [reconstruct_field(mask_b, integrated)]
        # The NE frames of a stored step are snapshotted on the device and downloaded by a helper thread while the
        # next steps run; the list entry is filled in when the download is joined (before the next snapshot / return).
        energy_frames = []
        pending = []   # [(thread, result holder, index into energy_frames)]

        def join_pending():
            while pending:
                th, box, idx = pending.pop()
                th.join()
                if "error" in box:
                    raise box["error"]
                energy_frames[idx] = list(box["frames"])

        def store_frames():
            if not store_energy_frames:
                energy_frames.append(None)
                return
            join_pending()                 # one snapshot buffer: the previous download has to be through
            ctx.frames_snapshot()
            box = {}

            def work():
                try:
                    box["frames"] = ctx.frames_download()
                except BaseException as exc:   # re-raised in the calling thread by join_pending
                    box["error"] = exc

            th = threading.Thread(target=work, daemon=True)
            energy_frames.append(None)
            pending.append((th, box, len(energy_frames) - 1))
            th.start()

        try:
            store_frames()
            mass = [float(np.sum(integrated) * dx * dx)]
            _callback(progress_callback, 0.0, frames[0])

            current_time = 0.0
            step = 0
            generation = _Generation(external_generation, E_bins, mask_b)
            generation.upload_program(ctx)
            while step < total_steps:
                nxt = min(((step // store_every) + 1) * store_every, total_steps)
                if nxt > full_steps and step < full_steps:
                    nxt = full_steps
                if generation.one_step_batches:
                    nxt = step + 1
                is_final = step >= full_steps
                h = remainder_dt if is_final else dt
                count = nxt - step
                gen_kwargs = generation.advance_args(current_time)
                try:
                    recs = ctx.advance(count, h, slot=1 if is_final else 0, t_start=current_time, want_pauli=True,
                                       **gen_kwargs)
                except capi.QpbError as exc:
                    # the device-side checks of a custom generation body carry the reference's messages (solver.py:954-962)
                    if exc.message.startswith("External generation mode 'custom'"):
                        raise ValueError(exc.message) from None
                    raise
                for k in range(count):
                    policy.check(recs[k], step + k + 1, current_time + h)
                    current_time += h
                step = nxt
                if step % store_every == 0 or step == total_steps:
                    integrated = ctx.get_integrated()
                    times.append(float(current_time))
                    frame = reconstruct_field(mask_b, integrated)
                    frames.append(frame)
                    # NE NaN-padded frames assembled on the device (one dense download instead of NE host scatters)
                    store_frames()
                    if want_ph_hist:
                        ph = ctx.get_state(want_qp=False, want_phonons=True)[1] if collisions else None
                        snapshot_phonons(ph if ph is not None else phonon_state)
                    mass.append(float(np.sum(integrated) * dx * dx))
                    _callback(progress_callback, float(current_time), frame)
            join_pending()
        finally:
            # an exception on the way (Pauli violation, callback, device error) must not leave a download running into a
            # context that is being destroyed
            for th, _, _ in pending:
                th.join()
        info.update(ctx.diag())
        info["generation_uploads"] = generation.uploads
        info["generation_on_device"] = bool(generation.on_device)

    limits = _color_limits(frames)
    if phonon_history_out is not None:
        phonon_history_out.clear()
        phonon_history_out.update({
            "phonon_frames": ph_frames,
            "phonon_energy_frames": ph_energy_frames,
            "phonon_energy_bins": np.asarray(omega_bins, dtype=float).copy(),
            "phonon_metadata": {
                "mode": "dynamic_local_coupled",
                "field_units": "integrated_occupation",
                "energy_frame_units": "occupation",
            },
        })
    return times, frames, mass, limits, energy_frames, E_bins


def _callback(cb, t, frame):
    if cb is None:
        return
    try:
        cb(t, np.array(frame, copy=True))
    except Exception:
        pass


def _color_limits(frames):
    stack = np.stack(frames)
    lo = float(np.nanmin(stack))
    hi = float(np.nanmax(stack))
    if abs(hi - lo) < 1e-12:
        hi = lo + 1e-9
    return [lo, hi]


# --------------------------------------------------------------------------------------------------------------
# in-place collision helpers (solver.py:794-875)
# --------------------------------------------------------------------------------------------------------------
def _collide_in_place(state, phonon_state, Kr_tab, Ks_tab, rho_tab, gap_id, idx_diff, idx_sum, sign, dE, dt,
                      enable_recombination, enable_scattering, update_phonons, device):
    ne, n = state.shape
    if phonon_state.shape[1] != n:
        raise ValueError("phonon_state shape does not match quasiparticle state.")
    nw = phonon_state.shape[0]
    flags = 0
    if enable_scattering and Ks_tab is not None:
        flags |= capi.F_SCATTERING
    if enable_recombination and Kr_tab is not None:
        flags |= capi.F_RECOMBINATION
    if not update_phonons:
        flags |= capi.F_FREEZE_PHONONS
    ngap = rho_tab.shape[0]
    with capi.Context(ny=1, nx=n, ne=ne, nw=nw, ncell=n, ngap=ngap, flags=flags, dx=1.0, dE=float(dE),
                      device=device) as ctx:
        ctx.upload_geometry(np.ones((1, n), dtype=np.uint8))
        ctx.upload_collision(Kr_tab if flags & capi.F_RECOMBINATION else None,
                             Ks_tab if flags & capi.F_SCATTERING else None, rho_tab, gap_id, idx_diff, idx_sum, sign)
        ctx.set_state(state, phonon_state)
        ctx.collide(dt)
        new_state, new_ph = ctx.get_state(want_phonons=True)
    state[...] = new_state
    if update_phonons and (flags & (capi.F_SCATTERING | capi.F_RECOMBINATION)):
        phonon_state[...] = new_ph


def apply_collision_step_fischer_catelani_uniform(state, phonon_state, K_r0, K_s0, rho_bins, omega_idx_diff,
                                                  omega_idx_sum, diff_sign, dE, dt, *, enable_recombination,
                                                  enable_scattering, update_phonons=True, device=0):
    """GPU version of solver.py:794-831: one coupled quasiparticle-phonon update of every column, in place."""
    rho_tab = np.asarray(rho_bins, dtype=float)[None, :]
    Kr = None if K_r0 is None else np.asarray(K_r0, dtype=float)[None]
    Ks = None if K_s0 is None else np.asarray(K_s0, dtype=float)[None]
    _collide_in_place(state, phonon_state, Kr, Ks, rho_tab, None, omega_idx_diff, omega_idx_sum, diff_sign, dE, dt,
                      enable_recombination, enable_scattering, update_phonons, device)


def apply_collision_step_fischer_catelani_nonuniform(state, phonon_state, K_r0_all, K_s0_all, rho_all,
                                                     omega_idx_diff, omega_idx_sum, diff_sign, dE, dt, *,
                                                     enable_recombination, enable_scattering,
                                                     update_phonons=True, device=0):
    """GPU version of solver.py:834-875.  Per-pixel tables are de-duplicated into one table per distinct
    (rho, K_r0, K_s0) triple before upload."""
    n = state.shape[1]
    rho_all = np.asarray(rho_all, dtype=float)
    if rho_all.shape[0] != n:
        raise ValueError("rho_all shape does not match quasiparticle state.")
    keys: dict[bytes, int] = {}
    gap_id = np.empty(n, dtype=np.int32)
    first: list[int] = []
    for px in range(n):
        key = rho_all[px].tobytes()
        if K_r0_all is not None:
            key += np.ascontiguousarray(K_r0_all[px]).tobytes()
        if K_s0_all is not None:
            key += np.ascontiguousarray(K_s0_all[px]).tobytes()
        gid = keys.get(key)
        if gid is None:
            gid = len(first)
            keys[key] = gid
            first.append(px)
        gap_id[px] = gid
    rho_tab = np.stack([rho_all[p] for p in first])
    Kr = None if K_r0_all is None else np.stack([np.asarray(K_r0_all[p], dtype=float) for p in first])
    Ks = None if K_s0_all is None else np.stack([np.asarray(K_s0_all[p], dtype=float) for p in first])
    _collide_in_place(state, phonon_state, Kr, Ks, rho_tab, gap_id, omega_idx_diff, omega_idx_sum, diff_sign, dE, dt,
                      enable_recombination, enable_scattering, update_phonons, device)


# --------------------------------------------------------------------------------------------------------------
# fixed-bath forward-Euler collision forms (solver.py:551-605)
# --------------------------------------------------------------------------------------------------------------
def _euler_in_place(state, kind, K, vec, dE, dt, device):
    state_arr = np.asarray(state)
    if state_arr.ndim != 2:
        raise ValueError("state must have shape (num_energy_bins, num_spatial_pts).")
    ne, n = state_arr.shape
    K = np.asarray(K, dtype=float)
    vec = np.asarray(vec, dtype=float)
    if K.shape != (ne, ne) or vec.shape != (ne,):
        raise ValueError("kernel / per-bin vector shapes do not match the state.")
    with capi.Context(ny=1, nx=n, ne=ne, nw=0, ncell=n, flags=0, dx=1.0, dE=float(dE), device=device) as ctx:
        ctx.upload_geometry(np.ones((1, n), dtype=np.uint8))
        ctx.set_state(state_arr)
        ctx.euler_step(kind, K, vec, dt)
        new_state, _ = ctx.get_state(want_phonons=False)
    state[...] = new_state


def apply_scattering_step(state, K_s, rho_bins, dE, dt, *, device=0):
    """GPU version of solver.py:551-581: one forward-Euler step of quasiparticle-phonon scattering against a fixed
    bath, in place.  scat_in = dE rho (1-f) (K_s^T n), scat_out = n dE (K_s rho (1-f)): the K(E,E') contraction over all
    cells as one FP64 tensor-core GEMM (64 bins and more) or a fused GEMV."""
    _euler_in_place(state, 1, K_s, rho_bins, dE, dt, device)


def apply_recombination_step(state, K_r, G_therm, dE, dt, *, device=0):
    """GPU version of solver.py:584-605: one forward-Euler step of recombination + thermal generation, in place:
    n += dt (G_therm - 2 n dE (K_r n)), the bilinear form n.R.n per cell."""
    _euler_in_place(state, 2, K_r, G_therm, dE, dt, device)
