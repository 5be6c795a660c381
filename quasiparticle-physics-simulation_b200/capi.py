"""ctypes binding of ``libqpb.so`` (C ABI declared in ``include/qpb.h``).

The library is built in-tree by :func:`build_library` (``nvcc`` for sm_100a only).  There is no Python or CPU
fallback: if the shared object is missing or no B200 is visible every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libqpb.so")
INCLUDE_DIR = os.path.join(REPO_DIR, "include")

ABI_VERSION = 1

F_DIFFUSION = 1 << 0
F_SCATTERING = 1 << 1
F_RECOMBINATION = 1 << 2
F_FREEZE_PHONONS = 1 << 3
F_VARIABLE_D = 1 << 4
F_PAULI = 1 << 5
F_SCALAR = 1 << 6

GEN_NONE, GEN_CONSTANT, GEN_PULSE, GEN_ARRAY, GEN_RESIDENT, GEN_PROGRAM = 0, 1, 2, 3, 4, 5

E_NOCONV = -5


class QpbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libqpb error {code}: {message}")
        self.code = code
        self.message = message


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
        ("ne", C.c_int32), ("nw", C.c_int32), ("ncell", C.c_int32), ("ngap", C.c_int32),
        ("flags", C.c_uint32), ("reserved", C.c_int32),
        ("dx", C.c_double), ("dE", C.c_double), ("diff_tol", C.c_double), ("pauli_floor", C.c_double),
    ]


class Diag(C.Structure):
    _fields_ = [
        ("steps_done", C.c_int64), ("sweeps", C.c_int64), ("bin_sweeps", C.c_int64),
        ("pr_iterations", C.c_int64), ("last_delta", C.c_double), ("direct_mode", C.c_int32),
        ("commuting", C.c_int32), ("kernel_launches", C.c_int64), ("last_advance_ms", C.c_double),
        ("sweep_path", C.c_int32), ("reserved", C.c_int32),
    ]


class PauliRec(C.Structure):
    _fields_ = [("max_occ", C.c_double), ("max_index", C.c_int64), ("forbidden", C.c_int64)]


class Generation(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("reserved", C.c_int32), ("rate", C.c_double), ("pulse_start", C.c_double),
        ("pulse_duration", C.c_double), ("array", C.c_void_p),
    ]


class GenOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("value", C.c_double)]


SOURCES = ["qpb_api.cu", "qpb_diffusion.cu", "qpb_sweep_fast.cu", "qpb_sweep_pipe.cu", "qpb_collision.cu", "qpb_aux.cu",
           "qpb_krylov.cu", "qpb_spectral.cu", "qpb_resident.cu"]

# every symbol include/qpb.h declares; tests check that the built library exports all of them
EXPORTED = [
    "qpb_last_error", "qpb_abi_version", "qpb_device_count", "qpb_create", "qpb_destroy",
    "qpb_upload_geometry", "qpb_upload_diffusion", "qpb_prepare_diffusion", "qpb_upload_collision",
    "qpb_set_state", "qpb_get_state", "qpb_get_integrated", "qpb_advance", "qpb_collide", "qpb_diffuse",
    "qpb_pauli", "qpb_get_diag", "qpb_synchronize", "qpb_enable_timers", "qpb_reset_timers", "qpb_get_timer",
    "qpb_device_ptr", "qpb_measure_fp64", "qpb_measure_copy", "qpb_scatter_block", "qpb_gather_block",
    "qpb_add_generation", "qpb_set_stream", "qpb_get_frames", "qpb_trim_cache", "qpb_set_state_uniform_phonons", "qpb_pauli_record", "qpb_pauli_fetch", "qpb_set_exchange", "qpb_collide_exchange",
    "qpb_ipc_export", "qpb_ipc_open", "qpb_ipc_close", "qpb_euler_step", "qpb_set_state_separable",
    "qpb_frames_snapshot", "qpb_frames_download", "qpb_upload_generation_program", "qpb_eval_generation_program",
    "qpb_add_generation_program", "qpb_add_generation_array", "qpb_generation_status",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libqpb cannot be built")


def library_is_current() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR)] + [os.path.join(INCLUDE_DIR, "qpb.h")]
    return all(os.path.getmtime(d) <= built for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package for sm_100a into lib/libqpb.so (cross-compiles without a GPU).
    The translation units are compiled in parallel into lib/obj/ and only when they (or a header) changed."""
    if not force and library_is_current():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(INCLUDE_DIR, "qpb.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    base = [
        _nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC", "-I", INCLUDE_DIR, "-I", CSRC_DIR,
    ]
    if verbose:
        base += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        path = os.path.join(CSRC_DIR, src)
        if (not force and not verbose and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(path), newest_header)):
            return obj, ""
        res = subprocess.run(base + ["-c", path, "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        print("".join(log for _, log in results))
    res = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"]
                         + [obj for obj, _ in results] + ["-o", LIB_PATH], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def load_library():
    """dlopen libqpb.so and declare the prototypes.  Raises when the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for this path)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.qpb_last_error.restype = C.c_char_p
    lib.qpb_last_error.argtypes = []
    lib.qpb_abi_version.restype = C.c_int
    lib.qpb_device_count.restype = C.c_int
    lib.qpb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.qpb_destroy.argtypes = [vp]
    lib.qpb_destroy.restype = None
    lib.qpb_upload_geometry.argtypes = [vp, vp, vp, vp, vp]
    lib.qpb_upload_diffusion.argtypes = [vp, vp]
    lib.qpb_prepare_diffusion.argtypes = [vp, C.c_int, dbl]
    lib.qpb_upload_collision.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.qpb_set_state.argtypes = [vp, vp, vp]
    lib.qpb_get_state.argtypes = [vp, vp, vp]
    lib.qpb_set_state_uniform_phonons.argtypes = [vp, vp, vp]
    lib.qpb_set_state_separable.argtypes = [vp, vp, vp, vp]
    lib.qpb_get_integrated.argtypes = [vp, vp]
    lib.qpb_get_frames.argtypes = [vp, vp]
    lib.qpb_frames_snapshot.argtypes = [vp]
    lib.qpb_frames_download.argtypes = [vp, vp]
    lib.qpb_trim_cache.argtypes = []
    lib.qpb_advance.argtypes = [vp, i32, dbl, i32, dbl, C.POINTER(Generation), vp]
    lib.qpb_upload_generation_program.argtypes = [vp, C.POINTER(GenOp), i32, vp, vp, vp]
    lib.qpb_eval_generation_program.argtypes = [vp, dbl, vp]
    lib.qpb_collide.argtypes = [vp, dbl]
    lib.qpb_diffuse.argtypes = [vp, i32]
    lib.qpb_pauli.argtypes = [vp, C.POINTER(PauliRec)]
    lib.qpb_pauli_record.argtypes = [vp, i32]
    lib.qpb_set_exchange.argtypes = [vp, i32, C.POINTER(vp), i64, vp, vp, vp]
    lib.qpb_collide_exchange.argtypes = [vp, dbl, i32]
    lib.qpb_euler_step.argtypes = [vp, i32, vp, vp, dbl]
    lib.qpb_ipc_export.argtypes = [vp, C.c_int, vp]
    lib.qpb_ipc_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.qpb_ipc_close.argtypes = [C.c_int, vp]
    lib.qpb_pauli_fetch.argtypes = [vp, i32, vp]
    lib.qpb_get_diag.argtypes = [vp, C.POINTER(Diag)]
    lib.qpb_synchronize.argtypes = [vp]
    lib.qpb_enable_timers.argtypes = [vp, C.c_int]
    lib.qpb_reset_timers.argtypes = [vp]
    lib.qpb_get_timer.argtypes = [vp, C.c_int, C.POINTER(dbl), C.POINTER(i64)]
    lib.qpb_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    lib.qpb_scatter_block.argtypes = [vp, vp, i32, i32]
    lib.qpb_gather_block.argtypes = [vp, vp, i32, i32]
    lib.qpb_add_generation.argtypes = [vp, dbl, dbl]
    lib.qpb_add_generation_program.argtypes = [vp, dbl, dbl]
    lib.qpb_add_generation_array.argtypes = [vp, dbl, vp]
    lib.qpb_generation_status.argtypes = [vp, C.POINTER(i32)]
    lib.qpb_set_stream.argtypes = [vp, vp]
    lib.qpb_measure_fp64.argtypes = [C.c_int, C.POINTER(dbl)]
    lib.qpb_measure_copy.argtypes = [C.c_int, i64, C.POINTER(dbl)]
    for name in EXPORTED:
        fn = getattr(lib, name)
        if name not in ("qpb_last_error", "qpb_destroy"):
            fn.restype = C.c_int
    if lib.qpb_abi_version() != ABI_VERSION:
        raise RuntimeError("libqpb.so ABI version does not match the Python binding; rebuild the library")
    _lib = lib
    return lib


# bytes that crossed PCIe through this binding, counted at the copy calls (bench.py reports them per step)
transfer_stats = {"h2d": 0, "d2h": 0}


def _up(*arrays):
    for a in arrays:
        if a is not None:
            transfer_stats["h2d"] += int(a.nbytes)


def _down(*arrays):
    for a in arrays:
        if a is not None:
            transfer_stats["d2h"] += int(a.nbytes)


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and out.shape != tuple(shape):
        raise ValueError(f"expected array of shape {tuple(shape)}, got {out.shape}")
    return out


class Context:
    """RAII wrapper of a qpb_ctx handle."""

    def __init__(self, *, ny, nx, ne, nw, ncell, ngap=1, flags=0, dx=1.0, dE=1.0, device=0, diff_tol=0.0,
                 pauli_floor=1e-18):
        self.lib = load_library()
        self.cfg = Config(ABI_VERSION, int(device), int(ny), int(nx), int(ne), int(nw), int(ncell), int(ngap),
                          int(flags), 0, float(dx), float(dE), float(diff_tol), float(pauli_floor))
        self.handle = C.c_void_p()
        self._check(self.lib.qpb_create(C.byref(self.cfg), C.byref(self.handle)))
        self.ne, self.nw, self.ncell, self.ny, self.nx = int(ne), int(nw), int(ncell), int(ny), int(nx)
        self.flags = int(flags)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise QpbError(rc, self.lib.qpb_last_error().decode("utf-8", "replace"))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.qpb_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads -------------------------------------------------------------------------------------
    def upload_geometry(self, mask, bcx=None, bcy=None, source=None):
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        if m.shape != (self.ny, self.nx):
            raise ValueError("mask shape does not match the context")
        arrs = [None if a is None else _f64(a, (self.ny, self.nx)) for a in (bcx, bcy, source)]
        _up(m, *arrs)
        self._check(self.lib.qpb_upload_geometry(self.handle, _ptr(m), *[_ptr(a) for a in arrs]))

    def upload_diffusion(self, D):
        shape = (self.ne, self.ncell) if self.flags & F_VARIABLE_D else (self.ne,)
        d = _f64(D, shape)
        _up(d)
        self._check(self.lib.qpb_upload_diffusion(self.handle, _ptr(d)))

    def prepare_diffusion(self, slot: int, dt: float):
        self._check(self.lib.qpb_prepare_diffusion(self.handle, int(slot), float(dt)))

    def upload_collision(self, K_r0, K_s0, rho, gap_id, idx_diff, idx_sum, sign):
        ng, ne = self.cfg.ngap, self.ne
        kr = None if K_r0 is None else _f64(K_r0).reshape(ng, ne, ne)
        ks = None if K_s0 is None else _f64(K_s0).reshape(ng, ne, ne)
        rh = _f64(rho).reshape(ng, ne)
        gid = None if gap_id is None else np.ascontiguousarray(gap_id, dtype=np.int32).reshape(self.ncell)
        idd = None if idx_diff is None else np.ascontiguousarray(idx_diff, dtype=np.int32).reshape(ne, ne)
        ids = None if idx_sum is None else np.ascontiguousarray(idx_sum, dtype=np.int32).reshape(ne, ne)
        sg = None if sign is None else np.ascontiguousarray(sign, dtype=np.int8).reshape(ne, ne)
        _up(kr, ks, rh, gid, idd, ids, sg)
        self._check(self.lib.qpb_upload_collision(self.handle, _ptr(kr), _ptr(ks), _ptr(rh), _ptr(gid), _ptr(idd),
                                                  _ptr(ids), _ptr(sg)))

    # ---- state ---------------------------------------------------------------------------------------
    def set_state(self, n, n_ph=None):
        a = _f64(n, (self.ne, self.ncell))
        p = None if (n_ph is None or self.nw == 0) else _f64(n_ph, (self.nw, self.ncell))
        _up(a, p)
        self._check(self.lib.qpb_set_state(self.handle, _ptr(a), _ptr(p)))

    def set_state_uniform_phonons(self, n, n_ph_bins):
        """State upload when every cell starts from the same phonon occupations (one value per phonon bin)."""
        a = _f64(n, (self.ne, self.ncell))
        p = None if self.nw == 0 else _f64(n_ph_bins, (self.nw,))
        _up(a, p)
        self._check(self.lib.qpb_set_state_uniform_phonons(self.handle, _ptr(a), _ptr(p)))

    def set_state_separable(self, weights, spatial, n_ph_bins=None):
        """Default initial state ``state[i] = spatial * weights[i]`` (solver.py:1281-1283) formed on the device from its
        factors, with the bath phonon occupations broadcast over the cells (NE + N + Nw doubles uploaded)."""
        w = _f64(weights, (self.ne,))
        sp = _f64(spatial, (self.ncell,))
        p = None if self.nw == 0 else _f64(n_ph_bins, (self.nw,))
        _up(w, sp, p)
        self._check(self.lib.qpb_set_state_separable(self.handle, _ptr(w), _ptr(sp), _ptr(p)))

    def get_state(self, want_phonons=True, want_qp=True):
        n = np.empty((self.ne, self.ncell)) if want_qp else None
        p = np.empty((self.nw, self.ncell)) if (want_phonons and self.nw > 0) else None
        self._check(self.lib.qpb_get_state(self.handle, _ptr(n), _ptr(p)))
        _down(n, p)
        return n, p

    def get_integrated(self):
        out = np.empty(self.ncell)
        self._check(self.lib.qpb_get_integrated(self.handle, _ptr(out)))
        _down(out)
        return out

    def get_frames(self):
        """The NE stored energy frames of one snapshot, (NE, ny, nx) with NaN outside the mask."""
        out = np.empty((self.ne, self.ny, self.nx))
        self._check(self.lib.qpb_get_frames(self.handle, _ptr(out)))
        _down(out)
        return out

    def frames_snapshot(self):
        """Assemble the NE frames of the current state into a device buffer of their own (stream ordered, returns at
        once).  :meth:`frames_download` fetches them; one snapshot is outstanding at a time."""
        self._check(self.lib.qpb_frames_snapshot(self.handle))

    def frames_download(self, out=None):
        """Copy the last snapshot to the host, (NE, ny, nx) with NaN outside the mask.  Safe to call from another
        thread while this context keeps stepping (ctypes releases the GIL)."""
        if out is None:
            out = np.empty((self.ne, self.ny, self.nx))
        self._check(self.lib.qpb_frames_download(self.handle, _ptr(out)))
        _down(out)
        return out

    # ---- stepping ------------------------------------------------------------------------------------
    def advance(self, nsteps, dt, slot=0, t_start=0.0, gen_mode=GEN_NONE, rate=0.0, pulse_start=0.0,
                pulse_duration=0.0, gen_array=None, want_pauli=False):
        g = Generation(int(gen_mode), 0, float(rate), float(pulse_start), float(pulse_duration), None)
        keep = None
        if gen_mode == GEN_ARRAY:
            keep = _f64(gen_array, (self.ne, self.ncell))
            _up(keep)
            g.array = keep.ctypes.data
        recs = (PauliRec * max(1, int(nsteps)))() if want_pauli else None
        rc = self.lib.qpb_advance(self.handle, int(nsteps), float(dt), int(slot), float(t_start), C.byref(g),
                                  C.cast(recs, C.c_void_p) if recs is not None else None)
        self._check(rc)
        if recs is None:
            return None
        transfer_stats["d2h"] += 24 * int(nsteps)
        return [(recs[k].max_occ, recs[k].max_index, recs[k].forbidden) for k in range(int(nsteps))]

    def upload_generation_program(self, program, E_bins, cell_x, cell_y):
        """A custom generation body as a postfix program ``[(op, value), ...]`` (userexpr.compile_program) with the
        inputs it is evaluated on: E per bin, normalised x / y per cell.  Afterwards gen_mode=GEN_PROGRAM."""
        ops = (GenOp * len(program))(*[GenOp(int(op), 0, float(v)) for op, v in program])
        e, x, y = _f64(E_bins, (self.ne,)), _f64(cell_x, (self.ncell,)), _f64(cell_y, (self.ncell,))
        for a in (e, x, y):
            _up(a)
        self._check(self.lib.qpb_upload_generation_program(self.handle, ops, len(program), _ptr(e), _ptr(x), _ptr(y)))

    def eval_generation_program(self, t):
        out = np.empty((self.ne, self.ncell), dtype=np.float64)
        self._check(self.lib.qpb_eval_generation_program(self.handle, float(t), _ptr(out)))
        transfer_stats["d2h"] += out.nbytes
        return out

    def collide(self, dt):
        self._check(self.lib.qpb_collide(self.handle, float(dt)))

    # ---- layout exchange fused into the collision kernel (multi-GPU) ----------------------------------
    def set_exchange(self, peer_ptrs, peer_ncd, bin_owner, bin_row, cell_dense):
        n = len(peer_ptrs)
        arr = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        own = np.ascontiguousarray(bin_owner, dtype=np.int16).reshape(self.ne)
        row = np.ascontiguousarray(bin_row, dtype=np.int16).reshape(self.ne)
        cd = np.ascontiguousarray(cell_dense, dtype=np.int32).reshape(self.ncell)
        self._check(self.lib.qpb_set_exchange(self.handle, n, arr, int(peer_ncd), _ptr(own), _ptr(row), _ptr(cd)))

    def collide_exchange(self, dt, mode):
        self._check(self.lib.qpb_collide_exchange(self.handle, float(dt), int(mode)))

    def ipc_export(self, which=0) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.qpb_ipc_export(self.handle, int(which), C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def euler_step(self, kind: int, K, vec, dt: float):
        """Fixed-bath forward-Euler form on the context's state: kind 1 scattering (K_s, rho), 2 recombination
        (K_r, G_therm)."""
        k = _f64(K, (self.ne, self.ne))
        v = _f64(vec, (self.ne,))
        self._check(self.lib.qpb_euler_step(self.handle, int(kind), _ptr(k), _ptr(v), float(dt)))

    def diffuse(self, slot=0):
        self._check(self.lib.qpb_diffuse(self.handle, int(slot)))

    def pauli(self):
        r = PauliRec()
        self._check(self.lib.qpb_pauli(self.handle, C.byref(r)))
        return r.max_occ, r.max_index, r.forbidden

    def pauli_record(self, slot: int):
        self._check(self.lib.qpb_pauli_record(self.handle, int(slot)))

    def pauli_fetch(self, count: int):
        recs = (PauliRec * max(1, int(count)))()
        self._check(self.lib.qpb_pauli_fetch(self.handle, int(count), C.cast(recs, C.c_void_p)))
        return [(recs[k].max_occ, recs[k].max_index, recs[k].forbidden) for k in range(int(count))]

    def diag(self) -> dict:
        d = Diag()
        self._check(self.lib.qpb_get_diag(self.handle, C.byref(d)))
        return {name: getattr(d, name) for name, _ in Diag._fields_}

    def synchronize(self):
        self._check(self.lib.qpb_synchronize(self.handle))

    def enable_timers(self, on=True):
        self._check(self.lib.qpb_enable_timers(self.handle, 1 if on else 0))

    def reset_timers(self):
        self._check(self.lib.qpb_reset_timers(self.handle))

    def timer(self, which: int):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.qpb_get_timer(self.handle, int(which), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def scatter_block(self, dev_ptr: int, cell0: int, count: int):
        self._check(self.lib.qpb_scatter_block(self.handle, C.c_void_p(int(dev_ptr)), int(cell0), int(count)))

    def gather_block(self, dev_ptr: int, cell0: int, count: int):
        self._check(self.lib.qpb_gather_block(self.handle, C.c_void_p(int(dev_ptr)), int(cell0), int(count)))

    def add_generation(self, scale: float, rate: float):
        self._check(self.lib.qpb_add_generation(self.handle, float(scale), float(rate)))

    def add_generation_program(self, scale: float, t: float):
        self._check(self.lib.qpb_add_generation_program(self.handle, float(scale), float(t)))

    def add_generation_array(self, scale: float, array=None):
        """state += scale * g with g[ne][ncell] a host array (kept on the device), or the array of the last call."""
        if array is None:
            self._check(self.lib.qpb_add_generation_array(self.handle, float(scale), None))
            return
        a = _f64(array, (self.ne, self.ncell))
        _up(a)
        self._check(self.lib.qpb_add_generation_array(self.handle, float(scale), _ptr(a)))

    def generation_status(self) -> int:
        v = C.c_int32(0)
        self._check(self.lib.qpb_generation_status(self.handle, C.byref(v)))
        return int(v.value)

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.qpb_set_stream(self.handle, C.c_void_p(int(cuda_stream)) if cuda_stream else None))

    def device_ptr(self, which: int):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.qpb_device_ptr(self.handle, int(which), C.byref(p), C.byref(n)))
        return p.value, n.value


def ipc_open(device: int, handle: bytes) -> int:
    lib = load_library()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    out = C.c_void_p()
    rc = lib.qpb_ipc_open(int(device), C.cast(buf, C.c_void_p), C.byref(out))
    if rc != 0:
        raise QpbError(rc, lib.qpb_last_error().decode())
    return int(out.value)


def ipc_close(device: int, ptr: int) -> None:
    lib = load_library()
    rc = lib.qpb_ipc_close(int(device), C.c_void_p(int(ptr)))
    if rc != 0:
        raise QpbError(rc, lib.qpb_last_error().decode())


def measure_fp64_tflops(device: int = 0) -> float:
    lib = load_library()
    out = C.c_double()
    rc = lib.qpb_measure_fp64(int(device), C.byref(out))
    if rc != 0:
        raise QpbError(rc, lib.qpb_last_error().decode())
    return out.value


def measure_copy_gbs(device: int = 0, nbytes: int = 1 << 30) -> float:
    lib = load_library()
    out = C.c_double()
    rc = lib.qpb_measure_copy(int(device), int(nbytes), C.byref(out))
    if rc != 0:
        raise QpbError(rc, lib.qpb_last_error().decode())
    return out.value
