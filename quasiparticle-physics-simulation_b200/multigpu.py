"""One-box multi-GPU driver of the time-stepping hot path.

The two halves of a time step parallelise on different axes: the Crank-Nicolson diffusion solve is independent
per energy bin (qpsim/solver.py:1443-1452), the collision update is independent per cell (solver.py:814-831).
One process per GPU (torchrun); rank g owns

  * diffusion layout:  bins  g, g+P, g+2P, ...  (interleaved: the sweep count of a bin grows with D(E), so a
    contiguous split would give the low-energy rank the cheap bins) of ALL cells — a qpb context with the full
    mask and NE_g bins;
  * collision layout:  a contiguous range of the compressed cell ordering, ALL bins, plus the phonon state of
    those cells (phonons never move) — a qpb context whose geometry is a 1 x N_g strip.

Per step:  g_ext, C(dt/2) [cells]  ->  all-to-all  ->  D(dt) [bins]  ->  all-to-all  ->  C(dt/2), Pauli [cells].
The all-to-all is the only data-path collective; it moves 8*NE*N*(P-1)/P^2 bytes out of every GPU per exchange
over NVLink (NCCL ``all_to_all_single`` with split sizes).  Packing is done by the library's own block kernels
(``qpb_gather_block`` / ``qpb_scatter_block``) and one row permutation; every kernel and the collective are
ordered on one CUDA stream (``qpb_set_stream``), so the host does not synchronise between stages.

``ShardPlan`` and ``ShardedStepper`` contain no device code and are exercised on CPU tensors over gloo by
tests/test_multigpu_gloo.py with stand-in stages; ``DeviceStages`` binds them to libqpb.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import capi


# ------------------------------------------------------------------------------------------------------------
# partition
# ------------------------------------------------------------------------------------------------------------
class ShardPlan:
    """Who owns which bins (diffusion layout) and which cells (collision layout)."""

    def __init__(self, ne: int, ncell: int, world: int, rank: int, interleave: bool = True):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad world/rank")
        if ne < world or ncell < world:
            raise ValueError(f"cannot shard {ne} bins / {ncell} cells over {world} ranks")
        self.ne, self.ncell, self.world, self.rank, self.interleave = int(ne), int(ncell), int(world), int(rank), interleave
        if interleave:
            self._bins = [np.arange(g, ne, world, dtype=np.int64) for g in range(world)]
        else:
            cuts = [(ne * g) // world for g in range(world + 1)]
            self._bins = [np.arange(cuts[g], cuts[g + 1], dtype=np.int64) for g in range(world)]
        self._cuts = [(ncell * g) // world for g in range(world + 1)]
        # rows of the cell-sharded state ordered by destination rank
        self.perm = np.concatenate(self._bins)

    def bins(self, g: int | None = None) -> np.ndarray:
        return self._bins[self.rank if g is None else g]

    def cells(self, g: int | None = None) -> tuple[int, int]:
        g = self.rank if g is None else g
        return self._cuts[g], self._cuts[g + 1]

    def ncells(self, g: int | None = None) -> int:
        c0, c1 = self.cells(g)
        return c1 - c0

    def nbins(self, g: int | None = None) -> int:
        return int(self.bins(g).size)

    # element counts of the all-to-all: to_bins sends (bins of g) x (my cells) to g, receives (my bins) x (cells of g)
    def splits_to_bins(self):
        send = [self.nbins(g) * self.ncells() for g in range(self.world)]
        recv = [self.nbins() * self.ncells(g) for g in range(self.world)]
        return send, recv

    def exchange_bytes(self) -> int:
        """Bytes this rank sends over the fabric per exchange (its own block stays local)."""
        send, _ = self.splits_to_bins()
        return 8 * (sum(send) - send[self.rank])


# ------------------------------------------------------------------------------------------------------------
# stepping
# ------------------------------------------------------------------------------------------------------------
class ShardedStepper:
    """The loop body of solver.py:1454-1478 on sharded state.  `stages` provides the local work:

        coll_state            torch tensor [NE, N_g], the live cell-sharded quasiparticle state
        collide(dt), add_generation(scale, rate), diffuse(slot), pauli() -> (max_occ, flat_index, forbidden)
        scatter_block(block[NE_g, count], cell0, count)   write cells cell0.. of my bins into the diffusion state
        gather_block(block[NE_g, count], cell0, count)    read them back
    """

    def __init__(self, plan: ShardPlan, stages: Any, *, diffusion: bool, collisions: bool, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.plan, self.stages, self.group = plan, stages, group
        self.diffusion, self.collisions = bool(diffusion), bool(collisions)
        st = stages.coll_state
        self.perm_t = torch.as_tensor(plan.perm, device=st.device)
        nloc = plan.ncells()
        send, recv = plan.splits_to_bins()
        self.send_elems, self.recv_elems = send, recv
        # cell-layout side buffer (rows grouped by destination) and bin-layout side buffer (blocks by source)
        self.buf_c = torch.empty((plan.ne, nloc), dtype=st.dtype, device=st.device)
        self.buf_d = torch.empty(sum(recv), dtype=st.dtype, device=st.device)
        self.exchanges = 0
        # fused exchange (DeviceStages.enable_fused_exchange): the collision kernel stores into / loads from the peers'
        # diffusion states; what is left between the stages is a barrier that orders writers and readers
        self.fused = bool(getattr(stages, "fused", False)) and self.diffusion and self.collisions
        self._flag = torch.zeros(1, dtype=torch.float32, device=st.device) if self.fused else None

    def _rank_barrier(self):
        if self.plan.world > 1:
            self.dist.all_reduce(self._flag, group=self.group)   # stream ordered, no host synchronisation

    # ---- layout exchange ------------------------------------------------------------------------------
    def to_bins(self):
        t, p = self.torch, self.plan
        st = self.stages.coll_state
        if p.interleave:
            t.index_select(st, 0, self.perm_t, out=self.buf_c)
            src = self.buf_c
        else:
            src = st
        if p.world > 1:
            self.dist.all_to_all_single(self.buf_d, src.reshape(-1), self.recv_elems, self.send_elems, group=self.group)
        else:
            self.buf_d.copy_(src.reshape(-1))
        off = 0
        for g in range(p.world):
            c0, c1 = p.cells(g)
            blk = self.buf_d[off:off + self.recv_elems[g]].view(p.nbins(), c1 - c0)
            self.stages.scatter_block(blk, c0, c1 - c0)
            off += self.recv_elems[g]
        self.exchanges += 1

    def to_cells(self):
        t, p = self.torch, self.plan
        off = 0
        for g in range(p.world):
            c0, c1 = p.cells(g)
            blk = self.buf_d[off:off + self.recv_elems[g]].view(p.nbins(), c1 - c0)
            self.stages.gather_block(blk, c0, c1 - c0)
            off += self.recv_elems[g]
        st = self.stages.coll_state
        dst = self.buf_c if p.interleave else st
        if p.world > 1:
            self.dist.all_to_all_single(dst.reshape(-1), self.buf_d, self.send_elems, self.recv_elems, group=self.group)
        else:
            dst.reshape(-1).copy_(self.buf_d)
        if p.interleave:
            st.index_copy_(0, self.perm_t, self.buf_c)
        self.exchanges += 1

    # ---- one time step --------------------------------------------------------------------------------
    def step(self, dt: float, slot: int = 0, gen_rate: float | None = None, want_pauli: bool = False,
             pauli_slot: int | None = None):
        """One time step.  want_pauli returns this rank's record (one host synchronisation); pauli_slot instead
        records it on the device (stages.pauli_record / pauli_fetch), so a batch of steps needs no extra sync."""
        s = self.stages
        if gen_rate is not None:
            s.add_generation(dt, gen_rate)          # solver.py:1459-1464
        if self.collisions and self.diffusion and self.fused:
            s.collide_exchange(0.5 * dt, 1)         # collide, then store into the owners' diffusion states
            self._rank_barrier()                    # every rank's cells have arrived
            s.diffuse(slot)
            self._rank_barrier()                    # every rank's bins are solved
            s.collide_exchange(0.5 * dt, 2)         # load my cells from the owners, then collide
            self._rank_barrier()                    # nobody still reads the diffusion states the next step overwrites
            self.exchanges += 2
        elif self.collisions and self.diffusion:    # solver.py:1469-1472
            s.collide(0.5 * dt)
            self.to_bins()
            s.diffuse(slot)
            self.to_cells()
            s.collide(0.5 * dt)
        else:                                       # solver.py:1474-1475
            if self.collisions:
                s.collide(dt)
            if self.diffusion:
                self.to_bins()
                s.diffuse(slot)
                self.to_cells()
        if pauli_slot is not None:
            s.pauli_record(pauli_slot)
            return None
        return s.pauli() if want_pauli else None

    # ---- collectives over small host values -----------------------------------------------------------
    def merge_pauli(self, recs: list[tuple[float, int, int]]):
        """Combine per-rank, per-step records (indices local: i*N_g + cell) into global records
        (i*N + cell, first maximum / first forbidden cell in the reference's np.argmax order)."""
        t, p = self.torch, self.plan
        k = len(recs)
        dev = self.stages.coll_state.device
        mine = t.tensor([[r[0], float(r[1]), float(r[2])] for r in recs], dtype=t.float64, device=dev).reshape(k, 3)
        if p.world > 1:
            parts = [t.empty_like(mine) for _ in range(p.world)]
            self.dist.all_gather(parts, mine, group=self.group)
        else:
            parts = [mine]
        parts = [x.cpu().numpy() for x in parts]
        out = []
        for s_ in range(k):
            best, best_idx, forb = -np.inf, -1, -1
            for g in range(p.world):
                c0, _ = p.cells(g)
                ng = p.ncells(g)
                mo, li, lf = parts[g][s_]
                i, q = divmod(int(li), ng)
                gi = i * p.ncell + c0 + q
                if mo > best or (mo == best and gi < best_idx):
                    best, best_idx = float(mo), gi
                if lf >= 0:
                    i, q = divmod(int(lf), ng)
                    gf = i * p.ncell + c0 + q
                    forb = gf if forb < 0 else min(forb, gf)
            out.append((best, best_idx, forb))
        return out

    def gather_state(self):
        """Full [NE, N] quasiparticle state on every rank (host numpy), for stored frames and tests."""
        t, p = self.torch, self.plan
        st = self.stages.coll_state
        if p.world == 1:
            return st.cpu().numpy().copy()
        nmax = max(p.ncells(g) for g in range(p.world))
        pad = t.zeros((p.ne, nmax), dtype=st.dtype, device=st.device)
        pad[:, :p.ncells()] = st
        parts = [t.empty_like(pad) for _ in range(p.world)]
        self.dist.all_gather(parts, pad, group=self.group)
        return np.concatenate([parts[g][:, :p.ncells(g)].cpu().numpy() for g in range(p.world)], axis=1)


# ------------------------------------------------------------------------------------------------------------
# device binding
# ------------------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of a libqpb device allocation for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape: tuple[int, ...]):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 2, "strides": None}


@dataclass
class ShardedProblem:
    """Host-side tables of one run (what solver.run_2d_crank_nicolson builds before its loop)."""
    mask: np.ndarray
    bcx: np.ndarray | None
    bcy: np.ndarray | None
    src: np.ndarray | None
    dx: float
    dE: float
    D: np.ndarray                 # [NE] or [NE, N] (variable)
    variable_D: bool
    rho: np.ndarray               # [ngap, NE]
    Kr: np.ndarray | None         # [ngap, NE, NE]
    Ks: np.ndarray | None
    gap_id: np.ndarray | None     # [N]
    idx_diff: np.ndarray | None
    idx_sum: np.ndarray | None
    sign: np.ndarray | None
    nw: int
    state: np.ndarray | None      # [NE, N] (None: only this rank's slice is held, see state_local)
    phonons: np.ndarray | None    # [Nw, N]
    diffusion: bool = True
    scattering: bool = True
    recombination: bool = True
    freeze_phonons: bool = False
    pauli_floor: float = 1e-18
    diff_tol: float = 0.0
    # large grids: a rank never materialises arrays of the whole run
    state_local: np.ndarray | None = None    # [NE, N_g] this rank's cells only (used when state is None)
    phonon_bins: np.ndarray | None = None    # [Nw] the same occupations in every cell (used when phonons is None)
    # default initial state as its factors, state[i] = spatial * weights[i] (solver.py:1281-1283): formed on the device
    weights: np.ndarray | None = None        # [NE]
    spatial_local: np.ndarray | None = None  # [N_g] this rank's cells

    def local_state(self, c0: int, c1: int):
        return self.state[:, c0:c1] if self.state is not None else self.state_local


class DeviceStages:
    """Two libqpb contexts on this rank's GPU: the bin-sharded diffusion context and the cell-sharded collision
    context, both enqueueing on one torch CUDA stream."""

    def __init__(self, plan: ShardPlan, prob: ShardedProblem, device: int, dt: float, remainder_dt: float = 0.0):
        import torch

        self.torch, self.plan, self.device = torch, plan, int(device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        ne, n = plan.ne, plan.ncell
        c0, c1 = plan.cells()
        nloc = c1 - c0
        bins = plan.bins()
        ny, nx = prob.mask.shape
        coll = prob.scattering or prob.recombination
        self.ctx_d = None
        if prob.diffusion:
            fl = capi.F_DIFFUSION | (capi.F_VARIABLE_D if prob.variable_D else 0)
            self.ctx_d = capi.Context(ny=ny, nx=nx, ne=bins.size, nw=0, ncell=n, flags=fl, dx=prob.dx, dE=prob.dE,
                                      device=self.device, diff_tol=prob.diff_tol)
            self.ctx_d.upload_geometry(prob.mask, prob.bcx, prob.bcy, prob.src)
            self.ctx_d.upload_diffusion(prob.D[bins])
            self.ctx_d.prepare_diffusion(0, dt)
            if remainder_dt > 0.0:
                self.ctx_d.prepare_diffusion(1, remainder_dt)
            self.ctx_d.set_stream(self.stream.cuda_stream)
        fl = capi.F_PAULI
        fl |= capi.F_SCATTERING if prob.scattering else 0
        fl |= capi.F_RECOMBINATION if prob.recombination else 0
        fl |= capi.F_FREEZE_PHONONS if prob.freeze_phonons else 0
        ngap = prob.rho.shape[0]
        self.ctx_c = capi.Context(ny=1, nx=nloc, ne=ne, nw=prob.nw if coll else 0, ncell=nloc, ngap=ngap, flags=fl,
                                  dx=prob.dx, dE=prob.dE, device=self.device, pauli_floor=prob.pauli_floor)
        self.ctx_c.upload_geometry(np.ones((1, nloc), dtype=np.uint8))
        gid = None if prob.gap_id is None else prob.gap_id[c0:c1]
        self.ctx_c.upload_collision(prob.Kr, prob.Ks, prob.rho, gid, prob.idx_diff if coll else None,
                                    prob.idx_sum if coll else None, prob.sign if coll else None)
        self.load_state(prob)
        self.ctx_c.set_stream(self.stream.cuda_stream)
        ptr, _ = self.ctx_c.device_ptr(0)
        # the strip context's dense state IS the compact [NE, N_g] array
        self.coll_state = torch.as_tensor(_DevArray(ptr, (ne, nloc)), device=f"cuda:{self.device}")
        self.collisions = coll

    def load_state(self, prob: "ShardedProblem"):
        """Upload this rank's cells (host -> device) into the collision layout."""
        c0, c1 = self.plan.cells()
        if (prob.state is None and prob.state_local is None and prob.weights is not None
                and prob.spatial_local is not None):
            # the outer product never exists on the host: NE + N_g + Nw doubles travel instead of NE x N_g
            if not self.collisions_on(prob):
                self.ctx_c.set_state_separable(prob.weights, prob.spatial_local, None)
                return
            if prob.phonons is None and prob.phonon_bins is not None:
                self.ctx_c.set_state_separable(prob.weights, prob.spatial_local, prob.phonon_bins)
                return
            prob.state_local = np.ascontiguousarray(prob.weights[:, None] * prob.spatial_local[None, :])
        if prob.phonons is None and prob.phonon_bins is not None and self.collisions_on(prob):
            self.ctx_c.set_state_uniform_phonons(prob.local_state(c0, c1), prob.phonon_bins)
        else:
            self.ctx_c.set_state(prob.local_state(c0, c1), None if prob.phonons is None else prob.phonons[:, c0:c1])

    @staticmethod
    def collisions_on(prob: "ShardedProblem") -> bool:
        return bool(prob.scattering or prob.recombination)

    def close(self):
        for ptr in getattr(self, "_ipc_open", []):
            try:
                capi.ipc_close(self.device, ptr)
            except Exception:
                pass
        self._ipc_open = []
        for c in (self.ctx_d, self.ctx_c):
            if c is not None:
                c.close()

    def enable_fused_exchange(self, prob: "ShardedProblem", group=None) -> bool:
        """Map every rank's diffusion state into this process (cudaIpc over NVLink) and route the collision kernel's
        stores / staging loads through it (qpb_set_exchange).  Returns False (and leaves the all-to-all path in place)
        when the run has no diffusion or collisions, or the structured collision kernel is not the one in use."""
        import torch.distributed as dist

        self.fused = False
        self._ipc_open = []
        if self.ctx_d is None or not self.collisions or os.environ.get("QPB_NO_FUSED_EXCHANGE") == "1":
            return False
        plan = self.plan
        own_ptr, _ = self.ctx_d.device_ptr(0)
        peers = [own_ptr] * plan.world
        if plan.world > 1:
            handles = [None] * plan.world
            dist.all_gather_object(handles, self.ctx_d.ipc_export(0), group=group)
            for r in range(plan.world):
                if r != plan.rank:
                    peers[r] = capi.ipc_open(self.device, handles[r])
                    self._ipc_open.append(peers[r])
        owner = np.empty(plan.ne, dtype=np.int16)
        row = np.empty(plan.ne, dtype=np.int16)
        for g in range(plan.world):
            b = plan.bins(g)
            owner[b] = g
            row[b] = np.arange(b.size)
        c0, c1 = plan.cells()
        cell_dense = np.flatnonzero(np.asarray(prob.mask, dtype=bool).ravel())[c0:c1]
        ok = True
        try:
            self.ctx_c.set_exchange(peers, prob.mask.size, owner, row, cell_dense)
        except capi.QpbError:
            ok = False
        if plan.world > 1:   # all ranks or none
            flags = [None] * plan.world
            dist.all_gather_object(flags, ok, group=group)
            ok = all(flags)
        self.fused = ok
        return ok

    def collide(self, dt):
        self.ctx_c.collide(dt)

    def collide_exchange(self, dt, mode):
        self.ctx_c.collide_exchange(dt, mode)

    def add_generation(self, scale, rate):
        self.ctx_c.add_generation(scale, rate)

    def upload_generation_program(self, program, E_bins, x, y):
        self.ctx_c.upload_generation_program(program, E_bins, x, y)

    def add_generation_program(self, scale, t):
        self.ctx_c.add_generation_program(scale, t)

    def add_generation_array(self, scale, array=None):
        self.ctx_c.add_generation_array(scale, array)

    def generation_status(self):
        return self.ctx_c.generation_status()

    def diffuse(self, slot):
        self.ctx_d.diffuse(slot)

    def pauli(self):
        return self.ctx_c.pauli()

    def pauli_record(self, slot):
        self.ctx_c.pauli_record(slot)

    def pauli_fetch(self, count):
        return self.ctx_c.pauli_fetch(count)

    def scatter_block(self, block, cell0, count):
        self.ctx_d.scatter_block(block.data_ptr(), cell0, count)

    def gather_block(self, block, cell0, count):
        self.ctx_d.gather_block(block.data_ptr(), cell0, count)

    def launches(self) -> int:
        n = self.ctx_c.diag()["kernel_launches"]
        if self.ctx_d is not None:
            n += self.ctx_d.diag()["kernel_launches"]
        return int(n)


def init_process_group(backend: str | None = None):
    """torchrun rendezvous (env://, 127.0.0.1); returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device(f"cuda:{local}")
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


# ------------------------------------------------------------------------------------------------------------
# run_2d_crank_nicolson(..., devices=[...]): the sharded loop behind the drop-in
# ------------------------------------------------------------------------------------------------------------
def _gather_rows(local, plan: ShardPlan, dst: int = 0, group=None):
    """[rows, N_g] device tensors of all ranks -> host [rows, N] on rank `dst` (None elsewhere).  One block travels at
    a time, so the receiver never holds more than one rank's slice on the device."""
    import torch
    import torch.distributed as dist

    if plan.world == 1:
        return local.cpu().numpy().copy()
    rows = local.shape[0]
    if plan.rank != dst:
        dist.send(local.contiguous(), dst=dst, group=group)
        return None
    out = np.empty((rows, plan.ncell))
    for g in range(plan.world):
        c0, c1 = plan.cells(g)
        if g == dst:
            out[:, c0:c1] = local.cpu().numpy()
        else:
            buf = torch.empty((rows, c1 - c0), dtype=local.dtype, device=local.device)
            dist.recv(buf, src=g, group=group)
            out[:, c0:c1] = buf.cpu().numpy()
    return out


def _run_spmd(su: dict, progress_callback=None) -> dict:
    """The loop of solver.run_2d_crank_nicolson (energy-resolved mode) on the ranks of the current process group.
    Called by every rank with identical arguments; rank r drives su["devices"][r]."""
    import torch
    import torch.distributed as dist

    from .solver import reconstruct_field, _callback
    from . import physics

    rank, world = dist.get_rank(), dist.get_world_size()
    device = int(su["devices"][rank])
    mask, ne, n, nw = su["mask"], su["ne"], su["n"], su["nw"]
    gen = su["generation"]
    gmode = "none" if gen is None else gen.mode.strip().lower()
    plan = ShardPlan(ne, n, world, rank, interleave=True)
    c0, c1 = plan.cells()
    custom = None
    if gmode == "custom":
        # evaluate_external_generation (solver.py:918-962) on this rank's cells: a time-dependent body as a device program,
        # a time-independent one as an array that is evaluated and uploaded once; a body without a per-value meaning on
        # the host, one slice per step
        from . import userexpr
        custom = userexpr.CustomGeneration(gen, su["E_bins"], mask)
        if os.environ.get("QPB_NO_GEN_PROGRAM", "0") == "1":
            custom.program = None
    coll = su["scattering"] or su["recombination"]
    separable = su["state"] is None
    state_local = su["state"][:, c0:c1] if not separable else None
    prob = ShardedProblem(
        mask=mask, bcx=su["bcx"], bcy=su["bcy"], src=su["src"], dx=su["dx"], dE=su["dE"], D=np.asarray(su["D"]),
        variable_D=su["variable_D"], rho=su["rho"], Kr=su["Kr"], Ks=su["Ks"], gap_id=su["gap_id"],
        idx_diff=su["idx_diff"], idx_sum=su["idx_sum"], sign=su["sign"], nw=nw, state=None,
        phonons=su["phonon_state"], diffusion=su["diffusion"], scattering=su["scattering"],
        recombination=su["recombination"], freeze_phonons=su["freeze_phonons"], pauli_floor=su["pauli_floor"],
        diff_tol=su["diff_tol"], state_local=None if separable else np.ascontiguousarray(state_local),
        phonon_bins=su["phonon_bins"] if su["phonon_state"] is None else None,
        weights=np.asarray(su["weights"], dtype=float) if separable else None,
        spatial_local=np.ascontiguousarray(su["spatial"][c0:c1], dtype=float) if separable else None)
    dt, rem = su["dt"], su["remainder_dt"]
    import time
    t_enter = time.perf_counter()
    stages = DeviceStages(plan, prob, device, dt, rem)
    t_stages = time.perf_counter()
    policy = su["policy"]
    try:
        fused = stages.enable_fused_exchange(prob)
        with torch.cuda.stream(stages.stream):
            stepper = ShardedStepper(plan, stages, diffusion=su["diffusion"], collisions=coll)
            dev = stages.coll_state.device
            dE, dx = su["dE"], su["dx"]
            ph_view = None
            if coll and nw > 0:
                pptr, _ = stages.ctx_c.device_ptr(1)
                ph_view = torch.as_tensor(_DevArray(pptr, (nw, plan.ncells())), device=dev)
            widths = (physics.integration_widths_from_centers(su["omega_bins"], fallback_width=dE)
                      if su["want_phonon_history"] else None)

            def integrated_all():
                mine = torch.as_tensor(stages.ctx_c.get_integrated(), device=dev)
                if world == 1:
                    return mine.cpu().numpy()
                nmax = max(plan.ncells(g) for g in range(world))
                pad = torch.zeros(nmax, dtype=mine.dtype, device=dev)
                pad[:mine.numel()] = mine
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad)
                return np.concatenate([parts[g][:plan.ncells(g)].cpu().numpy() for g in range(world)])

            times, frames, mass, eframes = [0.0], [], [], []
            ph_frames, ph_eframes = [], []

            def store(t):
                integ = integrated_all()
                frame = reconstruct_field(mask, integ)
                frames.append(frame)
                mass.append(float(np.sum(integ) * dx * dx))
                if su["store_energy_frames"]:
                    full = _gather_rows(stages.coll_state, plan)
                    eframes.append(None if full is None else [reconstruct_field(mask, full[i]) for i in range(ne)])
                else:
                    eframes.append(None)
                if su["want_phonon_history"]:
                    if ph_view is not None:
                        ph = _gather_rows(ph_view, plan)
                    elif rank == 0:   # collisions off: the phonons stay what they were initialised to
                        ph = (np.asarray(su["phonon_state"], dtype=float) if su["phonon_state"] is not None
                              else su["phonon_bins"][:, None] * np.ones((1, n)))
                    else:
                        ph = None
                    if ph is not None:
                        ph_eframes.append([reconstruct_field(mask, ph[i]) for i in range(ph.shape[0])])
                        ph_frames.append(reconstruct_field(mask, np.sum(ph * widths[:, None], axis=0)))
                _callback(progress_callback, float(t), frame)

            t_ready = time.perf_counter()
            rec0 = stepper.merge_pauli([stages.pauli()])[0]
            policy.check(rec0, 0, 0.0)
            store(0.0)
            full_steps, total_steps, store_every = su["full_steps"], su["total_steps"], su["store_every"]
            t, step = 0.0, 0
            gen_uploads = 0
            if custom is not None and custom.program is not None:
                stages.upload_generation_program(custom.program, custom.E, custom.x[c0:c1], custom.y[c0:c1])
            while step < total_steps:
                nxt = min(((step // store_every) + 1) * store_every, total_steps)
                if nxt > full_steps and step < full_steps:
                    nxt = full_steps
                is_final = step >= full_steps
                h = rem if is_final else dt
                count = nxt - step
                tt = t
                for k in range(count):
                    rate = None
                    if custom is not None:
                        if custom.program is not None:
                            stages.add_generation_program(h, tt)
                        elif custom.time_dependent or gen_uploads == 0:
                            stages.add_generation_array(h, np.ascontiguousarray(custom(tt)[:, c0:c1]))
                            gen_uploads += 1
                        else:
                            stages.add_generation_array(h, None)
                    if gmode == "constant":
                        rate = float(gen.rate)
                    elif gmode == "pulse" and gen.pulse_start <= tt < gen.pulse_start + gen.pulse_duration:
                        rate = float(gen.pulse_rate)
                    stepper.step(h, 1 if is_final else 0, rate, pauli_slot=k)
                    tt += h
                merged = stepper.merge_pauli(stages.pauli_fetch(count))
                if custom is not None and custom.program is not None:
                    # the device-side checks of the body, agreed between the ranks (solver.py:954-962, same messages)
                    bad = stages.generation_status()
                    flag = torch.tensor([bad & 1, (bad >> 1) & 1], dtype=torch.int32, device=f"cuda:{device}")
                    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                    bad = int(flag[0].item()) | (int(flag[1].item()) << 1)
                    if bad & 1:
                        raise ValueError("External generation mode 'custom' produced non-finite values.")
                    if bad & 2:
                        raise ValueError("External generation mode 'custom' produced negative values. "
                                         "Generation rates must be non-negative.")
                for k in range(count):
                    policy.check(merged[k], step + k + 1, t + h)
                    t += h
                step = nxt
                if step % store_every == 0 or step == total_steps:
                    times.append(float(t))
                    store(t)
            info = dict(stages.ctx_c.diag())
            info.update(seconds={"contexts_tables_state": t_stages - t_enter, "exchange_setup": t_ready - t_stages,
                                 "loop_and_stores": time.perf_counter() - t_ready})
            info.update(generation_uploads=gen_uploads,
                        generation_on_device=bool(custom is not None and custom.program is not None))
            info.update(world=world, fused_exchange=bool(fused), exchanges=stepper.exchanges,
                        exchange_bytes_per_gpu=plan.exchange_bytes())
            if stages.ctx_d is not None:
                d = stages.ctx_d.diag()
                info.update({k: d[k] for k in ("sweeps", "bin_sweeps", "pr_iterations", "sweep_path", "commuting")})
                info["kernel_launches"] += d["kernel_launches"]
            torch.cuda.synchronize()
    finally:
        stages.close()
    hist = None
    if su["want_phonon_history"] and rank == 0:
        hist = {"phonon_frames": ph_frames, "phonon_energy_frames": ph_eframes,
                "phonon_energy_bins": np.asarray(su["omega_bins"], dtype=float).copy(),
                "phonon_metadata": {"mode": "dynamic_local_coupled", "field_units": "integrated_occupation",
                                    "energy_frame_units": "occupation"}}
    return {"times": times, "frames": frames, "mass": mass,
            "energy_frames": eframes if (rank == 0 and su["store_energy_frames"]) else [None] * len(times),
            "phonon_history": hist, "info": info}


def _spawn_worker(rank: int, world: int, setup: dict, port: int, outpath: str):
    import pickle

    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dev = int(setup["devices"][rank])
    torch.cuda.set_device(dev)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{dev}"))
    try:
        out = _run_spmd(setup, None)
        if rank == 0:
            with open(outpath, "wb") as f:
                pickle.dump(out, f, protocol=5)
    finally:
        dist.destroy_process_group()


def run_dropin(setup: dict, progress_callback=None) -> dict:
    """Entry point of run_2d_crank_nicolson(devices=[...]).  Inside an initialised process group (torchrun) the call
    is collective: every rank runs its share.  From a plain process one worker per device is spawned for the duration
    of the call (rendezvous on 127.0.0.1) and rank 0's result is handed back; progress callbacks are not forwarded
    across that boundary."""
    import torch.distributed as dist

    world = len(setup["devices"])
    if dist.is_available() and dist.is_initialized():
        if dist.get_world_size() != world:
            raise ValueError(f"devices lists {world} GPUs but the process group has {dist.get_world_size()} ranks")
        return _run_spmd(setup, progress_callback)
    import pickle
    import socket
    import tempfile

    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        outpath = os.path.join(tmp, "result.pkl")
        try:
            mp.spawn(_spawn_worker, args=(world, setup, port, outpath), nprocs=world, join=True)
        except mp.ProcessRaisedException as exc:
            # a ValueError of the loop (Pauli limits, custom generation checks) stays a ValueError for the caller, like
            # in the single-device path and in the reference
            last = str(exc).strip().splitlines()[-1]
            if last.startswith("ValueError: "):
                raise ValueError(last[len("ValueError: "):]) from None
            raise
        with open(outpath, "rb") as f:
            return pickle.load(f)


# ------------------------------------------------------------------------------------------------------------
# tilings of the weak-scaling workload (bench.py --workload c2 --gpus N)
# ------------------------------------------------------------------------------------------------------------
def weak_tiling(world: int) -> tuple[int, int]:
    """(tiles along y, tiles along x) of the per-GPU 256 x 256 mask: 1x1, 1x2, 2x2, 4x2.  Rows stay at most 512 cells
    long (whole-line x tiles); at 8 GPUs the columns are 1024 long and take the segmented y sweep, measured faster
    (1.68 ms per diffusion step on the full 1024 x 512 rectangle) than 512 x 1024 (2.14 ms)."""
    table = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (4, 2)}
    if world in table:
        return table[world]
    tx = 1
    while tx * 2 <= 2 and world % (tx * 2) == 0:
        tx *= 2
    return world // tx, tx
