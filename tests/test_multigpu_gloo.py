"""Host logic of the multi-GPU path on CPU: the partition, the all-to-all layout exchange and the sharded step
order, with world_size = 2 (and 3, uneven splits) over gloo.  The local stage work is done by the oracle here
(test stand-in for the two libqpb contexts), so the result must equal the single-process oracle run exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import helpers
import qpsim_b200 as Q
from oracle import qp_oracle as O
from qpsim_b200.multigpu import ShardPlan, ShardedStepper


@pytest.mark.parametrize("ne,ncell,world,interleave", [(8, 10, 2, True), (7, 11, 3, True), (16, 64, 4, False), (5, 9, 1, True)])
def test_shard_plan_partitions(ne, ncell, world, interleave):
    plans = [ShardPlan(ne, ncell, world, r, interleave) for r in range(world)]
    assert sorted(np.concatenate([p.bins() for p in plans]).tolist()) == list(range(ne))
    cuts = [p.cells() for p in plans]
    assert cuts[0][0] == 0 and cuts[-1][1] == ncell and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
    assert sorted(plans[0].perm.tolist()) == list(range(ne))
    # what r sends to g in to_bins is what g expects from r
    for r in range(world):
        send, _ = plans[r].splits_to_bins()
        for g in range(world):
            _, recv = plans[g].splits_to_bins()
            assert send[g] == recv[r]
    with pytest.raises(ValueError):
        ShardPlan(2, 100, 4, 0)


class OracleStages:
    """CPU stand-in of multigpu.DeviceStages: same interface, numpy arrays, arithmetic by the oracle."""

    def __init__(self, plan, case, tabs):
        c0, c1 = plan.cells()
        self.state_c = np.ascontiguousarray(tabs["state"][:, c0:c1])
        self.phon_c = np.ascontiguousarray(tabs["phonons"][:, c0:c1])
        self.coll_state = torch.from_numpy(self.state_c)
        self.state_d = np.zeros((plan.nbins(), plan.ncell))
        self.t = tabs
        mask = case["mask"]
        edges = Q.extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
        Dloc = tabs["D"][plan.bins()][:, None] * np.ones((1, plan.ncell))
        self.op = O.DiffusionCN(mask, edges, bcs, case["dx"], Dloc, case["dt"], False)
        self.rho_state = tabs["rho"][:, None] * np.ones((1, c1 - c0))

    def collide(self, dt):
        t = self.t
        O.collide(self.state_c, self.phon_c, t["Kr"], t["Ks"], t["rho"], t["idx_diff"], t["idx_sum"], t["sign"], t["dE"],
                  dt, recomb=True, scat=True)

    def add_generation(self, scale, rate):
        self.state_c += scale * rate

    def diffuse(self, slot):
        self.op.step(self.state_d)

    def pauli(self):
        mo, (i, q), forb = O.pauli_stats(self.state_c, self.rho_state)
        n = self.state_c.shape[1]
        return mo, i * n + q, -1 if forb is None else forb[0] * n + forb[1]

    def scatter_block(self, block, cell0, count):
        self.state_d[:, cell0:cell0 + count] = block.numpy()

    def gather_block(self, block, cell0, count):
        block.copy_(torch.from_numpy(np.ascontiguousarray(self.state_d[:, cell0:cell0 + count])))


def _tables(case):
    E, dE = Q.build_energy_grid(case["energy_gap"], case["energy_min_factor"], case["energy_max_factor"],
                                case["num_energy_bins"])
    rho = Q.density_of_states(E, case["energy_gap"], case["dynes_gamma"])
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    mask = case["mask"]
    n = int(mask.sum())
    w = rho / (np.sum(rho) * dE)
    return dict(E=E, dE=dE, rho=rho, Kr=Q.recombination_kernel_base(E, case["energy_gap"], case["tau_0"], case["T_c"]),
                Ks=Q.scattering_kernel_base(E, case["energy_gap"], case["tau_0"], case["T_c"]), idx_diff=idd, idx_sum=ids,
                sign=sg, state=w[:, None] * case["initial_field"][mask][None, :],
                phonons=Q.thermal_phonon_occupation(om, case["bath_temperature"])[:, None] * np.ones((1, n)),
                D=case["diffusion_coefficient"] * np.sqrt(np.maximum(0.0, 1.0 - (case["energy_gap"] / E) ** 2)), n=n)


def _case():
    return cases.meander_c2(ny=20, nx=22, ne=7, steps=3, pad=2, pitch=6, gap_len=6)


def _worker(rank, world, port, interleave, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = _case()
        tabs = _tables(case)
        plan = ShardPlan(case["num_energy_bins"], tabs["n"], world, rank, interleave)
        stages = OracleStages(plan, case, tabs)
        st = ShardedStepper(plan, stages, diffusion=True, collisions=True)
        g = case["generation"]
        t, recs = 0.0, []
        for _ in range(3):
            rate = g["pulse_rate"] if g["pulse_start"] <= t < g["pulse_start"] + g["pulse_duration"] else None
            recs.append(st.step(case["dt"], 0, rate, want_pauli=True))
            t += case["dt"]
        merged = st.merge_pauli(recs)
        full = st.gather_state()
        if rank == 0:
            np.savez(out, state=full, pauli=np.array(merged, dtype=float), exchanges=st.exchanges)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,interleave", [(2, True), (2, False), (3, True)])
def test_sharded_steps_equal_single_process_oracle(world, interleave, tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), interleave, out), nprocs=world, join=True)
    got = np.load(out)
    case = _case()
    want = helpers.run_oracle(case)
    # same arithmetic, same order per cell / per bin: agreement to rounding of the chunked einsum only
    helpers.assert_close(got["state"], want["state"][-1], "sharded n(E,cell)", rtol=1e-13)
    assert int(got["exchanges"]) == 6
    # merged Pauli records against the oracle's global ones
    edges = Q.extract_edge_segments(case["mask"])
    res = O.run(case["mask"], edges, cases.make_bcs(edges, case["bc"], Q.BoundaryCondition), case["initial_field"],
                case["diffusion_coefficient"], case["dt"], case["total_time"], case["dx"], store_every=3,
                gap=case["energy_gap"], fmin=case["energy_min_factor"], fmax=case["energy_max_factor"],
                ne=case["num_energy_bins"], diffusion=True, recomb=True, scat=True, gamma=case["dynes_gamma"],
                tau_s=case["tau_0"], tau_r=case["tau_0"], Tc=case["T_c"], T_bath=case["bath_temperature"],
                gext=helpers.gen_callable(case))
    n = int(case["mask"].sum())
    for k in range(3):
        mo, (i, q), forb = res.extra["pauli"][k + 1]
        assert abs(got["pauli"][k][0] - mo) <= 1e-12 * mo
        assert int(got["pauli"][k][1]) == i * n + q
        assert int(got["pauli"][k][2]) == (-1 if forb is None else forb[0] * n + forb[1])


# ---- ensembles (BASELINE configs[4]: independent runs, replicas only) ---------------------------------------
def _fake_run(mask=None, dt=1.0, total_time=1.0, tag=0, device=0, **_):
    """Stand-in for run_2d_crank_nicolson with the same return structure (times, frames, mass, limits, eframes, E)."""
    times = [0.0, float(total_time)]
    return times, [None, None], [float(tag), float(tag) * dt], [0.0, 1.0], None, None


def _ensemble_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qpsim_b200.ensemble import parameter_grid, run_ensemble

        members = parameter_grid(dict(total_time=2.0), tag=range(1, 6), dt=[0.5, 0.25])
        ran = []

        def runner(**kw):
            ran.append(kw["tag"] * 100 + int(kw["dt"] * 100))
            assert kw["device"] == rank     # default device = LOCAL_RANK
            return _fake_run(**kw)

        res = run_ensemble(members, runner=runner)
        np.savez(out + f".{rank}.npz", mass=np.array([r[1] for r in res]), ran=np.array(ran))
    finally:
        dist.destroy_process_group()


def test_ensemble_members_are_dealt_round_robin_and_gathered(tmp_path):
    from qpsim_b200.ensemble import member_indices, parameter_grid, run_ensemble

    assert member_indices(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((member_indices(64, 8, r) for r in range(8)), [])) == list(range(64))
    with pytest.raises(ValueError):
        member_indices(4, 2, 2)
    grid = parameter_grid(dict(a=1), b=[1, 2], c=[3, 4, 5])
    assert len(grid) == 6 and grid[0] == dict(a=1, b=1, c=3) and grid[-1] == dict(a=1, b=2, c=5)
    # single process: every member runs here, in order
    single = run_ensemble(parameter_grid(dict(total_time=2.0), tag=range(1, 6), dt=[0.5, 0.25]), runner=_fake_run)
    out = str(tmp_path / "ens")
    world = 3
    mp.spawn(_ensemble_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ran_all = []
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        np.testing.assert_array_equal(z["mass"], np.array([s[1] for s in single]))   # every rank has the full list
        ran_all += z["ran"].tolist()
    assert len(ran_all) == 10 and len(set(ran_all)) == 10                           # each member ran exactly once
