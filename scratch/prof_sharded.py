"""Per-section device times of the sharded step (run under torchrun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import bench, cases
import qpsim_b200 as Q
from qpsim_b200 import capi
from qpsim_b200.multigpu import *
rank, world, local = init_process_group("nccl")
ty, tx = weak_tiling(world)
w = bench.c2_workload(tile_y=ty, tile_x=tx); tabs = bench.build_tables(w)
mask = w["mask"]; n, ne, nw = tabs["n"], w["num_energy_bins"], int(tabs["omega"].size)
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, w["bc"], Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, w["dx"])
prob = ShardedProblem(mask=mask, bcx=bcx, bcy=bcy, src=src, dx=w["dx"], dE=tabs["dE"], D=tabs["D"], variable_D=False,
                      rho=tabs["rho"][None], Kr=tabs["Kr"][None], Ks=tabs["Ks"][None], gap_id=None, idx_diff=tabs["idx_diff"],
                      idx_sum=tabs["idx_sum"], sign=tabs["sign"], nw=nw, state=tabs["state"], phonons=tabs["phonons"])
plan = ShardPlan(ne, n, world, rank, interleave=True)
dt = w["dt"]
stages = DeviceStages(plan, prob, local, dt)
acc = {}
with torch.cuda.stream(stages.stream):
    st = ShardedStepper(plan, stages, diffusion=True, collisions=True)
    for _ in range(3):
        st.step(dt, 0, w["pulse_rate"])
    def timed(name, fn):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(stages.stream); fn(); e1.record(stages.stream); torch.cuda.synchronize()
        a = acc.setdefault(name, [0.0, 0.0]); a[0] += e0.elapsed_time(e1); a[1] += (time.perf_counter() - t0) * 1e3
    K = 10
    fused = stages.enable_fused_exchange(prob)
    st = ShardedStepper(plan, stages, diffusion=True, collisions=True)
    for _ in range(2):
        st.step(dt, 0, w["pulse_rate"])
    for _ in range(K):
        timed("generation", lambda: stages.add_generation(dt, w["pulse_rate"]))
        if fused:
            timed("collide_x1", lambda: stages.collide_exchange(0.5 * dt, 1))
            timed("barrier", st._rank_barrier)
            timed("diffuse", lambda: stages.diffuse(0))
            timed("barrier", st._rank_barrier)
            timed("collide_x2", lambda: stages.collide_exchange(0.5 * dt, 2))
            timed("barrier", st._rank_barrier)
        else:
            timed("collide", lambda: stages.collide(0.5 * dt))
            timed("to_bins", st.to_bins)
            timed("diffuse", lambda: stages.diffuse(0))
            timed("to_cells", st.to_cells)
            timed("collide", lambda: stages.collide(0.5 * dt))
        timed("pauli_record", lambda: stages.pauli_record(0))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stages.stream)
    for k in range(K):
        st.step(dt, 0, w["pulse_rate"], pauli_slot=k)
    e1.record(stages.stream); torch.cuda.synchronize()
    if rank == 0:
        print("world", world, "cells/rank", plan.ncells(), "bins/rank", plan.nbins(), "grid", mask.shape)
        for k, (dev, wall) in acc.items():
            print(f"{k:14s} device {dev / K:8.3f} ms/step   wall {wall / K:8.3f} ms/step")
        print("whole step (no per-section sync):", e0.elapsed_time(e1) / K, "ms")
stages.close(); dist.destroy_process_group()
