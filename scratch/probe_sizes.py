"""Timing probe of the stages at the BASELINE shapes (device-resident, events inside the library)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpsim_b200 as Q
from qpsim_b200 import capi
import cases
import _libswitch  # noqa: F401  (QPB_LIB=... selects another build of the library)

def diffusion_probe(ny, nx, ne, dt=0.2, fmax=3.0, steps=3):
    mask = np.ones((ny, nx), bool)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, fmax, ne)
    D = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
    edges = Q.extract_edge_segments(mask)
    bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
    bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
    n = ny * nx
    rng = np.random.default_rng(3)
    st = np.exp(rng.standard_normal((ne, n)) * 0.3) * 1e-4
    t0 = time.perf_counter()
    with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=n, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(mask, bcx, bcy, src)
        ctx.upload_diffusion(D)
        ctx.prepare_diffusion(0, dt)
        ctx.set_state(st)
        t1 = time.perf_counter()
        ctx.advance(1, dt)
        d0 = ctx.diag()
        ctx.advance(steps, dt)
        d1 = ctx.diag()
        ms = d1["last_advance_ms"] / steps
        bs = (d1["bin_sweeps"] - d0["bin_sweeps"]) / steps
        sw = (d1["sweeps"] - d0["sweeps"]) / steps
        m0 = st.sum(); m1 = ctx.get_state(want_phonons=False)[0].sum()
    print(f"diffusion {ny}x{nx}x{ne}: setup {t1-t0:.2f}s  {ms:.2f} ms/step  sweeps/step {sw:.0f}  "
          f"algorithmic {16.0*n*bs/(ms*1e-3)/1e9:.0f} GB/s  mass drift {abs(m1-m0)/m0:.2e}", flush=True)

def collision_probe(ne, ncell, fmax, frozen=False):
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, fmax, ne)
    rho = Q.density_of_states(E, cases.GAP, cases.GAMMA)
    Kr = Q.recombination_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    Ks = Q.scattering_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    nph = Q.thermal_phonon_occupation(om, 0.25 if frozen else cases.TBATH)
    rng = np.random.default_rng(5)
    st = (rho / (rho.sum() * dE))[:, None] * (1e-4 * np.exp(0.3 * rng.standard_normal((1, ncell))))
    flags = capi.F_SCATTERING | capi.F_RECOMBINATION | (capi.F_FREEZE_PHONONS if frozen else 0)
    with capi.Context(ny=1, nx=ncell, ne=ne, nw=om.size, ncell=ncell, flags=flags, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(np.ones((1, ncell), np.uint8))
        ctx.upload_collision(Kr[None], Ks[None], rho[None], None, idd, ids, sg)
        ctx.set_state_uniform_phonons(st, nph)
        ctx.collide(0.05); ctx.synchronize()
        ctx.enable_timers(True); ctx.reset_timers()
        for _ in range(3):
            ctx.collide(0.05)
        ctx.synchronize()
        ms, nl = ctx.timer(2)
    ms /= nl
    fl = (8.0 if frozen else 21.0) * ne * ne * ncell   # 8 NE^2 on the GEMM form
    print(f"collision ne={ne} nw={om.size} cells={ncell} frozen={frozen}: {ms:.3f} ms/call  {fl/(ms*1e-3)/1e12:.2f} TFLOP/s (algorithmic)", flush=True)

if __name__ == "__main__":
    what = sys.argv[1:] or ["coll", "diff"]
    if "coll" in what:
        collision_probe(128, 45952, 5.0)
        collision_probe(256, 32768, 3.0)
        collision_probe(512, 16384, 10.0)
        collision_probe(512, 16384, 10.0, frozen=True)
        collision_probe(512, 131072, 10.0, frozen=True)
        collision_probe(128, 45952, 5.0, frozen=True)
        os.environ["QPB_NO_GEMM"] = "1"
        collision_probe(512, 16384, 10.0, frozen=True)
        collision_probe(128, 45952, 5.0, frozen=True)
        del os.environ["QPB_NO_GEMM"]
        collision_probe(64, 65536, 3.0)
    if "diff" in what:
        diffusion_probe(256, 256, 128, dt=0.5, fmax=5.0)
        diffusion_probe(512, 512, 64)
        diffusion_probe(1024, 1024, 32, dt=0.05, fmax=10.0)
        diffusion_probe(2048, 2048, 16)
