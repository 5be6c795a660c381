"""A/B of the collision kernel on small energy grids (<= 64 bins): 256-thread CTAs, two per SM, against the 512-thread
CTA (QPB_COLL_NT=512); results compared bit for bit between the two and against the oracle on a few cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpsim_b200 as Q
from qpsim_b200 import capi
import cases


def run(ne, ncell, fmax, nt):
    if nt:
        os.environ["QPB_COLL_NT"] = str(nt)
    else:
        os.environ.pop("QPB_COLL_NT", None)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, fmax, ne)
    rho = Q.density_of_states(E, cases.GAP, cases.GAMMA)
    Kr = Q.recombination_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    Ks = Q.scattering_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    nph = Q.thermal_phonon_occupation(om, cases.TBATH)
    rng = np.random.default_rng(5)
    st = (rho / (rho.sum() * dE))[:, None] * (1e-4 * np.exp(0.3 * rng.standard_normal((1, ncell))))
    flags = capi.F_SCATTERING | capi.F_RECOMBINATION
    with capi.Context(ny=1, nx=ncell, ne=ne, nw=om.size, ncell=ncell, flags=flags, dx=1.0, dE=dE) as ctx:
        ctx.upload_geometry(np.ones((1, ncell), np.uint8))
        ctx.upload_collision(Kr[None], Ks[None], rho[None], None, idd, ids, sg)
        ctx.set_state_uniform_phonons(st, nph)
        ctx.collide(0.05); ctx.synchronize()
        ctx.enable_timers(True); ctx.reset_timers()
        for _ in range(5):
            ctx.collide(0.05)
        ctx.synchronize()
        ms, nl = ctx.timer(2)
        s, p = ctx.get_state()
    ms /= nl
    print(f"collision ne={ne} cells={ncell} QPB_COLL_NT={nt or 'default'}: {ms:.4f} ms/call "
          f"{21.0 * ne * ne * ncell / (ms * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)
    return s, p


for ne, ncell, fmax in ((64, 65536, 3.0), (64, 45952, 3.0), (32, 65536, 3.0), (48, 65536, 3.0), (128, 45952, 5.0)):
    a = run(ne, ncell, fmax, 512)
    b = run(ne, ncell, fmax, 0)
    print("   identical:", np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]),
          " max rel diff n:", float(np.max(np.abs(a[0] - b[0]) / np.abs(a[0]).max())), flush=True)
