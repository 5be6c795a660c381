"""A/B of the structured collision kernel between two builds of the library (QPB_LIB=... for the other one): time per
call at the C3 (256 bins, 16-cell CTAs), C2 (128 bins) and 64-bin shapes; prints a checksum so that the two runs can
be compared bit for bit."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpsim_b200 as Q
from qpsim_b200 import capi
import cases
import _libswitch  # noqa: F401


def run(ne, ncell, fmax, reps=5):
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
        ctx.collide(0.05); ctx.collide(0.05); ctx.synchronize()
        ctx.enable_timers(True); ctx.reset_timers()
        for _ in range(reps):
            ctx.collide(0.05)
        ctx.synchronize()
        ms, nl = ctx.timer(2)
        s, p = ctx.get_state()
    ms /= max(nl, 1)
    h = hashlib.sha1(s.tobytes() + p.tobytes()).hexdigest()[:12]
    print(f"collision ne={ne} cells={ncell}: {ms:.4f} ms/call {21.0 * ne * ne * ncell / (ms * 1e-3) / 1e12:.2f} TFLOP/s "
          f"launches {nl} sha {h}", flush=True)


run(256, 148 * 16 * 16, 3.0)
run(128, 45952, 5.0)
run(64, 65536, 3.0)
run(384, 148 * 16 * 4, 4.0, reps=3)
