"""Golden fixtures of the reference's fixed-bath Euler collision forms (qpsim/solver.py:493-605): the bath-dressed
kernels recombination_kernel / scattering_kernel, G_therm as qpsim/precompute.py:230-245 builds it, and
apply_scattering_step / apply_recombination_step on seeded states.  Needs /root/reference:
    python tests/golden/make_golden_euler.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from refimport import load_reference  # noqa: E402


def main():
    if load_reference() is None:
        raise SystemExit("reference not available")
    import qpsim.solver as S

    out = {}
    rng = np.random.default_rng(20260106)
    for tag, (ne, ncell, fmax, tbath) in {"small": (24, 37, 4.0, 0.25), "gemm": (72, 150, 6.0, 0.3),
                                         "wide": (130, 70, 10.0, 0.15)}.items():
        E, dE = S.build_energy_grid(cases.GAP, 1.0, fmax, ne)
        rho = S._dynes_density_of_states(E, cases.GAP, cases.GAMMA)
        Kr = S.recombination_kernel(E, cases.GAP, cases.TAU, cases.TC, tbath)
        Ks = S.scattering_kernel(E, cases.GAP, cases.TAU, cases.TC, tbath)
        n_eq = S.thermal_qp_weights(E, cases.GAP, tbath, cases.GAMMA)
        G_therm = 2.0 * n_eq * dE * (Kr @ n_eq)          # qpsim/precompute.py:240
        state = rho[:, None] * rng.uniform(0.0, 0.4, size=(ne, ncell))
        dt = 0.05
        s_scat = state.copy()
        S.apply_scattering_step(s_scat, Ks, rho, dE, dt)
        s_rec = state.copy()
        S.apply_recombination_step(s_rec, Kr, G_therm, dE, dt)
        s_both = state.copy()
        for _ in range(3):
            S.apply_scattering_step(s_both, Ks, rho, dE, dt)
            S.apply_recombination_step(s_both, Kr, G_therm, dE, dt)
        for k, v in dict(E=E, dE=np.array(dE), rho=rho, Kr=Kr, Ks=Ks, G_therm=G_therm, state=state, dt=np.array(dt),
                         after_scattering=s_scat, after_recombination=s_rec, after_3_pairs=s_both,
                         params=np.array([ne, ncell, fmax, tbath])).items():
            out[f"{tag}_{k}"] = v
    path = os.path.join(HERE, "euler_steps.npz")
    np.savez_compressed(path, **out)
    print("euler fixtures ->", os.path.getsize(path) / 1024, "KiB")


if __name__ == "__main__":
    main()
