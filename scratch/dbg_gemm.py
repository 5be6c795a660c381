import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpsim_b200 as Q
import cases, helpers
from oracle import qp_oracle as O
for (ne, n) in [(72, 300), (128, 257), (64, 128)]:
  for rec, sc in [(True, True), (True, False), (False, True)]:
    rng = np.random.default_rng(23)
    E, dE = Q.build_energy_grid(cases.GAP, 1.0, 6.0, ne)
    om, idd, ids, sg = Q.phonon_frequency_map(E)
    rho = Q.density_of_states(E, cases.GAP, 0.18)
    Kr = Q.recombination_kernel_base(E, cases.GAP, 300.0, 1.2)
    Ks = Q.scattering_kernel_base(E, cases.GAP, 500.0, 1.2)
    state0 = rho[:, None] * rng.uniform(0, 0.6, (ne, n))
    ph0 = Q.thermal_phonon_occupation(om, 0.35)[:, None] * np.ones((1, n))
    s_ref, p_ref = state0.copy(), ph0.copy()
    O.collide(s_ref, p_ref, Kr, Ks, rho, idd, ids, sg, dE, 0.4, recomb=rec, scat=sc, update_phonons=False)
    for tag, env in (("gemm", {}), ("gemv", {"QPB_NO_GEMM": "1"}), ("struct", {"QPB_NO_UNIFORM": "1"})):
        for k in ("QPB_NO_GEMM", "QPB_NO_UNIFORM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s, p = state0.copy(), ph0.copy()
        Q.apply_collision_step_fischer_catelani_uniform(s, p, Kr, Ks, rho, idd, ids, sg, dE, 0.4, enable_recombination=rec,
                                                        enable_scattering=sc, update_phonons=False)
        err = np.abs(s - s_ref) / np.max(np.abs(s_ref), axis=1, keepdims=True)
        i, q = np.unravel_index(np.argmax(err), err.shape)
        print(ne, n, rec, sc, tag, f"{err.max():.3e} at bin {i} cell {q}  bad entries {(err > 1e-11).sum()}", flush=True)
