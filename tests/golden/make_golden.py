"""Generate the golden fixtures by running the UNMODIFIED reference (needs /root/reference; run in the build
container):  python tests/golden/make_golden.py

For every case of tests/cases.golden_cases() the reference's run_2d_crank_nicolson is executed and its outputs
are stored compressed to mask cells: times, mass, the full n(E, cell) at every stored time and the phonon
occupations n_ph(omega, cell).  A second file pins the table builders (energy grid, DOS, base kernels, phonon
map) and direct calls of _apply_fischer_catelani_local_pixel on seeded random states.
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
    qpsim = load_reference()
    if qpsim is None:
        raise SystemExit("reference not available")
    import qpsim.solver as S
    from qpsim.geometry import extract_edge_segments
    from qpsim.models import BoundaryCondition, ExternalGenerationSpec

    class Phys:  # the reference's own builders under the names cases.py expects
        build_energy_grid = staticmethod(S.build_energy_grid)
        thermal_qp_weights = staticmethod(S.thermal_qp_weights)

    only_large = len(sys.argv) > 1 and sys.argv[1] == "large"
    for case in (cases.golden_cases_large() if only_large else cases.golden_cases()):
        mask = case["mask"]
        edges = extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, case["bc"], BoundaryCondition)
        gen = ExternalGenerationSpec(**case["generation"]) if case["generation"] else None
        kw = cases.solver_kwargs(case, edges, bcs, gen, Phys)
        hist = {}
        times, frames, mass, limits, eframes, E = S.run_2d_crank_nicolson(phonon_history_out=hist, **kw)
        out = {"times": np.array(times), "mass": np.array(mass), "limits": np.array(limits)}
        if eframes is not None:
            out["state"] = np.array([[f[mask] for f in t] for t in eframes])
            out["E"] = E
            out["phonons"] = np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]])
            out["omega"] = hist["phonon_energy_bins"]
        else:
            out["state"] = np.array([f[mask] for f in frames])[:, None, :]
        if only_large:   # thinned: two stored times in full, phonons of 512 seeded cells
            keep = np.array([1, len(times) - 1])
            cells = np.sort(np.random.default_rng(5).choice(out["state"].shape[2], 512, replace=False))
            out.update(keep=keep, state=out["state"][keep], ph_cells=cells, phonons=out["phonons"][keep][:, :, cells])
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']:40s} T={len(times)} state{out['state'].shape} -> {os.path.getsize(path)/1024:.1f} KiB")

    if only_large:
        return
    # table builders + per-pixel collision calls
    tabs = {}
    rng = np.random.default_rng(20260101)
    for tag, (ne, fmin, fmax, gamma) in {"a": (24, 1.0, 4.0, 0.18), "b": (50, 1.0, 10.0, 0.0), "c": (16, 1.0, 5.0, 0.18)}.items():
        E, dE = S.build_energy_grid(cases.GAP, fmin, fmax, ne)
        rho = S._dynes_density_of_states(E, cases.GAP, gamma)
        Kr = S.recombination_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
        Ks = S.scattering_kernel_base(E, cases.GAP, cases.TAU, cases.TC)
        om, idd, ids, sg = S._build_phonon_frequency_map(E)
        nph0 = S.thermal_phonon_occupation(om, 0.25)
        wts = S.thermal_qp_weights(E, cases.GAP, 0.25, gamma)
        ncell = 6
        n = rho[:, None] * rng.uniform(0.0, 0.45, size=(ne, ncell))
        ph = nph0[:, None] * rng.uniform(0.5, 2.0, size=(om.size, ncell)) + 1e-6 * rng.random((om.size, ncell))
        n_out = np.empty_like(n)
        ph_out = np.empty_like(ph)
        for c in range(ncell):
            a, b = S._apply_fischer_catelani_local_pixel(n[:, c], ph[:, c], Kr, Ks, rho, idd, ids, sg, dE, 0.3,
                                                         enable_recombination=True, enable_scattering=True)
            n_out[:, c], ph_out[:, c] = a, b
        for k, v in dict(E=E, dE=np.array(dE), rho=rho, Kr=Kr, Ks=Ks, omega=om, idx_diff=idd, idx_sum=ids, sign=sg,
                         nph_thermal=nph0, qp_weights=wts, n_in=n, ph_in=ph, n_out=n_out, ph_out=ph_out,
                         params=np.array([ne, fmin, fmax, gamma])).items():
            tabs[f"{tag}_{k}"] = v
    path = os.path.join(HERE, "tables_and_pixels.npz")
    np.savez_compressed(path, **tabs)
    print("tables ->", os.path.getsize(path) / 1024, "KiB")


if __name__ == "__main__":
    main()
