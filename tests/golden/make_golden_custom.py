"""Record the UNMODIFIED reference on the user-expression options of the drop-in boundary (needs /root/reference):

  python tests/golden/make_golden_custom.py

For every case of tests/cases.custom_mode_cases() the fixture tests/golden/custom_modes.npz keeps the run's outputs
(times, mass, n(E, cell) and n_ph(omega, cell) at every stored time) and what the reference's own evaluators made of
the expressions: g_ext at two times (evaluate_external_generation, solver.py:878-964), the initial states
(initial_conditions.py:494-507, 601-632) and the gap map with D(E, x) (precompute.py:171-228) - the CPU tests pin
the package's userexpr module against those bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from refimport import load_reference  # noqa: E402

GEN_TIMES = (0.0, 0.75)


def main():
    if load_reference() is None:
        raise SystemExit("reference not available")
    import qpsim.solver as S
    from qpsim import initial_conditions as IC
    from qpsim.geometry import extract_edge_segments
    from qpsim.models import BoundaryCondition, ExternalGenerationSpec, InitialConditionSpec
    from qpsim.initial_conditions import evaluate_gap_expression

    class Phys:
        build_energy_grid = staticmethod(S.build_energy_grid)
        thermal_qp_weights = staticmethod(S.thermal_qp_weights)

    out = {}
    for case in cases.custom_mode_cases():
        mask = case["mask"]
        edges = extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, case["bc"], BoundaryCondition)
        gen = ExternalGenerationSpec(**case["generation"]) if case["generation"] else None
        kw = cases.solver_kwargs(case, edges, bcs, gen, Phys, InitialConditionSpec)
        hist = {}
        times, frames, mass, limits, eframes, E = S.run_2d_crank_nicolson(phonon_history_out=hist, **kw)
        p = case["name"] + "/"
        out[p + "times"] = np.array(times)
        out[p + "mass"] = np.array(mass)
        out[p + "state"] = np.array([[f[mask] for f in t] for t in eframes])
        out[p + "phonons"] = np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]])
        out[p + "E"] = E
        n = int(mask.sum())
        if gen is not None:
            for k, t in enumerate(GEN_TIMES):
                out[p + f"gext_{k}"] = S.evaluate_external_generation(gen, E, n, t, mask)
        if case.get("ic_spec"):
            spec = kw["initial_condition_spec"]
            qp0 = IC.build_initial_qp_energy_state(mask=mask, E_bins=E, spec=spec)
            if qp0 is not None:
                out[p + "qp0"] = qp0
            out[p + "ph0"] = IC.build_initial_phonon_energy_state(
                mask=mask, omega_bins=hist["phonon_energy_bins"], spec=spec, bath_temperature=case["bath_temperature"])
            out[p + "omega"] = hist["phonon_energy_bins"]
        if case.get("gap_expression"):
            out[p + "gap_values"] = evaluate_gap_expression(case["gap_expression"], mask, case["energy_gap"])
        print(f"{case['name']:28s} T={len(times)} state{out[p + 'state'].shape} mass {mass[0]:.6e} -> {mass[-1]:.6e}")
    path = os.path.join(HERE, "custom_modes.npz")
    np.savez_compressed(path, **out)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
