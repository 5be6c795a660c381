"""Record the reference's own validation suite (qpsim/test_cases.py:1133-1178, ``generate_test_suite()``: 10 strip,
9 rectangle, 4 polygon-donut, 3 recombination and 2 scattering cases) as golden fixtures.

Needs /root/reference; run in the build container:  python tests/golden/make_golden_suite.py

``generate_test_suite`` is executed UNMODIFIED; the only intervention is a recording wrapper around the name
``run_2d_crank_nicolson`` inside ``qpsim.test_cases`` that forwards to the reference's solver and keeps the
arguments and results of every call.  What is stored per call (``suite_cases.npz``, keys prefixed ``NN_``):

* inputs: mask, initial_field, the boundary condition of every edge in edge order (kind code, value, aux), the
  scalar keyword arguments, energy_weights;
* outputs: all stored times and masses, and the full field at a thinned set of stored times (first, second, a
  few on the way, last) - the run is sequential, so the last frame pins the whole trajectory.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from refimport import load_reference  # noqa: E402

BC_KINDS = ("reflective", "neumann", "dirichlet", "absorbing", "robin")
SCALARS = ("diffusion_coefficient", "dt", "total_time", "dx", "store_every", "energy_gap", "energy_min_factor",
           "energy_max_factor", "num_energy_bins", "enable_diffusion", "enable_recombination", "enable_scattering",
           "dynes_gamma", "tau_0", "T_c", "bath_temperature")
DEFAULTS = dict(store_every=1, energy_gap=0.0, energy_min_factor=1.0, energy_max_factor=10.0, num_energy_bins=50,
                enable_diffusion=True, enable_recombination=False, enable_scattering=False, dynes_gamma=0.0,
                tau_0=440.0, T_c=1.2, bath_temperature=0.1)


def thin(count: int):
    keep = sorted({0, 1, 2, count // 4, count // 2, (3 * count) // 4, count - 2, count - 1} & set(range(count)))
    return np.array(keep, dtype=np.int64)


def main():
    if load_reference() is None:
        raise SystemExit("reference not available")
    import qpsim.solver as S
    import qpsim.test_cases as TC

    calls = []

    def recording(**kw):
        res = S.run_2d_crank_nicolson(**kw)
        calls.append((kw, res))
        return res

    TC.run_2d_crank_nicolson = recording
    suite = TC.generate_test_suite()
    ids = [c.case_id for g in suite.geometry_groups for c in g.cases]
    assert len(ids) == len(calls) == 28, (len(ids), len(calls))

    out = {"case_ids": np.array(ids)}
    for k, (cid, (kw, res)) in enumerate(zip(ids, calls)):
        p = f"{k:02d}_"
        times, frames, mass, limits, eframes, E = res
        mask = np.asarray(kw["mask"], dtype=bool)
        edges = kw["edges"]
        bc = np.zeros((len(edges), 3))
        for j, e in enumerate(edges):
            c = kw["edge_conditions"][e.edge_id]
            bc[j] = (BC_KINDS.index(c.kind.strip().lower()),
                     np.nan if c.value is None else c.value, np.nan if c.aux_value is None else c.aux_value)
        out[p + "mask"] = mask
        out[p + "initial_field"] = np.asarray(kw["initial_field"], dtype=float)
        out[p + "bc"] = bc
        out[p + "edge_ids"] = np.array([e.edge_id for e in edges])
        out[p + "scalars"] = np.array([float(kw.get(s, DEFAULTS.get(s))) for s in SCALARS])
        if kw.get("energy_weights") is not None:
            out[p + "energy_weights"] = np.asarray(kw["energy_weights"], dtype=float)
        keep = thin(len(times))
        out[p + "times"] = np.asarray(times)
        out[p + "mass"] = np.asarray(mass)
        out[p + "keep"] = keep
        if eframes is not None:
            out[p + "state"] = np.array([[f[mask] for f in eframes[t]] for t in keep])
        else:
            out[p + "state"] = np.array([frames[t][mask] for t in keep])[:, None, :]
        print(f"{k:02d} {cid:45s} cells={int(mask.sum()):5d} stored={len(times):5d} kept={len(keep)} "
              f"state{out[p + 'state'].shape}")
    path = os.path.join(HERE, "suite_cases.npz")
    np.savez_compressed(path, **out)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
