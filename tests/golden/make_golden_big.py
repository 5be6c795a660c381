"""Full-size reference runs stored as thinned fixtures (needs /root/reference; build container only; takes minutes):

  python tests/golden/make_golden_big.py [c2] [c1] [mkid]

* ``c2_full_256x256x128``: BASELINE configs[1] exactly as ``bench.py`` runs it (``bench.c2_workload()``), two full
  time steps of the UNMODIFIED reference.  The state of one stored time is 47 MB, so the
  fixture keeps: every cell of 4 bins, every bin of 2048 seeded cells, the integrated field, the masses and the
  phonon occupations of 256 of those cells.
* ``c1_strip_128x64_20steps``: BASELINE configs[0], 20 steps, every stored time in full.
* ``mkid_crosscheck_48x12``: the inputs of the reference's own cross-check test
  (tests/test_mkid_crosscheck.py:110-165), 12 steps, every stored time in full.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cases  # noqa: E402
from refimport import load_reference  # noqa: E402

C2_BINS = (0, 1, 64, 127)
C2_CELLS = 2048
C2_PH_CELLS = 256


def c2_sample(n):
    rng = np.random.default_rng(20260102)
    cells = np.sort(rng.choice(n, size=C2_CELLS, replace=False))
    return cells, cells[:: C2_CELLS // C2_PH_CELLS]


def mkid_case():
    """tests/test_mkid_crosscheck.py:110-165 of the reference: 1 x 48 strip, 12 bins, tau = 400 ns."""
    case = cases.strip_c1(steps=12, nx=48, ne=12)
    case.update(name="mkid_crosscheck_48x12", tau_0=400.0, store_every=1)
    return case


def main(which):
    if load_reference() is None:
        raise SystemExit("reference not available")
    import qpsim.solver as S
    from qpsim.geometry import extract_edge_segments
    from qpsim.models import BoundaryCondition, ExternalGenerationSpec

    class Phys:
        build_energy_grid = staticmethod(S.build_energy_grid)
        thermal_qp_weights = staticmethod(S.thermal_qp_weights)

    def run(case):
        mask = case["mask"]
        edges = extract_edge_segments(mask)
        bcs = cases.make_bcs(edges, case["bc"], BoundaryCondition)
        gen = ExternalGenerationSpec(**case["generation"]) if case["generation"] else None
        kw = cases.solver_kwargs(case, edges, bcs, gen, Phys)
        hist = {}
        t0 = time.time()
        res = S.run_2d_crank_nicolson(phonon_history_out=hist, **kw)
        print(f"{case['name']}: reference took {time.time() - t0:.1f} s")
        return res, hist

    if "c2" in which:
        import bench

        case = bench.c2_case(steps=2)
        (times, frames, mass, limits, eframes, E), hist = run(case)
        mask = case["mask"]
        n = int(mask.sum())
        cells, ph_cells = c2_sample(n)
        state = np.array([[f[mask] for f in t] for t in eframes])          # (T, NE, N)
        ph = np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]])
        out = dict(times=np.array(times), mass=np.array(mass), limits=np.array(limits), E=E,
                   integrated=np.array([f[mask] for f in frames]), bins=np.array(C2_BINS), cells=cells,
                   ph_cells=ph_cells, state_bins=state[:, list(C2_BINS), :], state_cells=state[:, :, cells],
                   phonons_cells=ph[:, :, ph_cells])
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")

    for tag, case in (("c1", cases.strip_c1(steps=20, nx=128, ne=64)), ("mkid", mkid_case())):
        if tag not in which:
            continue
        if tag == "c1":
            case["name"] += "_20steps"
        (times, frames, mass, limits, eframes, E), hist = run(case)
        mask = case["mask"]
        out = dict(times=np.array(times), mass=np.array(mass), limits=np.array(limits), E=E,
                   state=np.array([[f[mask] for f in t] for t in eframes]),
                   phonons=np.array([[f[mask] for f in t] for t in hist["phonon_energy_frames"]]))
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("->", path, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2", "c1", "mkid"])
