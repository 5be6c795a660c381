"""Output path (second half): vectorised frame_to_jsonable and the binary sidecar.  CPU."""
import json
import time

import numpy as np

import qpsim_b200 as Q
from refimport import load_reference


def _frames(rng, t, ne, shape, mask):
    out = []
    for _ in range(t):
        out.append([np.where(mask, rng.random(shape), np.nan) for _ in range(ne)])
    return out


def test_frame_to_jsonable_equals_the_reference_loop():
    rng = np.random.default_rng(1)
    mask = rng.random((17, 23)) > 0.3
    frame = np.where(mask, rng.normal(size=mask.shape), np.nan)
    got = Q.frame_to_jsonable(frame)
    want = [[None if np.isnan(v) else float(v) for v in row] for row in frame]     # qpsim/storage.py:57-61
    assert got == want and all(type(v) in (float, type(None)) for row in got for v in row)
    assert json.loads(json.dumps(got)) == want
    if load_reference() is not None:
        from qpsim.storage import frame_to_jsonable as ref
        assert ref(frame) == got
    np.testing.assert_array_equal(Q.output.frame_from_jsonable(got), frame)
    big = np.where(rng.random((512, 512)) > 0.2, 1.5, np.nan)
    t0 = time.perf_counter(); a = Q.frame_to_jsonable(big); t1 = time.perf_counter()
    b = [[None if np.isnan(v) else float(v) for v in row] for row in big]; t2 = time.perf_counter()
    assert a == b and (t1 - t0) < (t2 - t1)


def test_sidecar_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    mask = rng.random((9, 14)) > 0.25
    eframes = _frames(rng, 3, 5, mask.shape, mask)
    frames = [np.where(mask, rng.random(mask.shape), np.nan) for _ in range(3)]
    hist = {"phonon_frames": [np.where(mask, rng.random(mask.shape), np.nan) for _ in range(3)],
            "phonon_energy_frames": _frames(rng, 3, 7, mask.shape, mask), "phonon_energy_bins": np.arange(7.0),
            "phonon_metadata": {"mode": "dynamic_local_coupled"}}
    path = str(tmp_path / "run.json")
    Q.save_result(path, [0.0, 0.5, 1.0], frames, [1.0, 2.0, 3.0], [0.0, 1.0], eframes, np.linspace(180, 540, 5), hist,
                  metadata={"note": "x"})
    doc = json.load(open(path))
    assert doc["times"] == [0.0, 0.5, 1.0] and doc["mass_over_time"] == [1.0, 2.0, 3.0] and doc["energy_frames"]
    back = Q.load_result(path)
    for a, b in zip(back["frames"], frames):
        np.testing.assert_array_equal(a, b)
    for ta, tb in zip(back["energy_frames"], eframes):
        for a, b in zip(ta, tb):
            np.testing.assert_array_equal(a, b)
    for ta, tb in zip(back["phonon_history"]["phonon_energy_frames"], hist["phonon_energy_frames"]):
        for a, b in zip(ta, tb):
            np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(back["energy_bins"], np.linspace(180, 540, 5))
    assert back["phonon_history"]["phonon_metadata"] == {"mode": "dynamic_local_coupled"}
