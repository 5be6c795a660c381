"""Output path of a run: what the reference does with the frames a stepping loop
returns.  Once the steps are fast this is where a GUI run spends its time: ``frame_to_jsonable`` walks every value of
every frame in a Python loop (``qpsim/storage.py:57-61``, called T x NE times at ``qpsim/ui/main_app.py:1971-1975``)
and the result is stored as JSON text.

* :func:`frame_to_jsonable` / :func:`frames_to_jsonable` - the same nested lists (``None`` where the frame is NaN,
  Python floats elsewhere), built with array operations;
* :func:`save_result` / :func:`load_result` - the run's 6-tuple (plus the phonon history) as a small JSON document in
  the reference's field names (``times``, ``mass_over_time``, ``color_limits``, ``energy_bins``, ...) and a binary
  ``.npz`` sidecar holding every frame compressed to the mask cells (8 B per cell and frame instead of ~20 B of JSON
  text per value; NaN padding is re-created on load from the stored mask).
"""
from __future__ import annotations

import json
import os
from typing import Any

import numpy as np

SIDECAR_FORMAT = 1


def frame_to_jsonable(frame: np.ndarray) -> list:
    """qpsim/storage.py:57-61 without the per-value Python loop."""
    a = np.asarray(frame, dtype=float)
    obj = a.astype(object)
    obj[np.isnan(a)] = None
    return obj.tolist()


def frames_to_jsonable(frames) -> list:
    """A list of frames (or a list of lists of frames: energy_frames[t][i]) at once."""
    if frames is None:
        return None
    return [frames_to_jsonable(f) if isinstance(f, (list, tuple)) else frame_to_jsonable(f) for f in frames]


def frame_from_jsonable(frame) -> np.ndarray:
    """qpsim/storage.py:64-65."""
    return np.array([[np.nan if v is None else float(v) for v in row] for row in frame], dtype=float)


def _mask_of(frames) -> np.ndarray:
    return ~np.isnan(np.asarray(frames[0], dtype=float))


def save_result(path: str, times, frames, mass, limits, energy_frames=None, energy_bins=None,
                phonon_history: dict | None = None, metadata: dict | None = None) -> tuple[str, str]:
    """Write ``path`` (JSON) and ``path + '.npz'`` (frames on the mask cells).  Returns both file names."""
    mask = _mask_of(frames)
    arrays: dict[str, Any] = {"mask": mask, "frames": np.array([np.asarray(f)[mask] for f in frames])}
    doc: dict[str, Any] = {
        "sidecar_format": SIDECAR_FORMAT, "sidecar": os.path.basename(path) + ".npz",
        "times": [float(t) for t in times], "mass_over_time": [float(m) for m in mass],
        "color_limits": [float(v) for v in limits], "grid_shape": [int(s) for s in mask.shape],
        "energy_bins": None if energy_bins is None else np.asarray(energy_bins, dtype=float).tolist(),
        "frames": {"sidecar_key": "frames"}, "energy_frames": None, "metadata": metadata or {},
    }
    if energy_frames is not None:
        if any(t is None for t in energy_frames):
            raise ValueError("energy_frames holds unfilled entries (a run with store_energy_frames=False?)")
        arrays["energy_frames"] = np.array([[np.asarray(f)[mask] for f in t] for t in energy_frames])
        doc["energy_frames"] = {"sidecar_key": "energy_frames"}
    if phonon_history:
        doc["phonon_metadata"] = phonon_history.get("phonon_metadata")
        pb = phonon_history.get("phonon_energy_bins")
        doc["phonon_energy_bins"] = None if pb is None else np.asarray(pb, dtype=float).tolist()
        if phonon_history.get("phonon_frames") is not None:
            arrays["phonon_frames"] = np.array([np.asarray(f)[mask] for f in phonon_history["phonon_frames"]])
            doc["phonon_frames"] = {"sidecar_key": "phonon_frames"}
        if phonon_history.get("phonon_energy_frames") is not None:
            arrays["phonon_energy_frames"] = np.array(
                [[np.asarray(f)[mask] for f in t] for t in phonon_history["phonon_energy_frames"]])
            doc["phonon_energy_frames"] = {"sidecar_key": "phonon_energy_frames"}
    np.savez(path + ".npz", **arrays)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(doc, f, indent=2)
    return path, path + ".npz"


def _expand(mask: np.ndarray, values: np.ndarray):
    out = np.full(values.shape[:-1] + mask.shape, np.nan)
    out[..., mask] = values
    return out


def load_result(path: str) -> dict:
    """Inverse of :func:`save_result`: the lists ``run_2d_crank_nicolson`` returned, NaN padded."""
    with open(path, encoding="utf-8") as f:
        doc = json.load(f)
    if doc.get("sidecar_format") != SIDECAR_FORMAT:
        raise ValueError("not a sidecar result document of this format")
    with np.load(os.path.join(os.path.dirname(path), doc["sidecar"])) as z:
        mask = z["mask"]
        out = {"times": doc["times"], "mass": doc["mass_over_time"], "color_limits": doc["color_limits"],
               "energy_bins": None if doc["energy_bins"] is None else np.array(doc["energy_bins"]),
               "frames": list(_expand(mask, z["frames"])), "energy_frames": None, "metadata": doc.get("metadata", {})}
        if doc.get("energy_frames"):
            out["energy_frames"] = [list(t) for t in _expand(mask, z["energy_frames"])]
        hist = {}
        if doc.get("phonon_frames"):
            hist["phonon_frames"] = list(_expand(mask, z["phonon_frames"]))
        if doc.get("phonon_energy_frames"):
            hist["phonon_energy_frames"] = [list(t) for t in _expand(mask, z["phonon_energy_frames"])]
        if hist or doc.get("phonon_metadata"):
            hist["phonon_energy_bins"] = (None if doc.get("phonon_energy_bins") is None
                                          else np.array(doc["phonon_energy_bins"]))
            hist["phonon_metadata"] = doc.get("phonon_metadata")
        out["phonon_history"] = hist or None
    return out
