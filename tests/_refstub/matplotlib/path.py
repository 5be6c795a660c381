"""Stand-in for ``matplotlib.path`` so that ``import qpsim`` succeeds in the build container (tests only).

``Path.contains_points`` is a plain even-odd ray-casting test: enough for the reference's polygon-donut test
group (``qpsim/test_cases.py:523-541``), whose cases are recorded together with the mask they ran on - parity
is solver against solver on the same mask, so the rasterisation need not agree with matplotlib's bit for bit
(SURVEY.md section 8c).
"""
import numpy as np


class Path:  # pragma: no cover - never used on the hot path
    def __init__(self, vertices, codes=None, **kwargs):
        if codes is not None:
            raise RuntimeError("matplotlib is not installed; only plain closed polygons are supported by this stub")
        self.vertices = np.asarray(vertices, dtype=float).reshape(-1, 2)

    def contains_points(self, points, transform=None, radius=0.0):
        pts = np.asarray(points, dtype=float).reshape(-1, 2)
        x, y = pts[:, 0], pts[:, 1]
        inside = np.zeros(pts.shape[0], dtype=bool)
        v = self.vertices
        for k in range(v.shape[0]):
            x0, y0 = v[k]
            x1, y1 = v[(k + 1) % v.shape[0]]
            if y0 == y1:
                continue
            crosses = (y0 > y) != (y1 > y)
            xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
            inside ^= crosses & (x < xi)
        return inside
