"""Stub for matplotlib.path used only so `import qpsim` succeeds in the build container."""


class Path:  # pragma: no cover - never used on the hot path
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib is not installed; polygon rasterisation is unavailable in this stub")
