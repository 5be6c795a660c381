"""Locate and import the unmodified reference (`/root/reference`) when it exists.

The reference is present only in the build container; on the GPU box every caller
must cope with ``load_reference() is None`` (tests skip, golden fixtures are used).
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("QPSIM_REFERENCE_ROOT", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_refstub")


def load_reference():
    """Return the imported ``qpsim`` reference package or None."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "qpsim")):
        return None
    try:
        import matplotlib.path  # noqa: F401
    except Exception:
        if _STUB not in sys.path:
            sys.path.append(_STUB)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        return importlib.import_module("qpsim")
    except Exception:
        return None
