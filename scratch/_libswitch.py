"""QPB_LIB=path: run a probe script against another build of libqpb.so (missing entry points become dummies)."""
import os
from qpsim_b200 import capi
if os.environ.get("QPB_LIB"):   # A/B against an older build of the library (missing entry points become dummies)
    import ctypes
    capi.LIB_PATH = os.path.abspath(os.environ["QPB_LIB"])
    class _Dummy:
        pass
    class _Tol(ctypes.CDLL):
        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                if name.startswith("qpb_"):
                    return _Dummy()
                raise
    capi.C.CDLL = _Tol
    print("library:", capi.LIB_PATH)
