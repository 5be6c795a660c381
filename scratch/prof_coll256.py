"""ncu target: collision half steps at 256 bins (C3's energy grid) on 18 944 cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scratch"))
os.environ.setdefault("QPB_PROF_ONLY", "1")
import probe_coll_ab  # noqa: F401  (runs its four shapes; the first one is the 256-bin launch)
