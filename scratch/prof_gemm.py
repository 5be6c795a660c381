"""ncu driver: frozen uniform phonons at 512 bins (tensor-core GEMM path) and segmented 2048^2 sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_sizes import collision_probe, diffusion_probe
what = sys.argv[1]
if what == "gemm":
    collision_probe(512, 16384, 10.0, frozen=True)
else:
    diffusion_probe(2048, 2048, 8, steps=1)
