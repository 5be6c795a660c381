import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_sizes import diffusion_probe
for shape in [(2048, 256, 16), (256, 2048, 16), (512, 1024, 16), (1024, 512, 16), (512, 512, 32), (256, 1024, 32)]:
    diffusion_probe(shape[0], shape[1], shape[2], dt=0.5, fmax=5.0, steps=5)
