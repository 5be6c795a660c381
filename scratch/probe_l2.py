import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_sizes import diffusion_probe
for ne in (8, 16, 32, 64, 128, 256):
    diffusion_probe(256, 256, ne, dt=0.5, fmax=5.0, steps=5)
