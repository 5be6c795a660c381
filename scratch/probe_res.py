import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpsim_b200 as Q
from qpsim_b200 import capi
import cases
ny = nx = int(sys.argv[1]); ne = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mask = np.ones((ny, nx), bool)
E, dE = Q.build_energy_grid(cases.GAP, 1.0, 3.0, ne)
D = cases.D0 * np.sqrt(np.maximum(0.0, 1.0 - (cases.GAP / E) ** 2))
edges = Q.extract_edge_segments(mask)
bcs = cases.make_bcs(edges, "reflective", Q.BoundaryCondition)
bcx, bcy, src = Q.compile_boundaries(mask, edges, bcs, 1.0)
rng = np.random.default_rng(3)
st = np.exp(rng.standard_normal((ne, ny * nx)) * 0.3) * 1e-4
with capi.Context(ny=ny, nx=nx, ne=ne, nw=0, ncell=ny * nx, flags=capi.F_DIFFUSION, dx=1.0, dE=dE) as ctx:
    ctx.upload_geometry(mask, bcx, bcy, src); ctx.upload_diffusion(D); ctx.prepare_diffusion(0, 0.2); ctx.set_state(st)
    for _ in range(3):
        ctx.advance(1, 0.2)
    print(ctx.diag())
