import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, helpers
import qpsim_b200 as Q
warnings.simplefilter("ignore")
side, ne, steps = int(sys.argv[1]), int(sys.argv[2]), 5
mask = np.ones((side, side), bool)
yy, xx = np.mgrid[:side, :side]
gap_field = cases.GAP * (1.0 - 0.1 * xx / side)          # one gap value per column
case = dict(name="nonuni", mask=mask, bc="reflective", initial_field=cases.gaussian_field(mask, sigma=0.2),
            diffusion_coefficient=cases.D0, dt=0.3, total_time=0.3 * steps, dx=1.0, store_every=steps,
            energy_gap=cases.GAP, energy_min_factor=1.0, energy_max_factor=4.0, num_energy_bins=ne, weights=None,
            enable_diffusion=True, enable_recombination=True, enable_scattering=True, dynes_gamma=cases.GAMMA,
            tau_0=cases.TAU, T_c=cases.TC, bath_temperature=cases.TBATH, generation=None, gap_values=gap_field[mask])
edges = Q.extract_edge_segments(mask); bcs = cases.make_bcs(edges, case["bc"], Q.BoundaryCondition)
kw = cases.solver_kwargs(case, edges, bcs, None, Q.physics)
for variant in ("both", "collisions only", "diffusion only"):
    k2 = dict(kw)
    if variant == "collisions only": k2["enable_diffusion"] = False
    if variant == "diffusion only": k2["enable_recombination"] = k2["enable_scattering"] = False
    Q.run_2d_crank_nicolson(**k2, store_energy_frames=False)
    t0 = time.perf_counter(); Q.run_2d_crank_nicolson(**k2, store_energy_frames=False); dt = time.perf_counter() - t0
    info = Q.solver.last_run_info
    print(f"{side}x{side}x{ne} non-uniform gap ({len(np.unique(case['gap_values']))} gap values) {variant}: "
          f"{info['last_advance_ms']/steps:.2f} ms/step device, call {dt:.2f} s, sweeps/step {info['sweeps']/steps:.0f}, path {info['sweep_path']}", flush=True)
