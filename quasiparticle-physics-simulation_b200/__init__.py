"""qpsim-b200: the time-stepping hot path of qpsim (masked 2-D Crank-Nicolson diffusion per energy bin plus
the per-cell quasiparticle-phonon collision step) as hand-written sm_100a CUDA behind a C ABI.

Public surface = what the reference exposes for this path (qpsim/solver.py):
``run_2d_crank_nicolson`` and the two in-place collision helpers, plus the argument types they take.
Import this package as ``qpsim_b200`` (the directory name contains hyphens).
"""
from .models import (  # noqa: F401
    BoundaryAssignmentError,
    BoundaryCondition,
    BoundaryFace,
    EdgeSegment,
    ExternalGenerationSpec,
    InitialConditionSpec,
)
from .geometry import compile_boundaries, extract_edge_segments  # noqa: F401
from .physics import (  # noqa: F401
    build_energy_grid,
    density_of_states,
    phonon_frequency_map,
    recombination_kernel_base,
    scattering_kernel_base,
    recombination_kernel,
    scattering_kernel,
    thermal_generation,
    thermal_phonon_occupation,
    thermal_qp_weights,
)
from .solver import (  # noqa: F401
    apply_collision_step_fischer_catelani_nonuniform,
    apply_collision_step_fischer_catelani_uniform,
    apply_recombination_step,
    apply_scattering_step,
    reconstruct_field,
    run_2d_crank_nicolson,
)
from . import capi, output, userexpr  # noqa: F401
from .output import frame_to_jsonable, frames_to_jsonable, load_result, save_result  # noqa: F401
from .ensemble import parameter_grid, run_ensemble  # noqa: F401

__all__ = [
    "run_2d_crank_nicolson",
    "apply_collision_step_fischer_catelani_uniform",
    "apply_collision_step_fischer_catelani_nonuniform",
    "apply_scattering_step", "apply_recombination_step",
    "BoundaryCondition", "BoundaryFace", "EdgeSegment", "ExternalGenerationSpec", "InitialConditionSpec", "BoundaryAssignmentError",
    "extract_edge_segments", "compile_boundaries", "build_energy_grid", "capi", "run_ensemble", "parameter_grid",
    "frame_to_jsonable", "frames_to_jsonable", "save_result", "load_result",
]
