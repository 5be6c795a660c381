"""Argument types of the drop-in solver entry point.

These mirror the fields the reference's hot path reads from its dataclasses (``qpsim/models.py:33-130``:
BoundaryCondition, BoundaryFace, EdgeSegment, ExternalGenerationSpec).  The drop-in is duck-typed: objects
created by the reference's own ``qpsim.models`` work unchanged; these classes exist so the package is usable
(and testable) where the reference is not installed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

BOUNDARY_KINDS = ("reflective", "neumann", "dirichlet", "absorbing", "robin")
COLLISION_SOLVERS = ("fischer_catelani_local",)
GENERATION_MODES = ("none", "constant", "pulse", "custom")


class BoundaryAssignmentError(ValueError):
    """Same role as qpsim.solver.BoundaryAssignmentError (solver.py:21)."""


def check_collision_solver(name: str) -> str:
    """qpsim/models.py:23-30: only the local coupled solver exists."""
    key = str(name).strip().lower()
    if key not in COLLISION_SOLVERS:
        raise ValueError(
            f"Unsupported collision solver '{name}'. Supported values: {', '.join(sorted(COLLISION_SOLVERS))}."
        )
    return key


@dataclass
class BoundaryCondition:
    kind: str
    value: float | None = None
    aux_value: float | None = None

    def normalized_kind(self) -> str:
        return self.kind.strip().lower()

    def validate(self) -> None:
        kind = self.normalized_kind()
        if kind not in BOUNDARY_KINDS:
            raise ValueError(f"Unsupported boundary condition kind: {self.kind}")
        if kind in ("neumann", "dirichlet", "robin") and self.value is None:
            raise ValueError(f"Boundary condition '{kind}' requires a numeric value")


@dataclass
class BoundaryFace:
    row: int
    col: int
    direction: str  # "up" | "down" | "left" | "right"


@dataclass
class EdgeSegment:
    edge_id: str
    x0: float
    y0: float
    x1: float
    y1: float
    normal: str
    faces: list = field(default_factory=list)


@dataclass
class ExternalGenerationSpec:
    mode: str = "none"
    rate: float = 0.0
    pulse_start: float = 0.0
    pulse_duration: float = 10.0
    pulse_rate: float = 0.0
    custom_body: str = "return 0.0"
    custom_params: dict[str, Any] = field(default_factory=dict)

    def normalized_mode(self) -> str:
        return self.mode.strip().lower()

    def validate(self) -> None:
        if self.normalized_mode() not in GENERATION_MODES:
            raise ValueError(
                f"Unsupported external generation mode '{self.mode}'. "
                f"Supported: {', '.join(sorted(GENERATION_MODES))}."
            )
        if self.rate < 0:
            raise ValueError("External generation constant rate must be non-negative.")
        if self.pulse_rate < 0:
            raise ValueError("External generation pulse rate must be non-negative.")
        if self.pulse_duration < 0:
            raise ValueError("External generation pulse_duration must be non-negative.")


_FULL_BODY = "return np.exp(-((x-0.5)**2 + (y-0.5)**2) / 0.02) * np.exp(-E / 500.0)"


@dataclass
class InitialConditionSpec:
    """Field names of qpsim/models.py:82-108 (the drop-in reads them by attribute, so the reference's own object works
    as well): split spatial / energy profiles of the quasiparticles and the phonons, plus optional non-separable
    bodies F(x, y, E, params)."""
    spatial_kind: str = ""
    spatial_params: dict[str, Any] = field(default_factory=dict)
    spatial_custom_body: str = "return np.exp(-((x-0.5)**2 + (y-0.5)**2) / 0.02)"
    spatial_custom_params: dict[str, Any] = field(default_factory=dict)
    energy_kind: str = ""
    energy_params: dict[str, Any] = field(default_factory=dict)
    energy_custom_body: str = "return np.ones_like(E)"
    energy_custom_params: dict[str, Any] = field(default_factory=dict)
    qp_full_custom_enabled: bool = False
    qp_full_custom_body: str = _FULL_BODY
    qp_full_custom_params: dict[str, Any] = field(default_factory=dict)
    phonon_spatial_kind: str = ""
    phonon_spatial_params: dict[str, Any] = field(default_factory=dict)
    phonon_spatial_custom_body: str = "return 1.0"
    phonon_spatial_custom_params: dict[str, Any] = field(default_factory=dict)
    phonon_energy_kind: str = ""
    phonon_energy_params: dict[str, Any] = field(default_factory=dict)
    phonon_energy_custom_body: str = "return np.ones_like(E)"
    phonon_energy_custom_params: dict[str, Any] = field(default_factory=dict)
    phonon_full_custom_enabled: bool = False
    phonon_full_custom_body: str = _FULL_BODY
    phonon_full_custom_params: dict[str, Any] = field(default_factory=dict)
