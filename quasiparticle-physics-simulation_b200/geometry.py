"""Host-side geometry compiler: mask + edges + boundary conditions -> dense per-cell arrays for the device.

Replaces the per-cell Python loops of ``build_laplacian_with_boundaries`` / ``_apply_boundary_contribution``
(reference ``qpsim/solver.py:112-212``) with vectorised numpy; the Laplacian itself is never assembled — the
CUDA sweeps generate its coefficients from these arrays.  ``extract_edge_segments`` reproduces the edge ids
and face grouping of ``qpsim/geometry.py:150-242`` so boundary-condition dictionaries keyed by edge id mean
the same thing here.
"""
from __future__ import annotations

import numpy as np

from .models import BOUNDARY_KINDS, BoundaryAssignmentError, BoundaryFace, EdgeSegment

_FACE_DIRS = ("up", "down", "left", "right")  # lookup order of the reference (solver.py:25-30)


def boundary_face_masks(mask: np.ndarray) -> dict[str, np.ndarray]:
    """Boolean [ny,nx] arrays: cell has a domain boundary on that side."""
    m = np.asarray(mask, dtype=bool)
    up = m.copy()
    up[1:, :] &= ~m[:-1, :]
    down = m.copy()
    down[:-1, :] &= ~m[1:, :]
    left = m.copy()
    left[:, 1:] &= ~m[:, :-1]
    right = m.copy()
    right[:, :-1] &= ~m[:, 1:]
    return {"up": up, "down": down, "left": left, "right": right}


def boundary_face_indices(mask: np.ndarray) -> dict[str, np.ndarray]:
    """Sorted flat (row-major) indices of the cells that have a domain boundary on each side: the same sets as
    ``flatnonzero`` of :func:`boundary_face_masks`, found through the cells that are not interior (one grid-sized pass
    and a thin candidate list instead of four grid-sized selections)."""
    m = np.asarray(mask, dtype=bool)
    ny, nx = m.shape
    inner = m.copy()
    inner[0, :] = inner[-1, :] = False
    inner[:, 0] = inner[:, -1] = False
    if ny > 2 and nx > 2:
        core = inner[1:-1, 1:-1]
        core &= m[:-2, 1:-1]
        core &= m[2:, 1:-1]
        core &= m[1:-1, :-2]
        core &= m[1:-1, 2:]
    cand = np.flatnonzero((m & ~inner).ravel())      # mask cells with at least one missing neighbour
    r, c = np.divmod(cand, nx)
    flat = m.ravel()
    out = {}
    has = np.zeros(cand.size, dtype=bool)
    for name, ok, nb in (("up", r > 0, cand - nx), ("down", r < ny - 1, cand + nx),
                         ("left", c > 0, cand - 1), ("right", c < nx - 1, cand + 1)):
        has[:] = False
        has[ok] = flat[nb[ok]]
        out[name] = cand[~has]
    return out


def _runs(sorted_pos: np.ndarray):
    """Split a sorted integer array into maximal runs of consecutive values -> list of (start, end_exclusive)."""
    if sorted_pos.size == 0:
        return []
    breaks = np.nonzero(np.diff(sorted_pos) != 1)[0] + 1
    starts = np.concatenate(([0], breaks))
    ends = np.concatenate((breaks, [sorted_pos.size]))
    return [(int(sorted_pos[a]), int(sorted_pos[b - 1]) + 1) for a, b in zip(starts, ends)]


def extract_edge_segments(mask: np.ndarray) -> list[EdgeSegment]:
    """Group boundary faces into straight edges; same ids and order as qpsim/geometry.py:150-242.

    Horizontal faces first, ordered by their y coordinate then by normal name ("down" < "up"), each split into
    maximal runs of adjacent columns; then vertical faces ordered by x then normal ("left" < "right").
    """
    m = np.asarray(mask, dtype=bool)
    faces = boundary_face_indices(m)
    nx_grid = m.shape[1]
    out: list[EdgeSegment] = []

    def add(normal, line, a, b):
        eid = f"edge_{len(out) + 1:04d}"
        if normal in ("up", "down"):
            row = line if normal == "up" else line - 1
            fl = [BoundaryFace(row=row, col=c, direction=normal) for c in range(a, b)]
            out.append(EdgeSegment(eid, float(a), float(line), float(b), float(line), normal, fl))
        else:
            col = line if normal == "left" else line - 1
            fl = [BoundaryFace(row=r, col=col, direction=normal) for r in range(a, b)]
            out.append(EdgeSegment(eid, float(line), float(a), float(line), float(b), normal, fl))

    groups = []
    for normal in ("down", "up"):
        rows, cols = np.divmod(faces[normal], nx_grid)
        line = rows + (1 if normal == "down" else 0)
        for y in np.unique(line):
            groups.append((int(y), normal, np.sort(cols[line == y])))
    for y, normal, pos in sorted(groups, key=lambda g: (g[0], g[1])):
        for a, b in _runs(pos):
            add(normal, y, a, b)
    groups = []
    for normal in ("left", "right"):
        rows, cols = np.divmod(faces[normal], nx_grid)
        line = cols + (1 if normal == "right" else 0)
        for x in np.unique(line):
            groups.append((int(x), normal, np.sort(rows[line == x])))
    for x, normal, pos in sorted(groups, key=lambda g: (g[0], g[1])):
        for a, b in _runs(pos):
            add(normal, x, a, b)
    return out


def _face_arrays(edge):
    """[(direction, rows, cols)] of an edge's faces as index arrays.  The per-face attribute walk is the expensive part
    of compiling a geometry (7 000 faces at 256 x 256), so the result is kept on the edge object and reused as long as
    its face list is the same list with the same end faces (a GUI reruns one geometry many times)."""
    faces = edge.faces
    key = (id(faces), len(faces), id(faces[0]), id(faces[-1]))
    cached = getattr(edge, "_qpb_face_arrays", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    rows = np.fromiter((f.row for f in faces), dtype=np.int64, count=len(faces))
    cols = np.fromiter((f.col for f in faces), dtype=np.int64, count=len(faces))
    dirs = {f.direction for f in faces}
    if len(dirs) == 1:
        out = [(next(iter(dirs)), rows, cols)]
    else:
        out = []
        for d in dirs:
            sel = np.fromiter((f.direction == d for f in faces), dtype=bool, count=len(faces))
            out.append((d, rows[sel], cols[sel]))
    try:
        edge._qpb_face_arrays = (key, out)
    except Exception:   # objects that refuse new attributes (slots, frozen): just do not cache
        pass
    return out


def compile_boundaries(mask: np.ndarray, edges, edge_conditions, dx: float):
    """Return dense (bcx, bcy, source) arrays, each [ny,nx] float64.

    For every boundary face of kind (solver.py:112-149):
      reflective -> nothing;  absorbing -> diagonal 2 (units 1/dx^2);  dirichlet(g) -> diagonal 2, source 2g/dx^2;
      neumann(q) -> source q/dx;  robin(beta,gamma) -> diagonal beta*dx, source gamma/dx.
    ``bcx`` collects the diagonal terms of left/right faces, ``bcy`` of up/down faces.
    Raises BoundaryAssignmentError exactly where the reference does (solver.py:169-173, 196-199).
    """
    if dx <= 0:
        raise ValueError("dx must be positive.")
    m = np.asarray(mask)
    if m.ndim != 2:
        raise ValueError("mask must be 2D.")
    m = m.astype(bool)
    if not m.any():
        raise ValueError("Geometry mask has no interior points.")
    ny, nx = m.shape
    inv_dx = 1.0 / dx
    inv_dx2 = inv_dx * inv_dx
    # The boundary faces are a thin set: everything is kept on their sorted flat indices, one short array per
    # direction - no grid-sized scratch arrays (at 2048 x 2048 a dozen of them cost more in page faults than the rest).
    faces = boundary_face_indices(m)
    covered = {d: np.zeros(faces[d].size, dtype=bool) for d in _FACE_DIRS}
    # later edges overwrite earlier ones for a shared face (dict semantics of solver.py:37-50); accumulate per
    # face first, then add, so a face listed twice is not counted twice
    face_diag = {d: np.zeros(faces[d].size) for d in _FACE_DIRS}
    face_src = {d: np.zeros(faces[d].size) for d in _FACE_DIRS}

    def positions(d, flat):
        """Index into faces[d] of every flat cell index that is a boundary face of direction d (others dropped)."""
        fd = faces[d]
        if fd.size == 0:
            return np.zeros(0, dtype=np.int64)
        pos = np.searchsorted(fd, flat)
        pos[pos == fd.size] = 0
        return pos[fd[pos] == flat]

    for edge in edges:
        bc = edge_conditions.get(edge.edge_id)
        if bc is None:
            continue
        kind = bc.kind.strip().lower()
        if kind not in BOUNDARY_KINDS:  # qpsim/models.py:42-49 via solver.py:46-47
            raise ValueError(f"Unsupported boundary condition kind: {kind}")
        if kind in ("neumann", "dirichlet", "robin") and bc.value is None:
            raise ValueError(f"Boundary condition '{kind}' requires a numeric value")
        val = float(bc.value or 0.0)
        aux = float(getattr(bc, "aux_value", None) or 0.0)
        if kind == "reflective":
            diag, s = 0.0, 0.0
        elif kind == "absorbing":
            diag, s = 2.0, 0.0
        elif kind == "dirichlet":
            diag, s = 2.0, 2.0 * val * inv_dx2
        elif kind == "neumann":
            diag, s = 0.0, val * inv_dx
        elif kind == "robin":
            diag, s = val * dx, aux * inv_dx
        else:
            raise BoundaryAssignmentError(f"Unsupported boundary kind: {bc.kind}")
        if not edge.faces:
            continue
        for d, r, c in _face_arrays(edge):
            ok = (r >= 0) & (r < ny) & (c >= 0) & (c < nx)
            pos = positions(d, r[ok] * nx + c[ok])
            covered[d][pos] = True
            face_diag[d][pos] = diag
            face_src[d][pos] = s
    missing = [e.edge_id for e in edges if e.edge_id not in edge_conditions]
    if missing:
        raise BoundaryAssignmentError(
            f"All edges must be assigned boundary conditions before simulation. Missing: {len(missing)}"
        )
    uncovered = {d: faces[d][~covered[d]] for d in _FACE_DIRS}
    if any(u.size for u in uncovered.values()):
        first = min(int(u[0]) for u in uncovered.values() if u.size)   # first offending cell in row-major order
        r, c = divmod(first, nx)
        for d in _FACE_DIRS:
            if uncovered[d].size and first in uncovered[d]:
                raise BoundaryAssignmentError(
                    f"Missing boundary condition for face at cell ({r}, {c}) direction '{d}'."
                )
    bcx = np.zeros((ny, nx))
    bcy = np.zeros((ny, nx))
    src = np.zeros((ny, nx))
    flat_src = src.ravel()
    for d in _FACE_DIRS:
        idx = faces[d]
        tgt = (bcx if d in ("left", "right") else bcy).ravel()
        tgt[idx] += face_diag[d]
        flat_src[idx] += face_src[d]
    return bcx, bcy, src
