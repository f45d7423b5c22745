"""Flattening of an unstructured 2D grid (triangles and quadrilaterals) into the arrays of `vofx_mesh_desc`.

This is the host-side step `north_star` describes for unstructured grids ("flattens the mesh into device arrays ...
CSR vertex-neighbour stencils"): what a DUNE adaptor does by walking `elements(gridView)` / `intersections(gridView,
entity)` and copying `geometry().corner/center/volume`, `centerUnitOuterNormal`, `integrationOuterNormal` — here for a
mesh given as vertex coordinates and cell→vertex lists.  All geometric quantities are DATA for the library (the
reference also takes them from the grid, never recomputes them), so parity does not depend on how they were rounded.

Conventions (DUNE reference elements): triangle corners (v0, v1, v2), faces (v0,v1), (v0,v2), (v1,v2); quadrilateral
corners (v0, v1, v2, v3) with v2 above v0, faces (v0,v2), (v1,v3), (v0,v1), (v2,v3).  Intersections are stored in face
order; `face_neighbor` is -1 on the domain boundary.  The vertex-neighbour stencil follows
stencil/vertexneighborsstencil.hh:52-94: all cells sharing a vertex, ascending index, self excluded.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_TRI_FACES = ((0, 1), (0, 2), (1, 2))
_QUAD_FACES = ((0, 2), (1, 3), (0, 1), (2, 3))
_QUAD_POLYGON = (0, 1, 3, 2)


@dataclass
class FlatMesh:
    n_cells: int
    n_vertices: int
    corner_offsets: np.ndarray      # int32 [n_cells+1]
    corner_vertex: np.ndarray       # int32 [nnz]   vertex id of every cell corner
    corners: np.ndarray             # f64 [nnz, 2]
    centers: np.ndarray             # f64 [n_cells, 2]
    volumes: np.ndarray             # f64 [n_cells]
    face_offsets: np.ndarray        # int32 [n_cells+1]
    face_neighbor: np.ndarray       # int32 [nf]
    face_corners: np.ndarray        # f64 [nf, 2, 2]
    face_centers: np.ndarray        # f64 [nf, 2]
    face_unit_normals: np.ndarray   # f64 [nf, 2]
    face_int_normals: np.ndarray    # f64 [nf, 2]
    face_volumes: np.ndarray        # f64 [nf]
    stencil_offsets: np.ndarray     # int32 [n_cells+1]
    stencil: np.ndarray             # int32

    @property
    def n_faces(self) -> int:
        return int(self.face_offsets[-1])

    @classmethod
    def from_npz(cls, z, prefix: str = "mesh_") -> "FlatMesh":
        """a flattened mesh stored field by field (tests/golden/mesh2d_*.npz)"""
        kw = {}
        for k in cls.__dataclass_fields__:
            a = z[prefix + k]
            kw[k] = int(a) if a.ndim == 0 else np.ascontiguousarray(a)
        return cls(**kw)


def flatten(vertices, cells) -> FlatMesh:
    """vertices: (nv, 2) floats; cells: sequence of 3- or 4-tuples of vertex ids in DUNE corner order."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    n_cells = len(cells)
    corner_offsets = np.zeros(n_cells + 1, dtype=np.int32)
    face_offsets = np.zeros(n_cells + 1, dtype=np.int32)
    for c, vs in enumerate(cells):
        if len(vs) not in (3, 4):
            raise ValueError("only triangles and quadrilaterals")
        corner_offsets[c + 1] = corner_offsets[c] + len(vs)
        face_offsets[c + 1] = face_offsets[c] + len(vs)
    corner_vertex = np.fromiter((v for vs in cells for v in vs), dtype=np.int32, count=int(corner_offsets[-1]))
    corners = vertices[corner_vertex]

    centers = np.empty((n_cells, 2))
    volumes = np.empty(n_cells)
    edge_cells = {}
    face_local = []
    for c, vs in enumerate(cells):
        x = vertices[list(vs)]
        centers[c] = x.sum(axis=0) / len(vs)
        poly = x if len(vs) == 3 else x[list(_QUAD_POLYGON)]
        nxt = np.roll(poly, -1, axis=0)
        volumes[c] = abs(0.5 * np.sum(poly[:, 0] * nxt[:, 1] - nxt[:, 0] * poly[:, 1]))
        for f, (a, b) in enumerate(_TRI_FACES if len(vs) == 3 else _QUAD_FACES):
            edge_cells.setdefault((min(vs[a], vs[b]), max(vs[a], vs[b])), []).append(c)
            face_local.append((c, vs[a], vs[b]))

    nf = len(face_local)
    face_neighbor = np.full(nf, -1, dtype=np.int32)
    face_corners = np.empty((nf, 2, 2))
    for s, (c, va, vb) in enumerate(face_local):
        owners = edge_cells[(min(va, vb), max(va, vb))]
        if len(owners) > 2:
            raise ValueError("non-manifold edge")
        for o in owners:
            if o != c:
                face_neighbor[s] = o
        face_corners[s, 0] = vertices[va]
        face_corners[s, 1] = vertices[vb]
    face_centers = 0.5 * (face_corners[:, 0] + face_corners[:, 1])
    d = face_corners[:, 1] - face_corners[:, 0]
    face_volumes = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    normals = np.stack([d[:, 1], -d[:, 0]], axis=1) / face_volumes[:, None]
    owner = np.repeat(np.arange(n_cells), np.diff(face_offsets))
    outward = np.sum(normals * (face_centers - centers[owner]), axis=1) < 0
    normals[outward] *= -1.0
    face_int_normals = normals * face_volumes[:, None]

    vertex_cells = [[] for _ in range(len(vertices))]
    for c, vs in enumerate(cells):
        for v in vs:
            vertex_cells[v].append(c)
    stencil_offsets = np.zeros(n_cells + 1, dtype=np.int32)
    entries = []
    for c, vs in enumerate(cells):
        nb = sorted({o for v in vs for o in vertex_cells[v] if o != c})
        entries.extend(nb)
        stencil_offsets[c + 1] = len(entries)
    stencil = np.asarray(entries, dtype=np.int32)

    c_ = np.ascontiguousarray
    return FlatMesh(n_cells, len(vertices), corner_offsets, c_(corner_vertex), c_(corners), c_(centers), c_(volumes),
                    face_offsets, face_neighbor, c_(face_corners), c_(face_centers), c_(normals),
                    c_(face_int_normals), c_(face_volumes), stencil_offsets, stencil)


def perturbed_square(n: int, kind: str = "tri", seed: int = 0, jitter: float = 0.25):
    """A conforming mesh of the unit square from an n x n lattice whose interior vertices are moved by up to
    `jitter`·h: kind "tri" splits every quadrilateral into two triangles (alternating diagonals), "quad" keeps the
    quadrilaterals, "mixed" splits a seeded random half of them.  Returns (vertices, cells)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / n
    xs = np.arange(n + 1) * h
    vx, vy = np.meshgrid(xs, xs, indexing="xy")
    v = np.stack([vx.ravel(), vy.ravel()], axis=1)
    interior = ((vx > 0) & (vx < n * h) & (vy > 0) & (vy < n * h)).ravel()
    v[interior] += (rng.random((int(interior.sum()), 2)) - 0.5) * 2.0 * jitter * h
    vid = lambda i, j: j * (n + 1) + i
    cells = []
    split = rng.random((n, n)) < 0.5
    for j in range(n):
        for i in range(n):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            tri = kind == "tri" or (kind == "mixed" and split[j, i])
            if not tri:
                cells.append((a, b, c, d))
            elif (i + j) % 2 == 0:
                cells += [(a, b, d), (a, d, c)]
            else:
                cells += [(a, b, c), (b, d, c)]
    return v, cells


class MeshOperators:
    """dune-vof's FlagOperator / ModifiedYoungsReconstruction / Evolution on an unstructured 2D grid, computed on the GPU
    through the C ABI (vofx_mesh_*, include/vofx.h).  Host numpy arrays in and out; `run` keeps the colour function on
    the device for a whole time loop (algorithm.hh:68-95).  There is no CPU fallback."""

    def __init__(self, m: FlatMesh, device: int = -1):
        import ctypes as C

        from . import lib as _lib
        self._C, self._lib, self.L = C, _lib, _lib.load()
        self.m = m
        a = lambda x: x.ctypes.data
        self._desc = _lib.MeshDesc(2, m.n_cells, a(m.corner_offsets), a(m.corners), a(m.centers), a(m.volumes),
                                   a(m.face_offsets), a(m.face_neighbor), a(m.face_corners), a(m.face_centers),
                                   a(m.face_unit_normals), a(m.face_int_normals), a(m.face_volumes),
                                   a(m.stencil_offsets), a(m.stencil))
        h = C.c_void_p()
        _lib.check(self.L.vofx_mesh_create(C.byref(self._desc), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.L.vofx_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.L.vofx_mesh_launch_count(self._h))

    def flag(self, u, eps):
        u = np.ascontiguousarray(u, dtype=np.float64)
        flags = np.empty(self.m.n_cells, dtype=np.uint8)
        self._lib.check(self.L.vofx_mesh_flag_host(self._h, u.ctypes.data, eps, flags.ctypes.data))
        return flags

    def youngs(self, u, flags):
        u = np.ascontiguousarray(u, dtype=np.float64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        hs = np.empty((self.m.n_cells, 3))
        self._lib.check(self.L.vofx_mesh_reconstruct_host(self._h, u.ctypes.data, flags.ctypes.data, hs.ctypes.data))
        return hs

    def swartz(self, u, flags, max_iterations=10):
        u = np.ascontiguousarray(u, dtype=np.float64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        hs = np.empty((self.m.n_cells, 3))
        self._lib.check(self.L.vofx_mesh_reconstruct_swartz_host(self._h, u.ctypes.data, flags.ctypes.data, hs.ctypes.data,
                                                                 int(max_iterations)))
        return hs

    def evolve(self, hs, flags, face_velocity, dt):
        hs = np.ascontiguousarray(hs, dtype=np.float64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        vel = np.ascontiguousarray(face_velocity, dtype=np.float64)
        upd = np.empty(self.m.n_cells)
        est = self._C.c_double(0.0)
        self._lib.check(self.L.vofx_mesh_evolve_host(self._h, hs.ctypes.data, flags.ctypes.data, vel.ctypes.data, dt,
                                                     upd.ctypes.data, self._C.byref(est)))
        return upd, est.value

    def run(self, u, face_velocity, eps, cfl, passes, time=0.0, dt=0.0, swartz=False, max_iterations=10):
        """upload, `passes` device-resident loop passes, download.  Returns (u, time, dt)."""
        self._lib.check(self.L.vofx_mesh_set_reconstruction(self._h, self._lib.RECON_SWARTZ if swartz else self._lib.RECON_YOUNGS,
                                                            int(max_iterations)))
        u = np.ascontiguousarray(u, dtype=np.float64)
        vel = np.ascontiguousarray(face_velocity, dtype=np.float64)
        self._lib.check(self.L.vofx_mesh_color_upload(self._h, u.ctypes.data))
        self._lib.check(self.L.vofx_mesh_velocity_upload(self._h, vel.ctypes.data))
        t, d = self._C.c_double(time), self._C.c_double(dt)
        self._lib.check(self.L.vofx_mesh_run(self._h, eps, cfl, passes, self._C.byref(t), self._C.byref(d)))
        out = np.empty(self.m.n_cells)
        self._lib.check(self.L.vofx_mesh_color_download(self._h, out.ctypes.data, None))
        return out, t.value, d.value
