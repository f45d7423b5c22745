"""Flattening of an unstructured grid (2D: triangles and quadrilaterals; 3D: tetrahedra and hexahedra) into the arrays of
`vofx_mesh_desc`.

This is the host-side step `north_star` describes for unstructured grids ("flattens the mesh into device arrays ...
CSR vertex-neighbour stencils"): what a DUNE adaptor does by walking `elements(gridView)` / `intersections(gridView,
entity)` and copying `geometry().corner/center/volume`, `centerUnitOuterNormal`, `integrationOuterNormal` — here for a
mesh given as vertex coordinates and cell→vertex lists.  All geometric quantities are DATA for the library (the
reference also takes them from the grid, never recomputes them), so parity does not depend on how they were rounded.

Conventions (DUNE reference elements): triangle corners (v0, v1, v2), faces (v0,v1), (v0,v2), (v1,v2); quadrilateral
corners (v0, v1, v2, v3) with v2 above v0, faces (v0,v2), (v1,v3), (v0,v1), (v2,v3).  Intersections are stored in face
order; `face_neighbor` is -1 on the domain boundary.  The vertex-neighbour stencil follows
stencil/vertexneighborsstencil.hh:52-94: all cells sharing a vertex, ascending index, self excluded.

3D (`flatten3`): tetrahedron corners (v0..v3), faces (v0,v1,v2), (v0,v1,v3), (v0,v2,v3), (v1,v2,v3); hexahedron corners
v0..v7 (bit d of the corner number = upper side along axis d), faces x-, x+, y-, y+, z-, z+ with their corners in
ascending corner number.  The library forms makePolyhedron( geometry ) (3d/polyhedron.hh:366-404) from the corners and
upwindPolygon3d (3d/upwindpolygon.hh:24-74) from the intersection corners, so `face_corner_offsets` says how many
corners (3 or 4) each intersection has.
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
    dim: int = 2
    face_corner_offsets: np.ndarray = None   # int32 [nf+1], 3D only (2D: two corners per intersection)

    @property
    def n_faces(self) -> int:
        return int(self.face_offsets[-1])

    @classmethod
    def from_npz(cls, z, prefix: str = "mesh_") -> "FlatMesh":
        """a flattened mesh stored field by field (tests/golden/mesh2d_*.npz)"""
        kw = {}
        for k in cls.__dataclass_fields__:
            if prefix + k not in z:
                continue                      # 2D files carry neither dim nor face_corner_offsets
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


_TET_FACES = ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))
_HEX_FACES = ((0, 2, 4, 6), (1, 3, 5, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 1, 2, 3), (4, 5, 6, 7))


def flatten3(vertices, cells) -> FlatMesh:
    """vertices: (nv, 3) floats; cells: sequence of 4-tuples (tetrahedra) or 8-tuples (hexahedra with planar faces) of
    vertex ids in DUNE corner order."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    n_cells = len(cells)
    corner_offsets = np.zeros(n_cells + 1, dtype=np.int32)
    face_offsets = np.zeros(n_cells + 1, dtype=np.int32)
    for c, vs in enumerate(cells):
        if len(vs) not in (4, 8):
            raise ValueError("only tetrahedra and hexahedra")
        corner_offsets[c + 1] = corner_offsets[c] + len(vs)
        face_offsets[c + 1] = face_offsets[c] + (4 if len(vs) == 4 else 6)
    corner_vertex = np.fromiter((v for vs in cells for v in vs), dtype=np.int32, count=int(corner_offsets[-1]))
    corners = vertices[corner_vertex]

    centers = np.empty((n_cells, 3))
    volumes = np.empty(n_cells)
    face_cells = {}
    face_local = []
    for c, vs in enumerate(cells):
        x = vertices[list(vs)]
        centers[c] = x.sum(axis=0) / len(vs)
        if len(vs) == 4:
            volumes[c] = abs(np.linalg.det(np.stack([x[1] - x[0], x[2] - x[0], x[3] - x[0]]))) / 6.0
        else:
            # planar faces: the hexahedron is the union of six tetrahedra around its main diagonal
            vol = 0.0
            for a, b in ((1, 3), (3, 2), (2, 6), (6, 4), (4, 5), (5, 1)):
                vol += abs(np.linalg.det(np.stack([x[a] - x[0], x[b] - x[0], x[7] - x[0]]))) / 6.0
            volumes[c] = vol
        for loc in (_TET_FACES if len(vs) == 4 else _HEX_FACES):
            ids = tuple(vs[k] for k in loc)
            face_cells.setdefault(tuple(sorted(ids)), []).append(c)
            face_local.append((c, ids))

    nf = len(face_local)
    face_neighbor = np.full(nf, -1, dtype=np.int32)
    face_corner_offsets = np.zeros(nf + 1, dtype=np.int32)
    for s, (c, ids) in enumerate(face_local):
        face_corner_offsets[s + 1] = face_corner_offsets[s] + len(ids)
    face_corners = np.empty((int(face_corner_offsets[-1]), 3))
    face_centers = np.empty((nf, 3))
    face_volumes = np.empty(nf)
    normals = np.empty((nf, 3))
    for s, (c, ids) in enumerate(face_local):
        owners = face_cells[tuple(sorted(ids))]
        if len(owners) > 2:
            raise ValueError("non-manifold face")
        for o in owners:
            if o != c:
                face_neighbor[s] = o
        x = vertices[list(ids)]
        face_corners[face_corner_offsets[s]:face_corner_offsets[s + 1]] = x
        face_centers[s] = x.sum(axis=0) / len(ids)
        if len(ids) == 3:
            cr = np.cross(x[1] - x[0], x[2] - x[0]) * 0.5
        else:
            cr = np.cross(x[3] - x[0], x[2] - x[1]) * 0.5      # half the cross product of the diagonals (corners 0,1,2,3 = 00,10,01,11)
        face_volumes[s] = np.sqrt(cr @ cr)
        n = cr / face_volumes[s]
        if n @ (face_centers[s] - centers[c]) < 0:
            n = -n
        normals[s] = n
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
                    c_(face_int_normals), c_(face_volumes), stencil_offsets, stencil, 3, face_corner_offsets)


def perturbed_cube(n: int, kind: str = "tet", seed: int = 0, jitter: float = 0.2):
    """A conforming mesh of the unit cube from an n^3 lattice.  kind "tet": interior vertices moved by up to `jitter`·h, every
    lattice cube split into six tetrahedra around its main diagonal (conforming across cubes); "hex": a rectilinear lattice
    with seeded non-uniform spacings (planar faces); "mixed" is not offered: tetrahedra and hexahedra do not share faces.
    Returns (vertices, cells)."""
    rng = np.random.default_rng(seed)
    h = 1.0 / n
    if kind == "hex":
        axes = []
        for _ in range(3):
            w = 1.0 + (rng.random(n) - 0.5) * 2.0 * jitter
            x = np.concatenate([[0.0], np.cumsum(w)])
            axes.append(x / x[-1])
        vx, vy, vz = np.meshgrid(axes[0], axes[1], axes[2], indexing="ij")
    else:
        xs = np.arange(n + 1) * h
        vx, vy, vz = np.meshgrid(xs, xs, xs, indexing="ij")
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i
    v = np.empty(((n + 1) ** 3, 3))
    for k in range(n + 1):
        for j in range(n + 1):
            for i in range(n + 1):
                v[vid(i, j, k)] = (vx[i, j, k], vy[i, j, k], vz[i, j, k])
    if kind != "hex":
        for k in range(1, n):
            for j in range(1, n):
                for i in range(1, n):
                    v[vid(i, j, k)] += (rng.random(3) - 0.5) * 2.0 * jitter * h
    cells = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                c = [vid(i + (q & 1), j + ((q >> 1) & 1), k + ((q >> 2) & 1)) for q in range(8)]
                if kind == "hex":
                    cells.append(tuple(c))
                else:
                    # Kuhn's subdivision: one tetrahedron per monotone path from corner 0 to corner 7
                    # (positively oriented: makePolyhedron's simplex table then yields outer face normals)
                    for a, b in ((1, 3), (3, 2), (2, 6), (6, 4), (4, 5), (5, 1)):
                        t = (c[0], c[a], c[b], c[7])
                        x = v[list(t)]
                        if np.linalg.det(np.stack([x[1] - x[0], x[2] - x[0], x[3] - x[0]])) < 0:
                            t = (c[0], c[b], c[a], c[7])
                        cells.append(t)
    return v, cells


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
    """dune-vof's FlagOperator / ModifiedYoungsReconstruction / Evolution on an unstructured 2D or 3D grid, computed on the GPU
    through the C ABI (vofx_mesh_*, include/vofx.h).  Host numpy arrays in and out; `run` keeps the colour function on
    the device for a whole time loop (algorithm.hh:68-95).  There is no CPU fallback."""

    def __init__(self, m: FlatMesh, device: int = -1):
        import ctypes as C

        from . import lib as _lib
        self._C, self._lib, self.L = C, _lib, _lib.load()
        self.m = m
        a = lambda x: x.ctypes.data
        self._desc = _lib.MeshDesc(m.dim, m.n_cells, a(m.corner_offsets), a(m.corners), a(m.centers), a(m.volumes),
                                   a(m.face_offsets), a(m.face_neighbor), a(m.face_corners), a(m.face_centers),
                                   a(m.face_unit_normals), a(m.face_int_normals), a(m.face_volumes),
                                   a(m.stencil_offsets), a(m.stencil),
                                   a(m.face_corner_offsets) if m.face_corner_offsets is not None else None)
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
        hs = np.empty((self.m.n_cells, self.m.dim + 1))
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
