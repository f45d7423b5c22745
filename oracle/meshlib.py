"""TEST INFRASTRUCTURE ONLY (oracle/): the unstructured-mesh operators of the reference (oracle/_ref/libvofrefmesh.so for
2D meshes, libvofrefmesh3.so for 3D ones: the unmodified dune-vof headers on oracle/shim/dune/grid/common/meshgrid.hh) and of the plain-C restatement
(vo_mesh_* in oracle/vof_oracle.c), behind one surface.  Both take a flattened mesh (dune_vof_b200.mesh.FlatMesh).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oraclelib

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(HERE, "_ref", "libvofrefmesh.so")
REF3_LIB_PATH = os.path.join(HERE, "_ref", "libvofrefmesh3.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _b(a):
    return a.ctypes.data_as(_bp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def ref_available(dim: int = 2) -> bool:
    return os.path.exists(REF3_LIB_PATH if dim == 3 else REF_LIB_PATH)


class _RefLib:
    """the harness of one dimension behind dimension-free names (symbols refmesh_* / refmesh3_*)"""

    def __init__(self, dim):
        L = C.CDLL(REF3_LIB_PATH if dim == 3 else REF_LIB_PATH)
        pre = "refmesh3_" if dim == 3 else "refmesh_"
        f = lambda name: getattr(L, pre + name)
        self.last_error = f("last_error")
        self.last_error.restype = C.c_char_p
        self.create = f("create")
        self.create.restype = C.c_void_p
        self.create.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _ip]
        self.destroy = f("destroy")
        self.destroy.argtypes = [C.c_void_p]
        self.stencil = f("stencil")
        self.stencil.restype = C.c_int64
        self.stencil.argtypes = [C.c_void_p, _ip, _ip, C.c_int64]
        self.flag = f("flag")
        self.flag.argtypes = [C.c_void_p, _dp, C.c_double, _bp]
        self.youngs = f("youngs")
        self.youngs.argtypes = [C.c_void_p, _dp, _bp, _dp]
        self.evolve = f("evolve")
        self.evolve.argtypes = [C.c_void_p, _dp, _bp, _dp, C.c_double, _dp, _dp]
        if dim == 2:
            self.swartz = f("swartz")
            self.swartz.argtypes = [C.c_void_p, _dp, _bp, _dp, C.c_int]
        self.locate = f("locate")
        self.locate.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_double, _dp]


_ref = {}


def _reflib(dim=2):
    if dim not in _ref:
        _ref[dim] = _RefLib(dim)
    return _ref[dim]


class RefMesh:
    """the reference's FlagOperator / ModifiedYoungsReconstruction / Evolution on a flattened mesh"""

    def __init__(self, m):
        self.m = m
        self.dim = m.dim
        L = self.R = _reflib(m.dim)
        self._h = L.create(m.n_cells, m.n_vertices, _i(m.corner_offsets), _i(m.corner_vertex), _d(m.corners),
                           _d(m.centers), _d(m.volumes), _i(m.face_offsets), _i(m.face_neighbor),
                           _d(m.face_corners), _d(m.face_centers), _d(m.face_unit_normals),
                           _d(m.face_int_normals), _d(m.face_volumes),
                           _i(m.face_corner_offsets) if m.face_corner_offsets is not None else None)

    def close(self):
        if self._h:
            self.R.destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("reference mesh harness: " + self.R.last_error().decode())

    def stencil(self):
        offsets = np.empty(self.m.n_cells + 1, dtype=np.int32)
        n = self.R.stencil(self._h, _i(offsets), _i(np.empty(1, dtype=np.int32)), 0)
        entries = np.empty(max(int(n), 1), dtype=np.int32)
        self.R.stencil(self._h, _i(offsets), _i(entries), n)
        return offsets, entries[:n]

    def flag(self, u, eps):
        u = _f64(u)
        out = np.empty(self.m.n_cells, dtype=np.uint8)
        self._check(self.R.flag(self._h, _d(u), eps, _b(out)))
        return out

    def youngs(self, u, flags):
        u, flags = _f64(u), _u8(flags)
        hs = np.empty((self.m.n_cells, self.m.dim + 1))
        self._check(self.R.youngs(self._h, _d(u), _b(flags), _d(hs)))
        return hs

    def swartz(self, u, flags, max_iterations=10):
        u, flags = _f64(u), _u8(flags)
        hs = np.empty((self.m.n_cells, self.m.dim + 1))
        self._check(self.R.swartz(self._h, _d(u), _b(flags), _d(hs), max_iterations))
        return hs

    def evolve(self, hs, flags, face_velocity, dt):
        hs, flags, face_velocity = _f64(hs), _u8(flags), _f64(face_velocity)
        upd = np.empty(self.m.n_cells)
        est = C.c_double(0.0)
        self._check(self.R.evolve(self._h, _d(hs), _b(flags), _d(face_velocity), dt, _d(upd), C.byref(est)))
        return upd, est.value

    def locate(self, cell, normal, fraction):
        out = np.empty(self.m.dim + 1)
        self._check(self.R.locate(self._h, int(cell), _d(_f64(normal)), fraction, _d(out)))
        return out


class VoMesh(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("corner_offsets", _ip), ("corners", _dp), ("centers", _dp), ("volumes", _dp),
                ("face_offsets", _ip), ("face_neighbor", _ip), ("face_corners", _dp), ("face_centers", _dp),
                ("face_unit_normals", _dp), ("face_int_normals", _dp), ("face_volumes", _dp),
                ("stencil_offsets", _ip), ("stencil", _ip), ("dim", C.c_int32), ("face_corner_offsets", _ip)]


class OracleMesh:
    """the plain-C restatement (vo_mesh_*) on a flattened mesh"""

    def __init__(self, m):
        self.m = m
        L = oraclelib.lib()
        M = C.POINTER(VoMesh)
        L.vo_mesh_flag.argtypes = [M, _dp, C.c_double, _bp]
        L.vo_mesh_youngs.argtypes = [M, _dp, _bp, _dp]
        L.vo_mesh_evolve.argtypes = [M, _dp, _bp, _dp, C.c_double, _dp, _dp]
        L.vo_mesh_swartz.argtypes = [M, _dp, _bp, _dp, C.c_int]
        L.vo_mesh_locate.argtypes = [M, C.c_int64, _dp, C.c_double, _dp]
        self.L = L
        self.desc = VoMesh(m.n_cells, _i(m.corner_offsets), _d(m.corners), _d(m.centers), _d(m.volumes),
                           _i(m.face_offsets), _i(m.face_neighbor), _d(m.face_corners), _d(m.face_centers),
                           _d(m.face_unit_normals), _d(m.face_int_normals), _d(m.face_volumes),
                           _i(m.stencil_offsets), _i(m.stencil), m.dim,
                           _i(m.face_corner_offsets) if m.face_corner_offsets is not None else None)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle: {what} failed with status {rc}")

    def flag(self, u, eps):
        u = _f64(u)
        out = np.empty(self.m.n_cells, dtype=np.uint8)
        self._check(self.L.vo_mesh_flag(C.byref(self.desc), _d(u), eps, _b(out)), "vo_mesh_flag")
        return out

    def youngs(self, u, flags):
        u, flags = _f64(u), _u8(flags)
        hs = np.empty((self.m.n_cells, self.m.dim + 1))
        self._check(self.L.vo_mesh_youngs(C.byref(self.desc), _d(u), _b(flags), _d(hs)), "vo_mesh_youngs")
        return hs

    def swartz(self, u, flags, max_iterations=10):
        u, flags = _f64(u), _u8(flags)
        hs = np.empty((self.m.n_cells, self.m.dim + 1))
        self._check(self.L.vo_mesh_swartz(C.byref(self.desc), _d(u), _b(flags), _d(hs), max_iterations), "vo_mesh_swartz")
        return hs

    def evolve(self, hs, flags, face_velocity, dt):
        hs, flags, face_velocity = _f64(hs), _u8(flags), _f64(face_velocity)
        upd = np.empty(self.m.n_cells)
        est = C.c_double(0.0)
        self._check(self.L.vo_mesh_evolve(C.byref(self.desc), _d(hs), _b(flags), _d(face_velocity), dt, _d(upd),
                                          C.byref(est)), "vo_mesh_evolve")
        return upd, est.value

    def locate(self, cell, normal, fraction):
        out = np.empty(self.m.dim + 1)
        self._check(self.L.vo_mesh_locate(C.byref(self.desc), int(cell), _d(_f64(normal)), fraction, _d(out)),
                    "vo_mesh_locate")
        return out


def circle_fractions(m, center=(0.5, 0.75), radius=0.15, samples=24):
    """initial colour function of a disc by sub-sampling every cell (deterministic; only an INPUT for the parity
    tests — any values in [0, 1] with an O(h) band would do)"""
    u = np.zeros(m.n_cells)
    t = (np.arange(samples) + 0.5) / samples
    for c in range(m.n_cells):
        b, e = m.corner_offsets[c], m.corner_offsets[c + 1]
        x = m.corners[b:e]
        d = np.hypot(x[:, 0] - center[0], x[:, 1] - center[1])
        if d.max() < radius:
            u[c] = 1.0
        elif d.min() > radius + 0.2 * np.sqrt(m.volumes[c]) and np.hypot(*(m.centers[c] - center)) > radius:
            u[c] = 0.0
        else:
            if e - b == 3:
                s, r = np.meshgrid(t, t)
                keep = s + r <= 1.0
                pts = x[0] + s[keep][:, None] * (x[1] - x[0]) + r[keep][:, None] * (x[2] - x[0])
            else:
                s, r = np.meshgrid(t, t)
                s, r = s.ravel()[:, None], r.ravel()[:, None]
                pts = (1 - s) * (1 - r) * x[0] + s * (1 - r) * x[1] + (1 - s) * r * x[2] + s * r * x[3]
            u[c] = np.mean(np.hypot(pts[:, 0] - center[0], pts[:, 1] - center[1]) < radius)
    return u


def sphere_fractions(m, center=(0.5, 0.6, 0.45), radius=0.3, samples=200, seed=5):
    """initial colour function of a ball on a 3D mesh by seeded sampling of every cell near the surface (convex combinations
    of the corners; deterministic; only an INPUT for the parity tests, stored with the golden vectors)"""
    u = np.zeros(m.n_cells)
    rng = np.random.default_rng(seed)
    center = np.asarray(center)
    for c in range(m.n_cells):
        b, e = m.corner_offsets[c], m.corner_offsets[c + 1]
        x = m.corners[b:e]
        d = np.linalg.norm(x - center, axis=1)
        if d.max() < radius:
            u[c] = 1.0
        elif d.min() > radius + 0.5 * m.volumes[c] ** (1.0 / 3.0):
            u[c] = 0.0
        else:
            w = rng.dirichlet(np.ones(e - b), size=samples)
            u[c] = np.mean(np.linalg.norm(w @ x - center, axis=1) < radius)
    return u


def rotation_velocity3(m):
    """a rigid rotation about the z axis plus a drift along z, sampled at the intersection centres (an INPUT)"""
    fc = m.face_centers
    return np.ascontiguousarray(np.stack([fc[:, 1] - 0.5, 0.5 - fc[:, 0], 0.3 + 0.0 * fc[:, 0]], axis=1))


def rotation_velocity(m):
    """RotatingCircle's field (y - 1/2, 1/2 - x)·2π/10 sampled at the intersection centres (an INPUT, see above)"""
    fc = m.face_centers
    return np.ascontiguousarray(np.stack([fc[:, 1] - 0.5, 0.5 - fc[:, 0]], axis=1) * (2.0 * np.pi / 10.0))


def run(ops, m, u, face_velocity, eps, cfl, passes, dt=0.0, swartz=False):
    """algorithm.hh:68-95 on a mesh with any of RefMesh / OracleMesh / the device operators"""
    u = _f64(u).copy()
    time = 0.0
    for _ in range(passes):
        flags = ops.flag(u, eps)
        hs = ops.swartz(u, flags) if swartz else ops.youngs(u, flags)
        upd, est = ops.evolve(hs, flags, face_velocity, dt)
        u = u + upd
        if dt > 0.0:
            time += dt
        dt = est * cfl
    return u, time, dt
