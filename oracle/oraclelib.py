"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes loader for oracle/libvoforacle.so (plain-C restatement).

Same surface as oracle/reflib.py so tests can swap the real reference and the restatement.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvoforacle.so")

RECON_YOUNGS, RECON_HF, RECON_SWARTZ = 0, 1, 2
PROB_ROTATING_CIRCLE, PROB_SFLOW, PROB_SLOTTED_CYLINDER, PROB_LINEAR_WALL, PROB_LEVEQUE = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_bp = C.POINTER(C.c_uint8)


class VoGrid(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("origin", C.c_double * 3), ("h", C.c_double * 3)]


def build(force: bool = False) -> None:
    src = os.path.join(HERE, "vof_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile"), "oracle"], cwd=HERE)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        G = C.POINTER(VoGrid)
        L.vo_flag.argtypes = [G, _dp, C.c_double, _bp]
        L.vo_reconstruct.argtypes = [G, C.c_int, _dp, _bp, _dp, C.c_int]
        L.vo_evolve.argtypes = [G, _dp, _bp, C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.vo_curvature.argtypes = [G, _dp, _dp, _bp, _dp, _bp]
        L.vo_curvature_swartz.argtypes = [G, _dp, _bp, _dp]
        L.vo_init.argtypes = [G, C.c_int, C.c_double, _dp]
        L.vo_circle_interpolation.argtypes = [G, _dp, C.c_double, _dp]
        L.vo_l1error_circle.argtypes = [G, _dp, _bp, _dp, C.c_double, _dp]
        L.vo_face_velocity.argtypes = [G, C.c_int, C.c_double, C.c_int64, C.c_int, _dp]
        L.vo_run.argtypes = [G, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                             C.POINTER(C.c_int64)]
        L.vo_locate.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
        L.vo_volume_fraction.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.vo_flux.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.vo_interface.argtypes = [C.c_int, _dp, _dp, _dp, C.POINTER(C.c_int), _dp, _dp, _dp]
        L.vo_brent_cubic.restype = C.c_double
        L.vo_brent_cubic.argtypes = [C.c_double] * 7
        L.vo_det_cospi.restype = C.c_double
        L.vo_det_cospi.argtypes = [C.c_double]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _b(a):
    return a.ctypes.data_as(_bp)


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle: {what} failed with status {rc}")


class OracleGrid:
    """The restated operators on a Cartesian grid with n cells per direction (x fastest)."""

    def __init__(self, n, lo=None, h=None):
        self.n = tuple(int(x) for x in n)
        self.dim = len(self.n)
        self.lo = _vec(lo if lo is not None else [0.0] * self.dim)
        self.h = _vec(h if h is not None else [1.0 / x for x in self.n])
        self.cells = int(np.prod(self.n))
        g = VoGrid()
        g.dim = self.dim
        for d in range(3):
            g.n[d] = self.n[d] if d < self.dim else 1
            g.origin[d] = self.lo[d] if d < self.dim else 0.0
            g.h[d] = self.h[d] if d < self.dim else 1.0
        self._g = g

    def close(self):
        pass

    def flag(self, u, eps):
        u = _vec(u)
        out = np.empty(self.cells, dtype=np.uint8)
        _check(lib().vo_flag(C.byref(self._g), _d(u), eps, _b(out)), "flag")
        return out

    def reconstruct(self, kind, u, flags, max_iterations=10):
        u = _vec(u)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        hs = np.empty((self.cells, self.dim + 1), dtype=np.float64)
        _check(lib().vo_reconstruct(C.byref(self._g), kind, _d(u), _b(flags), _d(hs), max_iterations), "reconstruct")
        return hs

    def evolve(self, hs, flags, problem, time, dt):
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        upd = np.empty(self.cells, dtype=np.float64)
        dt_est = C.c_double(0.0)
        _check(lib().vo_evolve(C.byref(self._g), _d(hs), _b(flags), problem, time, dt, _d(upd), C.byref(dt_est)), "evolve")
        return upd, dt_est.value

    def curvature(self, u, hs, flags):
        u = _vec(u)
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        kappa = np.empty(self.cells, dtype=np.float64)
        valid = np.empty(self.cells, dtype=np.uint8)
        _check(lib().vo_curvature(C.byref(self._g), _d(u), _d(hs), _b(flags), _d(kappa), _b(valid)), "curvature")
        return kappa, valid

    def curvature_swartz(self, hs, flags):
        """SwartzCurvature (2D), curvature/swartzcurvature.hh:41-173"""
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        kappa = np.empty(self.cells, dtype=np.float64)
        _check(lib().vo_curvature_swartz(C.byref(self._g), _d(hs), _b(flags), _d(kappa)), "curvature_swartz")
        return kappa

    def circle_interpolation(self, center, radius):
        """circleInterpolation (2D), interpolation/circleinterpolation.hh:140-153"""
        u = np.empty(self.cells, dtype=np.float64)
        c = _vec(center)
        _check(lib().vo_circle_interpolation(C.byref(self._g), _d(c), float(radius), _d(u)), "circle_interpolation")
        return u

    def l1error_circle(self, hs, flags, center, radius):
        """l1error against the exact circle (2D), test/errors.hh:62-95"""
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        c = _vec(center)
        out = C.c_double(0.0)
        _check(lib().vo_l1error_circle(C.byref(self._g), _d(hs), _b(flags), _d(c), float(radius), C.byref(out)), "l1error_circle")
        return out.value

    def init(self, problem, time=0.0):
        u = np.empty(self.cells, dtype=np.float64)
        _check(lib().vo_init(C.byref(self._g), problem, time, _d(u)), "init")
        return u

    def face_velocity(self, problem, time, cell, face):
        v = np.empty(self.dim, dtype=np.float64)
        _check(lib().vo_face_velocity(C.byref(self._g), problem, time, int(cell), int(face), _d(v)), "face_velocity")
        return v

    def run(self, kind, problem, eps, cfl, passes, u, time=0.0, dt=0.0, max_iterations=10):
        u = _vec(u).copy()
        t = C.c_double(time)
        d = C.c_double(dt)
        phases = np.zeros(4, dtype=np.float64)
        mixed = C.c_int64(0)
        _check(lib().vo_run(C.byref(self._g), kind, problem, eps, cfl, max_iterations, passes, _d(u), C.byref(t),
                            C.byref(d), _d(phases), C.byref(mixed)), "run")
        return u, t.value, d.value, phases, mixed.value


def locate(lo, h, normal, fraction):
    lo, h, normal = _vec(lo), _vec(h), _vec(normal)
    dim = len(lo)
    out = np.empty(dim + 1)
    lib().vo_locate(dim, _d(lo), _d(h), _d(normal), float(fraction), _d(out))   # status 1 = degenerate (NaN)
    return out


def volume_fraction(lo, h, hs):
    lo, h, hs = _vec(lo), _vec(h), _vec(hs)
    out = C.c_double(0.0)
    lib().vo_volume_fraction(len(lo), _d(lo), _d(h), _d(hs), C.byref(out))
    return out.value


def flux(lo, h, face, v, hs):
    lo, h, v, hs = _vec(lo), _vec(h), _vec(v), _vec(hs)
    out = C.c_double(0.0)
    lib().vo_flux(len(lo), _d(lo), _d(h), int(face), _d(v), _d(hs), C.byref(out))
    return out.value


def interface(lo, h, hs):
    lo, h, hs = _vec(lo), _vec(h), _vec(hs)
    dim = len(lo)
    nv = C.c_int(0)
    verts = np.zeros(16 * dim)
    cen = np.zeros(dim)
    meas = C.c_double(0.0)
    lib().vo_interface(dim, _d(lo), _d(h), _d(hs), C.byref(nv), _d(verts), _d(cen), C.byref(meas))
    return nv.value, verts.reshape(16, dim)[: min(nv.value, 8)].copy(), cen, meas.value


def brent_cubic(A, B, Cc, target, a, b, tol):
    return lib().vo_brent_cubic(A, B, Cc, target, a, b, tol)


def det_cospi(s):
    return lib().vo_det_cospi(float(s))
