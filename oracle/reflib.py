"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes loader for oracle/_ref/libvofref.so.

libvofref.so is the UNMODIFIED reference (header-only dune-vof) instantiated on the oracle's serial Cartesian
grid model, built by oracle/Makefile where /root/reference is mounted.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libvofref.so")

RECON_YOUNGS, RECON_HF, RECON_SWARTZ = 0, 1, 2
PROB_ROTATING_CIRCLE, PROB_SFLOW, PROB_SLOTTED_CYLINDER, PROB_LINEAR_WALL, PROB_LEVEQUE = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_uint8)


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, _ip, _dp, _dp]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_flag.argtypes = [C.c_void_p, _dp, C.c_double, _bp]
        L.ref_reconstruct.argtypes = [C.c_void_p, C.c_int, _dp, _bp, _dp, C.c_int]
        L.ref_evolve.argtypes = [C.c_void_p, _dp, _bp, C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.ref_curvature.argtypes = [C.c_void_p, _dp, _dp, _bp, _dp, _bp]
        L.ref_curvature_swartz.argtypes = [C.c_void_p, _dp, _bp, _dp]
        L.ref_init.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp]
        L.ref_face_velocity.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_int, _dp]
        L.ref_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                              _dp, _dp, _dp, _dp, C.POINTER(C.c_int64)]
        L.ref_locate.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
        L.ref_volume_fraction.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.ref_flux.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp]
        L.ref_interface.argtypes = [C.c_int, _dp, _dp, _dp, _ip, _dp, _dp, _dp]
        L.ref_brent_cubic.restype = C.c_double
        L.ref_brent_cubic.argtypes = [C.c_double] * 7
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _b(a):
    return a.ctypes.data_as(_bp)


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference harness: " + lib().ref_last_error().decode())


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class RefGrid:
    """Reference operators on a Cartesian grid with n cells per direction (x fastest)."""

    def __init__(self, n, lo=None, h=None):
        self.n = tuple(int(x) for x in n)
        self.dim = len(self.n)
        self.lo = _vec(lo if lo is not None else [0.0] * self.dim)
        self.h = _vec(h if h is not None else [1.0 / x for x in self.n])
        self.cells = int(np.prod(self.n))
        nn = (C.c_int * self.dim)(*self.n)
        self._h = lib().ref_create(self.dim, nn, _d(self.lo), _d(self.h))
        if not self._h:
            raise RuntimeError("ref_create failed")

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def flag(self, u, eps):
        u = _vec(u)
        out = np.empty(self.cells, dtype=np.uint8)
        _check(lib().ref_flag(self._h, _d(u), eps, _b(out)))
        return out

    def reconstruct(self, kind, u, flags, max_iterations=10):
        u = _vec(u)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        hs = np.empty((self.cells, self.dim + 1), dtype=np.float64)
        _check(lib().ref_reconstruct(self._h, kind, _d(u), _b(flags), _d(hs), max_iterations))
        return hs

    def evolve(self, hs, flags, problem, time, dt):
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        upd = np.empty(self.cells, dtype=np.float64)
        dt_est = C.c_double(0.0)
        _check(lib().ref_evolve(self._h, _d(hs), _b(flags), problem, time, dt, _d(upd), C.byref(dt_est)))
        return upd, dt_est.value

    def curvature(self, u, hs, flags):
        u = _vec(u)
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        kappa = np.empty(self.cells, dtype=np.float64)
        valid = np.empty(self.cells, dtype=np.uint8)
        _check(lib().ref_curvature(self._h, _d(u), _d(hs), _b(flags), _d(kappa), _b(valid)))
        return kappa, valid

    def curvature_swartz(self, hs, flags):
        hs = _vec(hs)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        kappa = np.empty(self.cells, dtype=np.float64)
        _check(lib().ref_curvature_swartz(self._h, _d(hs), _b(flags), _d(kappa)))
        return kappa

    def init(self, problem, time=0.0):
        u = np.empty(self.cells, dtype=np.float64)
        _check(lib().ref_init(self._h, problem, time, _d(u)))
        return u

    def face_velocity(self, problem, time, cell, face):
        v = np.empty(self.dim, dtype=np.float64)
        _check(lib().ref_face_velocity(self._h, problem, time, int(cell), int(face), _d(v)))
        return v

    def run(self, kind, problem, eps, cfl, passes, u, time=0.0, dt=0.0, max_iterations=10):
        """algorithm.hh:68-95 replayed `passes` times.  Returns (u, time, dt, phase_seconds[4], mixed_cell_steps)."""
        u = _vec(u).copy()
        t = C.c_double(time)
        d = C.c_double(dt)
        phases = np.zeros(4, dtype=np.float64)
        mixed = C.c_int64(0)
        _check(lib().ref_run(self._h, kind, problem, eps, cfl, max_iterations, passes, _d(u), C.byref(t), C.byref(d),
                             _d(phases), C.byref(mixed)))
        return u, t.value, d.value, phases, mixed.value


def locate(lo, h, normal, fraction):
    lo, h, normal = _vec(lo), _vec(h), _vec(normal)
    dim = len(lo)
    out = np.empty(dim + 1)
    _check(lib().ref_locate(dim, _d(lo), _d(h), _d(normal), float(fraction), _d(out)))
    return out


def volume_fraction(lo, h, hs):
    lo, h, hs = _vec(lo), _vec(h), _vec(hs)
    out = C.c_double(0.0)
    _check(lib().ref_volume_fraction(len(lo), _d(lo), _d(h), _d(hs), C.byref(out)))
    return out.value


def flux(lo, h, face, v, hs):
    lo, h, v, hs = _vec(lo), _vec(h), _vec(v), _vec(hs)
    out = C.c_double(0.0)
    _check(lib().ref_flux(len(lo), _d(lo), _d(h), int(face), _d(v), _d(hs), C.byref(out)))
    return out.value


def interface(lo, h, hs):
    lo, h, hs = _vec(lo), _vec(h), _vec(hs)
    dim = len(lo)
    nv = C.c_int(0)
    verts = np.zeros(8 * dim)
    cen = np.zeros(dim)
    meas = C.c_double(0.0)
    _check(lib().ref_interface(dim, _d(lo), _d(h), _d(hs), C.byref(nv), _d(verts), _d(cen), C.byref(meas)))
    return nv.value, verts.reshape(8, dim)[: min(nv.value, 8)].copy(), cen, meas.value


def brent_cubic(A, B, Cc, target, a, b, tol):
    return lib().ref_brent_cubic(A, B, Cc, target, a, b, tol)
