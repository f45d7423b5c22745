"""The product's DEVICE geometry functions (dune-vof_b200/csrc/vofx_geom{2d,3d}.cuh), compiled for the host by
tests/host_geom (test-only), against the oracle: bit-exact on seeded inputs incl. degenerate normals.
This checks the kernels' logic where no GPU exists; the GPU tests repeat it on the device.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oraclelib as O
from util import ROOT, same_bits

HG = os.path.join(ROOT, "tests", "host_geom")


@pytest.fixture(scope="module")
def hg():
    subprocess.check_call(["make", "-s", "-f", os.path.join(HG, "Makefile")], cwd=ROOT)
    L = C.CDLL(os.path.join(HG, "libhostgeom.so"))
    L.hg_det_cospi.restype = C.c_double
    L.hg_det_cospi.argtypes = [C.c_double]
    return L


def _d(a):
    return C.c_void_p(a.ctypes.data)


def _v(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def hg_locate(L, lo, h, n, f):
    lo, h, n = _v(lo), _v(h), _v(n)
    out = np.empty(len(lo) + 1)
    L.hg_locate(len(lo), _d(lo), _d(h), _d(n), C.c_double(f), _d(out))
    return out


def hg_flux(L, lo, h, face, vel, hs):
    lo, h, vel, hs = _v(lo), _v(h), _v(vel), _v(hs)
    out = C.c_double(0)
    L.hg_flux(len(lo), _d(lo), _d(h), int(face), _d(vel), _d(hs), C.byref(out))
    return out.value


def hg_flux_coop(L, lo, h, face, vel, hs):
    """the 8-lane cooperative clip of vofx_coop3d.cuh, lanes emulated phase by phase; returns (flux, path)"""
    lo, h, vel, hs = _v(lo), _v(h), _v(vel), _v(hs)
    out = C.c_double(0)
    path = C.c_int(-1)
    L.hg_flux_coop(_d(lo), _d(h), int(face), _d(vel), _d(hs), C.byref(out), C.byref(path))
    return out.value, path.value


def hg_vf(L, lo, h, hs):
    lo, h, hs = _v(lo), _v(h), _v(hs)
    out = C.c_double(0)
    L.hg_volume_fraction(len(lo), _d(lo), _d(h), _d(hs), C.byref(out))
    return out.value


def hg_iface(L, lo, h, hs):
    lo, h, hs = _v(lo), _v(h), _v(hs)
    dim = len(lo)
    nv = C.c_int(0)
    verts = np.zeros(16 * dim)
    cen = np.zeros(dim)
    m = C.c_double(0)
    L.hg_interface(dim, _d(lo), _d(h), _d(hs), C.byref(nv), _d(verts), _d(cen), C.byref(m))
    return nv.value, verts.reshape(16, dim)[: min(nv.value, 8)].copy(), cen, m.value


@pytest.mark.parametrize("dim", [2, 3])
def test_geometry_bit_exact(hg, dim):
    rng = np.random.default_rng(70 + dim)
    paths = {}
    for it in range(6000):
        res = int(rng.choice([64, 128, 512, 1024, 4096, 100, 37]))
        h = np.full(dim, 1.0 / res)
        lo = rng.integers(0, res, dim) * h
        n = rng.normal(size=dim)
        mode = it % 12
        if mode == 1:
            n[rng.integers(0, dim)] = 0
        if mode == 2:
            n[rng.integers(0, dim)] *= 1e-9
        if mode == 3:
            k = rng.integers(0, dim); n[:] = 0; n[k] = rng.choice([-1.0, 1.0])
        if mode == 4:
            n = rng.choice([-1.0, 1.0], size=dim)
        if mode == 6 and dim == 3:   # SURVEY.md hard part 1(iii): nearly axis-aligned normals
            k = rng.integers(0, 3); n[:] = 0; n[k] = rng.choice([-1, 1])
            n[(k + 1) % 3] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-17, -5)
            if it % 2:
                n[(k + 2) % 3] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-17, -5)
        n /= np.linalg.norm(n)
        f = rng.random()
        if mode == 5:
            f = float(rng.choice([0.0, 1.0, 1e-7, 1 - 1e-7, 1e-6, 0.5, 1.2, -0.1]))
        a = O.locate(lo, h, n, f)
        assert same_bits(a, hg_locate(hg, lo, h, n, f)), (dim, mode, lo, h, n, f)
        hs = a.copy()
        if np.isnan(hs[-1]):
            hs[-1] = -np.dot(n, lo + 0.5 * h)
        if it % 5 == 0:
            hs[-1] += rng.normal() * h[0]
        face = int(rng.integers(0, 2 * dim))
        vel = rng.normal(size=dim) * h[0] * 0.5
        if it % 11 == 3:
            vel[rng.integers(0, dim)] = 0
        fo = O.flux(lo, h, face, vel, hs)
        assert same_bits(fo, hg_flux(hg, lo, h, face, vel, hs))
        if dim == 3:
            fc, path = hg_flux_coop(hg, lo, h, face, vel, hs)
            assert same_bits(fo, fc), (path, lo, h, face, vel, hs)
            paths[path] = paths.get(path, 0) + 1
        assert same_bits(O.volume_fraction(lo, h, hs), hg_vf(hg, lo, h, hs))
        na, va, ca, ma = O.interface(lo, h, hs)
        nb, vb, cb, mb = hg_iface(hg, lo, h, hs)
        assert na == nb and same_bits(va, vb) and same_bits(ca, cb) and same_bits(ma, mb)
    if dim == 3:
        # the tables carry the clips, including the patterns with nodes ON the plane (cells located at fraction 0 or 1);
        # the serial path is only for the few patterns the table cannot represent
        assert paths.get(0, 0) > 2000 and paths.get(2, 0) < 0.01 * sum(paths.values()), paths
        print("cooperative clip paths (0 table, 1 empty, 2 serial):", paths)


def test_clip_on_plane_patterns(hg):
    """nodes of the swept box exactly on the plane (3d/intersect.hh:104-108): planes through cell nodes, edges and faces"""
    irregular, empty = C.c_int(0), C.c_int(0)
    total = hg.hg_clip_boundary_table(C.byref(irregular), C.byref(empty))
    assert total == 6561 and irregular.value < 0.5 * total, (irregular.value, empty.value)
    rng = np.random.default_rng(5)
    paths = {}
    for it in range(3000):
        res = int(rng.choice([64, 128, 512]))
        h = np.full(3, 1.0 / res)
        lo = rng.integers(0, res, 3) * h
        face = int(rng.integers(0, 6))
        vel = rng.normal(size=3) * h[0] * 0.5
        mode = it % 4
        if mode == 0:      # a cell filled to fraction 0 or 1: the plane passes through one of its nodes
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            hs = O.locate(lo, h, n, float(rng.choice([0.0, 1.0, 1.3, -0.2])))
        elif mode == 1:    # axis-aligned plane through a cell face
            k = int(rng.integers(0, 3))
            n = np.zeros(3)
            n[k] = rng.choice([-1.0, 1.0])
            hs = np.append(n, -n[k] * (lo[k] + rng.integers(0, 2) * h[k]))
            vel[int(rng.integers(0, 3))] = 0.0
        elif mode == 2:    # diagonal plane through an edge of the cell (exactly representable)
            k = int(rng.integers(0, 3))
            n = np.ones(3) * rng.choice([-1.0, 1.0], size=3)
            n[k] = 0.0
            corner = lo + rng.integers(0, 2, 3) * h
            hs = np.append(n, -np.dot(n, corner))
        else:              # plane through three nodes of the swept box
            n = np.array([1.0, 1.0, 1.0]) * rng.choice([-1.0, 1.0], size=3)
            corner = lo + rng.integers(0, 2, 3) * h
            hs = np.append(n, -np.dot(n, corner))
            vel = np.zeros(3)
            vel[face // 2] = (1 if face % 2 else -1) * h[0] * 0.25
        if np.isnan(hs[-1]):
            continue
        fo = O.flux(lo, h, face, vel, hs)
        fc, path = hg_flux_coop(hg, lo, h, face, vel, hs)
        assert same_bits(fo, fc), (mode, path, lo, h, face, vel, hs)
        paths[(mode, path)] = paths.get((mode, path), 0) + 1
    print("on-plane clips by (mode, path):", dict(sorted(paths.items())))
    assert sum(v for (m, p), v in paths.items() if p == 0) > 1000


def test_det_cospi_matches_oracle(hg):
    for s in np.linspace(-2.0, 5.0, 2001):
        assert hg.hg_det_cospi(float(s)) == O.det_cospi(float(s))


def hg_locate_coop(L, lanes, lo, h, n, f):
    lo, h, n = _v(lo), _v(h), _v(n)
    d = C.c_double(0)
    st = L.hg_locate_coop(int(lanes), _d(lo), _d(h), _d(n), C.c_double(f), C.byref(d))
    return st, d.value


@pytest.mark.parametrize("lanes", [8, 4])
def test_cooperative_locate_bit_exact(hg, lanes):
    """locate_coop (vofx_coop3d.cuh: sweep topology from the table of 96 height orders), lanes emulated, against the
    oracle's locateHalfSpace on random and near-degenerate normals; the table path must take the generic cases"""
    assert hg.hg_sweep_table_cases() == 96
    rng = np.random.default_rng(500 + lanes)
    paths = {0: 0, 2: 0}
    for it in range(6000):
        res = int(rng.choice([64, 128, 512, 1024, 100, 37]))
        h = np.full(3, 1.0 / res)
        if it % 7 == 0:
            h = h * rng.uniform(0.5, 2.0, 3)            # anisotropic cells
        lo = rng.integers(0, res, 3) * h
        n = rng.normal(size=3)
        mode = it % 12
        if mode == 1:
            n[rng.integers(0, 3)] = 0
        if mode == 2:
            n[rng.integers(0, 3)] *= 1e-9
        if mode == 3:
            k = rng.integers(0, 3); n[:] = 0; n[k] = rng.choice([-1.0, 1.0])
        if mode == 4:
            n = rng.choice([-1.0, 1.0], size=3)
        if mode == 6:
            k = rng.integers(0, 3); n[:] = 0; n[k] = rng.choice([-1, 1])
            n[(k + 1) % 3] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-17, -5)
        n /= np.linalg.norm(n)
        f = rng.random()
        if mode == 5:
            f = float(rng.choice([0.0, 1.0, 1e-7, 1 - 1e-7, 1e-6, 0.5, 1.2, -0.1]))
        st, d = hg_locate_coop(hg, lanes, lo, h, n, f)
        paths[st] += 1
        if st == 0:
            ref = O.locate(lo, h, n, f)
            assert same_bits(ref[-1], d), (lanes, mode, lo, h, n, f, ref[-1], d)
    assert paths[0] > 3500, paths
    print("cooperative locate paths (0 table, 2 serial):", paths)


@pytest.mark.parametrize("lanes", [8, 4])
def test_cooperative_youngs_bit_exact(hg, lanes):
    """Young's normal + plane constant of every mixed cell of small 3D problems, cooperative path vs the oracle"""
    for n, problem in (((16, 16, 16), O.PROB_LEVEQUE), ((20, 12, 16), O.PROB_SFLOW), ((12, 12, 12), O.PROB_LINEAR_WALL)):
        og = O.OracleGrid(n)
        u = og.init(problem)
        u, *_ = og.run(O.RECON_YOUNGS, problem, 1e-6, 0.5, 3, u)          # a few passes: over/undershoots appear
        flags = og.flag(u, 1e-6)
        hs = og.reconstruct(O.RECON_YOUNGS, u, flags).reshape(-1, 4)
        nn = np.asarray(n, dtype=np.int32)
        uu = _v(u)
        status = {0: 0, 1: 0, 2: 0}
        for c in np.flatnonzero((flags == 1) | (flags == 2)):
            ijk = np.asarray([c % n[0], (c // n[0]) % n[1], c // (n[0] * n[1])], dtype=np.int32)
            out = np.empty(4)
            st = hg.hg_youngs_coop(int(lanes), C.c_void_p(nn.ctypes.data), _d(uu), C.c_void_p(ijk.ctypes.data), _d(out))
            assert st in (0, 1, 2)
            status[st] += 1
            assert same_bits(out, hs[c]), (n, problem, c, st, out, hs[c])
        print(n, problem, "cooperative Young's status counts:", status)
        og.close()


@pytest.mark.parametrize("lanes", [8, 4])
def test_cooperative_swartz_bit_exact(hg, lanes):
    """ModifiedSwartz (3D) of every mixed cell, cooperative path (lanes emulated) vs the oracle"""
    for n, problem in (((12, 12, 12), O.PROB_LEVEQUE), ((16, 10, 12), O.PROB_SFLOW)):
        og = O.OracleGrid(n)
        u = og.init(problem)
        u, *_ = og.run(O.RECON_YOUNGS, problem, 1e-6, 0.5, 2, u)
        flags = og.flag(u, 1e-6)
        hy = og.reconstruct(O.RECON_YOUNGS, u, flags).reshape(-1, 4)
        hs = og.reconstruct(O.RECON_SWARTZ, u, flags).reshape(-1, 4)
        nn = np.asarray(n, dtype=np.int32)
        uu = _v(u)
        ff = np.ascontiguousarray(flags, dtype=np.uint8)
        cells = np.flatnonzero((flags == 1) | (flags == 2))
        for c in cells:
            ijk = np.asarray([c % n[0], (c // n[0]) % n[1], c // (n[0] * n[1])], dtype=np.int32)
            rec = hy[c].copy()
            hg.hg_swartz_coop(int(lanes), C.c_void_p(nn.ctypes.data), _d(uu), C.c_void_p(ff.ctypes.data), C.c_void_p(ijk.ctypes.data), 10, _d(rec))
            assert same_bits(rec, hs[c]), (n, problem, c, rec, hs[c])
        print(n, problem, "cooperative Swartz cells checked:", len(cells))
        og.close()


def test_unstructured_3d_primitives_bit_exact(hg):
    """the topology policies of vofx_geom3d.cuh (topo::Tet = makePolyhedron( simplex ), topo::Prism = upwindPolygon3d of a
    triangle, topo::Cube) compiled for the host, against the oracle's general polyhedra: locateHalfSpace on tetrahedra and
    boxes, truncVolume of swept triangles and quadrilaterals; random, axis-aligned and degenerate inputs"""
    L = O.lib()
    rng = np.random.default_rng(11)
    checked = 0
    for it in range(3000):
        nc = 4 if it % 2 == 0 else 8
        if nc == 4:
            x = rng.random((4, 3)) * rng.choice([1.0, 0.01])
            if np.linalg.det(np.stack([x[1] - x[0], x[2] - x[0], x[3] - x[0]])) < 0:
                x[[1, 2]] = x[[2, 1]]
            if it % 10 == 0:
                x[:3, 2] = x[0, 2]              # a face at constant z: the N = 3 "face" start of the sweep
        else:
            lo, h = rng.random(3), rng.random(3) * 0.1 + 0.01
            x = np.array([[lo[d] + ((i >> d) & 1) * h[d] for d in range(3)] for i in range(8)])
        x = np.ascontiguousarray(x)
        n = rng.normal(size=3)
        if it % 7 == 0:
            n[rng.integers(0, 3)] = 0
        if it % 11 == 0:
            n[:] = 0
            n[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])
        n /= np.linalg.norm(n)
        f = rng.random()
        if it % 13 == 0:
            f = float(rng.choice([0.0, 1.0, 1e-7, 1 - 1e-7, 0.5]))
        o1, o2 = np.zeros(4), np.zeros(4)
        rc = L.vo_locate_corners(nc, _d(x), _d(n), C.c_double(f), _d(o1))
        hg.hg_locate_corners(nc, _d(x), _d(n), C.c_double(f), _d(o2))
        if rc == 0 and not (np.isnan(o1[3]) and np.isnan(o2[3])):
            assert np.array_equal(o1.view(np.uint64), o2.view(np.uint64)), (nc, it, o1, o2)
            checked += 1
        hs = np.array([n[0], n[1], n[2], o1[3] if rc == 0 and np.isfinite(o1[3]) else 0.0])
        for fnc in (3, 4):
            if fnc == 3:
                c = np.ascontiguousarray(rng.random((3, 3)) * 0.1 + x[0])
            else:
                a, b = rng.random(3) * 0.1, rng.random(3) * 0.1
                c = np.ascontiguousarray(np.array([x[0], x[0] + a, x[0] + b, x[0] + a + b]))
            v = rng.normal(size=3) * 0.02
            if it % 9 == 0:
                v[rng.integers(0, 3)] = 0
            q1, q2 = C.c_double(0), C.c_double(0)
            L.vo_flux_corners(fnc, _d(c), _d(v), _d(hs), C.byref(q1))
            hg.hg_flux_corners(fnc, _d(c), _d(v), _d(hs), C.byref(q2))
            assert same_bits(q1.value, q2.value), (fnc, it, q1.value, q2.value)
            checked += 1
    assert checked > 8000
