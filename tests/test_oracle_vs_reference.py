"""The plain-C restatement against the reference itself, live (oracle/_ref/libvofref.so = /root/reference's own
templates).  Skipped where the prebuilt reference harness is absent.  CPU only; bit-exact."""
import os

import numpy as np
import pytest

from oracle import oraclelib as O
from oracle import reflib as R
from util import same_bits

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libvofref.so not built (needs /root/reference)")


def _ref_locate_forked(lo, h, n, f):
    """the reference runs into undefined behaviour (segfault) for some near-degenerate normals: isolate it"""
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        os.write(w, R.locate(lo, h, n, f).tobytes())
        os._exit(0)
    os.close(w)
    data = b""
    while True:
        chunk = os.read(r, 64)
        if not chunk:
            break
        data += chunk
    os.close(r)
    _, status = os.waitpid(pid, 0)
    if status != 0 or len(data) != 8 * (len(lo) + 1):
        return None
    return np.frombuffer(data, dtype=np.float64).copy()


def test_locate_3d_fuzz():
    rng = np.random.default_rng(101)
    for it in range(6000):
        res = int(rng.choice([64, 128, 512, 1024, 100, 37]))
        h = np.full(3, 1.0 / res)
        lo = rng.integers(0, res, 3) * h
        n = rng.normal(size=3)
        mode = it % 10
        if mode == 1:
            n[rng.integers(0, 3)] = 0.0
        if mode == 2:
            n[rng.integers(0, 3)] *= 1e-9
        if mode == 3:
            k = rng.integers(0, 3); n[:] = 0; n[k] = rng.choice([-1.0, 1.0])
        n /= np.linalg.norm(n)
        f = rng.random() if mode != 5 else float(rng.choice([0.0, 1.0, 1e-7, 1 - 1e-7, 1.2, -0.1]))
        assert same_bits(R.locate(lo, h, n, f), O.locate(lo, h, n, f)), (lo, h, n, f)


def test_locate_3d_near_degenerate():
    """SURVEY.md hard part 1(iii): normals (1, eps, 0): the reference silently returns grossly wrong planes, or
    crashes; the restatement returns the same wrong plane, or NaN where the reference's behaviour is undefined."""
    rng = np.random.default_rng(102)
    R.lib()
    crashed = 0
    for it in range(400):
        res = int(rng.choice([64, 512, 1024]))
        h = np.full(3, 1.0 / res)
        lo = rng.integers(0, res, 3) * h
        k = rng.integers(0, 3)
        n = np.zeros(3)
        n[k] = rng.choice([-1, 1])
        n[(k + 1) % 3] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-17, -5)
        if it % 2:
            n[(k + 2) % 3] = rng.choice([-1, 1]) * 10.0 ** rng.uniform(-17, -5)
        n /= np.linalg.norm(n)
        f = rng.random()
        a = _ref_locate_forked(lo, h, n, f)
        b = O.locate(lo, h, n, f)
        if a is None:
            crashed += 1
            assert np.isnan(b[3])
        else:
            assert same_bits(a, b), (lo, h, n, f)


@pytest.mark.parametrize("dim", [2, 3])
def test_flux_interface_fuzz(dim):
    rng = np.random.default_rng(103 + dim)
    for it in range(4000):
        res = int(rng.choice([64, 128, 512, 4096, 100]))
        h = np.full(dim, 1.0 / res)
        lo = rng.integers(0, res, dim) * h
        n = rng.normal(size=dim)
        if it % 7 == 1:
            n[rng.integers(0, dim)] = 0
        n /= np.linalg.norm(n)
        hs = R.locate(lo, h, n, rng.random())
        if it % 5 == 0:
            hs[dim] += rng.normal() * h[0]
        face = int(rng.integers(0, 2 * dim))
        v = rng.normal(size=dim) * h[0] * 0.5
        assert same_bits(R.flux(lo, h, face, v, hs), O.flux(lo, h, face, v, hs))
        assert same_bits(R.volume_fraction(lo, h, hs), O.volume_fraction(lo, h, hs))
        na, va, ca, ma = R.interface(lo, h, hs)
        nb, vb, cb, mb = O.interface(lo, h, hs)
        assert na == nb and same_bits(va, vb) and same_bits(ca, cb) and same_bits(ma, mb)


@pytest.mark.parametrize("n,problem", [((24, 24), R.PROB_ROTATING_CIRCLE), ((24, 20), R.PROB_SFLOW),
                                       ((48, 48), R.PROB_SLOTTED_CYLINDER), ((12, 12, 12), R.PROB_LEVEQUE),
                                       ((14, 10, 12), R.PROB_SFLOW)])
def test_operators_live(n, problem):
    rg, og = R.RefGrid(n), O.OracleGrid(n)
    u0 = rg.init(problem)
    assert same_bits(u0, og.init(problem))
    eps = 1e-6
    for kind in (R.RECON_YOUNGS, R.RECON_HF, R.RECON_SWARTZ):
        passes = 3 if (kind == R.RECON_SWARTZ and len(n) == 3) else 7
        ur, tr, dr, _, mr = rg.run(kind, problem, eps, 0.5, passes, u0)
        uo, to, do, _, mo = og.run(kind, problem, eps, 0.5, passes, u0)
        assert same_bits(ur, uo) and tr == to and dr == do and mr == mo
    u = ur
    fl = rg.flag(u, eps)
    assert np.array_equal(fl, og.flag(u, eps))
    hs = rg.reconstruct(R.RECON_HF, u, fl)
    assert same_bits(hs, og.reconstruct(O.RECON_HF, u, fl))
    kr, vr = rg.curvature(u, hs, fl)
    ko, vo = og.curvature(u, hs, fl)
    assert np.array_equal(vr, vo) and same_bits(kr, ko)


@pytest.mark.parametrize("n,problem", [((32, 32), R.PROB_ROTATING_CIRCLE), ((40, 24), R.PROB_SFLOW), ((64, 64), R.PROB_SLOTTED_CYLINDER)])
def test_swartz_curvature_live(n, problem):
    """SwartzCurvature (2D, curvature/swartzcurvature.hh:41-173): the C restatement against the unmodified reference"""
    rg, og = R.RefGrid(n), O.OracleGrid(n)
    u = og.init(problem)
    u, *_ = og.run(O.RECON_HF, problem, 1e-6, 0.5, 3, u)
    flags = og.flag(u, 1e-6)
    for kind in (O.RECON_YOUNGS, O.RECON_HF):
        hs = og.reconstruct(kind, u, flags)
        kr = rg.curvature_swartz(hs, flags)
        ko = og.curvature_swartz(hs, flags)
        assert np.count_nonzero(kr) > 10
        assert same_bits(kr, ko), (n, problem, kind, np.flatnonzero(kr != ko)[:5])
