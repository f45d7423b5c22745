"""The plain-C restatement (oracle/vof_oracle.c) against the golden vectors that tests/golden/make_golden.py
dumped from the reference's own templates, and against the reference's own known-answer tests.
Bit-exact everywhere (the arithmetic order of the reference is part of the spec).
CPU only."""
import numpy as np
import pytest

from oracle import oraclelib as O
from util import OPERATOR_CASES, golden, same_bits


@pytest.mark.parametrize("dim", [2, 3])
def test_primitives_against_golden(dim):
    g = golden(f"primitives_{dim}d.npz")
    n = len(g["fraction"])
    bad = {"locate": 0, "flux": 0, "vf": 0, "iface": 0}
    for i in range(n):
        lo, h = g["lo"][i], g["h"][i]
        if not same_bits(O.locate(lo, h, g["normal"][i], g["fraction"][i]), g["hs_located"][i]):
            bad["locate"] += 1
        hs = g["hs"][i]
        if not same_bits(O.flux(lo, h, g["face"][i], g["v"][i], hs), g["flux"][i]):
            bad["flux"] += 1
        if not same_bits(O.volume_fraction(lo, h, hs), g["volume_fraction"][i]):
            bad["vf"] += 1
        nv, verts, cen, meas = O.interface(lo, h, hs)
        k = min(nv, 8)
        if nv != g["iface_nverts"][i] or not same_bits(verts[:k], g["iface_verts"][i][:k]) \
                or not same_bits(cen, g["iface_centroid"][i]) or not same_bits(meas, g["iface_measure"][i]):
            bad["iface"] += 1
    assert bad == {"locate": 0, "flux": 0, "vf": 0, "iface": 0}


@pytest.mark.parametrize("case", OPERATOR_CASES)
def test_operators_against_golden(case):
    g = golden(f"operators_{case}.npz")
    n = tuple(int(x) for x in g["n"])
    problem, eps, cfl = int(g["problem"]), float(g["eps"]), float(g["cfl"])
    og = O.OracleGrid(n)
    assert same_bits(og.init(problem), g["u0"])
    u, t, dt = g["u"], float(g["time"]), float(g["dt"])
    ur, tr, dr, _, _ = og.run(O.RECON_YOUNGS, problem, eps, cfl, 4, g["u0"])
    assert same_bits(ur, u) and tr == t and dr == dt
    fl = og.flag(u, eps)
    assert np.array_equal(fl, g["flags"])            # flags: bit-exact integers
    for kind, name in [(O.RECON_YOUNGS, "youngs"), (O.RECON_HF, "hf"), (O.RECON_SWARTZ, "swartz")]:
        if f"hs_{name}" not in g:
            continue
        hs = og.reconstruct(kind, u, fl)
        assert same_bits(hs, g[f"hs_{name}"]), name
        upd, de = og.evolve(hs, fl, problem, t, dt)
        assert same_bits(upd, g[f"update_{name}"]) and de == float(g[f"dtest_{name}"]), name
        if name == "hf":
            k, v = og.curvature(u, hs, fl)
            assert np.array_equal(v, g["kappa_valid"]) and same_bits(k, g["kappa"])
        passes = int(g[f"run_{name}_passes"])
        ur, tr, dr, _, _ = og.run(kind, problem, eps, cfl, passes, g["u0"])
        assert same_bits(ur, g[f"run_{name}_u"]) and tr == float(g[f"run_{name}_time"]) and dr == float(g[f"run_{name}_dt"]), name


def test_c1_known_answer():
    """128^2 Young's rotating circle, 20 passes -> time, next dt and mass recorded from the reference."""
    g = golden("run_c1_128.npz")
    og = O.OracleGrid((128, 128))
    u0 = og.init(O.PROB_ROTATING_CIRCLE)
    assert same_bits(u0, g["u0"])
    u, t, dt, _, mixed = og.run(O.RECON_YOUNGS, O.PROB_ROTATING_CIRCLE, 1e-6, 0.5, 20, u0)
    assert same_bits(u, g["u"])
    assert t == 0.12448621033571523 and dt == 0.0064175380278990047
    assert mixed == int(g["mixed_cell_steps"])
    s = 0.0
    for x in u:
        s += x
    assert s == 1158.1167158193416


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_geometry_kat(dim):
    """dune/vof/test/test-geometry.cc:50-71: first cell at level 2 (h = 1/32) clipped by four half-spaces;
    volumes 0, V/2, V, V/2 to machine epsilon."""
    h = np.full(dim, 1.0 / 32)
    lo = np.zeros(dim)
    vol = float(np.prod(h))
    eps = np.finfo(float).eps
    n1 = np.zeros(dim); n1[0] = -1.0
    corner0, corner1, center = lo.copy(), lo.copy(), lo + 0.5 * h
    corner1[0] += h[0]

    def hs(normal, point):
        return np.concatenate([normal, [-1.0 * float(np.dot(normal, point))]])

    assert O.volume_fraction(lo, h, hs(n1, corner0)) * vol < eps
    assert abs(O.volume_fraction(lo, h, hs(n1, center)) * vol - 0.5 * vol) < eps
    assert abs(O.volume_fraction(lo, h, hs(n1, corner1)) * vol - vol) < eps
    n2 = np.ones(dim); n2 /= np.sqrt(float(dim))
    assert abs(O.volume_fraction(lo, h, hs(n2, center)) * vol - 0.5 * vol) < eps


def test_survey_plane_constants():
    """known answers recorded from the reference in SURVEY.md section 8(c) (unit square / cube)"""
    n2 = np.array([1.0, 2.0]) / np.sqrt(5.0)
    for f, d in [(0.5, -0.67082039324993692), (0.1, -1.0587980740252418), (0.9, -0.282842712474632)]:
        assert O.locate([0, 0], [1, 1], n2, f)[2] == d
    n3 = np.array([0.3, -0.5, 0.8]); n3 /= np.linalg.norm(n3)
    assert O.locate([0, 0, 0], [1, 1, 1], n3, 0.37)[3] == -0.4094120010727027
    assert O.locate([0, 0, 0], [1, 1, 1], [0, 0, 1.0], 0.25)[3] == -0.75


def test_det_cospi_accuracy():
    """the fixed-polynomial cos(pi s) that defines the LeVeque time factor: ~1 ulp on [0, 1/2], exactly periodic"""
    from fractions import Fraction
    import math
    worst = 0.0
    for k in range(0, 513):
        s = k / 1024.0                     # exact binary fractions in [0, 1/2]
        worst = max(worst, abs(O.det_cospi(s) - math.cos(math.pi * s)))
    assert worst < 3e-16
    for k in range(0, 257):
        s = k / 256.0
        assert O.det_cospi(s + 2.0) == O.det_cospi(s) == O.det_cospi(-s)
        assert O.det_cospi(1.0 - s) == -O.det_cospi(s) or s > 0.5
    assert O.det_cospi(0.0) == 1.0 and O.det_cospi(1.0) == -1.0 and O.det_cospi(0.5) == 0.0


def test_flag_edge_cases():
    og = O.OracleGrid((4, 3))
    u = np.zeros(12)
    u[5] = np.nan
    u[6] = 1.0          # full with an empty face neighbour -> mixedfull
    u[1] = 1e-7         # below eps -> empty
    u[2] = 0.5
    fl = og.flag(u, 1e-6)
    assert fl[5] == 99 and fl[6] == 2 and fl[1] == 0 and fl[2] == 1
    u[:] = 1.0
    assert (og.flag(u, 1e-6) == 3).all()


def test_swartz_curvature_golden():
    """oracle SwartzCurvature (2D) against the vectors dumped from the unmodified reference"""
    gk = golden("curvature_swartz_2d.npz")
    for name in ("2d_circle_32", "2d_sflow_40x24", "2d_zalesak_64"):
        g = golden(f"operators_{name}.npz")
        og = O.OracleGrid(tuple(int(x) for x in g["n"]))
        for kname in ("youngs", "hf"):
            assert same_bits(og.curvature_swartz(g[f"hs_{kname}"], g["flags"]), gk[f"{name}_{kname}"]), (name, kname)
        og.close()
