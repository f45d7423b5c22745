"""Parity of the CUDA path (through the C ABI) with the golden vectors dumped from the reference and with the
oracle on seeded inputs.  Bit-exact: flags and indices are integers, and for the fp64 quantities the arithmetic
order of the reference is reproduced (tolerance 0 ulp; north_star asks for 1e-12 relative)."""
import ctypes as C

import numpy as np
import pytest

from util import OPERATOR_CASES, golden, n_mismatch, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vofx():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    import dune_vof_b200.operators as ops
    return ops


def _lib():
    from dune_vof_b200 import lib as vlib
    return vlib.load(), vlib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_division_and_sqrt_fast_paths():
    """vdiv / vdiv3 / vdiv2 / vsqrt (vofx_math.cuh) are the compiler's own division and square-root fast paths written as
    straight-line code: they must return the native operators' bits for every operand (2^27 operand sets per seed:
    random mantissas, exponents up to +-1020, signed zeros, powers of two, exact squares)"""
    L, vlib = _lib()
    for seed in (1, 0xDEADBEEF, 2 ** 63 + 12345):
        bad = (C.c_int64 * 3)()
        vlib.check(L.vofx_test_arith(1 << 27, seed, bad))
        assert list(bad) == [0, 0, 0], (seed, list(bad))


@pytest.mark.parametrize("dim", [2, 3])
def test_primitives_against_golden(dim):
    L, vlib = _lib()
    g = golden(f"primitives_{dim}d.npz")
    n = len(g["fraction"])
    lo, h = np.ascontiguousarray(g["lo"]), np.ascontiguousarray(g["h"])
    out = np.empty((n, dim + 1))
    normal, fraction = np.ascontiguousarray(g["normal"]), np.ascontiguousarray(g["fraction"])   # keep alive across the call
    vlib.check(L.vofx_test_locate(dim, n, _p(lo), _p(h), _p(normal), _p(fraction), _p(out)))
    assert n_mismatch(out, g["hs_located"]) == 0
    hs = np.ascontiguousarray(g["hs"])
    flux = np.empty(n)
    face = np.ascontiguousarray(g["face"], dtype=np.int32)
    vel = np.ascontiguousarray(g["v"])
    vlib.check(L.vofx_test_flux(dim, n, _p(lo), _p(h), _p(face), _p(vel), _p(hs), _p(flux)))
    assert n_mismatch(flux, g["flux"]) == 0
    nv = np.empty(n, dtype=np.int32); cen = np.empty((n, dim)); meas = np.empty(n)
    vlib.check(L.vofx_test_interface(dim, n, _p(lo), _p(h), _p(hs), _p(nv), _p(cen), _p(meas)))
    assert np.array_equal(nv, g["iface_nverts"])
    assert n_mismatch(cen, g["iface_centroid"]) == 0 and n_mismatch(meas, g["iface_measure"]) == 0


@pytest.mark.parametrize("case", OPERATOR_CASES)
def test_operators_against_golden(vofx, case):
    import torch
    g = golden(f"operators_{case}.npz")
    n = tuple(int(x) for x in g["n"])
    problem, eps, cfl = int(g["problem"]), float(g["eps"]), float(g["cfl"])
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    dev = grid.device
    u = torch.tensor(g["u"], device=dev)
    t, dt = float(g["time"]), float(g["dt"])

    flags = grid.empty("flags")
    vofx.FlagOperator(grid, eps)(u, flags)
    assert np.array_equal(flags.cpu().numpy(), g["flags"])

    recs = grid.empty("reconstructions")
    upd = grid.empty("update")
    for cls, name in [(vofx.ModifiedYoungsReconstruction, "youngs"), (vofx.HeightFunctionReconstruction, "hf"),
                      (vofx.ModifiedSwartzReconstruction, "swartz")]:
        if f"hs_{name}" not in g:
            continue
        cls(grid)(u, recs, flags)
        assert n_mismatch(recs.cpu().numpy(), g[f"hs_{name}"]) == 0, name
        de = vofx.Evolution(grid)(recs, flags, t, dt, upd)
        assert n_mismatch(upd.cpu().numpy(), g[f"update_{name}"]) == 0, name
        assert de == float(g[f"dtest_{name}"]), name
        if name == "hf":
            kappa = grid.empty("curvature")
            op = vofx.CartesianHeightFunctionCurvature(grid)
            op(u, recs, flags, kappa)
            assert np.array_equal(op.satisfies_constraint.cpu().numpy(), g["kappa_valid"])
            # curvature uses pow(): CUDA's and glibc's pow may differ in the last bit -> 1e-12 relative
            k_ref = g["kappa"]
            k_dev = kappa.cpu().numpy()
            assert np.allclose(k_dev, k_ref, rtol=1e-12, atol=0.0)
        # the fused device-resident loop against the reference's loop
        kind = {"youngs": vofx.RECON_YOUNGS, "hf": vofx.RECON_HF, "swartz": vofx.RECON_SWARTZ}[name]
        alg = vofx.Algorithm(grid, cfl, eps, reconstruction_kind=kind)
        uh = torch.tensor(g["u0"], device=dev)
        st = alg.run_passes(uh, int(g[f"run_{name}_passes"]))
        assert n_mismatch(uh.cpu().numpy(), g[f"run_{name}_u"]) == 0, name
        assert st.time == float(g[f"run_{name}_time"]) and st.dt == float(g[f"run_{name}_dt"]), name
    grid.close()


def test_c1_run(vofx):
    """config C1 (plumbing): 128^2 rotating circle, Young's, 20 loop passes; known answer of SURVEY.md 8(c)"""
    import torch
    g = golden("run_c1_128.npz")
    grid = vofx.CartesianGrid((128, 128))
    grid.set_velocity_builtin(vofx.PROB_ROTATING_CIRCLE)
    uh = torch.tensor(g["u0"], device=grid.device)
    alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS)
    st = alg.run_passes(uh, 20)
    assert n_mismatch(uh.cpu().numpy(), g["u"]) == 0
    assert st.time == 0.12448621033571523 and st.dt == 0.0064175380278990047
    # Algorithm::operator()( uh, start, end ): do { pass } while( time < end ) stops at the reference's pass (the 20th: the time
    # after 19 passes is 0.1181 < 0.12)
    uh2 = torch.tensor(g["u0"], device=grid.device)
    st2 = alg(uh2, 0.0, 0.12)
    assert n_mismatch(uh2.cpu().numpy(), g["u"]) == 0
    assert st2.time == 0.12448621033571523 and st2.dt == 0.0064175380278990047
    grid.close()


@pytest.mark.parametrize("n,problem", [((64, 64, 64), 4), ((48, 40, 56), 1), ((96, 96, 96), 0), ((256, 256), 2), ((200, 160), 4)])
def test_against_oracle_larger(vofx, n, problem):
    """seeded larger cases against the oracle (Young's + HF, several passes), incl. device initial data"""
    import torch
    from oracle import oraclelib as O
    og = O.OracleGrid(n)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    u0 = og.init(problem)
    if not (problem == 0 and len(n) == 2):
        assert n_mismatch(grid.init(problem).cpu().numpy(), u0) == 0
    for kind in (vofx.RECON_YOUNGS, vofx.RECON_HF):
        uo, to, do, _, _ = og.run(kind, problem, 1e-6, 0.5, 6, u0)
        uh = torch.tensor(u0, device=grid.device)
        st = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=kind).run_passes(uh, 6)
        assert n_mismatch(uh.cpu().numpy(), uo) == 0
        assert st.time == to and st.dt == do
    # face velocities
    rng = np.random.default_rng(5)
    for _ in range(50):
        c = int(rng.integers(0, og.cells)); f = int(rng.integers(0, 2 * len(n)))
        assert same_bits(grid.face_velocity(0.37, c, f), og.face_velocity(problem, 0.37, c, f))
    grid.close()


def test_host_buffer_variants(vofx):
    """the HOST-pointer entry points (what the DUNE-side std::vector containers bind to)"""
    L, vlib = _lib()
    g = golden("operators_3d_leveque_16.npz")
    grid = vofx.CartesianGrid((16, 16, 16))
    grid.set_velocity_builtin(4)
    u = np.ascontiguousarray(g["u"]); N = u.size
    flags = np.empty(N, dtype=np.uint8)
    vlib.check(L.vofx_flag_host(grid._ctx, _p(u), 1e-6, _p(flags)))
    assert np.array_equal(flags, g["flags"])
    hs = np.empty((N, 4))
    vlib.check(L.vofx_reconstruct_host(grid._ctx, 0, _p(u), _p(flags), _p(hs), 10))
    assert n_mismatch(hs, g["hs_youngs"]) == 0
    upd = np.empty(N); de = C.c_double()
    vlib.check(L.vofx_evolve_host(grid._ctx, _p(hs), _p(flags), float(g["time"]), float(g["dt"]), _p(upd), C.byref(de)))
    assert n_mismatch(upd, g["update_youngs"]) == 0 and de.value == float(g["dtest_youngs"])
    # one loop pass on host buffers == reference loop
    uh = np.ascontiguousarray(g["u0"]).copy()
    p = vlib.StepParams(0, 10, 1e-6, 0.5, 0)
    vlib.check(L.vofx_step_begin(grid._ctx, 0.0, 0.0))
    for _ in range(6):
        vlib.check(L.vofx_step_host(grid._ctx, C.byref(p), _p(uh), None))
    assert n_mismatch(uh, g["run_youngs_u"]) == 0
    grid.close()


def test_mass_and_bounds_512_slab(vofx):
    """size-independent properties at a BASELINE-sized slab (512x512x64 of the 512^3 grid is too thin for the
    sphere; use the full 256^3 here): rigid rotation conserves mass to round-off (discretely divergence-free face
    velocities), flags stay in {0,1,2,3}, colour stays within the reference's over/undershoot."""
    import torch
    n = (256, 256, 256)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(vofx.PROB_ROTATING_CIRCLE)
    uh = grid.init(vofx.PROB_ROTATING_CIRCLE)
    m0 = float(uh.sum())
    alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS)
    st = alg.run_passes(uh, 12)
    m1 = float(uh.sum())
    assert abs(m1 - m0) <= 1e-12 * m0
    assert st.mixed_cells > 10000
    assert int(alg.flags.max()) <= 3
    assert float(uh.min()) > -1e-2 and float(uh.max()) < 1 + 1e-2
    grid.close()


@pytest.mark.parametrize("case", ["2d_zalesak_64", "3d_sphere_16"])
def test_curvature_inside_time_loop(vofx, case):
    """The time loop with curvature (fused HF kernel in 2D, sparse clear of the previous step's curvature entries)
    against the operator-level calls on the same state: bit-identical, and zero wherever no mixed cell is."""
    import torch
    g = golden(f"operators_{case}.npz")
    n = tuple(int(x) for x in g["n"])
    problem, eps, cfl = int(g["problem"]), float(g["eps"]), float(g["cfl"])
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    u0 = torch.tensor(g["u0"], device=grid.device)
    passes = 5
    alg = vofx.Algorithm(grid, cfl, eps, reconstruction_kind=vofx.RECON_HF, with_curvature=True)
    ua = u0.clone()
    alg.run_passes(ua, passes)
    # state before the last pass, then the operators one by one
    ref = vofx.Algorithm(grid, cfl, eps, reconstruction_kind=vofx.RECON_HF)
    ub = u0.clone()
    ref.run_passes(ub, passes - 1)
    flags, recs, kappa = grid.empty("flags"), grid.empty("reconstructions"), grid.empty("curvature")
    vofx.FlagOperator(grid, eps)(ub, flags)
    vofx.HeightFunctionReconstruction(grid)(ub, recs, flags)
    vofx.CartesianHeightFunctionCurvature(grid)(ub, recs, flags, kappa)
    assert same_bits(alg.curvature.cpu().numpy(), kappa.cpu().numpy())
    mixed = (flags == 1) | (flags == 2)
    assert float(alg.curvature[~mixed].abs().max()) == 0.0
    assert same_bits(alg.reconstructions.cpu().numpy()[mixed.cpu().numpy()], recs.cpu().numpy()[mixed.cpu().numpy()])
    grid.close()


@pytest.mark.parametrize("case", ["2d_zalesak_64", "2d_sflow_40x24", "3d_leveque_32", "3d_sflow_20x12x16"])
def test_interface_extraction(vofx, case):
    """getInterfaceVertices / MixedCellMapper on the device against the oracle's per-cell
    interface(): ordered mixed cells, CSR offsets and vertices, bit for bit"""
    import torch
    from oracle import oraclelib as O
    g = golden(f"operators_{case}.npz")
    n = tuple(int(x) for x in g["n"])
    dim = len(n)
    grid = vofx.CartesianGrid(n)
    u = torch.tensor(g["u"], device=grid.device)
    flags, recs = grid.empty("flags"), grid.empty("reconstructions")
    vofx.FlagOperator(grid, float(g["eps"]))(u, flags)
    vofx.ModifiedYoungsReconstruction(grid)(u, recs, flags)
    cells, offsets, vertices = vofx.get_interface_vertices(grid, recs, flags)
    f = flags.cpu().numpy()
    expect = np.flatnonzero((f == 1) | (f == 2))
    assert np.array_equal(cells, expect)
    hs = recs.cpu().numpy()
    h = np.array([1.0 / x for x in n])
    for i, c in enumerate(cells):
        ijk = [c % n[0], (c // n[0]) % n[1]] + ([c // (n[0] * n[1])] if dim == 3 else [])
        lo = np.array(ijk, dtype=np.float64)
        lo = np.array([0.0 + ijk[d] * h[d] for d in range(dim)])
        nv, verts, _, _ = O.interface(lo, h, hs[c])
        assert offsets[i + 1] - offsets[i] == nv, (case, c)
        assert same_bits(vertices[offsets[i]:offsets[i + 1]], verts[:nv]), (case, c)
    assert offsets[-1] == len(vertices)
    grid.close()


def _torch_flags(u, n, eps):
    """flagging.hh:35-70 restated with torch element-wise ops (independent of the kernels): dense check at full size"""
    import torch
    U = u.view(n[2], n[1], n[0]) if len(n) == 3 else u.view(n[1], n[0])
    f = torch.full_like(U, 3, dtype=torch.uint8)
    f[U <= 1 - eps] = 1
    f[U < eps] = 0
    f[U != U] = 99
    empty = U < eps
    nb = torch.zeros_like(empty)
    for d in range(U.dim()):
        lo = [slice(None)] * U.dim(); hi = [slice(None)] * U.dim()
        lo[d] = slice(0, -1); hi[d] = slice(1, None)
        nb[tuple(hi)] |= empty[tuple(lo)]
        nb[tuple(lo)] |= empty[tuple(hi)]
    f[(f == 3) & nb] = 2
    return f.reshape(-1)


def test_full_size_512_properties(vofx):
    """configs[2] at its full size (512^3, the bench workload): the divergence-free LeVeque face velocities conserve
    mass to round-off, the flags equal an independent dense restatement bit for bit, every mixed cell carries a unit
    normal whose plane cuts off its colour (|vf(plane) - u| small), two runs are bit-identical (no atomics in any sum)."""
    import torch
    n = (512, 512, 512)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = grid.init(vofx.PROB_LEVEQUE)
    m0 = float(u0.sum())
    runs = []
    for _ in range(2):
        uh = u0.clone()
        alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS)
        st = alg.run_passes(uh, 10)
        runs.append((uh, st, alg))
    (ua, sa, alg), (ub, sb, _) = runs
    assert torch.equal(ua.view(torch.int64), ub.view(torch.int64)) and sa.time == sb.time and sa.dt == sb.dt
    assert abs(float(ua.sum()) - m0) <= 1e-11 * m0
    assert sa.mixed_cells > 100000
    # flags of the last pass belong to the colour function BEFORE its update: recompute that state
    uc = u0.clone()
    ref = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS)
    ref.run_passes(uc, 9)
    assert torch.equal(alg.flags, _torch_flags(uc, n, 1e-6))
    mixed = (alg.flags == 1) | (alg.flags == 2)
    hs = alg.reconstructions[mixed]
    assert int(mixed.sum()) == sa.mixed_cells
    nrm = (hs[:, :3] ** 2).sum(dim=1)
    ok = nrm > 0.5                                           # the zero half-space marks a degenerate Young's normal
    assert float((nrm[ok] - 1.0).abs().max()) < 1e-14 and int((~ok).sum()) < 10
    assert float(ua.min()) > -1e-2 and float(ua.max()) < 1 + 1e-2
    grid.close()


@pytest.mark.parametrize("case", ["2d_zalesak_64", "2d_sflow_40x24", "2d_circle_32"])
def test_swartz_curvature_2d(vofx, case):
    """SwartzCurvature (2D) against the oracle (which is pinned to the reference): bit-exact"""
    import torch
    from oracle import oraclelib as O
    g = golden(f"operators_{case}.npz")
    n = tuple(int(x) for x in g["n"])
    grid = vofx.CartesianGrid(n)
    u = torch.tensor(g["u"], device=grid.device)
    flags, recs, kappa = grid.empty("flags"), grid.empty("reconstructions"), grid.empty("curvature")
    vofx.FlagOperator(grid, float(g["eps"]))(u, flags)
    og = O.OracleGrid(n)
    for cls in (vofx.ModifiedYoungsReconstruction, vofx.HeightFunctionReconstruction):
        cls(grid)(u, recs, flags)
        vofx.SwartzCurvature(grid)(u, recs, flags, kappa)
        ko = og.curvature_swartz(recs.cpu().numpy(), flags.cpu().numpy())
        assert np.count_nonzero(ko) > 10
        assert n_mismatch(kappa.cpu().numpy(), ko) == 0
        # and against the vector dumped from the unmodified reference (tests/golden/make_golden.py)
        kname = "youngs" if cls is vofx.ModifiedYoungsReconstruction else "hf"
        assert n_mismatch(kappa.cpu().numpy(), golden("curvature_swartz_2d.npz")[f"{case}_{kname}"]) == 0
    og.close()
    grid.close()


@pytest.mark.parametrize("n,problem,passes", [((24, 20, 16), 3, 8), ((40, 32), 3, 8), ((20, 20, 20), 0, 6)])
def test_edge_cases_axis_aligned_and_walls(vofx, n, problem, passes):
    """SURVEY.md "edge cases": a planar interface aligned with an axis (LinearWall: normals exactly +-e_x -> the tied-height
    walk of the plane-constant solve, planes exactly on cell faces -> the serial clip for nodes on the plane), cells at
    the domain wall (ghost terms of Young's, wall faces skipped by the flux), all three reconstructions, against the
    oracle bit for bit"""
    import torch
    from oracle import oraclelib as O
    og = O.OracleGrid(n)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    u0 = og.init(problem)
    kinds = [vofx.RECON_YOUNGS, vofx.RECON_HF] + ([vofx.RECON_SWARTZ] if len(n) == 2 or problem == 3 else [])
    for kind in kinds:
        p = passes if kind != vofx.RECON_SWARTZ else 3
        uo, to, do, _, _ = og.run(kind, problem, 1e-6, 0.5, p, u0)
        uh = torch.tensor(u0, device=grid.device)
        st = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=kind).run_passes(uh, p)
        assert n_mismatch(uh.cpu().numpy(), uo) == 0, kind
        assert st.time == to and st.dt == do, kind
    grid.close()
    og.close()


def test_edge_cases_special_colours(vofx):
    """NaN colour -> flag 99 (never mixed); colours slightly outside [0,1] (over/undershoots) and just above eps (the
    corner-plane artefact of the absolute 1e-14 volume tolerance) reconstruct exactly like the oracle"""
    import torch
    from oracle import oraclelib as O
    n = (16, 16, 16)
    og = O.OracleGrid(n)
    grid = vofx.CartesianGrid(n, h=(1.0 / 512,) * 3)          # small cells: u * h^3 < 1e-14 for u < 1.3e-6
    og2 = O.OracleGrid(n, h=(1.0 / 512,) * 3)
    u = og.init(O.PROB_LEVEQUE)
    rng = np.random.default_rng(9)
    mixed = np.flatnonzero((u > 1e-6) & (u < 1 - 1e-6))
    u[mixed[::7]] = 10.0 ** rng.uniform(-6, -5, size=len(mixed[::7]))     # just above eps
    u[mixed[1::7]] = 1.0 + 3e-3
    u[mixed[2::7]] = -2e-3
    u[5] = np.nan
    fo = og2.flag(u, 1e-6)
    ut = torch.tensor(u, device=grid.device)
    flags, recs = grid.empty("flags"), grid.empty("reconstructions")
    vofx.FlagOperator(grid, 1e-6)(ut, flags)
    assert np.array_equal(flags.cpu().numpy(), fo) and fo[5] == 99
    for cls, kind in ((vofx.ModifiedYoungsReconstruction, O.RECON_YOUNGS), (vofx.HeightFunctionReconstruction, O.RECON_HF)):
        cls(grid)(ut, recs, flags)
        assert n_mismatch(recs.cpu().numpy(), og2.reconstruct(kind, u, fo)) == 0
    grid.close()


def test_256_cubed_against_oracle(vofx):
    """the largest case the oracle finishes in seconds (configs[3]'s grid, 2.8e4 mixed cells): 12 loop passes of the
    3D Young's step, every colour value, time and dt bit-identical"""
    import torch
    from oracle import oraclelib as O
    n = (256, 256, 256)
    og = O.OracleGrid(n)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = grid.init(vofx.PROB_LEVEQUE)
    uo, to, do, _, _ = og.run(O.RECON_YOUNGS, O.PROB_LEVEQUE, 1e-6, 0.5, 12, u0.cpu().numpy())
    uh = u0.clone()
    st = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS).run_passes(uh, 12)
    assert n_mismatch(uh.cpu().numpy(), uo) == 0
    assert st.time == to and st.dt == do
    grid.close()
    og.close()


@pytest.mark.parametrize("n,problem", [((128, 128), 4), ((48, 40, 56), 4), ((37 * 4, 50), 2)])
def test_device_diagnostics(vofx, n, problem):
    """vofx_diagnostics against sequential sums over the same arrays (what ColorFunction / cellwiseL1error,
    test/errors.hh:138-151, accumulate): min and max exact, the sums within 1e-12 relative (different summation tree)."""
    import math

    import torch
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(problem)
    u0 = grid.init(problem)
    alg = vofx.Algorithm(grid, 0.5, 1e-6, vofx.RECON_YOUNGS)
    uh = u0.clone()
    alg.run_passes(uh, 6)
    d = vofx.diagnostics(grid, uh, u0)
    a, e = uh.cpu().numpy(), u0.cpu().numpy()
    volume = float(np.prod([1.0 / k for k in n]))
    ref_sum = math.fsum(a.tolist())
    ref_l1 = math.fsum(np.abs(a - e).tolist()) * volume
    assert d["min"] == a.min() and d["max"] == a.max()
    assert abs(d["sum"] - ref_sum) <= 1e-12 * abs(ref_sum)
    assert abs(d["mass"] - ref_sum * volume) <= 1e-12 * abs(ref_sum * volume)
    assert abs(d["l1"] - ref_l1) <= 1e-12 * max(abs(ref_l1), volume)
    assert d["l1"] > 0.0
    # without an exact solution the error entry is zero
    d0 = vofx.diagnostics(grid, u0)
    assert d0["l1"] == 0.0 and d0["min"] >= 0.0 and d0["max"] <= 1.0
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [(128, 128), (1024, 768)])
def test_circle_interpolation_and_l1error(vofx, n):
    """circleInterpolation (interpolation/circleinterpolation.hh:75-153: the initialiser of RotatingCircle< double, 2 >, configs[0])
    and l1error against the exact circle (test/errors.hh:62-95) on the device, against the oracle.  Cells the circle does not cut
    are exact; cut cells go through asin / sin (CUDA's vs glibc's): <= 4e-16 absolute on a fraction, 1e-12 relative on the error."""
    import torch
    from oracle import oraclelib as O
    L, vlib = _lib()
    grid = vofx.CartesianGrid(n)
    og = O.OracleGrid(n)
    center = np.array([0.5, 0.75])
    uo = og.circle_interpolation(center, 0.15)
    ud = grid.empty("color")
    cc = (C.c_double * 2)(*center)
    vlib.check(L.vofx_init_circle(grid._ctx, cc, 0.15, grid.pc(ud)))
    u = ud.cpu().numpy()
    uncut = (uo == 0.0) | (uo == 1.0)
    assert np.array_equal(u[uncut], uo[uncut])
    assert np.abs(u - uo).max() <= 4e-16
    assert np.count_nonzero(~uncut) > 2 * min(n) * 0.15
    # the known answer of the reference's own initialiser: the cell values sum to the disc's area
    h2 = 1.0 / (n[0] * n[1])
    assert abs(u.sum() * h2 - np.pi * 0.15 ** 2) < 1e-13 * np.pi * 0.15 ** 2
    # error of a reconstruction against the exact circle
    grid.set_velocity_builtin(vofx.PROB_ROTATING_CIRCLE)
    uh = torch.tensor(uo, device=grid.device)
    alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_HF)
    st = alg.run_passes(uh, 12)
    # flags / reconstructions of the last pass belong to the state before its update
    u11, t11, d11, _, _ = og.run(O.RECON_HF, O.PROB_ROTATING_CIRCLE, 1e-6, 0.5, 11, uo)
    fo = og.flag(u11, 1e-6)
    ho = og.reconstruct(O.RECON_HF, u11, fo)
    omega = 2 * np.pi / 10
    c_t = np.array([0.5, 0.5]) + np.cos(omega * t11) * np.array([0.0, 0.25]) + np.sin(omega * t11) * np.array([0.25, 0.0])
    eo = og.l1error_circle(ho, fo, c_t, 0.15)
    ed = C.c_double(0.0)
    cc = (C.c_double * 2)(*c_t)
    vlib.check(L.vofx_l1error_circle_device(grid._ctx, grid.ph(alg.reconstructions), grid.pf(alg.flags), cc, 0.15, C.byref(ed)))
    assert eo > 0.0 and abs(ed.value - eo) <= 1e-12 * eo
    grid.close()
    og.close()


def test_checkpoint_binarywriter_format(vofx, tmp_path):
    """vofx_checkpoint_write / _read: the bytes of the reference's BinaryWriter (example/binarywriter.hh:36-60,
    colorfunction.hh:40-51: the time, then one raw double per cell in iteration order), the writer's file name, and a restart
    that continues the loop bit for bit"""
    import torch
    L, vlib = _lib()
    name = C.create_string_buffer(64)
    vlib.check(L.vofx_checkpoint_name(name, 64, b"vof", 1, 7, 3, 8))
    assert name.value == b"s0008-p0003-vof-1-00007.bin"
    n = (48, 40, 32)
    ga, gb = vofx.CartesianGrid(n), vofx.CartesianGrid(n)
    for g in (ga, gb):
        g.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = ga.init(vofx.PROB_LEVEQUE)
    p = vlib.StepParams(0, 10, 1e-6, 0.5, 0)
    host = u0.cpu().numpy().copy()
    vlib.check(L.vofx_step_begin(ga._ctx, 0.0, 0.0))
    vlib.check(L.vofx_color_upload(ga._ctx, C.c_void_p(host.ctypes.data)))
    for _ in range(6):
        vlib.check(L.vofx_step_resident(ga._ctx, C.byref(p)))
    t, dt = C.c_double(), C.c_double()
    vlib.check(L.vofx_step_state(ga._ctx, C.byref(t), C.byref(dt), None, None))
    path = str(tmp_path / name.value.decode())
    vlib.check(L.vofx_checkpoint_write(ga._ctx, path.encode(), t.value))
    raw = np.fromfile(path, dtype="<f8")
    mid = np.empty_like(host)
    vlib.check(L.vofx_color_download(ga._ctx, C.c_void_p(mid.ctypes.data), None))
    assert raw.size == 1 + host.size and raw[0] == t.value and np.array_equal(raw[1:].view(np.uint64), mid.view(np.uint64))
    # continue on the first context, restart on the second from the file: identical afterwards
    for _ in range(5):
        vlib.check(L.vofx_step_resident(ga._ctx, C.byref(p)))
    t2 = C.c_double()
    vlib.check(L.vofx_checkpoint_read(gb._ctx, path.encode(), C.byref(t2)))
    assert t2.value == t.value
    vlib.check(L.vofx_step_begin(gb._ctx, t2.value, dt.value))
    for _ in range(5):
        vlib.check(L.vofx_step_resident(gb._ctx, C.byref(p)))
    a, b = np.empty_like(host), np.empty_like(host)
    vlib.check(L.vofx_color_download(ga._ctx, C.c_void_p(a.ctypes.data), None))
    vlib.check(L.vofx_color_download(gb._ctx, C.c_void_p(b.ctypes.data), None))
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    ta, tb = C.c_double(), C.c_double()
    vlib.check(L.vofx_step_state(ga._ctx, C.byref(ta), None, None, None))
    vlib.check(L.vofx_step_state(gb._ctx, C.byref(tb), None, None, None))
    assert ta.value == tb.value
    # a file of the wrong size is refused
    np.zeros(10).tofile(str(tmp_path / "short.bin"))
    assert L.vofx_checkpoint_read(gb._ctx, str(tmp_path / "short.bin").encode(), C.byref(t2)) != 0
    ga.close()
    gb.close()


def test_step_host_sparse_matches_full_upload(vofx):
    """vofx_step_host_sparse: a host colour function whose changes the caller declares.  Plain loop (nothing touched): the host
    array after every pass is bit-identical to vofx_step_host's, with no host -> device colour traffic after the first call;
    then the caller overwrites a few cells on the host and lists them; then it uses another entry point in between (the mirror
    is dropped and the whole array goes up again)."""
    import numpy as np
    from dune_vof_b200 import lib as vlib
    L = vlib.load()
    n = (48, 40, 56)
    N = n[0] * n[1] * n[2]
    p = vlib.StepParams(vofx.RECON_YOUNGS, 10, 1e-6, 0.5, 0)

    def make():
        g = vofx.CartesianGrid(n)
        g.set_velocity_builtin(vofx.PROB_LEVEQUE)
        u = g.init(vofx.PROB_LEVEQUE).cpu().numpy().copy()
        vlib.check(L.vofx_step_begin(g._ctx, 0.0, 0.0))
        return g, np.ascontiguousarray(u)

    ga, ua = make()
    gb, ub = make()
    h2d, d2h = C.c_int64(0), C.c_int64(0)
    rng = np.random.default_rng(9)
    for k in range(14):
        touched = None
        if k in (5, 9):
            # the caller edits its host array: a few cells near the interface and a few far away
            cells = rng.integers(0, N, 40).astype(np.int32)
            vals = rng.random(40)
            ua[cells] = vals
            ub[cells] = vals
            touched = np.ascontiguousarray(cells)
        if k == 11:
            # another entry point uses the context's mirror in between
            f = np.empty(N, dtype=np.uint8)
            vlib.check(L.vofx_flag_host(gb._ctx, _p(ub), 1e-6, _p(f)))
        vlib.check(L.vofx_step_host(ga._ctx, C.byref(p), _p(ua), None))
        vlib.check(L.vofx_step_host_sparse(gb._ctx, C.byref(p), _p(ub), None, _p(touched) if touched is not None else None,
                                           0 if touched is None else len(touched)))
        vlib.check(L.vofx_host_transfer_bytes(gb._ctx, C.byref(h2d), C.byref(d2h)))
        assert n_mismatch(ua, ub) == 0, k
        if k == 0 or k == 11:
            assert h2d.value == 8 * N                      # first call / dropped mirror: the whole array
        elif touched is not None:
            assert h2d.value == 12 * len(touched)
        else:
            assert h2d.value == 0
        assert 0 < d2h.value < 8 * N
    ta, tb = C.c_double(), C.c_double()
    da, db = C.c_double(), C.c_double()
    e, m = C.c_double(), C.c_int64()
    vlib.check(L.vofx_step_state(ga._ctx, C.byref(ta), C.byref(da), C.byref(e), C.byref(m)))
    vlib.check(L.vofx_step_state(gb._ctx, C.byref(tb), C.byref(db), C.byref(e), C.byref(m)))
    assert ta.value == tb.value and da.value == db.value
    ga.close()
    gb.close()
