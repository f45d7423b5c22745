"""Unstructured meshes on the GPU (vofx_mesh_*, include/vofx.h; 2D triangles / quadrilaterals, 3D tetrahedra / hexahedra)
against the golden vectors of the unmodified reference and against the oracle on larger seeded meshes.  Bit-exact: flags, half-spaces, updates, dtEst, the device-resident loop."""
import numpy as np
import pytest

from dune_vof_b200 import mesh
from oracle import meshlib
from util import golden, n_mismatch

pytestmark = pytest.mark.gpu

KINDS = ["tri", "quad", "mixed"]


@pytest.mark.parametrize("kind", KINDS)
def test_operators_match_golden(kind):
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    D = mesh.MeshOperators(m)
    flags = D.flag(z["u"], float(z["eps"]))
    assert np.array_equal(flags, z["flags"])
    hs = D.youngs(z["u"], z["flags"])
    assert n_mismatch(hs, z["halfspaces"]) == 0
    upd, est = D.evolve(z["halfspaces"], z["flags"], z["velocity"], float(z["dt"]))
    assert n_mismatch(upd, z["update"]) == 0
    assert est == float(z["dt_est"])
    assert D.launches > 0


@pytest.mark.parametrize("kind", KINDS)
def test_resident_loop_matches_golden(kind):
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    D = mesh.MeshOperators(m)
    u, t, dt = D.run(z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, z["u_end"]) == 0
    assert t == float(z["time_end"]) and dt == float(z["dt_end"])
    # the operator-by-operator loop gives the same bits as the resident one
    v, t2, dt2 = meshlib.run(D, m, z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, v) == 0 and (t, dt) == (t2, dt2)


@pytest.mark.parametrize("kind", KINDS)
def test_swartz_matches_golden(kind):
    """ModifiedSwartzReconstruction in 2D on the mesh: operator and device-resident loop"""
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    D = mesh.MeshOperators(m)
    hs = D.swartz(z["u"], z["flags"], 10)
    assert n_mismatch(hs, z["halfspaces_swartz"]) == 0
    u, t, dt = D.run(z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]), swartz=True)
    assert n_mismatch(u, z["u_end_swartz"]) == 0
    assert t == float(z["time_end_swartz"]) and dt == float(z["dt_end_swartz"])
    # switching back to Young's on the same object rebuilds the captured pass
    u, t, dt = D.run(z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, z["u_end"]) == 0


def test_swartz_larger_mesh_matches_oracle():
    v, cells = mesh.perturbed_square(96, "mixed", 31)
    m = mesh.flatten(v, cells)
    O, D = meshlib.OracleMesh(m), mesh.MeshOperators(m)
    u = meshlib.circle_fractions(m, center=(0.5, 0.7), radius=0.2, samples=8)
    vel = meshlib.rotation_velocity(m)
    a = meshlib.run(O, m, u, vel, 1e-6, 0.5, 15, swartz=True)
    b = D.run(u, vel, 1e-6, 0.5, 15, swartz=True)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]


@pytest.mark.parametrize("kind,n,seed", [("tri", 160, 21), ("mixed", 192, 22), ("quad", 256, 23)])
def test_larger_meshes_match_oracle(kind, n, seed):
    v, cells = mesh.perturbed_square(n, kind, seed)
    m = mesh.flatten(v, cells)
    O, D = meshlib.OracleMesh(m), mesh.MeshOperators(m)
    u = meshlib.circle_fractions(m, center=(0.5, 0.7), radius=0.2, samples=8)
    vel = meshlib.rotation_velocity(m)
    a = meshlib.run(O, m, u, vel, 1e-6, 0.5, 40)
    b = D.run(u, vel, 1e-6, 0.5, 40)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]
    assert abs(((b[0] - u) * m.volumes).sum()) < 1e-14


def test_edge_cases():
    v, cells = mesh.perturbed_square(12, "mixed", 5)
    m = mesh.flatten(v, cells)
    O, D = meshlib.OracleMesh(m), mesh.MeshOperators(m)
    vel = meshlib.rotation_velocity(m)
    for u in (np.zeros(m.n_cells), np.ones(m.n_cells), np.full(m.n_cells, 0.5), np.full(m.n_cells, np.nan)):
        fo, fd = O.flag(u, 1e-6), D.flag(u, 1e-6)
        assert np.array_equal(fo, fd)
        ho, hd = O.youngs(u, fo), D.youngs(u, fd)
        assert n_mismatch(ho, hd) == 0
        uo, eo = O.evolve(ho, fo, vel, 0.02)
        ud, ed = D.evolve(hd, fd, vel, 0.02)
        assert n_mismatch(uo, ud) == 0 and eo == ed
    # zero velocity: dtEst = volume / 0 = inf never wins against DBL_MAX
    u = meshlib.circle_fractions(m)
    f = O.flag(u, 1e-6)
    h = O.youngs(u, f)
    uo, eo = O.evolve(h, f, np.zeros_like(vel), 0.02)
    ud, ed = D.evolve(h, f, np.zeros_like(vel), 0.02)
    assert n_mismatch(uo, ud) == 0 and eo == ed == np.finfo(np.float64).max


def test_descriptor_validation():
    import ctypes as C

    from dune_vof_b200 import lib
    v, cells = mesh.perturbed_square(4, "tri", 1)
    m = mesh.flatten(v, cells)
    bad = mesh.FlatMesh(**{k: getattr(m, k) for k in m.__dataclass_fields__})
    bad.stencil = m.stencil[::-1].copy()
    with pytest.raises(lib.VofxError):
        mesh.MeshOperators(bad)
    L = lib.load()
    h = C.c_void_p()
    assert L.vofx_mesh_create(None, -1, C.byref(h)) != 0


# ---- 3D: tetrahedra and hexahedra --------------------------------------------------------------------------------------------

KINDS3 = ["tet", "hex"]


@pytest.mark.parametrize("kind", KINDS3)
def test_operators_match_golden_3d(kind):
    z = golden(f"mesh3d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    D = mesh.MeshOperators(m)
    flags = D.flag(z["u"], float(z["eps"]))
    assert np.array_equal(flags, z["flags"])
    hs = D.youngs(z["u"], z["flags"])
    assert hs.shape == (m.n_cells, 4)
    assert n_mismatch(hs, z["halfspaces"]) == 0
    upd, est = D.evolve(z["halfspaces"], z["flags"], z["velocity"], float(z["dt"]))
    assert n_mismatch(upd, z["update"]) == 0
    assert est == float(z["dt_est"])


@pytest.mark.parametrize("kind", KINDS3)
def test_resident_loop_matches_golden_3d(kind):
    z = golden(f"mesh3d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    D = mesh.MeshOperators(m)
    u, t, dt = D.run(z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, z["u_end"]) == 0
    assert t == float(z["time_end"]) and dt == float(z["dt_end"])
    assert abs(((u - z["u_run"]) * m.volumes).sum()) < 1e-15


@pytest.mark.parametrize("kind,n,seed", [("tet", 14, 41), ("hex", 28, 42)])
def test_larger_meshes_match_oracle_3d(kind, n, seed):
    v, cells = mesh.perturbed_cube(n, kind, seed)
    m = mesh.flatten3(v, cells)
    O, D = meshlib.OracleMesh(m), mesh.MeshOperators(m)
    u = meshlib.sphere_fractions(m, center=(0.5, 0.55, 0.45), radius=0.28, samples=60, seed=seed)
    vel = meshlib.rotation_velocity3(m)
    a = meshlib.run(O, m, u, vel, 1e-6, 0.5, 12)
    b = D.run(u, vel, 1e-6, 0.5, 12)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]
    assert abs(((b[0] - u) * m.volumes).sum()) < 1e-14


def test_3d_mesh_descriptor_validation_and_swartz_scope():
    from dune_vof_b200 import lib
    v, cells = mesh.perturbed_cube(3, "tet", 1)
    m = mesh.flatten3(v, cells)
    bad = mesh.FlatMesh(**{k: getattr(m, k) for k in m.__dataclass_fields__})
    bad.face_corner_offsets = None                       # a 3D mesh needs the corner counts of its intersections
    with pytest.raises(lib.VofxError):
        mesh.MeshOperators(bad)
    D = mesh.MeshOperators(m)
    u = np.full(m.n_cells, 0.5)
    with pytest.raises(lib.VofxError):                   # the Swartz reconstruction on meshes is 2D only
        D.swartz(u, D.flag(u, 1e-6))


@pytest.mark.parametrize("kind", KINDS3)
def test_edge_cases_3d(kind):
    """uniform, full, empty and NaN colour functions, zero velocity, a colour function slightly outside [0, 1]"""
    v, cells = mesh.perturbed_cube(5, kind, 7)
    m = mesh.flatten3(v, cells)
    O, D = meshlib.OracleMesh(m), mesh.MeshOperators(m)
    vel = meshlib.rotation_velocity3(m)
    rng = np.random.default_rng(3)
    over = np.clip(meshlib.sphere_fractions(m, samples=40) + (rng.random(m.n_cells) - 0.5) * 4e-3, -2e-3, 1.002)
    for u in (np.zeros(m.n_cells), np.ones(m.n_cells), np.full(m.n_cells, 0.5), np.full(m.n_cells, np.nan), over):
        fo, fd = O.flag(u, 1e-6), D.flag(u, 1e-6)
        assert np.array_equal(fo, fd)
        ho, hd = O.youngs(u, fo), D.youngs(u, fd)
        assert n_mismatch(ho, hd) == 0
        uo, eo = O.evolve(ho, fo, vel, 0.02)
        ud, ed = D.evolve(hd, fd, vel, 0.02)
        assert n_mismatch(uo, ud) == 0 and eo == ed
    u = meshlib.sphere_fractions(m, samples=40)
    f = O.flag(u, 1e-6)
    h = O.youngs(u, f)
    uo, eo = O.evolve(h, f, np.zeros_like(vel), 0.02)
    ud, ed = D.evolve(h, f, np.zeros_like(vel), 0.02)
    assert n_mismatch(uo, ud) == 0 and eo == ed == np.finfo(np.float64).max
