"""Unstructured 2D meshes, CPU side: the plain-C restatement (vo_mesh_*, oracle/vof_oracle.c) against the golden vectors
dumped from the unmodified reference (tests/golden/mesh2d_*.npz, tests/golden/make_golden_mesh.py) and, where the
reference harness is present (oracle/_ref/libvofrefmesh.so, built only where /root/reference is mounted), against the
reference itself on fresh seeded meshes.  Bit-exact everywhere."""
import numpy as np
import pytest

from dune_vof_b200 import mesh
from oracle import meshlib
from util import golden, n_mismatch

KINDS = ["tri", "quad", "mixed"]


@pytest.mark.parametrize("kind", KINDS)
def test_flatten_reproduces_the_stored_mesh_topology(kind):
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    # the reference's own VertexNeighborsStencil (stencil/vertexneighborsstencil.hh:52-94) on this mesh
    assert np.array_equal(m.stencil_offsets, z["ref_stencil_offsets"])
    assert np.array_equal(m.stencil, z["ref_stencil"])
    sizes = np.diff(m.corner_offsets)
    assert set(sizes.tolist()) <= {3, 4}
    if kind == "mixed":
        assert set(sizes.tolist()) == {3, 4}
    # conforming: every interior intersection is seen from both sides
    owner = np.repeat(np.arange(m.n_cells), np.diff(m.face_offsets))
    inner = m.face_neighbor >= 0
    pairs = set(zip(owner[inner].tolist(), m.face_neighbor[inner].tolist()))
    assert all((b, a) in pairs for a, b in pairs)
    assert abs(m.volumes.sum() - 1.0) < 1e-12


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_golden_operators(kind):
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    O = meshlib.OracleMesh(m)
    flags = O.flag(z["u"], float(z["eps"]))
    assert np.array_equal(flags, z["flags"])
    assert 99 in flags
    hs = O.youngs(z["u"], flags)
    assert n_mismatch(hs, z["halfspaces"]) == 0
    upd, est = O.evolve(hs, flags, z["velocity"], float(z["dt"]))
    assert n_mismatch(upd, z["update"]) == 0
    assert est == float(z["dt_est"])


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_golden_time_loop(kind):
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    u, t, dt = meshlib.run(meshlib.OracleMesh(m), m, z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, z["u_end"]) == 0
    assert t == float(z["time_end"]) and dt == float(z["dt_end"])
    # mass is conserved to round-off (north_star): sum of u * volume
    assert abs(((u - z["u_run"]) * m.volumes).sum()) < 1e-15


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_golden_swartz(kind):
    """ModifiedSwartzReconstruction in 2D (reconstruction/modifiedswartz.hh:70-153) on the mesh"""
    z = golden(f"mesh2d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    O = meshlib.OracleMesh(m)
    hs = O.swartz(z["u"], z["flags"], 10)
    assert n_mismatch(hs, z["halfspaces_swartz"]) == 0
    assert n_mismatch(hs, z["halfspaces"]) > 0          # the iteration does change the Young's normals
    u, t, dt = meshlib.run(O, m, z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]), swartz=True)
    assert n_mismatch(u, z["u_end_swartz"]) == 0
    assert t == float(z["time_end_swartz"]) and dt == float(z["dt_end_swartz"])


@pytest.mark.skipif(not meshlib.ref_available(), reason="reference harness not built (needs /root/reference)")
@pytest.mark.parametrize("kind,n,seed", [("tri", 40, 1), ("quad", 48, 2), ("mixed", 36, 3), ("tri", 17, 4)])
def test_oracle_matches_live_reference(kind, n, seed):
    v, cells = mesh.perturbed_square(n, kind, seed)
    m = mesh.flatten(v, cells)
    R, O = meshlib.RefMesh(m), meshlib.OracleMesh(m)
    off, ent = R.stencil()
    assert np.array_equal(off, m.stencil_offsets) and np.array_equal(ent, m.stencil)
    u = meshlib.circle_fractions(m, center=(0.45, 0.6), radius=0.22)
    vel = meshlib.rotation_velocity(m)
    a = meshlib.run(R, m, u, vel, 1e-6, 0.5, 20)
    b = meshlib.run(O, m, u, vel, 1e-6, 0.5, 20)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]
    a = meshlib.run(R, m, u, vel, 1e-6, 0.5, 10, swartz=True)
    b = meshlib.run(O, m, u, vel, 1e-6, 0.5, 10, swartz=True)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]
    # the plane-constant solve on single cells, random normals and fractions (geometry/algorithm.hh:68-111)
    rng = np.random.default_rng(seed)
    for _ in range(200):
        c = int(rng.integers(0, m.n_cells))
        nrm = rng.normal(size=2)
        nrm /= np.linalg.norm(nrm)
        f = float(rng.choice([rng.random(), 0.0, 1.0, 1e-9, 0.5]))
        assert n_mismatch(R.locate(c, nrm, f), O.locate(c, nrm, f)) == 0


# ---- 3D: tetrahedra and hexahedra (makePolyhedron simplex / cube tables, upwindPolygon3d prisms) ------------------------------

KINDS3 = ["tet", "hex"]


@pytest.mark.parametrize("kind", KINDS3)
def test_flatten3_reproduces_the_stored_mesh_topology(kind):
    z = golden(f"mesh3d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    assert m.dim == 3
    assert np.array_equal(m.stencil_offsets, z["ref_stencil_offsets"])
    assert np.array_equal(m.stencil, z["ref_stencil"])
    assert set(np.diff(m.corner_offsets).tolist()) == ({4} if kind == "tet" else {8})
    assert set(np.diff(m.face_corner_offsets).tolist()) == ({3} if kind == "tet" else {4})
    owner = np.repeat(np.arange(m.n_cells), np.diff(m.face_offsets))
    inner = m.face_neighbor >= 0
    pairs = set(zip(owner[inner].tolist(), m.face_neighbor[inner].tolist()))
    assert all((b, a) in pairs for a, b in pairs)
    assert abs(m.volumes.sum() - 1.0) < 1e-12
    # closed cells: the integration outer normals of a cell sum to zero
    s = np.zeros((m.n_cells, 3))
    np.add.at(s, owner, m.face_int_normals)
    assert np.abs(s).max() < 1e-13


@pytest.mark.parametrize("kind", KINDS3)
def test_oracle_matches_golden_operators_3d(kind):
    z = golden(f"mesh3d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    O = meshlib.OracleMesh(m)
    flags = O.flag(z["u"], float(z["eps"]))
    assert np.array_equal(flags, z["flags"])
    assert 99 in flags and 1 in flags and 3 in flags and 0 in flags
    hs = O.youngs(z["u"], flags)
    assert n_mismatch(hs, z["halfspaces"]) == 0
    upd, est = O.evolve(hs, flags, z["velocity"], float(z["dt"]))
    assert n_mismatch(upd, z["update"]) == 0
    assert est == float(z["dt_est"])


@pytest.mark.parametrize("kind", KINDS3)
def test_oracle_matches_golden_time_loop_3d(kind):
    z = golden(f"mesh3d_{kind}.npz")
    m = mesh.FlatMesh.from_npz(z)
    u, t, dt = meshlib.run(meshlib.OracleMesh(m), m, z["u_run"], z["velocity"], float(z["eps"]), float(z["cfl"]), int(z["passes"]))
    assert n_mismatch(u, z["u_end"]) == 0
    assert t == float(z["time_end"]) and dt == float(z["dt_end"])
    assert abs(((u - z["u_run"]) * m.volumes).sum()) < 1e-15


@pytest.mark.skipif(not meshlib.ref_available(3), reason="3D reference harness not built (needs /root/reference)")
@pytest.mark.parametrize("kind,n,seed", [("tet", 6, 31), ("hex", 10, 32), ("tet", 5, 33)])
def test_oracle_matches_live_reference_3d(kind, n, seed):
    v, cells = mesh.perturbed_cube(n, kind, seed)
    m = mesh.flatten3(v, cells)
    R, O = meshlib.RefMesh(m), meshlib.OracleMesh(m)
    off, ent = R.stencil()
    assert np.array_equal(off, m.stencil_offsets) and np.array_equal(ent, m.stencil)
    u = meshlib.sphere_fractions(m, center=(0.45, 0.55, 0.5), radius=0.27, seed=seed)
    vel = meshlib.rotation_velocity3(m)
    a = meshlib.run(R, m, u, vel, 1e-6, 0.5, 6)
    b = meshlib.run(O, m, u, vel, 1e-6, 0.5, 6)
    assert n_mismatch(a[0], b[0]) == 0 and a[1:] == b[1:]
    # locateHalfSpace( Polyhedron ) on single cells, random and axis-aligned normals, extreme fractions
    rng = np.random.default_rng(seed)
    for it in range(150):
        c = int(rng.integers(0, m.n_cells))
        nrm = rng.normal(size=3)
        if it % 5 == 0:
            nrm[:] = 0
            nrm[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])
        nrm /= np.linalg.norm(nrm)
        f = float(rng.choice([rng.random(), 0.0, 1.0, 1e-9, 0.5]))
        assert n_mismatch(R.locate(c, nrm, f), O.locate(c, nrm, f)) == 0
