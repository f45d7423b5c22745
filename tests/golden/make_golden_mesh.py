"""Generates tests/golden/mesh2d_{tri,quad,mixed}.npz and tests/golden/mesh3d_{tet,hex}.npz from the UNMODIFIED reference on
the unstructured grid model (oracle/_ref/libvofrefmesh.so and libvofrefmesh3.so: /root/reference's FlagOperator,
ModifiedYoungsReconstruction, ModifiedSwartzReconstruction (2D), VertexNeighborsStencil and Evolution on
oracle/shim/dune/grid/common/meshgrid.hh).  Run where /root/reference is mounted:

    python tests/golden/make_golden_mesh.py

The inputs (mesh, colour function, sampled velocity) are seeded and stored with the outputs, so the GPU box only reads
the files.  Outputs are whatever the reference returns.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dune_vof_b200 import mesh  # noqa: E402
from oracle import meshlib  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EPS, CFL, DT, PASSES = 1e-6, 0.5, 0.01, 25


def main():
    for kind, n, seed in (("tri", 24, 11), ("quad", 32, 12), ("mixed", 28, 13)):
        vertices, cells = mesh.perturbed_square(n, kind, seed)
        m = mesh.flatten(vertices, cells)
        ref = meshlib.RefMesh(m)
        s_off, s_ent = ref.stencil()
        u = meshlib.circle_fractions(m)
        u[3] = np.nan                      # flag 99 away from the interface
        vel = meshlib.rotation_velocity(m)
        flags = ref.flag(u, EPS)
        hs = ref.youngs(u, flags)
        upd, est = ref.evolve(hs, flags, vel, DT)
        hs_swartz = ref.swartz(u, flags, 10)
        u_run = meshlib.circle_fractions(m)
        u_end, t_end, dt_end = meshlib.run(ref, m, u_run, vel, EPS, CFL, PASSES)
        u_end_sw, t_end_sw, dt_end_sw = meshlib.run(ref, m, u_run, vel, EPS, CFL, PASSES, swartz=True)
        assert np.array_equal(s_off, m.stencil_offsets) and np.array_equal(s_ent, m.stencil)
        # the flattened mesh is stored field by field (mesh_*): the fixtures do not depend on numpy's rounding elsewhere
        fields = {"mesh_" + k: getattr(m, k) for k in m.__dataclass_fields__ if k not in ("dim", "face_corner_offsets")}
        np.savez_compressed(os.path.join(OUT, f"mesh2d_{kind}.npz"), **fields, ref_stencil_offsets=s_off,
                            ref_stencil=s_ent, u=u, velocity=vel,
                            eps=EPS, dt=DT, cfl=CFL, passes=PASSES, flags=flags, halfspaces=hs, update=upd, dt_est=est,
                            u_run=u_run, u_end=u_end, time_end=t_end, dt_end=dt_end,
                            halfspaces_swartz=hs_swartz, u_end_swartz=u_end_sw, time_end_swartz=t_end_sw, dt_end_swartz=dt_end_sw)
        print(kind, m.n_cells, "cells", int(np.count_nonzero((flags == 1) | (flags == 2))), "mixed")

    # 3D: tetrahedra (Kuhn split of a perturbed lattice) and hexahedra (rectilinear lattice with seeded spacings)
    for kind, n, seed, passes in (("tet", 7, 21, 8), ("hex", 12, 22, 8)):
        vertices, cells = mesh.perturbed_cube(n, kind, seed)
        m = mesh.flatten3(vertices, cells)
        ref = meshlib.RefMesh(m)
        s_off, s_ent = ref.stencil()
        u = meshlib.sphere_fractions(m)
        u[3] = np.nan
        vel = meshlib.rotation_velocity3(m)
        flags = ref.flag(u, EPS)
        hs = ref.youngs(u, flags)
        upd, est = ref.evolve(hs, flags, vel, DT)
        u_run = meshlib.sphere_fractions(m)
        u_end, t_end, dt_end = meshlib.run(ref, m, u_run, vel, EPS, CFL, passes)
        assert np.array_equal(s_off, m.stencil_offsets) and np.array_equal(s_ent, m.stencil)
        fields = {"mesh_" + k: getattr(m, k) for k in m.__dataclass_fields__}
        np.savez_compressed(os.path.join(OUT, f"mesh3d_{kind}.npz"), **fields, ref_stencil_offsets=s_off,
                            ref_stencil=s_ent, u=u, velocity=vel,
                            eps=EPS, dt=DT, cfl=CFL, passes=passes, flags=flags, halfspaces=hs, update=upd, dt_est=est,
                            u_run=u_run, u_end=u_end, time_end=t_end, dt_end=dt_end)
        print(kind, m.n_cells, "cells", int(np.count_nonzero((flags == 1) | (flags == 2))), "mixed")


if __name__ == "__main__":
    main()
