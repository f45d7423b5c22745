"""Timing of the unstructured-mesh loop pass (vofx_mesh_run) on one GPU, next to the oracle on the host.

    python tools/mesh_probe.py [n] [kind]      # n x n lattice, kind tri|quad|mixed

Prints one JSON line: cells, mixed cells, ms per loop pass on the device (device-resident colour function, CUDA events
are not exposed for this path: wall time around a synchronising call over many passes), oracle ms per pass.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_vof_b200 import mesh  # noqa: E402
from oracle import meshlib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    kind = sys.argv[2] if len(sys.argv) > 2 else "tri"
    v, cells = mesh.perturbed_square(n, kind, 7)
    m = mesh.flatten(v, cells)
    u = meshlib.circle_fractions(m, center=(0.5, 0.7), radius=0.2, samples=6)
    vel = meshlib.rotation_velocity(m)
    D = mesh.MeshOperators(m)
    D.run(u, vel, 1e-6, 0.5, 5)                      # warm-up
    passes = 200
    import ctypes as C
    from dune_vof_b200 import lib
    L = lib.load()
    lib.check(L.vofx_mesh_color_upload(D._h, u.ctypes.data))
    lib.check(L.vofx_mesh_velocity_upload(D._h, vel.ctypes.data))
    t, d = C.c_double(0.0), C.c_double(0.0)
    t0 = time.perf_counter()
    lib.check(L.vofx_mesh_run(D._h, 1e-6, 0.5, passes, C.byref(t), C.byref(d)))
    gpu_ms = (time.perf_counter() - t0) * 1e3 / passes
    out = np.empty(m.n_cells)
    flags = np.empty(m.n_cells, dtype=np.uint8)
    lib.check(L.vofx_mesh_color_download(D._h, out.ctypes.data, flags.ctypes.data))
    O = meshlib.OracleMesh(m)
    t0 = time.perf_counter()
    ref = meshlib.run(O, m, u, vel, 1e-6, 0.5, 20)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 20
    ref200 = meshlib.run(O, m, u, vel, 1e-6, 0.5, passes)
    same = bool(np.array_equal(ref200[0].view(np.uint64), out.view(np.uint64)))
    print(json.dumps({"cells": m.n_cells, "kind": kind, "mixed_cells": int(np.count_nonzero((flags == 1) | (flags == 2))),
                      "gpu_ms_per_pass": gpu_ms, "gpu_cell_steps_per_s": m.n_cells / gpu_ms * 1e3,
                      "oracle_ms_per_pass_1core": cpu_ms, "bit_identical_after_200_passes": same}))


if __name__ == "__main__":
    main()
