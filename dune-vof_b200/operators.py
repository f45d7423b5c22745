"""Host-side mirror of dune-vof's operator API on top of the vofx C ABI.

Same names, argument meaning and call sequence as the reference (paths relative to the reference's dune/vof/):

    FlagOperator(eps)(color, flags)                                   flagging.hh:30-70
    ModifiedYoungsReconstruction / HeightFunctionReconstruction /
      ModifiedSwartzReconstruction (grid)(color, reconstructions, flags)   reconstruction/*.hh
    reconstruction(grid)            -> the default factory (height function with Young's initialiser), reconstruction.hh:29-37
    Evolution(grid)(reconstructions, flags, velocity, deltaT, update) -> dtEst      evolution/evolution.hh:48-76
    CartesianHeightFunctionCurvature(grid)(color, reconstructions, flags, curvature)   curvature/cartesianheightfunctioncurvature.hh:50-87
    ColorFunction.axpy(a, x)                                          colorfunction.hh:31-37
    Algorithm(grid, problem, cfl, eps)(uh, start, end)                algorithm.hh:37-120 (without l1error / writer)

Containers are torch CUDA tensors (device memory and streams are PyTorch's job here; the arithmetic is not):
colour float64[cells], flags uint8[cells], reconstructions float64[cells, dim+1], curvature float64[cells].
Nothing in this module computes on the CPU; without libvofx.so or a GPU every call raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import lib as _lib
from .lib import (GridDesc, StepParams, VofxError, check, RECON_YOUNGS, RECON_HF, RECON_SWARTZ,  # noqa: F401
                  PROB_ROTATING_CIRCLE, PROB_SFLOW, PROB_SLOTTED_CYLINDER, PROB_LINEAR_WALL, PROB_LEVEQUE)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise VofxError("no CUDA device: vofx has no CPU fallback")
    return torch


class CartesianGrid:
    """The flattened structured grid a DUNE YaspGrid/SPGrid maps to (include/vofx.h, "Grid convention").

    n: cells per axis of the whole grid.  For a z-slab (y-slab in 2D) pass lo/n_local/halo."""

    def __init__(self, n, origin=None, h=None, lo=None, n_local=None, halo=0, device=None):
        torch = _torch()
        self.dim = len(n)
        if self.dim not in (2, 3):
            raise VofxError("dim must be 2 or 3")
        self.n = tuple(int(x) for x in n)
        self.origin = tuple(float(x) for x in (origin if origin is not None else [0.0] * self.dim))
        self.h = tuple(float(x) for x in (h if h is not None else [1.0 / x for x in self.n]))
        self.lo = tuple(int(x) for x in (lo if lo is not None else [0] * self.dim))
        self.n_local = tuple(int(x) for x in (n_local if n_local is not None else self.n))
        self.halo = int(halo)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        d = GridDesc()
        d.dim = self.dim
        for k in range(3):
            used = k < self.dim
            d.n_global[k] = self.n[k] if used else 1
            d.origin[k] = self.origin[k] if used else 0.0
            d.h[k] = self.h[k] if used else 1.0
            d.lo[k] = self.lo[k] if used else 0
            d.n_local[k] = self.n_local[k] if used else 1
        d.halo = self.halo
        self._L = _lib.load()
        handle = C.c_void_p()
        check(self._L.vofx_create(C.byref(d), self.device.index, C.byref(handle)))
        self._ctx = handle
        self.cells = int(self._L.vofx_cells(self._ctx))
        self.storage_shape = tuple(reversed([self.n_local[k] + (2 * self.halo if k == self.dim - 1 else 0)
                                             for k in range(self.dim)]))   # (z, y, x): x fastest
        self.velocity_problem = None

    # -- plumbing ------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.vofx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_current_stream(self):
        torch = _torch()
        check(self._L.vofx_set_stream(self._ctx, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def empty(self, kind):
        torch = _torch()
        if kind == "color" or kind == "curvature" or kind == "update":
            return torch.zeros(self.cells, dtype=torch.float64, device=self.device)
        if kind == "flags" or kind == "valid":
            return torch.zeros(self.cells, dtype=torch.uint8, device=self.device)
        if kind == "reconstructions":
            return torch.zeros(self.cells, self.dim + 1, dtype=torch.float64, device=self.device)
        raise ValueError(kind)

    def _ptr(self, t, dtype, numel):
        torch = _torch()
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous() or t.numel() != numel:
            raise VofxError(f"expected a contiguous {dtype} CUDA tensor with {numel} entries on {self.device}")
        return C.c_void_p(t.data_ptr())

    def pc(self, t):
        return self._ptr(t, _torch().float64, self.cells)

    def pf(self, t):
        return self._ptr(t, _torch().uint8, self.cells)

    def ph(self, t):
        return self._ptr(t, _torch().float64, self.cells * (self.dim + 1))

    @property
    def launch_count(self):
        return int(self._L.vofx_launch_count(self._ctx))

    # -- velocity (Velocity< Problem, GridView >, velocity.hh) --------------------------------------------------------
    def set_velocity_builtin(self, problem):
        check(self._L.vofx_velocity_builtin(self._ctx, int(problem)))
        self.velocity_problem = int(problem)

    def set_velocity_separable(self, components):
        """components[a] = dict(coef=..., time=("const"|"cospi", period), terms=[[fx, fy(, fz)], ...]) with f callables
        of the coordinate; tables are sampled on the host at the coordinates include/vofx.h defines."""
        import numpy as np
        check(self._L.vofx_velocity_clear(self._ctx))
        for a, comp in enumerate(components):
            kind, period = comp.get("time", ("const", 1.0))
            terms = comp.get("terms", [])
            check(self._L.vofx_velocity_set_component(self._ctx, a, len(terms), float(comp.get("coef", 1.0)),
                                                       _lib.TIME_COSPI if kind == "cospi" else _lib.TIME_CONST, float(period)))
            for r, factors in enumerate(terms):
                for axis, f in enumerate(factors):
                    lo = np.array([self.origin[axis] + i * self.h[axis] for i in range(self.n[axis])])
                    tabs = [np.ascontiguousarray([f(x) for x in xs], dtype=np.float64)
                            for xs in (lo + 0.5 * self.h[axis], lo, lo + self.h[axis])]
                    check(self._L.vofx_velocity_set_factor(self._ctx, a, r, axis, *[C.c_void_p(t.ctypes.data) for t in tabs]))
        self.velocity_problem = -1

    def face_velocity(self, time, cell, face):
        import numpy as np
        v = np.zeros(self.dim)
        check(self._L.vofx_face_velocity(self._ctx, float(time), int(cell), int(face), C.c_void_p(v.ctypes.data)))
        return v

    # -- initial data (Average< Problem >, test/average.hh) -----------------------------------------------------------
    def init(self, problem, time=0.0, out=None):
        u = out if out is not None else self.empty("color")
        self.use_current_stream()
        check(self._L.vofx_init(self._ctx, int(problem), float(time), self.pc(u)))
        return u


class ColorFunction:
    """dune-vof ColorFunction (colorfunction.hh:20-52): cell values + axpy."""

    def __init__(self, grid: CartesianGrid, values=None):
        self.grid = grid
        self.values = values if values is not None else grid.empty("color")

    def axpy(self, a, x: "ColorFunction"):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_axpy(g._ctx, g.pc(self.values), float(a), g.pc(x.values)))

    def clear(self):
        self.values.zero_()


class FlagOperator:
    """flagging.hh:24-74"""

    def __init__(self, grid: CartesianGrid, eps: float):
        self.grid, self.eps = grid, float(eps)

    def __call__(self, color, flags, communicate=True):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_flag(g._ctx, g.pc(color), self.eps, g.pf(flags)))


class _Reconstruction:
    kind = RECON_YOUNGS

    def __init__(self, grid: CartesianGrid, max_iterations: int = 10):
        self.grid, self.max_iterations = grid, int(max_iterations)

    def __call__(self, color, reconstructions, flags, communicate=True):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_reconstruct(g._ctx, self.kind, g.pc(color), g.pf(flags), g.ph(reconstructions), self.max_iterations))


class ModifiedYoungsReconstruction(_Reconstruction):
    """reconstruction/modifiedyoungs.hh:30-141"""
    kind = RECON_YOUNGS


class HeightFunctionReconstruction(_Reconstruction):
    """reconstruction/heightfunction.hh:35-287 with the ModifiedYoungs initialiser"""
    kind = RECON_HF


class ModifiedSwartzReconstruction(_Reconstruction):
    """reconstruction/modifiedswartz.hh:36-252 with the ModifiedYoungs initialiser; maxIterations default 10 (:51,55)"""
    kind = RECON_SWARTZ


def reconstruction(grid: CartesianGrid):
    """the reference's default factory: height function with Young's initialiser (reconstruction.hh:29-37)"""
    return HeightFunctionReconstruction(grid)


class Evolution:
    """evolution/evolution.hh:37-181.  `velocity` is the grid's separable velocity (set_velocity_*) frozen at `time`."""

    def __init__(self, grid: CartesianGrid):
        self.grid = grid

    def __call__(self, reconstructions, flags, time, deltaT, update):
        g = self.grid
        g.use_current_stream()
        dt_est = C.c_double(0.0)
        check(g._L.vofx_evolve(g._ctx, g.ph(reconstructions), g.pf(flags), float(time), float(deltaT), g.pc(update), C.byref(dt_est)))
        return dt_est.value


def evolution(grid: CartesianGrid):
    """evolution.hh:25-30"""
    return Evolution(grid)


class CartesianHeightFunctionCurvature:
    """curvature/cartesianheightfunctioncurvature.hh:30-190"""

    def __init__(self, grid: CartesianGrid):
        self.grid = grid
        self.satisfies_constraint = grid.empty("valid")

    def __call__(self, color, reconstructions, flags, curvature, communicate=True):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_curvature(g._ctx, g.pc(color), g.ph(reconstructions), g.pf(flags), g.pc(curvature), g.pf(self.satisfies_constraint)))


@dataclass
class StepResult:
    time: float
    dt: float
    dt_est: float
    mixed_cells: int


class SwartzCurvature:
    """curvature/swartzcurvature.hh:27-181 (2D)"""

    def __init__(self, grid: CartesianGrid):
        self.grid = grid

    def __call__(self, color, reconstructions, flags, curvature, communicate=True):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_curvature_swartz(g._ctx, g.ph(reconstructions), g.pf(flags), g.pc(curvature)))


def diagnostics(grid: CartesianGrid, color, exact=None):
    """sum, min, max, cellwiseL1error (test/errors.hh:138-151) and mass of the owned cells, reduced on the device"""
    out = (C.c_double * 5)()
    ex = grid.pc(exact) if exact is not None else C.c_void_p(0)
    check(grid._L.vofx_diagnostics(grid._ctx, grid.pc(color), ex, out))
    return {"sum": out[0], "min": out[1], "max": out[2], "l1": out[3], "mass": out[4]}


def get_interface_vertices(grid: CartesianGrid, reconstructions, flags):
    """getInterfaceVertices (utility.hh:19-48) + MixedCellMapper::update (mixedcellmapper.hh:82-86): returns numpy arrays
    (cells, offsets, vertices): the owned mixed cells in ascending index and the CSR of their interface polygons."""
    import numpy as np
    g = grid
    g.use_current_stream()
    nm, nv = C.c_int64(0), C.c_int64(0)
    check(g._L.vofx_interface(g._ctx, g.ph(reconstructions), g.pf(flags), C.byref(nm), C.byref(nv)))
    cells = np.empty(nm.value, dtype=np.int32)
    offsets = np.zeros(nm.value + 1, dtype=np.int64)
    vertices = np.empty((nv.value, g.dim), dtype=np.float64)
    check(g._L.vofx_interface_fetch(g._ctx, C.c_void_p(cells.ctypes.data), C.c_void_p(offsets.ctypes.data), C.c_void_p(vertices.ctypes.data)))
    return cells, offsets, vertices


class Algorithm:
    """The time loop of algorithm.hh:37-120 on a device-resident colour function (no l1error, no writer).

    `reconstruction_kind` selects the operator (the reference hard-wires it in reconstruction.hh); `step()` is one
    loop pass, fully on the device (time and dt live there); `__call__(uh, start, end)` replays Algorithm::operator()
    and stops at the reference's pass."""

    def __init__(self, grid: CartesianGrid, cfl: float, eps: float, reconstruction_kind=RECON_HF, max_iterations=10,
                 with_curvature=False, incremental=False):
        """incremental: the caller promises that between two passes only the passes themselves (and a halo exchange)
        change the colour function, so flags are maintained from the previous band instead of a dense pass per step
        (vofx_set_incremental_flagging, include/vofx.h)"""
        self.grid = grid
        if incremental:
            check(grid._L.vofx_set_incremental_flagging(grid._ctx, 1))
        self.params = StepParams(int(reconstruction_kind), int(max_iterations), float(eps), float(cfl), 1 if with_curvature else 0)
        self.flags = grid.empty("flags")
        self.reconstructions = grid.empty("reconstructions")
        self.curvature = grid.empty("curvature") if with_curvature else None

    def begin(self, time=0.0, dt=0.0):
        g = self.grid
        g.use_current_stream()
        check(g._L.vofx_step_begin(g._ctx, float(time), float(dt)))

    def step(self, uh):
        g = self.grid
        kap = C.c_void_p(self.curvature.data_ptr()) if self.curvature is not None else C.c_void_p(0)
        check(g._L.vofx_step(g._ctx, C.byref(self.params), g.pc(uh), g.pf(self.flags), g.ph(self.reconstructions), kap))

    def band_list(self):
        """storage indices of the mixed cells of the last finished pass (numpy int32, no defined order)"""
        import numpy as np
        g = self.grid
        n = C.c_int64(0)
        check(g._L.vofx_band_list_fetch(g._ctx, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.int32)
        if n.value:
            check(g._L.vofx_band_list_fetch(g._ctx, C.c_void_p(out.ctypes.data), n.value, C.byref(n)))
        return out

    def state(self) -> StepResult:
        g = self.grid
        t, dt, de, m = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        check(g._L.vofx_step_state(g._ctx, C.byref(t), C.byref(dt), C.byref(de), C.byref(m)))
        return StepResult(t.value, dt.value, de.value, m.value)

    def __call__(self, uh, start, end):
        """do { pass } while( time < end ), algorithm.hh:68-95: the loop condition is read back after EVERY pass (8 bytes),
        so the loop stops at exactly the pass the reference stops at and (uh, time, dt) carry its bits; use step() /
        capture() + replay() for loops whose length is known in advance."""
        self.begin(start, 0.0)
        self.grid.use_current_stream()
        while True:
            self.step(uh)
            st = self.state()
            if st.time >= end:
                return st

    def capture(self, uh):
        """capture one loop pass into a CUDA graph (it only reads device-resident state: time, dt, counters); replay()
        then costs one host launch per pass -- what small grids need, whose passes are launch-bound (C1: 5 kernels
        of a few microseconds each).  Run at least one pass eagerly first (dense flags; the batched 3D Swartz pass sizes its buffers there)."""
        torch = _torch()
        g = self.grid
        self._graph_stream = torch.cuda.Stream(device=g.device)
        self._graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(g.device)
        with torch.cuda.graph(self._graph, stream=self._graph_stream):
            g.use_current_stream()
            self.step(uh)
        torch.cuda.synchronize(g.device)
        return self._graph

    def replay(self):
        self._graph.replay()

    def run_passes(self, uh, passes, time=0.0, dt=0.0):
        self.begin(time, dt)
        self.grid.use_current_stream()
        for _ in range(passes):
            self.step(uh)
        return self.state()
