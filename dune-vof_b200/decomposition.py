"""Slab decomposition of the structured grid over the GPUs of one node (one process per GPU).

Replaces the reference's MPI overlap decomposition (dune-grid's GridView::communicate; call sites flagging.hh:69,
modifiedyoungs.hh:75, evolution/evolution.hh:73,75 -- SURVEY.md section 2.3) by:
  * one halo exchange of the COLOUR function per step (width 2 for Young's, 3 for Swartz, 4 for height functions):
    flags and reconstructions of the first halo layer are recomputed locally from it, so the reference's exchanges of
    flags / reconstructions / update (All_All + Add) disappear (owner-computes gather);
  * one MIN all-reduce of the time-step estimate (gridView().comm().min, evolution.hh:75).
The last axis is the slowest-varying one, so a halo is a contiguous block of memory: no pack kernels; the exchange is
a grouped NCCL send/recv over NVLink (torch.distributed is the plumbing).  Results are identical to the serial run
because coordinates are generated from global indices (include/vofx.h).
"""
from __future__ import annotations

import ctypes as C

HALO_WIDTH = {0: 2, 1: 4, 2: 3}   # reconstruction kind -> colour halo width (DESIGN.md "Multi-GPU")


def split_last_axis(n_last: int, world: int, rank: int):
    """balanced contiguous split of n_last layers: returns (first layer, number of layers) of `rank`"""
    base, rem = divmod(n_last, world)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


class HaloExchanger:
    """Exchanges `halo` layers of a [layers, layer_size] tensor with the two neighbouring ranks.
    Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, owned_layers: int, halo: int, rank: int, world: int, group=None):
        self.owned, self.halo, self.rank, self.world, self.group = owned_layers, halo, rank, world, group
        if world > 1 and owned_layers < halo:
            raise ValueError("slab thinner than the halo")

    def exchange(self, t):
        """t: tensor whose leading dimension is owned_layers + 2*halo"""
        import torch.distributed as dist
        H, n = self.halo, self.owned
        if self.world == 1 or H == 0:
            return
        assert t.shape[0] == n + 2 * H and t.is_contiguous()
        ops = []
        lower, upper = self.rank - 1, self.rank + 1
        if lower >= 0:
            ops.append(dist.P2POp(dist.isend, t[H:2 * H], lower, self.group))        # my lowest owned layers
            ops.append(dist.P2POp(dist.irecv, t[0:H], lower, self.group))            # its highest owned layers
        if upper < self.world:
            ops.append(dist.P2POp(dist.isend, t[n:n + H], upper, self.group))        # my highest owned layers
            ops.append(dist.P2POp(dist.irecv, t[n + H:n + 2 * H], upper, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _DevicePtr:
    """exposes a raw device address as a 1-element float64 CUDA array (for torch.as_tensor)"""

    def __init__(self, ptr: int):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class SlabAlgorithm:
    """Algorithm (operators.Algorithm) on the slab of this rank: exchange colour halo -> local step -> MIN all-reduce
    of the time-step estimate -> finish."""

    def __init__(self, n_global, rank, world, cfl, eps, reconstruction_kind=0, max_iterations=10, origin=None, h=None,
                 device=None, group=None, with_curvature=False, overlap=True, incremental=False):
        import torch
        from . import operators as ops
        self.rank, self.world, self.group = rank, world, group
        dim = len(n_global)
        halo = HALO_WIDTH[int(reconstruction_kind)] if world > 1 else 0
        if with_curvature:
            halo = max(halo, 4) if world > 1 else 0
        lo_last, n_last = split_last_axis(n_global[-1], world, rank)
        lo = [0] * dim
        lo[-1] = lo_last
        n_local = list(n_global)
        n_local[-1] = n_last
        self.grid = ops.CartesianGrid(n_global, origin=origin, h=h, lo=lo, n_local=n_local, halo=halo, device=device)
        self.alg = ops.Algorithm(self.grid, cfl, eps, reconstruction_kind, max_iterations, with_curvature, incremental)
        self.exchanger = HaloExchanger(n_last, halo, rank, world, group)
        self.layer_cells = self.grid.cells // (n_last + 2 * halo)
        self._dt_est = None
        self._comm = None
        self.overlap = overlap
        if world > 1:
            ptr = self.grid._L.vofx_dt_est_device(self.grid._ctx)
            self._dt_est = torch.as_tensor(_DevicePtr(ptr), device=self.grid.device)

    def layers(self, t):
        return t.view(self.exchanger.owned + 2 * self.exchanger.halo, -1)

    def owned(self, t):
        H = self.exchanger.halo
        v = self.layers(t)
        return v[H:H + self.exchanger.owned] if H else v

    def begin(self, time=0.0, dt=0.0):
        self.alg.begin(time, dt)

    def step(self, uh):
        """one loop pass: the colour halos travel while the interior layers are flagged, the time-step estimate is
        MIN-reduced while the update runs (vofx_step_phase, include/vofx.h)"""
        import torch
        import torch.distributed as dist
        from .lib import check
        g, a = self.grid, self.alg
        if self.world == 1:
            a.step(uh)
            return
        kap = C.c_void_p(a.curvature.data_ptr()) if a.curvature is not None else C.c_void_p(0)
        args = (g._ctx, C.byref(a.params), g.pc(uh), g.pf(a.flags), g.ph(a.reconstructions), kap)
        if not self.overlap:
            self.exchanger.exchange(self.layers(uh))
            check(g._L.vofx_step_local(*args))
            dist.all_reduce(self._dt_est, op=dist.ReduceOp.MIN, group=self.group)
            check(g._L.vofx_step_finish(g._ctx, C.byref(a.params)))
            return
        main = torch.cuda.current_stream(g.device)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=g.device)
        comm = self._comm
        comm.wait_stream(main)                                   # the previous pass has updated the owned layers
        with torch.cuda.stream(comm):
            self.exchanger.exchange(self.layers(uh))
        check(g._L.vofx_step_phase(*args, 0))                    # interior flags: no halo value is read
        main.wait_stream(comm)
        check(g._L.vofx_step_phase(*args, 1))                    # boundary flags, reconstruction, fluxes
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            dist.all_reduce(self._dt_est, op=dist.ReduceOp.MIN, group=self.group)
        check(g._L.vofx_step_phase(*args, 2))                    # update
        main.wait_stream(comm)
        check(g._L.vofx_step_finish(g._ctx, C.byref(a.params)))

    def capture(self, uh):
        """capture one loop pass (kernels, halo exchange, MIN reduction, both streams) into a CUDA graph; replay() then
        costs one launch per pass on the host.  The pass only reads device-resident state (time, dt, counters), so the
        captured work is the same every step."""
        import torch
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

    def state(self):
        return self.alg.state()


class PipelinedAlgorithm:
    """The time loop on ONE GPU with the grid cut into `chunks` z-slabs that share the caller's contiguous arrays.

    Every slab is a context of its own (with halo layers that alias the neighbouring slab's owned layers, so nothing is
    copied) on a stream of its own: while the band kernels of one slab run (latency-bound, few resident warps), the
    dense flag pass of the next slab streams through HBM.  Slabs only meet twice per pass: before the in-place update
    (all reads of the old colour function are done) and for the MIN of the time-step estimates.  The result is the
    single-context result bit for bit (the slab machinery is the multi-GPU one, tests/test_gpu_multirank.py).

    MEASURED (B200, 512^3 sphere, graph replay, tools/pipeline_probe.py): 1.062 / 1.092 / 1.054 / 1.048 ms per step for
    1 / 2 / 4 / 8 chunks -- no gain: the band kernels fill the register file and shared memory of every SM (9 blocks x
    23.7 KB, 55 K registers), so the flag pass of the next slab gets one block per SM and streams at a fraction of the
    HBM rate.  bench.py therefore runs the plain single-context loop; the class stays as the tested way to run several
    contexts on windows of one contiguous array."""

    def __init__(self, n, cfl, eps, reconstruction_kind=0, chunks=2, origin=None, h=None, device=None, max_iterations=10):
        import torch
        from . import operators as ops
        self.n = tuple(int(x) for x in n)
        self.dim = len(self.n)
        self.chunks = int(chunks)
        halo = HALO_WIDTH[int(reconstruction_kind)]
        self.grids, self.params, self.offsets = [], [], []
        self.layer = 1
        for x in self.n[:-1]:
            self.layer *= x
        for k in range(self.chunks):
            lo_last, n_last = split_last_axis(self.n[-1], self.chunks, k)
            if n_last < halo + 2:
                raise ValueError("chunk thinner than its halo")
            lo = [0] * self.dim
            lo[-1] = lo_last
            n_local = list(self.n)
            n_local[-1] = n_last
            g = ops.CartesianGrid(self.n, origin=origin, h=h, lo=lo, n_local=n_local, halo=halo, device=device)
            self.grids.append(g)
            self.params.append(ops.StepParams(int(reconstruction_kind), int(max_iterations), float(eps), float(cfl), 0))
            self.offsets.append((lo_last - halo) * self.layer)         # first storage cell of the slab in the global arrays
        self.device = self.grids[0].device
        self.cells = self.layer * self.n[-1]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.chunks)]
        self._dt = [torch.as_tensor(_DevicePtr(g._L.vofx_dt_est_device(g._ctx)), device=self.device) for g in self.grids]
        for g, st in zip(self.grids, self.streams):
            with torch.cuda.stream(st):
                g.use_current_stream()
        self.flags = torch.zeros(self.cells, dtype=torch.uint8, device=self.device)
        self.reconstructions = torch.zeros(self.cells, self.dim + 1, dtype=torch.float64, device=self.device)
        self._graph = None

    def set_velocity_builtin(self, problem):
        for g in self.grids:
            g.set_velocity_builtin(problem)

    def begin(self, time=0.0, dt=0.0):
        from .lib import check
        for g in self.grids:
            check(g._L.vofx_step_begin(g._ctx, float(time), float(dt)))

    def _args(self, k, uh):
        off = self.offsets[k]
        g = self.grids[k]
        return (g._ctx, C.byref(self.params[k]), C.c_void_p(uh.data_ptr() + 8 * off), C.c_void_p(self.flags.data_ptr() + off),
                C.c_void_p(self.reconstructions.data_ptr() + 8 * (self.dim + 1) * off), C.c_void_p(0))

    def step(self, uh):
        import torch
        from .lib import check
        if uh.numel() != self.cells or uh.device != self.device or not uh.is_contiguous():
            raise ValueError("expected the contiguous colour function of the whole grid on the algorithm's device")
        main = torch.cuda.current_stream(self.device)
        L = self.grids[0]._L
        for k, st in enumerate(self.streams):                    # flags, reconstruction, fluxes of every slab
            st.wait_stream(main)
            with torch.cuda.stream(st):
                args = self._args(k, uh)
                check(L.vofx_step_phase(*args, 0))
                check(L.vofx_step_phase(*args, 1))
        for st in self.streams:                                  # nobody reads the old colour function any more
            main.wait_stream(st)
        for k, st in enumerate(self.streams):                    # in-place update of the owned cells
            st.wait_stream(main)
            with torch.cuda.stream(st):
                check(L.vofx_step_phase(*self._args(k, uh), 2))
        for st in self.streams:
            main.wait_stream(st)
        m = self._dt[0]                                          # gridView().comm().min( dtEst ), evolution.hh:75
        for d in self._dt[1:]:
            torch.minimum(m, d, out=m)
        for d in self._dt[1:]:
            d.copy_(m)
        for k, g in enumerate(self.grids):
            check(L.vofx_set_stream(g._ctx, C.c_void_p(main.cuda_stream)))
            check(L.vofx_step_finish(g._ctx, C.byref(self.params[k])))
            check(L.vofx_set_stream(g._ctx, C.c_void_p(self.streams[k].cuda_stream)))

    def capture(self, uh):
        import torch
        self._graph_stream = torch.cuda.Stream(device=self.device)
        self._graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(self.device)
        with torch.cuda.graph(self._graph, stream=self._graph_stream):
            self.step(uh)
        torch.cuda.synchronize(self.device)
        return self._graph

    def replay(self):
        self._graph.replay()

    def state(self):
        from .operators import StepResult
        g = self.grids[0]
        t, dt, de, m = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        from .lib import check
        check(g._L.vofx_step_state(g._ctx, C.byref(t), C.byref(dt), C.byref(de), C.byref(m)))
        return StepResult(t.value, dt.value, de.value, m.value)

    @property
    def launch_count(self):
        return sum(g.launch_count for g in self.grids)

    def close(self):
        for g in self.grids:
            g.close()
