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
                 device=None, group=None, with_curvature=False, overlap=True):
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
        self.alg = ops.Algorithm(self.grid, cfl, eps, reconstruction_kind, max_iterations, with_curvature)
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
