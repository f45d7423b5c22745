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
    """exposes a raw device address as a CUDA array of `count` float64 / uint8 entries (for torch.as_tensor)"""

    def __init__(self, ptr: int, count: int = 1, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def split_by_cost(costs, world: int, min_layers: int = 1):
    """contiguous split of len(costs) layers into `world` slabs with (nearly) equal summed cost: returns the list of layer
    counts.  costs[k] = estimated work of layer k (e.g. ns_per_cell * cells + us_per_mixed_cell * mixed cells of the layer);
    every slab gets at least `min_layers` layers (the halo width).  Replaces gridPtr->loadBalance() (example/vof.cc:63)."""
    n = len(costs)
    if world * min_layers > n:
        raise ValueError("more slabs x minimal thickness than layers")
    total = float(sum(costs))
    counts, start, acc = [], 0, 0.0
    for r in range(world):
        remaining_slabs = world - r - 1
        if remaining_slabs == 0:
            counts.append(n - start)
            break
        target = total * (r + 1) / world
        end = start + min_layers
        acc_end = acc + float(sum(costs[start:end]))
        # extend while the next layer brings the running sum closer to the target, leaving room for the other slabs
        while end < n - remaining_slabs * min_layers and abs(acc_end + costs[end] - target) <= abs(acc_end - target):
            acc_end += costs[end]
            end += 1
        counts.append(end - start)
        start, acc = end, acc_end
    return counts


class SlabAlgorithm:
    """Algorithm (operators.Algorithm) on the slab of this rank.

    transport "peer" (default for world > 1): the colour function lives in the context's window, the update kernel stores
    every new boundary value straight into the neighbours' halos over NVLink and the time-step MIN travels through the same
    peer-mapped windows (vofx_comm_*, include/vofx.h) -- torch.distributed only carries the 80-byte window handles at start-up
    and the two barriers of begin().  transport "nccl": colour halo exchange (grouped send/recv) and MIN all-reduce per
    pass over torch.distributed on a caller-owned colour tensor (the round-1 path; kept for host-resident colour functions
    and as the fallback where peer mapping is unavailable)."""

    def __init__(self, n_global, rank, world, cfl, eps, reconstruction_kind=0, max_iterations=10, origin=None, h=None,
                 device=None, group=None, with_curvature=False, overlap=True, incremental=False, transport=None, layers=None):
        import torch
        from . import operators as ops
        from .lib import check
        self.rank, self.world, self.group = rank, world, group
        self.transport = transport if transport is not None else ("peer" if world > 1 else "local")
        dim = len(n_global)
        halo = HALO_WIDTH[int(reconstruction_kind)] if world > 1 else 0
        if with_curvature:
            halo = max(halo, 4) if world > 1 else 0
        if layers is None:
            layers = [split_last_axis(n_global[-1], world, r)[1] for r in range(world)]
        if len(layers) != world or sum(layers) != n_global[-1]:
            raise ValueError("layers must hold one count per rank and sum to the grid's last extent")
        self.layer_counts = [int(x) for x in layers]
        lo_last, n_last = sum(self.layer_counts[:rank]), self.layer_counts[rank]
        lo = [0] * dim
        lo[-1] = lo_last
        n_local = list(n_global)
        n_local[-1] = n_last
        self.grid = ops.CartesianGrid(n_global, origin=origin, h=h, lo=lo, n_local=n_local, halo=halo, device=device)
        self.alg = ops.Algorithm(self.grid, cfl, eps, reconstruction_kind, max_iterations, with_curvature,
                                 incremental or self.transport == "peer")
        self.exchanger = HaloExchanger(n_last, halo, rank, world, group)
        self.layer_cells = self.grid.cells // (n_last + 2 * halo)
        self._dt_est = None
        self._comm = None
        self.overlap = overlap
        self._color = None
        g = self.grid
        if self.transport == "peer" and world > 1:
            import torch.distributed as dist
            handle = (C.c_ubyte * 80)()
            check(g._L.vofx_comm_export(g._ctx, handle))
            gathered = [None] * world
            dist.all_gather_object(gathered, bytes(handle), group=group)
            blob = b"".join(gathered)
            counts = (C.c_int32 * world)(*self.layer_counts)
            check(g._L.vofx_comm_connect(g._ctx, rank, world, counts, C.c_char_p(blob)))
            # the algorithm's containers are the context-resident ones
            self.alg.flags = torch.as_tensor(_DevicePtr(g._L.vofx_flags_device(g._ctx), g.cells, "|u1"), device=g.device)
            self.alg.reconstructions = torch.as_tensor(_DevicePtr(g._L.vofx_halfspaces_device(g._ctx), g.cells * (dim + 1)),
                                                       device=g.device).view(g.cells, dim + 1)
        elif world > 1:
            ptr = g._L.vofx_dt_est_device(g._ctx)
            self._dt_est = torch.as_tensor(_DevicePtr(ptr), device=g.device)

    def color(self):
        """the colour tensor of this rank (storage size incl. halos): the context-resident one for the peer transport"""
        import torch
        g = self.grid
        if self._color is None:
            if self.transport == "peer" and self.world > 1:
                self._color = torch.as_tensor(_DevicePtr(g._L.vofx_color_device(g._ctx), g.cells), device=g.device)
            else:
                self._color = g.empty("color")
        return self._color

    def layers(self, t):
        return t.view(self.exchanger.owned + 2 * self.exchanger.halo, -1)

    def owned(self, t):
        H = self.exchanger.halo
        v = self.layers(t)
        return v[H:H + self.exchanger.owned] if H else v

    def begin(self, time=0.0, dt=0.0):
        """start of a loop; for the peer transport the owned layers of color() must hold the initial data: the halos are
        filled here (two host barriers, the only ones of the run)"""
        import torch
        from .lib import check
        self.alg.begin(time, dt)
        if self.transport == "peer" and self.world > 1:
            import torch.distributed as dist
            g = self.grid
            torch.cuda.synchronize(g.device)
            dist.barrier(group=self.group)
            g.use_current_stream()
            check(g._L.vofx_halo_push(g._ctx))
            torch.cuda.synchronize(g.device)
            dist.barrier(group=self.group)

    def step(self, uh):
        """one loop pass.  peer: one C call, no host communication.  nccl: the colour halos travel while the interior layers
        are flagged, the time-step estimate is MIN-reduced while the update runs (vofx_step_phase, include/vofx.h)"""
        import torch
        from .lib import check
        g, a = self.grid, self.alg
        if self.world == 1:
            a.step(uh)
            return
        if self.transport == "peer":
            if uh.data_ptr() != self.color().data_ptr():
                raise ValueError("the peer transport steps the context-resident colour function: pass SlabAlgorithm.color()")
            check(g._L.vofx_step_distributed(g._ctx, C.byref(a.params)))
            return
        import torch.distributed as dist
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
        """capture one loop pass (every kernel incl. the peer signals, or the NCCL operations on both streams) into a CUDA
        graph; replay() then costs one launch per pass on the host.  The pass only reads device-resident state (time, dt,
        counters, epoch), so the captured work is the same every step."""
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

    def close(self):
        """collective: unmap the peers' windows on every rank, barrier, then free (an exported window must outlive its
        mappings in the other processes)"""
        import torch
        from .lib import check
        g = self.grid
        if getattr(g, "_ctx", None) is None:
            return
        torch.cuda.synchronize(g.device)
        if hasattr(self, "_graph"):
            del self._graph
        if self.transport == "peer" and self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
            check(g._L.vofx_comm_disconnect(g._ctx))
            dist.barrier(group=self.group)
        self._color = None
        g.close()

    def set_velocity_builtin(self, problem):
        self.grid.set_velocity_builtin(problem)
        self._velocity_problem = int(problem)

    def layer_histogram(self):
        """mixed cells per OWNED layer from the flags of the last pass (numpy int64): the input of split_by_cost"""
        import torch
        f = self.owned(self.alg.flags)
        return ((f == 1) | (f == 2)).sum(dim=1).cpu().numpy().astype("int64")

    def global_layer_costs(self, ns_per_cell=1e-4, ns_per_mixed_cell=5.0):
        """estimated cost of every layer of the whole grid (all ranks): ns_per_cell * cells + ns_per_mixed_cell * mixed cells,
        from the flags of the last pass; the defaults are the measured costs of the resident loop on B200 (1e-4 ns per cell: the dense
        part of a pass is the mark scan only, a mixed cell costs about 5 ns)"""
        import numpy as np
        hist = self.layer_histogram()
        if self.world > 1:
            import torch.distributed as dist
            parts = [None] * self.world
            dist.all_gather_object(parts, hist, group=self.group)
            hist = np.concatenate(parts)
        return ns_per_cell * self.layer_cells + ns_per_mixed_cell * hist

    def rebalanced(self, uh, ns_per_cell=1e-4, ns_per_mixed_cell=5.0, min_gain=0.05):
        """Load balancing of the band (the reference calls gridPtr->loadBalance() once, example/vof.cc:63): a new slab
        decomposition with equal estimated cost per rank, the colour function migrated over torch.distributed (contiguous
        layer ranges, rank to rank) and the loop state (time, dt) carried over.  Returns (algorithm, colour tensor): the
        same objects if the estimated gain of the slowest rank is below `min_gain`.  Collective: every rank must call it
        between two passes.  The results do not depend on the decomposition (bit-identical to the serial run)."""
        import numpy as np
        import torch
        costs = self.global_layer_costs(ns_per_cell, ns_per_mixed_cell)
        counts = split_by_cost(costs, self.world, min_layers=max(self.exchanger.halo, 1))
        edges_old = np.concatenate([[0], np.cumsum(self.layer_counts)])
        edges_new = np.concatenate([[0], np.cumsum(counts)])
        worst_old = max(costs[edges_old[r]:edges_old[r + 1]].sum() for r in range(self.world))
        worst_new = max(costs[edges_new[r]:edges_new[r + 1]].sum() for r in range(self.world))
        if list(counts) == list(self.layer_counts) or worst_new > (1.0 - min_gain) * worst_old:
            return self, uh
        st = self.state()
        g = self.grid
        p = self.alg.params
        new = SlabAlgorithm(g.n, self.rank, self.world, p.cfl, p.eps, p.reconstruction, p.max_iterations, origin=g.origin, h=g.h,
                            device=g.device.index, group=self.group, with_curvature=bool(p.with_curvature), overlap=self.overlap,
                            incremental=True, transport=self.transport, layers=counts)
        if getattr(self, "_velocity_problem", None) is not None:
            new.set_velocity_builtin(self._velocity_problem)
        u_new = new.color()
        # layer ranges that move from rank s (old owner) to rank d (new owner)
        import torch.distributed as dist
        src, dst = self.owned(uh), new.owned(u_new)
        ops_list = []
        for other in range(self.world):
            lo = max(edges_old[self.rank], edges_new[other])          # I send [lo, hi) to `other`
            hi = min(edges_old[self.rank + 1], edges_new[other + 1])
            if hi > lo:
                a = src[lo - edges_old[self.rank]:hi - edges_old[self.rank]]
                if other == self.rank:
                    dst[lo - edges_new[self.rank]:hi - edges_new[self.rank]].copy_(a)
                else:
                    ops_list.append(dist.P2POp(dist.isend, a, other, self.group))
            lo = max(edges_new[self.rank], edges_old[other])          # I receive [lo, hi) from `other`
            hi = min(edges_new[self.rank + 1], edges_old[other + 1])
            if hi > lo and other != self.rank:
                ops_list.append(dist.P2POp(dist.irecv, dst[lo - edges_new[self.rank]:hi - edges_new[self.rank]], other, self.group))
        if ops_list:
            for req in dist.batch_isend_irecv(ops_list):
                req.wait()
        torch.cuda.synchronize(g.device)
        new.begin(st.time, st.dt)
        return new, u_new


