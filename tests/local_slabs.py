"""Test helper: the distributed step (vofx_comm_*, include/vofx.h) with ALL slabs in one process -- one context and one
stream per slab, connected by vofx_comm_connect_local.  On a one-GPU box every context sits on the same device, so the
whole multi-GPU code path (halo pushes from the update kernel, epoch signals, MIN through the windows, dense flags on
the halo layers + incremental flags inside) is exercised without NCCL; with several GPUs pass `devices`."""
import ctypes as C

import torch

import dune_vof_b200.operators as ops
from dune_vof_b200 import lib as vlib
from dune_vof_b200.decomposition import HALO_WIDTH, _DevicePtr, split_last_axis


class LocalSlabs:
    def __init__(self, n, world, cfl, eps, recon, problem, with_curvature=False, layers=None, devices=None, max_iterations=10):
        self.n, self.world, self.dim = tuple(n), world, len(n)
        self.L = vlib.load()
        halo = HALO_WIDTH[int(recon)]
        if with_curvature:
            halo = max(halo, 4)
        if layers is None:
            layers = [split_last_axis(n[-1], world, r)[1] for r in range(world)]
        assert sum(layers) == n[-1]
        self.layer_counts, self.halo = list(layers), halo
        self.layer_cells = 1
        for x in n[:-1]:
            self.layer_cells *= x
        self.params = vlib.StepParams(int(recon), int(max_iterations), float(eps), float(cfl), 1 if with_curvature else 0)
        self.grids, self.streams, self.colors = [], [], []
        for r in range(world):
            lo = [0] * self.dim
            lo[-1] = sum(layers[:r])
            n_local = list(n)
            n_local[-1] = layers[r]
            dev = devices[r] if devices else torch.cuda.current_device()
            g = ops.CartesianGrid(n, lo=lo, n_local=n_local, halo=halo, device=dev)
            g.set_velocity_builtin(problem)
            self.grids.append(g)
            self.streams.append(torch.cuda.Stream(device=g.device))
        arr = (C.c_void_p * world)(*[g._ctx for g in self.grids])
        vlib.check(self.L.vofx_comm_connect_local(arr, world))
        for g in self.grids:
            self.colors.append(torch.as_tensor(_DevicePtr(self.L.vofx_color_device(g._ctx), g.cells), device=g.device))

    def owned(self, r):
        H = self.halo
        return self.colors[r].view(self.layer_counts[r] + 2 * H, -1)[H:H + self.layer_counts[r]]

    def load(self, u_global):
        layers = u_global.view(self.n[-1], -1)
        for r in range(self.world):
            lo = sum(self.layer_counts[:r])
            self.colors[r].zero_()
            self.owned(r).copy_(layers[lo:lo + self.layer_counts[r]])

    def begin(self, time=0.0, dt=0.0):
        for g, st in zip(self.grids, self.streams):
            with torch.cuda.stream(st):
                g.use_current_stream()
                vlib.check(self.L.vofx_step_begin(g._ctx, float(time), float(dt)))
        for g in self.grids:
            torch.cuda.synchronize(g.device)
        for g in self.grids:
            vlib.check(self.L.vofx_halo_push(g._ctx))
        for g in self.grids:
            torch.cuda.synchronize(g.device)

    def step(self):
        for g, st in zip(self.grids, self.streams):
            with torch.cuda.stream(st):
                g.use_current_stream()
                vlib.check(self.L.vofx_step_distributed(g._ctx, C.byref(self.params)))

    def capture(self):
        self.graphs = []
        for g in self.grids:
            torch.cuda.synchronize(g.device)
        for g in self.grids:
            gr = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(device=g.device)
            with torch.cuda.graph(gr, stream=cap):
                g.use_current_stream()
                vlib.check(self.L.vofx_step_distributed(g._ctx, C.byref(self.params)))
            self.graphs.append(gr)
        for g in self.grids:
            torch.cuda.synchronize(g.device)

    def replay(self):
        # every graph on its slab's stream: on one stream the first graph's waits would never see the second graph's signals
        for gr, st in zip(self.graphs, self.streams):
            with torch.cuda.stream(st):
                gr.replay()

    def gather(self):
        for g in self.grids:
            torch.cuda.synchronize(g.device)
        return torch.cat([self.owned(r).reshape(-1).to(self.grids[0].device) for r in range(self.world)])

    def states(self):
        out = []
        for g in self.grids:
            t, dt, de, m = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
            vlib.check(self.L.vofx_step_state(g._ctx, C.byref(t), C.byref(dt), C.byref(de), C.byref(m)))
            out.append((t.value, dt.value, de.value, m.value))
        return out

    def close(self):
        for g in self.grids:
            g.close()
