"""Worker of tests/test_gpu_multirank.py, launched by torch.distributed.run with one rank per GPU: the slab-decomposed
run must reproduce the single-context run BIT FOR BIT (coordinates come from global indices, the update is an
owner-computes gather and the time step is a MIN all-reduce)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dune_vof_b200.operators as ops
    from dune_vof_b200.decomposition import SlabAlgorithm, split_last_axis

    failures = []
    cases = [((48, 40, 64), ops.PROB_LEVEQUE, ops.RECON_YOUNGS, 6), ((32, 32, 48), ops.PROB_SFLOW, ops.RECON_HF, 4),
             ((64, 96), ops.PROB_SLOTTED_CYLINDER, ops.RECON_HF, 5), ((24, 24, 32), ops.PROB_ROTATING_CIRCLE, ops.RECON_SWARTZ, 2)]
    for n, problem, recon, passes in cases:
        # serial reference on this rank's GPU
        g1 = ops.CartesianGrid(n, device=local)
        g1.set_velocity_builtin(problem)
        if len(n) == 2 and problem == ops.PROB_SLOTTED_CYLINDER:
            u1 = g1.init(problem)
        else:
            u1 = g1.init(problem)
        u0 = u1.clone()
        a1 = ops.Algorithm(g1, cfl=0.5, eps=1e-6, reconstruction_kind=recon)
        st1 = a1.run_passes(u1, passes)

        slab = SlabAlgorithm(n, rank, world, 0.5, 1e-6, recon, device=local)
        slab.grid.set_velocity_builtin(problem)
        uh = slab.grid.empty("color")
        lo, cnt = split_last_axis(n[-1], world, rank)
        layer = u0.numel() // n[-1]
        slab.owned(uh).copy_(u0.view(n[-1], layer)[lo:lo + cnt])
        slab.begin(0.0, 0.0)
        for _ in range(passes):
            slab.step(uh)
        st = slab.state()
        torch.cuda.synchronize()
        mine = slab.owned(uh).contiguous()
        ref = u1.view(n[-1], layer)[lo:lo + cnt]
        same = torch.equal(mine.view(torch.int64), ref.contiguous().view(torch.int64))
        ok = torch.tensor([1 if (same and st.time == st1.time and st.dt == st1.dt) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) != 1:
            failures.append((n, problem, recon))
        if rank == 0:
            print(f"case n={n} problem={problem} recon={recon}: {'bit-identical' if int(ok.item()) == 1 else 'MISMATCH'} "
                  f"(time={st.time!r} dt={st.dt!r})", flush=True)
        # the same passes replayed from a CUDA graph of one captured pass (what bench.py times); Swartz sizes its buffers
        # from a host read-back and is not capturable
        if recon != ops.RECON_SWARTZ:
            slab2 = SlabAlgorithm(n, rank, world, 0.5, 1e-6, recon, device=local)
            slab2.grid.set_velocity_builtin(problem)
            u2 = slab2.grid.empty("color")
            slab2.owned(u2).copy_(u0.view(n[-1], layer)[lo:lo + cnt])
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                slab2.grid.use_current_stream()
                slab2.begin(0.0, 0.0)
                slab2.step(u2)                                   # first pass eagerly (allocations, NCCL warm-up)
            torch.cuda.synchronize()
            slab2.capture(u2)
            for _ in range(passes - 1):
                slab2.replay()
            torch.cuda.synchronize()
            st2 = slab2.state()
            same2 = torch.equal(slab2.owned(u2).contiguous().view(torch.int64), ref.contiguous().view(torch.int64))
            ok2 = torch.tensor([1 if (same2 and st2.time == st1.time and st2.dt == st1.dt) else 0], device="cuda")
            dist.all_reduce(ok2, op=dist.ReduceOp.MIN)
            if int(ok2.item()) != 1:
                failures.append((n, problem, recon, "graph"))
            if rank == 0:
                print(f"   graph replay: {'bit-identical' if int(ok2.item()) == 1 else 'MISMATCH'}", flush=True)
        # host-resident slab: H2D every pass, changed (cell, value) pairs written back (what bench.py's e2e does for N > 1)
        if recon == ops.RECON_YOUNGS:
            import ctypes as C
            from dune_vof_b200.lib import check
            slab3 = SlabAlgorithm(n, rank, world, 0.5, 1e-6, recon, device=local)
            g3 = slab3.grid
            g3.set_velocity_builtin(problem)
            u3 = g3.empty("color")
            own3 = slab3.owned(u3)
            host = u0.view(n[-1], layer)[lo:lo + cnt].contiguous().cpu().pin_memory()
            check(g3._L.vofx_track_changes(g3._ctx, 1))
            slab3.begin(0.0, 0.0)
            nchg = C.c_int64(0)
            for _ in range(passes):
                own3.copy_(host.view_as(own3), non_blocking=True)
                slab3.step(u3)
                check(g3._L.vofx_changed_fetch(g3._ctx, C.c_void_p(host.data_ptr()), C.c_int64(-slab3.exchanger.halo * slab3.layer_cells), C.byref(nchg)))
            same3 = torch.equal(host.view(torch.int64), ref.contiguous().cpu().view(torch.int64))
            ok3 = torch.tensor([1 if same3 else 0], device="cuda")
            dist.all_reduce(ok3, op=dist.ReduceOp.MIN)
            if int(ok3.item()) != 1:
                failures.append((n, problem, recon, "host slab"))
            if rank == 0:
                print(f"   host-resident slab with sparse write-back: {'bit-identical' if int(ok3.item()) == 1 else 'MISMATCH'}", flush=True)
        g1.close()
    dist.barrier()
    torch.cuda.synchronize()
    if failures:
        print("MULTIRANK FAILED", failures, flush=True)
    elif rank == 0:
        print("MULTIRANK PASSED", flush=True)
    sys.stdout.flush()
    os._exit(1 if failures else 0)     # no teardown of captured graphs / communicators


if __name__ == "__main__":
    sys.exit(main())
