"""Worker of tests/test_gpu_multirank.py, launched by torch.distributed.run with one rank per GPU: the slab-decomposed
run must reproduce the single-context run BIT FOR BIT (coordinates come from global indices, the update is an
owner-computes gather and the time step is an exact MIN), for both transports of SlabAlgorithm:
  peer  the distributed step of the C ABI (vofx_comm_*): CUDA IPC windows, halo pushes from the update kernel
  nccl  colour halo exchange + MIN all-reduce over torch.distributed
eagerly, replayed from a CUDA graph, after a rebalancing of the slabs, and with a host-resident slab."""
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
    from dune_vof_b200.decomposition import SlabAlgorithm

    failures = []

    def agree(flag):
        ok = torch.tensor([1 if flag else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return int(ok.item()) == 1

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    cases = [((48, 40, 64), ops.PROB_LEVEQUE, ops.RECON_YOUNGS, 8), ((32, 32, 48), ops.PROB_SFLOW, ops.RECON_HF, 4),
             ((64, 96), ops.PROB_SLOTTED_CYLINDER, ops.RECON_HF, 5), ((24, 24, 32), ops.PROB_ROTATING_CIRCLE, ops.RECON_SWARTZ, 4)]
    for n, problem, recon, passes in cases:
        # serial reference on this rank's GPU
        g1 = ops.CartesianGrid(n, device=local)
        g1.set_velocity_builtin(problem)
        u1 = g1.init(problem)
        u0 = u1.clone()
        st1 = ops.Algorithm(g1, cfl=0.5, eps=1e-6, reconstruction_kind=recon).run_passes(u1, passes)
        layer = u0.numel() // n[-1]

        def reference_slab(slab):
            lo = sum(slab.layer_counts[:rank])
            return u1.view(n[-1], layer)[lo:lo + slab.layer_counts[rank]].contiguous()

        def fresh(transport, layers=None):
            slab = SlabAlgorithm(n, rank, world, 0.5, 1e-6, recon, device=local, transport=transport, layers=layers)
            slab.set_velocity_builtin(problem)
            uh = slab.color()
            uh.zero_()
            lo = sum(slab.layer_counts[:rank])
            slab.owned(uh).copy_(u0.view(n[-1], layer)[lo:lo + slab.layer_counts[rank]])
            torch.cuda.synchronize()
            return slab, uh

        for transport in ("peer", "nccl"):
            slab, uh = fresh(transport)
            slab.begin(0.0, 0.0)
            for _ in range(passes):
                slab.step(uh)
            st = slab.state()
            torch.cuda.synchronize()
            same = torch.equal(slab.owned(uh).contiguous().view(torch.int64), reference_slab(slab).view(torch.int64))
            ok = agree(same and st.time == st1.time and st.dt == st1.dt)
            if not ok:
                failures.append((n, problem, recon, transport))
            say(f"case n={n} problem={problem} recon={recon} [{transport}]: {'bit-identical' if ok else 'MISMATCH'} (time={st.time!r} dt={st.dt!r})")
            # the same passes replayed from a CUDA graph of one captured pass (what bench.py times)
            if passes > 2:
                slab2, u2 = fresh(transport)
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    slab2.grid.use_current_stream()
                    slab2.begin(0.0, 0.0)
                    slab2.step(u2)                                   # first passes eagerly (dense flags, NCCL warm-up)
                    slab2.step(u2)
                torch.cuda.synchronize()
                slab2.capture(u2)
                for _ in range(passes - 2):
                    slab2.replay()
                torch.cuda.synchronize()
                st2 = slab2.state()
                same2 = torch.equal(slab2.owned(u2).contiguous().view(torch.int64), reference_slab(slab2).view(torch.int64))
                ok2 = agree(same2 and st2.time == st1.time and st2.dt == st1.dt)
                if not ok2:
                    failures.append((n, problem, recon, transport, "graph"))
                say(f"   graph replay [{transport}]: {'bit-identical' if ok2 else 'MISMATCH'}")
                slab2.close()
            slab.close()

        # rebalancing between two passes: uneven start, cost-weighted split from the flags, migration, continue
        if recon == ops.RECON_YOUNGS and world == 2:
            first = max(4, n[-1] // 8)
            slab3, u3 = fresh("peer", layers=[first, n[-1] - first])
            slab3.begin(0.0, 0.0)
            half = passes // 2
            for _ in range(half):
                slab3.step(u3)
            slab4, u4 = slab3.rebalanced(u3, min_gain=0.0)
            moved = list(slab4.layer_counts) != list(slab3.layer_counts)
            for _ in range(passes - half):
                slab4.step(u4)
            st4 = slab4.state()
            torch.cuda.synchronize()
            same4 = torch.equal(slab4.owned(u4).contiguous().view(torch.int64), reference_slab(slab4).view(torch.int64))
            ok4 = agree(same4 and st4.time == st1.time and st4.dt == st1.dt and moved)
            if not ok4:
                failures.append((n, problem, recon, "rebalanced"))
            say(f"   rebalanced {slab3.layer_counts} -> {slab4.layer_counts}: {'bit-identical' if ok4 else 'MISMATCH'}")
            if slab4 is not slab3:
                slab4.close()
            slab3.close()

        # host-resident slab: H2D every pass, changed (cell, value) pairs written back (what bench.py's e2e does for N > 1)
        if recon == ops.RECON_YOUNGS:
            import ctypes as C
            from dune_vof_b200.lib import check
            slab5 = SlabAlgorithm(n, rank, world, 0.5, 1e-6, recon, device=local, transport="nccl")
            g5 = slab5.grid
            g5.set_velocity_builtin(problem)
            u5 = g5.empty("color")
            own5 = slab5.owned(u5)
            lo = sum(slab5.layer_counts[:rank])
            host = u0.view(n[-1], layer)[lo:lo + slab5.layer_counts[rank]].contiguous().cpu().pin_memory()
            check(g5._L.vofx_track_changes(g5._ctx, 1))
            slab5.begin(0.0, 0.0)
            nchg = C.c_int64(0)
            for _ in range(passes):
                own5.copy_(host.view_as(own5), non_blocking=True)
                slab5.step(u5)
                check(g5._L.vofx_changed_fetch(g5._ctx, C.c_void_p(host.data_ptr()), C.c_int64(-slab5.exchanger.halo * slab5.layer_cells), C.byref(nchg)))
            same5 = torch.equal(host.view(torch.int64), reference_slab(slab5).cpu().view(torch.int64))
            ok5 = agree(same5)
            if not ok5:
                failures.append((n, problem, recon, "host slab"))
            say(f"   host-resident slab with sparse write-back: {'bit-identical' if ok5 else 'MISMATCH'}")
            slab5.close()
        g1.close()
    dist.barrier()
    torch.cuda.synchronize()
    if failures:
        print("MULTIRANK FAILED", failures, flush=True)
    else:
        say("MULTIRANK PASSED")
    sys.stdout.flush()
    dist.destroy_process_group()
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
