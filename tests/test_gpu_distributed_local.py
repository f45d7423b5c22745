"""The distributed step over peer-mapped windows (vofx_comm_*, include/vofx.h) with all slabs in one process
(tests/local_slabs.py): bit-identical to the single-context run -- every colour value, time and dt -- for all three
reconstructions, 2D and 3D, uneven slabs, an interface that crosses the slab boundaries, eagerly and replayed from one CUDA
graph per slab.  Runs on one GPU (all contexts on the same device); tests/test_gpu_multirank.py covers the IPC variant."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vofx():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    import dune_vof_b200.operators as ops
    return ops


CASES = [((48, 40, 64), 4, 0, 2, None, 12, False), ((48, 40, 64), 4, 0, 4, [10, 22, 14, 18], 12, False),
         ((32, 32, 48), 1, 1, 3, None, 6, False), ((64, 96), 2, 1, 2, None, 8, True), ((96, 128), 4, 0, 3, [30, 58, 40], 10, False),
         ((24, 24, 32), 0, 2, 2, None, 3, False), ((40, 32), 4, 2, 2, None, 4, False)]


@pytest.mark.parametrize("n,problem,recon,world,layers,passes,curv", CASES)
def test_local_slabs_match_serial(vofx, n, problem, recon, world, layers, passes, curv):
    import torch
    from local_slabs import LocalSlabs
    g1 = vofx.CartesianGrid(n)
    g1.set_velocity_builtin(problem)
    u0 = g1.init(problem)
    ref = vofx.Algorithm(g1, 0.5, 1e-6, reconstruction_kind=recon, with_curvature=curv)
    u1 = u0.clone()
    st = ref.run_passes(u1, passes)

    slabs = LocalSlabs(n, world, 0.5, 1e-6, recon, problem, with_curvature=curv, layers=layers)
    slabs.load(u0)
    slabs.begin(0.0, 0.0)
    for _ in range(passes):
        slabs.step()
    ud = slabs.gather()
    assert torch.equal(ud.view(torch.int64), u1.view(torch.int64))
    for t, dt, _, _ in slabs.states():
        assert t == st.time and dt == st.dt
    assert sum(s[3] for s in slabs.states()) >= st.mixed_cells        # the halo layer's mixed cells are reconstructed twice

    if True:                            # every reconstruction is capturable (the batched 3D Swartz pass sizes its buffers in the eager passes)
        slabs.load(u0)
        slabs.begin(0.0, 0.0)
        slabs.step()
        slabs.step()
        slabs.capture()
        for _ in range(passes - 2):
            slabs.replay()
        ug = slabs.gather()
        assert torch.equal(ug.view(torch.int64), u1.view(torch.int64))
        for t, dt, _, _ in slabs.states():
            assert t == st.time and dt == st.dt
    slabs.close()
    g1.close()


def test_local_slabs_256_long_run(vofx):
    """256^3 LeVeque sphere on 4 uneven slabs over 40 passes (the band crosses two slab boundaries and moves across them)"""
    import torch
    from local_slabs import LocalSlabs
    n = (256, 256, 256)
    g1 = vofx.CartesianGrid(n)
    g1.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = g1.init(vofx.PROB_LEVEQUE)
    u1 = u0.clone()
    st = vofx.Algorithm(g1, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS, incremental=True).run_passes(u1, 40)
    slabs = LocalSlabs(n, 4, 0.5, 1e-6, vofx.RECON_YOUNGS, vofx.PROB_LEVEQUE, layers=[70, 20, 30, 136])
    slabs.load(u0)
    slabs.begin(0.0, 0.0)
    for _ in range(40):
        slabs.step()
    assert torch.equal(slabs.gather().view(torch.int64), u1.view(torch.int64))
    assert all(s[0] == st.time and s[1] == st.dt for s in slabs.states())
    slabs.close()
    g1.close()
