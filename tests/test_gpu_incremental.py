"""Incremental flagging (vofx_set_incremental_flagging, include/vofx.h): after every loop pass the flags, the band list
(as a set) and the colour function must be bit-identical to the dense flag kernel's -- 3D Young's at 256^3 over 60 passes
(the band moves through several cell layers), 2D height functions with curvature, eagerly and replayed from a CUDA
graph, and through the context-resident loop of the C ABI (vofx_step_resident)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vofx():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    import dune_vof_b200.operators as ops
    return ops


def _compare_every_pass(vofx, n, problem, kind, passes, with_curvature=False):
    import torch
    ga, gb = vofx.CartesianGrid(n), vofx.CartesianGrid(n)
    for g in (ga, gb):
        g.set_velocity_builtin(problem)
    u0 = ga.init(problem)
    dense = vofx.Algorithm(ga, 0.5, 1e-6, reconstruction_kind=kind, with_curvature=with_curvature)
    incr = vofx.Algorithm(gb, 0.5, 1e-6, reconstruction_kind=kind, with_curvature=with_curvature, incremental=True)
    ua, ub = u0.clone(), u0.clone()
    dense.begin(0.0, 0.0)
    incr.begin(0.0, 0.0)
    grew = 0
    for k in range(passes):
        dense.step(ua)
        incr.step(ub)
        assert torch.equal(dense.flags, incr.flags), f"flags differ after pass {k}"
        la, lb = np.sort(dense.band_list()), np.sort(incr.band_list())
        assert np.array_equal(la, lb), f"band lists differ after pass {k}"
        assert len(np.unique(lb)) == len(lb), f"duplicate band entries after pass {k}"
        assert torch.equal(ua.view(torch.int64), ub.view(torch.int64)), f"colour functions differ after pass {k}"
        if with_curvature:
            assert torch.equal(dense.curvature.view(torch.int64), incr.curvature.view(torch.int64))
        sa, sb = dense.state(), incr.state()
        assert sa.time == sb.time and sa.dt == sb.dt and sa.mixed_cells == sb.mixed_cells == len(lb)
        grew = max(grew, len(lb))
    ga.close()
    gb.close()
    return grew


def test_incremental_3d_youngs_256(vofx):
    assert _compare_every_pass(vofx, (256, 256, 256), vofx.PROB_LEVEQUE, vofx.RECON_YOUNGS, 60) > 25000


@pytest.mark.parametrize("n,problem,kind,curv", [((512, 512), 2, 1, True), ((96, 80, 64), 1, 1, False), ((200, 160), 4, 0, False),
                                                  ((24, 20, 16), 3, 0, False), ((37, 29), 4, 2, False)])
def test_incremental_other_configs(vofx, n, problem, kind, curv):
    """2D height functions with curvature (C2's operators), 3D height functions, rows that are not a multiple of 4 (the
    generic dense kernel), an axis-aligned wall touching the domain boundary"""
    _compare_every_pass(vofx, n, problem, kind, 25, with_curvature=curv)


def test_incremental_graph_replay_and_restart(vofx):
    """a captured pass replays the incremental kernels; vofx_step_begin makes the next pass dense again"""
    import torch
    n = (96, 96, 96)
    ga, gb = vofx.CartesianGrid(n), vofx.CartesianGrid(n)
    for g in (ga, gb):
        g.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = ga.init(vofx.PROB_LEVEQUE)
    dense = vofx.Algorithm(ga, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS)
    incr = vofx.Algorithm(gb, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS, incremental=True)
    ua, ub = u0.clone(), u0.clone()
    sa = dense.run_passes(ua, 30)
    incr.begin(0.0, 0.0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gb.use_current_stream()
        incr.step(ub)
        incr.step(ub)
    torch.cuda.synchronize()
    before = gb.launch_count
    incr.capture(ub)
    captured = gb.launch_count - before
    for _ in range(28):
        incr.replay()
    torch.cuda.synchronize()
    sb = incr.state()
    assert torch.equal(ua.view(torch.int64), ub.view(torch.int64)) and sa.time == sb.time and sa.dt == sb.dt
    assert torch.equal(dense.flags, incr.flags)
    # restart on a different colour function: the first pass must not trust the old flags
    ub2 = u0.clone()
    gb.use_current_stream()
    sb2 = incr.run_passes(ub2, 30)
    assert torch.equal(ua.view(torch.int64), ub2.view(torch.int64)) and sb2.time == sa.time
    assert captured >= 6
    ga.close()
    gb.close()


def test_resident_loop_is_incremental_and_exact(vofx):
    """vofx_color_upload / vofx_step_resident / vofx_color_download (what the C++ adaptor's Algorithm uses) against the
    plain loop on device arrays"""
    import torch
    from dune_vof_b200 import lib as vlib
    L = vlib.load()
    n = (64, 64, 64)
    ga, gb = vofx.CartesianGrid(n), vofx.CartesianGrid(n)
    for g in (ga, gb):
        g.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = ga.init(vofx.PROB_LEVEQUE)
    ua = u0.clone()
    sa = vofx.Algorithm(ga, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS).run_passes(ua, 20)
    host = u0.cpu().numpy().copy()
    p = vlib.StepParams(0, 10, 1e-6, 0.5, 0)
    vlib.check(L.vofx_step_begin(gb._ctx, 0.0, 0.0))
    vlib.check(L.vofx_color_upload(gb._ctx, C.c_void_p(host.ctypes.data)))
    before = gb.launch_count
    for _ in range(20):
        vlib.check(L.vofx_step_resident(gb._ctx, C.byref(p)))
    flags = np.empty(host.size, dtype=np.uint8)
    vlib.check(L.vofx_color_download(gb._ctx, C.c_void_p(host.ctypes.data), C.c_void_p(flags.ctypes.data)))
    assert np.array_equal(host.view(np.uint64), ua.cpu().numpy().view(np.uint64))
    t, dt = C.c_double(), C.c_double()
    vlib.check(L.vofx_step_state(gb._ctx, C.byref(t), C.byref(dt), None, None))
    assert t.value == sa.time and dt.value == sa.dt
    ga.close()
    gb.close()
