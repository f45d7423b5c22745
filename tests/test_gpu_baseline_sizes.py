"""Bit-exact parity with the plain-C oracle at the sizes BASELINE.json quotes its configs on (the oracle is pinned to the
unmodified reference by tests/test_oracle_vs_reference.py):
  C3  512^3 sphere in the LeVeque field, Young's -- the bench workload itself
  C2  Zalesak disk at 1024^2 and 4096^2, height functions + curvature every pass
  C4  3D modified Swartz at 64^3, 128^3 and 256^3, batched and per-cell formulations
Every colour value, time and dt must carry the oracle's bits (tolerance 0; curvature uses pow(): 1e-12 relative)."""
import os

import numpy as np
import pytest

from util import n_mismatch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vofx():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    import dune_vof_b200.operators as ops
    return ops


@pytest.mark.parametrize("incremental", [False, True])
def test_c3_512_youngs_against_oracle(vofx, incremental):
    """configs[2] exactly: 512^3, LeVeque sphere, Young's, eps 1e-6, cfl 0.5; 3 loop passes (the first is the reference's
    dry run with dt = 0, so two real advections) -- about 12 s of host time for the oracle"""
    import torch
    from oracle import oraclelib as O
    n = (512, 512, 512)
    grid = vofx.CartesianGrid(n)
    grid.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = grid.init(vofx.PROB_LEVEQUE)
    og = O.OracleGrid(n)
    uo, to, do, _, mixed = og.run(O.RECON_YOUNGS, O.PROB_LEVEQUE, 1e-6, 0.5, 3, u0.cpu().numpy())
    uh = u0.clone()
    alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_YOUNGS, incremental=incremental)
    st = alg.run_passes(uh, 3)
    assert n_mismatch(uh.cpu().numpy(), uo) == 0
    assert st.time == to and st.dt == do
    assert st.mixed_cells > 100000 and mixed > 300000
    grid.close()
    og.close()


@pytest.mark.parametrize("n", [1024, 4096])
def test_c2_hf_curvature_against_oracle(vofx, n):
    """configs[1]: Zalesak disk (SlottedCylinder), height-function reconstruction + height-function curvature in every
    pass, 3 passes.  The colour function, time and dt are compared after the loop; the reconstructions, the curvature and
    its validity flags of the last pass against the oracle's operators on the state before that pass."""
    import torch
    from oracle import oraclelib as O
    nn = (n, n)
    grid = vofx.CartesianGrid(nn)
    grid.set_velocity_builtin(vofx.PROB_SLOTTED_CYLINDER)
    u0 = grid.init(vofx.PROB_SLOTTED_CYLINDER)
    og = O.OracleGrid(nn)
    u0h = u0.cpu().numpy()
    assert n_mismatch(u0h, og.init(O.PROB_SLOTTED_CYLINDER)) == 0
    for incremental in (False, True):
        alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_HF, with_curvature=True, incremental=incremental)
        uh = u0.clone()
        st = alg.run_passes(uh, 3)
        u2, t2, d2, _, _ = og.run(O.RECON_HF, O.PROB_SLOTTED_CYLINDER, 1e-6, 0.5, 2, u0h)
        u3, t3, d3, _, _ = og.run(O.RECON_HF, O.PROB_SLOTTED_CYLINDER, 1e-6, 0.5, 1, u2, t2, d2)
        assert n_mismatch(uh.cpu().numpy(), u3) == 0
        assert st.time == t3 and st.dt == d3
        fo = og.flag(u2, 1e-6)
        assert np.array_equal(alg.flags.cpu().numpy(), fo)
        ho = og.reconstruct(O.RECON_HF, u2, fo)
        mixed = (fo == 1) | (fo == 2)
        assert mixed.sum() > 3 * n
        assert n_mismatch(alg.reconstructions.cpu().numpy()[mixed], ho[mixed]) == 0
        ko, vo = og.curvature(u2, ho, fo)
        kd = alg.curvature.cpu().numpy()
        assert np.count_nonzero(ko) > n
        assert np.array_equal(kd != 0.0, ko != 0.0)
        assert np.allclose(kd, ko, rtol=1e-12, atol=0.0)        # pow(): CUDA's and glibc's may differ in the last bit
    # validity flags through the operator-level call on the same state
    flags, recs, kappa = grid.empty("flags"), grid.empty("reconstructions"), grid.empty("curvature")
    ut = torch.tensor(u2, device=grid.device)
    vofx.FlagOperator(grid, 1e-6)(ut, flags)
    vofx.HeightFunctionReconstruction(grid)(ut, recs, flags)
    op = vofx.CartesianHeightFunctionCurvature(grid)
    op(ut, recs, flags, kappa)
    assert np.array_equal(op.satisfies_constraint.cpu().numpy(), vo)
    assert np.allclose(kappa.cpu().numpy(), ko, rtol=1e-12, atol=0.0)
    grid.close()
    og.close()


@pytest.mark.parametrize("n,mode", [(64, "batched"), (64, "cell"), (128, "batched"), (128, "cell"), (256, "batched")])
def test_c4_swartz3d_against_oracle(vofx, n, mode, monkeypatch):
    """configs[3]: 3D ModifiedSwartzReconstruction (maxIterations 10) on the LeVeque sphere; 256^3 is the configured size
    (3.1e4 mixed cells per pass, about half a minute of host time for the oracle's two passes)"""
    import torch
    from oracle import oraclelib as O
    monkeypatch.setenv("VOFX_SWARTZ_MODE", mode)
    nn = (n, n, n)
    grid = vofx.CartesianGrid(nn)
    grid.set_velocity_builtin(vofx.PROB_LEVEQUE)
    u0 = grid.init(vofx.PROB_LEVEQUE)
    og = O.OracleGrid(nn)
    uo, to, do, _, mixed = og.run(O.RECON_SWARTZ, O.PROB_LEVEQUE, 1e-6, 0.5, 2, u0.cpu().numpy())
    uh = u0.clone()
    alg = vofx.Algorithm(grid, 0.5, 1e-6, reconstruction_kind=vofx.RECON_SWARTZ)
    st = alg.run_passes(uh, 2)
    assert n_mismatch(uh.cpu().numpy(), uo) == 0
    assert st.time == to and st.dt == do
    # the reconstructions of the last pass against the operator on the state before it
    u1, t1, d1, _, _ = og.run(O.RECON_SWARTZ, O.PROB_LEVEQUE, 1e-6, 0.5, 1, u0.cpu().numpy())
    fo = og.flag(u1, 1e-6)
    if n <= 128:
        ho = og.reconstruct(O.RECON_SWARTZ, u1, fo)
        m = (fo == 1) | (fo == 2)
        assert n_mismatch(alg.reconstructions.cpu().numpy()[m], ho[m]) == 0
    assert mixed > 2 * n * n / 8
    grid.close()
    og.close()
