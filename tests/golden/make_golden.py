"""Generates the golden vectors in this directory from the UNMODIFIED reference (oracle/_ref/libvofref.so, i.e.
/root/reference's own templates on the oracle's grid shim).  Run where /root/reference is mounted:

    python tests/golden/make_golden.py

The .npz files are committed; the GPU box has no /root/reference and only reads them.
Inputs are seeded (numpy default_rng) so the script is reproducible; outputs are whatever the reference returns.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reflib as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def primitives(dim, count, seed):
    rng = np.random.default_rng(seed)
    nface = 2 * dim
    lo = np.empty((count, dim)); h = np.empty((count, dim)); nrm = np.empty((count, dim)); frac = np.empty(count)
    hs_loc = np.empty((count, dim + 1)); hs = np.empty((count, dim + 1)); face = np.empty(count, dtype=np.int32)
    v = np.empty((count, dim)); flux = np.empty(count); vf = np.empty(count)
    nv = np.empty(count, dtype=np.int32); cen = np.empty((count, dim)); meas = np.empty(count)
    verts = np.zeros((count, 8, dim))
    for it in range(count):
        res = int(rng.choice([64, 128, 512, 1024, 4096, 100, 37]))
        h[it] = 1.0 / res
        lo[it] = rng.integers(0, res, dim) * h[it]
        n = rng.normal(size=dim)
        mode = it % 10
        if mode == 1:
            n[rng.integers(0, dim)] = 0.0
        if mode == 2:
            n[rng.integers(0, dim)] *= 1e-9
        if mode == 3:
            k = rng.integers(0, dim); n[:] = 0; n[k] = rng.choice([-1.0, 1.0])
        if mode == 4:
            k = rng.integers(0, dim); n[:] = rng.choice([-1.0, 1.0], size=dim); n[k] = 0.0 if dim == 3 else n[k]
        n /= np.linalg.norm(n)
        f = rng.random()
        if mode == 5:
            f = float(rng.choice([0.0, 1.0, 1e-7, 1 - 1e-7, 1e-6, 0.5, 1.2, -0.1]))
        nrm[it] = n; frac[it] = f
        hs_loc[it] = R.locate(lo[it], h[it], n, f)
        hs[it] = hs_loc[it]
        if it % 5 == 0:
            hs[it, dim] += rng.normal() * h[it, 0]
        face[it] = rng.integers(0, nface)
        v[it] = rng.normal(size=dim) * h[it, 0] * 0.5
        if it % 11 == 3:
            v[it, rng.integers(0, dim)] = 0.0
        flux[it] = R.flux(lo[it], h[it], face[it], v[it], hs[it])
        vf[it] = R.volume_fraction(lo[it], h[it], hs[it])
        nv[it], vv, cen[it], meas[it] = R.interface(lo[it], h[it], hs[it])
        verts[it, : len(vv)] = vv
    np.savez_compressed(os.path.join(OUT, f"primitives_{dim}d.npz"), lo=lo, h=h, normal=nrm, fraction=frac,
                        hs_located=hs_loc, hs=hs, face=face, v=v, flux=flux, volume_fraction=vf,
                        iface_nverts=nv, iface_verts=verts, iface_centroid=cen, iface_measure=meas)


def operators(name, n, problem, steps_before=4, eps=1e-6, cfl=0.5, swartz=True):
    g = R.RefGrid(n)
    u0 = g.init(problem)
    u, t, dt, _, _ = g.run(R.RECON_YOUNGS, problem, eps, cfl, steps_before, u0)
    out = dict(n=np.array(n), problem=problem, eps=eps, cfl=cfl, u0=u0, u=u, time=t, dt=dt)
    fl = g.flag(u, eps)
    out["flags"] = fl
    kinds = [(R.RECON_YOUNGS, "youngs"), (R.RECON_HF, "hf")] + ([(R.RECON_SWARTZ, "swartz")] if swartz else [])
    for kind, kname in kinds:
        hs = g.reconstruct(kind, u, fl)
        upd, de = g.evolve(hs, fl, problem, t, dt)
        out[f"hs_{kname}"] = hs
        out[f"update_{kname}"] = upd
        out[f"dtest_{kname}"] = de
        if kind == R.RECON_HF:
            k, v = g.curvature(u, hs, fl)
            out["kappa"] = k
            out["kappa_valid"] = v
        passes = 3 if (kind == R.RECON_SWARTZ and len(n) == 3) else 6
        ur, tr, dr, _, _ = g.run(kind, problem, eps, cfl, passes, u0)
        out[f"run_{kname}_passes"] = passes
        out[f"run_{kname}_u"] = ur
        out[f"run_{kname}_time"] = tr
        out[f"run_{kname}_dt"] = dr
    np.savez_compressed(os.path.join(OUT, f"operators_{name}.npz"), **out)
    g.close()


def swartz_curvature_2d():
    """SwartzCurvature (2D, curvature/swartzcurvature.hh:41-173) of the reference on the reconstructions of the 2D operator cases"""
    out = {}
    for name in ("2d_circle_32", "2d_sflow_40x24", "2d_zalesak_64"):
        g0 = np.load(os.path.join(OUT, f"operators_{name}.npz"))
        grid = R.RefGrid(tuple(int(x) for x in g0["n"]))
        for kname in ("youngs", "hf"):
            out[f"{name}_{kname}"] = grid.curvature_swartz(g0[f"hs_{kname}"], g0["flags"])
        grid.close()
    np.savez_compressed(os.path.join(OUT, "curvature_swartz_2d.npz"), **out)


def c1_run():
    """SURVEY.md section 8(c) known answer: 128^2, Young's, RotatingCircle, eps 1e-6, cfl 0.5, 20 loop passes."""
    g = R.RefGrid((128, 128))
    u0 = g.init(R.PROB_ROTATING_CIRCLE)
    u, t, dt, _, mixed = g.run(R.RECON_YOUNGS, R.PROB_ROTATING_CIRCLE, 1e-6, 0.5, 20, u0)
    np.savez_compressed(os.path.join(OUT, "run_c1_128.npz"), u0=u0, u=u, time=t, dt=dt, mixed_cell_steps=mixed)


if __name__ == "__main__":
    primitives(3, 3000, 11)
    primitives(2, 3000, 12)
    operators("2d_circle_32", (32, 32), R.PROB_ROTATING_CIRCLE)
    operators("2d_sflow_40x24", (40, 24), R.PROB_SFLOW)
    operators("2d_zalesak_64", (64, 64), R.PROB_SLOTTED_CYLINDER)
    operators("2d_leveque_32", (32, 32), R.PROB_LEVEQUE)
    operators("3d_sphere_16", (16, 16, 16), R.PROB_ROTATING_CIRCLE)
    operators("3d_sflow_20x12x16", (20, 12, 16), R.PROB_SFLOW)
    operators("3d_leveque_16", (16, 16, 16), R.PROB_LEVEQUE)
    operators("3d_leveque_32", (32, 32, 32), R.PROB_LEVEQUE, swartz=False)
    c1_run()
    swartz_curvature_2d()
    print("golden vectors written to", OUT)
