"""N>1 path on CPU: world_size-2 (and 4) gloo runs of the decomposed driver
(geometricvofext_b200/multigpu.py) with the CPU oracle as the per-rank engine, checked against
the single-domain oracle.  Covers the host-side logic of the multi-GPU leg: box decomposition,
halo index plans, the per-step alpha exchange and the owned-cell reductions."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, n, steps, port, out_dir):
    import torch.distributed as dist
    from common import LEVEQUE_CONTROLS, capi, fields, oracle_lib
    from geometricvofext_b200.multigpu import DecomposedSolveVofEqu
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ds = DecomposedSolveVofEqu(n, LEVEQUE_CONTROLS, rank, world, lib=oracle_lib())
    s = ds.s
    a0 = fields.sphere_alpha_quadrature(ds.mesh)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C_), fields.face_flux(Cf, Sf)
    ds.setAlpha(a0)
    ds.exchange_alpha()
    dt = 0.25 / n
    v0 = ds.volume()
    for k in range(steps):
        ds.setPhi(phi0)
        ds.setU(U0)
        if k % 2:
            ds.step(dt)          # reconstruct + advect + halo swap in one call (the graph-replayed call on the device)
        else:
            ds.reconstruct()
            ds.advect(dt)
    v1 = ds.volume()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), gid=ds.owned_global_ids(), alpha=ds.alpha_owned(), v0=v0, v1=v1)
    dist.barrier()
    dist.destroy_process_group()


def _single(n, steps):
    from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, fields, meshmod, oracle_lib
    m = meshmod.hex_block(n)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    a0 = fields.sphere_alpha_quadrature(m)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C_), fields.face_flux(Cf, Sf)
    s.setAlpha(a0)
    dt = 0.25 / n
    for k in range(steps):
        s.setPhi(phi0)
        s.setU(U0)
        s.reconstruct()
        s.advect(dt)
    return s.alpha(), s.volume()


def test_decomposition_plans_are_consistent():
    from geometricvofext_b200.multigpu import Decomposition, block_grid
    assert block_grid(1) == (1, 1, 1) and block_grid(2) == (1, 1, 2) and block_grid(4) == (1, 2, 2) and block_grid(8) == (2, 2, 2)
    for world in (2, 4, 8):
        dec = Decomposition(24, world, halo=3)
        seen = np.zeros(24 ** 3, dtype=int)
        plans = [dec.plan(r) for r in range(world)]
        meshes = [dec.rank_mesh(r) for r in range(world)]
        for r in range(world):
            seen[meshes[r].cell_global[plans[r]["owned"]]] += 1
            for q, idx in plans[r]["recv"].items():
                # what r receives from q is exactly what q sends to r, in the same global order
                g_recv = meshes[r].cell_global[idx]
                g_send = meshes[q].cell_global[plans[q]["send"][r]]
                assert np.array_equal(g_recv, g_send)
        assert np.all(seen == 1), "every global cell is owned by exactly one rank"


@pytest.mark.parametrize("world", [2, 4])
def test_two_rank_gloo_matches_single_domain(tmp_path, world):
    import torch.multiprocessing as mp
    n, steps = 20, 6
    port = 29500 + (os.getpid() % 500) + world
    mp.spawn(_worker, args=(world, n, steps, port, str(tmp_path)), nprocs=world, join=True)
    ref, vref = _single(n, steps)
    got = np.full(n ** 3, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got[d["gid"]] = d["alpha"]
        assert abs(d["v1"] - vref) <= 1e-13 * abs(vref)
        assert abs(d["v1"] - d["v0"]) <= 1e-13 * abs(d["v0"]), "volume conserved across ranks"
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-12, "decomposed run differs from single domain by %g" % np.abs(got - ref).max()
