"""N>1 on real GPUs (NCCL): skipped unless the box has >= 2 CUDA devices.  Two ranks advance the tiled LeVeque
problem with the stream-ordered halo exchange; every rank's owned cells must reproduce the single-GPU run of the
global mesh (<= 1e-12, in practice bitwise) and the total volume must be conserved."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, n, steps, port, out_dir):
    import torch
    import torch.distributed as dist
    from common import LEVEQUE_CONTROLS, capi, fields
    from geometricvofext_b200.multigpu import DecomposedSolveVofEqu, block_grid
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    grid = np.array(block_grid(world))
    ds = DecomposedSolveVofEqu((n * grid).tolist(), LEVEQUE_CONTROLS, rank, world, device=rank, length=grid.astype(float).tolist())
    s = ds.s
    a0 = sum(fields.sphere_alpha_quadrature(ds.mesh, centre=(0.35 + i, 0.35 + j, 0.35 + k))
             for i in range(grid[0]) for j in range(grid[1]) for k in range(grid[2]))
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    s.setAlpha(a0)
    s.setPhi(fields.face_flux(Cf, Sf))
    s.setU(fields.leveque_velocity(C_), np.zeros((s.nBF, 3)))
    ds.exchange_alpha()
    v0 = ds.volume()
    for k in range(steps):
        ds.reconstruct()
        ds.advect(0.25 / n)
    s.synchronize()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), gid=ds.owned_global_ids(), alpha=ds.alpha_owned(), v0=v0, v1=ds.volume())
    dist.barrier()
    ds.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 CUDA devices")
def test_two_gpus_match_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, fields, meshmod
    from geometricvofext_b200.multigpu import block_grid
    n, steps, world = 32, 8, 2
    mp.spawn(_worker, args=(world, n, steps, 29700 + os.getpid() % 200, str(tmp_path)), nprocs=world, join=True)
    grid = np.array(block_grid(world))
    m = meshmod.hex_block((n * grid).tolist(), length=grid.astype(float).tolist())
    s = SolveVofEqu(m, LEVEQUE_CONTROLS)
    a0 = sum(fields.sphere_alpha_quadrature(m, centre=(0.35 + i, 0.35 + j, 0.35 + k))
             for i in range(grid[0]) for j in range(grid[1]) for k in range(grid[2]))
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    s.setAlpha(a0)
    s.setPhi(fields.face_flux(Cf, Sf))
    s.setU(fields.leveque_velocity(C_))
    for k in range(steps):
        s.reconstruct()
        s.advect(0.25 / n)
    ref = s.alpha()
    got = np.full(m.n_cells, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got[d["gid"]] = d["alpha"]
        assert abs(d["v1"] - d["v0"]) <= 1e-13 * abs(d["v0"])
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-12
