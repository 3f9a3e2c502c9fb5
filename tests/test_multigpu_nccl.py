"""N>1 on real GPUs: skipped unless the box has >= 2 CUDA devices.  Every rank holds its sub-mesh with ghost layers;
the ghost refresh is the LIBRARY's (pack kernel + grouped ncclSend/ncclRecv + scatter kernel on the solver's stream,
bootstrapped with a 128-byte NCCL id) -- torch.distributed only broadcasts that id.  The owned cells of all ranks must
reproduce the single-GPU run of the global mesh (<= 1e-12, in practice bitwise): on a hex box through the graph-replayed
step, and on an RCB-split Kelvin-cell (polyDualMesh-like) mesh."""
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


def _setup(kind, n):
    from common import LEVEQUE_CONTROLS, fields, meshmod
    if kind == "hex":
        return meshmod.hex_block(n), dict(LEVEQUE_CONTROLS), fields.leveque_velocity, 0.25 / n
    m = meshmod.kelvin_mesh(n)
    return m, dict(LEVEQUE_CONTROLS), fields.rotation_velocity, 0.25 * np.cbrt(m.meta["cell_volume"]) / 3.2


def _alpha0(kind, m, C_, V):
    from common import fields
    if kind == "hex":
        return fields.sphere_alpha_quadrature(m)
    return np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.62, 0.5]), axis=1) - 0.2) / np.cbrt(V), 0.0, 1.0)


def _run(s, step, kind, m_global, steps, velocity, dt, nIF, a0):
    from common import capi, fields
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    s.setAlpha(a0)
    s.setPhi(fields.face_flux(Cf, Sf, velocity))
    s.setU(velocity(C_), velocity(Cf[nIF:]))
    for _ in range(steps):
        step(dt)
    s.synchronize()


def _worker(rank, world, kind, n, steps, port, out_dir):
    import torch
    import torch.distributed as dist
    from common import capi
    from geometricvofext_b200 import multigpu as mg
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    m, controls, velocity, dt = _setup(kind, n)
    layers = mg.default_layers(controls)
    if kind == "hex":
        sub, maps = mg.BoxDecomposition(n, world, layers).rank_mesh(rank)
    else:
        sub, maps = mg.decompose(m, mg.partition_rcb(m, world), rank, layers)
    ds = mg.DecomposedSolveVofEqu(sub, maps, controls, rank, world, device=rank)
    s = ds.s
    if kind == "hex":
        a0 = _alpha0(kind, m, None, None)[maps["cell_global"]]
    else:
        a0 = _alpha0(kind, None, s.field(capi.F_C), s.field(capi.F_V))
    _run(s, ds.step, kind, m, steps, velocity, dt, sub.n_internal_faces, a0)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), gid=ds.owned_global_ids(), alpha=ds.alpha_owned(), v1=ds.volume(),
             err=s.info(capi.I_ERROR_FLAGS), halo=ds.halo_bytes)
    dist.barrier()
    ds.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 CUDA devices")
@pytest.mark.parametrize("kind,n,steps", [("hex", 32, 8), ("kelvin", 10, 6)])
def test_two_gpus_match_single_gpu(tmp_path, kind, n, steps):
    import torch.multiprocessing as mp
    from common import SolveVofEqu, capi
    world = 2
    mp.spawn(_worker, args=(world, kind, n, steps, 29700 + os.getpid() % 200 + (11 if kind != "hex" else 0), str(tmp_path)),
             nprocs=world, join=True)
    m, controls, velocity, dt = _setup(kind, n)
    s = SolveVofEqu(m, controls)
    a0 = _alpha0(kind, m, s.field(capi.F_C), s.field(capi.F_V))
    _run(s, s.step, kind, m, steps, velocity, dt, m.n_internal_faces, a0)
    ref, vref = s.alpha(), s.volume()
    got = np.full(m.n_cells, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got[d["gid"]] = d["alpha"]
        assert d["err"] == 0 and d["halo"] > 0
        assert abs(d["v1"] - vref) <= 1e-13 * abs(vref)
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-12


def _case_worker(rank, world, case_dir, port):
    import torch
    import torch.distributed as dist
    from geometricvofext_b200 import foamcase
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    out = foamcase.run_plic_vof_advection_decomposed(case_dir, rank, world, device=rank)
    np.savez(os.path.join(case_dir, "run_rank%d.npz" % rank), steps=out["steps"], written=np.array(out["written"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 CUDA devices")
def test_two_gpus_run_a_decomposed_case_directory(tmp_path):
    """processor*/ directories (cellProcAddressing, 0/alpha.water) on two GPUs with the library's NCCL ghost refresh: the
    reconstructed field equals the single-GPU run of the same case directory."""
    import torch.multiprocessing as mp
    from common import fields
    from geometricvofext_b200 import foamcase, foamfile
    from test_foam_formats import make_case
    case = make_case(str(tmp_path / "c"), n=24, end=0.01, wi=0.005)
    mesh = case.mesh()
    a0 = fields.sphere_alpha_quadrature(mesh)
    case.write_alpha(mesh, "0", a0)
    foamfile.write_polymesh(mesh, case.dir)
    foamcase.decompose_case(case, 2, weights=1.0 + 50.0 * ((a0 > 0) & (a0 < 1)))
    serial = foamcase.run_plic_vof_advection(case)
    mp.spawn(_case_worker, args=(2, case.dir, 29950 + os.getpid() % 40), nprocs=2, join=True)
    for r in range(2):
        d = np.load(os.path.join(case.dir, "run_rank%d.npz" % r))
        assert int(d["steps"]) == serial["steps"] and d["written"].tolist() == serial["written"]
    for name in serial["written"]:
        got = foamcase.reconstruct_par(case, name)
        ref = foamfile.read_field(os.path.join(case.dir, name, case.alpha_name)).internal_array(mesh.n_cells)
        assert np.abs(got - ref).max() <= 1e-12
