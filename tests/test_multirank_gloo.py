"""N>1 path on CPU: world_size 2..4 gloo runs of the decomposed driver (geometricvofext_b200/multigpu.py) with
the CPU oracle as the per-rank engine, checked against the single-domain oracle.  Covers the host-side logic of the
multi-GPU leg: the library's partitioner and sub-domain extraction (svof_partition_rcb / svof_decompose, called here
without a GPU), the box fast path, the ghost plan, the per-step ghost refresh and the owned-cell reductions -- on a
hex box and on an RCB-split polyhedral (Kelvin-cell) mesh."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _case(kind, n):
    """mesh, controls, alpha0(C, V), velocity, dt -- the same on every rank and in the single-domain run"""
    from common import LEVEQUE_CONTROLS, fields, meshmod
    if kind in ("hex", "hexI"):   # hexI: the interface-aware boxes of the strong-scaling bench (multigpu.interface_boxes)
        m = meshmod.hex_block(n)
        return m, dict(LEVEQUE_CONTROLS), None, fields.leveque_velocity, 0.25 / n
    if kind == "hex10":   # default nAlphaBounds 10, Courant ~1: several effective bounding sweeps near the cuts
        m = meshmod.hex_block(n)
        return m, dict(LEVEQUE_CONTROLS, nAlphaBounds=10), None, fields.leveque_velocity, 0.5 / n
    m = meshmod.kelvin_mesh(n)
    return m, dict(LEVEQUE_CONTROLS), "smeared", fields.rotation_velocity, None


def _alpha0(kind, mesh_like, C_, V):
    from common import fields
    if kind.startswith("hex"):
        return fields.sphere_alpha_quadrature(mesh_like)
    h = np.cbrt(V)
    return np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.62, 0.5]), axis=1) - 0.2) / h, 0.0, 1.0)


def _fields(s, velocity, nIF):
    from common import capi, fields
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    return velocity(C_), fields.face_flux(Cf, Sf, velocity), velocity(Cf[nIF:])


def _worker(rank, world, kind, n, steps, port, out_dir, layers):
    import torch.distributed as dist
    from common import capi, oracle_lib
    from geometricvofext_b200 import multigpu as mg
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    m, controls, _, velocity, dt = _case(kind, n)
    layers = layers or mg.default_layers(controls)
    custom = mg.interface_boxes(n, world) if kind == "hexI" else None
    assert kind != "hexI" or custom is not None
    if kind.startswith("hex") and rank % 2 == 0:
        sub, maps = mg.BoxDecomposition(n, world, layers, boxes=custom).rank_mesh(rank)      # box fast path (no global mesh)
    elif kind.startswith("hex"):
        boxes = mg.BoxDecomposition(n, world, layers, boxes=custom).boxes                     # the same boxes through svof_decompose
        cr = np.empty(n ** 3, np.int32)
        for r, (lo, hi) in enumerate(boxes):
            k, j, i = np.meshgrid(np.arange(lo[2], hi[2]), np.arange(lo[1], hi[1]), np.arange(lo[0], hi[0]), indexing="ij")
            cr[(i + n * (j + n * k)).reshape(-1)] = r
        sub, maps = mg.decompose(m, cr, rank, layers)
    else:
        sub, maps = mg.decompose(m, mg.partition_rcb(m, world), rank, layers)
    ds = mg.DecomposedSolveVofEqu(sub, maps, controls, rank, world, lib=oracle_lib())
    s = ds.s
    if kind.startswith("hex"):
        sub.meta.update({k: v for k, v in m.meta.items() if k in ("N", "length", "origin", "h")})
        if "lo" not in sub.meta or sub.meta.get("kind") == "submesh":
            a0 = _alpha0(kind, m, None, None)[maps["cell_global"]]              # restrict the global field
        else:
            a0 = _alpha0(kind, sub, None, None)
    else:
        a0 = _alpha0(kind, None, s.field(capi.F_C), s.field(capi.F_V))
    U0, phi0, Ub = _fields(s, velocity, sub.n_internal_faces)
    if dt is None:
        dt = 0.25 * np.cbrt(m.meta["cell_volume"]) / 3.2
    ds.setAlpha(a0)
    ds.exchange_alpha()
    v0 = ds.volume()
    for k in range(steps):
        ds.setPhi(phi0)
        ds.setU(U0, Ub)
        if k % 2:
            ds.step(dt)          # reconstruct + advect + ghost refresh in one call (the graph-replayed call on the device)
        else:
            ds.reconstruct()
            ds.advect(dt)
    v1 = ds.volume()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), gid=ds.owned_global_ids(), alpha=ds.alpha_owned(), v0=v0, v1=v1,
             sweeps=s.info(capi.I_N_BOUND_SWEEPS))
    dist.barrier()
    dist.destroy_process_group()


def _single(kind, n, steps):
    from common import SolveVofEqu, capi, oracle_lib
    m, controls, _, velocity, dt = _case(kind, n)
    s = SolveVofEqu(m, controls, lib=oracle_lib())
    a0 = _alpha0(kind, m, s.field(capi.F_C), s.field(capi.F_V))
    U0, phi0, Ub = _fields(s, velocity, m.n_internal_faces)
    if dt is None:
        dt = 0.25 * np.cbrt(m.meta["cell_volume"]) / 3.2
    s.setAlpha(a0)
    max_sweeps = 0
    for k in range(steps):
        s.setPhi(phi0)
        s.setU(U0, Ub)
        s.reconstruct()
        s.advect(dt)
        max_sweeps = max(max_sweeps, int(s.info(capi.I_N_BOUND_SWEEPS)))
    return s.alpha(), s.volume(), m.n_cells, max_sweeps


def test_partition_and_decomposition_are_consistent():
    """library partitioner + sub-domain extraction, no GPU: every cell owned once, ghost plans mirror each other,
    local numbering monotone in the global one, the cut faces closed by one extra patch"""
    from common import meshmod
    from geometricvofext_b200 import capi, multigpu as mg
    assert mg.block_grid(1) == (1, 1, 1) and mg.block_grid(2) == (1, 1, 2) and mg.block_grid(4) == (1, 2, 2) and mg.block_grid(8) == (2, 2, 2)
    for m, world in ((meshmod.hex_block(12), 4), (meshmod.kelvin_mesh(4), 3), (meshmod.prism_mesh(6), 2)):
        w = np.ones(m.n_cells)
        w[: m.n_cells // 4] = 5.0
        cr = mg.partition_rcb(m, world, w)
        loads = np.bincount(cr, weights=w, minlength=world)
        assert loads.min() > 0 and loads.max() / loads.mean() < 1.25, "weighted parts are balanced: %s" % loads
        seen = np.zeros(m.n_cells, int)
        subs = [mg.decompose(m, cr, r, 2) for r in range(world)]
        for r, (sub, maps) in enumerate(subs):
            g = maps["cell_global"]
            assert np.all(np.diff(g) > 0) and np.all(np.diff(maps["point_global"]) > 0)
            nIF = sub.n_internal_faces
            assert np.all(np.diff(maps["face_global"][:nIF]) > 0), "internal faces keep their global order"
            assert np.all(sub.owner[:nIF] < sub.neighbour), "upper-triangular order survives the renumbering"
            assert np.array_equal(maps["cell_owner_rank"], cr[g])
            seen[g[maps["owned_local"]]] += 1
            assert sub.patches[-1].name == "cut" and sub.patches[-1].kind == capi.PATCH_GENERIC
            assert sum(p.size for p in sub.patches) == sub.n_faces - nIF
            # layer-1 ghosts are exactly the non-owned cells sharing a point with an owned cell
            owned_pts = np.zeros(m.n_points, bool)
            own_g = set(g[maps["owned_local"]].tolist())
            for f in range(m.n_faces):
                cells_f = [m.owner[f]] + ([m.neighbour[f]] if f < m.n_internal_faces else [])
                if any(c in own_g for c in cells_f):
                    owned_pts[m.face_points[m.face_offsets[f]:m.face_offsets[f + 1]]] = True
            ring = set()
            for f in range(m.n_faces):
                if owned_pts[m.face_points[m.face_offsets[f]:m.face_offsets[f + 1]]].any():
                    ring.add(int(m.owner[f]))
                    if f < m.n_internal_faces:
                        ring.add(int(m.neighbour[f]))
            assert set(g[maps["cell_layer"] == 1].tolist()) == ring - own_g
        assert np.all(seen == 1), "every global cell is owned by exactly one rank"


@pytest.mark.parametrize("kind,world,n,steps,layers", [("hex", 2, 20, 6, 0), ("hex", 4, 20, 6, 0), ("kelvin", 3, 6, 5, 0),
                                                       ("hex10", 2, 24, 10, 0), ("hexI", 4, 24, 6, 0), ("hexI", 8, 24, 6, 0)])
def test_gloo_ranks_match_single_domain(tmp_path, kind, world, n, steps, layers):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500) + world + (7 if kind != "hex" else 0) + (13 if kind == "hex10" else 0) + (23 if kind == "hexI" else 0)
    mp.spawn(_worker, args=(world, kind, n, steps, port, str(tmp_path), layers), nprocs=world, join=True)
    ref, vref, ncells, max_sweeps = _single(kind, n, steps)
    got = np.full(ncells, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got[d["gid"]] = d["alpha"]
        assert abs(d["v1"] - vref) <= 1e-13 * abs(vref)
        if kind.startswith("hex"):   # (the faceted outer boundary of the Kelvin mesh is open to the rotation flux)
            assert abs(d["v1"] - d["v0"]) <= 1e-13 * abs(d["v0"]), "volume conserved across ranks"
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() <= 1e-12, "decomposed run differs from single domain by %g" % np.abs(got - ref).max()
    if kind == "hex10":
        assert max_sweeps >= 2, "the case must exercise several bounding sweeps (got %d)" % max_sweeps


# ---- a decomposed CASE DIRECTORY: processor*/ with cellProcAddressing, run the way plicVofAdvectionFoam -parallel runs -----------
def _case_worker(rank, world, case_dir, port):
    import torch.distributed as dist
    from common import oracle_lib
    from geometricvofext_b200 import foamcase
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    out = foamcase.run_plic_vof_advection_decomposed(case_dir, rank, world, lib=oracle_lib())
    np.savez(os.path.join(case_dir, "run_rank%d.npz" % rank), steps=out["steps"], written=np.array(out["written"]), volume=out["volume"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_case_directory_run_equals_the_serial_case_run(tmp_path, world):
    """decomposePar's files (processorR/constant/polyMesh/cellProcAddressing, processorR/0/alpha.water) -> one process per
    rank over gloo, adaptive time loop with globally reduced Courant numbers -> processorR/<time>/alpha.water ->
    reconstructPar: the field equals the serial run of the same case (same steps, same write times, alpha <= 1e-12)."""
    import torch.multiprocessing as mp
    from common import fields, oracle_lib
    from geometricvofext_b200 import foamcase, foamfile
    from test_foam_formats import make_case
    case = make_case(str(tmp_path / "c"), n=14, end=0.01, wi=0.005)
    mesh = case.mesh()
    a0 = fields.sphere_alpha_quadrature(mesh)
    case.write_alpha(mesh, "0", a0)
    foamfile.write_polymesh(mesh, case.dir)            # the undecomposed mesh every rank cuts its part from
    # interface cells weigh more, as in the strong-scaling bench: the parts are not equal slabs
    cell_rank = foamcase.decompose_case(case, world, weights=1.0 + 50.0 * ((a0 > 0) & (a0 < 1)))
    assert sorted(np.unique(cell_rank).tolist()) == list(range(world))
    cr2, addr = foamcase.read_cell_proc_addressing(case)
    assert np.array_equal(cr2, cell_rank) and all(np.all(np.diff(a) > 0) for a in addr)
    serial = foamcase.run_plic_vof_advection(case, lib=oracle_lib())
    port = 29900 + (os.getpid() % 400) + world
    mp.spawn(_case_worker, args=(world, case.dir, port), nprocs=world, join=True)
    for r in range(world):
        d = np.load(os.path.join(case.dir, "run_rank%d.npz" % r))
        assert int(d["steps"]) == serial["steps"] and d["written"].tolist() == serial["written"]
        assert abs(float(d["volume"]) - serial["volume"]) <= 1e-13 * abs(serial["volume"])
    for name in serial["written"]:
        got = foamcase.reconstruct_par(case, name)
        ref = foamfile.read_field(os.path.join(case.dir, name, case.alpha_name)).internal_array(mesh.n_cells)
        assert np.abs(got - ref).max() <= 1e-12, "time %s: decomposed case run differs from the serial one by %g" % (name, np.abs(got - ref).max())
    assert np.abs(got - a0).max() > 1e-3          # the interface really moved
