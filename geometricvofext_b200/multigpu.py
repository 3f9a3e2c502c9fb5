"""Decomposed (one process per GPU) SimPLIC step: host-side driver.

The reference runs in parallel the OpenFOAM way: one MPI rank per scotch sub-domain, processor
patches, ~10 halo swaps / reductions per step.  Here every rank holds its cells
plus `layers` point-neighbour layers of ghost cells (svof_decompose in the library, or -- for
uniform boxes that are too big to materialise globally -- the same sub-mesh generated directly by
`hex_block`), runs the unchanged single-domain step on that sub-mesh, and ONE exchange per step
refreshes ghost alpha from the owning ranks.

On GPUs the exchange lives in the library (include/svof.h "decomposed runs": pack kernel, grouped
ncclSend/ncclRecv on the handle's stream, scatter kernel); Python only broadcasts the 128-byte NCCL
id (torch.distributed is the bootstrap transport, as MPI_Bcast would be inside OpenFOAM).  On CPUs
(tests: gloo, world_size 2..4, the oracle as per-rank engine) the same plan is executed with
torch.distributed send/recv of host arrays.

Dependency radius of one step: LS normal (1 point layer) -> plane/flux of the upwind cell (1 face
layer) -> one layer per bounding sweep, hence the default layers = nAlphaBounds + 2.
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import capi, fields
from .mesh import Patch, PolyMesh, hex_block
from .solver import SolveVofEqu


def default_layers(controls):
    return int((controls or {}).get("nAlphaBounds", 10)) + 2


# ------------------------------------------------------------------------ general meshes ----
def partition_rcb(mesh, n_parts, weights=None, lib=None):
    """cell -> rank by the library's weighted recursive coordinate bisection (scotch stand-in)."""
    lib = lib or capi.load_product()
    cm, keep = mesh.to_c()
    out = np.empty(mesh.n_cells, np.int32)
    w = capi.f64(weights, (mesh.n_cells,)) if weights is not None else None
    rc = lib.svof_partition_rcb(C.byref(cm), capi.dptr(w), int(n_parts), capi.iptr(out))
    del keep
    if rc:
        raise capi.SvofError(rc, lib.svof_decomp_last_error().decode())
    return out


def _np_from(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def decompose(mesh, cell_rank, rank, layers, lib=None):
    """(sub-mesh PolyMesh, maps) of `rank`: owned cells + `layers` ghost layers (svof_decompose)."""
    lib = lib or capi.load_product()
    cm, keep = mesh.to_c()
    cr = capi.i32(cell_rank)
    h = C.c_void_p()
    rc = lib.svof_decompose(C.byref(cm), capi.iptr(cr), int(rank), int(layers), C.byref(h))
    del keep
    if rc:
        raise capi.SvofError(rc, lib.svof_decomp_last_error().decode())
    try:
        sm = capi.SvofMesh()
        lib.svof_submesh_mesh(h, C.byref(sm))
        n_owned = C.c_int32()
        ptrs = [capi.c_int32_p() for _ in range(6)]
        lib.svof_submesh_maps(h, C.byref(n_owned), *[C.byref(p) for p in ptrs])
        nC, nF, nP, nIF = sm.n_cells, sm.n_faces, sm.n_points, sm.n_internal_faces
        fo = _np_from(sm.face_offsets, nF + 1, np.int32)
        patches = []
        for i in range(sm.n_patches):
            p = sm.patches[i]
            src = mesh.patches[i] if i < len(mesh.patches) else None
            patches.append(Patch(src.name if src else "cut", p.start, p.size, p.kind, p.nbr_rank, p.alpha_bc, p.alpha_value))
        sub = PolyMesh(points=_np_from(sm.points, 3 * nP, np.float64).reshape(-1, 3), face_offsets=fo,
                       face_points=_np_from(sm.face_points, int(fo[-1]), np.int32),
                       owner=_np_from(sm.owner, nF, np.int32), neighbour=_np_from(sm.neighbour, nIF, np.int32),
                       patches=patches, n_cells=nC, meta=dict(mesh.meta, kind="submesh", rank=int(rank), layers=int(layers)))
        maps = {"n_owned": int(n_owned.value),
                "cell_global": _np_from(ptrs[0], nC, np.int32), "cell_owner_rank": _np_from(ptrs[1], nC, np.int32),
                "cell_layer": _np_from(ptrs[2], nC, np.int32), "owned_local": _np_from(ptrs[3], n_owned.value, np.int32),
                "face_global": _np_from(ptrs[4], nF, np.int32), "point_global": _np_from(ptrs[5], nP, np.int32)}
        sub.cell_global = maps["cell_global"].astype(np.int64)
    finally:
        lib.svof_submesh_free(h)
    return sub, maps


# --------------------------------------------------------------------------- uniform boxes ----
def block_grid(p):
    """(px, py, pz) with px*py*pz == p, as cubic as possible, z split first (slabs for small p)."""
    best = None
    for px in range(1, p + 1):
        if p % px:
            continue
        for py in range(1, p // px + 1):
            if (p // px) % py:
                continue
            pz = p // px // py
            if px <= py <= pz:
                cand = (pz - px, (px, py, pz))
                if best is None or cand < best:
                    best = cand
    return best[1]


def box_rcb(n, world, weight_fn=None):
    """Recursive bisection of the index box [0,n) into `world` boxes of (nearly) equal weight.
    weight_fn(axis, lo, hi) -> 1-D array of the summed cell weights of each index plane of box [lo,hi) along
    `axis` (None: uniform).  Returns a list of (lo, hi) integer triples; box r belongs to rank r."""
    N = np.array(n if np.ndim(n) else (n, n, n), dtype=np.int64)

    def split(lo, hi, parts):
        if parts == 1:
            return [(lo.copy(), hi.copy())]
        ext = hi - lo
        ax = int(np.argmax(ext))   # first longest axis
        w = weight_fn(ax, lo, hi) if weight_fn is not None else np.full(int(ext[ax]), float(np.prod(ext) // ext[ax]))
        pl = parts // 2
        cum = np.cumsum(w)
        target = cum[-1] * pl / parts
        cut = int(np.searchsorted(cum, target, side="left")) + 1
        cut = max(1, min(int(ext[ax]) - 1, cut))
        hi_l, lo_r = hi.copy(), lo.copy()
        hi_l[ax] = lo[ax] + cut
        lo_r[ax] = lo[ax] + cut
        return split(lo, hi_l, pl) + split(lo_r, hi, parts - pl)

    return split(np.zeros(3, np.int64), N, int(world))


class BoxDecomposition:
    """Decomposition of an N=(Nx,Ny,Nz) uniform hex box into `world` boxes with `layers` ghost layers, generated
    per rank without ever materialising the global mesh (512^3 = 134 M cells).  Produces exactly what svof_decompose
    would: cells in ascending global label, ghost addressing, cut faces closed by the physical patches."""

    def __init__(self, n, world, layers, length=(1.0, 1.0, 1.0), boxes=None):
        self.N = np.array(n if np.ndim(n) else (n, n, n), dtype=np.int64)
        self.world, self.G = int(world), int(layers)
        self.length = tuple(float(x) for x in length)
        if boxes is None:
            grid = np.array(block_grid(world), dtype=np.int64)
            boxes = []
            for r in range(world):
                c = np.array([r % grid[0], (r // grid[0]) % grid[1], r // (grid[0] * grid[1])], dtype=np.int64)
                boxes.append(((self.N * c) // grid, (self.N * (c + 1)) // grid))
        self.boxes = [(np.array(lo, dtype=np.int64), np.array(hi, dtype=np.int64)) for lo, hi in boxes]

    def owned_box(self, rank):
        return self.boxes[rank]

    def ext_box(self, rank):
        lo, hi = self.boxes[rank]
        return np.maximum(lo - self.G, 0), np.minimum(hi + self.G, self.N)

    def rank_mesh(self, rank):
        """(sub-mesh, maps) of `rank`."""
        elo, ehi = self.ext_box(rank)
        m = hex_block(self.N, lo=elo, hi=ehi, length=self.length, cut_as_wall=True)
        nx, ny, nz = (ehi - elo).tolist()
        owner = np.full((nz, ny, nx), -1, dtype=np.int32)
        for r, (lo, hi) in enumerate(self.boxes):
            a, b = np.maximum(lo, elo) - elo, np.minimum(hi, ehi) - elo
            if np.all(b > a):
                owner[a[2]:b[2], a[1]:b[1], a[0]:b[0]] = r
        owner = owner.reshape(-1)
        assert owner.min() >= 0
        maps = {"cell_global": m.cell_global.astype(np.int32) if int(np.prod(self.N)) < 2 ** 31 else m.cell_global,
                "cell_owner_rank": owner, "owned_local": np.nonzero(owner == rank)[0].astype(np.int32)}
        maps["n_owned"] = int(maps["owned_local"].size)
        return m, maps


# ------------------------------------------------------------------------------ the driver ----
class DecomposedSolveVofEqu:
    """solveVofEqu on one rank of a decomposed mesh; same member names as SolveVofEqu.

    sub, maps   the rank's sub-mesh and addressing (from decompose() or BoxDecomposition.rank_mesh())
    device      CUDA ordinal -> the library's own NCCL exchange; None -> CPU engine (`lib`), gloo exchange
    """

    def __init__(self, sub, maps, controls, rank, world, lib=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = rank, world
        self.mesh, self.maps = sub, maps
        self.owned = np.asarray(maps["owned_local"], dtype=np.int64)
        self.cell_global = np.asarray(maps["cell_global"])
        self.owner_rank = np.asarray(maps["cell_owner_rank"], dtype=np.int32)
        self.on_gpu = device is not None
        self.device = device
        if self.on_gpu:
            import torch
            self.torch = torch
            lib = lib or capi.load_product()
            idbuf = (C.c_char * 128)()
            if world > 1:
                if rank == 0:
                    rc = lib.svof_comm_unique_id(idbuf)
                    if rc:
                        raise capi.SvofError(rc, lib.svof_last_error(None).decode())
                t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).clone()
                if dist.get_backend() == "nccl":
                    t = t.cuda(device)
                dist.broadcast(t, 0)
                idbuf = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
            self._idbuf = idbuf
            self.s = SolveVofEqu(sub, controls, lib=lib, comm=(rank, world, device, C.addressof(idbuf) if world > 1 else None))
            cg, co = capi.i32(self.cell_global), capi.i32(self.owner_rank)
            self.s._chk(lib.svof_halo_setup(self.s._h, capi.iptr(cg), capi.iptr(co)))
            self.halo_bytes = int(self.s.info(capi.I_HALO_BYTES))
        else:
            self.s = SolveVofEqu(sub, controls, lib=lib, comm=(rank, 1, -1, None))   # per-rank single-domain engine
            self._plan_host()

    # -- CPU path: the plan the library builds with NCCL, built with torch.distributed objects ------------
    def _plan_host(self):
        dist = self.dist
        ghosts = np.nonzero(self.owner_rank != self.rank)[0]
        need = {}
        for r in np.unique(self.owner_rank[ghosts]):
            idx = ghosts[self.owner_rank[ghosts] == r]
            need[int(r)] = (idx, self.cell_global[idx])
        self.recv = {r: idx for r, (idx, _) in need.items()}
        wants = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(wants, {r: g for r, (_, g) in need.items()})
        self.send = {}
        for q in range(self.world):
            if q == self.rank or not wants[q] or self.rank not in wants[q]:
                continue
            g = np.asarray(wants[q][self.rank])
            pos = np.searchsorted(self.cell_global, g)
            assert np.array_equal(self.cell_global[pos], g) and np.all(self.owner_rank[pos] == self.rank)
            self.send[q] = pos
        self.halo_bytes = 8 * sum(len(v) for v in self.recv.values())

    def exchange_alpha(self):
        """ghost alpha <- owning rank"""
        if self.world == 1:
            return
        if self.on_gpu:
            self.s._chk(self.s.lib.svof_halo_exchange(self.s._h))
            return
        import torch
        dist = self.dist
        a = self.s.alpha()
        reqs, bufs = [], {}
        for r, idx in self.send.items():
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[idx])), r))
        for r, idx in self.recv.items():
            bufs[r] = torch.empty(len(idx), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[r], r))
        for q in reqs:
            q.wait()
        for r, idx in self.recv.items():
            a[idx] = bufs[r].numpy()
        self.s.setAlpha(a)

    # -- the reference's member functions -----------------------------------------------------------
    def setAlpha(self, alpha_local):
        self.s.setAlpha(alpha_local)

    def setPhi(self, phi):
        self.s.setPhi(phi)

    def setU(self, U, Ub=None):
        self.s.setU(U, Ub)

    def reconstruct(self):
        self.s.reconstruct()

    def advect(self, dt, Sp=None, Su=None):
        self.s.advect(dt, Sp, Su)          # on the device the library appends the ghost refresh itself
        if not self.on_gpu:
            self.exchange_alpha()

    def step(self, dt):
        """reconstruct() + advect(dt) (one CUDA-graph launch on the device) + the ghost refresh."""
        self.s.step(dt)
        if not self.on_gpu:
            self.exchange_alpha()

    def alpha_owned(self):
        return self.s.alpha()[self.owned]

    def owned_global_ids(self):
        return self.cell_global[self.owned]

    def volume(self):
        """gSum(alpha*V) over the owned cells of all ranks."""
        import torch
        if self.on_gpu:
            v = self.s.info(capi.I_VOLUME_OWNED)
        else:
            v = float(np.sum(self.s.alpha()[self.owned] * self.s.field(capi.F_V)[self.owned]))
        t = torch.tensor([v], dtype=torch.float64)
        if self.on_gpu and self.dist.get_backend() == "nccl":
            t = t.cuda(self.device)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def close(self):
        self.s.close()


# ------------------------------------------------------------------------------------ bench ----
def sphere_plane_weights(n, mixed_weight, centre=(0.35, 0.35, 0.35), radius=0.15):
    """weight_fn for box_rcb on the LeVeque initial field: 1 per cell + `mixed_weight` per interface cell, where the
    interface cells of a sphere are counted per index plane from the spherical-shell area (no global field needed)."""
    n = int(n)
    h = 1.0 / n

    def fn(axis, lo, hi):
        ext = (hi - lo).astype(np.float64)
        base = float(np.prod(ext) / ext[axis])
        x = (np.arange(lo[axis], hi[axis]) + 0.5) * h
        # area of the sphere surface inside the slab of plane i and inside the box's other two extents ~ a band of a sphere:
        # 2*pi*r*h per plane for |x - c| < r (Archimedes), restricted by the fraction of the band inside the box
        w = np.full(x.shape, base)
        inside = np.abs(x - centre[axis]) < radius
        if inside.any():
            oth = [d for d in range(3) if d != axis]
            frac = np.ones(x.shape)
            # fraction of the circle of latitude inside the box (sampled)
            th = np.linspace(0.0, 2 * np.pi, 256, endpoint=False)
            for k in np.nonzero(inside)[0]:
                rr = np.sqrt(max(radius ** 2 - (x[k] - centre[axis]) ** 2, 0.0))
                p0 = centre[oth[0]] + rr * np.cos(th)
                p1 = centre[oth[1]] + rr * np.sin(th)
                ok = (p0 >= lo[oth[0]] * h) & (p0 < hi[oth[0]] * h) & (p1 >= lo[oth[1]] * h) & (p1 < hi[oth[1]] * h)
                frac[k] = ok.mean()
            band_cells = 2 * np.pi * radius * h / (h * h) * 1.5   # interface cells per plane (shell ~1.5 cells thick)
            w = w + mixed_weight * band_cells * frac * inside
        return w

    return fn


def interface_boxes(n, world, centre=(0.35, 0.35, 0.35), radius=0.15, margin=3, k_bulk=25.7e-9, chain=(0.35, 17.2e-6), hide=0.76):
    """Boxes for the strong-scaling run that put the INTERFACE on half of the ranks and give those ranks few bulk cells.

    Cost model of one rank (fitted on the single-GPU profiles): the interface chain costs
    chain[0] + chain[1] * interface cells [ms] however few cells the rank streams, and the streaming kernel k_bulk ms per
    cell, of which `hide` still shows when both run on one GPU.  A bisection that balances (cells + w * interface cells)
    hands every rank a piece of the sphere, so EVERY rank pays the chain on top of its share of the bulk.  Here the
    corner region [0,qx) x [0,qy) x [0,qz) that contains the sphere is split, through the sphere's centre, over world/2
    "interface ranks"; the rest of the box -- three slabs -- goes to the other ranks as pure streaming work:
        world 8:  4 interface boxes (x, y split)  |  x >= qx in two halves, {x < qx, y >= qy}, {x < qx, y < qy, z >= qz}
        world 4:  2 interface boxes (x split) of the column [0,qx) x [0,qy) x [0,n)  |  x >= qx, {x < qx, y >= qy}
    qx, qy, qz >= the sphere's extent + margin are chosen to even out the bulk ranks.  Other rank counts: None (the caller
    falls back to the weighted bisection).  Returns a list of (lo, hi) integer triples, box r for rank r."""
    n = int(n)
    qmin = int(np.ceil((max(centre) + radius) * n)) + margin
    c = [int(round(x * n)) for x in centre]
    if qmin >= n - 1:
        return None
    best = None
    if world == 8:
        for qz in range(qmin, n, max(1, n // 64)):
            for qy in range(qmin, n, max(1, n // 64)):
                for qx in range(qmin, n, max(1, n // 64)):
                    A = (n - qx) * n * n / 2.0
                    B = qx * (n - qy) * n
                    Cc = qx * qy * (n - qz)
                    R = qx * qy * qz / 4.0
                    shell = 4 * np.pi * (radius * n) ** 2 * 1.5 / 4.0
                    t = max(k_bulk * max(A, B, Cc), chain[0] + chain[1] * shell + hide * k_bulk * R)
                    if best is None or t < best[0]:
                        best = (t, qx, qy, qz)
        _, qx, qy, qz = best
        h = n // 2
        return [((0, 0, 0), (c[0], c[1], qz)), ((c[0], 0, 0), (qx, c[1], qz)), ((0, c[1], 0), (c[0], qy, qz)), ((c[0], c[1], 0), (qx, qy, qz)),
                ((qx, 0, 0), (n, h, n)), ((qx, h, 0), (n, n, n)), ((0, qy, 0), (qx, n, n)), ((0, 0, qz), (qx, qy, n))]
    if world == 4:
        for qy in range(qmin, n, max(1, n // 64)):
            for qx in range(qmin, n, max(1, n // 64)):
                A = (n - qx) * n * n
                B = qx * (n - qy) * n
                R = qx * qy * n / 2.0
                shell = 4 * np.pi * (radius * n) ** 2 * 1.5 / 2.0
                t = max(k_bulk * max(A, B), chain[0] + chain[1] * shell + hide * k_bulk * R)
                if best is None or t < best[0]:
                    best = (t, qx, qy)
        _, qx, qy = best
        return [((0, 0, 0), (c[0], qy, n)), ((c[0], 0, 0), (qx, qy, n)), ((qx, 0, 0), (n, n, n)), ((0, qy, 0), (qx, n, n))]
    return None


def bench(args, controls, metric, unit):
    """N-GPU leg of bench.py: STRONG scaling of one LeVeque problem (default 512^3, BASELINE.json configs[4]) over
    `world` ranks, or (--scaling weak) one 256^3 unit cube per GPU."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL's version banner goes to stdout: keep rank 0's stdout to the one JSON line
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    strong = args.scaling == "strong"
    workload_name = getattr(args, "workload", "leveque")
    if workload_name == "dambreak":
        # BASELINE.json configs[3] substitute: the damBreakWithObstacle box and alpha controls, 8-way
        import bench as _bench
        controls = dict(_bench.DAMBREAK_CONTROLS)
        strong = True
    layers = args.layers if args.layers > 0 else default_layers(controls)
    t0 = time.perf_counter()
    if workload_name == "dambreak":
        n = args.n
        dec = BoxDecomposition(n, world, layers, boxes=box_rcb(n, world, None))
        centre = None
        dt = 0.2 / n
        workload = ("damBreakWithObstacle substitute, advection only: %d^3 hex box (%.1f M cells), water box (0.6 x 0.1875 x 0.75), clip true / "
                    "snapTol 1e-12 / mixedCellTol 1e-10, inletOutlet top patch, prescribed 3-D deformation velocity, split over %d ranks "
                    "(BASELINE.json configs[3])" % (n, n ** 3 / 1e6, world))
    elif strong:
        n = args.strong_n
        wfn = sphere_plane_weights(n, args.mixed_weight) if args.mixed_weight > 0 else None
        # the margin keeps the sphere out of the GHOST layers of the bulk ranks too (a rank with any interface cell pays the chain)
        boxes = interface_boxes(n, world, margin=layers + 2) if getattr(args, "partition", "rcb") == "interface" else None
        how = "into interface boxes (the sphere on %d ranks with few bulk cells) and bulk boxes (streaming only)" % (world // 2)
        if boxes is None:
            boxes = box_rcb(n, world, wfn)
            how = "by weighted recursive bisection into boxes (interface cells weigh %g)" % args.mixed_weight
        dec = BoxDecomposition(n, world, layers, boxes=boxes)
        centre = (0.35, 0.35, 0.35)
        dt = 0.2 / n
        workload = ("LeVeque 3-D deformation, sphere r=0.15, %d^3 hex blockMesh (BASELINE.json configs[4]), ONE problem split over "
                    "%d ranks %s" % (n, world, how))
    else:
        grid = np.array(block_grid(world))
        n_global = (args.n * grid).tolist()
        dec = BoxDecomposition(n_global, world, layers, length=grid.astype(float).tolist())
        c = np.array([rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1])], dtype=float)
        centre = tuple(c + 0.35)
        dt = 0.2 / args.n
        workload = ("LeVeque 3-D deformation tiled %dx%dx%d: one unit cube (%d^3 cubic hex cells, sphere r=0.15) per GPU" %
                    (grid[0], grid[1], grid[2], args.n))
    sub, maps = dec.rank_mesh(rank)
    if workload_name == "dambreak":
        from .mesh import cell_centres_hex
        for p_ in sub.patches:
            if p_.name == "top":
                p_.alpha_bc, p_.alpha_value = capi.BC_INLET_OUTLET, 0.0
    ds = DecomposedSolveVofEqu(sub, maps, controls, rank, world, device=local)
    s = ds.s
    if workload_name == "dambreak":
        Cc = cell_centres_hex(sub)
        a0 = ((Cc[:, 0] < 0.6) & (Cc[:, 1] < 0.1875) & (Cc[:, 2] < 0.75)).astype(np.float64)
        del Cc
    else:
        a0 = fields.sphere_alpha_quadrature(sub, centre=centre)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    f = fields.u_factor(dt, dt, 6.0)
    U = fields.leveque_velocity(C_) * f
    phi = fields.face_flux(Cf, Sf) * f
    del C_, Cf, Sf
    s.setAlpha(a0)
    s.setPhi(phi)
    s.setU(U, np.zeros((s.nBF, 3)))
    del U, phi
    if args.overlap >= 0:
        s.setOption("overlap", args.overlap)
    ds.exchange_alpha()
    setup_s = time.perf_counter() - t0

    for _ in range(2 + max(3, args.warmup) + (getattr(args, "develop_steps", 0) if workload_name == "dambreak" else 0)):
        ds.step(dt)
    s.synchronize()
    l0 = s.info(capi.I_GPU_LAUNCHES)
    dist.barrier()
    torch.cuda.synchronize()
    s.lib.svof_mark(s._h, 0)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        ds.step(dt)
    s.lib.svof_mark(s._h, 1)
    s.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    ms = C.c_double()
    s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms))
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)       # job time = slowest rank (CUDA events on each handle's stream)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    total_ms = float(t.item())
    # ---- end-to-end leg: every rank drives its sub-domain through svof_step_host with pinned host buffers (phi, U in;
    #      alpha, alphaPhi out; the ghost refresh is part of the call); max over ranks of the host wall time of the calls
    e2e = None
    e2e_steps = max(1, min(args.steps, getattr(args, "e2e_steps", 5)))
    try:
        if not getattr(args, "e2e_multi", False):   # opt-in (--e2e-multi): a rank that fails inside this leg leaves its peers in NCCL calls
            raise RuntimeError("not requested (--e2e-multi); the host-buffer end-to-end number is measured at N=1")
        lib = s.lib
        C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
        U0 = fields.leveque_velocity(C_) * f
        phi0 = fields.face_flux(Cf, Sf) * f
        del C_, Cf, Sf
        phi_h, U_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3))
        Ub_h = capi.pinned_array(lib, (max(s.nBF, 1), 3))
        a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
        phi_h[:] = phi0
        U_h[:] = U0
        Ub_h[:] = 0
        s.setAlpha(a0)              # the same steps of the same problem as the device-resident leg: from t = 0
        ds.exchange_alpha()
        for _ in range(1 + max(3, args.warmup)):
            s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
        dist.barrier()
        e2e_s = 0.0
        for k in range(e2e_steps):
            fk = 1.0 - 1e-3 * (k + 1)
            np.multiply(phi0, fk, out=phi_h)
            np.multiply(U0, fk, out=U_h)
            dist.barrier()
            t0 = time.perf_counter()
            s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
            e2e_s += time.perf_counter() - t0
        te = torch.tensor([e2e_s, float(s.info(capi.I_H2D_BYTES)), float(s.info(capi.I_D2H_BYTES))], dtype=torch.float64, device="cuda")
        tsum = te.clone()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e = {"seconds": float(te[0].item()), "h2d": int(tsum[1].item()), "d2h": int(tsum[2].item()), "steps": e2e_steps}
        del U0, phi0
    except Exception as ex:   # the device-resident numbers above stand on their own
        e2e = {"error": str(ex)}
    stats = torch.tensor([float(len(ds.owned)), float(s.nC), float(s.info(capi.I_N_MIXED)), float(ds.halo_bytes)],
                         dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)
    per_rank = [[float(x) for x in g.tolist()] for g in gathered]
    cells = sum(p[0] for p in per_rank)
    launches = int(s.info(capi.I_GPU_LAUNCHES) - l0)
    value = cells * args.steps / (total_ms * 1e-3)
    vol = ds.volume()
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "cells": int(cells), "dt": dt, "controls": controls, "ghost_layers": layers,
                       "per_rank": [{"owned_cells": int(p[0]), "local_cells": int(p[1]), "mixed_cells_local": int(p[2]),
                                     "halo_bytes_per_step": int(p[3])} for p in per_rank],
                       "exchange": "1 ghost-alpha refresh per step: library-side pack kernel + grouped ncclSend/ncclRecv + scatter kernel "
                                   "on the solver's stream (no torch in the data path)",
                       "l2": "inputs larger than L2", "timing": "CUDA events on each rank's stream, max over ranks (min %.4f ms/step); "
                       "host wall %.3f ms/step" % (float(tmin.item()) / args.steps, wall_ms / args.steps),
                       "setup_s": setup_s, "volume": vol},
            "gpu_launches": launches,
            "e2e": ({"value": cells * e2e["steps"] / e2e["seconds"], "unit": unit, "h2d_bytes_per_step": e2e["h2d"],
                     "d2h_bytes_per_step": e2e["d2h"], "steps": e2e["steps"], "ms_per_step": 1e3 * e2e["seconds"] / e2e["steps"],
                     "note": "every rank calls svof_step_host on its sub-domain with pinned host buffers (phi, U in; alpha, alphaPhi "
                             "out; phi and U rescaled on the host between calls; the NCCL ghost refresh is inside the call); host "
                             "wall time of the calls, max over ranks; bytes summed over ranks, last call"}
                    if e2e and "seconds" in e2e else
                    {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                     "note": "multi-GPU leg is device resident (no host copies declared): %s" % (e2e or {}).get("error")}),
        }
        if strong and workload_name != "dambreak":
            base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "strong_base_%d.json" % args.strong_n)
            if os.path.exists(base):
                try:
                    line["config"]["single_gpu_same_problem"] = json.load(open(base))
                except Exception:
                    pass
        print(json.dumps(line))
    dist.barrier()
    ds.close()
    dist.barrier()
    dist.destroy_process_group()
