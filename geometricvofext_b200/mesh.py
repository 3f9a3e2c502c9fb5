"""Host-side mesh mirror: OpenFOAM polyMesh layout as numpy arrays.

`PolyMesh` holds exactly what constant/polyMesh/{points,faces,owner,neighbour,
boundary} hold; `hex_block` is the blockMesh equivalent used by the reference's
test case (tutorials/test/plicVofAdvectionFoam/system/blockMeshDict:17-90) and
can also emit one sub-block of a decomposed box with processor patches in
OpenFOAM's convention (physical patches first, processor patches after, faces
of a processor patch in the same order on both sides).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class Patch:
    name: str
    start: int
    size: int
    kind: int = capi.PATCH_GENERIC
    nbr_rank: int = -1
    alpha_bc: int = capi.BC_ZERO_GRADIENT
    alpha_value: float = 0.0


@dataclass
class PolyMesh:
    points: np.ndarray          # [nP,3] f64
    face_offsets: np.ndarray    # [nF+1] i32
    face_points: np.ndarray     # [sum] i32
    owner: np.ndarray           # [nF] i32
    neighbour: np.ndarray       # [nIF] i32
    patches: list = field(default_factory=list)
    n_cells: int = 0
    # global addressing of a decomposed block (cellProcAddressing-style), optional
    cell_global: np.ndarray = None
    meta: dict = field(default_factory=dict)

    @property
    def n_points(self):
        return self.points.shape[0]

    @property
    def n_faces(self):
        return self.owner.shape[0]

    @property
    def n_internal_faces(self):
        return self.neighbour.shape[0]

    @property
    def n_boundary_faces(self):
        return self.n_faces - self.n_internal_faces

    def to_c(self):
        """svof_mesh struct (+ keep-alive list for the arrays it points into)."""
        pts = capi.f64(self.points.reshape(-1))
        fo, fp = capi.i32(self.face_offsets), capi.i32(self.face_points)
        own, nei = capi.i32(self.owner), capi.i32(self.neighbour)
        parr = (capi.SvofPatch * max(1, len(self.patches)))()
        for i, p in enumerate(self.patches):
            parr[i] = capi.SvofPatch(p.start, p.size, p.kind, p.nbr_rank, p.alpha_bc, 0, p.alpha_value)
        m = capi.SvofMesh()
        m.n_points, m.n_faces, m.n_internal_faces = self.n_points, self.n_faces, self.n_internal_faces
        m.n_cells, m.n_patches = self.n_cells, len(self.patches)
        m.points, m.face_offsets, m.face_points = capi.dptr(pts), capi.iptr(fo), capi.iptr(fp)
        m.owner, m.neighbour = capi.iptr(own), capi.iptr(nei)
        m.patches = C.cast(parr, C.POINTER(capi.SvofPatch))
        m.Cf = m.Sf = m.C = m.V = None
        return m, [pts, fo, fp, own, nei, parr]

    def face_centres_simple(self):
        """Vertex-mean face centres (exact for planar parallelograms); used only to
        sample analytic velocity fields in the harness, never by the solver."""
        n = np.diff(self.face_offsets)
        s = np.add.reduceat(self.points[self.face_points], self.face_offsets[:-1], axis=0)
        return s / n[:, None]


# patch order and names of tutorials/test/plicVofAdvectionFoam/system/blockMeshDict:44-90
_HEX_PATCHES = [("top", 2, 1), ("left", 1, 0), ("back", 0, 0), ("right", 1, 1), ("bottom", 2, 0), ("front", 0, 1)]


def hex_block(n, lo=(0, 0, 0), hi=None, length=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), proc_nbr=None,
              cut_as_wall=False):
    """Uniform hex mesh of cells [lo,hi) out of a global n=(Nx,Ny,Nz) box.

    Natural cell order c = i + nx*j + nx*ny*k, upper-triangular face order (what
    blockMesh emits).  `proc_nbr` maps (axis, side) -> neighbour rank for sides
    that are cut by a decomposition; those become processor patches listed after
    the six physical patches (zero-sized where the side is not physical).
    `cut_as_wall` closes cut sides with the physical patch instead (stand-alone sub-domain).
    """
    N = np.array(n if np.ndim(n) else (n, n, n), dtype=np.int64)
    lo = np.array(lo, dtype=np.int64)
    hi = N.copy() if hi is None else np.array(hi, dtype=np.int64)
    nx, ny, nz = (hi - lo).tolist()
    proc_nbr = dict(proc_nbr or {})
    h = np.array(length, dtype=np.float64) / N

    # points: coordinates computed from GLOBAL indices so sub-blocks are bit-identical to the full mesh
    # (blockMesh-like: vertices at exact fractions i/N of the edge)
    gx = origin[0] + length[0] * ((lo[0] + np.arange(nx + 1)) / N[0])
    gy = origin[1] + length[1] * ((lo[1] + np.arange(ny + 1)) / N[1])
    gz = origin[2] + length[2] * ((lo[2] + np.arange(nz + 1)) / N[2])
    px, py, pz = nx + 1, ny + 1, nz + 1
    pts = np.empty((pz, py, px, 3), dtype=np.float64)
    pts[..., 0] = gx[None, None, :]
    pts[..., 1] = gy[None, :, None]
    pts[..., 2] = gz[:, None, None]
    pts = pts.reshape(-1, 3)

    def pid(i, j, k):
        return (i + px * (j + py * k)).astype(np.int32)

    def cid(i, j, k):
        return (i + nx * (j + ny * k)).astype(np.int32)

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    c = cid(I, J, K)

    # internal faces: per cell (ascending) the faces towards +x, +y, +z neighbours
    def xface(i, j, k):   # plane i (point index), normal +x
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)

    def yface(i, j, k):   # plane j, normal +y
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], axis=1)

    def zface(i, j, k):   # plane k, normal +z
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    mx, my, mz = I < nx - 1, J < ny - 1, K < nz - 1
    # slot = 3*cell + axis keeps (owner asc, neighbour asc) order
    slot = np.concatenate([3 * c[mx].astype(np.int64), 3 * c[my].astype(np.int64) + 1, 3 * c[mz].astype(np.int64) + 2])
    fpts = np.concatenate([xface(I[mx] + 1, J[mx], K[mx]), yface(I[my], J[my] + 1, K[my]), zface(I[mz], J[mz], K[mz] + 1)])
    own = np.concatenate([c[mx], c[my], c[mz]])
    nei = np.concatenate([cid(I[mx] + 1, J[mx], K[mx]), cid(I[my], J[my] + 1, K[my]), cid(I[mz], J[mz], K[mz] + 1)])
    order = np.argsort(slot, kind="stable")
    fpts, own, nei = fpts[order], own[order], nei[order]
    n_if = own.shape[0]

    # boundary faces (outward normals), per side
    def side_faces(axis, side):
        if axis == 0:
            k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
            j, k = j.reshape(-1), k.reshape(-1)
            i = np.full_like(j, 0 if side == 0 else nx - 1)
            f = xface(i + side, j, k)
        elif axis == 1:
            k, i = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
            i, k = i.reshape(-1), k.reshape(-1)
            j = np.full_like(i, 0 if side == 0 else ny - 1)
            f = yface(i, j + side, k)
        else:
            j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
            i, j = i.reshape(-1), j.reshape(-1)
            k = np.full_like(i, 0 if side == 0 else nz - 1)
            f = zface(i, j, k + side)
        if side == 0:
            f = f[:, [0, 3, 2, 1]]  # reverse to point outward (-axis)
        return f, cid(i, j, k)

    b_fpts, b_own, patches = [], [], []
    start = n_if
    for name, axis, side in _HEX_PATCHES:
        physical = cut_as_wall or ((lo[axis] == 0) if side == 0 else (hi[axis] == N[axis]))
        if physical and (axis, side) not in proc_nbr:
            f, o = side_faces(axis, side)
        else:
            f, o = np.zeros((0, 4), np.int32), np.zeros((0,), np.int32)
        patches.append(Patch(name, start, f.shape[0]))
        b_fpts.append(f)
        b_own.append(o)
        start += f.shape[0]
    for (axis, side), nbr in sorted(proc_nbr.items(), key=lambda kv: kv[1]):
        f, o = side_faces(axis, side)
        patches.append(Patch("procBoundaryTo%d" % nbr, start, f.shape[0], kind=capi.PATCH_PROCESSOR, nbr_rank=int(nbr)))
        b_fpts.append(f)
        b_own.append(o)
        start += f.shape[0]

    face_points = np.concatenate([fpts] + b_fpts).astype(np.int32).reshape(-1)
    owner = np.concatenate([own] + b_own).astype(np.int32)
    n_f = owner.shape[0]
    face_offsets = (4 * np.arange(n_f + 1, dtype=np.int64)).astype(np.int32)
    gi, gj, gk = I + lo[0], J + lo[1], K + lo[2]
    cell_global = (gi + N[0] * (gj + N[1] * gk)).astype(np.int64)
    return PolyMesh(points=pts, face_offsets=face_offsets, face_points=face_points, owner=owner,
                    neighbour=nei.astype(np.int32), patches=patches, n_cells=nx * ny * nz, cell_global=cell_global,
                    meta={"kind": "hex_block", "N": N.tolist(), "lo": lo.tolist(), "hi": hi.tolist(),
                          "h": h.tolist(), "origin": list(origin), "length": list(length)})


def cell_centres_hex(mesh):
    """Cell centres of a hex_block mesh from its metadata (harness use only)."""
    N, lo, hi = (np.array(mesh.meta[k]) for k in ("N", "lo", "hi"))
    L, o = np.array(mesh.meta["length"]), np.array(mesh.meta["origin"])
    nx, ny, nz = (hi - lo).tolist()
    cx = o[0] + L[0] * ((lo[0] + np.arange(nx) + 0.5) / N[0])
    cy = o[1] + L[1] * ((lo[1] + np.arange(ny) + 0.5) / N[1])
    cz = o[2] + L[2] * ((lo[2] + np.arange(nz) + 0.5) / N[2])
    C3 = np.empty((nz, ny, nx, 3))
    C3[..., 0] = cx[None, None, :]
    C3[..., 1] = cy[None, :, None]
    C3[..., 2] = cz[:, None, None]
    return C3.reshape(-1, 3)


# ---------------------------------------------------------------------------------------------
# general polyhedral meshes (the arbitrary-polyhedron PLIC path)
# ---------------------------------------------------------------------------------------------
def perturb_points(mesh, amplitude=0.15, seed=0):
    """Jitter the INTERIOR points of a hex_block mesh by `amplitude` x cell size: every interior
    face becomes a warped (non-planar) quad -> face flatness < 1, the warped branch of
    timeIntegratedFaceFlux (cutFace.C:312-347) and splitWarpedFace (cutCell.C:140-236)."""
    rng = np.random.default_rng(seed)
    h = np.array(mesh.meta["h"])
    lo = np.array(mesh.meta["origin"])
    hi = lo + np.array(mesh.meta["length"])
    pts = mesh.points.copy()
    interior = np.all((pts > lo + 1e-12) & (pts < hi - 1e-12), axis=1)
    pts[interior] += amplitude * h * rng.uniform(-1.0, 1.0, size=(int(interior.sum()), 3))
    out = PolyMesh(points=pts, face_offsets=mesh.face_offsets, face_points=mesh.face_points, owner=mesh.owner,
                   neighbour=mesh.neighbour, patches=list(mesh.patches), n_cells=mesh.n_cells,
                   cell_global=mesh.cell_global, meta=dict(mesh.meta, kind="perturbed_hex"))
    return out


def build_polymesh(points, cells, patch_of=None, patch_names=("walls",)):
    """Assemble an OpenFOAM-ordered polyMesh from cells given as lists of outward-oriented faces.

    Internal faces come out in upper-triangular order (sorted by owner, then neighbour, oriented
    owner -> neighbour); boundary faces are grouped into patches by `patch_of(face_centre) -> index`.
    """
    face_map = {}
    for c, faces in enumerate(cells):
        for f in faces:
            key = tuple(sorted(f))
            if key in face_map:
                face_map[key][2] = c
            else:
                face_map[key] = [c, list(f), -1]
    internal = [(o, n, v) for (o, v, n) in face_map.values() if n >= 0]
    internal.sort(key=lambda t: (t[0], t[1]))
    boundary = [(o, v) for (o, v, n) in face_map.values() if n < 0]
    pts = np.asarray(points, dtype=np.float64)
    npatch = len(patch_names)
    groups = [[] for _ in range(npatch)]
    for o, v in boundary:
        pi = 0 if patch_of is None else int(patch_of(pts[v].mean(axis=0)))
        groups[pi].append((o, v))
    for g in groups:
        g.sort(key=lambda t: t[0])
    owner = [t[0] for t in internal]
    neighbour = [t[1] for t in internal]
    fverts = [t[2] for t in internal]
    patches = []
    start = len(internal)
    for name, g in zip(patch_names, groups):
        patches.append(Patch(name, start, len(g)))
        owner += [t[0] for t in g]
        fverts += [t[1] for t in g]
        start += len(g)
    off = np.zeros(len(fverts) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(v) for v in fverts])
    return PolyMesh(points=pts, face_offsets=off, face_points=np.concatenate(fverts).astype(np.int32),
                    owner=np.array(owner, dtype=np.int32), neighbour=np.array(neighbour, dtype=np.int32),
                    patches=patches, n_cells=len(cells), meta={"kind": "poly"})


def prism_mesh(n):
    """Unit cube of n^3 hexes, each split into two triangular prisms along the x-y diagonal:
    triangular and quadrilateral faces, 5-face / 6-point cells."""
    g = np.linspace(0.0, 1.0, n + 1)
    P = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)

    def pid(i, j, k):
        return (i * (n + 1) + j) * (n + 1) + k

    cells = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                a, b, c, d = pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)
                e, f, g_, h = pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)
                # prism 1: triangle a-b-c (bottom) / e-f-g (top);  prism 2: a-c-d / e-g-h
                cells.append([[a, c, b], [e, f, g_], [a, b, f, e], [b, c, g_, f], [c, a, e, g_]])
                cells.append([[a, d, c], [e, g_, h], [a, c, g_, e], [c, d, h, g_], [d, a, e, h]])
    return build_polymesh(P, cells)


def kelvin_mesh(n):
    """2 n^3 truncated octahedra (Kelvin cells, the Voronoi cells of a BCC lattice): 14 faces (6 squares, 8 hexagons)
    and 24 points per cell -- the cell population of a polyDualMesh (BASELINE.json configs[2]: ~14 faces per cell,
    ~5 vertices per face).  Whole cells only, so the outer boundary is faceted; it fits inside the unit cube."""
    import itertools
    # one cell about the origin in integer coordinates: vertices = all permutations of (0, +-1, +-2)
    verts = sorted({p for perm in itertools.permutations((0, 1, 2)) for p in
                    itertools.product(*[(v, -v) if v else (0,) for v in perm])})
    V = np.array(verts, dtype=np.int64)
    normals = [np.array(e) * sgn for e in ((1, 0, 0), (0, 1, 0), (0, 0, 1)) for sgn in (1, -1)]
    normals += [np.array(sg) for sg in itertools.product((1, -1), repeat=3)]
    faces = []
    for nrm in normals:
        d = 2 if np.abs(nrm).sum() == 1 else 3
        on = np.nonzero(V @ nrm == d)[0]
        fc = V[on].mean(axis=0)
        u = (V[on[0]] - fc).astype(np.float64)
        u /= np.linalg.norm(u)
        w = np.cross(nrm / np.linalg.norm(nrm), u)
        ang = np.arctan2((V[on] - fc) @ w, (V[on] - fc) @ u)
        faces.append([int(i) for i in on[np.argsort(ang)]])       # counter-clockwise about the outward normal
    ids, pts, cells = {}, [], []
    centres = [(4 * i + o, 4 * j + o, 4 * k + o) for o in (0, 2) for k in range(n) for j in range(n) for i in range(n)]
    for c in centres:
        loc = []
        for v in verts:
            key = (v[0] + c[0], v[1] + c[1], v[2] + c[2])
            if key not in ids:
                ids[key] = len(pts)
                pts.append(key)
            loc.append(ids[key])
        cells.append([[loc[i] for i in f] for f in faces])
    P = (np.array(pts, dtype=np.float64) + 2.0) / (4.0 * n + 2.0)
    m = build_polymesh(P, cells)
    m.meta["kind"] = "kelvin"
    m.meta["cell_volume"] = 32.0 / (4.0 * n + 2.0) ** 3
    return m


def kelvin_mesh_fast(n):
    """kelvin_mesh(n) built with numpy (2 n^3 cells in seconds instead of Python loops): the same cells and faces
    with a different -- but deterministic and OpenFOAM-conforming -- point numbering (ascending lattice key) and the
    internal faces in upper-triangular order.  n = 95 gives the 1.7 M polyhedral cells of BASELINE.json configs[2]."""
    import itertools
    verts = sorted({p for perm in itertools.permutations((0, 1, 2)) for p in
                    itertools.product(*[(v, -v) if v else (0,) for v in perm])})
    V = np.array(verts, dtype=np.int64)
    normals = [np.array(e) * sgn for e in ((1, 0, 0), (0, 1, 0), (0, 0, 1)) for sgn in (1, -1)]
    normals += [np.array(sg) for sg in itertools.product((1, -1), repeat=3)]
    faces, fcentre = [], []
    for nrm in normals:
        d = 2 if np.abs(nrm).sum() == 1 else 3
        on = np.nonzero(V @ nrm == d)[0]
        fc = V[on].mean(axis=0)
        u = (V[on[0]] - fc).astype(np.float64)
        u /= np.linalg.norm(u)
        w = np.cross(nrm / np.linalg.norm(nrm), u)
        ang = np.arctan2((V[on] - fc) @ w, (V[on] - fc) @ u)
        faces.append(np.array([int(i) for i in on[np.argsort(ang)]]))
        fcentre.append(np.rint(fc).astype(np.int64))          # integer lattice point (2,0,0)- or (1,1,1)-type
    g = np.arange(n, dtype=np.int64)
    K, J, I = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([I.reshape(-1), J.reshape(-1), K.reshape(-1)], axis=1) * 4
    centres = np.concatenate([base, base + 2])                  # same cell order as kelvin_mesh
    nC = centres.shape[0]
    M = 4 * n + 8

    def key(xyz):                                               # lattice coordinates (>= -2) -> unique integer
        q = xyz + 2
        return q[..., 0] + M * (q[..., 1] + M * q[..., 2])

    # points
    pk = key(centres[:, None, :] + V[None, :, :])               # [nC, 24]
    uniq, inv = np.unique(pk.reshape(-1), return_inverse=True)
    cell_pts = inv.reshape(nC, 24).astype(np.int32)
    z, rem = np.divmod(uniq, M * M)
    y, x = np.divmod(rem, M)
    P = (np.stack([x, y, z], axis=1).astype(np.float64) - 2.0 + 2.0) / (4.0 * n + 2.0)
    # faces: (cell, local face) records, matched through the lattice key of the face centre
    recs = []
    for lf, (fl, fc) in enumerate(zip(faces, fcentre)):
        recs.append((key(centres + fc[None, :]), np.arange(nC, dtype=np.int64), np.full(nC, lf, np.int64)))
    fkey = np.concatenate([r[0] for r in recs])
    fcell = np.concatenate([r[1] for r in recs])
    floc = np.concatenate([r[2] for r in recs])
    order = np.lexsort((fcell, fkey))                           # by key, lower cell first
    fkey, fcell, floc = fkey[order], fcell[order], floc[order]
    same_next = np.concatenate([fkey[1:] == fkey[:-1], [False]])
    same_prev = np.concatenate([[False], fkey[1:] == fkey[:-1]])
    int_first = np.nonzero(same_next)[0]                        # owner side of an internal face
    bnd = np.nonzero(~same_next & ~same_prev)[0]
    own_i, nei_i, loc_i = fcell[int_first], fcell[int_first + 1], floc[int_first]
    o2 = np.lexsort((nei_i, own_i))                             # upper-triangular order
    own_i, nei_i, loc_i = own_i[o2], nei_i[o2], loc_i[o2]
    own_b, loc_b = fcell[bnd], floc[bnd]
    o3 = np.argsort(own_b, kind="stable")
    own_b, loc_b = own_b[o3], loc_b[o3]
    owner = np.concatenate([own_i, own_b])
    loc = np.concatenate([loc_i, loc_b])
    nv = np.array([len(f) for f in faces], dtype=np.int64)[loc]
    face_offsets = np.concatenate([[0], np.cumsum(nv)])
    face_points = np.empty(int(face_offsets[-1]), dtype=np.int32)
    for lf, fl in enumerate(faces):
        sel = np.nonzero(loc == lf)[0]
        if sel.size:
            dst = face_offsets[sel][:, None] + np.arange(len(fl))[None, :]
            face_points[dst.reshape(-1)] = cell_pts[owner[sel]][:, fl].reshape(-1)
    nIF = own_i.shape[0]
    m = PolyMesh(points=P, face_offsets=face_offsets.astype(np.int32), face_points=face_points, owner=owner.astype(np.int32),
                 neighbour=nei_i.astype(np.int32), patches=[Patch("walls", nIF, int(own_b.shape[0]))], n_cells=nC,
                 meta={"kind": "kelvin", "cell_volume": 32.0 / (4.0 * n + 2.0) ** 3, "n": n})
    return m


def refined_interface_mesh(n):
    """2:1 refinement interface (what dynamicRefineFvMesh produces in the reference's AMR cases): the half
    x < 0.5 is meshed with n^3/2 coarse hexes, the half x > 0.5 with 8x finer ones; the coarse cells on the
    interface are genuine polyhedra (9 faces, 13 points, four coplanar quads on one side)."""
    assert n % 2 == 0
    H, hf = 1.0 / n, 0.5 / n
    pts, index = [], {}

    def P(x, y, z):
        key = (int(round(x / hf)), int(round(y / hf)), int(round(z / hf)))
        if key not in index:
            index[key] = len(pts)
            pts.append((key[0] * hf, key[1] * hf, key[2] * hf))
        return index[key]

    def hex_faces(x0, y0, z0, s, split_xplus=False):
        x1, y1, z1 = x0 + s, y0 + s, z0 + s
        f = [[P(x0, y0, z0), P(x0, y0, z1), P(x0, y1, z1), P(x0, y1, z0)],      # x-
             [P(x0, y0, z0), P(x1, y0, z0), P(x1, y0, z1), P(x0, y0, z1)],      # y-
             [P(x0, y1, z0), P(x0, y1, z1), P(x1, y1, z1), P(x1, y1, z0)],      # y+
             [P(x0, y0, z0), P(x0, y1, z0), P(x1, y1, z0), P(x1, y0, z0)],      # z-
             [P(x0, y0, z1), P(x1, y0, z1), P(x1, y1, z1), P(x0, y1, z1)]]      # z+
        if not split_xplus:
            f.append([P(x1, y0, z0), P(x1, y1, z0), P(x1, y1, z1), P(x1, y0, z1)])
        else:
            ym, zm = y0 + s / 2, z0 + s / 2
            for (ya, yb) in ((y0, ym), (ym, y1)):
                for (za, zb) in ((z0, zm), (zm, z1)):
                    f.append([P(x1, ya, za), P(x1, yb, za), P(x1, yb, zb), P(x1, ya, zb)])
            # the neighbouring coarse faces sharing the split edges need the mid-edge points too
            f[1] = [P(x0, y0, z0), P(x1, y0, z0), P(x1, y0, zm), P(x1, y0, z1), P(x0, y0, z1)]
            f[2] = [P(x0, y1, z0), P(x0, y1, z1), P(x1, y1, z1), P(x1, y1, zm), P(x1, y1, z0)]
            f[3] = [P(x0, y0, z0), P(x0, y1, z0), P(x1, y1, z0), P(x1, ym, z0), P(x1, y0, z0)]
            f[4] = [P(x0, y0, z1), P(x1, y0, z1), P(x1, ym, z1), P(x1, y1, z1), P(x0, y1, z1)]
        return f

    cells = []
    nc = n // 2
    for k in range(n):
        for j in range(n):
            for i in range(nc):
                cells.append(hex_faces(i * H, j * H, k * H, H, split_xplus=(i == nc - 1)))
    for k in range(2 * n):
        for j in range(2 * n):
            for i in range(n, 2 * n):
                cells.append(hex_faces(i * hf, j * hf, k * hf, hf))
    return build_polymesh(np.array(pts), cells)
