"""Run the reference's own case directories on the device.

The reference's test case is driven by `Allrun` (tutorials/test/plicVofAdvectionFoam/Allrun:8-22):

    blockMesh; renumberMesh -overwrite; ln -s exactSolutions/0/alpha.water.exact 0/alpha.water
    plicVofAdvectionFoam; calcVofAdvectionErrors

This module is those steps without OpenFOAM, on the files OpenFOAM reads and writes:

    block_mesh()                 system/blockMeshDict (one axis-aligned hex block, uniform grading)
    renumber_mesh()              renumberMesh's default Cuthill-McKee cell order + upper-triangular faces
    FoamCase                     controlDict / fvSolution solvers."alpha.*" / 0/alpha.water / constant/polyMesh
    run_plic_vof_advection()     the plicVofAdvectionFoam time loop (plicVof.H:13-57, updateU.H, setDeltaT.H)
                                 around SolveVofEqu, writing <time>/alpha.water in OpenFOAM binary format
    calc_vof_advection_errors()  E_v, alphaMin, 1-alphaMax, E_s of calcVofAdvectionErrors/updateErrors.H
                                 (E_p needs OpenFOAM's isoSurface and is not computed)

    python -m geometricvofext_b200.foamcase run <caseDir> [--end-time T] [--renumber] [--exact DIR]

Host-side plumbing; the step itself is SolveVofEqu (the CUDA library unless another `lib` is passed).
"""
import argparse
import os
import shutil
import sys
from collections import deque

import numpy as np

from . import capi, fields, foamfile
from .mesh import Patch, PolyMesh, hex_block
from .solver import SolveVofEqu

EXACT_INITIAL_VOL = 0.0141366879746714   # calcVofAdvectionErrors.C: exactInitialVol

# block-local vertex positions of the six sides of `hex (v0 .. v7)`: (axis, side) -> vertex slots
_SIDE_SLOTS = {(0, 0): {0, 3, 7, 4}, (0, 1): {1, 2, 6, 5}, (1, 0): {0, 1, 5, 4}, (1, 1): {3, 2, 6, 7},
               (2, 0): {0, 1, 2, 3}, (2, 1): {4, 5, 6, 7}}


def block_mesh(block_mesh_dict):
    """blockMeshDict (path or parsed FoamDict) -> PolyMesh, for one axis-aligned hex block with uniform grading.

    Cells in blockMesh's natural order c = i + Nx (j + Ny k), internal faces upper-triangular, patches in the
    order of the `boundary` list, unlisted sides collected in `defaultFaces` (type empty) as blockMesh does."""
    d = foamfile.read_dict(block_mesh_dict) if isinstance(block_mesh_dict, (str, os.PathLike)) else block_mesh_dict
    scale = float(d.get("scale", d.get("convertToMeters", 1.0)))
    verts = np.asarray(d["vertices"], dtype=np.float64) * scale
    blocks = d["blocks"]
    if len([b for b in blocks if b == "hex"]) != 1:
        raise foamfile.FoamFormatError("blockMeshDict: exactly one hex block is supported")
    i = blocks.index("hex")
    hexv, ncells = [int(v) for v in blocks[i + 1]], [int(round(float(v))) for v in blocks[i + 2]]
    grading = blocks[i + 4] if len(blocks) > i + 4 else [1, 1, 1]
    if any(float(g) != 1.0 for g in np.asarray(grading, dtype=np.float64).reshape(-1)):
        raise foamfile.FoamFormatError("blockMeshDict: only simpleGrading (1 1 1) is supported")
    if d.get("edges"):
        raise foamfile.FoamFormatError("blockMeshDict: curved edges are not supported")
    p = verts[hexv]
    origin = p[0]
    ex, ey, ez = p[1] - p[0], p[3] - p[0], p[4] - p[0]
    for a, e in enumerate((ex, ey, ez)):
        off = np.delete(e, a)
        if np.abs(off).max() > 1e-12 * max(1.0, np.abs(e).max()) or e[a] <= 0:
            raise foamfile.FoamFormatError("blockMeshDict: the block must be axis aligned and right handed")
    length = (ex[0], ey[1], ez[2])
    expect = origin + np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]]) * length
    if np.abs(p - expect).max() > 1e-12 * max(1.0, np.abs(p).max()):
        raise foamfile.FoamFormatError("blockMeshDict: the block is not a box")
    m = hex_block(tuple(ncells), length=length, origin=tuple(origin))
    side_range = {(ax, sd): (pt.start, pt.size) for pt, (_, ax, sd) in zip(m.patches, _hex_patches())}
    # regroup the six sides into the dictionary's patches
    slot_of = {v: s for s, v in enumerate(hexv)}
    new_patches, order, used, ptypes = [], [], set(), {}
    start = m.n_internal_faces
    for item in d.get("boundary", []):
        name, pd = item
        sides = []
        for quad in pd.get("faces", []):
            slots = {slot_of[int(v)] for v in quad}
            key = next((k for k, s in _SIDE_SLOTS.items() if s == slots), None)
            if key is None or key in used:
                raise foamfile.FoamFormatError("blockMeshDict: patch %s: face %s is not a (free) block side" % (name, quad))
            used.add(key)
            sides.append(key)
        size = 0
        for key in sides:
            s0, n = side_range[key]
            order.append(np.arange(s0, s0 + n))
            size += n
        ptype = str(pd.get("type", "patch"))
        ptypes[name] = ptype
        new_patches.append(Patch(name, start, size, kind=capi.PATCH_EMPTY if ptype == "empty" else capi.PATCH_GENERIC))
        start += size
    rest = [k for k in _SIDE_SLOTS if k not in used]
    if rest:
        size = 0
        for key in rest:
            s0, n = side_range[key]
            order.append(np.arange(s0, s0 + n))
            size += n
        new_patches.append(Patch("defaultFaces", start, size, kind=capi.PATCH_EMPTY))
        ptypes["defaultFaces"] = "empty"
    order = np.concatenate(order) if order else np.zeros(0, np.int64)
    nIF = m.n_internal_faces
    fp = m.face_points.reshape(-1, 4)
    m.face_points = np.concatenate([fp[:nIF], fp[order]]).reshape(-1).astype(np.int32)
    m.owner = np.concatenate([m.owner[:nIF], m.owner[order]]).astype(np.int32)
    m.patches = new_patches
    m.meta["patch_types"] = ptypes
    return m


def _hex_patches():
    from .mesh import _HEX_PATCHES
    return _HEX_PATCHES


def cuthill_mckee(mesh):
    """OpenFOAM's default renumberMesh order (CuthillMcKee = bandCompression of cellCells): newToOld.

    cellCells of a cell in ascending internal-face order; start from the lowest-index unvisited cell of minimum
    neighbour count; breadth first, the unvisited neighbours of a cell appended in ascending neighbour count (stable)."""
    nC, own, nei = mesh.n_cells, mesh.owner[:mesh.n_internal_faces].astype(np.int64), mesh.neighbour.astype(np.int64)
    nIF = own.shape[0]
    # CSR of cellCells in ascending face order
    cell = np.concatenate([own, nei])
    other = np.concatenate([nei, own])
    face = np.concatenate([np.arange(nIF), np.arange(nIF)])
    o = np.lexsort((face, cell))
    cell, other = cell[o], other[o]
    deg = np.bincount(cell, minlength=nC)
    off = np.zeros(nC + 1, dtype=np.int64)
    np.cumsum(deg, out=off[1:])
    visited = np.zeros(nC, dtype=bool)
    new_to_old = np.empty(nC, dtype=np.int64)
    n_done = 0
    by_deg = np.lexsort((np.arange(nC), deg))   # candidates for new components: min degree, then lowest index
    cand = 0
    degl, offl, otherl = deg.tolist(), off.tolist(), other.tolist()
    while n_done < nC:
        while visited[by_deg[cand]]:
            cand += 1
        q = deque([int(by_deg[cand])])
        while q:
            c = q.popleft()
            if visited[c]:
                continue
            visited[c] = True
            new_to_old[n_done] = c
            n_done += 1
            nb = [x for x in otherl[offl[c]:offl[c + 1]] if not visited[x]]
            nb.sort(key=degl.__getitem__)
            q.extend(nb)
    return new_to_old


def renumber_mesh(mesh, new_to_old=None):
    """renumberMesh -overwrite: cells in `new_to_old` order (default: cuthill_mckee), internal faces re-ordered
    upper-triangular (flipped where the new owner would exceed the new neighbour), boundary faces and points kept.
    Returns (new mesh, new_to_old); fields map as new[i] = old[new_to_old[i]]."""
    if new_to_old is None:
        new_to_old = cuthill_mckee(mesh)
    nC, nIF, nF = mesh.n_cells, mesh.n_internal_faces, mesh.n_faces
    old_to_new = np.empty(nC, dtype=np.int64)
    old_to_new[new_to_old] = np.arange(nC)
    o, n = old_to_new[mesh.owner[:nIF]], old_to_new[mesh.neighbour]
    flip = o > n
    new_own, new_nei = np.minimum(o, n), np.maximum(o, n)
    forder = np.lexsort((new_nei, new_own))
    off = mesh.face_offsets.astype(np.int64)
    sizes = np.diff(off)
    face_order = np.concatenate([forder, np.arange(nIF, nF)])
    flip_all = np.concatenate([flip, np.zeros(nF - nIF, dtype=bool)])
    new_off = np.zeros(nF + 1, dtype=np.int64)
    np.cumsum(sizes[face_order], out=new_off[1:])
    lab = np.empty(int(new_off[-1]), dtype=np.int32)
    fp = mesh.face_points
    if np.all(sizes == sizes[0]):       # fast path (all quads, all triangles, ...)
        k = int(sizes[0])
        rows = fp.reshape(nF, k)[face_order].copy()
        fl = flip_all[face_order]
        rows[fl, 1:] = rows[fl, :0:-1]  # face::reverseFace keeps the first vertex
        lab[:] = rows.reshape(-1)
    else:
        for i, f in enumerate(face_order):
            v = fp[off[f]:off[f + 1]]
            if flip_all[f]:
                v = np.concatenate([v[:1], v[:0:-1]])
            lab[new_off[i]:new_off[i + 1]] = v
    owner = np.concatenate([new_own[forder], old_to_new[mesh.owner[nIF:]]]).astype(np.int32)
    m = PolyMesh(points=mesh.points, face_offsets=new_off.astype(np.int32), face_points=lab, owner=owner,
                 neighbour=new_nei[forder].astype(np.int32), patches=[Patch(**vars(p)) for p in mesh.patches], n_cells=nC,
                 meta=dict(mesh.meta, kind="renumbered", cell_map=new_to_old))
    return m, new_to_old


class FoamCase:
    """An OpenFOAM case directory, as far as the volume-fraction transport step needs it."""

    def __init__(self, case_dir, alpha_name="alpha.water"):
        self.dir, self.alpha_name = os.path.abspath(case_dir), alpha_name
        self.control_dict = foamfile.read_dict(os.path.join(self.dir, "system", "controlDict"))
        self.fv_solution = foamfile.read_dict(os.path.join(self.dir, "system", "fvSolution"))

    def alpha_controls(self):
        """fvSolution solvers."alpha.*" (solveVofEqu.C:68: mesh.solverDict(alpha1.name()))."""
        d = self.fv_solution["solvers"].lookup(self.alpha_name)
        if d is None:
            raise foamfile.FoamFormatError("fvSolution: no solvers entry matches %s" % self.alpha_name)
        ctl = {k: v for k, v in d.items() if not isinstance(v, (dict, list, tuple))}
        # orientationMethod alphaGrad evaluates fvc::grad(alpha1, "grad(alpha1)") (reconstruction.C:78) with the case's gradient
        # scheme: system/fvSchemes gradSchemes { grad(alpha1) ...; } or its default entry
        if str(ctl.get("orientationMethod", "")) == "alphaGrad":
            sch = self.grad_alpha_scheme()
            if sch:
                ctl["gradSchemes"] = sch
        return ctl

    def grad_alpha_scheme(self):
        """The gradSchemes entry that applies to grad(alpha1) ('Gauss linear', 'Gauss pointLinear', ...) or None."""
        path = os.path.join(self.dir, "system", "fvSchemes")
        if not os.path.isfile(path):
            return None
        g = foamfile.read_dict(path).get("gradSchemes")
        if g is None:
            return None
        v = g.lookup("grad(alpha1)") if hasattr(g, "lookup") else g.get("grad(alpha1)")
        if v is None:   # this reader splits `grad(alpha1) Gauss pointLinear;` into the keyword grad and the list (alpha1)
            w = g.get("grad")
            if isinstance(w, (list, tuple)) and len(w) > 1 and list(w[0]) == ["alpha1"]:
                v = list(w[1:])
        if v is None:
            v = g.get("default")
        if v is None or str(v) == "none":
            return None
        return " ".join(str(x) for x in v) if isinstance(v, (list, tuple)) else str(v)

    def has_poly_mesh(self):
        return os.path.isfile(os.path.join(self.dir, "constant", "polyMesh", "faces"))

    def mesh(self, renumber=False):
        if self.has_poly_mesh():
            return foamfile.read_polymesh(self.dir)
        m = block_mesh(os.path.join(self.dir, "system", "blockMeshDict"))
        if renumber:
            m, _ = renumber_mesh(m)
        return m

    def time_dirs(self):
        out = []
        for n in os.listdir(self.dir):
            try:
                out.append((float(n), n))
            except ValueError:
                pass
        return sorted(out)

    def start_dir(self):
        for n in ("0", "0.orig"):
            if os.path.isfile(os.path.join(self.dir, n, self.alpha_name)):
                return os.path.join(self.dir, n)
        raise FileNotFoundError("no 0/%s in %s" % (self.alpha_name, self.dir))

    def read_alpha(self, mesh, time_dir=None):
        f = foamfile.read_field(os.path.join(time_dir or self.start_dir(), self.alpha_name))
        foamfile.apply_alpha_boundary(mesh, f)
        return f.internal_array(mesh.n_cells), f

    def write_alpha(self, mesh, time_name, alpha, template=None, fmt=None, name=None):
        fmt = fmt or str(self.control_dict.get("writeFormat", "binary"))
        bnd = {}
        for p in mesh.patches:
            if p.kind == capi.PATCH_EMPTY:
                bnd[p.name] = {"type": "empty"}
            elif p.kind == capi.PATCH_PROCESSOR:
                bnd[p.name] = {"type": "processor", "value": 0.0}
            elif p.alpha_bc == capi.BC_FIXED_VALUE:
                bnd[p.name] = {"type": "fixedValue", "value": p.alpha_value}
            elif p.alpha_bc == capi.BC_INLET_OUTLET:
                bnd[p.name] = {"type": "inletOutlet", "inletValue": p.alpha_value, "value": p.alpha_value}
            else:
                bnd[p.name] = {"type": "zeroGradient"}
        path = os.path.join(self.dir, time_name, name or self.alpha_name)
        foamfile.write_field(path, "volScalarField", name or self.alpha_name, np.asarray(alpha), bnd, fmt=fmt, location=time_name)
        return path


def plic_surface_functions(control_dict):
    """(function object name, surface name) of every `type surfaces` function object that samples a `plicSurface`
    (tutorials/test/plicVofAdvectionFoam/system/controlDict:53-72)."""
    out = []
    fns = control_dict.get("functions") or {}
    for fname, fd in fns.items():
        if not isinstance(fd, dict) or str(fd.get("type")) != "surfaces":
            continue
        for item in fd.get("surfaces", []) or []:
            if isinstance(item, tuple) and isinstance(item[1], dict) and str(item[1].get("type")) == "plicSurface":
                out.append((fname, item[0]))
    return out


def _write_plic_fields(case, mesh, s, time_name):
    """writePlicFields true (reconstruction.C:520-570): interfaceN [1/m], interfaceD [m], interfaceC [m], interfaceS [m^2]
    are AUTO_WRITE fields of the reconstruction; `calculated` patches with the zero-gradient value."""
    fmt = str(case.control_dict.get("writeFormat", "binary"))
    nIF = mesh.n_internal_faces
    for name, which, cls, dims in (("interfaceN", capi.F_INTERFACE_N, "volVectorField", (0, -1, 0, 0, 0, 0, 0)),
                                   ("interfaceD", capi.F_INTERFACE_D, "volScalarField", (0, 1, 0, 0, 0, 0, 0)),
                                   ("interfaceC", capi.F_INTERFACE_C, "volVectorField", (0, 1, 0, 0, 0, 0, 0)),
                                   ("interfaceS", capi.F_INTERFACE_S, "volVectorField", (0, 2, 0, 0, 0, 0, 0))):
        f = s.field(which)
        bnd = {}
        for p in mesh.patches:
            if p.kind == capi.PATCH_EMPTY:
                bnd[p.name] = {"type": "empty"}
            else:
                bnd[p.name] = {"type": "calculated", "value": np.ascontiguousarray(f[mesh.owner[p.start:p.start + p.size]])}
        foamfile.write_field(os.path.join(case.dir, time_name, name), cls, name, f, bnd, dimensions=dims, fmt=fmt, location=time_name)


def _write_surfaces(case, s, time_name, surfaces):
    if not surfaces:
        return
    pts, off, cells = s.interface()
    a = s.alpha()
    for fname, sname in surfaces:
        foamfile.write_vtk_polydata(os.path.join(case.dir, "postProcessing", fname, time_name, sname + ".vtk"), pts, off,
                                    {"cellIds": cells, case.alpha_name: a[cells]})


def _time_name(t, precision=6):
    """Time::timeName with timeFormat general, timePrecision 6."""
    s = "%.*g" % (precision, t)
    return "0" if float(s) == 0 else s


def run_plic_vof_advection(case, lib=None, end_time=None, renumber=False, alpha0=None, write=True, log=None):
    """The plicVofAdvectionFoam application on a case directory.  Returns a dict with the run summary."""
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    log = log or (lambda *a: None)
    cd = case.control_dict
    mesh = case.mesh(renumber=renumber)
    if alpha0 is None:
        alpha0, _ = case.read_alpha(mesh)
    else:
        try:
            case.read_alpha(mesh)          # boundary conditions still come from the case
        except FileNotFoundError:
            pass
    controls = case.alpha_controls()
    s = SolveVofEqu(mesh, controls, lib=lib)
    s.setAlpha(np.ascontiguousarray(alpha0, dtype=np.float64))
    adjust = str(cd.get("adjustTimeStep", "no")).lower() in ("yes", "true", "on", "1")
    drv = fields.AdvectionDriver(
        s, period=float(controls.get("period", 0.0)), max_co=float(cd.get("maxCo", 1.0)),
        max_alpha_co=float(cd.get("maxAlphaCo", 1.0)), max_delta_t=float(cd.get("maxDeltaT", 1e30)),
        delta_t0=float(cd.get("deltaT", 1e-3)), fixed_dt=None if adjust else float(cd.get("deltaT", 1e-3)))
    drv.t = drv.start_time = float(cd.get("startTime", 0.0))
    t_end = float(end_time if end_time is not None else cd.get("endTime"))
    w_int = float(cd.get("writeInterval", t_end))
    by_time = str(cd.get("writeControl", "adjustableRunTime")) in ("adjustableRunTime", "runTime")
    adjustable = adjust and str(cd.get("writeControl", "adjustableRunTime")) == "adjustableRunTime"
    if adjustable:
        drv.write_interval = w_int        # Time::adjustDeltaT: equal steps up to each write time
    s.reconstruct()                       # plicVof.H:8-9: interface at the initial time, function objects executed
    surfaces = plic_surface_functions(cd) if write else []
    write_plic = write and str(controls.get("writePlicFields", "false")).lower() in ("true", "yes", "on", "1")
    _write_surfaces(case, s, _time_name(drv.t), surfaces)
    written, vols = [], []
    next_write = drv.t + w_int if by_time else None
    while (drv.running(t_end) if adjustable else drv.t < t_end - 1e-12):
        stop = min(t_end, next_write) if by_time else t_end
        drv.step(end_time=stop)
        s.synchronize()          # device capacity flags (the reference would FatalError) surface here, once per step
        vols.append(s.volume())
        due = (drv.write_now or not drv.running(t_end)) if adjustable else ((by_time and drv.t >= next_write - 1e-12) or (not by_time and drv.steps % max(1, int(w_int)) == 0) or drv.t >= t_end - 1e-12)
        if due:
            name = _time_name(drv.t)
            if write:
                case.write_alpha(mesh, name, s.alpha())
                if write_plic:
                    _write_plic_fields(case, mesh, s, name)
                _write_surfaces(case, s, name, surfaces)   # the polygons of the step's reconstruct()
            written.append(name)
            log("Time = %s  steps %d  Phase-1 volume = %.15g" % (name, drv.steps, vols[-1]))
            if by_time:
                next_write += w_int
    out = {"steps": drv.steps, "end_time": drv.t, "written": written, "volume": vols[-1] if vols else s.volume(),
           "n_cells": mesh.n_cells, "reconstruction_time": s.reconstructionTime(), "advection_time": s.advectionTime(),
           "alpha": s.alpha(), "mesh": mesh}
    s.close()
    return out


# ---- decomposed cases: processor*/ directories --------------------------------------------------------------------------
def read_cell_proc_addressing(case, n_procs=None):
    """cell -> rank map of a decomposed case from processor*/constant/polyMesh/cellProcAddressing (the file of processor r
    lists, for each of its cells in local order, the cell's label in the undecomposed mesh)."""
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    procs = sorted(int(d[len("processor"):]) for d in os.listdir(case.dir)
                   if d.startswith("processor") and d[len("processor"):].isdigit())
    if n_procs is not None:
        procs = [r for r in procs if r < n_procs]
    if not procs or procs != list(range(len(procs))):
        raise foamfile.FoamFormatError("%s: processor directories %s are not 0..N-1" % (case.dir, procs))
    addr = [foamfile.read_labels(os.path.join(case.dir, "processor%d" % r, "constant", "polyMesh", "cellProcAddressing")) for r in procs]
    n = sum(len(a) for a in addr)
    cell_rank = np.full(n, -1, np.int32)
    for r, a in enumerate(addr):
        cell_rank[a] = r
    if (cell_rank < 0).any():
        raise foamfile.FoamFormatError("cellProcAddressing files do not cover every cell exactly once")
    return cell_rank, addr


def decompose_case(case, n_procs, cell_rank=None, weights=None, lib=None, fmt=None):
    """What this driver needs of `decomposePar`: processorR/constant/polyMesh/cellProcAddressing and processorR/0/alpha.water for
    R = 0..n_procs-1.  The cell -> rank map is given (a scotch map read elsewhere) or comes from the library's weighted recursive
    bisection (svof_partition_rcb).  The processor meshes themselves are not written: the device decomposition
    (svof_decompose: ghost layers instead of processor patches) is cut from the undecomposed mesh by every rank."""
    from . import multigpu as mg
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    mesh = case.mesh()
    if cell_rank is None:
        cell_rank = mg.partition_rcb(mesh, n_procs, weights=weights, lib=lib)
    cell_rank = np.asarray(cell_rank, np.int32)
    alpha0, _ = case.read_alpha(mesh)
    fmt = fmt or str(case.control_dict.get("writeFormat", "binary"))
    for r in range(n_procs):
        own = np.nonzero(cell_rank == r)[0].astype(np.int32)       # ascending, as decomposePar numbers a processor's cells
        pdir = os.path.join(case.dir, "processor%d" % r)
        foamfile.write_labels(os.path.join(pdir, "constant", "polyMesh", "cellProcAddressing"), own, "cellProcAddressing",
                              "constant/polyMesh", fmt=fmt)
        foamfile.write_field(os.path.join(pdir, "0", case.alpha_name), "volScalarField", case.alpha_name, alpha0[own], {},
                             fmt=fmt, location="0")
    return cell_rank


class _RankSolver:
    """The members fields.AdvectionDriver uses, on one rank of a decomposed run (multigpu.DecomposedSolveVofEqu)."""

    def __init__(self, ds):
        self.ds, self.s = ds, ds.s
        self.mesh, self.nC, self.nF, self.nIF, self.nBF = ds.s.mesh, ds.s.nC, ds.s.nF, ds.s.nIF, ds.s.nBF

    def field(self, which):
        return self.s.field(which)

    def alpha(self):
        return self.s.alpha()

    def setPhi(self, phi):
        self.ds.setPhi(phi)

    def setU(self, U, Ub=None):
        self.ds.setU(U, Ub)

    def reconstruct(self):
        self.ds.reconstruct()

    def advect(self, dt):
        self.ds.advect(dt)


def run_plic_vof_advection_decomposed(case, rank, world, lib=None, device=None, end_time=None, write=True, log=None):
    """plicVofAdvectionFoam -parallel on a decomposed case, one call per rank (torch.distributed initialised by the caller:
    gloo for the CPU engine, nccl on GPUs).  Rank r reads processor<r>/constant/polyMesh/cellProcAddressing and
    processor<r>/0/alpha.water, cuts its part (owned cells + ghost layers) out of the undecomposed mesh, runs the time loop
    with globally reduced Courant numbers and writes processor<r>/<time>/alpha.water for its own cells in processor order --
    what reconstructPar reads."""
    import torch
    import torch.distributed as dist
    from . import multigpu as mg
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    log = log or (lambda *a: None)
    cd = case.control_dict
    mesh = case.mesh()
    cell_rank, addr = read_cell_proc_addressing(case, world)
    if len(addr) != world:
        raise foamfile.FoamFormatError("%d processor directories for %d ranks" % (len(addr), world))
    try:
        case.read_alpha(mesh)            # the patch conditions come from the undecomposed 0/alpha.water
    except FileNotFoundError:
        pass
    controls = case.alpha_controls()
    sub, maps = mg.decompose(mesh, cell_rank, rank, mg.default_layers(controls))   # host utility of the product library
    ds = mg.DecomposedSolveVofEqu(sub, maps, controls, rank, world, lib=lib, device=device)
    pdir = os.path.join(case.dir, "processor%d" % rank)
    f0 = foamfile.read_field(os.path.join(pdir, "0", case.alpha_name))
    mine = np.asarray(addr[rank], np.int64)                     # processor-local order -> global label
    a_proc = f0.internal_array(len(mine))
    cell_global = np.asarray(maps["cell_global"], np.int64)
    pos = np.searchsorted(cell_global, mine)                     # sub-mesh numbering preserves the global order
    assert np.array_equal(cell_global[pos], mine), "cellProcAddressing does not match the rank's owned cells"
    a_local = np.zeros(sub.n_cells)
    a_local[pos] = a_proc
    ds.setAlpha(a_local)
    ds.exchange_alpha()                                          # ghost values from their owners

    def gmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64)
        if device is not None and dist.get_backend() == "nccl":
            t = t.cuda(device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rs = _RankSolver(ds)
    adjust = str(cd.get("adjustTimeStep", "no")).lower() in ("yes", "true", "on", "1")
    drv = fields.AdvectionDriver(
        rs, period=float(controls.get("period", 0.0)), max_co=float(cd.get("maxCo", 1.0)),
        max_alpha_co=float(cd.get("maxAlphaCo", 1.0)), max_delta_t=float(cd.get("maxDeltaT", 1e30)),
        delta_t0=float(cd.get("deltaT", 1e-3)), fixed_dt=None if adjust else float(cd.get("deltaT", 1e-3)), reduce_max=gmax)
    drv.t = drv.start_time = float(cd.get("startTime", 0.0))
    t_end = float(end_time if end_time is not None else cd.get("endTime"))
    w_int = float(cd.get("writeInterval", t_end))
    if not (adjust and str(cd.get("writeControl", "adjustableRunTime")) == "adjustableRunTime"):
        raise NotImplementedError("the decomposed driver follows writeControl adjustableRunTime with adjustTimeStep (the reference's cases)")
    drv.write_interval = w_int
    ds.reconstruct()
    fmt = str(cd.get("writeFormat", "binary"))
    written = []
    while drv.running(t_end):
        drv.step(end_time=t_end)
        ds.s.synchronize()
        if drv.write_now or not drv.running(t_end):
            name = _time_name(drv.t)
            vol = ds.volume()
            if write:
                foamfile.write_field(os.path.join(pdir, name, case.alpha_name), "volScalarField", case.alpha_name,
                                     ds.s.alpha()[pos], {}, fmt=fmt, location=name)
            written.append(name)
            if rank == 0:
                log("Time = %s  steps %d  Phase-1 volume = %.15g" % (name, drv.steps, vol))
    out = {"steps": drv.steps, "end_time": drv.t, "written": written, "n_owned": len(mine), "volume": ds.volume()}
    ds.close()
    return out


def reconstruct_par(case, time_name, n_procs=None):
    """reconstructPar for alpha: the undecomposed field of <time> from processor*/<time>/alpha.water."""
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    cell_rank, addr = read_cell_proc_addressing(case, n_procs)
    out = np.empty(len(cell_rank))
    for r, a in enumerate(addr):
        f = foamfile.read_field(os.path.join(case.dir, "processor%d" % r, time_name, case.alpha_name))
        out[a] = f.internal_array(len(a))
    return out


def calc_vof_advection_errors(case, exact_dir=None, exact_name=None, mesh=None, cell_volumes=None):
    """calcVofAdvectionErrors (updateErrors.H): per written time with an exact field,
    (time, E_v, alphaMin, 1-alphaMax, E_s).  The exact field is <time>/alpha.water.exact in the case, or
    <exact_dir>/<time>/alpha.water.exact (what Allrun.linkExactSolutions links)."""
    case = case if isinstance(case, FoamCase) else FoamCase(case)
    exact_name = exact_name or case.alpha_name + ".exact"
    rows = []
    for t, name in case.time_dirs():
        a_path = os.path.join(case.dir, name, case.alpha_name)
        e_path = os.path.join(exact_dir, name, exact_name) if exact_dir else os.path.join(case.dir, name, exact_name)
        if not (os.path.isfile(a_path) and os.path.isfile(e_path)):
            continue
        fa, fe = foamfile.read_field(a_path), foamfile.read_field(e_path)
        if not isinstance(fa.internal, np.ndarray):
            continue
        a = fa.internal
        e = fe.internal_array(a.shape[0])
        V = cell_volumes if cell_volumes is not None else np.full(a.shape[0], 1.0 / a.shape[0])
        rows.append((t, (np.sum(a * V) - EXACT_INITIAL_VOL) / EXACT_INITIAL_VOL, float(a.min()), float(1.0 - a.max()),
                     float(np.sum(np.abs(a - e) * V) / np.sum(e * V))))
    return rows


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m geometricvofext_b200.foamcase")
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("run", help="blockMesh [+ renumberMesh] + plicVofAdvectionFoam [+ calcVofAdvectionErrors]")
    r.add_argument("case")
    r.add_argument("--end-time", type=float, default=None)
    r.add_argument("--renumber", action="store_true", help="renumberMesh -overwrite (Cuthill-McKee), as the reference's Allrun does")
    r.add_argument("--exact", default=None, help="directory of <time>/alpha.water.exact fields; its time 0 field is the initial condition")
    r.add_argument("--write-mesh", action="store_true", help="also write constant/polyMesh")
    b = sub.add_parser("blockMesh", help="system/blockMeshDict -> constant/polyMesh")
    b.add_argument("case")
    b.add_argument("--renumber", action="store_true")
    d = sub.add_parser("decompose", help="processor*/constant/polyMesh/cellProcAddressing + processor*/0/alpha.water (weighted bisection)")
    d.add_argument("case")
    d.add_argument("n", type=int)
    pr = sub.add_parser("parallel", help="one rank of a decomposed case; launch with python -m torch.distributed.run --nproc-per-node N "
                                        "-m geometricvofext_b200.foamcase parallel <case> (nccl on GPUs, gloo with --cpu-engine <oracle .so>)")
    pr.add_argument("case")
    pr.add_argument("--end-time", type=float, default=None)
    pr.add_argument("--cpu-engine", default=None, help="path of a library exporting include/svof.h to run on the host (tests)")
    args = ap.parse_args(argv)
    case = FoamCase(args.case)
    if args.cmd == "decompose":
        if not case.has_poly_mesh():
            foamfile.write_polymesh(case.mesh(), case.dir)
        cr = decompose_case(case, args.n)
        print("cells per processor:", np.bincount(cr, minlength=args.n).tolist())
        return 0
    if args.cmd == "parallel":
        import torch.distributed as dist
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if args.cpu_engine:
            dist.init_process_group("gloo", rank=rank, world_size=world)
            out = run_plic_vof_advection_decomposed(case, rank, world, lib=capi.load(args.cpu_engine), end_time=args.end_time,
                                                    log=print)
        else:
            import torch
            dev = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(dev)
            dist.init_process_group("nccl", rank=rank, world_size=world)
            out = run_plic_vof_advection_decomposed(case, rank, world, device=dev, end_time=args.end_time, log=print)
        if rank == 0:
            print("End: %d steps" % out["steps"])
        dist.barrier()
        dist.destroy_process_group()
        return 0
    if args.cmd == "blockMesh":
        m = case.mesh(renumber=args.renumber)
        print(foamfile.write_polymesh(m, case.dir, fmt=str(case.control_dict.get("writeFormat", "binary"))))
        return 0
    if args.exact:     # Allrun:14: ln -rsf ../exactSolutions/0/alpha.water.exact 0/alpha.water
        os.makedirs(os.path.join(case.dir, "0"), exist_ok=True)
        shutil.copyfile(os.path.join(args.exact, "0", case.alpha_name + ".exact"), os.path.join(case.dir, "0", case.alpha_name))
    if args.write_mesh and not case.has_poly_mesh():
        foamfile.write_polymesh(case.mesh(renumber=args.renumber), case.dir)
    out = run_plic_vof_advection(case, end_time=args.end_time, renumber=args.renumber, log=print)
    print("End: %d steps, reconstruction %.3f s, advection %.3f s" % (out["steps"], out["reconstruction_time"], out["advection_time"]))
    if args.exact:
        print("%8s %14s %14s %14s %14s" % ("Time", "E_v", "alphaMin", "1-alphaMax", "E_s"))
        for row in calc_vof_advection_errors(case, exact_dir=args.exact):
            print("%8g %14.6e %14.6e %14.6e %14.6e" % row)
    return 0


if __name__ == "__main__":
    sys.exit(main())
