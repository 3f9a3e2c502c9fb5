"""Host-side mirror of the reference's facade class

    geometricVofExt::SimPLIC::solveVofEqu   (src/SimPLIC/solveVofEqu/solveVofEqu.H:117-211)

with the same member names (reconstruct, advect, getRhoPhi, alphaPhi,
reconstructionTime, advectionTime, alphaMappingTime, ...) and the same control
dictionary (`fvSolution` solvers."alpha.*", solveVofEqu.C:68), on top of the C
ABI in include/svof.h.  All numerics happen behind that ABI in the CUDA library;
this file only marshals numpy arrays.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import SvofError


def make_params(lib, controls=None):
    """fvSolution dictionary (python dict of key -> value) -> svof_params."""
    p = capi.SvofParams()
    rc = lib.svof_params_default(C.byref(p))
    if rc:
        raise SvofError(rc, "svof_params_default")
    for k, v in (controls or {}).items():
        if isinstance(v, bool):
            v = "true" if v else "false"
        rc = lib.svof_params_set(C.byref(p), str(k).encode(), str(v).encode())
        if rc == capi.ERR_BAD_CONFIG:
            # same text as reconstruction.C:612-624
            raise SvofError(rc, "Orientation vector calculation method '%s' is not valid. Valid methods are "
                                "(alphaGrad isoAlphaGrad isoRDF)" % v)
        if rc:
            raise SvofError(rc, "unknown or malformed control '%s %s'" % (k, v))
    return p


class SolveVofEqu:
    """solveVofEqu(alpha1, phi, U) on a PolyMesh.

    lib      -- a library loaded with capi.load(); default: the CUDA product
    controls -- the fvSolution solvers."alpha.*" dictionary
    comm     -- (rank, world_size, device[, address of the NCCL unique id]) for decomposed runs (multigpu.py)
    """

    typeName = "solveVofEqu"

    def __init__(self, mesh, controls=None, lib=None, comm=None):
        self.lib = lib if lib is not None else capi.load_product()
        self.mesh = mesh
        self.controls = dict(controls or {})
        self._params = make_params(self.lib, self.controls)
        cm, keep = mesh.to_c()
        c = tuple(comm) if comm is not None else (0, 1, -1, None)
        if len(c) == 3:
            c = c + (None,)
        cc = capi.SvofComm(int(c[0]), int(c[1]), int(c[2]), 0, c[3])   # c[3]: address of the 128-byte NCCL id (world_size > 1)
        h = C.c_void_p()
        rc = self.lib.svof_create(C.byref(cm), C.byref(self._params), C.byref(cc), C.byref(h))
        del keep
        if rc:
            raise SvofError(rc, self.lib.svof_last_error(None).decode())
        self._h = h
        self.nC, self.nF, self.nIF = mesh.n_cells, mesh.n_faces, mesh.n_internal_faces
        self.nBF = self.nF - self.nIF

    # -- life cycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.svof_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise SvofError(rc, self.lib.svof_last_error(self._h).decode())
        return rc

    # -- fields in ------------------------------------------------------------
    def setAlpha(self, alpha):
        a = capi.f64(alpha, (self.nC,))
        self._chk(self.lib.svof_set_alpha(self._h, capi.dptr(a)))

    def setPhi(self, phi):
        a = capi.f64(phi, (self.nF,))
        self._chk(self.lib.svof_set_phi(self._h, capi.dptr(a)))

    def setU(self, U, Ub=None):
        u = capi.f64(U, (self.nC, 3))
        ub = capi.f64(Ub, (self.nBF, 3)) if Ub is not None else np.zeros((self.nBF, 3))
        self._chk(self.lib.svof_set_U(self._h, capi.dptr(u), capi.dptr(ub)))

    # -- the reference's member functions ----------------------------------------
    def reconstruct(self):
        """solveVofEqu::reconstruct() (solveVofEqu.C:96-99)."""
        self._chk(self.lib.svof_reconstruct(self._h))

    def advect(self, dt, Sp=None, Su=None):
        """solveVofEqu::advect(Sp, Su) (solveVofEquTemplates.C:35-43); None == zeroField."""
        sp = capi.f64(Sp, (self.nC,)) if Sp is not None else None
        su = capi.f64(Su, (self.nC,)) if Su is not None else None
        self._chk(self.lib.svof_advect(self._h, float(dt), capi.dptr(sp), capi.dptr(su)))

    def step(self, dt):
        """reconstruct() + advect(dt) on the device-resident fields, one CUDA-graph launch in the steady state."""
        self._chk(self.lib.svof_step_device(self._h, float(dt)))

    def step_host(self, dt, phi, U, Ub=None, alpha_out=None, alpha_phi_out=None):
        """set phi/U from host buffers, reconstruct, advect, read alpha back."""
        ub = Ub if Ub is not None else np.zeros((self.nBF, 3))
        self._chk(self.lib.svof_step_host(self._h, float(dt), capi.dptr(phi), capi.dptr(U), capi.dptr(ub),
                                          capi.dptr(alpha_out), capi.dptr(alpha_phi_out)))

    def mapAlphaField(self, lower=0.01, upper=0.99):
        """solveVofEqu::mapAlphaField() (reconstruction.C:725-784): after a refinement step, alpha in the cells whose
        value lies within the refinement levels is re-derived from the mapped PLIC plane.  The caller decides when
        (mesh.changing() && mapAlphaField) and maps interfaceN/D first (setInterface)."""
        self._chk(self.lib.svof_map_alpha_field(self._h, float(lower), float(upper)))

    def setCellTypes(self, cell_types):
        """dynamicOversetFvMesh: the overset stencil's cell types (0 = CALCULATED); interface cells are listed only among the
        CALCULATED cells (reconstruction.C:649-662).  None switches the filter off.  phi must be the masked flux."""
        if cell_types is None:
            self._chk(self.lib.svof_set_cell_types(self._h, None))
            return
        t = capi.i32(cell_types)
        assert t.shape == (self.nC,)
        self._chk(self.lib.svof_set_cell_types(self._h, capi.iptr(t)))

    def setInterface(self, N, D):
        n, d = capi.f64(N, (self.nC, 3)), capi.f64(D, (self.nC,))
        self._chk(self.lib.svof_set_interface(self._h, capi.dptr(n), capi.dptr(d)))

    def updatePoints(self, points):
        """mesh.moving(): new point positions, same topology"""
        p = capi.f64(points, (self.mesh.n_points, 3))
        self._chk(self.lib.svof_update_points(self._h, capi.dptr(p), None, None, None, None))
        self.mesh.points = p

    def updateMesh(self, mesh):
        """mesh.topoChanging(): rebuild the handle's mesh tables for `mesh`; fields must be set again"""
        cm, keep = mesh.to_c()
        self._chk(self.lib.svof_update_mesh(self._h, C.byref(cm)))
        del keep
        self.mesh = mesh
        self.nC, self.nF, self.nIF = mesh.n_cells, mesh.n_faces, mesh.n_internal_faces
        self.nBF = self.nF - self.nIF

    def alpha(self):
        return self.field(capi.F_ALPHA)

    def alphaPhi(self):
        return self.field(capi.F_ALPHA_PHI)

    def dict(self):
        return self.controls

    def getRhoPhi(self, rho1, rho2):
        """advection::getRhoPhi(rho1, rho2) for dimensionedScalar densities
        (advection.H:329-343): (rho1 - rho2)*alphaPhi + rho2*phi."""
        raise NotImplementedError("getRhoPhi needs phi on the host; use rhoPhi(phi, rho1, rho2)")

    def rhoPhi(self, phi, rho1, rho2):
        return (rho1 - rho2) * self.alphaPhi() + rho2 * np.asarray(phi)

    def reconstructionTime(self):
        return self.info(capi.I_RECONSTRUCTION_TIME)

    def advectionTime(self):
        return self.info(capi.I_ADVECTION_TIME)

    def alphaMappingTime(self):
        return self.info(capi.I_ALPHA_MAPPING_TIME)

    # -- reconstruction:: accessors (reconstruction.H) ------------------------------
    def mixedCells(self):
        return self.field(capi.F_MIXED_CELLS)

    def cellStatus(self):
        return self.field(capi.F_CELL_STATUS)

    def interfaceN(self):
        return self.field(capi.F_INTERFACE_N)

    def interfaceD(self):
        return self.field(capi.F_INTERFACE_D)

    def interfaceC(self):
        return self.field(capi.F_INTERFACE_C)

    def interfaceS(self):
        return self.field(capi.F_INTERFACE_S)

    def faceFlatness(self):
        return self.field(capi.F_FACE_FLATNESS)

    def interface(self):
        """reconstruction::interface() (reconstruction.C:787-835): the PLIC polygons of the last reconstruct() as
        (points [nP,3], face_offsets [nFaces+1], meshCells [nFaces]) -- what the plicSurface sampler reads."""
        nP, nFc = C.c_int64(), C.c_int64()
        self._chk(self.lib.svof_plic_surface(self._h, 0, 0, None, None, None, C.byref(nP), C.byref(nFc)))
        pts = np.empty((nP.value, 3))
        off, cells = np.zeros(nFc.value + 1, np.int32), np.empty(nFc.value, np.int32)
        self._chk(self.lib.svof_plic_surface(self._h, nP.value, nFc.value, capi.dptr(pts), capi.iptr(off), capi.iptr(cells),
                                             C.byref(nP), C.byref(nFc)))
        return pts, off, cells

    def subCellFaces(self):
        """reconstruction::subCellFaces() (reconstruction.C:838-891): faces of the submerged sub-cells of the cut cells of the
        last reconstruct() as (points [nP,3], face_offsets [nF+1], face_points, face_cell [nF])."""
        nP, nF, nFP = C.c_int64(), C.c_int64(), C.c_int64()
        self._chk(self.lib.svof_subcell_faces(self._h, 0, 0, 0, None, None, None, None, C.byref(nP), C.byref(nF), C.byref(nFP)))
        pts = np.empty((nP.value, 3))
        off, fp, fc = np.zeros(nF.value + 1, np.int32), np.empty(nFP.value, np.int32), np.empty(nF.value, np.int32)
        self._chk(self.lib.svof_subcell_faces(self._h, nP.value, nF.value, nFP.value, capi.dptr(pts), capi.iptr(off), capi.iptr(fp),
                                              capi.iptr(fc), C.byref(nP), C.byref(nF), C.byref(nFP)))
        return pts, off, fp, fc

    # -- generic access ---------------------------------------------------------------
    _SHAPES = {
        capi.F_ALPHA: ("nC", 1, np.float64), capi.F_ALPHA_PHI: ("nF", 1, np.float64),
        capi.F_DVF: ("nF", 1, np.float64), capi.F_INTERFACE_N: ("nC", 3, np.float64),
        capi.F_INTERFACE_D: ("nC", 1, np.float64), capi.F_INTERFACE_C: ("nC", 3, np.float64),
        capi.F_INTERFACE_S: ("nC", 3, np.float64), capi.F_MIXED_CELLS: ("nM", 1, np.int32),
        capi.F_CELL_STATUS: ("nM", 1, np.int32), capi.F_FACE_FLATNESS: ("nF", 1, np.float64),
        capi.F_CF: ("nF", 3, np.float64), capi.F_SF: ("nF", 3, np.float64), capi.F_C: ("nC", 3, np.float64),
        capi.F_V: ("nC", 1, np.float64), capi.F_ALPHA_BOUNDARY: ("nBF", 1, np.float64),
        capi.F_UN0: ("nM", 1, np.float64),
    }

    def field(self, which, out=None):
        dim, ncomp, dt = self._SHAPES[which]
        n = int(self.info(capi.I_N_MIXED)) if dim == "nM" else getattr(self, dim)
        if out is None:
            out = np.empty((n, ncomp) if ncomp > 1 else (n,), dtype=dt)
        got = self.lib.svof_get_field(self._h, which, out.ctypes.data_as(C.c_void_p), out.size)
        self._chk(got)
        return out

    def info(self, which):
        v = C.c_double()
        self._chk(self.lib.svof_get_info(self._h, which, C.byref(v)))
        return v.value

    def volume(self):
        """gSum(alpha*V) as printed by the reference driver (plicVof.H:44-46)."""
        return self.info(capi.I_VOLUME)

    def setOption(self, name, value):
        self._chk(self.lib.svof_set_option(self._h, name.encode(), int(value)))

    def synchronize(self):
        self._chk(self.lib.svof_synchronize(self._h))

    def last_step_ms(self):
        r, a = C.c_double(), C.c_double()
        self._chk(self.lib.svof_last_step_ms(self._h, C.byref(r), C.byref(a)))
        return r.value, a.value

    # -- geometry primitives ----------------------------------------------------------------
    def cutFaces(self, pts, normals, dists):
        pts = capi.f64(pts)
        n_polys, n_verts = pts.shape[0], pts.shape[1]
        normals, dists = capi.f64(normals, (n_polys, 3)), capi.f64(dists, (n_polys,))
        st = np.empty(n_polys, np.int32)
        ce, ar = np.empty((n_polys, 3)), np.empty((n_polys, 3))
        self._chk(self.lib.svof_cut_faces(self._h, n_polys, n_verts, capi.dptr(pts), capi.dptr(normals),
                                          capi.dptr(dists), capi.iptr(st), capi.dptr(ce), capi.dptr(ar)))
        return st, ce, ar

    def cutCells(self, cells, normals, dists):
        cells = capi.i32(cells)
        n = cells.shape[0]
        normals, dists = capi.f64(normals, (n, 3)), capi.f64(dists, (n,))
        st, vof, sv = np.empty(n, np.int32), np.empty(n), np.empty(n)
        ic, ia = np.empty((n, 3)), np.empty((n, 3))
        self._chk(self.lib.svof_cut_cells(self._h, n, capi.iptr(cells), capi.dptr(normals), capi.dptr(dists),
                                          capi.iptr(st), capi.dptr(vof), capi.dptr(sv), capi.dptr(ic), capi.dptr(ia)))
        return st, vof, sv, ic, ia

    def findSignedDistance(self, cells, alphas, normals):
        cells = capi.i32(cells)
        n = cells.shape[0]
        alphas, normals = capi.f64(alphas, (n,)), capi.f64(normals, (n, 3))
        st, D = np.empty(n, np.int32), np.empty(n)
        ic, ia = np.empty((n, 3)), np.empty((n, 3))
        self._chk(self.lib.svof_find_signed_distance(self._h, n, capi.iptr(cells), capi.dptr(alphas),
                                                     capi.dptr(normals), capi.iptr(st), capi.dptr(D), capi.dptr(ic),
                                                     capi.dptr(ia)))
        return st, D, ic, ia

    def faceFluxes(self, faces, normals, dists, Un0, dt, phi):
        faces = capi.i32(faces)
        n = faces.shape[0]
        normals, dists, Un0, phi = capi.f64(normals, (n, 3)), capi.f64(dists, (n,)), capi.f64(Un0, (n,)), capi.f64(phi, (n,))
        out = np.empty(n)
        self._chk(self.lib.svof_face_fluxes(self._h, n, capi.iptr(faces), capi.dptr(normals), capi.dptr(dists),
                                            capi.dptr(Un0), float(dt), capi.dptr(phi), capi.dptr(out)))
        return out
