"""ctypes view of oracle/_ref/libref_solver.so -- TEST INFRASTRUCTURE ONLY (used by tests/ and by bench.py's CPU legs).

The library is the REFERENCE's own solveVofEqu class compiled unmodified against an OpenFOAM stand-in
(oracle/ref_solver.cpp, oracle/of_stub_rec/, recipe oracle/build.py:build_ref_solver)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(HERE))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import build as oracle_build  # noqa: E402
from geometricvofext_b200 import capi  # noqa: E402  (struct layouts of include/svof.h and array helpers only)


class RefSolver:
    """The REFERENCE's own solveVofEqu class -- solveVofEqu.C, reconstruction.C, advection.C + advectionTemplates.C,
    cutFace.C, cutCell.C compiled unmodified into oracle/_ref/libref_solver.so against the OpenFOAM stand-in
    oracle/of_stub_rec/ -- on a mesh: reconstruct(), advect(Sp, Su), mapAlphaField(), interface(), subCellFaces()."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            p = oracle_build.build_ref_solver()
            if p is None:
                return None
            L = C.CDLL(p)
            dp, ip, vp = capi.c_double_p, capi.c_int32_p, C.c_void_p
            i32p = C.POINTER(C.c_int32)
            L.ref_solver_create.restype = vp
            L.ref_solver_create.argtypes = [C.POINTER(capi.SvofMesh), C.POINTER(capi.SvofParams), C.c_int32]
            L.ref_solver_destroy.argtypes = [vp]
            L.ref_solver_face_flatness.argtypes = [vp, dp]
            L.ref_solver_set_state.argtypes = [vp, dp, dp, dp, dp]
            L.ref_solver_set_cell_types.argtypes = [vp, ip]
            L.ref_solver_reconstruct.argtypes = [vp, i32p]
            L.ref_solver_get_recon.argtypes = [vp, ip, ip, dp, dp, dp, dp]
            L.ref_solver_advect.argtypes = [vp, C.c_double, dp, dp]
            L.ref_solver_get_fields.argtypes = [vp, dp, dp, dp]
            L.ref_solver_map_alpha.argtypes = [vp, C.c_double, C.c_double]
            L.ref_solver_surface.argtypes = [vp, C.c_int32, i32p, i32p, i32p, i32p]
            L.ref_solver_surface_copy.argtypes = [vp, dp, ip, ip, ip]
            for f in ("ref_solver_log", "ref_solver_ctor_log", "ref_solver_error"):
                getattr(L, f).restype = C.c_char_p
                getattr(L, f).argtypes = [vp]
            cls._lib = L
        return cls._lib

    PLAIN, REFINE, OVERSET = 0, 1, 2

    def __init__(self, m, params, mesh_kind=0):
        self.L = self.lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libref_solver.so unavailable")
        self.m = m
        self.nC, self.nF, self.nBF = m.n_cells, m.n_faces, m.n_faces - m.n_internal_faces
        self._cm, self._keep = m.to_c()
        self._h = self.L.ref_solver_create(C.byref(self._cm), C.byref(params), mesh_kind)
        assert self._h
        self.nM = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_solver_destroy(self._h)
            self._h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("reference solveVofEqu failed: %s" % self.L.ref_solver_error(self._h).decode())

    def faceFlatness(self):
        z = np.empty(self.nF)
        self._chk(self.L.ref_solver_face_flatness(self._h, capi.dptr(z)))
        return z

    def ctorLog(self):
        return self.L.ref_solver_ctor_log(self._h).decode()

    def log(self):
        return self.L.ref_solver_log(self._h).decode()

    def setState(self, alpha=None, phi=None, U=None, Ub=None):
        """phi is set before alpha so that inletOutlet patch values see it (as correctBoundaryConditions would)."""
        a = capi.f64(alpha, (self.nC,)) if alpha is not None else None
        p = capi.f64(phi, (self.nF,)) if phi is not None else None
        u = capi.f64(U, (self.nC, 3)) if U is not None else None
        ub = capi.f64(Ub, (self.nBF, 3)) if Ub is not None else None
        opt = lambda x: capi.dptr(x) if x is not None else None
        self._chk(self.L.ref_solver_set_state(self._h, opt(a), opt(p), opt(u), opt(ub)))

    def setCellTypes(self, types):
        t = capi.i32(types)
        self._chk(self.L.ref_solver_set_cell_types(self._h, capi.iptr(t)))

    def reconstruct(self):
        n = C.c_int32()
        self._chk(self.L.ref_solver_reconstruct(self._h, C.byref(n)))
        self.nM = n.value

    def recon(self):
        """-> mixedCells, cellStatus, interfaceN, interfaceD, interfaceC, interfaceS of the last reconstruct()"""
        mc, st = np.empty(self.nM, np.int32), np.empty(self.nM, np.int32)
        iN, iD, iC, iS = np.empty((self.nC, 3)), np.empty(self.nC), np.empty((self.nC, 3)), np.empty((self.nC, 3))
        self._chk(self.L.ref_solver_get_recon(self._h, capi.iptr(mc), capi.iptr(st), capi.dptr(iN), capi.dptr(iD), capi.dptr(iC),
                                              capi.dptr(iS)))
        return mc, st, iN, iD, iC, iS

    def advect(self, dt, Sp=None, Su=None):
        sp = capi.f64(Sp, (self.nC,)) if Sp is not None else None
        su = capi.f64(Su, (self.nC,)) if Su is not None else None
        self._chk(self.L.ref_solver_advect(self._h, float(dt), capi.dptr(sp) if sp is not None else None,
                                           capi.dptr(su) if su is not None else None))

    def fields(self):
        """-> alpha [nC], alphaPhi [nF], alpha patch values [nBF]"""
        a, ap, ab = np.empty(self.nC), np.empty(self.nF), np.empty(self.nBF)
        self._chk(self.L.ref_solver_get_fields(self._h, capi.dptr(a), capi.dptr(ap), capi.dptr(ab)))
        return a, ap, ab

    def mapAlphaField(self, lower, upper):
        self._chk(self.L.ref_solver_map_alpha(self._h, float(lower), float(upper)))

    def surface(self, which):
        """which = 0: interface(), 1: subCellFaces() -> points [nP,3], face offsets, face points, meshCells"""
        nP, nFc, nFP, nCl = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._chk(self.L.ref_solver_surface(self._h, which, C.byref(nP), C.byref(nFc), C.byref(nFP), C.byref(nCl)))
        pts = np.empty((nP.value, 3))
        off, fp, cells = np.zeros(nFc.value + 1, np.int32), np.empty(nFP.value, np.int32), np.empty(nCl.value, np.int32)
        self._chk(self.L.ref_solver_surface_copy(self._h, capi.dptr(pts), capi.iptr(off), capi.iptr(fp), capi.iptr(cells)))
        return pts, off, fp, cells
