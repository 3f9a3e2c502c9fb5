"""Shared test helpers.  This is the only place (besides bench.py's CPU legs and
smoke()) that touches oracle/: it loads the oracle library as the CHECKER."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from geometricvofext_b200 import capi, fields, mesh as meshmod  # noqa: E402
from geometricvofext_b200.solver import SolveVofEqu  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build as oracle_build  # noqa: E402

_oracle = None
_ref = None

LEVEQUE_CONTROLS = {  # tutorials/test/plicVofAdvectionFoam/system/fvSolution:21-35
    "nAlphaBounds": 3, "snapTol": 0, "clip": False, "mixedCellTol": 1e-8, "orientationMethod": "LS",
    "splitWarpedFace": False, "writePlicFields": True, "nAlphaSubCycles": 1, "cAlpha": 1, "period": 6.0,
    "reverseTime": 0.0,
}


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = capi.load(oracle_build.build_oracle())
    return _oracle


def ref_overlap_lib():
    """The reference's own overlap.hpp compiled into oracle/_ref (None if unavailable)."""
    global _ref
    if _ref is None:
        p = oracle_build.build_ref()
        if p is None:
            return None
        _ref = C.CDLL(p)
        _ref.ref_sphere_hex_overlap.restype = C.c_int
        _ref.ref_sphere_hex_overlap.argtypes = [C.c_long, capi.c_double_p, capi.c_double_p, C.c_double, capi.c_double_p]
    return _ref


class RefCut:
    """The REFERENCE's own cutFace / cutCell classes (src/SimPLIC/cut, compiled unmodified into oracle/_ref/libref_cut.so
    against the OpenFOAM stand-in oracle/of_stub/), on a mesh.  Geometry and flatness come from `geom` (a SolveVofEqu of
    the same mesh): they are OpenFOAM / reconstruction.C quantities, outside those four files."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            p = oracle_build.build_ref_cut()
            if p is None:
                return None
            L = C.CDLL(p)
            dp, ip = capi.c_double_p, capi.c_int32_p
            L.ref_cut_create.restype = C.c_void_p
            L.ref_cut_create.argtypes = [C.POINTER(capi.SvofMesh), dp, dp, dp, dp, dp]
            L.ref_cut_destroy.argtypes = [C.c_void_p]
            L.ref_cut_faces.argtypes = [C.c_void_p, C.c_int32, C.c_int32, dp, dp, dp, ip, dp, dp]
            L.ref_cut_cells.argtypes = [C.c_void_p, C.c_int32, ip, dp, dp, ip, dp, dp, dp, dp]
            L.ref_find_signed_distance.argtypes = [C.c_void_p, C.c_int32, ip, dp, dp, C.c_int32, ip, dp, dp, dp]
            L.ref_face_fluxes.argtypes = [C.c_void_p, C.c_int32, ip, dp, dp, dp, C.c_double, dp, dp]
            L.ref_interface_points.argtypes = [C.c_void_p, C.c_int32, dp, C.c_double, C.c_int32, dp]
            cls._lib = L
        return cls._lib

    def __init__(self, m, geom, split=False):
        self.L = self.lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libref_cut.so unavailable")
        self.split = int(bool(split))
        self._cm, self._keep = m.to_c()
        Cf, Sf = geom.field(capi.F_CF), geom.field(capi.F_SF)
        self._g = [capi.f64(Cf), capi.f64(geom.field(capi.F_C)), capi.f64(geom.field(capi.F_V)),
                   capi.f64(np.sqrt(Sf[:, 0] * Sf[:, 0] + Sf[:, 1] * Sf[:, 1] + Sf[:, 2] * Sf[:, 2])),
                   capi.f64(geom.faceFlatness())]
        self._h = self.L.ref_cut_create(C.byref(self._cm), *[capi.dptr(a) for a in self._g])
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_cut_destroy(self._h)
            self._h = None

    def cutFaces(self, pts, normals, dists):
        pts = capi.f64(pts)
        n, nv = pts.shape[0], pts.shape[1]
        normals, dists = capi.f64(normals, (n, 3)), capi.f64(dists, (n,))
        st, ce, ar = np.empty(n, np.int32), np.empty((n, 3)), np.empty((n, 3))
        self.L.ref_cut_faces(self._h, n, nv, capi.dptr(pts), capi.dptr(normals), capi.dptr(dists), capi.iptr(st), capi.dptr(ce),
                             capi.dptr(ar))
        return st, ce, ar

    def cutCells(self, cells, normals, dists):
        cells = capi.i32(cells)
        n = cells.shape[0]
        normals, dists = capi.f64(normals, (n, 3)), capi.f64(dists, (n,))
        st, vof, sv, ic, ia = np.empty(n, np.int32), np.empty(n), np.empty(n), np.empty((n, 3)), np.empty((n, 3))
        self.L.ref_cut_cells(self._h, n, capi.iptr(cells), capi.dptr(normals), capi.dptr(dists), capi.iptr(st), capi.dptr(vof),
                             capi.dptr(sv), capi.dptr(ic), capi.dptr(ia))
        return st, vof, sv, ic, ia

    def findSignedDistance(self, cells, alphas, normals):
        cells = capi.i32(cells)
        n = cells.shape[0]
        alphas, normals = capi.f64(alphas, (n,)), capi.f64(normals, (n, 3))
        st, D, ic, ia = np.empty(n, np.int32), np.empty(n), np.empty((n, 3)), np.empty((n, 3))
        self.L.ref_find_signed_distance(self._h, n, capi.iptr(cells), capi.dptr(alphas), capi.dptr(normals), self.split,
                                        capi.iptr(st), capi.dptr(D), capi.dptr(ic), capi.dptr(ia))
        return st, D, ic, ia

    def faceFluxes(self, faces, normals, dists, Un0, dt, phi):
        faces = capi.i32(faces)
        n = faces.shape[0]
        normals, dists, Un0, phi = capi.f64(normals, (n, 3)), capi.f64(dists, (n,)), capi.f64(Un0, (n,)), capi.f64(phi, (n,))
        out = np.empty(n)
        self.L.ref_face_fluxes(self._h, n, capi.iptr(faces), capi.dptr(normals), capi.dptr(dists), capi.dptr(Un0), float(dt),
                               capi.dptr(phi), capi.dptr(out))
        return out

    def interfacePoints(self, cell, normal, dist, cap=64):
        out = np.empty((cap, 3))
        n = self.L.ref_interface_points(self._h, int(cell), capi.dptr(capi.f64(normal, (3,))), float(dist), cap, capi.dptr(out))
        return out[:n].copy()


class RefAdvect:
    """The REFERENCE's own advection class (src/SimPLIC/advection/advection.{H,C}, advectionTemplates.C, + cutFace.C,
    compiled unmodified into oracle/_ref/libref_advect.so against the OpenFOAM stand-in oracle/of_stub_adv/), on a mesh.
    One call = one advection::advect(Sp, Su) on the state handed in; the reconstruction it reads is an input."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            p = oracle_build.build_ref_advect()
            if p is None:
                return None
            L = C.CDLL(p)
            dp, ip = capi.c_double_p, capi.c_int32_p
            L.ref_advect_create.restype = C.c_void_p
            L.ref_advect_create.argtypes = [C.POINTER(capi.SvofMesh), C.POINTER(capi.SvofParams)]
            L.ref_advect_destroy.argtypes = [C.c_void_p]
            L.ref_advect_step.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int32, ip, ip, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp]
            L.ref_advect_log.restype = C.c_char_p
            L.ref_advect_log.argtypes = [C.c_void_p]
            L.ref_advect_error.restype = C.c_char_p
            L.ref_advect_error.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, m, params):
        self.L = self.lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libref_advect.so unavailable")
        self.m = m
        self._cm, self._keep = m.to_c()
        self._h = self.L.ref_advect_create(C.byref(self._cm), C.byref(params))
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_advect_destroy(self._h)
            self._h = None

    def step(self, alpha, phi, U, Ub, mixed, status, iN, iD, iC, dt, Sp=None, Su=None):
        """-> alpha_new [nC], alphaPhi [nF], alpha boundary values [nBF], the Info lines of the step"""
        m = self.m
        nC, nF, nBF = m.n_cells, m.n_faces, m.n_faces - m.n_internal_faces
        alpha, phi = capi.f64(alpha, (nC,)), capi.f64(phi, (nF,))
        U, Ub = capi.f64(U, (nC, 3)), capi.f64(Ub, (nBF, 3))
        mixed, status = capi.i32(mixed), capi.i32(status)
        iN, iD, iC = capi.f64(iN, (nC, 3)), capi.f64(iD, (nC,)), capi.f64(iC, (nC, 3))
        sp = capi.f64(Sp, (nC,)) if Sp is not None else None
        su = capi.f64(Su, (nC,)) if Su is not None else None
        a, ap, ab = np.empty(nC), np.empty(nF), np.empty(nBF)
        rc = self.L.ref_advect_step(self._h, capi.dptr(alpha), capi.dptr(phi), capi.dptr(U), capi.dptr(Ub), len(mixed),
                                    capi.iptr(mixed), capi.iptr(status), capi.dptr(iN), capi.dptr(iD), capi.dptr(iC), float(dt),
                                    capi.dptr(sp) if sp is not None else None, capi.dptr(su) if su is not None else None,
                                    capi.dptr(a), capi.dptr(ap), capi.dptr(ab))
        if rc != 0:
            raise RuntimeError("reference advect failed: %s" % self.L.ref_advect_error(self._h).decode())
        return a, ap, ab, self.L.ref_advect_log(self._h).decode()


from refsolver import RefSolver  # noqa: E402,F401  (oracle/refsolver.py: ctypes view of oracle/_ref/libref_solver.so)


def exact_sphere_alpha(m, centre=(0.35, 0.35, 0.35), radius=0.15):
    """Exact sphere/hex volume fractions on a hex_block mesh via the reference's
    overlap library (calcExactVofFieldForSphericalShapeInHexMesh/functions.H:1-27)."""
    lib = ref_overlap_lib()
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_overlap.so unavailable")
    Cc = meshmod.cell_centres_hex(m)
    N = np.array(m.meta["N"])
    h = np.array(m.meta["length"]) / N
    d = np.linalg.norm(Cc - np.array(centre), axis=1)
    hd = 0.5 * np.linalg.norm(h)
    alpha = np.zeros(m.n_cells)
    alpha[d + hd <= radius] = 1.0
    cut = np.nonzero(np.abs(d - radius) < hd * (1 + 1e-9))[0]
    off = 0.5 * h * np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                              [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
    hexes = np.ascontiguousarray(Cc[cut][:, None, :] + off[None, :, :])
    vol = np.empty(cut.size)
    c = np.array(centre, dtype=np.float64)
    lib.ref_sphere_hex_overlap(cut.size, capi.dptr(hexes), capi.dptr(c), float(radius), capi.dptr(vol))
    alpha[cut] = vol / np.prod(h)
    return alpha


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
