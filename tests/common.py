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
