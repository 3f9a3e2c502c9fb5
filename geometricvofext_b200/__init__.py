"""B200-native SimPLIC volume-fraction transport step (geometricVofExt's
solveVofEqu::reconstruct + advect) behind the C ABI of include/svof.h."""
from . import capi, fields, mesh  # noqa: F401
from .solver import SolveVofEqu, SvofError, make_params  # noqa: F401

__all__ = ["SolveVofEqu", "SvofError", "make_params", "capi", "fields", "mesh"]
