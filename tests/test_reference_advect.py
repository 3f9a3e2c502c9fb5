"""Pin the oracle (and, on a GPU, the CUDA product) to the REFERENCE's own advection class.

oracle/_ref/libref_advect.so is src/SimPLIC/advection/advection.{H,C} + advectionTemplates.C + cut/cutFace/cutFace.{H,C},
compiled unmodified from the reference tree against a stand-in for the OpenFOAM types they use (oracle/of_stub_adv/,
recipe oracle/build.py:build_ref_advect).  Each step the implementation under test reconstructs, its reconstruction
(interface-cell list, cut status, interfaceN/D/C) and its fields are handed to the reference's advection::advect(Sp, Su)
(advectionTemplates.C:352-418), and the bar is BITWISE equality of the new alpha, alphaPhi and the alpha patch values --
i.e. of everything advection.C / advectionTemplates.C do: downwind-face selection and time-integrated face fluxes
(advection.C:85-219), the alpha update (:399-404), limitFlux / boundFlux (:113-349, with the "Before / After
conservative bounding" numbers of its Info lines), snap / clip (advection.C:291-308), alphaPhi (:417).
What stays outside (OpenFOAM services, written from their published definitions in the stand-in): field algebra,
upwind::flux, fvc::surfaceIntegrate, patch-field evaluation, interpolationCellPoint, mesh geometry.
"""
import re

import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, RefAdvect, SolveVofEqu, capi, fields, meshmod, oracle_lib
from test_edge_cases import _case as edge_case

pytestmark = pytest.mark.skipif(RefAdvect.lib() is None, reason="oracle/_ref/libref_advect.so not built (no reference tree)")


def _smeared_sphere(C_, V, centre=(0.5, 0.62, 0.5), radius=0.15):
    h = np.cbrt(V)
    return np.clip(0.5 - (np.linalg.norm(C_ - np.array(centre), axis=1) - radius) / h, 0.0, 1.0)


def _poly(make, extra=None, cfl=0.25, steps=5):
    return lambda: (make(), dict(extra or {}), _smeared_sphere, fields.rotation_velocity, steps, cfl)


CASES = {
    # the reference's own test case, scaled down: LeVeque deformation of a sphere, hexahedra
    "LeVeque hexes 16^3": lambda: (meshmod.hex_block(16), {}, None, fields.leveque_velocity, 8, 0.5),
    "LeVeque hexes, Courant 1.2, nAlphaBounds 10": lambda: (meshmod.hex_block(12), {"nAlphaBounds": 10}, None, fields.leveque_velocity, 5, 1.2),
    "damBreak controls (clip, snapTol, mixedCellTol)": lambda: (meshmod.hex_block(14), {"clip": True, "snapTol": 1e-12, "mixedCellTol": 1e-10,
                                                                                   "nAlphaBounds": 5}, None, fields.leveque_velocity, 6, 0.6),
    "prisms": _poly(lambda: meshmod.prism_mesh(8)),
    "refinement-interface polyhedra": _poly(lambda: meshmod.refined_interface_mesh(8)),
    "warped hexes": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3)),
    "warped hexes, Courant 0.9": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.2, 5), cfl=0.9, steps=4),
    "Kelvin cells": _poly(lambda: meshmod.kelvin_mesh(8)),
}
for _n in ("every cell cut", "no bounding sweeps", "Courant number 1.5", "2-D: empty front and back", "inflow and outflow patches",
           "no interface: full", "single cell"):
    CASES["edge: " + _n] = (lambda n=_n: edge_case(n))


def _check_against_reference(case, lib, what, sources=False):
    from common import exact_sphere_alpha
    m, extra, alpha0, vel, steps, cfl = CASES[case]()
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, **extra), lib=lib)
    ref = RefAdvect(m, s._params)
    C_, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    a0 = exact_sphere_alpha(m) if alpha0 is None else alpha0(C_, V)
    U0, phi0 = vel(C_), fields.face_flux(Cf, Sf, vel)
    Ub = vel(Cf[m.n_internal_faces:])
    umax = max(np.abs(U0).max(), 1e-30)
    dt = cfl * np.cbrt(V.min()) / umax
    rng = np.random.default_rng(5)
    Sp = -0.3 * rng.random(m.n_cells) if sources else None
    Su = 0.05 * rng.random(m.n_cells) * (a0 > 0) if sources else None
    s.setAlpha(a0)
    s.setPhi(phi0)
    s.setU(U0, Ub)
    n_mixed = n_swept = 0
    for k in range(steps):
        s.reconstruct()
        a_old = s.alpha()
        mixed, status = s.mixedCells(), s.cellStatus()
        iN, iD, iC = s.interfaceN(), s.interfaceD(), s.field(capi.F_INTERFACE_C)
        r_alpha, r_alphaPhi, r_alphaB, log = ref.step(a_old, phi0, U0, Ub, mixed, status, iN, iD, iC, dt, Sp, Su)
        s.advect(dt, Sp=Sp, Su=Su)
        a, ap, ab = s.alpha(), s.alphaPhi(), s.field(capi.F_ALPHA_BOUNDARY)
        assert np.array_equal(a, r_alpha), "%s, step %d: alpha differs from the reference's advect() by %g" % (what, k, np.abs(a - r_alpha).max())
        assert np.array_equal(ap, r_alphaPhi), "%s, step %d: alphaPhi differs by %g" % (what, k, np.abs(ap - r_alphaPhi).max())
        assert np.array_equal(ab, r_alphaB), "%s, step %d: alpha patch values differ" % (what, k)
        # the two Info lines of limitFlux (advectionTemplates.C:133-134, 211-212), printed with 17 digits by the stand-in
        nums = re.findall(r"min\(alpha\) = (\S+), max\(alpha\) = 1 \+ (\S+)", log)
        assert len(nums) == 2, log
        (mn_b, mx_b), (mn_a, mx_a) = [(float(x), float(y)) for x, y in nums]
        assert mn_b == s.info(capi.I_MIN_ALPHA_BEFORE) and mx_b == s.info(capi.I_MAX_ALPHA_M1_BEFORE), "%s, step %d: 'Before' line" % (what, k)
        assert mn_a == s.info(capi.I_MIN_ALPHA_AFTER) and mx_a == s.info(capi.I_MAX_ALPHA_M1_AFTER), "%s, step %d: 'After' line" % (what, k)
        n_mixed = max(n_mixed, len(mixed))
        n_swept += int(s.info(capi.I_N_BOUND_SWEEPS))
    flags = s.info(capi.I_ERROR_FLAGS)
    s.close()
    assert flags == 0
    return n_mixed, n_swept


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_advect_matches_reference_class(case):
    n_mixed, n_swept = _check_against_reference(case, oracle_lib(), "oracle, " + case)
    if not case.startswith("edge: no interface") and case != "edge: single cell":
        assert n_mixed > 20
    if "Courant" in case:
        assert n_swept > 0          # boundFlux really ran


@pytest.mark.parametrize("case", ["damBreak controls (clip, snapTol, mixedCellTol)", "warped hexes", "edge: inflow and outflow patches"])
def test_oracle_advect_with_sources_matches_reference_class(case):
    """Sp / Su as fields (interPlicPhaseChangeFoam/alphaSuSp.H:14-15): the volScalarField::Internal instantiation."""
    _check_against_reference(case, oracle_lib(), "oracle with Sp/Su, " + case, sources=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_advect_matches_reference_class(case, product):
    """The CUDA library against the reference's advection class directly (no oracle in between)."""
    _check_against_reference(case, product, "CUDA, " + case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["damBreak controls (clip, snapTol, mixedCellTol)", "warped hexes"])
def test_gpu_advect_with_sources_matches_reference_class(case, product):
    _check_against_reference(case, product, "CUDA with Sp/Su, " + case, sources=True)
