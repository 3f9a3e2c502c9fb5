"""Pin the oracle (and, on a GPU, the CUDA product) to the REFERENCE's own solveVofEqu class, end to end.

oracle/_ref/libref_solver.so is every file of the reference's src/SimPLIC outside sampling/ --
solveVofEqu/solveVofEqu.{H,C} + solveVofEquTemplates.C, reconstruction/reconstruction.{H,C}, advection/advection.{H,C} +
advectionTemplates.C, cut/cutFace/cutFace.{H,C}, cut/cutCell/cutCell.{H,C} -- compiled unmodified from the reference tree
against a stand-in for the OpenFOAM types they use (oracle/of_stub_rec/, recipe oracle/build.py:build_ref_solver).
The reference object and the implementation under test get the same mesh, controls and initial fields and then run
reconstruct() / advect(Sp, Su) step after step, each on its OWN state (nothing is handed across between steps);
the bar is BITWISE equality, every step, of
  * the interface-cell list and its order, the cut status (reconstruction::initialize, reconstruction.C:634-677),
  * interfaceN at the interface cells (calcInterfaceNFromRegAlphaGrad / IsoAlphaGrad / IsoRDF, :74-405),
    interfaceD / interfaceC / interfaceS everywhere (reconstruct, :680-722 -> cutCell::findSignedDistance),
  * alpha, alphaPhi, the alpha patch values after advect (advectionTemplates.C:352-418),
  * the face flatness of the constructor (updateFaceFlatness, :408-473) and the numbers of its Info line,
  * mapAlphaField (:725-784), and the overset filter of initialize (:649-662) where the product has it.
interface() / subCellFaces() (:787-891) are compared as geometry (point order within a polygon goes through atan2).
What stays outside (OpenFOAM services written from their published definitions in the stand-in, or answered by the
oracle's restatement): field algebra, Gauss-linear fvc::grad, leastSquareGrad's fill order and LU solve, the
cell-point-cell stencil membership and order, reconstructedDistanceFunction, interpolationCellPoint, patch evaluation,
mesh geometry.
"""
import re

import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, RefSolver, SolveVofEqu, capi, exact_sphere_alpha, fields, meshmod, oracle_lib
from test_edge_cases import _case as edge_case

pytestmark = pytest.mark.skipif(RefSolver.lib() is None, reason="oracle/_ref/libref_solver.so not built (no reference tree)")


def _smeared_sphere(C_, V, centre=(0.5, 0.62, 0.5), radius=0.15):
    h = np.cbrt(V)
    return np.clip(0.5 - (np.linalg.norm(C_ - np.array(centre), axis=1) - radius) / h, 0.0, 1.0)


def _poly(make, extra=None, cfl=0.25, steps=5):
    return lambda: (make(), dict(extra or {}), _smeared_sphere, fields.rotation_velocity, steps, cfl)


CASES = {
    # the reference's own test case, scaled down: LeVeque deformation of a sphere, hexahedra, LS normals
    "LeVeque hexes 16^3": lambda: (meshmod.hex_block(16), {}, None, fields.leveque_velocity, 8, 0.5),
    "LeVeque hexes, Courant 1.2, nAlphaBounds 10": lambda: (meshmod.hex_block(12), {"nAlphaBounds": 10}, None, fields.leveque_velocity, 5, 1.2),
    "damBreak controls (clip, snapTol, mixedCellTol)": lambda: (meshmod.hex_block(14), {"clip": True, "snapTol": 1e-12, "mixedCellTol": 1e-10,
                                                                                   "nAlphaBounds": 5}, None, fields.leveque_velocity, 6, 0.6),
    "prisms": _poly(lambda: meshmod.prism_mesh(8)),
    "refinement-interface polyhedra": _poly(lambda: meshmod.refined_interface_mesh(8)),
    "warped hexes": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3)),
    "warped hexes, splitWarpedFace": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.25, 7), {"splitWarpedFace": True}, steps=4),
    "Kelvin cells": _poly(lambda: meshmod.kelvin_mesh(8)),
    # the other orientation methods
    "hexes, alphaGrad": lambda: (meshmod.hex_block(14), {"orientationMethod": "alphaGrad"}, None, fields.leveque_velocity, 5, 0.5),
    "warped hexes, alphaGrad": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.2, 3), {"orientationMethod": "alphaGrad"}, steps=4),
    "Kelvin cells, alphaGrad": _poly(lambda: meshmod.kelvin_mesh(6), {"orientationMethod": "alphaGrad"}, steps=3),
    # ... with the NAG test's gradient scheme (Gauss pointLinear: OpenFOAM's, re-stated; the stand-in evaluates it over the whole
    # field as pointLinear.C does, the implementations lazily per face: the same numbers)
    "hexes, alphaGrad, Gauss pointLinear": lambda: (meshmod.hex_block(14), {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"},
                                                    None, fields.leveque_velocity, 5, 0.5),
    "warped hexes, alphaGrad, Gauss pointLinear": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.2, 3),
                                                        {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"}, steps=4),
    "Kelvin cells, alphaGrad, Gauss pointLinear": _poly(lambda: meshmod.kelvin_mesh(6),
                                                        {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"}, steps=3),
    "hexes, isoRDF": lambda: (meshmod.hex_block(16), {"orientationMethod": "isoRDF"}, None, fields.leveque_velocity, 5, 0.5),
    "warped hexes, isoRDF": _poly(lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3), {"orientationMethod": "isoRDF"}, steps=4),
    "Kelvin cells, RDF, one iteration": _poly(lambda: meshmod.kelvin_mesh(7), {"orientationMethod": "RDF", "iterations": 1}, steps=3),
    "prisms, isoRDF, tight tolerances": _poly(lambda: meshmod.prism_mesh(8), {"orientationMethod": "isoRDF", "tol": 1e-10, "relTol": 1e-6,
                                                                           "iterations": 8}, steps=3),
}
for _n in ("every cell cut", "no bounding sweeps", "Courant number 1.5", "2-D: empty front and back", "inflow and outflow patches",
           "no interface: full", "no interface: empty", "single cell", "zero flux"):
    CASES["edge: " + _n] = (lambda n=_n: edge_case(n))
CASES["edge: 2-D, isoRDF"] = lambda: (lambda c: (c[0], {"orientationMethod": "isoRDF"}) + c[2:])(edge_case("2-D: empty front and back"))
CASES["edge: inflow and outflow patches, alphaGrad"] = \
    lambda: (lambda c: (c[0], {"orientationMethod": "alphaGrad"}) + c[2:])(edge_case("inflow and outflow patches"))
CASES["edge: inflow and outflow patches, alphaGrad, Gauss pointLinear"] = \
    lambda: (lambda c: (c[0], {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"}) + c[2:])(edge_case("inflow and outflow patches"))
CASES["edge: 2-D, alphaGrad, Gauss pointLinear"] = \
    lambda: (lambda c: (c[0], {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"}) + c[2:])(edge_case("2-D: empty front and back"))


def _check_against_reference(case, lib, what, sources=False):
    m, extra, alpha0, vel, steps, cfl = CASES[case]()
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, **extra), lib=lib)
    ref = RefSolver(m, s._params)
    # constructor: updateFaceFlatness and its Info line (reconstruction.C:408-442)
    z = s.faceFlatness()
    assert np.array_equal(z, ref.faceFlatness()), "%s: face flatness differs from the reference's updateFaceFlatness" % what
    mn, mx, _avg = re.search(r"face flatness: min/max/avg = ([^/]+)/([^/]+)/(\S+)", ref.ctorLog()).groups()
    assert float(mn) == z.min() and float(mx) == z.max()
    C_, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    a0 = exact_sphere_alpha(m) if alpha0 is None else alpha0(C_, V)
    U0, phi0 = vel(C_), fields.face_flux(Cf, Sf, vel)
    Ub = vel(Cf[m.n_internal_faces:])
    umax = max(np.abs(U0).max(), 1e-30)
    dt = cfl * np.cbrt(V.min()) / umax if umax > 1e-20 else 0.01
    rng = np.random.default_rng(5)
    Sp = -0.3 * rng.random(m.n_cells) if sources else None
    Su = 0.05 * rng.random(m.n_cells) * (a0 > 0) if sources else None
    s.setPhi(phi0)
    s.setAlpha(a0)
    s.setU(U0, Ub)
    ref.setState(a0, phi0, U0, Ub)
    n_mixed = 0
    for k in range(steps):
        s.reconstruct()
        ref.reconstruct()
        mc, st, iN, iD, iC, iS = ref.recon()
        tag = "%s, step %d" % (what, k)
        assert np.array_equal(mc, s.mixedCells()), tag + ": interface-cell list differs from the reference's initialize()"
        assert ("Number of mixed cells = %d" % len(mc)) in ref.log()
        assert np.array_equal(st, s.cellStatus()), tag + ": cut status"
        # interfaceN: compared where the path reads it (the reference also fills cells the step never touches: the Gauss
        # gradient of every cell for alphaGrad)
        gN = s.interfaceN()
        assert np.array_equal(iN[mc], gN[mc]), tag + ": interfaceN differs by %g" % np.abs(iN[mc] - gN[mc]).max()
        assert np.array_equal(iD, s.interfaceD()), tag + ": interfaceD"
        assert np.array_equal(iC, s.field(capi.F_INTERFACE_C)), tag + ": interfaceC"
        assert np.array_equal(iS, s.interfaceS()), tag + ": interfaceS"
        s.advect(dt, Sp=Sp, Su=Su)
        ref.advect(dt, Sp, Su)
        r_alpha, r_alphaPhi, r_alphaB = ref.fields()
        a, ap, ab = s.alpha(), s.alphaPhi(), s.field(capi.F_ALPHA_BOUNDARY)
        assert np.array_equal(a, r_alpha), tag + ": alpha differs from the reference's solveVofEqu by %g" % np.abs(a - r_alpha).max()
        assert np.array_equal(ap, r_alphaPhi), tag + ": alphaPhi differs by %g" % np.abs(ap - r_alphaPhi).max()
        assert np.array_equal(ab, r_alphaB), tag + ": alpha patch values differ"
        nums = re.findall(r"min\(alpha\) = (\S+), max\(alpha\) = 1 \+ (\S+)", ref.log())
        assert len(nums) == 2, ref.log()
        (mn_b, mx_b), (mn_a, mx_a) = [(float(x), float(y)) for x, y in nums]
        assert mn_b == s.info(capi.I_MIN_ALPHA_BEFORE) and mx_b == s.info(capi.I_MAX_ALPHA_M1_BEFORE), tag + ": 'Before' line"
        assert mn_a == s.info(capi.I_MIN_ALPHA_AFTER) and mx_a == s.info(capi.I_MAX_ALPHA_M1_AFTER), tag + ": 'After' line"
        n_mixed = max(n_mixed, len(mc))
    flags = s.info(capi.I_ERROR_FLAGS)
    s.close()
    assert flags == 0
    return n_mixed


_FEW = ("edge: no interface", "edge: single cell")


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_steps_match_reference_solveVofEqu(case):
    n_mixed = _check_against_reference(case, oracle_lib(), "oracle, " + case)
    if not case.startswith(_FEW):
        assert n_mixed > 20


@pytest.mark.parametrize("case", ["damBreak controls (clip, snapTol, mixedCellTol)", "warped hexes", "hexes, isoRDF"])
def test_oracle_steps_with_sources_match_reference_solveVofEqu(case):
    _check_against_reference(case, oracle_lib(), "oracle with Sp/Su, " + case, sources=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_gpu_steps_match_reference_solveVofEqu(case, product):
    """The CUDA library against the reference's own class directly (no oracle in between)."""
    _check_against_reference(case, product, "CUDA, " + case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["damBreak controls (clip, snapTol, mixedCellTol)", "warped hexes"])
def test_gpu_steps_with_sources_match_reference_solveVofEqu(case, product):
    _check_against_reference(case, product, "CUDA with Sp/Su, " + case, sources=True)


# ---- reconstruction::mapAlphaField (reconstruction.C:725-784) ---------------------------------------------------------
def _map_alpha_against_reference(lib, make_mesh, band):
    m = make_mesh()
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, mapAlphaField=True), lib=lib)
    ref = RefSolver(m, s._params, RefSolver.REFINE)
    C_, V = s.field(capi.F_C), s.field(capi.F_V)
    a0 = _smeared_sphere(C_, V, centre=(0.45, 0.5, 0.55), radius=0.23)
    s.setAlpha(a0)
    ref.setState(alpha=a0)
    s.reconstruct()
    ref.reconstruct()
    # a "mapped" field: the planes stay, the volume fractions are disturbed (what mapFields leaves behind after refinement)
    a1 = np.clip(a0 + 0.2 * (np.random.default_rng(11).random(m.n_cells) - 0.5) * ((a0 > 0) & (a0 < 1)), 0.0, 1.0)
    N, D = s.interfaceN(), s.interfaceD()
    s.setAlpha(a1)
    s.setInterface(N, D)
    ref.setState(alpha=a1)
    s.mapAlphaField(*band)
    ref.mapAlphaField(*band)
    r_alpha, _, r_alphaB = ref.fields()
    a = s.alpha()
    sel = (a1 >= band[0]) & (a1 <= band[1])
    assert sel.sum() > 30 and np.any(a[sel] != a1[sel])
    assert np.array_equal(a, r_alpha), "mapAlphaField differs from the reference's by %g" % np.abs(a - r_alpha).max()
    assert np.array_equal(s.field(capi.F_ALPHA_BOUNDARY), r_alphaB)
    s.close()


@pytest.mark.parametrize("kind", ["hexes", "Kelvin cells"])
def test_oracle_map_alpha_field_matches_reference(kind):
    _map_alpha_against_reference(oracle_lib(), {"hexes": lambda: meshmod.hex_block(12), "Kelvin cells": lambda: meshmod.kelvin_mesh(6)}[kind],
                                 (0.01, 0.99))


def test_reference_map_alpha_field_needs_a_refining_mesh():
    """reconstruction.C:727-730: on anything but a dynamicRefineFvMesh the call returns at once."""
    m = meshmod.hex_block(6)
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, mapAlphaField=True), lib=oracle_lib())
    ref = RefSolver(m, s._params, RefSolver.PLAIN)
    a0 = _smeared_sphere(s.field(capi.F_C), s.field(capi.F_V), radius=0.25)
    ref.setState(alpha=a0)
    ref.reconstruct()
    ref.setState(alpha=np.clip(a0 * 0.9, 0, 1))
    ref.mapAlphaField(0.01, 0.99)
    assert np.array_equal(ref.fields()[0], np.clip(a0 * 0.9, 0, 1))
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hexes", "Kelvin cells"])
def test_gpu_map_alpha_field_matches_reference(kind, product):
    _map_alpha_against_reference(product, {"hexes": lambda: meshmod.hex_block(12), "Kelvin cells": lambda: meshmod.kelvin_mesh(6)}[kind],
                                 (0.01, 0.99))


# ---- reconstruction::interface() / subCellFaces() (reconstruction.C:787-891) ------------------------------------------
def _surfaces_against_reference(lib, make_mesh, planar=True):
    m = make_mesh()
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=lib)
    ref = RefSolver(m, s._params)
    C_, V = s.field(capi.F_C), s.field(capi.F_V)
    a0 = _smeared_sphere(C_, V, centre=(0.45, 0.5, 0.55), radius=0.23)
    s.setAlpha(a0)
    ref.setState(alpha=a0)
    s.reconstruct()
    ref.reconstruct()
    # interface(): one polygon per cut cell, cells ascending, points in the order of cutCell::interfacePoints
    pts, off, cells = s.interface()
    rp, roff, rfp, rcells = ref.surface(0)
    assert len(cells) > 40 and np.array_equal(cells, rcells) and np.array_equal(off, roff)
    assert np.array_equal(rfp, np.arange(len(rp))), "reference polygons are identity faces over their own points"
    assert np.abs(pts - rp).max() < 1e-13       # atan2 ordering of coincident points: round-off, not bitwise (svof.h)
    # subCellFaces(): the same cells, closed polyhedra; the reference merges points with OpenFOAM's mergePoints (a stand-in
    # here), so compare what does not depend on the merge: per cell the enclosed volume and the polygon count
    spts, soff, sfp, sfc = s.subCellFaces()
    qp, qoff, qfp, qcells = ref.surface(1)
    assert np.array_equal(np.unique(sfc), qcells)

    def volumes(P, O, F, n_per_cell):
        out, k = [], 0
        for n in n_per_cell:
            vol = 0.0
            for i in range(k, k + n):
                p = P[F[O[i]:O[i + 1]]]
                for j in range(1, len(p) - 1):
                    vol += np.dot(p[0], np.cross(p[j], p[j + 1])) / 6.0
            out.append(vol)
            k += n
        return np.array(out)

    n_mine = np.bincount(sfc)[qcells]
    v_mine = volumes(spts, soff, sfp, n_mine)
    # the reference lists the faces cell by cell in mixed-cell order; its per-cell face count is recovered from the volume
    # check itself: same counts as ours must reproduce alpha * V
    assert len(qoff) - 1 == len(soff) - 1, "same number of sub-cell faces"
    v_ref = volumes(qp, qoff, qfp, n_mine)
    assert np.abs(v_mine - v_ref).max() <= 1e-13 * V.max(), "sub-cell polyhedra differ from the reference's"
    if planar:      # a fan over a warped face is not the cutter's face decomposition
        alphaV = (a0 * V)[qcells]
        assert np.abs(v_ref - alphaV).max() <= 5e-12 * V.max()
    s.close()


@pytest.mark.parametrize("kind", ["hexes", "warped hexes", "Kelvin cells"])
def test_oracle_surfaces_match_reference(kind):
    _surfaces_against_reference(oracle_lib(), {"hexes": lambda: meshmod.hex_block(14),
                                               "warped hexes": lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.2, 3),
                                               "Kelvin cells": lambda: meshmod.kelvin_mesh(6)}[kind], planar=(kind != "warped hexes"))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hexes", "Kelvin cells"])
def test_gpu_surfaces_match_reference(kind, product):
    _surfaces_against_reference(product, {"hexes": lambda: meshmod.hex_block(14), "Kelvin cells": lambda: meshmod.kelvin_mesh(6)}[kind])


# ---- the reference's test case as a whole: plicVofAdvectionFoam's time loop around the reference's own class ----------
class _RefAsSolver:
    """The few members fields.AdvectionDriver uses, served by the reference's solveVofEqu (geometry from `geom`)."""

    def __init__(self, m, geom):
        self.ref = RefSolver(m, geom._params)
        self.mesh, self.nC, self.nF, self.nIF, self.nBF = m, geom.nC, geom.nF, geom.nIF, geom.nBF
        self._geom = geom

    def field(self, which):
        return self._geom.field(which)

    def alpha(self):
        return self.ref.fields()[0]

    def setAlpha(self, a):
        self.ref.setState(alpha=a)

    def setPhi(self, phi):
        self.ref.setState(phi=phi)

    def setU(self, U, Ub):
        self.ref.setState(U=U, Ub=Ub)

    def reconstruct(self):
        self.ref.reconstruct()

    def advect(self, dt):
        self.ref.advect(dt)


def _drive_both(m, lib, a0, t0, dt0, n_steps=None, end_time=None, what=""):
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=lib)
    r = _RefAsSolver(m, s)
    drivers = []
    for x in (s, r):
        x.setAlpha(a0)
        d = fields.AdvectionDriver(x)
        if t0 > 0:
            d.t, d.dt = t0, dt0
            d.phi = d.phi0 * fields.u_factor(t0, d.dt, 6.0)
        drivers.append(d)
    k = n_mixed = 0
    while (k < n_steps) if n_steps is not None else drivers[0].running(end_time):
        for d in drivers:
            d.step(end_time)
        assert drivers[0].dt == drivers[1].dt, "%s step %d: the two time loops chose different time steps" % (what, k)
        mc, st, iN, iD, iC, iS = r.ref.recon()
        assert np.array_equal(mc, s.mixedCells()) and np.array_equal(st, s.cellStatus()), "%s step %d: interface cells" % (what, k)
        assert np.array_equal(iN[mc], s.interfaceN()[mc]) and np.array_equal(iD, s.interfaceD()), "%s step %d: planes" % (what, k)
        ra, rap, _ = r.ref.fields()
        assert np.array_equal(ra, s.alpha()), "%s step %d (t = %.4f): alpha differs from the reference's by %g" % (
            what, k, drivers[0].t, np.abs(ra - s.alpha()).max())
        assert np.array_equal(rap, s.alphaPhi()), "%s step %d: alphaPhi" % (what, k)
        n_mixed = max(n_mixed, len(mc))
        k += 1
    flags = s.info(capi.I_ERROR_FLAGS)
    s.close()
    assert flags == 0
    return k, n_mixed


def test_oracle_full_deformation_run_32_equals_reference_class_every_step():
    """The reference's test case (tutorials/test/plicVofAdvectionFoam: LeVeque deformation, period 6, maxCo = maxAlphaCo = 0.5,
    adaptive dt) at 32^3 from t = 0 to maximum deformation and back (t = 3): under-resolved filaments, slivers, bounding
    chains -- the oracle equals the reference's own solveVofEqu bitwise at every one of the ~300 steps."""
    m = meshmod.hex_block(32)
    steps, n_mixed = _drive_both(m, oracle_lib(), exact_sphere_alpha(m), 0.0, None, end_time=3.0, what="oracle 32^3")
    assert steps > 250 and n_mixed > 1200


@pytest.mark.gpu
def test_gpu_64_from_the_golden_t15_field_equals_reference_class(product):
    """CUDA library vs the reference's own class on the reference's 64^3 case at its hardest instant (golden exact field at
    t = 1.5: 6 094 interface cells in thin sheets), 12 adaptive steps."""
    import os
    N = 64
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "exact_alpha_64.npz"))
    a0 = np.zeros(N ** 3)
    a0[G["full_idx_1.5"]] = 1.0
    a0[G["part_idx_1.5"]] = G["part_val_1.5"]
    steps, n_mixed = _drive_both(meshmod.hex_block(N), product, a0, 1.5, 2.0e-3, n_steps=12, what="CUDA 64^3")
    assert n_mixed > 5000


@pytest.mark.gpu
def test_gpu_deformation_run_32_equals_reference_class(product):
    m = meshmod.hex_block(32)
    steps, n_mixed = _drive_both(m, product, exact_sphere_alpha(m), 0.0, None, n_steps=120, what="CUDA 32^3")
    assert n_mixed > 500


# ---- overset meshes: the cell-type filter of initialize() and the face mask of advect() -------------------------------
def _overset_against_reference(lib, what):
    """A dynamicOversetFvMesh scenario on one component mesh: a block of HOLE cells with a layer of INTERPOLATED cells
    around it, the sphere's interface crossing both; phi masked on the faces of the holes, as the reference demands of
    its callers (advectionTemplates.C:370).  The interface-cell list keeps CALCULATED cells only (reconstruction.C:649-662),
    dVf *= faceMask (advectionTemplates.C:383-396) runs in the reference; everything bitwise, five steps."""
    n = 14
    m = meshmod.hex_block(n)
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS), lib=lib)
    ref = RefSolver(m, s._params, RefSolver.OVERSET)
    C_, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    ijk = np.floor(C_ * n).astype(int)
    lo, hi = np.array([6, 4, 4]), np.array([8, 6, 6])
    hole = np.all((ijk >= lo) & (ijk <= hi), axis=1)
    fringe = np.all((ijk >= lo - 1) & (ijk <= hi + 1), axis=1) & ~hole
    types = np.zeros(m.n_cells, np.int32)
    types[fringe], types[hole] = 1, 2
    a0 = exact_sphere_alpha(m)
    mixed_by_alpha = (a0 > 1e-8) & (a0 < 1 - 1e-8)
    assert (mixed_by_alpha & fringe).sum() > 10 and (mixed_by_alpha & hole).sum() > 5 and (mixed_by_alpha & (types == 0)).sum() > 30
    vel = fields.leveque_velocity
    U0, phi0 = vel(C_), fields.face_flux(Cf, Sf, vel)
    Ub = vel(Cf[m.n_internal_faces:])
    # faceMask = localMin(cellMask): 0 where a hole touches the face
    mask = np.ones(m.n_faces)
    nIF = m.n_internal_faces
    mask[:nIF] = np.minimum(~hole[m.owner[:nIF]], ~hole[m.neighbour[:nIF]])
    mask[nIF:] = ~hole[m.owner[nIF:]]
    phi0 = phi0 * mask
    dt = 0.5 * np.cbrt(V.min()) / np.abs(U0).max()
    s.setPhi(phi0)
    s.setAlpha(a0)
    s.setU(U0, Ub)
    s.setCellTypes(types)
    ref.setState(a0, phi0, U0, Ub)
    ref.setCellTypes(types)
    for k in range(5):
        s.reconstruct()
        ref.reconstruct()
        mc, st, iN, iD, iC, iS = ref.recon()
        assert np.all(types[mc] == 0)
        a_now = s.alpha()
        assert len(mc) < int(((a_now > 1e-8) & (a_now < 1 - 1e-8)).sum()), "the filter must have removed cells"
        assert np.array_equal(mc, s.mixedCells()) and np.array_equal(st, s.cellStatus()), "%s step %d: interface-cell list" % (what, k)
        assert np.array_equal(iN[mc], s.interfaceN()[mc]) and np.array_equal(iD, s.interfaceD())
        assert np.array_equal(iC, s.field(capi.F_INTERFACE_C)) and np.array_equal(iS, s.interfaceS())
        s.advect(dt)
        ref.advect(dt)
        ra, rap, rab = ref.fields()
        assert np.array_equal(ra, s.alpha()), "%s step %d: alpha differs by %g" % (what, k, np.abs(ra - s.alpha()).max())
        assert np.array_equal(rap, s.alphaPhi()), "%s step %d: alphaPhi" % (what, k)
        assert np.array_equal(rab, s.field(capi.F_ALPHA_BOUNDARY))
    # switching the filter off again restores the plain list
    s.setCellTypes(None)
    s.reconstruct()
    a_now = s.alpha()
    assert len(s.mixedCells()) == int(((a_now > 1e-8) & (a_now < 1 - 1e-8)).sum())
    assert s.info(capi.I_ERROR_FLAGS) == 0
    s.close()


def test_oracle_overset_cell_types_match_reference():
    _overset_against_reference(oracle_lib(), "oracle")


@pytest.mark.gpu
def test_gpu_overset_cell_types_match_reference(product):
    _overset_against_reference(product, "CUDA")


@pytest.mark.gpu
def test_gpu_overset_filter_inside_the_captured_step(product, oracle):
    """svof_step_device (CUDA graph) with the cell-type filter: set, stepped, cleared, stepped -- against the oracle."""
    n = 12
    m = meshmod.hex_block(n)
    out = []
    for lib in (oracle, product):
        s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS), lib=lib)
        C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
        types = (np.linalg.norm(C_ - np.array([0.35, 0.35, 0.5]), axis=1) < 0.12).astype(np.int32)
        s.setPhi(fields.face_flux(Cf, Sf))
        s.setAlpha(exact_sphere_alpha(m))
        s.setU(fields.leveque_velocity(C_))
        rec = []
        for k in range(9):
            if k == 2:
                s.setCellTypes(types)
            if k == 6:
                s.setCellTypes(None)
            s.step(0.01)
            rec.append((s.alpha(), s.mixedCells()))
        out.append(rec)
        s.close()
    for k, ((ao, mo), (ag, mg)) in enumerate(zip(*out)):
        assert np.array_equal(mo, mg), "step %d: interface-cell list" % k
        assert np.array_equal(ao, ag), "step %d: alpha" % k
    assert len(out[0][3][1]) < len(out[0][1][1]) or len(out[0][3][1]) < len(out[0][7][1])
