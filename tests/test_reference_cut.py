"""Pin the oracle (and, on a GPU, the CUDA product) to the REFERENCE's own cutFace / cutCell code.

oracle/_ref/libref_cut.so is src/SimPLIC/cut/cutFace/cutFace.{H,C} + cutCell/cutCell.{H,C}, compiled unmodified from the
reference tree against a small stand-in for the OpenFOAM types they use (oracle/of_stub/OpenFOAMCutStub.H, recipe
oracle/build.py:build_ref_cut).  The bar is BITWISE equality of every output of
    cutFace::calcSubFace (cutFace.C:136-259), cutCell::calcSubCell (cutCell.C:343-542),
    cutCell::findSignedDistance (cutCell.C:611-799, with and without splitWarpedFace),
    cutFace::timeIntegratedFaceFlux / timeIntegratedArea (cutFace.C:262-508)
on random batteries over hexahedra, warped hexahedra, prisms, refinement-interface polyhedra and 14-face Kelvin cells,
including planes through vertices, axis-aligned normals, tiny volume fractions and stationary interfaces.
cutCell::interfacePoints (cutCell.C:545-608) is compared to 1e-13 (atan2 ordering of coincident points).
"""
import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, RefCut, SolveVofEqu, capi, have_gpu, meshmod, oracle_lib

pytestmark = pytest.mark.skipif(RefCut.lib() is None, reason="oracle/_ref/libref_cut.so not built (no reference tree)")

MESHES = {
    "hexes": lambda: meshmod.hex_block(8),
    "warped hexes": lambda: meshmod.perturb_points(meshmod.hex_block(8), 0.2, 3),
    "prisms": lambda: meshmod.prism_mesh(6),
    "refinement-interface polyhedra": lambda: meshmod.refined_interface_mesh(6),
    "Kelvin cells": lambda: meshmod.kelvin_mesh(6),
}


def _unit(v):
    return v / np.linalg.norm(v, axis=1)[:, None]


def _battery(m, s, rng, n):
    """Random inputs of the four primitives on mesh m (s: any SolveVofEqu of it, for the geometry)."""
    C_, Cf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_V)
    cells = rng.integers(0, m.n_cells, n).astype(np.int32)
    h = np.cbrt(V[cells])
    nrm = _unit(rng.normal(size=(n, 3)))
    nrm[:80] = np.eye(3)[rng.integers(0, 3, 80)] * rng.choice([-1.0, 1.0], 80)[:, None]  # planes along faces / through vertices
    Dc = -(nrm * (C_[cells] + rng.uniform(-0.7, 0.7, size=(n, 3)) * h[:, None])).sum(1)
    # planes exactly through a vertex of the cell
    for i in range(80, 120):
        f = np.nonzero(m.owner == cells[i])[0][0]
        p = m.points[m.face_points[m.face_offsets[f]]]
        Dc[i] = -(nrm[i] * p).sum()
    al = rng.uniform(1e-8, 1 - 1e-8, n)
    al[:20] = 10.0 ** rng.uniform(-8, -3, 20)
    al[20:40] = 1.0 - 10.0 ** rng.uniform(-8, -3, 20)
    faces = rng.integers(0, m.n_faces, n).astype(np.int32)
    hf = np.cbrt(V[m.owner[faces]])
    Df = -(nrm * (Cf[faces] + rng.uniform(-0.6, 0.6, size=(n, 3)) * hf[:, None])).sum(1)
    Un0 = rng.uniform(-2, 2, n)
    Un0[:30] = 0.0
    Un0[30:40] = 1e-16
    phi = rng.uniform(-1, 1, n) * hf * hf
    phi[40:45] = 0.0
    return dict(cells=cells, nrm=nrm, Dc=Dc, al=al, faces=faces, Df=Df, Un0=Un0, phi=phi, dt=0.3 * float(hf.min()))


def _check_primitives(m, impl, ref_plain, ref_split, split, rng, n, what):
    b = _battery(m, impl, rng, n)
    if not split:
        for name, a, r in zip(("status", "VOF", "subVolume", "interfaceCentre", "interfaceArea"),
                              impl.cutCells(b["cells"], b["nrm"], b["Dc"]), ref_plain.cutCells(b["cells"], b["nrm"], b["Dc"])):
            assert np.array_equal(a, r), "%s: calcSubCell %s differs from the reference" % (what, name)
    ref = ref_split if split else ref_plain
    for name, a, r in zip(("status", "D", "C", "S"), impl.findSignedDistance(b["cells"], b["al"], b["nrm"]),
                          ref.findSignedDistance(b["cells"], b["al"], b["nrm"])):
        assert np.array_equal(a, r), "%s: findSignedDistance %s differs from the reference" % (what, name)
    a = impl.faceFluxes(b["faces"], b["nrm"], b["Df"], b["Un0"], b["dt"], b["phi"])
    r = ref.faceFluxes(b["faces"], b["nrm"], b["Df"], b["Un0"], b["dt"], b["phi"])
    assert np.array_equal(a, r), "%s: timeIntegratedFaceFlux differs from the reference" % what


def _check_polygons(impl, ref, rng, n):
    for nv in (3, 4, 5, 7):
        ang = np.sort(rng.uniform(0, 2 * np.pi, size=(n, nv)), axis=1)
        r = rng.uniform(0.5, 1.0, size=(n, nv))
        pts = np.stack([r * np.cos(ang), r * np.sin(ang), 0.05 * rng.normal(size=(n, nv))], axis=2)
        nrm = _unit(rng.normal(size=(n, 3)))
        D = rng.uniform(-0.8, 0.8, n)
        D[:50] = -(nrm[:50] * pts[:50, 0]).sum(1)  # plane exactly through a vertex
        for name, a, b in zip(("status", "centre", "area"), impl.cutFaces(pts, nrm, D), ref.cutFaces(pts, nrm, D)):
            assert np.array_equal(a, b), "calcSubFace %s differs from the reference (%d vertices)" % (name, nv)


@pytest.mark.parametrize("case", list(MESHES))
@pytest.mark.parametrize("split", [False, True])
def test_oracle_matches_reference_cut_classes(case, split):
    m = MESHES[case]()
    so = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, splitWarpedFace=split), lib=oracle_lib())
    ref_plain, ref_split = RefCut(m, so, split=False), RefCut(m, so, split=True)
    rng = np.random.default_rng(11)
    _check_primitives(m, so, ref_plain, ref_split, split, rng, 1500, "oracle, " + case)
    if not split:
        _check_polygons(so, ref_plain, rng, 1500)


def test_oracle_interface_polygon_matches_reference():
    """cutCell::interfacePoints of the cut cells of a reconstructed sphere, cell by cell."""
    from common import exact_sphere_alpha
    m = meshmod.hex_block(16)
    so = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    ref = RefCut(m, so)
    so.setAlpha(exact_sphere_alpha(m))
    so.reconstruct()
    pts, off, cells = so.interface()
    N, D = so.interfaceN(), so.interfaceD()
    assert len(cells) > 80
    for i, c in enumerate(cells):
        r = ref.interfacePoints(c, N[c], D[c])
        mine = pts[off[i]:off[i + 1]]
        assert mine.shape == r.shape, "cell %d: %d points, reference %d" % (c, len(mine), len(r))
        assert np.abs(mine - r).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(MESHES))
@pytest.mark.parametrize("split", [False, True])
def test_gpu_matches_reference_cut_classes(case, split):
    """The CUDA library against the reference's own classes, without the oracle in between."""
    assert have_gpu()
    m = MESHES[case]()
    sg = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, splitWarpedFace=split), lib=capi.load_product())
    ref_plain, ref_split = RefCut(m, sg, split=False), RefCut(m, sg, split=True)
    rng = np.random.default_rng(12)
    _check_primitives(m, sg, ref_plain, ref_split, split, rng, 3000, "CUDA, " + case)
    if not split:
        _check_polygons(sg, ref_plain, rng, 3000)
    assert sg.info(capi.I_GPU_LAUNCHES) > 0
