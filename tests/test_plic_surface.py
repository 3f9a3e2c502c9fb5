"""PLIC surface extraction: reconstruction::interface() + cutCell::interfacePoints."""
import os

import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, exact_sphere_alpha, fields, meshmod, oracle_lib
from geometricvofext_b200 import foamfile


def _areas(pts, off):
    out = []
    for i in range(len(off) - 1):
        p = pts[off[i]:off[i + 1]]
        c = p.mean(axis=0)
        a = np.zeros(3)
        for k in range(len(p)):
            a += 0.5 * np.cross(p[k] - c, p[(k + 1) % len(p)] - c)
        out.append(a)
    return np.array(out)


def test_planar_interface_polygons_are_cell_cross_sections():
    m = meshmod.hex_block(8)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    C = s.field(capi.F_C)
    x0 = 0.43
    s.setAlpha(np.clip((x0 - (C[:, 0] - 0.0625)) / 0.125, 0.0, 1.0))   # liquid for x < x0
    s.reconstruct()
    pts, off, cells = s.interface()
    assert len(cells) == 64 and np.array_equal(cells, s.mixedCells()) and np.all(np.diff(off) == 4)
    assert np.abs(pts[:, 0] - x0).max() < 1e-12
    A = _areas(pts, off)
    assert np.abs(np.abs(A[:, 0]) - 0.125 ** 2).max() < 1e-14 and np.abs(A[:, 1:]).max() < 1e-15
    s.close()


def test_sphere_surface_area_and_planarity(tmp_path):
    m = meshmod.hex_block(32)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    s.setAlpha(exact_sphere_alpha(m))
    s.reconstruct()
    pts, off, cells = s.interface()
    st, mc = s.cellStatus(), s.mixedCells()
    assert np.array_equal(cells, mc[st == 0]) and off[-1] == len(pts)
    nv = np.diff(off)
    assert nv.min() >= 3 and nv.max() <= 6
    N, D = s.interfaceN(), s.interfaceD()
    face_of_pt = np.repeat(np.arange(len(cells)), nv)
    assert np.abs(np.sum(pts * N[cells][face_of_pt], axis=1) + D[cells][face_of_pt]).max() < 1e-12   # on their planes
    A = _areas(pts, off)
    area = np.linalg.norm(A, axis=1).sum()
    assert abs(area - 4 * np.pi * 0.15 ** 2) / (4 * np.pi * 0.15 ** 2) < 0.03
    # orientation: polygon normals agree with the stored interface area vectors up to sign convention and size
    S = s.interfaceS()[cells]
    assert np.abs(np.linalg.norm(A, axis=1) - np.linalg.norm(S, axis=1)).max() < 1e-12
    p = foamfile.write_vtk_polydata(str(tmp_path / "s.vtk"), pts, off, {"cellIds": cells})
    txt = open(p).read().split("\n")
    assert txt[4] == "POINTS %d double" % len(pts) and ("POLYGONS %d %d" % (len(cells), len(cells) + len(pts))) in txt
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["hex", "polyhedra", "warped"])
def test_gpu_surface_matches_oracle(case):
    from geometricvofext_b200 import capi as _capi
    m = {"hex": lambda: meshmod.hex_block(24), "polyhedra": lambda: meshmod.refined_interface_mesh(8),
         "warped": lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3)}[case]()
    so, sg = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib()), SolveVofEqu(m, LEVEQUE_CONTROLS, lib=_capi.load_product())
    C, V = so.field(capi.F_C), so.field(capi.F_V)
    r = np.linalg.norm(C - np.array([0.4, 0.45, 0.5]), axis=1)
    a0 = np.clip(0.5 + (0.27 - r) / np.cbrt(V), 0.0, 1.0)
    for s in (so, sg):
        s.setAlpha(a0)
        s.reconstruct()
    (po, oo, co), (pg, og, cg) = so.interface(), sg.interface()
    assert len(co) > 50 and np.array_equal(co, cg) and np.array_equal(oo, og)
    assert np.abs(po - pg).max() < 1e-13      # atan2 tie-breaks between coincident points: round-off, not bitwise (svof.h)
    so.close()
    sg.close()


# ---- reconstruction::subCellFaces() (reconstruction.C:838-891) ------------------------------------------------------
def _subcell_checks(s, m):
    """every cut cell's sub-cell faces form a closed, outward-oriented polyhedron whose volume is alpha*V"""
    pts, off, fp, fc = s.subCellFaces()
    assert len(off) == len(fc) + 1 and off[-1] == len(fp)
    alpha, V = s.alpha(), s.field(capi.F_V)
    mixed, status = s.mixedCells(), s.cellStatus()
    assert np.array_equal(np.unique(fc), mixed[status == 0]), "one sub-cell per cut cell, ascending"
    worst = 0.0
    for c in np.unique(fc)[:400]:
        faces = [fp[off[i]:off[i + 1]] for i in np.nonzero(fc == c)[0]]
        edges = {}
        vol = 0.0
        for f in faces:
            assert len(f) >= 3
            p = pts[f]
            for a, b in zip(f, np.roll(f, -1)):
                edges[(a, b)] = edges.get((a, b), 0) + 1
            # divergence theorem with a fan about the first vertex: V = 1/6 sum (p0 . (pi x pi+1))
            for i in range(1, len(f) - 1):
                vol += np.dot(p[0], np.cross(p[i], p[i + 1])) / 6.0
        for (a, b), k in edges.items():
            assert k == 1 and edges.get((b, a), 0) == 1, "closed surface: every edge is shared by two faces, in opposite directions"
        assert vol > 0, "outward orientation"
        worst = max(worst, abs(vol - alpha[c] * V[c]) / V[c])
    assert worst <= 5e-12, "sub-cell volume differs from alpha*V by %g V" % worst
    return pts, off, fp, fc


def test_subcell_faces_oracle_closed_polyhedra(oracle):
    m = meshmod.hex_block(10)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle)
    s.setAlpha(fields.sphere_alpha_quadrature(m))
    s.reconstruct()
    pts, off, fp, fc = _subcell_checks(s, m)
    assert len(fc) > 100


def test_subcell_faces_oracle_polyhedral_cells(oracle):
    m = meshmod.kelvin_mesh(5)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle)
    C_, V = s.field(capi.F_C), s.field(capi.F_V)
    s.setAlpha(np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.55, 0.5]), axis=1) - 0.22) / np.cbrt(V), 0.0, 1.0))
    s.reconstruct()
    _subcell_checks(s, m)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "kelvin"])
def test_subcell_faces_gpu_parity(oracle, product, kind):
    m = meshmod.hex_block(16) if kind == "hex" else meshmod.kelvin_mesh(6)
    out = []
    for lib in (oracle, product):
        s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=lib)
        if kind == "hex":
            a0 = fields.sphere_alpha_quadrature(m)
        else:
            C_, V = s.field(capi.F_C), s.field(capi.F_V)
            a0 = np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.55, 0.5]), axis=1) - 0.22) / np.cbrt(V), 0.0, 1.0)
        s.setAlpha(a0)
        s.reconstruct()
        out.append(_subcell_checks(s, m) if lib is product else s.subCellFaces())
    (po, oo, fo, co), (pg, og, fg, cg) = out
    assert np.array_equal(oo, og) and np.array_equal(fo, fg) and np.array_equal(co, cg), "same topology"
    assert np.abs(po - pg).max() <= 1e-13
