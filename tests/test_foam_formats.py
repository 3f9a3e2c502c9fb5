"""OpenFOAM on-disk formats and the case driver: dictionaries, polyMesh, fields,
blockMesh / renumberMesh equivalents and the plicVofAdvectionFoam loop on a case directory (CPU: against the oracle)."""
import os
import sys

import numpy as np
import pytest

from common import ROOT, SolveVofEqu, capi, fields, meshmod, oracle_lib
from geometricvofext_b200 import foamcase, foamfile

REF_TUT = "/root/reference/tutorials"

FV_SOLUTION = """
FoamFile { version 2.0; format ascii; class dictionary; object fvSolution; }
// comment
tolBase 1e-8;
solvers
{
    "alpha.*"
    {
        nAlphaBounds        3;
        snapTol             0;
        clip                false;      /* block
                                           comment */
        mixedCellTol        $tolBase;
        orientationMethod   LS;
        splitWarpedFace     false;
        writePlicFields     true;
        period              6.0;
        reverseTime         0.0;
    };
    "(p_rgh|pcorr)" { solver PCG; tolerance 1e-9; }
    p_rghFinal { $p_rgh; relTol 0; }
    "alpha.oil.*" { nAlphaBounds 5; }
}
PIMPLE {}
"""

CONTROL_DICT = """
FoamFile { version 2.0; format ascii; class dictionary; location "system"; object controlDict; }
application plicVofAdvectionFoam;
startTime 0.0;  endTime %(end)g;  deltaT 0.001;
writeControl adjustableRunTime;  writeInterval %(wi)g;  writeFormat binary;
adjustTimeStep yes;  maxCo 0.5;  maxAlphaCo 0.5;  maxDeltaT 0.2;
functions
{
    probe { type coded; codeWrite #{ Info<< "x" << endl; #}; }
    plicInterface { type surfaces; libs (SimPLIC sampling); surfaceFormat vtp; fields ( alpha.water cellIds );
                    surfaces ( plicSurf { type plicSurface; interpolate false; } ); }
}
"""

BLOCK_MESH_DICT = """
FoamFile { version 2.0; format ascii; class dictionary; object blockMeshDict; }
Nx  %(n)d;  Ny $Nx;  Nz #eval{ $Nx + 0 };
scale 1;
vertices ( (0 0 0) (1 0 0) (1 1 0) (0 1 0) (0 0 1) (1 0 1) (1 1 1) (0 1 1) );
blocks ( hex (0 1 2 3 4 5 6 7) ($Nx $Ny $Nz) simpleGrading (1 1 1) );
edges ( );
boundary
(
    top    { type wall; faces ( (4 5 6 7) ); }
    left   { type wall; faces ( (0 1 5 4) ); }
    back   { type wall; faces ( (0 4 7 3) ); }
    right  { type wall; faces ( (2 3 7 6) ); }
    bottom { type wall; faces ( (0 3 2 1) ); }
    front  { type wall; faces ( (1 2 6 5) ); }
);
"""


def make_case(tmp, n=16, end=0.004, wi=0.002):
    os.makedirs(os.path.join(tmp, "system"))
    for name, text in (("fvSolution", FV_SOLUTION), ("controlDict", CONTROL_DICT % {"end": end, "wi": wi}),
                       ("blockMeshDict", BLOCK_MESH_DICT % {"n": n})):
        with open(os.path.join(tmp, "system", name), "w") as f:
            f.write(text)
    return foamcase.FoamCase(tmp)


def test_dictionary_syntax_and_pattern_lookup(tmp_path):
    hdr, d = foamfile.parse_bytes(FV_SOLUTION.encode())
    assert hdr["class"] == "dictionary" and hdr["format"] == "ascii"
    s = d["solvers"]
    a = s.lookup("alpha.water")
    assert a["nAlphaBounds"] == 3 and a["mixedCellTol"] == 1e-8 and a["clip"] == "false" and a["orientationMethod"] == "LS"
    assert s.lookup("alpha.oil.1")["nAlphaBounds"] == 5          # the LAST matching pattern wins
    assert s.lookup("pcorr")["solver"] == "PCG"
    assert s["p_rghFinal"]["solver"] == "PCG" and s["p_rghFinal"]["relTol"] == 0   # $p_rgh; merged through the pattern
    assert s.lookup("T") is None
    _, c = foamfile.parse_bytes((CONTROL_DICT % {"end": 3, "wi": 0.25}).encode())
    assert c["endTime"] == 3 and c["writeInterval"] == 0.25 and c["functions"]["probe"]["type"] == "coded"
    _, b = foamfile.parse_bytes((BLOCK_MESH_DICT % {"n": 8}).encode())
    assert b["Ny"] == 8 and b["Nz"] == 8.0 and b["blocks"][2] == [8, 8, 8.0] and b["boundary"][4][0] == "bottom"
    with pytest.raises(foamfile.FoamFormatError):
        foamfile.parse_bytes(b"FoamFile { format ascii; class dictionary; } a { b 1; ")


@pytest.mark.parametrize("fmt", ["ascii", "binary"])
def test_polymesh_round_trip(tmp_path, fmt):
    for m in (meshmod.hex_block(4), meshmod.prism_mesh(3)):     # quads only / mixed triangle+quad faces
        d = str(tmp_path / ("case_%s_%d" % (fmt, m.n_cells)))
        foamfile.write_polymesh(m, d, fmt=fmt, patch_types={p.name: "wall" for p in m.patches})
        r = foamfile.read_polymesh(d)
        assert r.n_cells == m.n_cells
        assert np.array_equal(r.points, m.points)                # repr()/binary: bit exact either way
        assert np.array_equal(r.face_offsets, m.face_offsets) and np.array_equal(r.face_points, m.face_points)
        assert np.array_equal(r.owner, m.owner) and np.array_equal(r.neighbour, m.neighbour)
        assert [(p.name, p.start, p.size) for p in r.patches] == [(p.name, p.start, p.size) for p in m.patches]
        assert set(r.meta["patch_types"].values()) == {"wall"}


@pytest.mark.parametrize("fmt", ["ascii", "binary"])
def test_field_round_trip_and_boundary_mapping(tmp_path, fmt):
    m = meshmod.hex_block(4)
    rng = np.random.default_rng(0)
    a = rng.random(m.n_cells)
    bnd = {"top": {"type": "fixedValue", "value": 1.0}, "left": {"type": "inletOutlet", "inletValue": 0.25, "value": 0.25},
           '"(back|right|bottom|front)"': {"type": "zeroGradient"}}
    p = str(tmp_path / "0" / "alpha.water")
    foamfile.write_field(p, "volScalarField", "alpha.water", a, bnd, fmt=fmt, location="0")
    f = foamfile.read_field(p)
    assert f.cls == "volScalarField" and f.dimensions == (0, 0, 0, 0, 0, 0, 0)
    assert np.array_equal(f.internal_array(m.n_cells), a)
    foamfile.apply_alpha_boundary(m, f)
    bc = {q.name: (q.alpha_bc, q.alpha_value) for q in m.patches}
    assert bc["top"] == (capi.BC_FIXED_VALUE, 1.0) and bc["left"] == (capi.BC_INLET_OUTLET, 0.25)
    assert bc["front"] == (capi.BC_ZERO_GRADIENT, 0.0)
    # uniform internal field, vector field, unsupported alpha boundary type
    foamfile.write_field(p, "volScalarField", "alpha.water", 0.5, {'".*"': {"type": "zeroGradient"}}, fmt=fmt)
    assert np.array_equal(foamfile.read_field(p).internal_array(7), np.full(7, 0.5))
    U = rng.random((m.n_cells, 3))
    pu = str(tmp_path / "0" / "U")
    foamfile.write_field(pu, "volVectorField", "U", U, {'".*"': {"type": "fixedValue", "value": np.zeros(3)}}, fmt=fmt,
                         dimensions=(0, 1, -1, 0, 0, 0, 0))
    fu = foamfile.read_field(pu)
    assert np.array_equal(fu.internal_array(m.n_cells), U) and fu.boundary.lookup("top")["value"] == (0.0, 0.0, 0.0)
    foamfile.write_field(p, "volScalarField", "alpha.water", a, {'".*"': {"type": "codedFixedValue"}}, fmt=fmt)
    with pytest.raises(foamfile.FoamFormatError):
        foamfile.apply_alpha_boundary(m, foamfile.read_field(p))


def test_block_mesh_matches_hex_block_and_regroups_patches(tmp_path):
    case = make_case(str(tmp_path / "c"), n=6)
    m, h = case.mesh(), meshmod.hex_block(6)
    assert np.array_equal(m.points, h.points) and np.array_equal(m.neighbour, h.neighbour)
    assert np.array_equal(m.owner, h.owner) and np.array_equal(m.face_points, h.face_points)
    assert [p.name for p in m.patches] == ["top", "left", "back", "right", "bottom", "front"]
    # a dictionary that lists only two sides: the rest goes to defaultFaces (empty), as blockMesh does
    txt = (BLOCK_MESH_DICT % {"n": 4}).replace("type wall; faces ( (0 1 5 4) );", "type patch; faces ( (0 1 5 4) (2 3 7 6) );")
    for name in ("top", "back", "right", "bottom", "front"):
        i = txt.index("    %s " % name)
        txt = txt[:i] + txt[txt.index("\n", i) + 1:]
    _, d = foamfile.parse_bytes(txt.encode())
    m2 = foamcase.block_mesh(d)
    assert [(p.name, p.size, p.kind) for p in m2.patches] == [("left", 32, capi.PATCH_GENERIC), ("defaultFaces", 64, capi.PATCH_EMPTY)]
    s = SolveVofEqu(m2, {}, lib=oracle_lib())
    assert abs(s.field(capi.F_V).sum() - 1.0) < 1e-14             # still a closed, consistently oriented mesh
    s.close()


def test_renumber_mesh_is_cuthill_mckee_and_upper_triangular():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    m = meshmod.hex_block(8)
    n2o = foamcase.cuthill_mckee(m)
    assert np.array_equal(n2o, make_golden.cuthill_mckee_hex(8))   # the restatement the golden mapping was validated with
    r, _ = foamcase.renumber_mesh(m, n2o)
    nIF = r.n_internal_faces
    assert np.all(r.owner[:nIF] < r.neighbour)
    key = r.owner[:nIF].astype(np.int64) * r.n_cells + r.neighbour
    assert np.all(np.diff(key) > 0)                                 # upper-triangular order
    so, sr = SolveVofEqu(m, {}, lib=oracle_lib()), SolveVofEqu(r, {}, lib=oracle_lib())
    assert np.array_equal(sr.field(capi.F_V), so.field(capi.F_V)[n2o])
    assert np.allclose(sr.field(capi.F_C), so.field(capi.F_C)[n2o], atol=1e-15)
    so.close()
    sr.close()


def test_case_run_matches_direct_driver_and_writes_openfoam_fields(tmp_path):
    case = make_case(str(tmp_path / "c"), n=16, end=0.004, wi=0.002)
    mesh = case.mesh()
    a0 = fields.sphere_alpha_quadrature(mesh)
    case.write_alpha(mesh, "0", a0)
    out = foamcase.run_plic_vof_advection(case, lib=oracle_lib())
    assert out["written"] == ["0.002", "0.004"] and out["steps"] >= 2
    # the same loop driven directly
    s = SolveVofEqu(meshmod.hex_block(16), case.alpha_controls(), lib=oracle_lib())
    s.setAlpha(a0)
    drv = fields.AdvectionDriver(s, write_interval=0.002)   # adjustableRunTime: Time::adjustDeltaT lands on the write times
    s.reconstruct()
    for te in (0.002, 0.004):
        while True:
            drv.step()
            if drv.write_now:
                break
        assert abs(drv.t - te) < 1e-12
    assert drv.steps == out["steps"]
    # writePlicFields true: the four reconstruction fields next to alpha, as OpenFOAM fields
    fN = foamfile.read_field(os.path.join(case.dir, "0.004", "interfaceN"))
    fD = foamfile.read_field(os.path.join(case.dir, "0.004", "interfaceD"))
    assert fN.cls == "volVectorField" and fN.dimensions == (0, -1, 0, 0, 0, 0, 0) and fD.dimensions == (0, 1, 0, 0, 0, 0, 0)
    assert np.array_equal(fN.internal, s.interfaceN()) and np.array_equal(fD.internal, s.interfaceD())
    assert fN.boundary.lookup("top")["type"] == "calculated"
    # the plicSurface sampler of the controlDict: polygons written at t = 0 and at every write time
    assert foamcase.plic_surface_functions(case.control_dict) == [("plicInterface", "plicSurf")]
    for t in ("0", "0.002", "0.004"):
        vtk = os.path.join(case.dir, "postProcessing", "plicInterface", t, "plicSurf.vtk")
        assert os.path.isfile(vtk) and "POLYGONS" in open(vtk).read()
    f = foamfile.read_field(os.path.join(case.dir, "0.004", "alpha.water"))
    assert f.header["format"] == "binary" and np.array_equal(f.internal, s.alpha()) and np.array_equal(out["alpha"], s.alpha())
    assert abs(out["volume"] - (a0 / 16 ** 3).sum()) < 1e-15
    s.close()
    # calcVofAdvectionErrors against a stored "exact" field (here: the result itself -> E_s = 0)
    case.write_alpha(mesh, "0.004", out["alpha"], name="alpha.water.exact")
    rows = foamcase.calc_vof_advection_errors(case)
    assert len(rows) == 1 and rows[0][0] == 0.004 and rows[0][4] == 0.0 and rows[0][2] >= -1e-12


@pytest.mark.skipif(not os.path.isdir(REF_TUT), reason="reference tree not present")
def test_every_reference_tutorial_file_parses_and_golden_field_reads_in_place():
    n = 0
    for dp, _, fn in os.walk(REF_TUT):
        for f in fn:
            p = os.path.join(dp, f)
            if os.path.getsize(p) > 5e6:
                continue
            with open(p, "rb") as fh:
                if b"FoamFile" not in fh.read(2000):
                    continue
            foamfile.parse_file(p)
            n += 1
    assert n > 400
    case = foamcase.FoamCase(os.path.join(REF_TUT, "test", "plicVofAdvectionFoam"))
    c = case.alpha_controls()
    assert c["nAlphaBounds"] == 3 and c["mixedCellTol"] == 1e-8 and c["orientationMethod"] == "LS" and c["period"] == 6.0
    assert case.control_dict["maxAlphaCo"] == 0.5 and case.control_dict["writeFormat"] == "binary"
    # the golden t=0 field, read in place in the RENUMBERED cell order, is the committed natural-order fixture permuted
    m, n2o = foamcase.renumber_mesh(foamcase.block_mesh(os.path.join(case.dir, "system", "blockMeshDict")))
    g = foamfile.read_field(os.path.join(REF_TUT, "test", "exactSolutions", "0", "alpha.water.exact"))
    foamfile.apply_alpha_boundary(m, g)
    z = np.load(os.path.join(ROOT, "tests", "golden", "exact_alpha_64.npz"))
    nat = np.zeros(64 ** 3)
    nat[z["full_idx_0"]] = 1.0
    nat[z["part_idx_0"]] = z["part_val_0"]
    assert np.array_equal(g.internal_array(m.n_cells), nat[n2o])


@pytest.mark.gpu
def test_gpu_case_run_on_renumbered_mesh_is_bitwise_the_oracle(tmp_path):
    """renumberMesh changes every label-order-dependent choice of the step (LS stencil order, the Gauss-Seidel
    order of the bounding, face order of the sums): the device must still match the oracle bit for bit."""
    from geometricvofext_b200 import capi as _capi
    outs = []
    for tag, lib in (("gpu", _capi.load_product()), ("cpu", oracle_lib())):
        case = make_case(str(tmp_path / tag), n=24, end=0.012, wi=0.006)
        mesh = case.mesh(renumber=True)
        nat = fields.sphere_alpha_quadrature(meshmod.hex_block(24))
        case.write_alpha(mesh, "0", nat[mesh.meta["cell_map"]])
        outs.append(foamcase.run_plic_vof_advection(case, lib=lib, renumber=True))
    g, c = outs
    assert g["steps"] == c["steps"] and g["written"] == c["written"] == ["0.006", "0.012"]
    assert np.array_equal(g["alpha"], c["alpha"])
    assert abs(g["volume"] - c["volume"]) < 1e-15      # gSum(alpha*V): tree sum on the device, sequential in the oracle


@pytest.mark.gpu
def test_gpu_reference_case_64_first_golden_instant(tmp_path):
    """The reference's Allrun on its 64^3 case up to the first golden instant: blockMesh, renumberMesh, the golden t=0
    field (stored in renumbered order, as the reference's file is), plicVofAdvectionFoam to t = 0.25, calcVofAdvectionErrors."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "exact_alpha_64.npz"))

    def natural(t):
        a = np.zeros(64 ** 3)
        a[z["full_idx_" + t]] = 1.0
        a[z["part_idx_" + t]] = z["part_val_" + t]
        return a
    case = make_case(str(tmp_path / "c"), n=64, end=0.25, wi=0.25)
    mesh = case.mesh(renumber=True)
    n2o = mesh.meta["cell_map"]
    case.write_alpha(mesh, "0", natural("0")[n2o])
    out = foamcase.run_plic_vof_advection(case, renumber=True)
    assert out["written"] == ["0.25"]
    case.write_alpha(mesh, "0.25", natural("0.25")[n2o], name="alpha.water.exact")
    (t, ev, amin, omax, es), = foamcase.calc_vof_advection_errors(case)
    assert t == 0.25 and abs(ev) < 1e-10 and amin > -1e-5 and omax > -1e-5
    assert abs(es - 2.309e-2) / 2.309e-2 < 0.05          # the oracle's pinned E_s(0.25) (tests/test_oracle_golden.py)


def test_cli_block_mesh_writes_a_readable_polymesh(tmp_path):
    case = make_case(str(tmp_path / "c"), n=5)
    assert foamcase.main(["blockMesh", case.dir, "--renumber"]) == 0
    assert case.has_poly_mesh()
    m = case.mesh()                                  # now read back from constant/polyMesh
    ref, _ = foamcase.renumber_mesh(meshmod.hex_block(5))
    assert m.n_cells == 125 and np.array_equal(m.owner, ref.owner) and np.array_equal(m.neighbour, ref.neighbour)
    assert np.array_equal(m.face_points, ref.face_points) and np.array_equal(m.points, ref.points)
    assert [p.name for p in m.patches] == ["top", "left", "back", "right", "bottom", "front"]
    assert set(m.meta["patch_types"].values()) == {"wall"}


def test_case_gradient_scheme_reaches_the_alpha_grad_controls(tmp_path):
    """orientationMethod alphaGrad runs fvc::grad(alpha1, "grad(alpha1)") with the CASE's scheme (reconstruction.C:78): the
    fvSchemes entry of the NAG orientation test (tutorials/test/plicVofOrientationFoam/NAG/system/fvSchemes:32-36) must
    arrive in the solver's controls; LS cases are left alone; `default` applies when grad(alpha1) has no entry of its own."""
    case = make_case(str(tmp_path / "c"), n=8)
    schemes = ('FoamFile { version 2.0; format ascii; class dictionary; location "system"; object fvSchemes; }\n'
               "ddtSchemes { default Euler; }\n"
               "gradSchemes\n{\n    default         none;\n    grad(alpha1)    Gauss pointLinear;\n}\n")
    with open(os.path.join(case.dir, "system", "fvSchemes"), "w") as f:
        f.write(schemes)
    assert case.grad_alpha_scheme() == "Gauss pointLinear"
    assert "gradSchemes" not in case.alpha_controls()          # orientationMethod LS in this fvSolution
    case.fv_solution["solvers"].lookup(case.alpha_name)["orientationMethod"] = "alphaGrad"
    ctl = case.alpha_controls()
    assert ctl["gradSchemes"] == "Gauss pointLinear"
    s = SolveVofEqu(case.mesh(), ctl, lib=oracle_lib())
    assert s._params.orientation_method == 0 and s._params.alpha_grad_scheme == 1   # SVOF_ORIENT_ALPHA_GRAD, Gauss pointLinear
    s.close()
    with open(os.path.join(case.dir, "system", "fvSchemes"), "w") as f:
        f.write(schemes.replace("default         none;", "default         Gauss linear;").replace("    grad(alpha1)    Gauss pointLinear;\n", ""))
    assert foamcase.FoamCase(case.dir).grad_alpha_scheme() == "Gauss linear"
    if os.path.isdir(REF_TUT):
        assert foamcase.FoamCase(os.path.join(REF_TUT, "test/plicVofOrientationFoam/NAG")).alpha_controls()["gradSchemes"] == "Gauss pointLinear"
