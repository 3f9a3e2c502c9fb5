"""The C-ABI libraries load and export every symbol include/svof.h declares (no compute calls,
no GPU needed), and the configuration surface behaves like the reference's dictionary reads."""
import ctypes as C
import os
import re

import pytest

from common import ROOT, capi, oracle_lib
from geometricvofext_b200.solver import SvofError, make_params


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "svof.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svof_[A-Za-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    hdr = _header_symbols()
    bound = sorted(n for n, _, _ in capi.SYMBOLS)
    assert hdr == bound, "include/svof.h and capi.SYMBOLS differ: %s" % (set(hdr) ^ set(bound))


def test_product_library_exports_every_symbol():
    from geometricvofext_b200 import build
    so = build.build()          # nvcc cross-compiles for sm_100a without a GPU
    lib = capi.load(so)         # raises AttributeError on a missing symbol
    for name in _header_symbols():
        assert hasattr(lib, name)


def test_oracle_library_exports_every_symbol():
    lib = oracle_lib()
    for name in _header_symbols():
        assert hasattr(lib, name)


def test_product_has_no_cpu_fallback():
    """Without a CUDA device svof_create must fail loudly (never route to a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from common import meshmod
    from geometricvofext_b200.solver import SolveVofEqu
    with pytest.raises(SvofError) as e:
        SolveVofEqu(meshmod.hex_block(4), {}, lib=capi.load_product())
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU path" in str(e.value)


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_fvsolution_keys(which):
    lib = oracle_lib() if which == "oracle" else capi.load_product()
    p = make_params(lib, {})
    # defaults of reconstruction.C:502-516 / advection.C:455-457
    assert (p.mixed_cell_tol, p.snap_tol, p.n_alpha_bounds, p.clip, p.split_warped_face) == (1e-8, 0.0, 10, 1, 0)
    assert p.orientation_method == 1 and p.map_alpha_field == 0 and p.write_plic_fields == 0
    # the tutorial dictionary (tutorials/test/plicVofAdvectionFoam/system/fvSolution:21-35)
    p = make_params(lib, {"nAlphaBounds": 3, "snapTol": 0, "clip": "false", "mixedCellTol": 1e-8, "orientationMethod": "LS",
                          "splitWarpedFace": "false", "writePlicFields": "true", "nAlphaSubCycles": 1, "cAlpha": 1,
                          "period": 6.0, "reverseTime": 0.0})
    assert (p.n_alpha_bounds, p.clip, p.write_plic_fields) == (3, 0, 1)
    # isoAdvector names of the north star: surfCellTol aliases mixedCellTol unless that is given; isoFaceTol is recorded only
    p = make_params(lib, {"surfCellTol": 1e-6, "isoFaceTol": 1e-9})
    assert p.mixed_cell_tol == 1e-6 and p.iso_face_tol == 1e-9
    p = make_params(lib, {"mixedCellTol": 1e-10, "surfCellTol": 1e-6})
    assert p.mixed_cell_tol == 1e-10
    for name, code in (("isoAlphaGrad", 1), ("LS", 1), ("alphaGrad", 0), ("isoRDF", 2), ("RDF", 2)):
        assert make_params(lib, {"orientationMethod": name}).orientation_method == code
    with pytest.raises(SvofError) as e:  # reconstruction.C:610-625
        make_params(lib, {"orientationMethod": "smoothIsoRDF"})
    assert e.value.code == capi.ERR_BAD_CONFIG
    with pytest.raises(SvofError):
        make_params(lib, {"noSuchKey": 1})
    with pytest.raises(SvofError):
        make_params(lib, {"clip": "maybe"})


def test_bad_mesh_is_an_error_not_an_abort():
    from common import LEVEQUE_CONTROLS, SolveVofEqu, meshmod
    m = meshmod.hex_block(3)
    m.owner = m.owner.copy()
    m.owner[5] = 10 ** 6
    with pytest.raises(SvofError) as e:
        SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    assert e.value.code == capi.ERR_BAD_MESH
    m = meshmod.hex_block(3)
    m.patches[2].size -= 1   # patches no longer tile the boundary
    with pytest.raises(SvofError):
        SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle_lib())
    s = SolveVofEqu(meshmod.hex_block(3), LEVEQUE_CONTROLS, lib=oracle_lib())
    with pytest.raises(SvofError) as e:
        s.reconstruct()      # alpha not set
    assert e.value.code == capi.ERR_STATE
