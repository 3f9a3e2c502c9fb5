"""The OpenFOAM-side adapter (adapter/solveVofEquB200.{H,C}) is real source: it is compiled here against a minimal
stub of the OpenFOAM headers it uses (adapter/stubs) and RUN the way plicVofAdvectionFoam drives the reference class
(createFields.H:138, plicVof.H:37-41), linked against a library that implements include/svof.h.  The result must be
bitwise what the Python mirror of the same facade gets through the same ABI, the log must carry the three lines users
grep, and the 'reconstruction' registry object must serve the PLIC polygons."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, fields, meshmod, oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AD = os.path.join(ROOT, "adapter")


def _build(lib_path, out_dir, tag):
    exe = os.path.join(out_dir, "test_adapter_" + tag)
    lib_dir, lib_file = os.path.dirname(lib_path), os.path.basename(lib_path)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(AD, "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(AD, "solveVofEquB200.C"), os.path.join(AD, "stubs", "OpenFOAMStub.C"), os.path.join(AD, "test_adapter_main.C"),
           "-o", exe, "-L", lib_dir, "-l:" + lib_file, "-Wl,-rpath," + lib_dir]
    subprocess.check_call(cmd)
    return exe


def _dump(path, m, s, a0, phi, U, dt):
    with open(path, "wb") as f:
        np.array([m.n_points, m.n_faces, m.n_internal_faces, m.n_cells, len(m.patches), m.face_points.size], np.int32).tofile(f)
        m.points.astype(np.float64).tofile(f)
        m.face_offsets.astype(np.int32).tofile(f)
        m.face_points.astype(np.int32).tofile(f)
        m.owner.astype(np.int32).tofile(f)
        m.neighbour.astype(np.int32).tofile(f)
        for p in m.patches:
            np.array([p.start, p.size, p.kind], np.int32).tofile(f)
        for fld in (capi.F_CF, capi.F_SF, capi.F_C, capi.F_V):
            s.field(fld).tofile(f)
        np.array([dt]).tofile(f)
        a0.tofile(f)
        phi.tofile(f)
        U.tofile(f)


def _run_case(lib, lib_path, tmp_path, tag, n=10, steps=4):
    m = meshmod.hex_block(n)
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=lib)
    a0 = fields.sphere_alpha_quadrature(m)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    U, phi = fields.leveque_velocity(C_), fields.face_flux(Cf, Sf)
    dt = 0.25 / n
    fin, fout = os.path.join(str(tmp_path), "in.bin"), os.path.join(str(tmp_path), "out.bin")
    _dump(fin, m, s, a0, phi, U, dt)
    exe = _build(lib_path, str(tmp_path), tag)
    out = subprocess.run([exe, fin, fout, str(steps)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    log = out.stdout
    # the Python mirror through the same ABI
    s.setAlpha(a0)
    s.setPhi(phi)
    s.setU(U)
    for _ in range(steps):
        s.reconstruct()
        s.advect(dt)
    nC, nIF = m.n_cells, m.n_internal_faces
    raw = np.fromfile(fout, dtype=np.float64, count=nC + 3 * nIF)
    alpha, aphi, rphi, rphi2 = raw[:nC], raw[nC:nC + nIF], raw[nC + nIF:nC + 2 * nIF], raw[nC + 2 * nIF:]
    assert np.array_equal(alpha, s.alpha())
    assert np.array_equal(aphi, s.alphaPhi()[:nIF])
    want = (1000.0 - 1.0) * s.alphaPhi()[:nIF] + 1.0 * phi[:nIF]
    assert np.array_equal(rphi, want), "getRhoPhi(dimensionedScalar, dimensionedScalar) = (rho1 - rho2) alphaPhi + rho2 phi"
    assert np.allclose(rphi2, want, rtol=1e-14, atol=0), "getRhoPhi(volScalarField, volScalarField) with uniform densities"
    npoly = int(np.fromfile(fout, dtype=np.int32, offset=8 * (nC + 3 * nIF), count=1)[0])
    s.reconstruct()
    assert npoly == len(s.interface()[2]) and npoly > 0
    nsub = int(np.fromfile(fout, dtype=np.int32, offset=8 * (nC + 3 * nIF) + 4, count=1)[0])
    assert nsub == len(s.subCellFaces()[3]) and nsub > npoly
    assert log.count("SimPLIC::reconstruction: Number of mixed cells = ") == steps + 3
    import re
    kept, n_all, only, restored = re.search(r"overset filter: (\d+) of (\d+) interface cells kept, CALCULATED only (\w+), restored (\d+)", log).groups()
    assert 0 < int(kept) < int(n_all) == int(restored) == npoly and only == "yes"
    assert log.count("SimPLIC::advection: Before conservative bounding: min(alpha) = ") == steps
    assert log.count("SimPLIC::advection: After  conservative bounding: min(alpha) = ") == steps
    assert "SimPLIC::Mesh face flatness: min/max/avg = " in log


def test_adapter_compiles_and_runs_against_the_abi(tmp_path):
    from common import oracle_build
    _run_case(oracle_lib(), oracle_build.build_oracle(), tmp_path, "cpu")


@pytest.mark.gpu
def test_adapter_runs_on_the_device(tmp_path, product):
    _run_case(product, capi.PRODUCT_LIB, tmp_path, "gpu", n=16, steps=5)
