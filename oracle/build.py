"""Build recipe for the test oracle (TEST INFRASTRUCTURE ONLY).

  oracle/_build/libsvof_oracle.so  -- CPU restatement of the reference algorithm
                                      (oracle/*.hpp, svof_oracle.cpp), g++ only.
  oracle/_ref/libref_overlap.so    -- the reference's own vendored overlap.hpp +
                                      Eigen compiled from /root/reference (only
                                      when that tree is present, i.e. in the build
                                      container; the GPU box uses the prebuilt file).
  oracle/_ref/libref_cut.so        -- the reference's own cutFace.{H,C} and cutCell.{H,C}
                                      (src/SimPLIC/cut), compiled UNMODIFIED from
                                      /root/reference against oracle/of_stub/ (a stand-in
                                      for the few OpenFOAM types they use) + ref_cut.cpp
                                      (C entry points).  Pins the oracle's restatement of
                                      the geometric core to the reference's statements.

  oracle/_ref/libref_advect.so     -- the reference's own advection.{H,C} + advectionTemplates.C (src/SimPLIC/advection)
                                      and cutFace.{H,C}, compiled UNMODIFIED from /root/reference against
                                      oracle/of_stub_adv/ (a larger stand-in: GeometricField algebra, patches,
                                      upwind, fvc::surfaceIntegrate, bitSet, zeroField ...) + ref_advect.cpp.
                                      Pins the oracle's restatement of advect() to the reference's statements.

  oracle/_ref/libref_solver.so     -- the reference's WHOLE solveVofEqu class: solveVofEqu.{H,C} + solveVofEquTemplates.C,
                                      reconstruction.{H,C}, advection.{H,C} + advectionTemplates.C, cutFace.{H,C},
                                      cutCell.{H,C} (every file of src/SimPLIC outside sampling/), compiled UNMODIFIED from
                                      /root/reference against oracle/of_stub_rec/ (+ leastSquareGrad, Gauss-linear fvc::grad,
                                      zoneDistribute, reconstructedDistanceFunction, IOdictionary ...) + ref_solver.cpp.
                                      Pins reconstruct() and the chained reconstruct + advect steps to the reference's statements.

Both directories are git-ignored and are NOT gpurun-ignored.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_APP = "/root/reference/applications/test/calcExactVofFieldForSphericalShapeInHexMesh"
ORACLE_SO = os.path.join(HERE, "_build", "libsvof_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_overlap.so")
REF_CUT = "/root/reference/src/SimPLIC/cut"
REF_CUT_SO = os.path.join(HERE, "_ref", "libref_cut.so")
REF_ADV = "/root/reference/src/SimPLIC/advection"
REF_ADV_SO = os.path.join(HERE, "_ref", "libref_advect.so")
REF_REC = "/root/reference/src/SimPLIC/reconstruction"
REF_SOLVER = "/root/reference/src/SimPLIC/solveVofEqu"
REF_SOLVER_SO = os.path.join(HERE, "_ref", "libref_solver.so")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    srcs = [os.path.join(HERE, f) for f in
            ("svof_oracle.cpp", "ora_vec.hpp", "ora_mesh.hpp", "ora_cut.hpp", "ora_solver.hpp")]
    srcs.append(os.path.join(HERE, "..", "include", "svof.h"))
    if force or _stale(ORACLE_SO, srcs):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
               "-Wl,-Bsymbolic", "-Wall", "-o", ORACLE_SO, srcs[0]]
        subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref(force=False):
    """Compile the reference's overlap.hpp where it lies; returns path or None."""
    if not os.path.isdir(REF_APP):
        return REF_SO if os.path.exists(REF_SO) else None
    src = os.path.join(HERE, "ref_sphere_overlap.cpp")
    if force or _stale(REF_SO, [src]):
        os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I", REF_APP, "-o", REF_SO, src]
        subprocess.check_call(cmd)
    return REF_SO


def build_ref_cut(force=False):
    """Compile the reference's cutFace.C / cutCell.C where they lie, against the OpenFOAM stand-in; path or None."""
    if not os.path.isdir(REF_CUT):
        return REF_CUT_SO if os.path.exists(REF_CUT_SO) else None
    wrapper = os.path.join(HERE, "ref_cut.cpp")
    stub = os.path.join(HERE, "of_stub", "OpenFOAMCutStub.H")
    ref_srcs = [os.path.join(REF_CUT, "cutFace", "cutFace.C"), os.path.join(REF_CUT, "cutCell", "cutCell.C")]
    if force or _stale(REF_CUT_SO, [wrapper, stub, os.path.join(HERE, "..", "include", "svof.h")] + ref_srcs):
        os.makedirs(os.path.dirname(REF_CUT_SO), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w",
               "-I", os.path.join(HERE, "of_stub"), "-I", os.path.join(REF_CUT, "cutFace"), "-I", os.path.join(REF_CUT, "cutCell"),
               "-o", REF_CUT_SO, wrapper] + ref_srcs
        subprocess.check_call(cmd)
    return REF_CUT_SO


def build_ref_advect(force=False):
    """Compile the reference's advection.C / advectionTemplates.C / cutFace.C where they lie, against the larger
    OpenFOAM stand-in (oracle/of_stub_adv/); path or None."""
    if not os.path.isdir(REF_ADV):
        return REF_ADV_SO if os.path.exists(REF_ADV_SO) else None
    wrapper = os.path.join(HERE, "ref_advect.cpp")
    stubs = [os.path.join(HERE, "of_stub_adv", f) for f in ("OpenFOAMAdvectStub.H", "reconstruction.H")] + \
            [os.path.join(HERE, "of_stub", "OpenFOAMCutStub.H")]
    ora = [os.path.join(HERE, f) for f in ("ora_vec.hpp", "ora_mesh.hpp", "ora_cut.hpp", "ora_solver.hpp")]
    ref_srcs = [os.path.join(REF_ADV, "advection.C"), os.path.join(REF_CUT, "cutFace", "cutFace.C")]
    ref_hdrs = [os.path.join(REF_ADV, "advection.H"), os.path.join(REF_ADV, "advectionTemplates.C")]
    if force or _stale(REF_ADV_SO, [wrapper, os.path.join(HERE, "..", "include", "svof.h")] + stubs + ora + ref_srcs + ref_hdrs):
        os.makedirs(os.path.dirname(REF_ADV_SO), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-DNoRepository",
               "-I", os.path.join(HERE, "of_stub_adv"), "-I", REF_ADV, "-I", os.path.join(REF_CUT, "cutFace"),
               "-I", os.path.join(REF_CUT, "cutCell"), "-I", HERE, "-o", REF_ADV_SO, wrapper] + ref_srcs
        subprocess.check_call(cmd)
    return REF_ADV_SO


def build_ref_solver(force=False):
    """Compile the reference's solveVofEqu.C / reconstruction.C / advection.C / cutFace.C / cutCell.C where they lie, against
    oracle/of_stub_rec/ (which builds on the two smaller stand-ins); path or None."""
    if not os.path.isdir(REF_REC):
        return REF_SOLVER_SO if os.path.exists(REF_SOLVER_SO) else None
    wrapper = os.path.join(HERE, "ref_solver.cpp")
    stubs = [os.path.join(HERE, "of_stub_rec", "OpenFOAMReconStub.H"), os.path.join(HERE, "of_stub_adv", "OpenFOAMAdvectStub.H"),
             os.path.join(HERE, "of_stub", "OpenFOAMCutStub.H")]
    ora = [os.path.join(HERE, f) for f in ("ora_vec.hpp", "ora_mesh.hpp", "ora_cut.hpp", "ora_solver.hpp")]
    ref_srcs = [os.path.join(REF_SOLVER, "solveVofEqu.C"), os.path.join(REF_REC, "reconstruction.C"),
                os.path.join(REF_ADV, "advection.C"), os.path.join(REF_CUT, "cutFace", "cutFace.C"),
                os.path.join(REF_CUT, "cutCell", "cutCell.C")]
    ref_hdrs = [os.path.join(REF_SOLVER, "solveVofEqu.H"), os.path.join(REF_SOLVER, "solveVofEquTemplates.C"),
                os.path.join(REF_REC, "reconstruction.H"), os.path.join(REF_ADV, "advection.H"),
                os.path.join(REF_ADV, "advectionTemplates.C")]
    if force or _stale(REF_SOLVER_SO, [wrapper, os.path.join(HERE, "..", "include", "svof.h")] + stubs + ora + ref_srcs + ref_hdrs):
        os.makedirs(os.path.dirname(REF_SOLVER_SO), exist_ok=True)
        # the reference's own directories come first: reconstruction.H must be the real one
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-DNoRepository",
               "-I", REF_REC, "-I", REF_SOLVER, "-I", REF_ADV, "-I", os.path.join(REF_CUT, "cutFace"),
               "-I", os.path.join(REF_CUT, "cutCell"), "-I", os.path.join(HERE, "of_stub_rec"), "-I", HERE,
               "-o", REF_SOLVER_SO, wrapper] + ref_srcs
        subprocess.check_call(cmd)
    return REF_SOLVER_SO


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
    print(build_ref_cut(force="--force" in sys.argv))
    print(build_ref_advect(force="--force" in sys.argv))
    print(build_ref_solver(force="--force" in sys.argv))
