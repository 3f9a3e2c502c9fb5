"""Generate tests/golden/reference_solver_steps.npz: outputs of the REFERENCE's own solveVofEqu class.

Run in the build container (needs /root/reference, from which oracle/build.py:build_ref_solver compiles the reference's
src/SimPLIC -- solveVofEqu, reconstruction, advection, cutFace, cutCell -- unmodified against the OpenFOAM stand-in
oracle/of_stub_rec/ into oracle/_ref/libref_solver.so):
    python tests/golden/make_reference_solver_golden.py

For every case below the reference object runs reconstruct() + advect(dt) step after step on its own state; stored per step:
the interface-cell list, the cut status, interfaceN / interfaceD at the interface cells, alpha (sparse: the cells that are
not exactly 0 or 1 + the list of full cells) and alphaPhi (sparse: non-zero faces).  tests/test_reference_golden.py replays the
same cases with the oracle (CPU) and with the CUDA library (GPU) and demands bitwise equality -- so the pin also holds where
oracle/_ref/libref_solver.so is absent.  Meshes and fields are re-created by the same deterministic generators
(geometricvofext_b200/mesh.py, fields.py); the initial alpha is stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from common import LEVEQUE_CONTROLS, RefSolver, SolveVofEqu, capi, fields, meshmod, oracle_lib  # noqa: E402

OUT = os.path.join(HERE, "reference_solver_steps.npz")


def smeared_sphere(C_, V, centre=(0.5, 0.62, 0.5), radius=0.15):
    return np.clip(0.5 - (np.linalg.norm(C_ - np.array(centre), axis=1) - radius) / np.cbrt(V), 0.0, 1.0)


# name -> (mesh, extra controls, sphere centre / radius of the initial field, velocity, steps, Courant number)
CASES = {
    "hexes_LS": (lambda: meshmod.hex_block(16), {}, ((0.35, 0.35, 0.35), 0.15), fields.leveque_velocity, 6, 0.5),
    "hexes_Co1.2_nAlphaBounds10": (lambda: meshmod.hex_block(12), {"nAlphaBounds": 10}, ((0.35, 0.35, 0.35), 0.15), fields.leveque_velocity, 4, 1.2),
    "hexes_damBreak_controls": (lambda: meshmod.hex_block(14), {"clip": True, "snapTol": 1e-12, "mixedCellTol": 1e-10, "nAlphaBounds": 5},
                                ((0.35, 0.35, 0.35), 0.15), fields.leveque_velocity, 5, 0.6),
    "warped_hexes_split": (lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.25, 7), {"splitWarpedFace": True},
                           ((0.5, 0.62, 0.5), 0.15), fields.rotation_velocity, 3, 0.25),
    "prisms": (lambda: meshmod.prism_mesh(8), {}, ((0.5, 0.62, 0.5), 0.15), fields.rotation_velocity, 4, 0.25),
    "refinement_polyhedra": (lambda: meshmod.refined_interface_mesh(8), {}, ((0.5, 0.62, 0.5), 0.15), fields.rotation_velocity, 4, 0.25),
    "kelvin_cells": (lambda: meshmod.kelvin_mesh(7), {}, ((0.5, 0.62, 0.5), 0.15), fields.rotation_velocity, 3, 0.25),
    "hexes_alphaGrad_pointLinear": (lambda: meshmod.hex_block(14), {"orientationMethod": "alphaGrad", "gradSchemes": "Gauss pointLinear"},
                                    ((0.35, 0.35, 0.35), 0.15), fields.leveque_velocity, 4, 0.5),
    "warped_hexes_alphaGrad": (lambda: meshmod.perturb_points(meshmod.hex_block(10), 0.2, 3), {"orientationMethod": "alphaGrad"},
                               ((0.5, 0.62, 0.5), 0.15), fields.rotation_velocity, 3, 0.25),
    "hexes_isoRDF": (lambda: meshmod.hex_block(16), {"orientationMethod": "isoRDF"}, ((0.35, 0.35, 0.35), 0.15), fields.leveque_velocity, 4, 0.5),
    "kelvin_cells_RDF_1_iteration": (lambda: meshmod.kelvin_mesh(6), {"orientationMethod": "RDF", "iterations": 1}, ((0.5, 0.62, 0.5), 0.2),
                                     fields.rotation_velocity, 3, 0.25),
}


def case_inputs(name, s):
    """(a0, U, Ub, phi, dt) of a case on a solver `s` of its mesh -- shared with tests/test_reference_golden.py"""
    _, _, (centre, radius), vel, _, cfl = CASES[name]
    m = s.mesh
    C_, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    a0 = smeared_sphere(C_, V, centre, radius)
    U0, phi0 = vel(C_), fields.face_flux(Cf, Sf, vel)
    Ub = vel(Cf[m.n_internal_faces:])
    dt = cfl * np.cbrt(V.min()) / max(np.abs(U0).max(), 1e-30)
    return a0, U0, Ub, phi0, dt


def main():
    if RefSolver.lib() is None:
        raise SystemExit("oracle/_ref/libref_solver.so cannot be built here (no /root/reference)")
    out = {}
    for name, (make, extra, _, _, steps, _) in CASES.items():
        m = make()
        geom = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, **extra), lib=oracle_lib())   # geometry + parameter block only
        ref = RefSolver(m, geom._params)
        a0, U0, Ub, phi0, dt = case_inputs(name, geom)
        ref.setState(a0, phi0, U0, Ub)
        out[name + "/a0"] = a0
        out[name + "/dt"] = np.array([dt])
        out[name + "/sizes"] = np.array([m.n_cells, m.n_faces, m.n_points, steps])
        for k in range(steps):
            ref.reconstruct()
            mc, st, iN, iD, _, _ = ref.recon()
            ref.advect(dt)
            a, ap, _ = ref.fields()
            part = np.nonzero((a != 0.0) & (a != 1.0))[0].astype(np.int32)
            nz = np.nonzero(ap != 0.0)[0].astype(np.int32)
            pre = "%s/%d/" % (name, k)
            out[pre + "mixed"], out[pre + "status"] = mc, st
            out[pre + "N"], out[pre + "D"] = iN[mc], iD[mc]
            out[pre + "full"] = np.nonzero(a == 1.0)[0].astype(np.int32)
            out[pre + "part_idx"], out[pre + "part_val"] = part, a[part]
            out[pre + "aphi_idx"], out[pre + "aphi_val"] = nz, ap[nz]
        geom.close()
        print("%-34s cells %6d  steps %d  interface cells %d" % (name, m.n_cells, steps, len(mc)))
    np.savez_compressed(OUT, **out)
    print(OUT, "%.2f MB" % (os.path.getsize(OUT) / 1e6))


if __name__ == "__main__":
    main()
