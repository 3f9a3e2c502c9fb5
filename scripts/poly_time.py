"""Step time of the polyhedral capacity variants (BASELINE.json configs[2] surrogates): python scripts/poly_time.py
Rotating smeared sphere on (a) triangular prisms, (b) warped hexes, (c) warped hexes with splitWarpedFace."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from geometricvofext_b200 import capi, fields, mesh as meshmod
from geometricvofext_b200.solver import SolveVofEqu

CASES = {
    "Kelvin cells 2x48^3 (14 faces/cell)": (lambda: meshmod.kelvin_mesh(48), {}),
    "prisms 2x64^3": (lambda: meshmod.prism_mesh(64), {}),
    "warped hexes 128^3": (lambda: meshmod.perturb_points(meshmod.hex_block(128), 0.2, 3), {}),
    "warped hexes 128^3, splitWarpedFace": (lambda: meshmod.perturb_points(meshmod.hex_block(128), 0.2, 3), {"splitWarpedFace": True}),
    "2:1 refinement interface 48": (lambda: meshmod.refined_interface_mesh(48), {}),
}
base = {"nAlphaBounds": 3, "snapTol": 0, "clip": False, "mixedCellTol": 1e-8, "orientationMethod": "LS"}
for name, (make, extra) in CASES.items():
    t0 = time.time()
    m = make()
    s = SolveVofEqu(m, dict(base, **extra))
    C_, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    h = np.cbrt(V)
    a0 = np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.62, 0.5]), axis=1) - 0.15) / h, 0.0, 1.0)
    U0 = fields.rotation_velocity(C_)
    phi0 = fields.face_flux(Cf, Sf, fields.rotation_velocity)
    Ub = fields.rotation_velocity(Cf[m.n_internal_faces:])
    dt = 0.25 * np.cbrt(V.min()) / np.abs(U0).max()
    s.setAlpha(a0); s.setPhi(phi0); s.setU(U0, Ub)
    if os.environ.get("PROFILE"):
        for _ in range(3): s.reconstruct(); s.advect(dt)
        s.setOption("profile", 1)
        for _ in range(5): s.reconstruct(); s.advect(dt)
        s.close()
        continue
    for _ in range(5): s.step(dt)
    s.synchronize()
    s.lib.svof_mark(s._h, 0)
    K = 10
    for _ in range(K): s.step(dt)
    s.lib.svof_mark(s._h, 1)
    ms = C.c_double(); s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms)); s.synchronize()
    print("%-40s cells %8d mixed %6d  %.3f ms/step  %.2f G cell-updates/s  err %d  (setup %.0f s)" % (
        name, m.n_cells, int(s.info(capi.I_N_MIXED)), ms.value / K, m.n_cells * K / ms.value / 1e6, int(s.info(capi.I_ERROR_FLAGS)), time.time() - t0))
    s.close()
