"""svof_step_host diagnostics: bytes moved and wall time per call (python scripts/e2e_diag.py [N] [calls]); SVOF_LIB selects the library."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
lib = s.lib
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
phi_h, U_h, Ub_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3)), capi.pinned_array(lib, (max(s.nBF, 1), 3))
a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
phi_h[:] = phi; U_h[:] = U; Ub_h[:] = 0
s.setAlpha(a0)
for k in range(calls):
    fk = 1.0 - 1e-3 * (k + 1)
    np.multiply(phi, fk, out=phi_h); np.multiply(U, fk, out=U_h)
    t0 = time.perf_counter()
    s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
    t = time.perf_counter() - t0
    print("call %d: %.3f ms  h2d %d  d2h %d  mixed %d  sweeps %d  flags %d" % (k, t * 1e3, s.info(capi.I_H2D_BYTES), s.info(capi.I_D2H_BYTES),
          s.info(capi.I_N_MIXED), s.info(capi.I_N_BOUND_SWEEPS), s.info(capi.I_ERROR_FLAGS)), flush=True)
s.close()
