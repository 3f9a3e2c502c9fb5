"""Host-phase breakdown of svof_step_host at 256^3 (profile option): python scripts/profile_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = 256
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
lib = s.lib
phi_h, U_h, Ub_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3)), capi.pinned_array(lib, (max(s.nBF, 1), 3))
a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
phi_h[:] = phi; U_h[:] = U; Ub_h[:] = 0
s.setAlpha(a0)
for _ in range(25): s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
s.setOption("profile", 1)
t0 = time.perf_counter()
K = 10
for _ in range(K): s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
print("e2e %.3f ms/step (profiled, %d steps)" % (1e3 * (time.perf_counter() - t0) / K, K))
s.close()
