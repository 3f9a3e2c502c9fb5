"""svof_step_host with pinned buffers at N^3 (default 256): the overlapped zero-copy form against the serial one
(option zc_overlap), same inputs, wall time per call and bitwise comparison of the outputs.
python scripts/e2e_ab.py [N] [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
lib = s.lib
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
phi_h, U_h, Ub_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3)), capi.pinned_array(lib, (max(s.nBF, 1), 3))
a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
Ub_h[:] = 0
ref = None
for mode in [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '0,7,0,7').split(',')]:
    s.setOption("zc_overlap", mode % 100)
    if mode >= 100:
        s.setOption("zc_push_ctas", mode // 100)
    s.setAlpha(a0)
    ts = []
    for k in range(6 + calls):
        fk = 1.0 - 1e-3 * (k + 1)
        np.multiply(phi, fk, out=phi_h); np.multiply(U, fk, out=U_h)
        t0 = time.perf_counter()
        s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
        ts.append(time.perf_counter() - t0)
    res = (np.array(a_out), np.array(ap_out))
    if ref is None:
        ref = res
    same = np.array_equal(res[0], ref[0]) and np.array_equal(res[1], ref[1])
    print("zc_overlap %d: %.3f ms/call (median %.3f, last %d calls)  h2d %d  d2h %d  bitwise-same-as-first %s  flags %d" % (
        mode, 1e3 * np.mean(ts[6:]), 1e3 * np.median(ts[6:]), calls, s.info(capi.I_H2D_BYTES), s.info(capi.I_D2H_BYTES), same,
        s.info(capi.I_ERROR_FLAGS)), flush=True)
s.close()
