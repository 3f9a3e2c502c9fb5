"""Run-time schedule selection of svof_step_device at N^3 (default 256): step time with the selection off (default schedule),
with the alternative forced, and with the selection on (what it picks, early and in the t = 1.5 window).
python scripts/sched_auto_check.py [N] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
s.setPhi(phi); s.setU(U, np.zeros((s.nBF, 3)))


def timed(label, warm):
    s.setAlpha(a0)
    for _ in range(warm): s.step(dt)
    s.synchronize()
    s.lib.svof_mark(s._h, 0)
    for _ in range(steps): s.step(dt)
    s.lib.svof_mark(s._h, 1)
    ms = C.c_double(); s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms)); s.synchronize()
    print("%-34s %.4f ms/step  schedule %d  flags %d" % (label, ms.value / steps, int(s.info(capi.I_SCHEDULE)), int(s.info(capi.I_ERROR_FLAGS))), flush=True)
    return s.alpha()


s.setOption("sched_auto", 0)
ref = timed("selection off (default schedule)", 6)
s.setOption("fork", 2); s.setOption("dense_ctas", 4)
a = timed("fork 2, dense_ctas 4 forced", 6)
print("  bitwise same:", np.array_equal(a, ref))
s.setOption("fork", 1); s.setOption("dense_ctas", 0)
s.setOption("sched_auto", 1)
timed("selection on (settling + timed)", 24)      # (24 + steps steps: a different end state, not compared)
a = timed("selection on (settled)", 6)
print("  bitwise same:", np.array_equal(a, ref))
s.close()
