"""Schedule sweep at N^3 (default 256): device-resident graph step time for combinations of the run-time options
(overlap, fork, plic_ctas, dense_ctas[, dense_threads]).  python scripts/sweep_step.py "ov,fork,plic,dense[,thr[,l2]];..." [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = int(os.environ.get("N", "256"))
combos = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1].split(";")]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
s.setPhi(phi); s.setU(U, np.zeros((s.nBF, 3)))
ref = None
for combo in combos:
    ov, fork, plic, dense = combo[:4]
    thr = combo[4] if len(combo) > 4 else 256
    l2 = combo[5] if len(combo) > 5 else 0
    s.setOption("dense_l2", l2)
    split = combo[6] if len(combo) > 6 else 0
    s.setOption("dense_split", split)
    s.setAlpha(a0)
    s.setOption("overlap", ov); s.setOption("fork", fork); s.setOption("plic_ctas", plic); s.setOption("dense_ctas", dense)
    s.setOption("dense_threads", thr)
    for _ in range(6): s.step(dt)
    s.synchronize()
    s.lib.svof_mark(s._h, 0)
    for _ in range(steps): s.step(dt)
    s.lib.svof_mark(s._h, 1)
    ms = C.c_double(); s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms)); s.synchronize()
    a = s.alpha()
    if ref is None: ref = a
    print("overlap %d fork %d plic_ctas %d dense_ctas %d dense_threads %d dense_l2 %d dense_split %d: %.4f ms/step  bitwise-same-as-first %s  err %d" %
          (ov, fork, plic, dense, thr, l2, split, ms.value / steps, np.array_equal(a, ref), int(s.info(capi.I_ERROR_FLAGS))), flush=True)
s.close()
