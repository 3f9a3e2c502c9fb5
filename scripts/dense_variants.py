"""Streaming-kernel variants at N^3 (default 256): kernel time alone (plain launches, single stream), graph-replayed step
with the two-stream schedule, and bitwise equality of alpha/alphaPhi against the default kernel.
    [SVOF_LIB=...] [N=256] python scripts/dense_variants.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi
from geometricvofext_b200.solver import SolveVofEqu
n = int(os.environ.get("N", "256"))
steps = 20
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
s.setPhi(phi); s.setU(U, np.zeros((s.nBF, 3)))
ref = None
for v3 in (0, 1, 2):
    s.setOption("dense_v3", 1 if v3 == 1 else 0)
    s.setOption("dense_v4", 1 if v3 == 2 else 0)
    s.setAlpha(a0)
    s.setOption("overlap", 0)
    for _ in range(3): s.reconstruct(); s.advect(dt)
    s.synchronize()
    d0, n0 = s.info(capi.I_DENSE_KERNEL_MS), s.info(capi.I_DENSE_KERNEL_LAUNCHES)
    for _ in range(steps): s.reconstruct(); s.advect(dt)
    s.synchronize()
    k_ms = (s.info(capi.I_DENSE_KERNEL_MS) - d0) / (s.info(capi.I_DENSE_KERNEL_LAUNCHES) - n0)
    out = (s.alpha(), s.alphaPhi())
    same = True if ref is None else (np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1]))
    if ref is None: ref = out
    s.setOption("overlap", 1)
    s.setAlpha(a0)
    for _ in range(6): s.step(dt)
    s.synchronize()
    s.lib.svof_mark(s._h, 0)
    for _ in range(steps): s.step(dt)
    s.lib.svof_mark(s._h, 1)
    ms = C.c_double(); s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms)); s.synchronize()
    print("variant %d (0 rows, 1 batched, 2 sliced rows): kernel alone %.1f us   graph step (two streams) %.4f ms   bitwise-same-as-default %s   err %d" %
          (v3, 1e3 * k_ms, ms.value / steps, same, int(s.info(capi.I_ERROR_FLAGS))), flush=True)
s.close()
