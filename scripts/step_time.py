"""Device-resident step time at 256^3 for a schedule variant: OV=<overlap mode> python scripts/step_time.py [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200.solver import SolveVofEqu
n = int(os.environ.get("N", "256"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt = 0.2 / n
U, phi = bench.velocity_fields(s, dt, dt)
s.setAlpha(a0); s.setPhi(phi); s.setU(U, np.zeros((s.nBF, 3)))
for ov in [int(x) for x in os.environ.get("OV", "0").split(",")]:
    s.setOption("overlap", ov)
    graph = int(os.environ.get("GRAPH", "0"))
    def one():
        if graph: s.step(dt)
        else: s.reconstruct(); s.advect(dt)
    for _ in range(5): one()
    s.synchronize()
    s.lib.svof_mark(s._h, 0)
    for _ in range(steps): one()
    s.lib.svof_mark(s._h, 1)
    ms = C.c_double(); s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms)); s.synchronize()
    print("overlap %d: %.4f ms/step  volume %.17g" % (ov, ms.value / steps, s.volume()))
s.close()
