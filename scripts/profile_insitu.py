"""In-situ per-launch-site CUDA-event timing of the step (SVOF profile option): python scripts/profile_insitu.py  [OV=0|1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, numpy as np
from geometricvofext_b200 import capi, fields
from geometricvofext_b200.solver import SolveVofEqu
n=256
m, a0 = bench.build_case(n)
s = SolveVofEqu(m, bench.CONTROLS)
dt=0.2/n
U, phi = bench.velocity_fields(s, dt, dt)
s.setAlpha(a0); s.setPhi(phi); s.setU(U, np.zeros((s.nBF,3)))
s.setOption("overlap", int(os.environ.get("OV","1")))
s.setOption("dense_ctas", int(os.environ.get("DENSE_CTAS","0"))); s.setOption("dense_threads", int(os.environ.get("DENSE_THREADS","256")))
s.setOption("plic_ctas", int(os.environ.get("PLIC_CTAS","0")))
for _ in range(5): s.reconstruct(); s.advect(dt)
s.synchronize()
s.setOption("profile", 1)
for _ in range(20): s.reconstruct(); s.advect(dt)
s.synchronize()
s.close()
