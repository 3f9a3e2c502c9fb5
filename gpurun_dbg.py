import sys; sys.path.insert(0,'tests')
from common import *
N=64
m = meshmod.hex_block(N)
s = SolveVofEqu(m, LEVEQUE_CONTROLS)
a0 = fields.sphere_alpha_quadrature(m)
C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
U0, phi0 = fields.leveque_velocity(C_), fields.face_flux(Cf, Sf)
s.setAlpha(a0)
out = np.zeros(m.n_cells); aphi = np.zeros(m.n_faces)
prev_a = None
for k in range(4):
    s.step_host(0.2/N, phi0, U0, None, out, aphi)
    a = out.copy(); p = aphi.copy()
    if prev_a is not None:
        da = (a.view(np.int64) != prev_a.view(np.int64)); dp = (p.view(np.int64) != prev_p.view(np.int64))
        print(k, "alpha changed", da.sum(), "alphaPhi changed", dp.sum(), "of", m.n_faces, "h2d", s.info(capi.I_H2D_BYTES), "d2h", s.info(capi.I_D2H_BYTES))
        idx = np.nonzero(dp)[0]
        own = m.owner[idx]
        print("   changed faces: owner alpha==0:", (a[own]==0).sum(), " ==1:", (a[own]==1).sum(), "sample", p[idx[:5]], prev_p[idx[:5]])
    prev_a, prev_p = a, p
