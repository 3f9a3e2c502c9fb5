"""Caller-side pieces of the reference's advection-only driver
(applications/test/plicVofAdvectionFoam): the prescribed velocity field, the
time-step control and the time loop.  They sit either side of the hot path and
are plain numpy; nothing here is timed as part of the SimPLIC step.
"""
import math

import numpy as np

from . import capi


def leveque_velocity(x):
    """3-D deformation field of updateU.H:13-15 (cells) / :27-29 (faces), unscaled."""
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    pi = math.pi
    u = np.empty_like(x)
    u[:, 0] = 2 * np.sin(pi * X) ** 2 * np.sin(2 * pi * Y) * np.sin(2 * pi * Z)
    u[:, 1] = -np.sin(2 * pi * X) * np.sin(pi * Y) ** 2 * np.sin(2 * pi * Z)
    u[:, 2] = -np.sin(2 * pi * X) * np.sin(2 * pi * Y) * np.sin(pi * Z) ** 2
    return u


def rotation_velocity(x, omega=2 * math.pi, centre=(0.5, 0.5)):
    """Solid-body rotation about the z axis through `centre`."""
    u = np.zeros_like(x)
    u[:, 0] = -omega * (x[:, 1] - centre[1])
    u[:, 1] = omega * (x[:, 0] - centre[0])
    return u


def u_factor(t, dt, period):
    """updateU.H:59-69."""
    if period > 0.0:
        return 0.5 * (math.cos(2.0 * math.pi * t / period) + math.cos(2.0 * math.pi * (t + dt) / period))
    return 1.0


def face_flux(Cf, Sf, velocity=leveque_velocity):
    """phi = U(Cf) & Sf (updateU.H:31-57)."""
    Uf = velocity(Cf)
    return Uf[:, 0] * Sf[:, 0] + Uf[:, 1] * Sf[:, 1] + Uf[:, 2] * Sf[:, 2]


def sphere_alpha_quadrature(mesh, centre=(0.35, 0.35, 0.35), radius=0.15, order=24):
    """Volume fractions of a sphere on a hex_block mesh by Gauss quadrature in (x,y) and
    exact integration in z.  ~1e-6 accurate: good enough for the benchmark harness; the
    parity tests use the reference's exact overlap library (oracle/_ref) instead."""
    from .mesh import cell_centres_hex
    N, lo, hi = (np.array(mesh.meta[k]) for k in ("N", "lo", "hi"))
    h = np.array(mesh.meta["length"]) / N
    Cc = cell_centres_hex(mesh)
    d = Cc - np.array(centre)
    half_diag = 0.5 * np.linalg.norm(h)
    dist = np.linalg.norm(d, axis=1)
    alpha = np.zeros(Cc.shape[0])
    alpha[dist + half_diag <= radius] = 1.0
    cut = np.nonzero(np.abs(dist - radius) < half_diag)[0]
    if cut.size:
        g, w = np.polynomial.legendre.leggauss(order)
        gx = Cc[cut, 0][:, None] + 0.5 * h[0] * g[None, :]
        gy = Cc[cut, 1][:, None] + 0.5 * h[1] * g[None, :]
        rho2 = radius ** 2 - (gx[:, :, None] - centre[0]) ** 2 - (gy[:, None, :] - centre[1]) ** 2
        rho = np.sqrt(np.maximum(rho2, 0.0))
        z0 = (Cc[cut, 2] - 0.5 * h[2])[:, None, None]
        z1 = (Cc[cut, 2] + 0.5 * h[2])[:, None, None]
        length = np.maximum(0.0, np.minimum(z1, centre[2] + rho) - np.maximum(z0, centre[2] - rho))
        length = np.where(rho2 > 0, length, 0.0)
        frac = np.einsum("cij,i,j->c", length, w, w) * 0.25 / h[2]
        alpha[cut] = np.clip(frac, 0.0, 1.0)
    return alpha


class AdvectionDriver:
    """Time loop of plicVofAdvectionFoam (plicVof.H:13-57) around a SolveVofEqu.

    Courant numbers of step k use phi of step k-1 (CourantNo.H / alphaCourantNo.H:40-52
    run before ++runTime and updateU.H); dt follows setDeltaT.H:34-53 unless fixed_dt.
    """

    def __init__(self, solver, velocity=leveque_velocity, period=6.0, max_co=0.5, max_alpha_co=0.5,
                 max_delta_t=0.2, delta_t0=0.001, fixed_dt=None, write_interval=None, start_time=0.0, reduce_max=None):
        self.s = solver
        # decomposed runs: the Courant numbers are gMax over all ranks (CourantNo.H / alphaCourantNo.H reduce)
        self.reduce_max = reduce_max or (lambda x: x)
        # writeControl adjustableRunTime: Time::adjustDeltaT spreads the time to the next write over equal steps
        self.write_interval, self.start_time, self.write_index, self.write_now = write_interval, start_time, 0, False
        self.period, self.max_co, self.max_alpha_co = period, max_co, max_alpha_co
        self.max_delta_t, self.fixed_dt = max_delta_t, fixed_dt
        self.t, self.dt = 0.0, (fixed_dt if fixed_dt else delta_t0)
        self.C = solver.field(capi.F_C)
        self.V = solver.field(capi.F_V)
        Cf, Sf = solver.field(capi.F_CF), solver.field(capi.F_SF)
        self.U0 = velocity(self.C)
        self.phi0 = face_flux(Cf, Sf, velocity)
        # wall patches: fixedValue (0 0 0)  (0.orig/U); other patch kinds would be sampled by the caller
        self.Ub = np.zeros((solver.nBF, 3))
        m = solver.mesh
        self.own, self.nei = m.owner, m.neighbour
        self.phi = np.zeros(solver.nF)      # createPhi.H from U = 0
        self.steps = 0

    def _sum_mag_phi(self):
        s = np.zeros(self.s.nC)
        a = np.abs(self.phi)
        nIF = self.s.nIF
        np.add.at(s, self.own[:nIF], a[:nIF])
        np.add.at(s, self.nei, a[:nIF])
        np.add.at(s, self.own[nIF:], a[nIF:])
        return s

    def set_delta_t(self, alpha):
        if self.fixed_dt:
            self.dt = self.fixed_dt
            return
        SMALL = 1e-15
        sp = self._sum_mag_phi()
        co = 0.5 * self.reduce_max(np.max(sp / self.V)) * self.dt
        mask = (alpha - 0.01 >= 0) & (0.99 - alpha >= 0)
        aco = 0.5 * self.reduce_max(np.max(mask * sp / self.V)) * self.dt
        f = min(self.max_co / (co + SMALL), self.max_alpha_co / (aco + SMALL))
        fact = min(min(f, 1.0 + 0.1 * f), 1.2)
        self.dt = min(fact * self.dt, self.max_delta_t)
        self.adjust_delta_t()

    def adjust_delta_t(self):
        """Time::adjustDeltaT (OpenFOAM Time.C, called by Time::setDeltaT(dt, adjust = true) from setDeltaT.H:50):
        the remaining time to the next write is covered by round(timeToNextWrite/deltaT) equal steps; the step may grow
        by at most a factor 2 and shrink by at most a factor 5."""
        if not self.write_interval:
            return
        to_next = max(0.0, (self.write_index + 1) * self.write_interval - (self.t - self.start_time))
        n_steps = to_next / self.dt
        if n_steps < 2 ** 31 - 1:
            n = max(1, int(np.floor(n_steps + 0.5)))      # label(round(nSteps)), at least 1
            new_dt = to_next / n
            self.dt = min(new_dt, 2.0 * self.dt) if new_dt >= self.dt else max(new_dt, 0.2 * self.dt)

    def running(self, end_time):
        """Time::run(): value() < endTime - 0.5*deltaT."""
        return self.t < end_time - 0.5 * self.dt

    def step(self, end_time=None):
        alpha = self.s.alpha() if not self.fixed_dt else None
        self.set_delta_t(alpha)
        if end_time is not None and not self.write_interval and self.t + self.dt > end_time - 1e-12:
            self.dt = end_time - self.t
        self.t += self.dt                       # ++runTime
        self.write_now = False
        if self.write_interval:                 # Time::operator++: writeTime when the write index advances
            wi = int(((self.t - self.start_time) + 0.5 * self.dt) / self.write_interval)
            if wi > self.write_index:
                self.write_now, self.write_index = True, wi
        f = u_factor(self.t, self.dt, self.period)   # updateU.H uses the NEW time value
        self.phi = self.phi0 * f
        self.s.setPhi(self.phi)
        self.s.setU(self.U0 * f, self.Ub)
        self.s.reconstruct()
        self.s.advect(self.dt)
        self.steps += 1
        return self.t
