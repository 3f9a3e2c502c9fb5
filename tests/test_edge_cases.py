"""Edge cases of the step (empty interface, a single cell, every cell cut, zero flux, no bounding sweeps, Courant
number above one, 2-D cases with empty patches, inflow/outflow patches): the oracle's behaviour on CPU, and the
device against the oracle, bit for bit, on the GPU."""
import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, fields, meshmod, oracle_lib


def _sphere(C, V, centre=(0.5, 0.5, 0.5), radius=0.25):
    return np.clip(0.5 - (np.linalg.norm(C - np.array(centre), axis=1) - radius) / np.cbrt(V), 0.0, 1.0)


def _case(name):
    """-> (mesh, controls, alpha0(C, V), velocity(x), n_steps, dt_factor)"""
    rot = fields.rotation_velocity
    if name == "no interface: empty":
        return meshmod.hex_block(8), {}, lambda C, V: np.zeros(len(V)), rot, 3, 0.25
    if name == "no interface: full":
        return meshmod.hex_block(8), {}, lambda C, V: np.ones(len(V)), rot, 3, 0.25
    if name == "single cell":
        return meshmod.hex_block(1), {}, lambda C, V: np.full(len(V), 0.4), rot, 2, 0.25
    if name == "every cell cut":
        return meshmod.hex_block(10), {}, lambda C, V: np.random.default_rng(3).uniform(0.05, 0.95, len(V)), rot, 3, 0.25
    if name == "zero flux":
        return meshmod.hex_block(10), {}, _sphere, (lambda x: np.zeros_like(x)), 2, 0.25
    if name == "no bounding sweeps":
        return meshmod.hex_block(12), {"nAlphaBounds": 0}, _sphere, fields.leveque_velocity, 4, 0.25
    if name == "Courant number 1.5":
        return meshmod.hex_block(12), {}, _sphere, fields.leveque_velocity, 3, 1.5
    if name == "2-D: empty front and back":
        m = meshmod.hex_block((16, 16, 1), length=(1.0, 1.0, 1.0 / 16))
        for p in m.patches:
            if p.name in ("top", "bottom"):     # the z sides of hex_block
                p.kind = capi.PATCH_EMPTY
        a0 = lambda C, V: np.clip(0.5 - (np.linalg.norm(C[:, :2] - np.array([0.5, 0.62]), axis=1) - 0.2) * 16, 0.0, 1.0)
        return m, {}, a0, rot, 4, 0.25
    if name == "inflow and outflow patches":
        m = meshmod.hex_block(10)
        for p in m.patches:
            if p.name == "back":                # x = 0: liquid enters
                p.alpha_bc, p.alpha_value = capi.BC_FIXED_VALUE, 1.0
            if p.name == "front":               # x = 1: inletOutlet
                p.alpha_bc, p.alpha_value = capi.BC_INLET_OUTLET, 0.0
        uni = lambda x: np.tile(np.array([1.0, 0.1, 0.0]), (len(x), 1))
        return m, {}, (lambda C, V: np.clip(0.5 - (C[:, 0] - 0.33) * 10, 0.0, 1.0)), uni, 5, 0.3
    raise KeyError(name)


CASES = ["no interface: empty", "no interface: full", "single cell", "every cell cut", "zero flux", "no bounding sweeps",
         "Courant number 1.5", "2-D: empty front and back", "inflow and outflow patches"]


def _run(name, lib, record):
    m, extra, alpha0, vel, steps, cfl = _case(name)
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, **extra), lib=lib)
    C, Cf, Sf, V = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF), s.field(capi.F_V)
    a0 = alpha0(C, V)
    U0, phi0 = vel(C), fields.face_flux(Cf, Sf, vel)
    Ub = vel(Cf[m.n_internal_faces:])
    umax = max(np.abs(U0).max(), 1e-30)
    dt = cfl * np.cbrt(V.min()) / umax if umax > 1e-20 else 0.01
    s.setAlpha(a0)
    s.setPhi(phi0)
    s.setU(U0, Ub)
    out = {"a0": a0, "V": V, "phi": phi0, "mesh": m, "dt": dt, "steps": []}
    for _ in range(steps):
        s.reconstruct()
        rec = {"mixed": s.mixedCells(), "status": s.cellStatus(), "N": s.interfaceN(), "D": s.interfaceD()}
        s.advect(dt)
        rec.update(alpha=s.alpha(), alphaPhi=s.alphaPhi(), alphaB=s.field(capi.F_ALPHA_BOUNDARY), sweeps=s.info(capi.I_N_BOUND_SWEEPS))
        out["steps"].append(rec)
    out["flags"] = s.info(capi.I_ERROR_FLAGS) if record else 0
    s.close()
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_edge_case(name):
    r = _run(name, oracle_lib(), False)
    a0, V, last = r["a0"], r["V"], r["steps"][-1]
    m = r["mesh"]
    closed = name not in ("inflow and outflow patches", "every cell cut")   # (the rotation crosses the walls of the box)
    if closed:
        assert abs((last["alpha"] * V).sum() - (a0 * V).sum()) <= 1e-13 * max((a0 * V).sum(), V.sum() * 1e-3)
    if name.startswith("no interface"):
        assert all(len(st["mixed"]) == 0 for st in r["steps"])
        assert np.abs(last["alpha"] - a0).max() < 1e-14 and np.abs(last["alphaPhi"] - r["phi"] * a0[0]).max() < 1e-16
    if name == "single cell":
        assert np.array_equal(last["alpha"], a0) and len(r["steps"][0]["mixed"]) == 1
    if name == "every cell cut":
        assert len(r["steps"][0]["mixed"]) == m.n_cells
    if name == "zero flux":
        assert np.abs(last["alpha"] - a0).max() < 1e-15 and not np.any(last["alphaPhi"])
    if name == "no bounding sweeps":
        assert all(st["sweeps"] == 0 for st in r["steps"])
    if name == "2-D: empty front and back":
        N = np.concatenate([st["N"][st["mixed"]] for st in r["steps"]])
        assert np.all(N[:, 2] == 0.0) and np.abs(np.linalg.norm(N, axis=1) - 1).max() < 1e-12    # reconstruction stays in-plane
    if name == "inflow and outflow patches":
        assert (last["alpha"] * V).sum() > (a0 * V).sum() + 1e-3      # liquid came in through x = 0
        nIF = m.n_internal_faces
        back = next(p for p in m.patches if p.name == "back")
        assert np.all(last["alphaB"][back.start - nIF:back.start - nIF + back.size] == 1.0)
    assert last["alpha"].min() > -1e-9 and last["alpha"].max() < 1 + 1e-9 or name in ("Courant number 1.5", "no bounding sweeps", "every cell cut")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_edge_case_matches_oracle(name, product):
    ro, rg = _run(name, oracle_lib(), False), _run(name, product, True)
    for k, (so, sg) in enumerate(zip(ro["steps"], rg["steps"])):
        assert np.array_equal(so["mixed"], sg["mixed"]), "step %d: interface-cell set" % k
        assert np.array_equal(so["status"], sg["status"])
        for key in ("N", "D", "alpha", "alphaPhi", "alphaB"):
            assert np.array_equal(so[key], sg[key]), "step %d: %s differs by %g" % (k, key, np.abs(so[key] - sg[key]).max())
        assert so["sweeps"] == sg["sweeps"]
    assert rg["flags"] == 0
