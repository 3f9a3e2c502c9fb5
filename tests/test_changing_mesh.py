"""the hooks for meshes that change under the solver.
  svof_update_points   mesh.moving(): geometry + face flatness follow the points (reconstruction.C:643-647)
  svof_update_mesh     mesh.topoChanging(): the handle is rebuilt for the refined mesh, parameters kept
  svof_set_interface + svof_map_alpha_field   reconstruction::mapAlphaField (reconstruction.C:725-784)
CPU tests drive the oracle; the gpu-marked ones compare the CUDA library with it, bitwise."""
import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, fields, meshmod, oracle_lib


def _fresh_run(lib, m, a0, steps=3, dt=0.01):
    s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=lib)
    _drive(s, a0, steps, dt)
    return s


def _drive(s, a0, steps, dt):
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    s.setAlpha(a0)
    s.setPhi(fields.face_flux(Cf, Sf))
    s.setU(fields.leveque_velocity(C_))
    for _ in range(steps):
        s.reconstruct()
        s.advect(dt)


def _moving_and_refining(lib):
    """one handle: static hex mesh -> points move -> topology changes; each stage must equal a fresh handle"""
    out = {}
    m0 = meshmod.hex_block(10)
    a0 = fields.sphere_alpha_quadrature(m0)
    s = SolveVofEqu(m0, LEVEQUE_CONTROLS, lib=lib)
    _drive(s, a0, 2, 0.01)
    m1 = meshmod.perturb_points(m0, 0.12, seed=5)
    s.updatePoints(m1.points)
    for f in (capi.F_CF, capi.F_SF, capi.F_C, capi.F_V, capi.F_FACE_FLATNESS):
        out["geom%d" % f] = s.field(f)
    _drive(s, a0, 3, 0.01)
    out["alpha_moved"], out["flat_min"] = s.alpha(), s.info(capi.I_FLATNESS_MIN)
    m2 = meshmod.hex_block(14)
    s.updateMesh(m2)
    a2 = fields.sphere_alpha_quadrature(m2)
    _drive(s, a2, 3, 0.008)
    out["alpha_refined"], out["mixed_refined"] = s.alpha(), s.mixedCells()
    s.close()
    return out, (m1, a0), (m2, a2)


def _check_against_fresh(lib):
    got, (m1, a1), (m2, a2) = _moving_and_refining(lib)
    f1 = _fresh_run(lib, m1, a1, 3, 0.01)
    for f in (capi.F_CF, capi.F_SF, capi.F_C, capi.F_V, capi.F_FACE_FLATNESS):
        assert np.array_equal(got["geom%d" % f], f1.field(f)), "geometry field %d after svof_update_points" % f
    assert got["flat_min"] == f1.info(capi.I_FLATNESS_MIN) and got["flat_min"] < 1.0
    assert np.array_equal(got["alpha_moved"], f1.alpha())
    f2 = _fresh_run(lib, m2, a2, 3, 0.008)
    assert np.array_equal(got["alpha_refined"], f2.alpha())
    assert np.array_equal(got["mixed_refined"], f2.mixedCells())
    return got


def _children(n):
    """fine cell -> coarse parent for hex_block(2n) over hex_block(n)"""
    k, j, i = np.meshgrid(np.arange(2 * n), np.arange(2 * n), np.arange(2 * n), indexing="ij")
    return ((i // 2) + n * ((j // 2) + n * (k // 2))).reshape(-1)


def _map_alpha_case(lib, n=12):
    """dynamicRefineFvMesh scenario: reconstruct on the coarse mesh, refine every cell 2x2x2, map alpha / interfaceN / interfaceD
    to the children the way OpenFOAM's mapFields does (parent value), then mapAlphaField."""
    mc, mf = meshmod.hex_block(n), meshmod.hex_block(2 * n)
    s = SolveVofEqu(mc, LEVEQUE_CONTROLS, lib=lib)
    ac = fields.sphere_alpha_quadrature(mc)
    s.setAlpha(ac)
    s.reconstruct()
    N, D, Vc = s.interfaceN(), s.interfaceD(), s.field(capi.F_V)
    parent = _children(n)
    s.updateMesh(mf)
    s.setAlpha(ac[parent])
    s.setInterface(N[parent], D[parent])
    s.mapAlphaField(0.01, 0.99)
    af, Vf = s.alpha(), s.field(capi.F_V)
    t = s.alphaMappingTime()
    # a second step on the refined mesh must run (the dense interface fields are cleared by the next reconstruct)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    s.setPhi(fields.face_flux(Cf, Sf))
    s.setU(fields.leveque_velocity(C_))
    s.reconstruct()
    s.advect(0.005)
    a_next = s.alpha()
    s.close()
    return ac, af, Vc, Vf, parent, t, a_next


def test_oracle_moving_points_and_topology_change_equal_fresh_handles():
    _check_against_fresh(oracle_lib())


def test_oracle_map_alpha_field_redistributes_the_parent_volume():
    ac, af, Vc, Vf, parent, t, _ = _map_alpha_case(oracle_lib())
    sel = (ac >= 0.01) & (ac <= 0.99)
    vol_children = np.bincount(parent, weights=af * Vf, minlength=ac.size)
    # inside the refinement band the children share out exactly what the parent's plane cuts off: the parent's volume
    assert sel.sum() > 30
    assert np.abs(vol_children[sel] - ac[sel] * Vc[sel]).max() <= 1e-12 * Vc.max() * 50, \
        "children of a cut cell must hold the parent's liquid volume (plane positioned to 1e-14 in alpha)"
    # outside the band the mapped (parent) value is kept
    keep = ~sel[parent]
    assert np.array_equal(af[keep], ac[parent][keep])
    # and inside it the field is sharper than the piecewise-constant map: some children are now exactly full or empty
    band = sel[parent]
    assert ((af[band] == 0.0) | (af[band] == 1.0)).sum() > 0.3 * band.sum()
    assert t > 0.0


@pytest.mark.gpu
def test_gpu_moving_points_and_topology_change(product):
    got_g = _check_against_fresh(product)
    got_o, _, _ = _moving_and_refining(oracle_lib())
    for k in got_g:
        assert np.array_equal(np.asarray(got_g[k]), np.asarray(got_o[k])), k


@pytest.mark.gpu
def test_gpu_map_alpha_field_parity(product):
    o = _map_alpha_case(oracle_lib())
    g = _map_alpha_case(product)
    assert np.array_equal(o[1], g[1]), "mapAlphaField result differs by %g" % np.abs(o[1] - g[1]).max()
    assert np.array_equal(o[6], g[6]), "step after mapAlphaField differs"
