"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: interface-cell set and cut-face topology
bit-exact, alpha within 1e-12 absolute per step, volume conserved to 1e-13 relative.
In practice the two agree BITWISE (same operation order, no FMA contraction); the asserts
state the contractual tolerance and the bitwise result is reported."""
import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, exact_sphere_alpha, fields, meshmod

pytestmark = pytest.mark.gpu

ATOL_ALPHA = 1e-12


def _pair(m, controls, oracle, product):
    return SolveVofEqu(m, controls, lib=oracle), SolveVofEqu(m, controls, lib=product)


def _setup_leveque(m, so, sg, t=0.0, dt=None, alpha0=None):
    a0 = exact_sphere_alpha(m) if alpha0 is None else alpha0
    C, Cf, Sf = so.field(capi.F_C), so.field(capi.F_CF), so.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C), fields.face_flux(Cf, Sf)
    for s in (so, sg):
        s.setAlpha(a0)
    return a0, U0, phi0


def test_mesh_geometry_matches(oracle, product):
    m = meshmod.hex_block(12)
    so, sg = _pair(m, LEVEQUE_CONTROLS, oracle, product)
    for f in (capi.F_CF, capi.F_SF, capi.F_C, capi.F_V, capi.F_FACE_FLATNESS):
        a, b = so.field(f), sg.field(f)
        assert np.array_equal(a, b), "mesh field %d differs: max %g" % (f, np.abs(a - b).max())
    for i in (capi.I_FLATNESS_MIN, capi.I_FLATNESS_MAX):
        assert so.info(i) == sg.info(i)


@pytest.mark.parametrize("N", [16, 32])
def test_reconstruct_parity(oracle, product, N):
    m = meshmod.hex_block(N)
    so, sg = _pair(m, LEVEQUE_CONTROLS, oracle, product)
    _setup_leveque(m, so, sg)
    so.reconstruct()
    sg.reconstruct()
    mo, mg = so.mixedCells(), sg.mixedCells()
    assert np.array_equal(mo, mg), "interface-cell list must be bit-exact (and in the same order)"
    assert np.array_equal(so.cellStatus(), sg.cellStatus())
    for name, f in (("N", capi.F_INTERFACE_N), ("D", capi.F_INTERFACE_D), ("C", capi.F_INTERFACE_C), ("S", capi.F_INTERFACE_S)):
        a, b = so.field(f), sg.field(f)
        assert np.abs(a - b).max() <= 1e-13, "interface%s differs by %g" % (name, np.abs(a - b).max())
        assert np.array_equal(a, b), "interface%s not bitwise equal (max %g)" % (name, np.abs(a - b).max())


@pytest.mark.parametrize("N,steps", [(16, 12), (32, 8)])
def test_step_parity_leveque(oracle, product, N, steps):
    m = meshmod.hex_block(N)
    so, sg = _pair(m, LEVEQUE_CONTROLS, oracle, product)
    a0, U0, phi0 = _setup_leveque(m, so, sg)
    dt = 0.5 / N / 2.0
    t = 0.0
    v0 = sg.volume()
    for k in range(steps):
        t += dt
        f = fields.u_factor(t, dt, 6.0)
        for s in (so, sg):
            s.setPhi(phi0 * f)
            s.setU(U0 * f)
            s.reconstruct()
            s.advect(dt)
        ao, ag = so.alpha(), sg.alpha()
        assert np.array_equal(so.mixedCells(), sg.mixedCells()), "step %d: interface-cell set differs" % k
        assert np.array_equal(so.cellStatus(), sg.cellStatus()), "step %d: cut status differs" % k
        d = np.abs(ao - ag).max()
        assert d <= ATOL_ALPHA, "step %d: alpha differs by %g" % (k, d)
        assert np.array_equal(ao, ag), "step %d: alpha not bitwise equal (max %g)" % (k, d)
        assert np.array_equal(so.field(capi.F_UN0), sg.field(capi.F_UN0))
        assert np.array_equal(so.alphaPhi(), sg.alphaPhi()), "step %d: alphaPhi differs" % k
        assert np.array_equal(so.field(capi.F_DVF), sg.field(capi.F_DVF)), "step %d: dVf differs" % k
        assert so.info(capi.I_N_BOUND_SWEEPS) == sg.info(capi.I_N_BOUND_SWEEPS)
        for i in (capi.I_MIN_ALPHA_BEFORE, capi.I_MAX_ALPHA_M1_BEFORE, capi.I_MIN_ALPHA_AFTER, capi.I_MAX_ALPHA_M1_AFTER):
            assert so.info(i) == sg.info(i), "info %d differs: %r vs %r" % (i, so.info(i), sg.info(i))
        assert abs(sg.volume() - v0) <= 1e-13 * abs(v0), "volume not conserved"
        assert abs(sg.volume() - so.volume()) <= 1e-13 * abs(v0)
    assert sg.info(capi.I_ERROR_FLAGS) == 0


def test_step_parity_64_from_the_golden_t15_field(oracle, product):
    """The reference's own 64^3 case at its hardest instant: the golden exact field at t = 1.5 (maximum deformation,
    6 094 interface cells in thin sheets -- where boundFlux chains, sliver sub-cells and < 3-point sub-faces occur),
    advanced 40 steps with the driver's adaptive dt (maxCo = maxAlphaCo = 0.5) and phi(t), U(t) of updateU.H.
    Every step: interface-cell list, cut status, N/D/C/S, Un0, alpha, alphaPhi, dVf and the bounding log bitwise."""
    import os
    N = 64
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "exact_alpha_64.npz"))
    a0 = np.zeros(N ** 3)
    a0[G["full_idx_1.5"]] = 1.0
    a0[G["part_idx_1.5"]] = G["part_val_1.5"]
    assert int(((a0 > 1e-8) & (a0 < 1 - 1e-8)).sum()) == 6094
    m = meshmod.hex_block(N)
    so, sg = _pair(m, LEVEQUE_CONTROLS, oracle, product)
    drivers = []
    for s in (so, sg):
        s.setAlpha(a0)
        d = fields.AdvectionDriver(s)
        d.t, d.dt = 1.5, 2.0e-3
        d.phi = d.phi0 * fields.u_factor(1.5, d.dt, 6.0)
        drivers.append(d)
    n_sweeps = 0
    for k in range(40):
        for d in drivers:
            d.step()
        assert drivers[0].dt == drivers[1].dt, "step %d: the two drivers chose different time steps" % k
        assert np.array_equal(so.mixedCells(), sg.mixedCells()), "step %d: interface-cell set differs" % k
        assert np.array_equal(so.cellStatus(), sg.cellStatus()), "step %d: cut status differs" % k
        for name, f in (("N", capi.F_INTERFACE_N), ("D", capi.F_INTERFACE_D), ("C", capi.F_INTERFACE_C), ("S", capi.F_INTERFACE_S),
                        ("Un0", capi.F_UN0), ("alpha", capi.F_ALPHA), ("alphaPhi", capi.F_ALPHA_PHI), ("dVf", capi.F_DVF)):
            assert np.array_equal(so.field(f), sg.field(f)), "step %d: %s differs" % (k, name)
        assert so.info(capi.I_N_BOUND_SWEEPS) == sg.info(capi.I_N_BOUND_SWEEPS)
        n_sweeps += int(sg.info(capi.I_N_BOUND_SWEEPS))
        for i in (capi.I_MIN_ALPHA_BEFORE, capi.I_MAX_ALPHA_M1_BEFORE, capi.I_MIN_ALPHA_AFTER, capi.I_MAX_ALPHA_M1_AFTER):
            assert so.info(i) == sg.info(i)
    assert len(sg.mixedCells()) > 5000 and n_sweeps > 0
    assert sg.info(capi.I_ERROR_FLAGS) == 0


def test_step_parity_clip_snap_sources(oracle, product):
    """damBreak-style controls (clip true, snapTol 1e-12, mixedCellTol 1e-10, nAlphaBounds 5:
    tutorials/solvers/interPlicFoam/damBreakWithObstacle/system/fvSolution) and non-zero Sp/Su
    (interPlicPhaseChangeFoam/alphaSuSp.H:14-15)."""
    N = 16
    m = meshmod.hex_block(N)
    ctl = dict(LEVEQUE_CONTROLS, clip=True, snapTol=1e-12, mixedCellTol=1e-10, nAlphaBounds=5)
    so, sg = _pair(m, ctl, oracle, product)
    a0, U0, phi0 = _setup_leveque(m, so, sg)
    rng = np.random.default_rng(7)
    Sp = -0.3 * rng.random(m.n_cells)
    Su = 0.05 * rng.random(m.n_cells) * (a0 > 0)
    dt = 0.01
    for k in range(6):
        for s in (so, sg):
            s.setPhi(phi0)
            s.setU(U0)
            s.reconstruct()
            s.advect(dt, Sp=Sp, Su=Su)
        ao, ag = so.alpha(), sg.alpha()
        assert np.array_equal(so.mixedCells(), sg.mixedCells())
        assert np.abs(ao - ag).max() <= ATOL_ALPHA
        assert np.array_equal(ao, ag)
        assert ag.min() >= 0.0 and ag.max() <= 1.0
        assert np.array_equal(so.field(capi.F_ALPHA_BOUNDARY), sg.field(capi.F_ALPHA_BOUNDARY))


def test_primitives_parity(oracle, product):
    N = 8
    m = meshmod.hex_block(N)
    so, sg = _pair(m, LEVEQUE_CONTROLS, oracle, product)
    rng = np.random.default_rng(3)
    n = 4000
    h = 1.0 / N
    # polygons: random quads/pentagons near the unit square, random planes
    for nv in (3, 4, 5, 7):
        ang = np.sort(rng.uniform(0, 2 * np.pi, size=(n, nv)), axis=1)
        r = rng.uniform(0.5, 1.0, size=(n, nv))
        pts = np.stack([r * np.cos(ang), r * np.sin(ang), 0.05 * rng.normal(size=(n, nv))], axis=2)
        nrm = rng.normal(size=(n, 3))
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        D = rng.uniform(-0.8, 0.8, n)
        D[:50] = -(nrm[:50] * pts[:50, 0]).sum(1)  # plane exactly through a vertex
        ro, rg = so.cutFaces(pts, nrm, D), sg.cutFaces(pts, nrm, D)
        for a, b in zip(ro, rg):
            assert np.array_equal(a, b)
    cells = rng.integers(0, m.n_cells, n).astype(np.int32)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    nrm[:100] = np.eye(3)[rng.integers(0, 3, 100)]  # axis-aligned: planes through faces/vertices
    C = so.field(capi.F_C)[cells]
    D = -(nrm * (C + rng.uniform(-0.7 * h, 0.7 * h, size=(n, 3)))).sum(1)
    for a, b in zip(so.cutCells(cells, nrm, D), sg.cutCells(cells, nrm, D)):
        assert np.array_equal(a, b)
    al = rng.uniform(1e-8, 1 - 1e-8, n)
    al[:20] = 10.0 ** rng.uniform(-8, -3, 20)
    for a, b in zip(so.findSignedDistance(cells, al, nrm), sg.findSignedDistance(cells, al, nrm)):
        assert np.array_equal(a, b)
    faces = rng.integers(0, m.n_faces, n).astype(np.int32)
    Cf = so.field(capi.F_CF)[faces]
    D = -(nrm * (Cf + rng.uniform(-0.6 * h, 0.6 * h, size=(n, 3)))).sum(1)
    Un0 = rng.uniform(-2, 2, n)
    Un0[:30] = 0.0
    phi = rng.uniform(-1, 1, n) * h * h
    assert np.array_equal(so.faceFluxes(faces, nrm, D, Un0, 0.3 * h, phi), sg.faceFluxes(faces, nrm, D, Un0, 0.3 * h, phi))


def test_step_host_matches_split_calls(oracle, product):
    N = 16
    m = meshmod.hex_block(N)
    s1 = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
    s2 = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
    a0 = exact_sphere_alpha(m)
    C, Cf, Sf = s1.field(capi.F_C), s1.field(capi.F_CF), s1.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C), fields.face_flux(Cf, Sf)
    s1.setAlpha(a0)
    s2.setAlpha(a0)
    out = np.empty(m.n_cells)
    aphi = np.empty(m.n_faces)
    for k in range(3):
        s1.setPhi(phi0)
        s1.setU(U0)
        s1.reconstruct()
        s1.advect(0.01)
        s2.step_host(0.01, phi0, U0, None, out, aphi)
        assert np.array_equal(s1.alpha(), out)
        assert np.array_equal(s1.alphaPhi(), aphi)
    assert s1.info(capi.I_GPU_LAUNCHES) > 0


@pytest.mark.parametrize("pinned", [False, True, "serial"], ids=["pageable caller buffers (staged)", "pinned caller buffers (zero copy)",
                                                              "pinned caller buffers (zero copy, zc_overlap 0)"])
@pytest.mark.parametrize("case", ["sphere in an empty box", "half-filled box with inflow/outflow patches",
                                  "sphere in an empty box, Courant number 2", "sphere with negative slivers next to empty cells"])
def test_step_host_sparse_phi_upload_is_bitwise_the_full_upload(product, case, pinned):
    """svof_step_host uploads only the phi entries the step can depend on (faces next to a cell with alpha != 0, and all
    boundary faces) and reads alphaPhi back packed by the same face bitmap; phi changes on EVERY face between calls,
    including its sign.  Results must be bitwise those of the split calls with full fields, and of svof_step_host with
    sparse_phi off.  At Courant 2 empty cells overfill and bounding corrections land on faces outside the bitmap: the
    library must notice and return the full alphaPhi.  Negative slivers (what nAlphaBounds 3 without snapping leaves
    behind in the LeVeque run: 256^3 reaches this state after three steps) flow into EMPTY cells next to the interface,
    which then go out of bounds and are swept by boundFlux although phi was never uploaded on their other faces: with
    pinned buffers the sweep reads those entries from the caller's phi (no second pass), with pageable ones the step is
    redone with the full flux field."""
    N = 16
    dt = 0.06 if case.endswith("2") else 0.01
    m = meshmod.hex_block(N)
    if case.startswith("half"):
        for p in m.patches:
            if p.name == "back":
                p.alpha_bc, p.alpha_value = capi.BC_FIXED_VALUE, 1.0
            if p.name == "front":
                p.alpha_bc, p.alpha_value = capi.BC_INLET_OUTLET, 1.0
    ctl = dict(LEVEQUE_CONTROLS, clip=True, snapTol=1e-12) if case.startswith("half") else LEVEQUE_CONTROLS
    s1, s2, s3 = (SolveVofEqu(m, ctl, lib=product) for _ in range(3))
    C, Cf, Sf = s1.field(capi.F_C), s1.field(capi.F_CF), s1.field(capi.F_SF)
    if case.startswith("half"):
        a0 = (C[:, 2] < 0.5).astype(np.float64)
        vel = lambda x: np.tile(np.array([1.0, 0.2, 0.1]), (len(x), 1))
        U0, phi0 = vel(C), fields.face_flux(Cf, Sf, vel)
        Ub = vel(Cf[m.n_internal_faces:])
    else:
        a0 = exact_sphere_alpha(m)
        U0, phi0 = fields.leveque_velocity(C), fields.face_flux(Cf, Sf)
        Ub = np.zeros((s1.nBF, 3))
    slivers = case.startswith("sphere with negative slivers")
    if slivers:   # every empty face-neighbour of an interface cell carries a small negative value
        mixed = (a0 > 1e-8) & (a0 < 1 - 1e-8)
        nIF = m.n_internal_faces
        own, nei = m.owner[:nIF], m.neighbour
        touch = np.zeros(m.n_cells, bool)
        touch[own[mixed[nei]]] = True
        touch[nei[mixed[own]]] = True
        sel = touch & (a0 == 0.0)
        assert sel.sum() > 50
        a0 = a0.copy()
        a0[sel] = -1e-4 * np.random.default_rng(2).random(int(sel.sum()))
    for s in (s1, s2, s3):
        s.setAlpha(a0)
    s3.setOption("sparse_phi", 0)
    if pinned == "serial":   # the one-stream order of the pulls and pushes (the default overlaps them with the interface kernels)
        s2.setOption("zc_overlap", 0)
    redone = False
    out2, aphi2, out3, aphi3 = np.empty(m.n_cells), np.empty(m.n_faces), np.empty(m.n_cells), np.empty(m.n_faces)
    if pinned:   # page-locked buffers: the kernels read phi / U and write alpha / alphaPhi in the caller's memory directly
        out2, aphi2 = capi.pinned_array(product, (m.n_cells,)), capi.pinned_array(product, (m.n_faces,))
        phi_p, U_p, Ub_p = (capi.pinned_array(product, (m.n_faces,)), capi.pinned_array(product, (m.n_cells, 3)),
                            capi.pinned_array(product, (max(s1.nBF, 1), 3)))
    rng = np.random.default_rng(5)
    full = 8 * (m.n_faces + 3 * m.n_cells + 3 * s1.nBF)
    for k, f in enumerate((1.0, 0.7, -0.4, -1.0, 0.3, 0.9)):
        phi = phi0 * f + 1e-9 * rng.normal(size=phi0.shape) * np.abs(phi0).max()   # every face changes every step
        U = U0 * f
        s1.setPhi(phi)
        s1.setU(U, Ub * f)
        s1.reconstruct()
        s1.advect(dt)
        if pinned:
            phi_p[:], U_p[:], Ub_p[:s1.nBF] = phi, U, Ub * f
            s2.step_host(dt, phi_p, U_p, Ub_p, out2, aphi2)
        else:
            s2.step_host(dt, phi, U, Ub * f, out2, aphi2)
        s3.step_host(dt, phi, U, Ub * f, out3, aphi3)
        assert np.array_equal(s1.alpha(), out2), "step %d: alpha differs (sparse phi)" % k
        assert np.array_equal(s1.alphaPhi(), aphi2), "step %d: alphaPhi differs (sparse phi)" % k
        assert np.array_equal(out3, out2) and np.array_equal(aphi3, aphi2)
        assert np.array_equal(s1.field(capi.F_ALPHA_BOUNDARY), s2.field(capi.F_ALPHA_BOUNDARY))
        if slivers:
            if pinned:
                assert s2.info(capi.I_H2D_BYTES) < 0.5 * full, "step %d: the zero-copy path must not fall back to the full flux field" % k
            redone = redone or s2.info(capi.I_H2D_BYTES) >= 8 * m.n_faces
        if case == "sphere in an empty box":
            assert s2.info(capi.I_H2D_BYTES) < 0.5 * full < s3.info(capi.I_H2D_BYTES)
            if k >= 2:   # alphaPhi comes back on the marked faces only: less than the full fields
                assert s2.info(capi.I_D2H_BYTES) < 8 * (m.n_cells + m.n_faces)
    if slivers and not pinned:
        assert redone, "the case must drive an empty cell out of bounds (the staged path then redoes the step with the full phi)"
    # the device's phi is not a full field after the sparse upload: the split calls refuse it until svof_set_phi
    with pytest.raises(capi.SvofError):
        s2.advect(dt)
    s2.setPhi(phi0)
    s2.setU(U0, Ub)
    s2.reconstruct()
    s2.advect(dt)
    assert s2.info(capi.I_ERROR_FLAGS) == 0


def _smeared_sphere(C_, V, centre=(0.5, 0.62, 0.5), radius=0.15):
    """A sphere-like field on ANY mesh: one layer of mixed cells (parity input, not an exact shape)."""
    h = np.cbrt(V)
    d = np.linalg.norm(C_ - np.array(centre), axis=1) - radius
    return np.clip(0.5 - d / h, 0.0, 1.0)


_POLY_CASES = {
    "prisms": (lambda: meshmod.prism_mesh(8), {}),
    "refinement-interface polyhedra": (lambda: meshmod.refined_interface_mesh(8), {}),
    "warped hexes": (lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3), {}),
    "warped hexes, splitWarpedFace": (lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3), {"splitWarpedFace": True}),
    "Kelvin cells (14 faces, 24 points: polyDualMesh population)": (lambda: meshmod.kelvin_mesh(10), {}),
    # orientationMethod isoRDF (reconstruction.C:196-405): host-driven iteration on the device, bitwise the oracle's
    "hexes, isoRDF": (lambda: meshmod.hex_block(20), {"orientationMethod": "isoRDF"}),
    "warped hexes, isoRDF": (lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3), {"orientationMethod": "isoRDF"}),
    "Kelvin cells, isoRDF, one iteration": (lambda: meshmod.kelvin_mesh(8), {"orientationMethod": "RDF", "iterations": 1}),
    # orientationMethod alphaGrad (reconstruction.C:74-82) on hexes and on polyhedra with non-orthogonal faces
    "hexes, alphaGrad": (lambda: meshmod.hex_block(14), {"orientationMethod": "alphaGrad"}),
    "refinement-interface polyhedra, alphaGrad": (lambda: meshmod.refined_interface_mesh(8), {"orientationMethod": "alphaGrad"}),
    "warped hexes, alphaGrad": (lambda: meshmod.perturb_points(meshmod.hex_block(12), 0.2, 3), {"orientationMethod": "alphaGrad"}),
}


@pytest.mark.parametrize("case", list(_POLY_CASES))
def test_step_parity_polyhedral(oracle, product, case):
    """BASELINE.json configs[2]: the arbitrary-polyhedron PLIC path (triangular/pentagonal faces, 5..9-face
    cells, warped faces with and without splitWarpedFace), rotating-disc style advection."""
    make, extra = _POLY_CASES[case]
    m = make()
    ctl = dict(LEVEQUE_CONTROLS, **extra)
    so, sg = _pair(m, ctl, oracle, product)
    for f in (capi.F_CF, capi.F_SF, capi.F_C, capi.F_V, capi.F_FACE_FLATNESS):
        assert np.array_equal(so.field(f), sg.field(f)), "mesh field %d differs" % f
    C_, Cf, Sf, V = so.field(capi.F_C), so.field(capi.F_CF), so.field(capi.F_SF), so.field(capi.F_V)
    a0 = _smeared_sphere(C_, V)
    U0 = fields.rotation_velocity(C_)
    phi0 = fields.face_flux(Cf, Sf, fields.rotation_velocity)
    Ub = fields.rotation_velocity(Cf[m.n_internal_faces:])
    dt = 0.25 * np.cbrt(V.min()) / np.abs(U0).max()
    for s in (so, sg):
        s.setAlpha(a0)
        s.setPhi(phi0)
        s.setU(U0, Ub)
    v0 = sg.volume()
    n_mixed = 0
    for k in range(6):
        for s in (so, sg):
            s.reconstruct()
            s.advect(dt)
        assert np.array_equal(so.mixedCells(), sg.mixedCells()), "step %d: interface-cell set differs" % k
        assert np.array_equal(so.cellStatus(), sg.cellStatus())
        for name, f in (("N", capi.F_INTERFACE_N), ("D", capi.F_INTERFACE_D), ("C", capi.F_INTERFACE_C), ("S", capi.F_INTERFACE_S)):
            assert np.array_equal(so.field(f), sg.field(f)), "step %d: interface%s differs" % (k, name)
        assert np.array_equal(so.field(capi.F_UN0), sg.field(capi.F_UN0))
        ao, ag = so.alpha(), sg.alpha()
        assert np.abs(ao - ag).max() <= ATOL_ALPHA
        assert np.array_equal(ao, ag), "step %d: alpha not bitwise equal (max %g)" % (k, np.abs(ao - ag).max())
        assert np.array_equal(so.alphaPhi(), sg.alphaPhi())
        assert so.info(capi.I_N_BOUND_SWEEPS) == sg.info(capi.I_N_BOUND_SWEEPS)
        assert so.info(capi.I_RDF_ITERATIONS) == sg.info(capi.I_RDF_ITERATIONS)
        n_mixed = max(n_mixed, len(sg.mixedCells()))
    assert n_mixed > 50
    assert abs(sg.volume() - v0) <= 1e-12 * abs(v0)   # rotation keeps the shape inside the domain
    assert sg.info(capi.I_ERROR_FLAGS) == 0


def test_overlap_schedule_is_bitwise_identical(product):
    """The two-stream schedule (streaming kernel concurrent with the sparse chain) must not change results."""
    m = meshmod.hex_block(24)
    a0 = exact_sphere_alpha(m)
    res = []
    for overlap in (0, 1):
        s = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
        s.setOption("overlap", overlap)
        C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
        s.setAlpha(a0)
        s.setPhi(fields.face_flux(Cf, Sf))
        s.setU(fields.leveque_velocity(C_))
        for k in range(10):
            s.reconstruct()
            s.advect(0.01)
        res.append((s.alpha(), s.alphaPhi(), s.mixedCells()))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_schedule_selection_of_the_graph_step_is_bitwise_neutral(product, oracle):
    """svof_step_device measures its two schedules at run time (6 + 12 steps of the default, 6 + 12 of the alternative, then the
    decision; `sched_retune` repeats it): every step on the way -- plain-launch capture steps, both schedules' graphs, the
    switch back -- must equal the oracle bitwise, an explicit schedule option must switch the selection off, and the
    result must be reported."""
    m = meshmod.hex_block(24)
    so, sg = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=oracle), SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
    a0 = exact_sphere_alpha(m)
    C, Cf, Sf = so.field(capi.F_C), so.field(capi.F_CF), so.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C), fields.face_flux(Cf, Sf)
    dt = 0.25 / 24
    for s in (so, sg):
        s.setAlpha(a0)
        s.setPhi(phi0)
        s.setU(U0)
    seen = set()
    for k in range(84):
        if k == 42:
            sg.setOption("sched_retune", 1)
        if k == 80:
            sg.setOption("fork", 2)          # explicit choice: selection off, schedule as given
        so.reconstruct()
        so.advect(dt)
        sg.step(dt)
        seen.add(int(sg.info(capi.I_SCHEDULE)))
        assert np.array_equal(so.alpha(), sg.alpha()), "step %d (schedule %d)" % (k, int(sg.info(capi.I_SCHEDULE)))
        assert np.array_equal(so.alphaPhi(), sg.alphaPhi()), "step %d" % k
        if k in (40, 79):
            assert int(sg.info(capi.I_SCHEDULE)) in (100, 204), "settled after 37 calls: %d" % int(sg.info(capi.I_SCHEDULE))
    assert -100 in seen and -204 in seen, "both schedules were measured: %s" % sorted(seen)
    assert int(sg.info(capi.I_SCHEDULE)) in (200, 204)
    assert sg.info(capi.I_ERROR_FLAGS) == 0


def test_graph_step_is_bitwise_the_two_calls(product):
    """svof_step_device (one CUDA-graph launch per step in the steady state) against reconstruct() + advect()."""
    m = meshmod.hex_block(32)
    a, b = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product), SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
    a0 = exact_sphere_alpha(m)
    C, Cf, Sf = a.field(capi.F_C), a.field(capi.F_CF), a.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(C), fields.face_flux(Cf, Sf)
    dt = 0.25 / 32
    for s in (a, b):
        s.setAlpha(a0)
        s.setPhi(phi0)
        s.setU(U0)
    l0 = b.info(capi.I_GPU_LAUNCHES)
    for k in range(9):
        if k == 5:   # inputs refreshed between graph launches
            for s in (a, b):
                s.setPhi(0.5 * phi0)
                s.setU(0.5 * U0)
        a.reconstruct()
        a.advect(dt)
        b.step(dt)
        assert np.array_equal(a.alpha(), b.alpha()), "step %d" % k
        assert np.array_equal(a.alphaPhi(), b.alphaPhi())
    assert np.array_equal(a.mixedCells(), b.mixedCells()) and np.array_equal(a.interfaceD(), b.interfaceD())
    assert a.info(capi.I_N_BOUND_SWEEPS) == b.info(capi.I_N_BOUND_SWEEPS)
    assert b.info(capi.I_GPU_LAUNCHES) - l0 > 9 * 20      # graph launches are counted by the kernels they contain
    b.step(0.5 * dt)                                       # a new dt re-captures
    a.reconstruct()
    a.advect(0.5 * dt)
    assert np.array_equal(a.alpha(), b.alpha())
    assert b.info(capi.I_ERROR_FLAGS) == 0


def test_scatter_alpha_device_matches_set_alpha(product):
    """svof_scatter_alpha_device (the halo refresh of decomposed runs): alpha, patch values and the mixed-cell bitmap
    end up exactly as after svof_set_alpha with the same field, and the graph-replayed steps that follow agree."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    m = meshmod.hex_block(24)
    a, b = SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product), SolveVofEqu(m, LEVEQUE_CONTROLS, lib=product)
    a0 = exact_sphere_alpha(m)
    Cc, Cf, Sf = a.field(capi.F_C), a.field(capi.F_CF), a.field(capi.F_SF)
    U0, phi0 = fields.leveque_velocity(Cc), fields.face_flux(Cf, Sf)
    dt = 0.25 / 24
    for s in (a, b):
        s.setAlpha(a0)
        s.setPhi(phi0)
        s.setU(U0)
        for _ in range(3):
            s.step(dt)
    # overwrite a slab of cells (some become mixed, some stop being mixed) in both solvers, two ways
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(m.n_cells, 3000, replace=False)).astype(np.int32)
    vals = np.where(rng.random(idx.size) < 0.5, rng.random(idx.size), np.round(rng.random(idx.size)))
    cur = a.alpha()
    cur[idx] = vals
    a.setAlpha(cur)
    d_idx, d_vals = torch.as_tensor(idx, device="cuda"), torch.as_tensor(vals, device="cuda")
    torch.cuda.synchronize()
    b._chk(b.lib.svof_scatter_alpha_device(b._h, d_idx.data_ptr(), d_vals.data_ptr(), idx.size))
    assert np.array_equal(a.alpha(), b.alpha())
    assert np.array_equal(a.field(capi.F_ALPHA_BOUNDARY), b.field(capi.F_ALPHA_BOUNDARY))
    for k in range(4):
        a.reconstruct()
        a.advect(dt)
        b.step(dt)
        assert np.array_equal(a.mixedCells(), b.mixedCells()), "step %d" % k
        assert np.array_equal(a.alpha(), b.alpha())
    assert b.info(capi.I_ERROR_FLAGS) == 0
