"""The reference's own solveVofEqu class, from a committed fixture.

tests/golden/reference_solver_steps.npz holds what the REFERENCE's class (src/SimPLIC compiled unmodified against the OpenFOAM
stand-in: oracle/_ref/libref_solver.so) produced on eleven cases -- per step the interface-cell list, the cut status,
interfaceN / interfaceD at the interface cells, alpha and alphaPhi -- written by tests/golden/make_reference_solver_golden.py
in the build container.  Replayed here with the oracle (CPU) and the CUDA library (GPU): BITWISE equal, every step.
Unlike tests/test_reference_solver.py this needs neither /root/reference nor the prebuilt reference library."""
import os
import sys

import numpy as np
import pytest

from common import LEVEQUE_CONTROLS, SolveVofEqu, capi, oracle_lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_reference_solver_golden as gen  # noqa: E402

G = np.load(gen.OUT)


def _replay(name, lib, what):
    make, extra, _, _, steps, _ = gen.CASES[name]
    m = make()
    nC, nF, nP, n_steps = [int(x) for x in G[name + "/sizes"]]
    assert (m.n_cells, m.n_faces, m.n_points, steps) == (nC, nF, nP, n_steps), "the mesh generator changed: regenerate the fixture"
    s = SolveVofEqu(m, dict(LEVEQUE_CONTROLS, **extra), lib=lib)
    a0, U0, Ub, phi0, dt = gen.case_inputs(name, s)
    assert np.array_equal(a0, G[name + "/a0"]) and dt == float(G[name + "/dt"][0]), "%s: inputs differ from the fixture's" % what
    s.setPhi(phi0)
    s.setAlpha(a0)
    s.setU(U0, Ub)
    for k in range(steps):
        pre = "%s/%d/" % (name, k)
        s.reconstruct()
        mc = G[pre + "mixed"]
        tag = "%s, %s, step %d" % (what, name, k)
        assert np.array_equal(s.mixedCells(), mc), tag + ": interface-cell list"
        assert np.array_equal(s.cellStatus(), G[pre + "status"]), tag + ": cut status"
        assert np.array_equal(s.interfaceN()[mc], G[pre + "N"]), tag + ": interfaceN"
        assert np.array_equal(s.interfaceD()[mc], G[pre + "D"]), tag + ": interfaceD"
        s.advect(dt)
        a_ref = np.zeros(nC)
        a_ref[G[pre + "full"]] = 1.0
        a_ref[G[pre + "part_idx"]] = G[pre + "part_val"]
        ap_ref = np.zeros(nF)
        ap_ref[G[pre + "aphi_idx"]] = G[pre + "aphi_val"]
        a, ap = s.alpha(), s.alphaPhi()
        assert np.array_equal(a, a_ref), tag + ": alpha differs from the reference class's by %g" % np.abs(a - a_ref).max()
        assert np.array_equal(ap, ap_ref), tag + ": alphaPhi differs by %g" % np.abs(ap - ap_ref).max()
    flags = s.info(capi.I_ERROR_FLAGS)
    s.close()
    assert flags == 0


@pytest.mark.parametrize("name", list(gen.CASES))
def test_oracle_replays_the_reference_class_fixture(name):
    _replay(name, oracle_lib(), "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(gen.CASES))
def test_gpu_replays_the_reference_class_fixture(name, product):
    _replay(name, product, "CUDA")
