"""Decks read by the case front end (para + bmap -> particles and LIMITS blocks) run on the device against the CPU
oracle: the droplet deck (Gissler aero) and the jet deck (round inlet with BACK / BUFFER tables inside a Ghost-solver
pipe wall, "Rotation angles" 0,0,90) -- the layout of the reference's Examples/Droplet and Examples/Crossflow."""
import os

import numpy as np
import pytest

from fjsph_b200 import frontend
from tests.util import assert_fields_close, make_pair_from_deck, relerr

pytestmark = pytest.mark.gpu
DECKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks")


def test_droplet_deck_steps():
    case = frontend.read_case(os.path.join(DECKS, "droplet3d.para"), 3)
    o, e = make_pair_from_deck(case)
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt, step
    # the deck's own lattice + U(0, eps dx) positions are a tie-stress input (tests/test_gpu_parity.py): flags exact,
    # state 1e-8, rates 1e-3
    assert_fields_close(e, o, ("surf", "surfzone", "cellID", "b"), context="droplet deck")
    assert_fields_close(e, o, ("xi", "rho"), tol=1e-8, context="droplet deck")
    assert_fields_close(e, o, ("acc", "Rrho", "Af"), tol=1e-3, context="droplet deck")


def test_jet_deck_inlet_and_pipe():
    case = frontend.read_case(os.path.join(DECKS, "jet3d.para"), 3)
    o, e = make_pair_from_deck(case)
    n_add = 0
    for step in range(9):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "jet deck step %d" % step
        assert (se.n_add, se.n_del, se.total_points) == (so.n_add, so.n_del, so.total_points), ctx
        # dt follows max |acc| (a rate: 1e-6 bar); the sub-iteration loop stops on a residual threshold
        assert abs(se.iterations - so.iterations) <= 1 and abs(se.dt - so.dt) <= 1e-6 * so.dt, ctx
        got = e.download(("part_id", "b", "xi", "v", "rho"))
        assert np.array_equal(got["part_id"], o.get("part_id")), ctx
        assert np.array_equal(got["b"], o.get("b")), ctx
        # lattice + U(0, eps dx) positions with every sub-iteration loop running to its limit: a tie-stress input
        # (tests/test_gpu_parity.py), so the bars are the discrete ones above plus a loose one on the state
        # (1e-6 of the 1 mm domain = 1e-5 dx)
        assert relerr(got["xi"], o.get("xi")) <= 1e-6, ctx
        assert relerr(got["rho"], o.get("rho")) <= 1e-6, ctx
        assert relerr(got["v"], o.get("v")) <= 1e-4, ctx
        n_add += se.n_add
    assert n_add > 0  # the run did insert particles at the inlet


def _arch_pair(jitter):
    case = frontend.read_case(os.path.join(DECKS, "arch3d.para"), 3)
    assert [b["name"] for b in case["blocks"]] == ["Trough", "Vault", "Stub", "Water"]
    if jitter:
        rng = np.random.default_rng(11)
        case["xi"] = case["xi"] + rng.uniform(-jitter, jitter, case["xi"].shape) * case["params"].particle_step
    return make_pair_from_deck(case)


def test_arch_deck_steps():
    """Water resting in a trough of Arch blocks (arc.cpp in 3D: a Pressure-Gradient trough with straights, a Ghost vault
    in HCP order on a tilted plane, a stub of straights), particles culled where the blocks intersect: four steps, every
    one running its 20 sub-iterations out.  The oracle follows FJSPH's compiled sources on this deck to 1e-13
    (tests/test_frontend_vs_reference.py).  The deck's own positions -- lattice + U(0, eps dx) -- are a lattice-symmetric
    input: lam, the smallest eigenvalue of L, is a repeated eigenvalue for edge and face particles there, where Eigen's
    closed form takes the square root of an exact cancellation, so a last-bit difference in L (the engine sums its
    neighbours row by row, the reference in KD-tree order) moves lam by sqrt(eps) ~ 1.5e-8 (tests/test_gpu_parity.py,
    TOL_EIGEN_DEGENERATE), and lam feeds the surface normals and the shifting velocity.  Measured on a B200
    (profiles/r2h_arch_probe.txt, tools/arch_probe.py): lam 1.8e-8, xi 3.7e-10, rho 9.8e-8, p 6e-5 (the stiff EOS on rho),
    v 9e-6, rates 1.3e-4 after 4 x 20 sub-iterations; every flag, count and dt exact.  Bars one decade above that.  The same
    deck on generic positions holds the tight bars (xi 3e-17, rho 2e-16, rates 6e-13 measured): next test."""
    o, e = _arch_pair(0.0)
    for step in range(4):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt, step
        assert se.total_points == so.total_points
    assert_fields_close(e, o, ("surf", "surfzone", "b", "part_id"), context="arch deck")
    assert_fields_close(e, o, ("lam",), tol=2e-7, context="arch deck")
    assert_fields_close(e, o, ("xi",), tol=1e-8, context="arch deck")
    assert_fields_close(e, o, ("rho",), tol=1e-6, context="arch deck")
    assert_fields_close(e, o, ("p",), tol=1e-3, context="arch deck")
    assert_fields_close(e, o, ("v",), tol=1e-4, context="arch deck")
    assert_fields_close(e, o, ("acc", "Rrho"), tol=2e-3, context="arch deck")


def test_arch_deck_steps_generic_positions():
    """The Arch deck with every particle moved off its site by U(-0.02, 0.02) dx: no repeated eigenvalues, no neighbours
    on the support edge to the last bit.  Four steps of 20 sub-iterations: state 1e-10, velocity 1e-8, rates 1e-6."""
    o, e = _arch_pair(0.02)
    for step in range(4):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt, step
        assert se.total_points == so.total_points
    assert_fields_close(e, o, ("surf", "surfzone", "b", "part_id"), context="arch deck (generic)")
    assert_fields_close(e, o, ("xi", "rho", "p"), tol=1e-10, context="arch deck (generic)")
    assert_fields_close(e, o, ("v",), tol=1e-8, context="arch deck (generic)")
    assert_fields_close(e, o, ("acc", "Rrho"), tol=1e-6, context="arch deck (generic)")
