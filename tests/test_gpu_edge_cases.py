"""Edge cases of the neighbour build and the step against the CPU oracle: isolated particles, pairs on the support
edge (strict `<`, Neighbours.cpp:15-47), neighbour counts far above the default list capacity (both lists grow and the
build retries), ragged cells, and a case with no fluid at all."""
import numpy as np
import pytest

from fjsph_b200 import cases
from tests.util import assert_fields_close, make_pair

pytestmark = pytest.mark.gpu


def tiny_case(xi, dx):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    n = xi.shape[0]
    return dict(xi=xi, v=np.zeros_like(xi), rho=np.full(n, 1000.0), p=np.zeros(n), m=np.full(n, 1000.0 * dx**3),
                b=np.full(n, cases.FREE, dtype=np.int32), bound_points=0,
                params=dict(particle_step=dx, rho_rest=1000.0, speed_sound=100.0, delta_t_min=1e-9))


def same_lists(o, e):
    off_o, idx_o, _ = o.neighbours()
    off_e, idx_e = e.neighbours()
    return np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e)


def test_single_particle_and_isolated_pair():
    o, e, _ = make_pair(tiny_case([[0.0, 0.0, 0.0]], 1e-3))
    o.update_neighbours()
    e.update_neighbours()
    assert same_lists(o, e) and e.neighbour_counts().tolist() == [1]  # the list of the reference holds the particle itself
    so = o.integrate()[1]
    se = e.integrate()
    assert se.iterations == so.iterations and se.dt == so.dt
    assert_fields_close(e, o, ("xi", "v", "rho", "acc", "Rrho", "lam", "surf"), context="single particle")
    # two particles far outside each other's support: two lists of one
    o, e, _ = make_pair(tiny_case([[0.0, 0.0, 0.0], [0.05, 0.0, 0.0]], 1e-3))
    o.update_neighbours()
    e.update_neighbours()
    assert same_lists(o, e) and e.neighbour_counts().tolist() == [1, 1]


def test_support_edge_is_exclusive():
    """H = 2 dx = 1.0 exactly, sr = 4 H^2 = 4.0: a partner at distance exactly 2.0 is NOT a neighbour (d2 < sr), one ulp
    closer is."""
    dx = 0.5
    for x1, expect in ((2.0, [1, 1]), (np.nextafter(2.0, 0.0), [2, 2]), (np.nextafter(2.0, 3.0), [1, 1])):
        o, e, p = make_pair(tiny_case([[0.0, 0.0, 0.0], [x1, 0.0, 0.0]], dx))
        assert p.sr == 4.0
        o.update_neighbours()
        e.update_neighbours()
        assert same_lists(o, e), x1
        assert e.neighbour_counts().tolist() == expect, x1


def test_lists_grow_past_their_default_capacity():
    """A block compressed to 0.62 of the nominal spacing holds ~1100 neighbours per particle, against a default exact-list
    capacity of 288 and a superset capacity of ~416: both builds overflow, re-allocate and retry; the sets stay exact."""
    dx = 1e-3
    case = cases.synthetic_block((13, 12, 11), dx, jitter=0.1, seed=5)
    case["xi"] = case["xi"] * 0.62
    o, e, _ = make_pair(case)
    o.update_neighbours()
    e.update_neighbours()
    counts = e.neighbour_counts()
    assert counts.max() > 900
    assert same_lists(o, e)
    npd_o, npd_e = o.prestep(), e.dSPH_PreStep()
    assert abs(npd_e - npd_o) <= 1e-10 * abs(npd_o)
    assert_fields_close(e, o, ("L", "gradRho", "kernsum", "lam"), context="dense block prestep")
    # ... and shrink back to ordinary lists when the particles spread out again
    e.upload_level(1, xi=case["xi"] / 0.62)
    o.set("xi", case["xi"] / 0.62)
    o.update_neighbours()
    e.update_neighbours()
    assert same_lists(o, e) and e.neighbour_counts().max() < 300


def test_ragged_cells_and_clusters():
    """Three clusters of very different density far apart on a sparse cell grid, plus stragglers: empty cells, cells with
    one particle and cells with a hundred and more."""
    rng = np.random.default_rng(11)
    dx = 1e-3
    a = cases.lattice((9, 8, 7), dx, jitter=0.2, seed=1)
    b = cases.lattice((6, 6, 6), 0.7 * dx, start=(0.05, 0.0, 0.01), jitter=0.2, seed=2)
    c = cases.lattice((5, 4, 3), 1.6 * dx, start=(0.0, 0.04, 0.03), jitter=0.2, seed=3)
    stray = rng.uniform(0.0, 0.06, size=(25, 3))
    xi = np.concatenate([a, b, c, stray])
    o, e, _ = make_pair(tiny_case(xi[rng.permutation(len(xi))], dx))
    o.update_neighbours()
    e.update_neighbours()
    assert same_lists(o, e)
    counts = e.neighbour_counts()
    assert counts.min() == 1 and counts.max() > 200


def test_walls_only_case_steps_without_fluid():
    """No fluid particle at all: Integrator::integrate returns early (Integration.cpp:244-246); the engine must not
    divide by the zero fluid count or launch empty grids."""
    case = cases.box_with_walls(n=(5, 4, 4), dx=0.01, layers=2, jitter=0.05)
    nb = case["bound_points"]
    walls = {k: (v[:nb] if isinstance(v, np.ndarray) and v.shape[:1] == case["xi"].shape[:1] else v) for k, v in case.items()}
    walls["bound_points"] = nb
    o, e, _ = make_pair(walls)
    se = e.integrate()
    got = e.download(("xi", "v", "rho"))
    assert np.array_equal(got["xi"], walls["xi"]) and np.isfinite(got["rho"]).all() and se.total_points == nb


@pytest.mark.parametrize("solver", [0, 1], ids=["newmark_beta", "rk4"])
def test_step_host_round_trip_keeps_the_lists_and_the_answer(solver, monkeypatch):
    """fjsph_step_host (upload -> integrate -> download with host buffers): a host that feeds every step's output back as
    the next input must get the device-resident run, and the engine must keep its cell order and superset list across
    those uploads (no cell-list sweep after the first step).  A different particle set of the same size is noticed by the
    displacement test and handled like a fresh upload.  The upload crosses PCIe in three parts (x | rho, m, b | the rest)
    with the neighbour build and dSPH_PreStep running beside the later parts (abi.cu, upload_state_split): the results are
    those of the two-part form (FJSPH_B200_UPLOAD_PARTS=2: prestep after the whole upload), bit for bit, NB and RK4.  (With
    RK4 a host that hands back x, v, acc, rho, Rrho, p only does NOT get the device-resident run: Get_First_RK starts from
    pn's frozen terms of the step before, Runge_Kutta.cpp:462-476, which such a host does not carry -- so that comparison is
    made for Newmark-Beta only.)"""
    from fjsph_b200 import engine as eng

    case = cases.synthetic_block((14, 11, 9), 1e-3, jitter=0.1, seed=21)
    params = eng.default_params(3, **dict(case["params"], delta_t_min=1e-9, solver_type=solver))
    n = case["xi"].shape[0]
    a = eng.Engine(params, n)
    a.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    for _ in range(4):
        a.integrate()
    want = a.download(("xi", "v", "rho", "p", "acc", "Rrho"))
    out_fields = ("xi", "v", "acc", "rho", "Rrho", "p")

    def round_trips(engine):
        state = dict(xi=case["xi"], v=case["v"], acc=np.zeros_like(case["xi"]), rho=case["rho"], Rrho=np.zeros(n), p=case["p"],
                     m=case["m"], b=case["b"])
        sweeps = []
        for _ in range(4):
            out, st = engine.step_host(state, 0, 1, out_fields=out_fields)
            state.update(out)
            sweeps.append(st.skin_builds)
        assert sweeps[0] >= 1 and sweeps[1:] == [0, 0, 0], sweeps
        return state

    b = eng.Engine(params, n)
    state = round_trips(b)
    monkeypatch.setenv("FJSPH_B200_UPLOAD_PARTS", "2")
    two = eng.Engine(params, n)
    monkeypatch.delenv("FJSPH_B200_UPLOAD_PARTS")
    state2 = round_trips(two)
    for f in out_fields:
        assert np.array_equal(state[f], state2[f]), f
    from tests.util import relerr

    if solver == 0:
        for f, tol in (("xi", 1e-10), ("rho", 1e-10), ("v", 1e-8), ("p", 1e-8), ("acc", 1e-6), ("Rrho", 1e-6)):
            assert relerr(state[f], want[f]) <= tol, (f, relerr(state[f], want[f]))
    # another particle set of the same size: the kept list must not be trusted
    other = cases.synthetic_block((14, 11, 9), 1e-3, jitter=0.1, seed=22)
    st2 = dict(xi=other["xi"], v=other["v"], acc=np.zeros_like(other["xi"]), rho=other["rho"], Rrho=np.zeros(n), p=other["p"],
               m=other["m"], b=other["b"])
    out, st = b.step_host(dict(st2), 0, 1, out_fields=("xi", "v", "rho"))
    assert st.skin_builds >= 1
    out2, _ = two.step_host(dict(st2), 0, 1, out_fields=("xi", "v", "rho"))
    for f in ("xi", "v", "rho"):  # the re-sort waits for the whole upload: three parts or two, the same step
        assert np.array_equal(out[f], out2[f]), f
    if solver == 0:
        c = eng.Engine(params, n)
        c.upload_state(other["xi"], other["v"], other["rho"], other["p"], other["m"], other["b"])
        c.integrate()
        fresh = c.download(("xi", "v", "rho"))
        for f, tol in (("xi", 1e-10), ("rho", 1e-10), ("v", 1e-8)):
            assert relerr(out[f], fresh[f]) <= tol, f
    assert np.array_equal(b.download(("part_id",))["part_id"], np.arange(n))
