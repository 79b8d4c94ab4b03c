"""Parity of the sm_100a kernels (called through the C ABI) against the CPU oracle on the same seeded inputs.
Neighbour sets: bit-exact.  FP64 fields: 1e-10 normwise.  Integer flags: exact."""
import numpy as np
import pytest

from fjsph_b200 import cases
from tests.util import TOL, assert_fields_close, make_pair, relerr

pytestmark = pytest.mark.gpu

PRESTEP_FIELDS = ("L", "gradRho", "norm", "colourG", "kernsum", "colour")
EIGEN_FIELDS = ("lam", "lam_nb")
SURFACE_FIELDS = ("surf", "surfzone", "norm", "curve", "norm_curve", "woccl")
# lam = min eigenvalue from Eigen's closed-form computeDirect (Shifting.cpp:93-99).  Its trigonometric root
# formula takes sqrt(q) of a cancellation that is exactly 0 for a repeated eigenvalue, so on
# lattice-symmetric inputs (edge/face particles have lam repeated) a 1-ulp change of L moves lam by
# ~sqrt(eps) = 1.5e-8: the reference itself (KD-tree summation order) is only defined to that accuracy
# there.  Generic (jittered) inputs keep the 1e-10 bar.
TOL_EIGEN_DEGENERATE = 2e-7


def neighbour_cases():
    yield "lattice_eps", cases.synthetic_block((12, 10, 9), 1.5e-3, jitter="eps", seed=3)
    yield "lattice_exact", cases.synthetic_block((9, 9, 9), 1.0e-3, jitter=None)
    yield "jitter", cases.synthetic_block((14, 9, 11), 1.0e-3, jitter=0.2, seed=9)
    yield "droplet", cases.droplet(dx=0.005)
    yield "walls", cases.box_with_walls(n=(8, 7, 9), dx=0.01)


@pytest.mark.parametrize("name,case", list(neighbour_cases()), ids=[n for n, _ in neighbour_cases()])
def test_neighbour_sets_bit_exact(name, case):
    o, e, p = make_pair(case)
    o.update_neighbours()
    e.update_neighbours()
    off_o, idx_o, _ = o.neighbours()
    off_e, idx_e = e.neighbours()
    assert np.array_equal(off_o, off_e)
    assert np.array_equal(idx_o, idx_e)
    # re-sorting must not disturb what the caller sees
    got = e.download(("xi", "v", "rho", "p", "m", "b", "part_id"))
    assert np.array_equal(got["xi"], case["xi"]) and np.array_equal(got["b"], case["b"])
    assert np.array_equal(got["rho"], case["rho"]) and np.array_equal(got["m"], case["m"])
    assert np.array_equal(got["part_id"], np.arange(case["xi"].shape[0]))


@pytest.mark.parametrize("jitter", ["eps", 0.1], ids=["lattice_eps", "jitter"])
def test_neighbour_sets_through_skin_list(jitter):
    """update_neighbours on moved particles: the exact list filtered from the superset ("skin") list is the
    oracle's set, bit for bit, while the superset is still valid (moves < skin/2), after it has expired
    (moves > skin/2 force a cell-list sweep) and with the superset disabled (skin = 0)."""
    case = cases.synthetic_block((13, 10, 9), 1e-3, jitter=jitter, seed=11)
    o, e, p = make_pair(case)
    dx = 1e-3
    rng = np.random.default_rng(5)
    xi = case["xi"].copy()
    o.update_neighbours()
    e.update_neighbours()
    base = e.integrate_no_update  # noqa: F841  (keeps the engine alive in tracebacks)
    sweeps = []
    for step, amp in enumerate((0.05, 0.05, 0.05, 0.3, 0.01)):
        # amp*dx moves: three small ones (cumulative 0.15 dx < skin/2 = 0.2 dx), one big, one small again
        xi = xi + rng.uniform(-1.0, 1.0, size=xi.shape) * (amp * dx / np.sqrt(3.0))
        o.set("xi", xi)
        e.upload_level(1, xi=xi)
        o.update_neighbours()
        e.update_neighbours()
        off_o, idx_o, _ = o.neighbours()
        off_e, idx_e = e.neighbours()
        assert np.array_equal(off_o, off_e), step
        assert np.array_equal(idx_o, idx_e), step
    e.set_skin(0.0)
    e.update_neighbours()
    off_e, idx_e = e.neighbours()
    assert np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e)


def run_stages(o, e, ale=True, check=True, label="", tol_eigen=TOL):
    o.update_neighbours()
    e.update_neighbours()
    npd_o = o.prestep()
    npd_e = e.dSPH_PreStep()
    if check:
        assert abs(npd_e - npd_o) <= TOL * abs(npd_o)
        assert_fields_close(e, o, PRESTEP_FIELDS, context=label + " prestep")
        assert_fields_close(e, o, EIGEN_FIELDS, tol=tol_eigen, context=label + " prestep eigenvalues")
    o.aero_velocity()
    e.get_aero_velocity()
    if check:
        assert_fields_close(e, o, ("cellV", "cellID"), context=label + " aero_velocity")
    o.detect_surface()
    e.Detect_Surface()
    if check:
        assert_fields_close(e, o, SURFACE_FIELDS, context=label + " detect_surface")
        assert_fields_close(e, o, ("pDist",), tol=tol_eigen, context=label + " detect_surface pDist=lam")
    o.dissipation()
    e.dissipation_terms()
    if check:
        assert_fields_close(e, o, ("aVisc", "deltaD"), context=label + " dissipation")
    if ale:
        o.particle_shift()
        e.particle_shift()
        if check:
            assert_fields_close(e, o, ("vPert",), context=label + " shift")
    return npd_o, npd_e


@pytest.mark.parametrize("ale", [1, 0])
@pytest.mark.parametrize("which", ["block", "droplet", "block_eps"])
def test_stagewise_parity(which, ale):
    if which == "block":
        case = cases.synthetic_block((16, 12, 10), 1e-3, jitter=0.1, seed=1234)
    elif which == "block_eps":
        case = cases.synthetic_block((13, 11, 9), 1e-3, jitter="eps", seed=5)
    else:
        case = cases.droplet(dx=0.004)
    o, e, p = make_pair(case, ale=ale)
    npd_o, npd_e = run_stages(o, e, ale=bool(ale), label=which,
                              tol_eigen=TOL if which == "block" else TOL_EIGEN_DEGENERATE)
    o.forces(npd_o)
    e.get_acc_and_Rrho(npd_e)
    assert_fields_close(e, o, ("acc", "Rrho", "Af"), context=which + " forces")


@pytest.mark.parametrize("ale", [1, 0])
def test_nb_iteration_parity(ale):
    case = cases.synthetic_block((14, 12, 10), 1e-3, jitter=0.1, seed=77)
    o, e, p = make_pair(case, ale=ale, delta_t=1e-6, delta_t_min=1e-6)
    npd_o, npd_e = run_stages(o, e, ale=bool(ale), check=False)
    x0 = o.get("xi").copy()
    o.nb_iter(npd_o)
    err_e = e.Do_NB_Iter(npd_e)
    err_o = float(((o.get("xi") - x0) ** 2).sum())
    assert abs(err_e - err_o) <= 1e-9 * err_o
    assert_fields_close(e, o, ("xi", "v", "rho", "p", "acc", "Rrho"), context="nb_iter")
    # displacement itself (x - x_n) to 1e-10, not just x
    dx_e = e.get("xi") - case["xi"]
    dx_o = o.get("xi") - case["xi"]
    assert relerr(dx_e, dx_o) <= 1e-9


def same_residual(a, b, tol=2e-2):
    """rms_error = log10(rms |x - x_prev|) - logbase (Newmark_Beta.cpp:17-30).  Near convergence |x - x_prev| is
    ~1e-7 of the first displacement, i.e. a few hundred ulps of x, so the residual itself carries a relative
    rounding noise of ~1e-2; both run through +-inf when a displacement is exactly zero (log10 0)."""
    if np.isinf(a) or np.isinf(b) or np.isnan(a) or np.isnan(b):
        return (np.isnan(a) and np.isnan(b)) or a == b
    return abs(a - b) <= tol


# Multi-step bars.  Every stage agrees to 1e-10 on identical inputs (tests above); across sub-iterations the
# weakly-compressible EOS is stiff: a density difference d_rho moves the acceleration by c^2 d_rho / (rho dx),
# ~1e7 x d_rho for the test decks, so summation-order noise of 1e-14 in rho shows up as 1e-8..1e-7 in acc and,
# through dt, 1e-9 in v.  State (x, rho, p) 1e-10, velocity 1e-8, rates 1e-6, flags exact.
STATE_FIELDS = ("xi", "rho", "lam", "lam_nb")
RATE_FIELDS = ("acc", "Rrho", "Af", "aVisc", "deltaD", "vPert")
FLAG_FIELDS = ("surf", "surfzone", "cellID", "b")


def full_step_case(which):
    if which == "block":
        return cases.synthetic_block((14, 11, 9), 1e-3, jitter=0.1, seed=42)
    if which == "droplet":
        return cases.droplet(dx=0.005, jitter=0.05)
    if which == "walls":
        return cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4, jitter=0.05)
    raise KeyError(which)


@pytest.mark.parametrize("solver", [0, 1], ids=["newmark_beta", "runge_kutta"])
@pytest.mark.parametrize("which", ["block", "droplet", "walls"])
def test_full_step_parity(which, solver):
    """Three Integrator::integrate calls on generic (jittered) inputs: every field within 1e-9 of the oracle,
    flags and sub-iteration counts identical."""
    case = full_step_case(which)
    o, e, p = make_pair(case, solver_type=solver, delta_t_min=1e-9)
    for step in range(3):
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "%s solver %d step %d" % (which, solver, step)
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, ctx
        assert abs(se.npd - so.npd) <= TOL * abs(so.npd), ctx
        assert same_residual(se.rms_error, err_o), (ctx, se.rms_error, err_o)
        for a, b in ((se.maxf, so.maxf), (se.maxAf, so.maxAf), (se.maxRho_pc, so.maxRho_pc), (se.maxShift, so.maxShift)):
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-300), ctx
        assert_fields_close(e, o, FLAG_FIELDS, context=ctx)
        assert_fields_close(e, o, STATE_FIELDS, tol=1e-10, context=ctx)
        assert_fields_close(e, o, ("v", "p"), tol=1e-8, context=ctx)  # p = EOS(rho) is stiff: c^2 d_rho
        assert_fields_close(e, o, RATE_FIELDS, tol=1e-6, context=ctx)
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-10, level=0, context=ctx + " pn")
        assert_fields_close(e, o, ("v",), tol=1e-8, level=0, context=ctx + " pn")
        assert_fields_close(e, o, ("acc", "Rrho"), tol=1e-6, level=0, context=ctx + " pn")
    pe, po = e.params, o.params
    assert abs(pe.current_time - po.current_time) <= 1e-12 * po.current_time
    assert pe.cfl == po.cfl and pe.n_stable == po.n_stable and pe.n_unstable == po.n_unstable


@pytest.mark.parametrize("which,cfl", [("block", 6.0), ("walls", 8.0)])
def test_unstable_step_restart_parity(which, cfl):
    """Check_Error's unstable-step branch (Newmark_Beta.cpp:32-48; engine: integrate.cu, fj_integrate_no_update): a CFL
    number this large makes the sub-iterations diverge, so pnp1 = pn (every frozen-term array restored), the list is
    rebuilt mid-step -- with walls, between two wall treatments --, dt halves and the iteration counter restarts.  The
    engine and the oracle (itself pinned on this branch against the compiled reference, tests/test_oracle_vs_reference.py::
    test_unstable_step_restarts, and by tests/golden/ref_*_unstable_restart.npz) take the same number of halvings and
    land on the same state.  A diverging fixed-point iteration amplifies summation-order noise ~2.5x per sub-iteration, so
    the rate bars are those of the stiff jet deck (tests/test_golden_reference.py)."""
    case = full_step_case(which)
    o, e, p = make_pair(case, cfl=cfl, cfl_max=cfl, max_subits=2, delta_t_max=1.0, delta_t_min=1e-12)
    halvings = []
    for step in range(3):
        cfl_now = o.params.cfl
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "%s unstable step %d" % (which, step)
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-9 * so.dt, ctx
        assert se.neighbour_builds >= 2, ctx
        halvings.append(np.log2(cfl_now * so.safe_dt / so.dt))
        assert_fields_close(e, o, FLAG_FIELDS, context=ctx)
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-8, context=ctx)
        assert_fields_close(e, o, ("v", "p"), tol=1e-6, context=ctx)
        assert_fields_close(e, o, ("acc", "Rrho"), tol=1e-4, context=ctx)
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-8, level=0, context=ctx + " pn")
    assert max(halvings) >= 0.99, halvings  # at least one step was halved
    pe, po = e.params, o.params
    assert pe.cfl == po.cfl and pe.n_stable == po.n_stable and pe.n_unstable == po.n_unstable


def test_uniform_drift_keeps_the_superset_list():
    """A block drifting at 30 m/s moves every particle by more than skin/2 per step but no pair apart: the superset
    (skin) list must survive (no cell-list sweep after the first step), and the exact lists filtered from it must
    still be the oracle's sets, bit for bit; the state keeps the full-step parity bars."""
    case = cases.synthetic_block((14, 11, 9), 1e-3, jitter=0.1, seed=43)
    case["v"] = case["v"] + np.array([30.0, 0.0, 0.0])
    o, e, p = make_pair(case, delta_t_min=1e-9)
    x0 = case["xi"][:, 0].mean()
    for step in range(4):
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "drift step %d" % step
        assert se.iterations == so.iterations, ctx
        if step > 0:
            assert se.skin_builds == 0, (ctx, se.skin_builds)
        off_o, idx_o, _ = o.neighbours()
        off_e, idx_e = e.neighbours()
        assert np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e), ctx
        assert_fields_close(e, o, FLAG_FIELDS, context=ctx)
        # the velocity scale is 60x that of the other full-step cases while the block is as small: the same relative
        # noise in v integrates to a 60x larger share of |x| (measured 1.4e-10 after 4 steps)
        assert_fields_close(e, o, ("xi",), tol=1e-9, context=ctx)
        assert_fields_close(e, o, ("rho", "lam", "lam_nb"), tol=1e-10, context=ctx)
        assert_fields_close(e, o, ("v", "p"), tol=1e-8, context=ctx)
        assert_fields_close(e, o, RATE_FIELDS, tol=1e-6, context=ctx)
    moved = e.download(("xi",))["xi"][:, 0].mean() - x0
    assert moved > 0.5e-3, moved  # > skin/2 = 0.2 mm per step: the plain criterion would have swept every step


@pytest.mark.parametrize("which", ["droplet", "walls"])
def test_full_step_tie_stress(which):
    """The reference's own lattice + U(0, eps dx) inputs (square.cpp:103, circle.cpp:136): ~6 of ~257 neighbours sit
    on the support edge to the last bit.  On the SAME positions the sets are bit-exact (test above); once the two
    runs' positions differ by rounding (1e-12 after one sub-iteration) edge members flip.  They carry zero kernel
    weight, but the reference's non-smooth consumers see them (max_j |v_j - v_i| in particle_shift, Shifting.cpp:
    253-256; the Gissler occlusion max, Geometry.cpp:237-246), so per-particle agreement is limited to ~1e-5
    there -- for the reference against itself under a different summation order just the same.  Stated bar:
    integer flags and sub-iteration counts identical, positions/density 1e-8, rates 1e-3."""
    case = cases.droplet(dx=0.005) if which == "droplet" else cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4)
    o, e, p = make_pair(case, delta_t_min=1e-9)
    for step in range(2):
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "%s tie-stress step %d" % (which, step)
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-9 * so.dt, ctx
        assert_fields_close(e, o, ("surf", "surfzone", "cellID", "b"), context=ctx)
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-8, context=ctx)
        assert_fields_close(e, o, ("v", "acc", "Rrho", "vPert"), tol=1e-3, context=ctx)


@pytest.mark.parametrize("ale", [1, 0])
def test_walls_nb_iteration_parity(ale):
    """Do_NB_Iter with Adami pressure walls (Get_Boundary_Pressure, Resid.cpp:21-76) on a moved state: the pair
    distance r stays at its list-build value while Rji follows the positions (r = sqrt(jj.second))."""
    case = cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4, jitter=0.05)
    o, e, p = make_pair(case, ale=ale, delta_t=2e-4, delta_t_min=2e-4)
    npd_o, npd_e = run_stages(o, e, ale=bool(ale), check=False)
    for it in range(3):
        x0 = o.get("xi").copy()
        o.nb_iter(npd_o)
        err_e = e.Do_NB_Iter(npd_e)
        err_o = float(((o.get("xi")[case["bound_points"]:] - x0[case["bound_points"]:]) ** 2).sum())
        assert abs(err_e - err_o) <= 1e-6 * err_o, it
        assert_fields_close(e, o, ("xi", "v", "rho", "p", "acc", "Rrho"), tol=1e-9, context="walls nb_iter %d" % it)


@pytest.mark.parametrize("solver_type", [0, 1], ids=["newmark_beta", "rk4"])
@pytest.mark.parametrize("bound_solver,no_slip", [(0, 0), (0, 1), (2, 0)], ids=["dbc", "dbc_noslip", "ghost"])
def test_dbc_and_ghost_walls_full_steps(bound_solver, no_slip, solver_type):
    """Three full steps with DBC / Ghost walls under both integrators.  The wall densities of these two treatments are
    integrated -- after the forces in Newmark-Beta (Newmark_Beta.cpp:137-241), before them in every Runge-Kutta stage
    (Runge_Kutta.cpp:76-131, 276-349).  DBC has no reference vector (Boundary_DBC writes out of bounds in the reference,
    Resid.cpp:84-107; the contract is the oracle's: wall acc = 0), so the oracle is the yardstick here."""
    case = cases.box_with_walls(n=(8, 6, 7), jitter=0.05)
    nb, n = case["bound_points"], case["xi"].shape[0]
    o, e, p = make_pair(case, solver_type=solver_type, delta_t_min=1e-9)
    o.lib.orc_clear_blocks(o.h)
    o.add_block(0, 0, nb, bound_solver=bound_solver, no_slip=no_slip)
    o.add_block(1, nb, n)
    e.set_blocks([dict(first=0, second=nb, is_fluid=0, bound_solver=bound_solver, no_slip=no_slip),
                  dict(first=nb, second=n, is_fluid=1)])
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "walls %d/%d solver %d step %d" % (bound_solver, no_slip, solver_type, step)
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-9 * so.dt, ctx
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-9, context=ctx)
        assert_fields_close(e, o, ("v", "p"), tol=1e-7, context=ctx)
        assert_fields_close(e, o, ("acc", "Rrho"), tol=1e-5, context=ctx)


@pytest.mark.parametrize("acase", [2, 3], ids=["induced_pressure", "skin_friction"])
def test_other_aero_models(acase):
    """CalcAeroAcc's other two models (Aero.h:106-202 induced pressure, Aero.h:224-257 skin friction) on the droplet in
    a 21.55 m/s freestream: per-particle, no pair loop, evaluated in the prologue of the force kernel."""
    case = cases.droplet(dx=0.005, jitter=0.05)
    o, e, p = make_pair(case, acase=acase, delta_t_min=1e-9)
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations, step
        assert abs(se.maxAf - so.maxAf) <= 1e-6 * so.maxAf, step
        assert_fields_close(e, o, ("xi", "rho"), tol=1e-10, context="aero %d step %d" % (acase, step))
        assert_fields_close(e, o, ("v",), tol=1e-8, context="aero %d step %d" % (acase, step))
        assert_fields_close(e, o, ("Af", "acc"), tol=1e-6, context="aero %d step %d" % (acase, step))
    assert np.abs(o.get("Af")).max() > 0.1   # the model is really acting


def test_state_after_1000_steps():
    """north_star: "state after 1000 steps within a stated tolerance".  1000 Integrator::integrate calls on a jittered
    block (free surfaces on all sides, ALE shifting, surface tension): the sub-iteration count and dt of every step
    and every surface flag must agree, positions to 1e-8 of the domain, density 1e-10, velocity 1e-7 (measured:
    2e-9, 5e-13, 3e-9; tools/long_run.py prints the growth curve, also for the droplet and the walled tank)."""
    case = cases.synthetic_block((10, 9, 8), 1e-3, jitter=0.1, seed=8)
    o, e, p = make_pair(case, delta_t_min=1e-9)
    for step in range(1000):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations, step
        assert abs(se.dt - so.dt) <= 1e-9 * so.dt, step
    assert_fields_close(e, o, ("surf", "surfzone", "b"), context="1000 steps")
    assert_fields_close(e, o, ("xi",), tol=1e-8, context="1000 steps")
    assert_fields_close(e, o, ("rho",), tol=1e-10, context="1000 steps")
    assert_fields_close(e, o, ("v",), tol=1e-7, context="1000 steps")
    displacement = np.abs(o.get("xi") - case["xi"]).max()
    assert displacement > 5 * 1e-3   # the block really moved (several particle spacings)


def test_engine_errors_are_reported_not_fatal():
    from fjsph_b200 import engine as eng
    from fjsph_b200._lib import FjsphError

    p = eng.default_params(3, particle_step=1e-3)
    e = eng.Engine(p, 150)
    with pytest.raises(FjsphError):
        e.dSPH_PreStep()  # no particles / no list
    case = cases.synthetic_block((6, 5, 4), 1e-3)
    with pytest.raises(FjsphError):
        e.upload_state(np.zeros((200, 3)), None, 1000.0, 0.0, 1.0, 5)  # over capacity
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    with pytest.raises(FjsphError):
        e.get_acc_and_Rrho(1.0)  # list not built
    with pytest.raises(FjsphError):
        eng.default_params(4, particle_step=1e-3)  # SIMDIM is 2 or 3
    p2 = eng.default_params(2, particle_step=1e-3)
    p2.grav[2] = -9.81  # a third gravity component in a 2D build
    with pytest.raises(FjsphError):
        eng.Engine(p2, 10)
