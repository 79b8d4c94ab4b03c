"""Parity of the sm_100a kernels (called through the C ABI) against the CPU oracle on the same seeded inputs.
Neighbour sets: bit-exact.  FP64 fields: 1e-10 normwise.  Integer flags: exact."""
import numpy as np
import pytest

from fjsph_b200 import cases
from tests.util import TOL, assert_fields_close, make_pair, relerr

pytestmark = pytest.mark.gpu

PRESTEP_FIELDS = ("L", "gradRho", "norm", "lam", "lam_nb", "colourG", "kernsum", "colour")
SURFACE_FIELDS = ("surf", "surfzone", "norm", "curve", "norm_curve", "woccl", "pDist")


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


def run_stages(o, e, ale=True, check=True, label=""):
    o.update_neighbours()
    e.update_neighbours()
    npd_o = o.prestep()
    npd_e = e.dSPH_PreStep()
    if check:
        assert abs(npd_e - npd_o) <= TOL * abs(npd_o)
        assert_fields_close(e, o, PRESTEP_FIELDS, context=label + " prestep")
    o.aero_velocity()
    e.get_aero_velocity()
    if check:
        assert_fields_close(e, o, ("cellV", "cellID"), context=label + " aero_velocity")
    o.detect_surface()
    e.Detect_Surface()
    if check:
        assert_fields_close(e, o, SURFACE_FIELDS, context=label + " detect_surface")
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
    npd_o, npd_e = run_stages(o, e, ale=bool(ale), label=which)
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


@pytest.mark.parametrize("solver", [0, 1], ids=["newmark_beta", "runge_kutta"])
@pytest.mark.parametrize("which", ["block", "droplet"])
def test_full_step_parity(which, solver):
    if which == "block":
        case = cases.synthetic_block((14, 11, 9), 1e-3, jitter=0.1, seed=42)
    else:
        case = cases.droplet(dx=0.005)
    o, e, p = make_pair(case, solver_type=solver, delta_t_min=1e-9)
    for step in range(3):
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "%s solver %d step %d" % (which, solver, step)
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, ctx
        assert abs(se.npd - so.npd) <= TOL * abs(so.npd), ctx
        assert abs(se.rms_error - err_o) <= 1e-6, ctx
        for a, b in ((se.maxf, so.maxf), (se.maxAf, so.maxAf), (se.maxRho_pc, so.maxRho_pc), (se.maxShift, so.maxShift)):
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-300), ctx
        fields = ("xi", "v", "rho", "p", "acc", "Rrho", "Af", "aVisc", "deltaD", "vPert", "surf", "surfzone", "cellID", "b")
        assert_fields_close(e, o, fields, tol=1e-9, context=ctx)
        assert_fields_close(e, o, ("xi", "v", "rho", "acc", "Rrho"), tol=1e-9, level=0, context=ctx + " pn")
    pe, po = e.params, o.params
    assert abs(pe.current_time - po.current_time) <= 1e-12 * po.current_time
    assert pe.cfl == po.cfl and pe.n_stable == po.n_stable and pe.n_unstable == po.n_unstable


@pytest.mark.parametrize("solver", [0, 1], ids=["newmark_beta", "runge_kutta"])
def test_walls_adami_pressure_parity(solver):
    case = cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4)
    o, e, p = make_pair(case, ale=1, solver_type=solver)
    for step in range(2):
        err_o, so = o.integrate()
        se = e.integrate()
        ctx = "walls solver %d step %d" % (solver, step)
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, ctx
        assert_fields_close(e, o, ("xi", "v", "rho", "p", "acc", "Rrho", "lam", "lam_nb", "surf", "surfzone"), tol=1e-9,
                            context=ctx)


def test_engine_errors_are_reported_not_fatal():
    from fjsph_b200 import engine as eng
    from fjsph_b200._lib import FjsphError

    p = eng.default_params(3, particle_step=1e-3)
    e = eng.Engine(p, 100)
    with pytest.raises(FjsphError):
        e.dSPH_PreStep()  # no particles / no list
    case = cases.synthetic_block((6, 5, 4), 1e-3)
    with pytest.raises(FjsphError):
        e.upload_state(np.zeros((200, 3)), None, 1000.0, 0.0, 1.0, 5)  # over capacity
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    with pytest.raises(FjsphError):
        e.get_acc_and_Rrho(1.0)  # list not built
    with pytest.raises(FjsphError):
        eng.Engine(eng.default_params(2, particle_step=1e-3), 10)  # 2D is oracle-only
