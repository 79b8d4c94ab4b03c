"""Inlet buffer regions, insertion and delete planes on the device against the CPU oracle
(Newmark_Beta.cpp:243-297, Runge_Kutta.cpp:175-228, shapes/inlet.cpp:578-640, Integration.cpp:109-226).
The caller's particle order must stay the reference's through insert (end of the block) and erase."""
import numpy as np
import pytest

from fjsph_b200 import cases, engine as eng
from oracle import oracle as orc
from tests.util import relerr

pytestmark = pytest.mark.gpu


def make_inlet_pair(case, **kw):
    params = dict(case["params"], **kw)
    B = case["block"]
    o = orc.Oracle(orc.default_params(3, **params))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    o.lib.orc_clear_blocks(o.h)
    o.add_block(1, B["first"], B["second"], block_type=6, fixed_vel_or_dynamic=B["fixed_vel_or_dynamic"],
                insert_norm=B["insert_norm"], insconst=B["insconst"], delete_norm=B.get("delete_norm"),
                delconst=B.get("delconst", 9999999.0), aero_norm=B["aero_norm"], aeroconst=B["aeroconst"],
                back=B["back"], buffer=B["buffer"])
    e = eng.Engine(eng.default_params(3, **params), 4 * case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    e.set_blocks([B])
    return o, e


@pytest.mark.parametrize("fixed,solver", [(0, 0), (1, 0), (1, 1)], ids=["dynamic-nb", "fixed-nb", "fixed-rk"])
def test_inlet_insertion_and_delete_plane(fixed, solver):
    case = cases.inlet_jet(n=(5, 5, 4), fixed=fixed, delete_x=2.5, jitter=0.03)
    o, e = make_inlet_pair(case, solver_type=solver)
    n_add = n_del = 0
    for step in range(14):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "fixed %d solver %d step %d" % (fixed, solver, step)
        assert (se.n_add, se.n_del, se.total_points) == (so.n_add, so.n_del, so.total_points), ctx
        assert se.iterations == so.iterations, ctx
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, ctx
        assert e.n == o.n, ctx
        got = e.download(("part_id", "b", "xi", "v", "rho", "p"))
        assert np.array_equal(got["part_id"], o.get("part_id")), ctx   # same particles in the same order
        assert np.array_equal(got["b"], o.get("b")), ctx
        assert relerr(got["xi"], o.get("xi")) <= 1e-10, ctx
        assert relerr(got["rho"], o.get("rho")) <= 1e-10, ctx
        assert relerr(got["v"], o.get("v")) <= 1e-8, ctx
        n_add += se.n_add
        n_del += se.n_del
    assert n_add >= 50 and n_del >= 25   # the run did insert columns and did erase particles


def test_deleted_particles_are_handed_off():
    """IPT hand-off (Integration.cpp:151-169, Var.h:733-763): every particle past the delete plane is erased from the SPH
    state and handed to the host with what IPTPart(SPHPart, time, ...) copies -- id, time, position, velocity, mass, cell
    id, cell velocity and density -- in the reference's order (ascending index, step after step).  Checked against the
    oracle: the ids that leave its particle set at a step are the ids handed off at that step, with the state the
    oracle's particles had when they crossed the plane."""
    case = cases.inlet_jet(n=(5, 5, 4), fixed=1, delete_x=2.5, jitter=0.03)
    o, e = make_inlet_pair(case)
    handed = 0
    for step in range(14):
        _, so = o.integrate()
        t_before = e.params.current_time
        se = e.integrate()
        d = e.take_deleted()
        assert d["part_id"].shape[0] == se.n_del == so.n_del, step
        if se.n_del == 0:
            continue
        handed += se.n_del
        alive = set(o.get("part_id").tolist())
        assert not (set(d["part_id"].tolist()) & alive), step          # gone from the SPH set ...
        assert np.all(d["xi"][:, 0] > 2.5 * case["params"]["particle_step"]), step   # ... because past the plane
        assert np.all(np.diff(d["part_id"]) != 0) and np.allclose(d["t"], t_before), step
        assert np.allclose(d["mass"], case["m"][0]) and np.all(d["v"][:, 0] > 0.0), step
    assert handed >= 25
    assert e.take_deleted()["part_id"].shape[0] == 0                    # the queue was emptied


def test_runge_kutta_dynamic_inlet_is_rejected():
    from fjsph_b200._lib import FjsphError

    case = cases.inlet_jet(n=(3, 3, 3), fixed=0)
    o, e = make_inlet_pair(case, solver_type=1)
    with pytest.raises(FjsphError):
        e.integrate()
