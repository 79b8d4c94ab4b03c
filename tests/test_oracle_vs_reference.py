"""The oracle restatement against FJSPH's own sources, live.

oracle/_ref/liborc_ref*.so are the reference's time-step translation units compiled unmodified from /root/reference/src
against stand-in Eigen / nanoflann headers (oracle/Makefile.ref, oracle/ref_harness.cpp), exposing the same orc_* ABI as
the oracle.  Built in the container that has the reference; the libraries travel to the GPU box with the snapshot.  Where
neither the libraries nor the reference exist these tests skip and tests/test_golden_reference.py (committed vectors from
the same libraries) carries the pin.
"""
import numpy as np
import pytest

from fjsph_b200 import cases
from oracle import oracle as orc
from tests.util import relerr


def _have(kind):
    if orc.have_ref(kind):
        return True
    try:
        orc.build_ref()
    except Exception:
        return False
    return orc.have_ref(kind)


pytestmark = pytest.mark.skipif(not _have("ref3d"), reason="oracle/_ref not built (no /root/reference here)")

FLOATS = orc._VEC_FIELDS + ("L",) + tuple(orc._SCALAR_FIELDS)
INTS = ("part_id", "cellID", "b", "surf", "surfzone", "internal", "ipt_n_failed")


def pair(case, kind, dim=3, **kw):
    out = []
    for k in (None if dim == 3 else "2d", kind):
        o = orc.Oracle(orc.default_params(dim, **dict(case["params"], **kw)), kind=k)
        o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case.get("bound_points", 0))
        out.append(o)
    return out


def assert_same(a, r, fields, tol, ctx, level=1):
    for f in fields:
        x, y = a.get(f, level), r.get(f, level)
        if x.dtype.kind in "iu":
            assert np.array_equal(x, y), (ctx, f)
        else:
            e = relerr(x, y)
            assert e <= tol, "%s: %s differs from the reference by %.3e" % (ctx, f, e)


@pytest.mark.parametrize("dim,kind,kw", [
    (3, "ref3d", {}), (3, "ref3d", dict(particle_step=3e-5, speed_sound=300.0, rho_rest=810.0, press_pipe=2000.0, H_fac=1.7)),
    (3, "ref3d", dict(pressure_rel=1, press_pipe=500.0, press_back=100.0, gam=1.0)), (2, "ref2d", dict(particle_step=0.02)),
])
def test_set_values_constants(dim, kind, kw):
    """Every constant Set_Values derives (IO.cpp:26-128), through the reference's own get_density / Kernel / GetYcoef /
    get_n_full, equals the restatement's."""
    if not _have(kind):
        pytest.skip(kind)
    P = orc.default_params(dim, kind=None if dim == 3 else "2d", ale=1, **dict(dict(particle_step=1e-3), **kw))
    Q = orc.ref_set_values(P, kind)
    for name, _ in P._fields_:
        a, b = getattr(P, name), getattr(Q, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b or abs(a - b) <= 1e-14 * abs(b), (name, a, b)


def test_eos_and_kernel():
    P = orc.default_params(3, particle_step=1e-3, press_back=250.0)
    lib = orc._load("ref3d")
    import ctypes as C
    for rho in (950.0, 1000.0, 1000.5, 1049.0):
        p_ref = lib.orc_ref_pressure(C.byref(P), rho)
        assert abs(cases.cole_pressure(rho, 1000.0, P.speed_sound, P.gam, 250.0) - p_ref) <= 1e-12 * max(abs(p_ref), 1.0)
        assert abs(lib.orc_ref_density(C.byref(P), p_ref) - rho) <= 1e-12 * rho
    for r in (0.0, 0.3e-3, 1e-3, 3.9e-3):
        assert orc.kernel(r, P.H, P.W_correc) == lib.orc_kernel(r, P.H, P.W_correc)
    assert orc.get_n_full(1e-3, 2e-3) == lib.orc_get_n_full(1e-3, 2e-3)


@pytest.mark.parametrize("kind,ale", [("ref3d", 1), ("ref3d_dsph", 0)])
def test_every_stage_on_identical_inputs(kind, ale):
    """update_neighbours, dSPH_PreStep, get_aero_velocity, Detect_Surface, dissipation_terms, particle_shift,
    get_acc_and_Rrho, Do_NB_Iter and find_timestep, one after the other as integrate_no_update calls them."""
    if not _have(kind):
        pytest.skip(kind)
    case = cases.box_with_walls(n=(9, 7, 8), jitter=0.05)
    a, r = pair(case, kind, ale=ale, acase=1, v_inf=(3.0, 0.0, 0.0))
    for o in (a, r):
        o.update_neighbours()
    for x, y in zip(a.neighbours(), r.neighbours()):
        assert np.array_equal(x, y)          # offsets, indices and d^2, bit for bit
    na, nr = a.prestep(), r.prestep()
    assert abs(na - nr) <= 1e-13 * nr
    assert_same(a, r, ("L", "gradRho", "norm", "lam", "lam_nb", "colourG", "colour", "kernsum"), 1e-13, "prestep")
    for o in (a, r):
        o.aero_velocity()
    assert_same(a, r, ("cellV", "cellID"), 0.0, "aero velocity")
    for o in (a, r):
        o.detect_surface()
    assert_same(a, r, ("surf", "surfzone", "norm", "curve", "norm_curve", "woccl", "pDist"), 1e-13, "surface")
    for o in (a, r):
        o.dissipation()
    assert_same(a, r, ("aVisc", "deltaD"), 1e-13, "dissipation")
    for o in (a, r):
        o.particle_shift()
    assert_same(a, r, ("vPert",), 1e-13, "shifting")
    a.forces(na), r.forces(nr)
    assert_same(a, r, ("acc", "Af", "Rrho"), 1e-13, "forces")
    for o in (a, r):
        o.set_params(delta_t=1e-5)
    a.nb_iter(na), r.nb_iter(nr)
    assert_same(a, r, ("xi", "v", "rho", "p", "acc", "Rrho"), 1e-13, "Do_NB_Iter")
    assert abs(a.find_timestep() - r.find_timestep()) <= 1e-13 * r.find_timestep()


def step_cases():
    blk = cases.synthetic_block(n=(12, 10, 9), jitter=0.1)
    drop = cases.droplet(dx=0.006, jitter=0.05)
    tank = cases.box_with_walls(n=(8, 6, 7), jitter=0.05)
    yield "block_nb", blk, "ref3d", 3, dict(ale=1)
    yield "block_rk4", blk, "ref3d", 3, dict(ale=1, solver_type=1)
    yield "block_ties", cases.synthetic_block(n=(12, 10, 9), jitter="eps"), "ref3d", 3, dict(ale=1)
    yield "droplet_gissler", drop, "ref3d", 3, dict(ale=1)
    yield "droplet_tab", drop, "ref3d", 3, dict(ale=1, use_TAB_def=1)
    yield "droplet_neighbour_count_aero", drop, "ref3d", 3, dict(ale=1, use_lam=0)
    yield "droplet_induced_pressure", drop, "ref3d", 3, dict(ale=1, acase=2)
    yield "droplet_skin_friction", drop, "ref3d", 3, dict(ale=1, acase=3)
    yield "droplet_dsph", drop, "ref3d_dsph", 3, dict(ale=0)
    yield "tank_nb", tank, "ref3d", 3, dict(ale=1)
    yield "tank_rk4", tank, "ref3d", 3, dict(ale=1, solver_type=1)
    yield "tank_iso", tank, "ref3d", 3, dict(ale=1, pressure_rel=1)
    yield "tank_dsph", tank, "ref3d_dsph", 3, dict(ale=0)
    yield "dam_2d", cases.dam_2d(dx=0.05), "ref2d", 2, dict(ale=1)


@pytest.mark.parametrize("name,case,kind,dim,kw", list(step_cases()), ids=[c[0] for c in step_cases()])
def test_full_steps(name, case, kind, dim, kw):
    """Three Integrator::integrate calls: same sub-iterations and dt, every SPHPart field of both time levels."""
    if not _have(kind):
        pytest.skip(kind)
    a, r = pair(case, kind, dim=dim, **kw)
    for step in range(3):
        ea, sa = a.integrate()
        er, sr = r.integrate()
        ctx = "%s step %d" % (name, step)
        assert sa.iterations == sr.iterations and sa.total_points == sr.total_points, ctx
        assert abs(sa.dt - sr.dt) <= 1e-12 * sr.dt, ctx
        # maxShift exists in the -DALE binary only (Integration.h:66-68)
        for k in ("maxf", "maxAf", "maxRho_pc", "safe_dt") + (("maxShift",) if kw.get("ale") else ()):
            assert abs(getattr(sa, k) - getattr(sr, k)) <= 1e-9 * max(abs(getattr(sr, k)), 1e-300), (ctx, k)
        assert (ea == er) or abs(ea - er) <= 1e-6, ctx
    for level in (0, 1):
        assert_same(a, r, INTS, 0.0, name, level)
        assert_same(a, r, FLOATS, 1e-9, name, level)
    pa, pr = a.params, r.params
    assert pa.cfl == pr.cfl and pa.n_stable == pr.n_stable and pa.n_unstable == pr.n_unstable   # the CFL controller
    assert abs(pa.current_time - pr.current_time) <= 1e-12 * pr.current_time


def test_eighty_steps_track_the_reference():
    """A longer horizon: 80 Integrator::integrate calls on the jittered block (every particle near a free surface, the CFL
    controller stepping down on the way).  Same sub-iterations and time step at every step, same flags at the end, positions
    to 1e-12 (measured 2e-15 after 150 steps)."""
    blk = cases.synthetic_block(n=(9, 8, 7), jitter=0.1)
    a, r = pair(blk, "ref3d", ale=1, delta_t_min=1e-9)
    for step in range(80):
        _, sa = a.integrate()
        _, sr = r.integrate()
        assert sa.iterations == sr.iterations and abs(sa.dt - sr.dt) <= 1e-10 * sr.dt, step
    assert a.params.cfl == r.params.cfl and a.params.cfl < 1.0
    assert_same(a, r, INTS, 0.0, "80 steps")
    assert_same(a, r, ("xi", "rho"), 1e-12, "80 steps")
    assert_same(a, r, ("v", "p"), 1e-10, "80 steps")
    assert_same(a, r, ("acc", "Rrho", "vPert"), 1e-8, "80 steps")


@pytest.mark.parametrize("which,cfl,subits", [("block", 6.0, 2), ("block", 12.0, 4), ("tank", 8.0, 2)])
def test_unstable_step_restarts(which, cfl, subits):
    """Check_Error's unstable-step branch (Newmark_Beta.cpp:32-48), part of the integrator contract: with a CFL
    number this large the sub-iterations diverge (rms_error > 0 past max_subits), so pnp1 = pn, the list is rebuilt, dt
    halves, the iteration counter restarts -- as often as it takes.  dt = cfl * safe_dt / 2^k shows k >= 1 halvings; the
    restatement and the reference agree on k, on the restored state and on the CFL controller's reaction."""
    case = cases.synthetic_block(n=(10, 9, 8), jitter=0.1) if which == "block" else cases.box_with_walls(n=(7, 6, 6), jitter=0.05)
    a, r = pair(case, "ref3d", ale=1, cfl=cfl, cfl_max=cfl, max_subits=subits, delta_t_max=1.0, delta_t_min=1e-12)
    halvings = []
    for step in range(3):
        cfl_now = r.params.cfl
        _, sa = a.integrate()
        _, sr = r.integrate()
        ctx = "%s step %d" % (which, step)
        assert sa.iterations == sr.iterations and abs(sa.dt - sr.dt) <= 1e-12 * sr.dt, ctx
        assert abs(sa.rms_error - sr.rms_error) <= 1e-6, ctx
        halvings.append(np.log2(cfl_now * sr.safe_dt / sr.dt))
    assert max(halvings) >= 0.99 and all(abs(h - round(h)) < 1e-6 for h in halvings), halvings
    pa, pr = a.params, r.params
    assert pa.cfl == pr.cfl and pa.n_stable == pr.n_stable and pa.n_unstable == pr.n_unstable
    for level in (0, 1):
        assert_same(a, r, INTS, 0.0, which, level)
        assert_same(a, r, FLOATS, 1e-9, which, level)


WALLS = [("ghost", 2, 0, None, None, {}), ("ghost_noslip", 2, 1, None, None, {}), ("adami_noslip", 1, 1, None, None, {}),
         ("adami_moving", 1, 0, [0.0, 0.004, 0.009], [[0.1, 0, 0], [0, 0.2, 0], [0, 0, 0]], {}),
         ("ghost_noslip_moving_rk4", 2, 1, [0.0, 0.004], [[0.1, 0, 0], [0, 0.2, 0]], dict(solver_type=1)),
         ("adami_moving_rk4", 1, 0, [0.0, 0.004], [[0.1, 0, 0], [0, 0.2, 0]], dict(solver_type=1))]


@pytest.mark.parametrize("name,solver,no_slip,times,vels,kw", WALLS, ids=[w[0] for w in WALLS])
def test_wall_treatments(name, solver, no_slip, times, vels, kw):
    """Boundary_Ghost, Set_No_Slip, Get_Boundary_Pressure and the wall-velocity schedules of Do_NB_Iter and of the RK
    stages (Resid.cpp:21-186, Newmark_Beta.cpp:69-132, Runge_Kutta.cpp:36-131 with its own comparator and its wall
    densities advanced BEFORE the stage's forces).  Boundary_DBC is not compared: it sizes its scratch vector by the end
    of the WALL block and writes it at FLUID indices (Resid.cpp:84-107) -- heap corruption in the reference."""
    tank = cases.box_with_walls(n=(8, 6, 7), jitter=0.05)
    nb, n = tank["bound_points"], tank["xi"].shape[0]
    a, r = pair(tank, "ref3d", ale=1, **kw)
    for o in (a, r):
        o.lib.orc_clear_blocks(o.h)
        o.add_block(0, 0, nb, bound_solver=solver, no_slip=no_slip, times=times, vels=vels)
        o.add_block(1, nb, n)
    for step in range(3):
        _, sa = a.integrate()
        _, sr = r.integrate()
        assert sa.iterations == sr.iterations and abs(sa.dt - sr.dt) <= 1e-12 * sr.dt, (name, step)
    for level in (0, 1):
        assert_same(a, r, INTS, 0.0, name, level)
        assert_same(a, r, FLOATS, 1e-9, name, level)


def _with_block(o, B):
    o.lib.orc_clear_blocks(o.h)
    o.add_block(1, B["first"], B["second"], block_type=6, fixed_vel_or_dynamic=B["fixed_vel_or_dynamic"],
                insert_norm=B["insert_norm"], insconst=B["insconst"], delete_norm=B.get("delete_norm"),
                delconst=B.get("delconst", 9999999.0), aero_norm=B["aero_norm"], aeroconst=B["aeroconst"], back=B["back"],
                buffer=B["buffer"])


@pytest.mark.parametrize("fixed", [0, 1])
def test_inlet_insertion_and_delete_plane(fixed):
    """update_buffer_region (inlet.cpp:578-640), the BACK / BUFFER motion of Do_NB_Iter, Check_Pipe_Outlet and the delete
    plane of update_data: same particles in the same order with the same ids, step after step."""
    case = cases.inlet_jet(delete_x=2.5, fixed=fixed, jitter=0.02)
    a, r = pair(case, "ref3d", ale=1)
    for o in (a, r):
        _with_block(o, case["block"])
    added = deleted = 0
    for step in range(14):
        _, sa = a.integrate()
        _, sr = r.integrate()
        assert (sa.iterations, sa.n_add, sa.n_del, a.n) == (sr.iterations, sr.n_add, sr.n_del, r.n), step
        added, deleted = added + sr.n_add, deleted + sr.n_del
        assert_same(a, r, ("part_id", "b"), 0.0, "inlet step %d" % step)
        assert_same(a, r, ("xi", "v", "rho"), 1e-11, "inlet step %d" % step)
    assert added > 0 and deleted > 0
    assert_same(a, r, INTS, 0.0, "inlet")
    assert_same(a, r, FLOATS, 1e-9, "inlet")


@pytest.mark.parametrize("which", ["sheared", "inner_wall"])
def test_mesh_containment(which):
    """FindCell / CheckCell / Crossings3D on a tri-fanned hexahedral mesh: same cells, same `internal` flags, same
    failure counters, same aero force.  Particles start in cell 0: the reference indexes cells.cFaces[cellID] before any
    search (Containment.cpp:592-600).  Erasures are not compared: FindCell lists an escaped particle once per
    crossed cell-centre ray (its `break` leaves only the face loop) and get_aero_velocity then erases by those duplicated
    indices (Resid.cpp:486-500) -- undefined in the reference; the oracle's contract is "once" (oracle header)."""
    case = cases.droplet(dx=0.0125, jitter=0.05)
    if which == "sheared":
        mesh = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (6, 7, 5), p=100000.0, rho=1.1025,
                              vel=lambda c: np.stack([5 + 20 * c[:, 1], 21.55 + 0 * c[:, 0], 3 * c[:, 2]], 1))
    else:
        mesh = cases.hex_mesh((-0.1013, -0.1007, -0.03), (0.1009, 0.1003, 0.1017), (6, 7, 5), vel=(0.0, 21.55, 0.0),
                              p=100000.0, rho=1.1025, outer_marker=-1)
    sims = []
    for kind in (None, "ref3d"):
        o = orc.Oracle(orc.default_params(3, ale=1, asource=1, **dict(case["params"], delta_t_min=1e-9)), kind=kind)
        o.set_mesh(mesh)
        o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
        for lvl in (0, 1):
            o.set("cellID", np.zeros(o.n, dtype=np.int64), lvl)
        sims.append(o)
    a, r = sims
    for step in range(3):
        _, sa = a.integrate()
        _, sr = r.integrate()
        assert (sa.iterations, a.n) == (sr.iterations, r.n), step
        assert_same(a, r, ("cellID", "internal", "ipt_n_failed"), 0.0, "mesh step %d" % step)
    assert (r.get("cellID") >= 0).sum() > 20 and np.abs(r.get("Af")).max() > 1.0
    if which == "inner_wall":
        assert r.get("internal").sum() > 0
    assert_same(a, r, INTS, 0.0, which)
    assert_same(a, r, FLOATS, 1e-9, which)


@pytest.mark.parametrize("which", ["sheared", "inner_wall"])
def test_mesh_containment_2d(which):
    """The 2D build's containment: FindCell / CheckCell on Crossings2D (Geometry.cpp:354-399), the boundary test on
    get_line_intersection (Geometry.cpp:312-341, with its one-sided denominator test), on a quadrilateral mesh whose faces
    are edges -- a 2D droplet in a sheared stream, and the same droplet cut by an inner wall.  Same cells, `internal` flags,
    failure counters and aero force as FJSPH's own -DSIMDIM=2 objects."""
    if not _have("ref2d"):
        pytest.skip("ref2d")
    case = cases.droplet(dx=0.004, dim=2, jitter=0.05)
    if which == "sheared":
        mesh = cases.quad_mesh((-0.1013, -0.1007), (0.1009, 0.1003), (6, 7), p=100000.0, rho=1.1025,
                               vel=lambda c: np.stack([21.55 + 0 * c[:, 0], 5 + 20 * c[:, 0]], 1))
    else:
        mesh = cases.quad_mesh((-0.1013, -0.03), (0.1009, 0.1003), (6, 5), vel=(21.55, 0.0), p=100000.0, rho=1.1025,
                               outer_marker=-1)
    sims = []
    for kind in ("2d", "ref2d"):
        o = orc.Oracle(orc.default_params(2, ale=1, asource=1, **dict(case["params"], delta_t_min=1e-9)), kind=kind)
        o.set_mesh(mesh)
        o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
        for lvl in (0, 1):
            o.set("cellID", np.zeros(o.n, dtype=np.int64), lvl)
        sims.append(o)
    a, r = sims
    for step in range(3):
        _, sa = a.integrate()
        _, sr = r.integrate()
        assert (sa.iterations, a.n) == (sr.iterations, r.n), step
        assert_same(a, r, ("cellID", "internal", "ipt_n_failed"), 0.0, "2D mesh step %d" % step)
    assert (r.get("cellID") >= 0).sum() > 20 and np.abs(r.get("Af")).max() > 1.0
    assert len(np.unique(r.get("cellID"))) > 4
    if which == "inner_wall":
        assert r.get("internal").sum() > 0
    assert_same(a, r, INTS, 0.0, "2D " + which)
    assert_same(a, r, FLOATS, 1e-9, "2D " + which)
