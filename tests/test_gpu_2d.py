"""SIMDIM = 2 on the device (reference src/VarDefs.h:29-41): the same kernels on records whose z components are exact
zeros, with the 2 x 2 forms of the per-particle Eigen algorithms (dSPH_PreStep's L inverse and lam, Shifting.cpp:73-99), the
2D free-surface test (Geometry.cpp:76-84), the 2D Wendland normalisation and Gissler areas (IO.cpp:87, Var.h:244-266,
Aero.h:57-84).  Checker: the oracle's 2D build (oracle/lib/liborc2d.so), which follows FJSPH's own -DSIMDIM=2 objects on
Dam_2D and the 2D decks (tests/test_oracle_vs_reference.py, tests/golden/ref_dam_2d.npz).  Bars as in the 3D suite."""
import os

import numpy as np
import pytest

from fjsph_b200 import cases, engine as eng, frontend
from tests.test_gpu_parity import PRESTEP_FIELDS, TOL_EIGEN_DEGENERATE, run_stages
from tests.util import TOL, assert_fields_close, make_pair, make_pair_from_deck, relerr

pytestmark = pytest.mark.gpu
DECKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks")


block2d = cases.synthetic_block_2d


def cases2d():
    yield "block", block2d()
    yield "block_eps", block2d(n=(23, 19), jitter="eps", seed=4)
    yield "droplet", cases.droplet(dx=0.002, dim=2, jitter=0.05)
    yield "tank", cases.box_with_walls(n=(24, 14), dx=0.01, layers=4, jitter=0.05, dim=2)


@pytest.mark.parametrize("name,case", list(cases2d()), ids=[n for n, _ in cases2d()])
def test_2d_neighbour_sets_bit_exact(name, case):
    o, e, p = make_pair(case, dim=2)
    o.update_neighbours()
    e.update_neighbours()
    off_o, idx_o, _ = o.neighbours()
    off_e, idx_e = e.neighbours()
    assert np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e)
    got = e.download(("xi", "v", "rho", "m", "b", "part_id"))
    assert got["xi"].shape == case["xi"].shape == (case["xi"].shape[0], 2)
    assert np.array_equal(got["xi"], case["xi"]) and np.array_equal(got["b"], case["b"])
    assert np.array_equal(got["part_id"], np.arange(case["xi"].shape[0]))


@pytest.mark.parametrize("ale", [1, 0])
@pytest.mark.parametrize("which", ["block", "droplet", "block_eps"])
def test_2d_stagewise_parity(which, ale):
    case = dict(cases2d())[which]
    o, e, p = make_pair(case, dim=2, ale=ale)
    npd_o, npd_e = run_stages(o, e, ale=bool(ale), label="2D " + which,
                              tol_eigen=TOL if which == "block" else TOL_EIGEN_DEGENERATE)
    L = e.download(("L",))["L"]
    assert L.shape == (case["xi"].shape[0], 2, 2)
    o.forces(npd_o)
    e.get_acc_and_Rrho(npd_e)
    assert_fields_close(e, o, ("acc", "Rrho", "Af"), context="2D " + which + " forces")
    assert np.abs(e.download(("acc",))["acc"]).max() > 0.0


def _steps(o, e, steps, ctx, state_tol=1e-10, rate_tol=1e-6):
    for step in range(steps):
        _, so = o.integrate()
        se = e.integrate()
        c = "%s step %d" % (ctx, step)
        assert se.iterations == so.iterations and se.total_points == so.total_points, c
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, c
        assert np.array_equal(e.neighbour_counts(), o.neighbour_counts()), c
        assert_fields_close(e, o, ("surf", "surfzone", "b", "part_id"), context=c)
        assert_fields_close(e, o, ("xi", "rho", "lam", "lam_nb"), tol=state_tol, context=c)
        assert_fields_close(e, o, ("v", "p"), tol=100 * state_tol, context=c)
        assert_fields_close(e, o, ("acc", "Rrho", "Af", "aVisc", "deltaD", "vPert"), tol=rate_tol, context=c)


@pytest.mark.parametrize("solver", [0, 1], ids=["newmark_beta", "rk4"])
@pytest.mark.parametrize("which", ["block", "droplet", "tank"])
def test_2d_full_step_parity(which, solver):
    """Three Integrator::integrate steps in 2D on generic positions: free block, droplet in a gas stream (Gissler with the
    2D areas, TAB deformation on), tank with Adami walls under gravity along -y."""
    case = dict(cases2d())[which]
    kw = dict(solver_type=solver, delta_t_min=1e-9)
    if which == "droplet":
        kw.update(use_TAB_def=1)
    o, e, p = make_pair(case, dim=2, **kw)
    _steps(o, e, 3, "2D %s solver %d" % (which, solver))


def test_dam_2d_deck():
    """BASELINE.json configs[0]: Examples/Dam_2D as shipped, through the para / bmap front end (tests/decks/dam2d.para: a water
    column beside a Pressure-Gradient wall on a Ghost floor).  The deck's lattice + U(0, eps dx) positions are a tie-stress
    input (tests/test_gpu_decks.py), but no 2D particle sits on a repeated eigenvalue the way the 3D lattice decks' edge
    particles do: flags, counts and dt exact, the state to the bars of the generic cases."""
    case = frontend.read_case(os.path.join(DECKS, "dam2d.para"), 2)
    assert case["dim"] == 2 and case["xi"].shape[1] == 2
    o, e = make_pair_from_deck(case)
    for step in range(6):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "dam 2D step %d" % step
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-9 * so.dt, ctx
        assert se.total_points == so.total_points
    assert_fields_close(e, o, ("surf", "surfzone", "b", "part_id"), context="dam 2D")
    got = e.download(("xi", "rho", "p", "v", "acc", "Rrho"))
    report = {f: relerr(got[f], o.get(f)) for f in got}
    assert report["xi"] <= 1e-10 and report["rho"] <= 1e-10 and report["v"] <= 1e-8 and report["p"] <= 1e-8, report
    assert report["acc"] <= 1e-6 and report["Rrho"] <= 1e-6, report


def pair_with_mesh_2d(case, mesh, **kw):
    from oracle import oracle as orc

    params = dict(case["params"], delta_t_min=1e-9, asource=1, **kw)
    o = orc.Oracle(orc.default_params(2, **params), kind="2d")
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(mesh)
    e = eng.Engine(eng.default_params(2, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    e.upload_mesh(mesh)
    return o, e


def test_2d_mesh_sheared_solution_against_oracle():
    """2D aero mesh (faces are edges): CheckCell on Crossings2D (Geometry.cpp:354-399), FindCell's nearest-centre search;
    the oracle's 2D containment follows FJSPH's -DSIMDIM=2 objects (tests/test_oracle_vs_reference.py::
    test_mesh_containment_2d).  Cells identical, then three coupled steps."""
    case = cases.droplet(dx=0.002, dim=2, jitter=0.05)
    vel = lambda c: np.stack([21.55 * (1.0 + 4.0 * c[:, 1]), 2.0 * c[:, 0]], axis=1)
    mesh = cases.quad_mesh((-0.1013, -0.1007), (0.1009, 0.1003), (11, 9), vel=vel, p=lambda c: 100000.0 + 500.0 * c[:, 1],
                           rho=lambda c: 1.1025 + 0.1 * c[:, 0])
    o, e = pair_with_mesh_2d(case, mesh)
    o.update_neighbours(); e.update_neighbours()
    o.prestep(); e.dSPH_PreStep()
    o.aero_velocity(); e.get_aero_velocity()
    got = e.download(("cellID", "cellV", "cellP", "cellRho"))
    assert np.array_equal(got["cellID"], o.get("cellID"))
    hit = got["cellID"] >= 0
    assert hit.sum() > 50 and len(np.unique(got["cellID"][hit])) > 8
    for f in ("cellV", "cellP", "cellRho"):
        assert np.array_equal(got[f][hit], o.get(f)[hit]), f
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt
        got = e.download(("cellID", "xi", "v", "rho", "Af", "acc"))
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["rho"], o.get("rho")) <= 1e-10
        assert relerr(got["v"], o.get("v")) <= 1e-8
        assert relerr(got["Af"], o.get("Af")) <= 1e-6 and relerr(got["acc"], o.get("acc")) <= 1e-6
    assert np.abs(got["Af"]).max() > 1.0


@pytest.mark.parametrize("marker", [-2, -1], ids=["outer_boundary", "inner_wall"])
def test_2d_mesh_boundaries(marker):
    """A 2D mesh that does not cover the droplet's lower part: the free-surface particles out there are tested against the
    boundary edges with get_line_intersection (Geometry.cpp:312-341, one-sided denominator test) -- erased across an outer
    boundary (-2), flagged `internal` across an inner wall (-1) -- as in the oracle."""
    case = cases.droplet(dx=0.002, dim=2, jitter=0.05)
    mesh = cases.quad_mesh((-0.1013, -0.0303), (0.1009, 0.1003), (8, 6), vel=(21.55, 0.0), p=100000.0, rho=1.1025,
                           outer_marker=marker)
    o, e = pair_with_mesh_2d(case, mesh)
    n0 = e.n
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert e.n == o.n, step
        got = e.download(("part_id", "cellID", "internal", "xi", "rho"))
        assert np.array_equal(got["part_id"], o.get("part_id")), step
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        assert np.array_equal(got["internal"], o.get("internal")), step
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["rho"], o.get("rho")) <= 1e-10
    if marker == -2:
        assert e.n < n0  # something was erased
    else:
        assert got["internal"].sum() > 0


def test_2d_pipe_outlet_takes_its_first_cell_from_the_mesh():
    """Check_Pipe_Outlet with a 2D mesh (Containment.cpp:822-847): a PIPE particle crossing the aero plane becomes FREE and
    FirstCell assigns its cell from the 20 nearest cell centres (150 in 3D)."""
    from oracle import oracle as orc

    case = cases.inlet_jet(n=(7, 4), fixed=1, jitter=0.03, aero_x=0.5, dim=2)
    mesh = cases.quad_mesh((-0.0123, -0.0031), (0.0117, 0.0093), (12, 6), vel=(0.0, 30.0), p=100000.0, rho=1.2)
    params = dict(case["params"], asource=1, acase=1, v_inf=(0.0, 30.0, 0.0), p_ref=100000.0, rho_g=1.2)
    B = case["block"]
    o = orc.Oracle(orc.default_params(2, **params), kind="2d")
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    o.lib.orc_clear_blocks(o.h)
    o.add_block(1, B["first"], B["second"], block_type=6, fixed_vel_or_dynamic=1, insert_norm=B["insert_norm"],
                insconst=B["insconst"], aero_norm=B["aero_norm"], aeroconst=B["aeroconst"], back=B["back"], buffer=B["buffer"])
    o.set_mesh(mesh)
    e = eng.Engine(eng.default_params(2, **params), 4 * case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    e.set_blocks([B])
    e.upload_mesh(mesh)
    freed = 0
    for step in range(6):
        _, so = o.integrate()
        se = e.integrate()
        assert (se.n_add, se.total_points, se.iterations) == (so.n_add, so.total_points, so.iterations), step
        got = e.download(("part_id", "b", "cellID", "xi", "v"))
        assert np.array_equal(got["part_id"], o.get("part_id")) and np.array_equal(got["b"], o.get("b")), step
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["v"], o.get("v")) <= 1e-8
        freed = int((got["b"] == cases.FREE).sum())
    assert freed > 0 and o.first_cell_errors == 0


def test_2d_mesh_must_be_flat():
    """A 2D engine takes 2D meshes only: a vertex off the z = 0 plane is an error, not a wrong answer."""
    case = block2d(n=(12, 10))
    e = eng.Engine(eng.default_params(2, **case["params"]), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    mesh = cases.hex_mesh((-1e-2, -1e-2, -1e-2), (3e-2, 3e-2, 1e-2), (2, 2, 1), vel=lambda c: np.zeros_like(c), p=1e5, rho=1.2)
    with pytest.raises(Exception):
        e.upload_mesh(mesh)


def test_c1_dam_2d_at_the_examples_numbers():
    """BASELINE.json configs[0] at size: tests/decks/dam2d_example.para states Examples/Dam_2D's own numbers (0.02 m spacing,
    three Pressure-Gradient walls, 8371 particles; identical to the reference's deck particle for particle,
    tests/test_frontend_cpu.py::test_dam2d_example_deck_is_the_references).  Ten steps against the 2D oracle: flags, counts
    and dt exact; state 1e-10, rates 1e-6 (measured: x 3e-17, rho 2e-16, rates 9e-12)."""
    case = frontend.read_case(os.path.join(DECKS, "dam2d_example.para"), 2)
    assert case["xi"].shape == (8371, 2) and case["bound_points"] == 3520
    o, e = make_pair_from_deck(case)
    for step in range(10):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "Dam_2D step %d" % step
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-9 * so.dt, ctx
        assert se.total_points == so.total_points
    assert_fields_close(e, o, ("surf", "surfzone", "b", "part_id"), context="Dam_2D")
    got = e.download(("xi", "rho", "p", "v", "acc", "Rrho"))
    report = {f: relerr(got[f], o.get(f)) for f in got}
    print("Dam_2D, 10 steps:", {k: "%.1e" % v for k, v in report.items()})
    assert report["xi"] <= 1e-10 and report["rho"] <= 1e-10 and report["v"] <= 1e-8 and report["p"] <= 1e-8, report
    assert report["acc"] <= 1e-6 and report["Rrho"] <= 1e-6, report  # measured 3e-12 / 9e-12 (profiles/r2n_tests_gpu_2d.log)
    # the column has started to collapse: the free surface is detected along its top and its right side
    surf = e.download(("surf", "b", "xi"))
    fluid = surf["b"] != cases.BOUND
    assert surf["surf"][fluid].sum() > 100


def test_2d_state_after_500_steps():
    """The 3D suite's long run (tests/test_gpu_parity.py::test_state_after_1000_steps) in 2D: 500 Integrator::integrate
    calls on the jittered 2D block with free surfaces on all four sides; sub-iteration count and dt of every step and every
    surface flag agree, positions to 1e-8 of the domain, density 1e-10, velocity 1e-7."""
    case = block2d(n=(24, 18), seed=5)
    o, e, p = make_pair(case, dim=2, delta_t_min=1e-9)
    for step in range(500):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations, step
        assert abs(se.dt - so.dt) <= 1e-9 * so.dt, step
    assert_fields_close(e, o, ("surf", "surfzone", "b"), context="2D 500 steps")
    got = e.download(("xi", "rho", "v"))
    report = {f: relerr(got[f], o.get(f)) for f in got}
    print("2D, 500 steps:", {k: "%.1e" % v for k, v in report.items()})
    assert report["xi"] <= 1e-8 and report["rho"] <= 1e-10 and report["v"] <= 1e-7, report
    assert np.abs(o.get("xi") - case["xi"]).max() > 2e-3  # the block really moved


def test_2d_coupled_to_a_tau_edge_mesh(tmp_path):
    """A 2D droplet coupled to a TAU edge-based mesh and its two-layer solution read by fjsph_tau_read_edge (the 2D build's
    TAU::Read_tau_mesh_EDGE + Read_SOLUTION; pinned against the compiled reference in tests/test_frontend_vs_reference.py):
    the engine on the read mesh follows the 2D oracle on the same mesh -- cells identical, the cell data on the particles."""
    from tests.tau_case import write_tau_edge

    vel = lambda x: (21.55 * (1.0 + 2.0 * x[1]), 1.5 * x[0])
    mesh, sol, *_ = write_tau_edge(tmp_path, (-0.1013, -0.1007), (0.1009, 0.1003), (9, 8), vel,
                                   lambda x: 1.0e5 + 300.0 * x[1], lambda x: 1.1025 + 0.1 * x[0], plane="xz")
    tau = frontend.read_tau_edge(mesh, sol, offset_axis=2)
    case = cases.droplet(dx=0.002, dim=2, jitter=0.05)
    o, e = pair_with_mesh_2d(case, tau)
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt
        got = e.download(("cellID", "cellV", "cellP", "xi", "rho", "Af", "acc"))
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        hit = got["cellID"] >= 0
        assert np.array_equal(got["cellV"][hit], o.get("cellV")[hit]) and np.array_equal(got["cellP"][hit], o.get("cellP")[hit])
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["rho"], o.get("rho")) <= 1e-10
        assert relerr(got["Af"], o.get("Af")) <= 1e-6 and relerr(got["acc"], o.get("acc")) <= 1e-6
    assert hit.sum() > 50 and len(np.unique(got["cellID"][hit])) > 8 and np.abs(got["Af"]).max() > 1.0
