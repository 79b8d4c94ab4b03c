"""Aero-mesh containment on the device (FindCell / CheckCell / Crossings3D / FirstCell, Containment.cpp, Geometry.cpp)
against the CPU oracle, plus the built-in check of SURVEY 8d C4: a uniform mesh solution must reproduce the
constant-freestream run exactly."""
import numpy as np
import pytest

from fjsph_b200 import cases, engine as eng
from oracle import oracle as orc
from tests.util import relerr

pytestmark = pytest.mark.gpu
V_INF = (0.0, 21.55, 0.0)


def pair_with_mesh(case, mesh, **kw):
    params = dict(case["params"], delta_t_min=1e-9, asource=1, **kw)
    o = orc.Oracle(orc.default_params(3, **params))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(mesh)
    e = eng.Engine(eng.default_params(3, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    e.upload_mesh(mesh)
    return o, e


def test_uniform_mesh_reproduces_constant_freestream():
    case = cases.droplet(dx=0.01, jitter=0.05)
    mesh = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (8, 9, 7), vel=V_INF, p=100000.0, rho=1.1025)
    params = dict(case["params"], delta_t_min=1e-9)
    ec = eng.Engine(eng.default_params(3, **params), case["xi"].shape[0])
    ec.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    em = eng.Engine(eng.default_params(3, asource=1, **params), case["xi"].shape[0])
    em.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    em.upload_mesh(mesh)
    for step in range(3):
        sc, sm = ec.integrate(), em.integrate()
        assert sc.iterations == sm.iterations and sc.dt == sm.dt
        a, b = ec.download(("xi", "v", "rho", "acc", "Af")), em.download(("xi", "v", "rho", "acc", "Af", "cellID"))
        for f in ("xi", "v", "rho", "acc", "Af"):
            assert np.array_equal(a[f], b[f]), (step, f)
        assert (b["cellID"] >= 0).sum() > 50 and np.abs(b["Af"]).max() > 1.0   # the aero force is really on


def test_sheared_mesh_solution_against_oracle():
    case = cases.droplet(dx=0.01, jitter=0.05)
    vel = lambda c: np.stack([2.0 * c[:, 2], 21.55 * (1.0 + 4.0 * c[:, 0]), -3.0 * c[:, 1]], axis=1)
    mesh = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (11, 9, 10), vel=vel,
                          p=lambda c: 100000.0 + 500.0 * c[:, 1], rho=lambda c: 1.1025 + 0.1 * c[:, 2])
    o, e = pair_with_mesh(case, mesh)
    # stage by stage first: neighbours, prestep, containment
    o.update_neighbours(); e.update_neighbours()
    o.prestep(); e.dSPH_PreStep()
    o.aero_velocity(); e.get_aero_velocity()
    got = e.download(("cellID", "cellV", "cellP", "cellRho"))
    assert np.array_equal(got["cellID"], o.get("cellID"))
    hit = got["cellID"] >= 0
    assert hit.sum() > 50
    for f in ("cellV", "cellP", "cellRho"):
        assert np.array_equal(got[f][hit], o.get(f)[hit]), f
    for step in range(3):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations and abs(se.dt - so.dt) <= 1e-12 * so.dt
        got = e.download(("cellID", "xi", "v", "rho", "Af", "acc", "woccl"))
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["rho"], o.get("rho")) <= 1e-10
        assert relerr(got["v"], o.get("v")) <= 1e-8
        assert relerr(got["Af"], o.get("Af")) <= 1e-6 and relerr(got["acc"], o.get("acc")) <= 1e-6


def test_particles_outside_the_mesh_are_erased_like_the_oracle():
    """A mesh that does not cover the droplet's +y cap: the free-surface particles out there cross an outer boundary
    face (marker -2) and are erased from both time levels; the survivors keep the reference's order."""
    case = cases.droplet(dx=0.01, jitter=0.05)
    mesh = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.0303, 0.1017), (8, 6, 7), vel=V_INF, p=100000.0, rho=1.1025)
    o, e = pair_with_mesh(case, mesh)
    n0 = e.n
    for step in range(2):
        _, so = o.integrate()
        se = e.integrate()
        assert e.n == o.n, step
        got = e.download(("part_id", "cellID", "internal", "xi", "rho"))
        assert np.array_equal(got["part_id"], o.get("part_id")), step
        assert np.array_equal(got["cellID"], o.get("cellID")), step
        assert relerr(got["xi"], o.get("xi")) <= 1e-10 and relerr(got["rho"], o.get("rho")) <= 1e-10
    assert e.n < n0   # something was erased


def test_pipe_outlet_takes_its_first_cell_from_the_mesh():
    """Check_Pipe_Outlet with a mesh (Containment.cpp:822-847): a PIPE particle crossing the aero plane becomes FREE and
    FirstCell assigns its cell from the 150 nearest cell centres."""
    case = cases.inlet_jet(n=(5, 5, 4), fixed=1, jitter=0.03, aero_x=0.5)
    mesh = cases.hex_mesh((-0.0123, -0.0031, -0.0029), (0.0117, 0.0073, 0.0071), (12, 5, 5), vel=(0.0, 30.0, 0.0), p=100000.0,
                          rho=1.2)
    params = dict(case["params"], asource=1, acase=1, v_inf=(0.0, 30.0, 0.0), p_ref=100000.0, rho_g=1.2)
    B = case["block"]
    o = orc.Oracle(orc.default_params(3, **params))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    o.lib.orc_clear_blocks(o.h)
    o.add_block(1, B["first"], B["second"], block_type=6, fixed_vel_or_dynamic=1, insert_norm=B["insert_norm"],
                insconst=B["insconst"], aero_norm=B["aero_norm"], aeroconst=B["aeroconst"], back=B["back"], buffer=B["buffer"])
    o.set_mesh(mesh)
    e = eng.Engine(eng.default_params(3, **params), 4 * case["xi"].shape[0])
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


def test_tau_mesh_with_quadrilateral_faces(tmp_path):
    """A TAU face-based mesh + solution (NetCDF-3 classic, written with scipy, read by fjsph_tau_read without the NetCDF
    library): triangles AND four-cornered faces, which Crossings3D tests the reference's way, a sheared
    point solution averaged to the cells.  Two coupled steps against the oracle: cells identical, state 1e-10,
    velocity 1e-8, rates 1e-6 (measured on the box: x 1e-16, v 1e-12, Af 1e-15, profiles/r27_tau_probe.txt)."""
    from fjsph_b200 import frontend
    from tests.tau_case import write_tau

    lo, hi = np.array([-0.1013, -0.1007, -0.1011]), np.array([0.1009, 0.1003, 0.1017])
    mesh_file, sol_file, *_ = write_tau(tmp_path, lo, hi, (8, 9, 7), lambda x: (1.0 + 40 * x[0], 21.55 - 30 * x[2], 5 * x[1]),
                                        lambda x: 1.0e5 + 100 * x[1], lambda x: 1.1 + x[2])
    tau = frontend.read_tau(mesh_file, sol_file)
    assert set(np.diff(tau["face_ptr"])) == {3, 4}
    case = cases.droplet(dx=0.01, jitter=0.05)
    params = dict(case["params"], delta_t_min=1e-9)
    o = orc.Oracle(orc.default_params(3, asource=1, **params))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(tau)
    e = eng.Engine(eng.default_params(3, asource=1, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    e.upload_mesh(tau)
    for step in range(2):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations, step
        got = e.download(("cellID", "xi", "v", "rho", "Af", "acc"))
        assert np.array_equal(got["cellID"], o.get("cellID")) and (got["cellID"] >= 0).sum() > 50, step
        for f, tol in (("xi", 1e-10), ("rho", 1e-10), ("v", 1e-8), ("Af", 1e-6), ("acc", 1e-6)):
            ref = o.get(f)
            assert np.abs(got[f] - ref).max() <= tol * max(np.abs(ref).max(), 1e-300), (step, f)


def test_q6_crossed_quadrilaterals_keep_the_cell_below(tmp_path):
    """GPU twin of tests/test_tau_cpu.py::test_coupled_run_on_quadrilateral_faces_follows_the_reference: x-normal
    quadrilaterals, every particle started in the cell below its own; the three-edge crossing test keeps it there.  The
    engine must report the oracle's cells (which are FJSPH's, pinned on the CPU) and its state.  (Green on a B200 since
    round 1, GPUTEST_r01.json.)"""
    from fjsph_b200 import frontend
    from tests.tau_case import write_tau

    lo, hi, n = np.array([-0.1013, -0.1007, -0.1011]), np.array([0.1009, 0.1003, 0.1017]), (8, 9, 7)
    mesh_file, sol_file, *_ = write_tau(tmp_path, lo, hi, n, lambda x: (1.0 + 40 * x[0], 21.55 - 30 * x[2], 5 * x[1]),
                                        lambda x: 1.0e5 + 100 * x[1], lambda x: 1.1 + x[2], split="y")
    tau = frontend.read_tau(mesh_file, sol_file)
    case = cases.droplet(dx=0.0125, jitter=0.05)
    params = dict(case["params"], delta_t_min=1e-9)
    ijk = np.floor((case["xi"] - lo) / ((hi - lo) / np.array(n))).astype(int)
    own = (ijk[:, 2] * n[1] + ijk[:, 1]) * n[0] + ijk[:, 0]
    below = np.where(ijk[:, 2] > 0, own - n[0] * n[1], own).astype(np.int64)
    o = orc.Oracle(orc.default_params(3, asource=1, **params))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(tau)
    e = eng.Engine(eng.default_params(3, asource=1, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    e.upload_mesh(tau)
    for lvl in (0, 1):
        o.set("cellID", below, lvl)
        e.upload_level(lvl, cellID=below)
    for step in range(2):
        _, so = o.integrate()
        se = e.integrate()
        assert se.iterations == so.iterations, step
    got = e.download(("cellID", "xi", "v", "rho", "Af"))
    ref = o.get("cellID")
    assert np.array_equal(got["cellID"], ref)
    found = ref >= 0
    assert (ref[found] == below[found]).sum() > 0.9 * found.sum()
    for f, tol in (("xi", 1e-10), ("rho", 1e-10), ("v", 1e-8), ("Af", 1e-6)):
        r = o.get(f)
        assert np.abs(got[f] - r).max() <= tol * max(np.abs(r).max(), 1e-300), f
