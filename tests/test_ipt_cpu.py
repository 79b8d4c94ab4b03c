"""The implicit particle tracker (IPT::Integrate on the particles erased at a delete plane) on the CPU side:
the oracle restatement (oracle/ipt_oracle.inc) against FJSPH's own IPT.cpp / Containment.cpp / Geometry.cpp compiled in
oracle/_ref -- live where the reference is mounted, and through the committed vectors of tests/golden/ipt_*.npz everywhere
-- and the host-side settings functions of the product (fjsph_ipt_default_settings, fjsph_read_para_ipt,
fjsph_mesh_max_length).  The device tracker is compared with the oracle in tests/test_gpu_ipt.py."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

from fjsph_b200 import _lib, cases, engine as eng
from oracle import oracle as orc
from tests import ipt_case

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ipt_*.npz")))
FIELDS = ipt_case.INT_FIELDS + ipt_case.FLOAT_FIELDS


def run_oracle(dim, mesh, settings, start, kind=None, particle_step=1e-3, record_cap=ipt_case.RECORD_CAP):
    p = orc.default_params(dim, asource=1, particle_step=particle_step)
    o = orc.Oracle(p, kind=kind if kind else ("2d" if dim == 2 else None))
    o.set_mesh(mesh)
    return o.ipt_integrate(orc.ipt_settings(p, **settings), start, record_cap=record_cap)


def assert_same_tracks(a, b, what):
    """identical: every branch the tracker takes hangs on these numbers, and both sides evaluate the same expressions in the
    same order without FMA contraction"""
    assert (a["n_success"], a["n_failed"]) == (b["n_success"], b["n_failed"]), what
    assert np.array_equal(a["n_records"], b["n_records"]), what
    for f in FIELDS:
        assert np.array_equal(a["last"][f], b["last"][f]), (what, "last", f)
        assert np.array_equal(a["records"][f], b["records"][f]), (what, "records", f)


def test_fixtures_present():
    assert len(FIXTURES) == len(ipt_case.CASES), "tests/golden/ipt_*.npz missing: run tests/golden/make_ipt_vectors.py"


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_oracle_reproduces_the_reference_s_tracks(path):
    """tests/golden/ipt_*.npz hold what IPT::Integrate of the compiled reference made of each case: the restatement lands on
    the same cells, faces, outcomes and records, bit for bit -- including the particles the reference fails because
    MollerTrumbore accepts only half of a parallelogram face (note Q9) and the extra time step of note Q10."""
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    mesh = {k[5:]: z[k] for k in z.files if k.startswith("mesh_")}
    got = run_oracle(meta["dim"], mesh, meta["settings"], z["start"], particle_step=meta["particle_step"], record_cap=meta["record_cap"])
    ref = dict(last=z["last"], records=z["records"], n_records=z["n_records"], n_success=meta["n_success"], n_failed=meta["n_failed"])
    assert_same_tracks(got, ref, os.path.basename(path))
    assert got["n_steps"].max() >= 6 and (got["n_steps"] >= 1).all()
    # the record of a successful particle: start, one entry per completed step, the final state once more (Terminate_Particle)
    ok = (got["last"]["failed"] == 0)
    assert np.array_equal(got["n_records"][ok], got["n_steps"][ok] + 1)
    assert np.array_equal(got["n_records"][~ok], got["n_steps"][~ok])


@pytest.mark.parametrize("name", list(ipt_case.CASES))
def test_oracle_follows_the_compiled_reference(name):
    """live: the same inputs through oracle/_ref (skipped where /root/reference was never mounted), without records too"""
    case = ipt_case.build(name, n=120, seed=11)
    dim = case["dim"]
    kind = "ref2d" if dim == 2 else "ref3d"
    if not orc.have_ref(kind):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    p = orc.default_params(dim, asource=1, particle_step=case["particle_step"])
    start = ipt_case.start_records(case, orc.IPT_START, p.sim_mass)
    for record in (1, 0):
        settings = dict(case["settings"], record=record, max_length=case["length_factor"] * ipt_case.longest_edge(case["mesh"], dim))
        a = run_oracle(dim, case["mesh"], settings, start)
        b = run_oracle(dim, case["mesh"], settings, start, kind=kind)
        assert_same_tracks(a, b, (name, record))
        if not record:   # without streaks the record holds the start and, for a success, the end (IPT.cpp:881, 742-743)
            assert np.array_equal(a["n_records"], 1 + (a["last"]["failed"] == 0))


@pytest.mark.parametrize("diam", [6.0e-6, 4.0e-5, 9.9e-5, 1.0e-4])
@pytest.mark.parametrize("name", ["tet_o2_scatter", "hex_quad_o1_lane", "quad2d_o2_long"])
def test_small_droplets_start_every_step_from_the_gas_velocity(name, diam):
    """IPT.cpp:995-1003: below 10 microns a droplet starts every step after the first with the cell's gas velocity, up to 100
    microns with a blend of the two, above with its own.  The four diameters sit in, and on the edges of, the three regimes."""
    case = ipt_case.build(name, n=80, seed=31)
    dim = case["dim"]
    kind = "ref2d" if dim == 2 else "ref3d"
    if not orc.have_ref(kind):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    p = orc.default_params(dim, asource=1, particle_step=case["particle_step"])
    start = ipt_case.start_records(case, orc.IPT_START, p.sim_mass)
    settings = dict(case["settings"], diam=diam, area=np.pi * diam * diam / 4.0,
                    max_length=case["length_factor"] * ipt_case.longest_edge(case["mesh"], dim))
    a = run_oracle(dim, case["mesh"], settings, start)
    b = run_oracle(dim, case["mesh"], settings, start, kind=kind)
    assert_same_tracks(a, b, (name, diam))
    assert a["n_steps"].max() >= 5
    if diam < 1.0e-5:   # every record after the first step's carries the velocity the sub-iterations left, near the gas's
        moving = a["n_records"] > 3
        assert moving.any()


def test_outcomes_and_the_product_s_own_bounds():
    """What ends a track: the outer boundary (cellID = the marker), max_x, the speed and step-length bounds, a start outside
    the mesh, and max_steps (the reference loops unbounded: note Q11)."""
    mesh = cases.quad_mesh([-0.1, -0.1], [0.5, 0.1], (12, 5), vel=(40.0, 0.0), rho=1.2)
    p = orc.default_params(2, asource=1, particle_step=1e-3)
    start = np.zeros(4, dtype=orc.IPT_START)
    start["part_id"] = np.arange(4)
    start["xi"][:, :2] = [(-0.08, 0.013), (-0.08, 0.013), (-0.08, 0.013), (0.7, 0.0)]
    start["cellID"] = [0 + 12 * 2, 0 + 12 * 2, 0 + 12 * 2, -3]
    start["v"][:, 0] = [10.0, 10.0, 2000.0, 10.0]
    start["mass"] = p.sim_mass
    start["cellV"][:, 0] = 40.0
    start["cellRho"] = 1.2
    base = dict(eq_order=1, max_length=1.0, grav=[0.0, 0.0, 0.0])
    out = run_oracle(2, mesh, dict(base, max_x=0.2), start)
    last = out["last"]
    assert list(last["failed"]) == [0, 0, 1, 1] and (last["going"] == 0).all()
    assert last["xi"][0, 0] > 0.2 and last["cellID"][0] >= 0                   # stopped by max_x, still inside the mesh
    assert (out["n_success"], out["n_failed"]) == (2, 2)
    assert np.array_equal(last["xi"][3], start["xi"][3]) and out["n_steps"][3] == 1   # outside the mesh: no face, failed at once
    out = run_oracle(2, mesh, dict(base, max_x=9.0), start[:1])
    assert out["last"]["cellID"][0] == -2 and out["last"]["failed"][0] == 0 and out["n_steps"][0] == 12   # left through the outer boundary
    assert abs(out["last"]["xi"][0, 0] - 0.5) < 1e-6 and abs(out["last"]["xi"][0, 1] - 0.013) < 1e-12
    out = run_oracle(2, mesh, dict(base, max_x=9.0, max_steps=5), start[:1])
    assert out["last"]["failed"][0] == 2 and out["n_steps"][0] == 5 and out["n_failed"] == 1
    out = run_oracle(2, mesh, dict(base, max_x=9.0, max_length=0.049), start[:1])   # one cell is 0.05 long
    assert out["last"]["failed"][0] == 1


def test_ipt_settings_of_the_host(tmp_path):
    """fjsph_ipt_default_settings: IPT_SETT's defaults and ipt_diam / ipt_area as Set_Values derives them (IO.cpp:126-127);
    fjsph_read_para_ipt: the keys of IO.cpp:447-453, max_x scaled (IO.cpp:29), no tracking when max_x lies upstream of the SPH
    conversion coordinate (IO.cpp:674-679), an equation order other than 1 or 2 is an error (the reference exits)."""
    p = eng.default_params(3, particle_step=2e-3, rho_rest=800.0, mu_g=1.8e-5, max_subits=17, grav=(0.0, -3.0, -9.0))
    s, use = eng.ipt_settings(p)
    assert (s.eq_order, s.max_subits, s.record, use) == (2, 17, 1, 1) and (s.relax, s.n_relax, s.max_x) == (0.6, 5.0, 9999999.0)
    diam = ((6.0 * p.sim_mass) / (np.pi * 800.0)) ** (1.0 / 3.0)
    assert abs(s.diam - diam) <= 2e-16 * diam and abs(s.area - np.pi * diam * diam / 4.0) <= 1e-15 * s.area
    assert list(s.grav) == [0.0, -3.0, -9.0] and (s.mu_g, s.rho_rest) == (1.8e-5, 800.0)
    o = orc.ipt_settings(orc.default_params(3, particle_step=2e-3, rho_rest=800.0, mu_g=1.8e-5, max_subits=17))
    assert abs(o.diam - s.diam) <= 2e-16 * diam   # the oracle's helper and the product's agree
    para = tmp_path / "para"
    para.write_text(" Transition to IPT (0/1): 1\n Velocity equation order (1/2): 1\n SPH tracking conversion x coordinate: 0.3\n"
                    " Maximum x trajectory coordinate: 1.5   # metres of the unscaled grid\n Particle streak output (0/1/2): 0\n")
    s, use = eng.ipt_settings(p, para=para, scale=0.5)
    assert (use, s.eq_order, s.record) == (1, 1, 0) and s.max_x == 0.75
    para.write_text(" Transition to IPT (0/1): 1\n Particle streak output (0/1/2): 0\n Particle cell intersection output (0/1/2): 1\n"
                    " SPH tracking conversion x coordinate: 2\n Maximum x trajectory coordinate: 1.5\n")
    s, use = eng.ipt_settings(p, para=para)
    assert (use, s.record) == (0, 1)
    para.write_text(" Transition to IPT (0/1): 1\n Velocity equation order (1/2): 3\n")
    with pytest.raises(_lib.FjsphError, match="Equation order not 1 or 2"):
        eng.ipt_settings(p, para=para)
    para.write_text(" Transition to IPT (0/1): 0\n Velocity equation order (1/2): 3\n")   # only checked when tracking is on
    assert eng.ipt_settings(p, para=para)[1] == 0
    with pytest.raises(_lib.FjsphError, match="could not open"):
        eng.ipt_settings(p, para=tmp_path / "absent")
    # through the case front end: a deck that names an aero mesh and switches the tracker on
    from fjsph_b200 import frontend

    (tmp_path / "f.bmap").write_text("   Name: W\n  Shape: Sphere\n Centre coordinate: 0,0,0\n Radius: 0.03\n Particle spacing: 0.01\n block end\n")
    (tmp_path / "none.bmap").write_text("\n")
    deck = (" Input boundary definition filename: %s\n Input fluid definition filename: %s\n SPH initial spacing: 0.01\n"
            " SPH frame time interval: 1\n Transition to IPT (0/1): 1\n Velocity equation order (1/2): 1\n Grid scale: 0.5\n"
            " SPH aerodynamic case: Gissler\n"
            " SPH tracking conversion x coordinate: 0.2\n Maximum x trajectory coordinate: 3\n" % (tmp_path / "none.bmap", tmp_path / "f.bmap"))
    para.write_text(deck)
    s, use = frontend.read_case(str(para), 3)["ipt"]
    assert use == 0 and (s.eq_order, s.max_x) == (1, 1.5)            # constant free stream: nothing to track through (Integration.cpp:151)
    para.write_text(deck + " OpenFOAM input directory: foam\n OpenFOAM solution directory: 100\n")
    c = frontend.read_case(str(para), 3)
    s, use = c["ipt"]
    assert c["params"].asource == 1 and use == 1 and s.max_x == 1.5 and s.diam > 0.0   # max_x *= scale whatever the mesh (IO.cpp:29)


def test_mesh_max_length():
    """cells.maxlength as the TAU readers leave it: 5 x the longest edge of a triangle / longer diagonal of a quadrilateral
    (CDFIO.cpp:1117-1183, 1214), 4 x the longest edge in 2D (CDFIO.cpp:867-898, 931).  Pinned against the compiled readers in
    tests/test_frontend_vs_reference.py; here the arithmetic, and that the tests' own tight bound is the un-multiplied length."""
    lo, hi, n = [0.0, 0.0, 0.0], [1.0, 2.0, 3.0], (2, 2, 2)
    diag = np.sqrt(1.0 + 1.5 ** 2)
    for mesh in (cases.hex_mesh(lo, hi, n, triangulate=False), cases.hex_mesh(lo, hi, n)):   # the diagonal is an edge of both halves
        assert eng.mesh_max_length(mesh) == 5.0 * diag and ipt_case.longest_edge(mesh, 3) == diag
    quad = cases.quad_mesh(lo[:2], hi[:2], n[:2])
    assert eng.mesh_max_length(quad, 2) == 4.0 and ipt_case.longest_edge(quad, 2) == 1.0
    tets, _ = cases.tet_mesh(lo, hi, n)
    assert eng.mesh_max_length(tets) == 5.0 * ipt_case.longest_edge(tets, 3) == 5.0 * np.sqrt(0.25 + 1.0 + 2.25)
    bad = cases.quad_mesh(lo[:2], hi[:2], n[:2])
    bad["face_vtx"] = bad["face_vtx"].copy()
    bad["face_vtx"][3] = 999
    with pytest.raises(_lib.FjsphError, match="vertex index out of range"):
        eng.mesh_max_length(bad, 2)


def test_record_types_match_the_header():
    assert eng.IPT_START.itemsize == C.sizeof(_lib.FjsphDeleted) == orc.IPT_START.itemsize
    assert eng.IPT_POINT.itemsize == C.sizeof(_lib.FjsphIptPoint) == orc.IPT_POINT.itemsize
    for name in ("eq_order", "max_subits", "record", "max_steps", "relax", "n_relax", "max_x", "max_length", "diam", "area",
                 "grav", "mu_g", "rho_rest"):
        assert getattr(_lib.FjsphIptSettings, name).offset == getattr(orc.OrcIptSettings, name).offset


def test_tracks_on_the_reference_s_own_rae2822_mesh():
    """Examples/RAE2822 (the one mesh the reference ships: 23 552 cells around an aerofoil, its 2D edge-based TAU layout read by
    fjsph_tau_read_edge) with the flow solution beside it: 300 droplets started at cell centres upstream of and around the
    aerofoil at 30 % of the local gas velocity.  The restatement and the compiled reference follow every one of them through
    up to ~230 cells to the same end -- the aerofoil's wall (marker -1) or the end plane at x = 1.5 -- bit for bit."""
    from fjsph_b200 import frontend

    rae = "/root/reference/Examples/RAE2822"
    if not os.path.exists(rae + "/mesh.grid.conf.edges"):
        pytest.skip("the reference's Examples are not mounted here")
    if not orc.have_ref("ref2d"):
        pytest.skip("oracle/_ref not built")
    mesh = frontend.read_tau_edge(rae + "/mesh.grid.conf.edges", rae + "/sol.pval.10000", scale=1.0, offset_axis=2)
    c = mesh["cCentre"]
    near = np.where((c[:, 0] > -0.5) & (c[:, 0] < 0.2) & (np.abs(c[:, 1]) < 0.4))[0]
    pick = np.random.default_rng(1).choice(near, 300, replace=False)
    p = orc.default_params(2, asource=1, particle_step=1e-4)
    start = np.zeros(300, dtype=orc.IPT_START)
    start["part_id"], start["cellID"], start["mass"] = np.arange(300), pick, p.sim_mass
    start["xi"][:, :2], start["v"][:, :2] = c[pick], 0.3 * mesh["cVel"][pick]
    start["cellV"][:, :2], start["cellRho"] = mesh["cVel"][pick], mesh["cRho"][pick]
    # (the shipped points_zc ends in 1354 NetCDF fill values, so cells.maxlength is 1e37 here: no bound on a step)
    settings = dict(eq_order=2, max_x=1.5, max_length=eng.mesh_max_length(mesh, 2), grav=[0.0, -9.81, 0.0], max_steps=5000)
    a = run_oracle(2, mesh, settings, start, particle_step=1e-4, record_cap=8)
    b = run_oracle(2, mesh, settings, start, kind="ref2d", particle_step=1e-4, record_cap=8)
    assert_same_tracks(a, b, "RAE2822")
    last = a["last"]
    assert a["n_failed"] == 0 and a["n_steps"].max() > 150
    hit_wall, passed = last["cellID"] == -1, last["xi"][:, 0] > 1.5
    assert hit_wall.sum() > 20 and passed.sum() > 200 and (hit_wall | passed).all()


def test_tracks_on_a_tau_file_read_by_either_side(tmp_path):
    """The whole chain as the reference runs it: a TAU face mesh (triangles and quadrilaterals) and its solution in NetCDF files,
    read by TAU::Read_tau_mesh_FACE + Read_SOLUTION on one side and by fjsph_tau_read on the other; the bound on one step is
    each reader's own cells.maxlength (5 x the longest face diagonal).  IPT::Integrate on the reference's MESH and the restatement
    on the product's: the same tracks, bit for bit."""
    from fjsph_b200 import frontend
    from tests.tau_case import write_tau

    if not orc.have_ref("ref3d"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    lo, hi, n = np.array([-0.1, -0.1, -0.1]), np.array([0.5, 0.1, 0.1]), (12, 5, 4)
    vel = lambda x: (30 + 40 * x[0] + 20 * x[2], 0.4 * np.sin(9 * x[0]) - 0.3, -0.5 + 1.5 * x[1])
    mesh_file, sol_file, *_ = write_tau(tmp_path, lo, hi, n, vel, lambda x: 1.0e5, lambda x: 1.1 + x[2] + 0.3 * x[0], split="y")
    mine = frontend.read_tau(mesh_file, sol_file)
    assert set(np.diff(mine["face_ptr"])) == {3, 4}
    p = orc.default_params(3, asource=1, particle_step=1e-3)
    ref = orc.Oracle(p, kind="ref3d")
    theirs = orc.ref_read_tau(ref, mesh_file, sol_file, 1.0)
    assert np.array_equal(mine["cVel"], theirs["cVel"])
    rng = np.random.default_rng(41)
    k = 200
    cells = rng.integers(0, mine["cCentre"].shape[0] // 3, size=k)          # the upstream third
    start = np.zeros(k, dtype=orc.IPT_START)
    start["part_id"], start["cellID"], start["t"], start["mass"] = np.arange(k), cells, 0.1, p.sim_mass
    start["xi"] = mine["cCentre"][cells] + rng.uniform(-0.2, 0.2, size=(k, 3)) * (hi - lo) / np.array(n)
    start["v"] = rng.normal(scale=1.0, size=(k, 3)) + np.array([10.0, 0.0, 0.0])
    start["cellV"], start["cellRho"] = mine["cVel"][cells], mine["cRho"][cells]
    settings = orc.ipt_settings(p, eq_order=2, max_x=0.45, max_length=eng.mesh_max_length(mine), max_steps=4000)
    theirs_settings = orc.ipt_settings(p, eq_order=2, max_x=0.45, max_length=-1.0, max_steps=4000)   # -1: the reader's own maxlength
    b = ref.ipt_integrate(theirs_settings, start, record_cap=40)              # on the MESH the reference read itself
    o = orc.Oracle(p)
    o.set_mesh(mine)
    a = o.ipt_integrate(settings, start, record_cap=40)
    assert_same_tracks(a, b, "tau file")
    assert a["n_success"] > 5 and a["n_steps"].max() >= 8


def test_tracks_on_a_tau_edge_file_read_by_either_side(tmp_path):
    """The 2D build's chain: an edge-based TAU mesh and its two-layer solution read by TAU::Read_tau_mesh_EDGE on one side and
    by fjsph_tau_read_edge on the other, each side's own cells.maxlength (4 x the longest edge), then the tracker."""
    from fjsph_b200 import frontend
    from tests.tau_case import write_tau_edge

    if not orc.have_ref("ref2d"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    lo, hi, n = np.array([-0.1, -0.1]), np.array([0.5, 0.1]), (12, 5)
    vel = lambda x: (30 + 40 * x[0] + 20 * x[1], 5 * np.sin(9 * x[0]) + 3 * x[1])
    mesh_file, sol_file, *_ = write_tau_edge(tmp_path, lo, hi, n, vel, lambda x: 1.0e5, lambda x: 1.1 + x[1] + 0.3 * x[0], plane="xz")
    mine = frontend.read_tau_edge(mesh_file, sol_file, scale=1.0, offset_axis=2)
    p = orc.default_params(2, asource=1, particle_step=1e-3)
    ref = orc.Oracle(p, kind="ref2d")
    theirs = orc.ref_read_tau_edge(ref, mesh_file, sol_file, 1.0, 2)
    assert np.array_equal(mine["cVel"], theirs["cVel"])
    rng = np.random.default_rng(43)
    k = 200
    cells = rng.integers(0, mine["cCentre"].shape[0], size=k)
    cells = cells[mine["cCentre"][cells, 0] < 0.1][:120]
    k = len(cells)
    start = np.zeros(k, dtype=orc.IPT_START)
    start["part_id"], start["cellID"], start["t"], start["mass"] = np.arange(k), cells, 0.1, p.sim_mass
    start["xi"][:, :2] = mine["cCentre"][cells] + rng.uniform(-0.3, 0.3, size=(k, 2)) * (hi - lo) / np.array(n)
    start["v"][:, :2] = rng.normal(scale=3.0, size=(k, 2)) + np.array([10.0, 0.0])
    start["cellV"][:, :2], start["cellRho"] = mine["cVel"][cells], mine["cRho"][cells]
    common = dict(eq_order=1, max_x=0.45, max_steps=4000, grav=[0.0, -9.81, 0.0])
    b = ref.ipt_integrate(orc.ipt_settings(p, max_length=-1.0, **common), start, record_cap=40)
    o = orc.Oracle(p, kind="2d")
    o.set_mesh(mine)
    a = o.ipt_integrate(orc.ipt_settings(p, max_length=eng.mesh_max_length(mine, 2), **common), start, record_cap=40)
    assert_same_tracks(a, b, "tau edge file")
    assert a["n_failed"] == 0 and a["n_steps"].max() >= 8      # at the readers' 4 x bound nothing trips the step-length test
