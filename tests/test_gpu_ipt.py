"""The implicit particle tracker on the device (fjsph_ipt_integrate, csrc/ipt.cu: IPT::Integrate IPT.cpp:871-1107, FindFace
Containment.cpp:944-1079, Cross_Plane / MollerTrumbore / RayNormalIntersection Geometry.cpp:399-744) against the CPU oracle
and against the vectors FJSPH's own sources produced (tests/golden/ipt_*.npz).  ipt.cu is compiled without FMA contraction
and keeps the reference's order of operations, so cells, faces, outcomes, step and record counts must be identical and the
FP64 columns differ only through pow() / log10(): the bar is 1e-12 relative, far below the 1e-10 of the brief."""
import glob
import json
import os

import numpy as np
import pytest

from fjsph_b200 import _lib, cases, engine as eng
from oracle import oracle as orc
from tests import ipt_case
from tests.test_gpu_inlet import make_inlet_pair

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ipt_*.npz")))
TOL = 1e-12


def device_tracks(dim, mesh, settings, start, particle_step=1e-3, record_cap=ipt_case.RECORD_CAP):
    p = eng.default_params(dim, asource=1, particle_step=particle_step)
    e = eng.Engine(p, 64)
    e.upload_mesh(mesh)
    s, _ = eng.ipt_settings(p, **settings)
    return e.ipt_integrate(s, start, record_cap=record_cap)


def oracle_tracks(dim, mesh, settings, start, particle_step=1e-3, record_cap=ipt_case.RECORD_CAP):
    p = orc.default_params(dim, asource=1, particle_step=particle_step)
    o = orc.Oracle(p, kind="2d" if dim == 2 else None)
    o.set_mesh(mesh)
    return o.ipt_integrate(orc.ipt_settings(p, **settings), start, record_cap=record_cap)


def assert_tracks_close(got, ref, what, leftright, steps=True):
    """faceID: on a mesh whose quadrilaterals are split into two coplanar triangles a ray crosses both halves' plane at the
    same distance, and which half `dt < mindist` keeps is decided by the last bit of dt (one ulp of noise in Cd moves it on
    the CPU as well: the two libms differ in pow()).  The halves separate the same two cells, so the track is the same;
    faceID is compared up to that: where it differs, both faces must have the same cells on both sides."""
    leftright = np.asarray(leftright)
    assert (got["n_success"], got["n_failed"]) == (ref["n_success"], ref["n_failed"]), what
    assert np.array_equal(got["n_records"], ref["n_records"]), what
    if steps:
        assert np.array_equal(got["n_steps"], ref["n_steps"]), what
    worst = 0.0
    for part in ("last", "records"):
        for f in ipt_case.INT_FIELDS:
            if f == "faceID":
                a, b = got[part][f].ravel(), ref[part][f].ravel()
                dif = a != b
                assert (a[dif] >= 0).all() and (b[dif] >= 0).all() and np.array_equal(leftright[a[dif]], leftright[b[dif]]), (what, part, f)
                assert dif.sum() <= 0.05 * dif.size, (what, part, f, int(dif.sum()))
                continue
            assert np.array_equal(got[part][f], ref[part][f]), (what, part, f)
        for f in ipt_case.FLOAT_FIELDS:
            scale = max(np.abs(ref[part][f]).max(), 1e-300)
            err = np.abs(got[part][f] - ref[part][f]).max() / scale
            worst = max(worst, err)
            assert err <= TOL, (what, part, f, err)
    return worst


@pytest.mark.parametrize("name", list(ipt_case.CASES))
def test_device_tracker_against_the_oracle(name):
    case = ipt_case.build(name, n=400, seed=23)
    dim = case["dim"]
    settings = dict(case["settings"], max_length=case["length_factor"] * ipt_case.longest_edge(case["mesh"], dim))
    p = eng.default_params(dim, particle_step=case["particle_step"])
    for record in (1, 0):
        s = dict(settings, record=record)
        got = device_tracks(dim, case["mesh"], s, ipt_case.start_records(case, eng.IPT_START, p.sim_mass))
        ref = oracle_tracks(dim, case["mesh"], s, ipt_case.start_records(case, orc.IPT_START, p.sim_mass))
        worst = assert_tracks_close(got, ref, (name, record), case["mesh"]["leftright"])
        print("%s record=%d: %d left the mesh / passed max_x, %d failed, longest track %d steps, %d faceIDs on the twin half, "
              "worst relative difference %.2e" % (name, record, got["n_success"], got["n_failed"], got["n_steps"].max(),
                                                  int((got["last"]["faceID"] != ref["last"]["faceID"]).sum()), worst))
    assert ref["n_steps"].max() >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[4:-4] for p in FIXTURES])
def test_device_tracker_against_the_reference_vectors(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    mesh = {k[5:]: z[k] for k in z.files if k.startswith("mesh_")}
    got = device_tracks(meta["dim"], mesh, meta["settings"], z["start"], particle_step=meta["particle_step"], record_cap=meta["record_cap"])
    ref = dict(last=z["last"], records=z["records"], n_records=z["n_records"], n_success=meta["n_success"], n_failed=meta["n_failed"])
    assert_tracks_close(got, ref, os.path.basename(path), mesh["leftright"], steps=False)


def test_record_cap_bounds_and_errors():
    """Records beyond record_cap are dropped but counted; max_steps stops a march (failed = 2); the call needs a mesh and an
    equation order of 1 or 2; no particles is not an error."""
    case = ipt_case.build("quad2d_o2_long", n=50)
    settings = dict(case["settings"], max_length=4.0 * ipt_case.longest_edge(case["mesh"], 2))
    p = eng.default_params(2, asource=1, particle_step=1e-3)
    start = ipt_case.start_records(case, eng.IPT_START, p.sim_mass)
    full = device_tracks(2, case["mesh"], settings, start)
    short = device_tracks(2, case["mesh"], settings, start, record_cap=3)
    assert np.array_equal(short["n_records"], full["n_records"]) and full["n_records"].max() > 3
    for f in ipt_case.INT_FIELDS + ipt_case.FLOAT_FIELDS:
        assert np.array_equal(short["records"][f], full["records"][f][:, :3]), f
        assert np.array_equal(short["last"][f], full["last"][f]), f
    capped = device_tracks(2, case["mesh"], dict(settings, max_steps=2), start)
    ref = oracle_tracks(2, case["mesh"], dict(settings, max_steps=2), ipt_case.start_records(case, orc.IPT_START, p.sim_mass))
    assert (capped["last"]["failed"] == 2).sum() > 10 and capped["n_steps"].max() == 2
    assert_tracks_close(capped, ref, "max_steps", case["mesh"]["leftright"])
    e = eng.Engine(p, 64)
    s, _ = eng.ipt_settings(p, **settings)
    with pytest.raises(_lib.FjsphError, match="needs a mesh"):
        e.ipt_integrate(s, start)
    e.upload_mesh(case["mesh"])
    out = e.ipt_integrate(s, start[:0])
    assert (out["n_success"], out["n_failed"]) == (0, 0) and out["last"].shape == (0,)
    s.eq_order = 3
    with pytest.raises(_lib.FjsphError, match="Equation order not 1 or 2"):
        e.ipt_integrate(s, start)


def test_hand_off_from_the_delete_plane_to_the_tracker():
    """Integration.cpp:151-169 end to end: a jet whose FREE particles are coupled to a cross-flow mesh passes its delete plane;
    what fjsph_take_deleted hands over (id, time, state, the cell FindCell found and its solution) goes to the tracker, on the
    same mesh, and every particle is followed to the outer boundary or failed exactly as the oracle does it."""
    case = cases.inlet_jet(n=(5, 5, 4), fixed=1, jitter=0.03, aero_x=0.5, delete_x=2.5)
    case["params"] = dict(case["params"], acase=1, asource=1, ale=1, v_inf=(0.0, 30.0, 0.0), p_ref=100000.0, rho_g=1.2)
    mesh = cases.hex_mesh((-0.0123, -0.0031, -0.0029), (0.0117, 0.0073, 0.0071), (12, 5, 5),
                          vel=lambda c: np.stack([4.0 + 0 * c[:, 0], 30.0 + 800.0 * c[:, 0], 300.0 * c[:, 1]], axis=1), p=100000.0,
                          rho=lambda c: 1.2 + 5.0 * c[:, 2])
    o, e = make_inlet_pair(case)
    o.set_mesh(mesh)
    e.upload_mesh(mesh)
    handed = []
    for step in range(14):
        se = e.integrate()
        d = e.take_deleted()
        assert d["part_id"].shape[0] == se.n_del
        if se.n_del:
            handed.append(d)
    start = {k: np.concatenate([d[k] for d in handed]) for k in handed[0]}
    n = start["part_id"].shape[0]
    assert n >= 25 and (start["cellID"] >= 0).all() and np.abs(start["cellV"][:, 1]).min() > 1.0   # found by FindCell, solution attached
    p = e.params
    s, _ = eng.ipt_settings(p, eq_order=2, max_length=ipt_case.longest_edge(mesh, 3), max_steps=2000)   # (a tight bound: ipt_case)
    got = e.ipt_integrate(s, start, record_cap=40)
    rec = np.zeros(n, dtype=orc.IPT_START)
    for k in orc.IPT_START.names:
        rec[k] = start[k]
    so = orc.ipt_settings(o.params, eq_order=2, max_length=s.max_length, max_steps=2000)
    assert (so.diam, so.area, so.mu_g, so.rho_rest, list(so.grav)) == (s.diam, s.area, s.mu_g, s.rho_rest, list(s.grav))
    ref = o.ipt_integrate(so, rec, record_cap=40)
    worst = assert_tracks_close(got, ref, "hand-off", mesh["leftright"])
    print("hand-off: %d particles tracked, %d left the mesh, %d failed, longest track %d steps, worst relative difference %.2e"
          % (n, got["n_success"], got["n_failed"], got["n_steps"].max(), worst))
    assert got["n_success"] + got["n_failed"] == n and (got["last"]["going"] == 0).all()
    out = got["last"]["failed"] == 0
    assert (got["last"]["cellID"][out] == -2).all()                          # the successes left through the outer boundary
    assert got["n_steps"].max() >= 2
