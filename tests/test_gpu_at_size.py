"""Parity at the sizes BASELINE.json's configurations state (C2, C3, C4) and of the paths only large inputs take.

The small parity cases (tests/test_gpu_parity.py: <= ~2000 particles) are almost all surface particles with short
lists.  Here the engine meets the oracle on a million particles: rows hundreds of particles long, the bulk / near-surface
split of the fused surface + shifting sweep decided from real counts, list capacities that grow, key tables that would
not fit.  The oracle is the parity build with its particle loops threaded (oracle/lib/liborc3d_mt.so: per-particle sums in
the serial order, so per-particle results equal the serial build's bit for bit).  max_subits is pinned to 3 so that one
step (1 + 4 force evaluations) of a million particles costs the CPU tens of seconds; bars as in test_full_step_parity.
"""
import os

import numpy as np
import pytest

from fjsph_b200 import cases, engine as eng, frontend
from oracle import oracle as orc
from tests.util import assert_fields_close, make_pair_from_deck, relerr

pytestmark = pytest.mark.gpu
DECKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks")
STATE, RATES, FLAGS = ("xi", "rho", "lam", "lam_nb"), ("acc", "Rrho", "Af", "aVisc", "deltaD", "vPert"), ("surf", "surfzone", "cellID", "b")


def pair_mt(case, **kw):
    params = dict(case["params"], **kw)
    o = orc.Oracle(orc.default_params(3, **params), kind="3d_mt")
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    e = eng.Engine(eng.default_params(3, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    return o, e


def step_and_compare(o, e, steps, ctx):
    for step in range(steps):
        _, so = o.integrate()
        se = e.integrate()
        c = "%s step %d" % (ctx, step)
        assert se.iterations == so.iterations, c
        assert abs(se.dt - so.dt) <= 1e-12 * so.dt, c
        assert abs(se.npd - so.npd) <= 1e-10 * abs(so.npd), c
        assert np.array_equal(e.neighbour_counts(), o.neighbour_counts()), c
        assert_fields_close(e, o, FLAGS, context=c)
        assert_fields_close(e, o, STATE, tol=1e-10, context=c)
        assert_fields_close(e, o, ("v", "p"), tol=1e-8, context=c)
        assert_fields_close(e, o, RATES, tol=1e-6, context=c)


def test_c2_droplet_one_million():
    """BASELINE.json configs[1]: Examples/Droplet at dx = 0.0008 -- 1 022 208 particles, pairwise surface tension, Gissler
    aero -- on generic positions (lattice + U(-0.05, 0.05) dx; the deck's own U(0, eps dx) perturbation is a tie-stress
    input with looser bars, tests/test_gpu_decks.py)."""
    case = cases.droplet(dx=0.0008, jitter=0.05)
    assert case["xi"].shape[0] > 1_000_000
    o, e = pair_mt(case, max_subits=3, delta_t_min=1e-9)
    step_and_compare(o, e, 1, "C2 droplet")
    assert e.download(("surf",))["surf"].sum() > 10_000  # a real free surface


def test_c3_standing_column_one_million():
    """BASELINE.json configs[2]: the 3D extrusion of Examples/Standing_Column -- 101 x 51 x 201 fluid particles in an
    open tank of five Pressure-Gradient (Adami) walls four particles thick, hydrostatic start.  Two steps against the
    oracle, and the example's own check: p / (rho g h) against depth / h is the line of Examples/Standing_Column/Ideal.dat
    (0 0 -> 1 1)."""
    case = cases.box_with_walls(n=(101, 51, 201), dx=0.01, layers=4, jitter=0.05)
    nb = case["bound_points"]
    assert case["xi"].shape[0] - nb == 101 * 51 * 201
    o, e = pair_mt(case, max_subits=3, delta_t_min=1e-9)
    step_and_compare(o, e, 2, "C3 column")
    got = e.download(("p", "xi", "b"))
    fluid = got["b"] != cases.BOUND
    h, g, rho0 = case["height"], 9.81, 1000.0
    depth = (h - got["xi"][fluid, 2]) / h
    pn = got["p"][fluid] / (rho0 * g * h)
    inner = depth > 0.05  # below the free-surface layer, where the kernel support is full
    dev = np.abs(pn[inner] - depth[inner])
    # jittered particles (+-0.05 dx) two steps after a hydrostatic start: the line holds in the mean and in the slope; single
    # particles scatter by the few per cent of rho g h their 1e-5 density noise is worth under the stiff EOS (c = 10 sqrt(g h))
    assert dev.mean() < 0.01 and np.percentile(dev, 99.0) < 0.05 and dev.max() < 0.15, (dev.mean(), np.percentile(dev, 99.0), dev.max())
    assert abs(np.polyfit(depth[inner], pn[inner], 1)[0] - 1.0) < 0.01  # slope of the ideal line


def test_c4_crossflow_deck_coupled_to_a_tau_mesh(tmp_path):
    """BASELINE.json configs[3]: Examples/Crossflow (3D) at the reference's own numbers (tests/decks/crossflow3d.para: round
    dynamic inlet, hollow Ghost pipe, 43 k particles at 3e-5 spacing, Gissler aero) coupled to a TAU mesh: the mesh and its
    solution are written as NetCDF-3 files and read back by fjsph_tau_read.  (i) four steps against the oracle with the mesh;
    (ii) run on until the jet has left the pipe: the particles the lookup finds carry the mesh's (uniform) solution."""
    from tests.tau_case import write_tau

    case = frontend.read_case(os.path.join(DECKS, "crossflow3d.para"), 3)
    P = case["params"]
    assert case["xi"].shape[0] > 40_000
    lo = case["xi"].min(0) - 2.0e-3
    hi = case["xi"].max(0) + np.array([2.0e-3, 6.0e-3, 2.0e-3])
    vinf, pref, rhog = tuple(P.v_inf), P.p_ref, P.rho_g
    mesh_file, sol_file, *_ = write_tau(tmp_path, lo, hi, (9, 11, 8), lambda x: vinf, lambda x: pref, lambda x: rhog,
                                        wall_marker=-2)
    tau = frontend.read_tau(mesh_file, sol_file)
    # (i) oracle parity with the mesh: the first steps, while the column is still inside the pipe
    o, e = make_pair_from_deck(case, kind="3d_mt", asource=1)
    o.set_mesh(tau)
    e.upload_mesh(tau)
    for step in range(4):
        _, so = o.integrate()
        se = e.integrate()
        ctx = "crossflow step %d" % step
        assert (se.n_add, se.n_del, se.total_points) == (so.n_add, so.n_del, so.total_points), ctx
        assert abs(se.iterations - so.iterations) <= 1 and abs(se.dt - so.dt) <= 1e-6 * so.dt, ctx
        got = e.download(("part_id", "b", "xi", "v", "rho", "cellID"))
        assert np.array_equal(got["part_id"], o.get("part_id")) and np.array_equal(got["b"], o.get("b")), ctx
        # the deck's lattice + U(0, eps dx) positions are a tie-stress input: the bars of tests/test_gpu_decks.py
        assert relerr(got["xi"], o.get("xi")) <= 1e-6 and relerr(got["rho"], o.get("rho")) <= 1e-6, ctx
        assert relerr(got["v"], o.get("v")) <= 1e-4, ctx
    # (ii) the jet leaves the pipe, crosses the aero entry plane (PIPE -> FREE, FirstCell on the device) and meets the cross
    # flow: every FREE particle the lookup placed in a cell carries that cell's solution -- here the uniform free stream --
    # and feels the Gissler drag.  (The constant-free-stream run is NOT the same run on this deck: get_aero_velocity zeroes
    # cellV of the PIPE / BUFFER particles every step there, the mesh branch leaves them alone, Resid.cpp:478-611, and the
    # aero term is evaluated for both, Resid.cpp:267-277.)
    # The deck's frame interval (1e-6 s) is a few steps long and find_timestep clamps dt to what is left of the frame
    # (Integration.cpp:433-440), so the host marches the frames as FJSPH's main does (FJSPH.cpp:262-330): step until the
    # frame is full, then last_frame_time += frame_time_interval.
    n_add, t_sim, step = 0, 0.0, 0
    dt_min, frame_dt = e.params.delta_t_min, e.params.frame_time_interval
    stept = e.params.current_time - e.params.last_frame_time  # part (i) ran inside the first frame
    for frame in range(60):
        while stept + 0.1 * dt_min < frame_dt:
            s1 = e.integrate()
            n_add += s1.n_add
            stept += s1.dt
            step += 1
            assert s1.dt > 0.0 and step < 1500, (frame, step, s1.dt)
        t_sim += stept
        stept = 0.0
        e.set_params(last_frame_time=e.params.last_frame_time + frame_dt)
        if (e.download(("b",))["b"] == cases.FREE).sum() > 200:
            break
    a = e.download(("xi", "v", "rho", "Af", "b", "cellID", "cellV", "cellP", "cellRho"))
    free = a["b"] == cases.FREE
    found = free & (a["cellID"] >= 0)
    diag = dict(steps=step, frames=frame + 1, t=t_sim, dt=s1.dt, free=int(free.sum()), found=int(found.sum()), n_add=n_add,
                ymax=float(a["xi"][:, 1].max()), cell_ids=np.unique(a["cellID"])[:8].tolist())
    assert free.sum() > 200 and n_add > 0 and found.sum() > 100, diag
    assert a["cellID"][found].max() < 9 * 11 * 8
    assert np.abs(a["cellV"][found] - np.asarray(vinf)).max() <= 1e-12 * max(abs(v) for v in vinf)
    assert np.abs(a["cellP"][found] - pref).max() <= 1e-9 * pref and np.abs(a["cellRho"][found] - rhog).max() <= 1e-12 * rhog
    assert np.isfinite(a["Af"]).all() and np.abs(a["Af"][found]).max() > 0.0 and np.isfinite(a["xi"]).all()
    assert a["Af"][found][:, 0].mean() > 0.0  # the cross flow pushes along +x


def test_key_table_doubling_and_wide_rows(monkeypatch):
    """A cross-section whose one-spacing-wide rows would need more cell keys than the table may hold makes the rows wider
    until it fits (neighbours.cu, rebuild_skin; 2^25 keys by default, i.e. cross-sections beyond ~500 x 500 rows).  The
    limit is lowered to 2^13 keys here so that a 96 k-particle block takes that path -- rows two and four spacings wide,
    windows of up to 40 particles split over several slots -- and the lists must still be the oracle's, bit for bit."""
    monkeypatch.setenv("FJSPH_B200_MAX_KEY_BITS", "13")
    case = cases.synthetic_block((60, 40, 40), 1e-3, jitter=0.2, seed=11)
    o, e = pair_mt(case, max_subits=3, delta_t_min=1e-9)
    o.update_neighbours()
    e.update_neighbours()
    off_o, idx_o, _ = o.neighbours()
    off_e, idx_e = e.neighbours()
    assert np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e)
    step_and_compare(o, e, 2, "wide rows")


@pytest.mark.parametrize("axis", [1, 2])
def test_rows_along_another_axis(monkeypatch, axis):
    """The row axis is the longest extent of the bounding box; a column that is tallest along z (C3) or a jet along y takes
    rows along that axis.  Same lists, same step, whatever the axis."""
    monkeypatch.setenv("FJSPH_B200_ROW_AXIS", str(axis))
    case = cases.synthetic_block((20, 14, 12), 1e-3, jitter=0.15, seed=3)
    o, e = pair_mt(case, max_subits=3, delta_t_min=1e-9)
    o.update_neighbours()
    e.update_neighbours()
    off_o, idx_o, _ = o.neighbours()
    off_e, idx_e = e.neighbours()
    assert np.array_equal(off_o, off_e) and np.array_equal(idx_o, idx_e)
    step_and_compare(o, e, 2, "row axis %d" % axis)


@pytest.mark.parametrize("below", ["2.0", "-1.0"], ids=["always_split", "never_split"])
def test_lean_and_near_surface_launches_agree(monkeypatch, below):
    """The fused surface + shifting sweep runs as two launches (bulk particles with the lean body, the near-surface rest
    with the full one) when few warps hold near-surface particles, as one launch otherwise (sweeps.cu, k_surf23_shift
    CLASS).  Forced either way on a block whose warps mix both kinds, the step is the oracle's."""
    monkeypatch.setenv("FJSPH_B200_SPLIT_BELOW", below)
    case = cases.synthetic_block((70, 18, 16), 1e-3, jitter=0.1, seed=5)
    o, e = pair_mt(case, max_subits=3, delta_t_min=1e-9)
    step_and_compare(o, e, 2, "split below " + below)


@pytest.mark.parametrize("flavour", ["rk4", "dsph"])
def test_quarter_million_block_other_solvers(flavour):
    """The C5 block workload at 262 144 particles (the CPU arm's sample, bench.py) through the two paths the million-particle
    cases above do not take: Runge-Kutta 4 (Runge_Kutta.cpp:462-476: four force evaluations on FROZEN pair distances, the
    stage accumulators) and the delta-SPH build without ALE shifting (ALE = 0: no vPert terms in the force sweep, the plain
    continuity equation).  One step against the threaded parity oracle."""
    case = cases.synthetic_block((64, 64, 64), 1e-3, jitter=0.1, seed=1234)
    kw = dict(delta_t_min=1e-9, max_subits=3)
    kw.update(dict(solver_type=1) if flavour == "rk4" else dict(ale=0))
    o, e = pair_mt(case, **kw)
    step_and_compare(o, e, 1, "block 64^3 " + flavour)
