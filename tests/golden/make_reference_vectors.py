"""Generates tests/golden/ref_*.npz: outputs of FJSPH's OWN time-step sources (compiled unmodified from /root/reference/src
into oracle/_ref/ by oracle/Makefile.ref, against the stand-in Eigen / nanoflann headers of oracle/shim/) on small seeded
cases.  Run here, where /root/reference exists; the fixtures travel, the reference does not:

    python tests/golden/make_reference_vectors.py

Each file holds the inputs (state arrays, settings as JSON, LIMITS blocks, mesh) and, after `steps` calls of
Integrator::integrate, the per-step table (sub-iterations, dt, rms error, maxima, insertions / deletions) and every
SPHPart field of pnp1.  tests/test_golden_reference.py replays the inputs through the CPU oracle (CPU suite) and through
the CUDA engine's C ABI (GPU suite) and compares.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fjsph_b200 import cases  # noqa: E402  (host-side input generation only)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FLOATS = orc._VEC_FIELDS + ("L",) + tuple(orc._SCALAR_FIELDS)
INTS = ("part_id", "cellID", "b", "surf", "surfzone", "internal")
STATS = ("dt", "cfl_ratio", "rms_error", "maxRho_pc", "maxf", "maxAf", "maxShift", "safe_dt")
ISTATS = ("iterations", "n_add", "n_del", "total_points")
BLOCK_KEYS = ("first", "second", "is_fluid", "block_type", "fixed_vel_or_dynamic", "insert_norm", "insconst", "delete_norm",
              "delconst", "aero_norm", "aeroconst", "back", "buffer")


def golden_cases():
    """name -> (case, kind of reference binary, dim, steps, extra settings, mesh or None, initial cellID or None)"""
    out = {}
    blk = cases.synthetic_block(n=(9, 8, 7), jitter=0.1)
    out["block_nb_ale"] = (blk, "ref3d", 3, 3, dict(ale=1), None, None)
    out["block_rk4_ale"] = (blk, "ref3d", 3, 3, dict(ale=1, solver_type=1), None, None)
    out["block_nb_dsph"] = (blk, "ref3d_dsph", 3, 3, dict(ale=0), None, None)
    out["block_nb_ale_ties"] = (cases.synthetic_block(n=(9, 8, 7), jitter="eps"), "ref3d", 3, 2, dict(ale=1), None, None)
    drop = cases.droplet(dx=0.008, jitter=0.05)
    out["droplet_gissler"] = (drop, "ref3d", 3, 3, dict(ale=1), None, None)
    out["droplet_gissler_tab_nolam"] = (drop, "ref3d", 3, 2, dict(ale=1, use_TAB_def=1, use_lam=0), None, None)
    out["droplet_induced_pressure"] = (drop, "ref3d", 3, 2, dict(ale=1, acase=2), None, None)
    out["droplet_skin_friction"] = (drop, "ref3d", 3, 2, dict(ale=1, acase=3), None, None)
    out["droplet_dsph"] = (drop, "ref3d_dsph", 3, 2, dict(ale=0), None, None)
    tank = cases.box_with_walls(n=(7, 6, 6), jitter=0.05)
    out["tank_adami_nb"] = (tank, "ref3d", 3, 3, dict(ale=1), None, None)
    out["tank_adami_rk4"] = (tank, "ref3d", 3, 3, dict(ale=1, solver_type=1), None, None)
    out["tank_iso_eos"] = (tank, "ref3d", 3, 2, dict(ale=1, pressure_rel=1), None, None)
    out["dam_2d"] = (cases.dam_2d(dx=0.05), "ref2d", 2, 3, dict(ale=1), None, None)
    for fixed in (0, 1):
        jet = cases.inlet_jet(delete_x=2.5, fixed=fixed, jitter=0.02)
        out["inlet_jet_fixed%d" % fixed] = (jet, "ref3d", 3, 14, dict(ale=1), None, None)
    # other wall treatments (Resid.cpp:122-186, Newmark_Beta.cpp:69-132): Ghost continuity, no-slip mirror velocities, a
    # wall moving on a schedule.  (DBC walls cannot be pinned: Boundary_DBC sizes its scratch vector by the END of the
    # wall block and then writes it at FLUID indices, Resid.cpp:84-107 -- heap corruption in the reference.)
    nb, n = tank["bound_points"], tank["xi"].shape[0]

    def walls(solver, no_slip, times=None, vels=None):
        return dict(tank, blocks=[dict(is_fluid=0, first=0, second=nb, bound_solver=solver, no_slip=no_slip, times=times,
                                       vels=vels, fixed_vel_or_dynamic=0),
                                  dict(is_fluid=1, first=nb, second=n)])

    half, mid = nb // 2, (nb + n) // 2
    multi = dict(tank, blocks=[dict(is_fluid=0, first=0, second=half, bound_solver=1, no_slip=0),
                               dict(is_fluid=0, first=half, second=nb, bound_solver=2, no_slip=1),
                               dict(is_fluid=1, first=nb, second=mid), dict(is_fluid=1, first=mid, second=n)])
    out["tank_two_wall_blocks_two_fluid_blocks"] = (multi, "ref3d", 3, 3, dict(ale=1), None, None)
    out["tank_ghost_noslip"] = (walls(2, 1), "ref3d", 3, 3, dict(ale=1), None, None)
    out["tank_adami_moving_wall"] = (walls(1, 0, [0.0, 0.004, 0.009], [[0.1, 0, 0], [0, 0.2, 0], [0, 0, 0]]), "ref3d", 3, 3,
                                     dict(ale=1), None, None)
    out["tank_ghost_rk4_moving"] = (walls(2, 1, [0.0, 0.004], [[0.1, 0, 0], [0, 0.2, 0]]), "ref3d", 3, 3,
                                    dict(ale=1, solver_type=1), None, None)
    # Runge-Kutta over the other ingredients (each has its own code path in Runge_Kutta.cpp)
    out["droplet_gissler_rk4"] = (drop, "ref3d", 3, 3, dict(ale=1, solver_type=1), None, None)
    out["block_rk4_dsph"] = (blk, "ref3d_dsph", 3, 3, dict(ale=0, solver_type=1), None, None)
    out["inlet_jet_fixed1_rk4"] = (cases.inlet_jet(delete_x=2.5, fixed=1, jitter=0.02), "ref3d", 3, 14,
                                   dict(ale=1, solver_type=1), None, None)
    dm = cases.droplet(dx=0.0125, jitter=0.05)
    sheared = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (6, 7, 5),
                             vel=lambda c: np.stack([5 + 20 * c[:, 1], 21.55 + 0 * c[:, 0], 3 * c[:, 2]], 1), p=100000.0,
                             rho=1.1025)
    # FindCell reads cells.cFaces[cellID] before anything else (Containment.cpp:592-600): the reference needs a valid
    # cell id on every FREE particle, so the fixture starts all of them in cell 0
    out["droplet_sheared_mesh"] = (dm, "ref3d", 3, 3, dict(ale=1, asource=1, delta_t_min=1e-9), sheared, 0)
    wall = cases.hex_mesh((-0.1013, -0.1007, -0.03), (0.1009, 0.1003, 0.1017), (6, 7, 5), vel=(0.0, 21.55, 0.0), p=100000.0,
                          rho=1.1025, outer_marker=-1)
    out["droplet_inner_wall_mesh"] = (dm, "ref3d", 3, 3, dict(ale=1, asource=1, delta_t_min=1e-9), wall, 0)
    # Check_Pipe_Outlet with a mesh (Containment.cpp:822-890): PIPE particles crossing the aero plane become FREE and take
    # their first cell from the 150 nearest cell centres (FirstCell, Containment.cpp:425-470)
    pj = cases.inlet_jet(n=(5, 5, 4), fixed=1, jitter=0.03, aero_x=0.5)
    pj["params"] = dict(pj["params"], acase=1, v_inf=(0.0, 30.0, 0.0), p_ref=100000.0, rho_g=1.2)
    pipe_mesh = cases.hex_mesh((-0.0123, -0.0031, -0.0029), (0.0117, 0.0073, 0.0071), (12, 5, 5), vel=(0.0, 30.0, 0.0),
                               p=100000.0, rho=1.2)
    out["inlet_jet_mesh_first_cell"] = (pj, "ref3d", 3, 6, dict(ale=1, asource=1), pipe_mesh, None)
    out["jet_deck_jittered"] = (jittered_jet_deck(), "ref3d", 3, 12, {}, None, None)
    out["droplet_sheared_mesh_rk4"] = (dm, "ref3d", 3, 3, dict(ale=1, asource=1, delta_t_min=1e-9, solver_type=1), sheared, 0)
    # Check_Error's unstable-step branch (Newmark_Beta.cpp:32-48): a CFL number of 6 makes the sub-iterations diverge, so
    # pnp1 = pn, the list is rebuilt, dt halves and the step starts over -- once at step 0, twice at the steps after it
    # (dt = cfl * safe_dt / 2^k shows k); with walls the rebuild happens between two wall treatments
    unstable = dict(ale=1, cfl=6.0, cfl_max=6.0, max_subits=2, delta_t_max=1.0, delta_t_min=1e-12)
    out["block_nb_unstable_restart"] = (cases.synthetic_block(n=(10, 9, 8), jitter=0.1), "ref3d", 3, 3, unstable, None, None)
    out["tank_nb_unstable_restart"] = (tank, "ref3d", 3, 3, dict(unstable, cfl=8.0, cfl_max=8.0), None, None)
    return out


def add_full_block(o, B):
    """One LIMITS entry with every field it may carry (wall treatment, schedule, inlet planes and tables)."""
    o.add_block(B["is_fluid"], B["first"], B["second"], bound_solver=B.get("bound_solver", 1), no_slip=B.get("no_slip", 0),
                block_type=B.get("block_type", 0), fixed_vel_or_dynamic=B.get("fixed_vel_or_dynamic", 0),
                times=B.get("times"), vels=B.get("vels"), insert_norm=B.get("insert_norm"),
                insconst=B.get("insconst", 9999999.0), delete_norm=B.get("delete_norm"), delconst=B.get("delconst", 9999999.0),
                aero_norm=B.get("aero_norm"), aeroconst=B.get("aeroconst", 9999999.0),
                back=None if B.get("back") is None else np.asarray(B["back"], dtype=np.int64),
                buffer=None if B.get("buffer") is None else np.asarray(B["buffer"], dtype=np.int64))


def jittered_jet_deck():
    """tests/decks/jet3d.para through the product's front end (round dynamic inlet rotated into +y, hollow Ghost pipe, Gissler
    cross flow), every particle moved off its lattice site by U(-0.03, 0.03) dx so that the case is a generic input and
    not a tie-stress one.  12 steps: three bursts of insertions, PIPE -> FREE transitions, the pipe wall's near-inlet
    logic."""
    from fjsph_b200 import frontend
    from tests.util import INPUT_PARAMS

    c = frontend.read_case(os.path.join(ROOT, "tests", "decks", "jet3d.para"), 3)
    P = c["params"]
    params = {k: (tuple(getattr(P, k)) if hasattr(getattr(P, k), "__len__") else getattr(P, k)) for k in INPUT_PARAMS}
    rng = np.random.default_rng(5)
    xi = c["xi"] + rng.uniform(-0.03, 0.03, c["xi"].shape) * P.particle_step
    blocks = []
    for B in c["blocks"]:
        Bk = {k: B[k] for k in ("is_fluid", "first", "second", "bound_solver", "no_slip", "block_type", "fixed_vel_or_dynamic",
                                 "insconst", "delconst", "aeroconst")}
        for k in ("insert_norm", "delete_norm", "aero_norm"):
            Bk[k] = [float(x) for x in B[k]]
        Bk["times"] = None if B["times"] is None else [float(t) for t in B["times"]]
        Bk["vels"] = None if B["vels"] is None else np.asarray(B["vels"]).tolist()
        if B.get("back") is not None:
            Bk["back"] = np.asarray(B["back"]).tolist()
            Bk["buffer"] = np.asarray(B["buffer"]).tolist()
        blocks.append(Bk)
    return dict(xi=xi, v=c["v"], rho=c["rho"], p=c["p"], m=c["m"], b=c["b"], bound_points=c["bound_points"], params=params,
                blocks=blocks)


def make_sim(case, kind, dim, extra, mesh, cell0):
    """One simulation on `kind` (None / '2d' = the oracle restatement, 'ref*' = the compiled reference)."""
    P = orc.default_params(dim, **dict(case["params"], **extra))
    o = orc.Oracle(P, kind=kind)
    if mesh is not None:
        o.set_mesh(mesh)
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case.get("bound_points", 0))
    B = case.get("block")
    if B is not None:
        o.lib.orc_clear_blocks(o.h)
        o.add_block(1, B["first"], B["second"], block_type=B["block_type"], fixed_vel_or_dynamic=B["fixed_vel_or_dynamic"],
                    insert_norm=B["insert_norm"], insconst=B["insconst"], delete_norm=B.get("delete_norm"),
                    delconst=B.get("delconst", 9999999.0), aero_norm=B["aero_norm"], aeroconst=B["aeroconst"], back=B["back"],
                    buffer=B["buffer"])
    if case.get("blocks") is not None:
        o.lib.orc_clear_blocks(o.h)
        for Bk in case["blocks"]:
            add_full_block(o, Bk)
    if cell0 is not None:
        for lvl in (0, 1):
            o.set("cellID", np.full(o.n, cell0, dtype=np.int64), lvl)
    return o


def run(o, steps):
    table = {k: [] for k in STATS + ISTATS}
    for _ in range(steps):
        _, s = o.integrate()
        for k in STATS + ISTATS:
            table[k].append(getattr(s, k))
    return table


def main():
    if not orc.have_ref():
        orc.build_ref()
    only = set(sys.argv[1:])  # optional: names of the vectors to (re)make
    for name, (case, kind, dim, steps, extra, mesh, cell0) in golden_cases().items():
        if only and name not in only:
            continue
        o = make_sim(case, kind, dim, extra, mesh, cell0)
        # stdout of the reference's step table is noise here
        table = run(o, steps)
        data = {"in_" + k: np.asarray(case[k]) for k in ("xi", "v", "rho", "p", "m", "b")}
        meta = dict(name=name, kind=kind, dim=dim, steps=steps, bound_points=int(case.get("bound_points", 0)),
                    params={k: (list(v) if hasattr(v, "__len__") else v) for k, v in dict(case["params"], **extra).items()},
                    cell0=cell0)
        if case.get("block") is not None:
            B = case["block"]
            meta["block"] = {k: (np.asarray(B[k]).tolist() if hasattr(B[k], "__len__") else B[k]) for k in BLOCK_KEYS if k in B}
        if case.get("blocks") is not None:
            meta["blocks"] = case["blocks"]
        if mesh is not None:
            for k, v in mesh.items():
                data["mesh_" + k] = np.asarray(v)
        for k in STATS:
            data["step_" + k] = np.asarray(table[k], dtype=np.float64)
        for k in ISTATS:
            data["step_" + k] = np.asarray(table[k], dtype=np.int64)
        for f in FLOATS + INTS:
            data["out_" + f] = o.get(f, 1)
        data["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, "ref_%s.npz" % name)
        np.savez_compressed(path, **data)
        print("%-28s %5d particles, %2d steps, sub-iterations %s, %6.1f kB" % (
            name, o.n, steps, table["iterations"], os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
