"""CPU checks that pin the oracle's mesh-containment and inlet restatements (no GPU):
  * a uniform mesh solution must reproduce the constant-freestream run bit for bit;
  * CheckCell's ray parity must agree with the analytic cell of a uniform hexahedral mesh;
  * the inlet inserts one particle per column whenever a BACK particle crosses the insertion plane, part_ids grow
    monotonically, and the delete plane erases exactly the particles beyond it."""
import numpy as np

from fjsph_b200 import cases
from oracle import oracle as orc


def test_uniform_mesh_equals_constant_freestream():
    case = cases.droplet(dx=0.0125, jitter=0.05)
    mesh = cases.hex_mesh((-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (6, 7, 5), vel=(0.0, 21.55, 0.0),
                          p=100000.0, rho=1.1025)
    pc = dict(case["params"], delta_t_min=1e-9)
    oc = orc.Oracle(orc.default_params(3, **pc))
    om = orc.Oracle(orc.default_params(3, asource=1, **pc))
    for o in (oc, om):
        o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    om.set_mesh(mesh)
    for _ in range(2):
        oc.integrate()
        om.integrate()
    for f in ("xi", "v", "rho", "Af", "acc"):
        assert np.array_equal(oc.get(f), om.get(f)), f
    assert np.abs(om.get("Af")).max() > 1.0 and (om.get("cellID") >= 0).sum() > 20


def test_containment_matches_the_analytic_cell_of_a_uniform_mesh():
    lo, hi, n = np.array([-0.1013, -0.1007, -0.1011]), np.array([0.1009, 0.1003, 0.1017]), (6, 7, 5)
    mesh = cases.hex_mesh(lo, hi, n, vel=(1.0, 2.0, 3.0))
    case = cases.droplet(dx=0.0125, jitter=0.05)
    o = orc.Oracle(orc.default_params(3, asource=1, **dict(case["params"], lam_cutoff=1e9)))  # every FREE particle is looked up
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(mesh)
    o.update_neighbours()
    o.prestep()
    o.aero_velocity()
    ijk = np.floor((case["xi"] - lo) / ((hi - lo) / np.array(n))).astype(int)
    want = (ijk[:, 2] * n[1] + ijk[:, 1]) * n[0] + ijk[:, 0]
    assert np.array_equal(o.get("cellID"), want)
    assert np.array_equal(o.get("cellV"), np.broadcast_to([1.0, 2.0, 3.0], case["xi"].shape))


def test_inlet_inserts_columns_and_delete_plane_erases():
    case = cases.inlet_jet(delete_x=2.5)
    B = case["block"]
    o = orc.Oracle(orc.default_params(3, **case["params"]))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    o.lib.orc_clear_blocks(o.h)
    o.add_block(1, B["first"], B["second"], block_type=6, fixed_vel_or_dynamic=0, insert_norm=B["insert_norm"],
                insconst=B["insconst"], delete_norm=B["delete_norm"], delconst=B["delconst"], aero_norm=B["aero_norm"],
                aeroconst=B["aeroconst"], back=B["back"], buffer=B["buffer"])
    n0, ncol = o.n, len(B["back"])
    added = deleted = 0
    for _ in range(14):
        _, st = o.integrate()
        assert st.n_add in (0, ncol)
        added += st.n_add
        deleted += st.n_del
        assert o.n == n0 + added - deleted
        pid = o.get("part_id")
        assert len(np.unique(pid)) == o.n
        assert (o.get("xi")[:, 0] <= B["delconst"] + 1e-12).all()
        b = o.get("b")
        assert (b == cases.BACK).sum() == ncol and (b == cases.BUFFER).sum() == 4 * ncol
    assert added >= 2 * ncol and deleted >= ncol and pid.max() == n0 + added - 1
