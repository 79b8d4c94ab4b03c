"""A few steps through the code paths added late in round 2, for compute-sanitizer: the stashed max |v_j - v_i| of the fused pass,
the prestep's wall vote (a tank), the 2D build with a 2D aero mesh, the three-part upload of fjsph_step_host."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fjsph_b200 import cases, engine as eng  # noqa: E402


def steps(case, dim, n, mesh=None, **kw):
    params = dict(case["params"], delta_t_min=1e-9, **kw)
    e = eng.Engine(eng.default_params(dim, **params), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    if mesh is not None:
        e.upload_mesh(mesh)
    for _ in range(n):
        s = e.integrate()
    x = e.download(("xi",))["xi"]
    assert np.isfinite(x).all()
    return e, s


steps(cases.droplet(dx=0.008, jitter=0.05), 3, 2)
steps(cases.box_with_walls(n=(8, 7, 9), dx=0.01, jitter=0.05), 3, 2)
steps(cases.box_with_walls(n=(8, 7, 9), dx=0.01, jitter=0.05), 3, 1, solver_type=1)
d2 = cases.droplet(dx=0.005, dim=2, jitter=0.05)
steps(d2, 2, 2, mesh=cases.quad_mesh((-0.1013, -0.0303), (0.1009, 0.1003), (8, 6), vel=(21.55, 0.0), p=1e5, rho=1.1), asource=1)
steps(cases.box_with_walls(n=(14, 10), dx=0.01, layers=4, jitter=0.05, dim=2), 2, 2)
case = cases.synthetic_block((10, 8, 7), 1e-3, jitter=0.1, seed=3)
n = case["xi"].shape[0]
e = eng.Engine(eng.default_params(3, **dict(case["params"], delta_t_min=1e-9)), n)
state = dict(xi=case["xi"], v=case["v"], acc=np.zeros_like(case["xi"]), rho=case["rho"], Rrho=np.zeros(n), p=case["p"], m=case["m"],
             b=case["b"])
for _ in range(3):
    out, st = e.step_host(state, 0, 1, out_fields=("xi", "v", "acc", "rho", "Rrho", "p"))
    state.update(out)
print("memcheck paths done")
