"""Divergence of the engine from the CPU oracle over a long run (north_star: "state after 1000 steps within a stated
tolerance").  Prints the normwise relative error of x, rho, v at checkpoints; GPU box only."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fjsph_b200 import cases
from tests.util import make_pair, relerr

which = sys.argv[1] if len(sys.argv) > 1 else "block"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if which == "block":
    case = cases.synthetic_block((10, 9, 8), 1e-3, jitter=0.1, seed=8)
elif which == "droplet":
    case = cases.droplet(dx=0.0085, jitter=0.05)
else:
    case = cases.box_with_walls(n=(6, 6, 8), dx=0.01, layers=4, jitter=0.05)
o, e, p = make_pair(case, delta_t_min=1e-9)
x0 = case["xi"].copy()
t0 = time.time()
marks = [1, 3, 10, 30, 100, 300, 1000, 3000]
its_mismatch = 0
for s in range(1, steps + 1):
    _, so = o.integrate()
    se = e.integrate()
    its_mismatch += so.iterations != se.iterations
    if s in marks or s == steps:
        got = e.download(("xi", "v", "rho", "surf"))
        disp = np.abs(o.get("xi") - x0).max()
        print("%s step %5d  t=%.3es  its %2d/%2d (mismatched so far %d)  x %.2e  (vs displacement %.2e: %.2e)  rho %.2e  v %.2e  surf flips %d  [%.0fs]" % (
            which, s, o.params.current_time, so.iterations, se.iterations, its_mismatch, relerr(got["xi"], o.get("xi")), disp,
            np.abs(got["xi"] - o.get("xi")).max() / max(disp, 1e-300), relerr(got["rho"], o.get("rho")),
            relerr(got["v"], o.get("v")), int((got["surf"] != o.get("surf")).sum()), time.time() - t0), flush=True)
