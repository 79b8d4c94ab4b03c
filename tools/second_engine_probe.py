"""Does an engine created after another one was used and freed in the same process run slower?  (A 2-GPU bench taken after
the slab parity check ran 2.3x slower, profiles/r2e_bench_n2_after_check.json.)  1 GPU: time 2 steps of the 12.5 M block on
a fresh engine, free it, run a 1 M engine, free it, and time the 12.5 M block again."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fjsph_b200 import cases, engine as eng


def timed(cells, label, steps=2):
    case = cases.synthetic_block(cells, 1e-3, jitter=0.1, seed=1234)
    p = dict(case["params"], delta_t_min=1e-9, frame_time_interval=1e9, solver_type=0, max_subits=3, min_residual=-30.0)
    e = eng.Engine(eng.default_params(3, **p), case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
    e.integrate()
    e.timers_reset()
    e.timers_enable(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e.integrate()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    tm = e.timers()
    print("%-28s %8.1f ms/step   force %.2f ms  prestep %.2f ms  free mem %.1f GB" % (
        label, ms, tm["force"]["ms"] / tm["force"]["calls"], tm["prestep"]["ms"] / tm["prestep"]["calls"],
        torch.cuda.mem_get_info()[0] / 1e9), flush=True)
    e.close()
    del e


timed((500, 250, 100), "fresh process, 12.5 M")
timed((208, 70, 70), "then 1 M")
timed((500, 250, 100), "then 12.5 M again")
timed((500, 250, 100), "and once more")
