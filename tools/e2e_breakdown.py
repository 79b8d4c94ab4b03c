"""Where the end-to-end step spends its time: upload_state, the step, download_state, each timed with host clocks
around a device sync, plus raw pinned H2D / D2H copies of the same byte counts for reference."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fjsph_b200 import cases, engine as eng  # noqa: E402

cells = tuple(int(k) for k in (sys.argv[1] if len(sys.argv) > 1 else "500,250,100").split(","))
case = cases.synthetic_block(n=cells, jitter=0.1)
params = dict(case["params"], ale=1, max_subits=3, min_residual=-30.0, delta_t_min=1e-9)
n = case["xi"].shape[0]
e = eng.Engine(eng.default_params(3, **params), n)
e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
for _ in range(2):
    e.integrate()


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


st = e.download(("xi", "v", "acc", "rho", "Rrho", "p", "m", "b"))
ins_t = {k: pinned(v) for k, v in st.items()}
ins = {k: v.numpy() for k, v in ins_t.items()}
out_fields = ("xi", "v", "acc", "rho", "Rrho", "p")
tu = ts = td = 0.0
K = 4
for it in range(K + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.upload_state(ins["xi"], ins["v"], ins["rho"], ins["p"], ins["m"], ins["b"], 0, acc=ins["acc"], Rrho=ins["Rrho"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e.integrate()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e.download(out_fields, out=ins)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if it:
        tu, ts, td = tu + t1 - t0, ts + t2 - t1, td + t3 - t2
h2d = sum(v.nbytes for v in ins.values())
d2h = sum(ins[k].nbytes for k in out_fields)
print("particles %d  upload %.1f ms (%.2f GB, %.1f GB/s)  step %.1f ms  download %.1f ms (%.2f GB, %.1f GB/s)" % (
    n, tu / K * 1e3, h2d / 1e9, h2d / (tu / K) / 1e9, ts / K * 1e3, td / K * 1e3, d2h / 1e9, d2h / (td / K) / 1e9))
big = torch.empty(h2d, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(h2d, dtype=torch.uint8, device="cuda")
for name, fn, nb in (("H2D", lambda: dev.copy_(big, non_blocking=True), h2d), ("D2H", lambda: big[:d2h].copy_(dev[:d2h], non_blocking=True), d2h)):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("raw pinned %s of %.2f GB: %.1f ms (%.1f GB/s)" % (name, nb / 1e9, dt * 1e3, nb / dt / 1e9))
