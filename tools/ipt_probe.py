"""Throughput of the device particle tracker (fjsph_ipt_integrate, csrc/ipt.cu) beside the CPU restatement on one host
thread: N particles started in the upstream end of a sheared flow on a quadrilateral-faced box mesh, followed to the outflow
boundary.  Diagnostic only (the oracle is loaded as the yardstick, as in tools/diag_steps.py); prints one JSON line.
    python tools/ipt_probe.py [--n 1000000] [--cells 96,40,40] [--cpu-sample 20000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fjsph_b200 import cases, engine as eng  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--cells", default="96,40,40")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-gpu", action="store_true")
    a = ap.parse_args()
    cells = tuple(int(k) for k in a.cells.split(","))
    lo, hi = np.array([0.0, -0.2, -0.2]), np.array([0.96, 0.2, 0.2])
    vel = lambda c: np.stack([30 + 40 * c[:, 0] + 20 * c[:, 2], 0.4 * np.sin(9 * c[:, 0]) - 0.3, -0.5 + 1.5 * c[:, 1]], 1)
    t0 = time.time()
    mesh = cases.hex_mesh(lo, hi, cells, vel=vel, rho=lambda c: 1.1 + c[:, 2] + 0.3 * c[:, 0], triangulate=False)
    t_mesh = time.time() - t0
    rng = np.random.default_rng(3)
    n = a.n
    width = (hi - lo) / np.array(cells)
    ijk = np.stack([rng.integers(0, 4, size=n)] + [rng.integers(0, cells[d], size=n) for d in (1, 2)], axis=1)
    # the corner of the cross-section MollerTrumbore accepts (note Q9): tracks that cross the whole mesh
    frac = np.concatenate([rng.uniform(0.05, 0.95, size=(n, 1)), rng.uniform(0.12, 0.38, size=(n, 2))], axis=1)
    cid = ijk[:, 0] + cells[0] * (ijk[:, 1] + cells[1] * ijk[:, 2])
    start = np.zeros(n, dtype=eng.IPT_START)
    start["part_id"], start["cellID"], start["t"] = np.arange(n), cid, 0.0
    start["xi"] = lo + (ijk + frac) * width
    start["v"] = np.concatenate([rng.uniform(8.0, 14.0, size=(n, 1)), rng.normal(scale=0.05, size=(n, 2))], axis=1)
    start["cellV"], start["cellRho"] = mesh["cVel"][cid], mesh["cRho"][cid]
    p = eng.default_params(3, asource=1, particle_step=1e-3)
    start["mass"] = p.sim_mass
    # the bound on one step: the longest face diagonal itself (a fifth of cells.maxlength), as the committed measurement ran
    s, _ = eng.ipt_settings(p, eq_order=2, record=0, max_steps=4000, max_length=eng.mesh_max_length(mesh) / 5.0)
    out = dict(particles=n, cells=int(np.prod(cells)), mesh_build_s=round(t_mesh, 2))
    if not a.no_gpu:
        e = eng.Engine(p, 64)
        e.upload_mesh(mesh)
        e.ipt_integrate(s, start[:1000])                       # warm-up
        e.timers_enable(True)
        e.timers_reset()
        t0 = time.time()
        got = e.ipt_integrate(s, start)
        wall = time.time() - t0
        ms = e.timers()["ipt_integrate"]["ms"]
        steps = int(got["n_steps"].sum())
        out.update(gpu_kernel_ms=round(ms, 3), gpu_call_ms=round(1e3 * wall, 1), cell_steps=steps, left_the_mesh=got["n_success"],
                   failed=got["n_failed"], longest_track=int(got["n_steps"].max()),
                   gpu_cell_steps_per_s=round(steps / (1e-3 * ms)), gpu_particles_per_s_end_to_end=round(n / wall))
    if a.cpu_sample > 0:
        from oracle import oracle as orc  # yardstick only

        k = min(a.cpu_sample, n)
        po = orc.default_params(3, asource=1, particle_step=1e-3)
        o = orc.Oracle(po)
        o.set_mesh(mesh)
        so = orc.ipt_settings(po, eq_order=2, record=0, max_steps=4000, max_length=s.max_length)
        rec = np.zeros(k, dtype=orc.IPT_START)
        for f in orc.IPT_START.names:
            rec[f] = start[f][:k]
        t0 = time.time()
        ref = o.ipt_integrate(so, rec)
        cpu = time.time() - t0
        csteps = int(ref["n_steps"].sum())
        out.update(cpu_sample=k, cpu_s=round(cpu, 3), cpu_cell_steps_per_s=round(csteps / cpu), cpu_threads=1)
        if not a.no_gpu:
            same = bool(np.array_equal(ref["n_steps"], got["n_steps"][:k]) and np.array_equal(ref["last"]["cellID"], got["last"]["cellID"][:k]))
            out.update(sample_identical_cells_and_steps=same,
                       worst_rel_position_difference=float(np.abs(ref["last"]["xi"] - got["last"]["xi"][:k]).max() / np.abs(ref["last"]["xi"]).max()),
                       speedup_over_one_thread=round(out["gpu_cell_steps_per_s"] / out["cpu_cell_steps_per_s"], 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
