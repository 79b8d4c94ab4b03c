"""Slab-decomposition parity: the same case on WORLD_SIZE GPUs (x-slabs, NCCL halos) and on one GPU.

Run under torchrun:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/slab_check.py
Rank 0 also runs the undecomposed engine and compares every owned particle by part_id.  Exit code 0 = parity.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fjsph_b200 import cases, engine as eng, slab

FIELDS = ("part_id", "xi", "v", "rho", "p", "acc", "Rrho", "Af", "vPert", "aVisc", "deltaD", "lam", "surf", "surfzone", "cellID")
TOL = {"xi": 1e-10, "rho": 1e-10, "lam": 1e-9, "v": 1e-8, "p": 1e-8, "acc": 1e-6, "Rrho": 1e-6, "vPert": 1e-6, "aVisc": 1e-6, "Af": 1e-6,
       "deltaD": 1e-6}


def relerr(a, b):
    s = np.abs(b).max()
    d = np.abs(a - b).max()
    return 0.0 if d == 0 else (d / s if s > 0 else np.inf)


def run(name, case, params, steps, rank, world, local_rank, mesh=None):
    n = case["xi"].shape[0]
    dx = case["params"]["particle_step"]
    xmin, xmax = case["xi"][:, 0].min() - 0.5 * dx, case["xi"][:, 0].max() + 0.5 * dx
    lo, hi = slab.slab_bounds(xmin, xmax, world)
    own = slab.partition(case["xi"], lo[rank], hi[rank])
    sub = {k: (v[own] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in case.items()}
    sub["bound_points"] = 0
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        e = slab.SlabEngine(eng.default_params(3, **params), sub, rank, world, lo[rank], hi[rank], device=local_rank,
                            stream=stream, capacity=n + 1000, part_id=own)
        if mesh is not None:
            e.upload_mesh(mesh)  # the aero mesh is replicated on every rank
        its = []
        for _ in range(steps):
            s = e.integrate()
            its.append((s.iterations, s.dt, s.npd, s.rms_error))
        got = e.download(FIELDS)
        stats = e.slab_stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (got, its, stats))
    ok = True
    if rank == 0:
        ref = eng.Engine(eng.default_params(3, **params), n, device=local_rank)
        ref.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
        if mesh is not None:
            ref.upload_mesh(mesh)
        rits = []
        for _ in range(steps):
            s = ref.integrate()
            rits.append((s.iterations, s.dt, s.npd, s.rms_error))
        want = ref.download(FIELDS)
        pid = np.concatenate([g[0]["part_id"] for g in gathered])
        n_left = want["part_id"].shape[0]  # the aero mesh erases the particles that escape it
        assert len(pid) == n_left and len(np.unique(pid)) == n_left and np.array_equal(np.sort(pid), np.sort(want["part_id"])), \
            "%s: particles lost or duplicated (%d of %d)" % (name, len(np.unique(pid)), n_left)
        if n_left != n:
            print("%s: %d of %d particles erased on both sides" % (name, n - n_left, n))
        order = np.argsort(want["part_id"])
        pid = order[np.searchsorted(want["part_id"][order], pid)]  # rows of the reference download
        print("%s: owned per rank %s, ghosts %s, exchanges %s (beside an interior sweep: %s), redecomps %s" % (
            name, [g[2]["n_owned"] for g in gathered], [g[2]["n_ghost"] for g in gathered],
            [g[2]["exchanges"] for g in gathered], [g[2]["overlapped"] for g in gathered],
            [g[2]["redecomps"] for g in gathered]))
        for r, g in enumerate(gathered):
            for a, b in zip(g[1], rits):
                if a[0] != b[0] or abs(a[1] - b[1]) > 1e-12 * b[1] or abs(a[2] - b[2]) > 1e-10 * abs(b[2]):
                    print("  rank %d step stats differ: %s vs %s" % (r, a, b))
                    ok = False
        for f in FIELDS[1:]:
            a = np.concatenate([g[0][f] for g in gathered])
            b = want[f][pid]
            if a.dtype.kind in "iu":
                bad = int((a != b).sum())
                print("  %-8s differing flags: %d" % (f, bad))
                ok &= bad == 0
            else:
                r = relerr(a, b)
                print("  %-8s relerr %.3e (tol %.0e)" % (f, r, TOL[f]))
                ok &= r <= TOL[f]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    # 1. quasi-static block: ghosts + forward exchanges, no migration
    case = cases.synthetic_block((24 * world, 12, 10), 1e-3, jitter=0.1, seed=3)
    ok &= run("block", case, dict(case["params"], delta_t_min=1e-9), 3, rank, world, local_rank)
    # 2. the same block drifting along +x at 30 m/s: re-decomposition with migration every other step
    case2 = cases.synthetic_block((24 * world, 12, 10), 1e-3, jitter=0.1, seed=4)
    case2["v"] = case2["v"] + np.array([30.0, 0.0, 0.0])
    ok &= run("drifting block", case2, dict(case2["params"], delta_t_min=1e-9), 6, rank, world, local_rank)
    # 3. Runge-Kutta
    ok &= run("block RK4", case, dict(case["params"], delta_t_min=1e-9, solver_type=1), 2, rank, world, local_rank)
    # 4. Gissler aero coupled to a replicated mesh that ends below the top of the block: the particles above it escape
    #    and are erased on whichever rank owns them (FindCell, Containment.cpp:735-777), together on all ranks
    Lx = 24 * world * 1e-3
    mesh = cases.hex_mesh((-2.1e-3, -2.2e-3, -2.3e-3), (Lx + 2.2e-3, 8.6e-3, 12.4e-3), (3 * world, 3, 4),
                          vel=lambda c: np.stack([20.0 + 1e3 * c[:, 1], 5.0 + 0 * c[:, 0], 1e3 * c[:, 0]], axis=1),
                          p=100000.0, rho=1.2)
    ok &= run("block in an aero mesh", case, dict(case["params"], delta_t_min=1e-9, acase=1, asource=1, lam_cutoff=1e9,
                                                  v_inf=(20.0, 5.0, 0.0), p_ref=100000.0, rho_g=1.2), 3, rank, world,
              local_rank, mesh=mesh)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB PARITY %s" % ("OK" if ok else "FAILED"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
