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


def run(name, case, params, steps, rank, world, local_rank, mesh=None, block=None, bounds=None, dim=3):
    n = case["xi"].shape[0]
    dx = case["params"]["particle_step"]
    xmin, xmax = case["xi"][:, 0].min() - 0.5 * dx, case["xi"][:, 0].max() + 0.5 * dx
    lo, hi = bounds if bounds is not None else slab.slab_bounds(xmin, xmax, world)
    own = slab.partition(case["xi"], lo[rank], hi[rank])
    sub = {k: (v[own] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in case.items()}
    sub["bound_points"] = 0
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        e = slab.SlabEngine(eng.default_params(dim, **params), sub, rank, world, lo[rank], hi[rank], device=local_rank,
                            stream=stream, capacity=2 * n + 1000, part_id=own)
        if mesh is not None:
            e.upload_mesh(mesh)  # the aero mesh is replicated on every rank
        if block is not None:
            # every rank carries the block; its back / buffer tables (local indices) only where the buffer region lives
            loc = -np.ones(n, dtype=np.int64)
            loc[own] = np.arange(len(own))
            mine = dict(block, first=0, second=len(own))
            if (loc[block["back"]] >= 0).all() and (loc[block["buffer"]] >= 0).all():
                mine.update(back=loc[block["back"]], buffer=loc[block["buffer"]])
            else:
                assert (loc[block["back"]] < 0).all() and (loc[block["buffer"]] < 0).all(), "buffer region split over ranks"
                mine.pop("back"), mine.pop("buffer")
            e.set_blocks([mine])
        its = []
        for _ in range(steps):
            s = e.integrate()
            its.append((s.iterations, s.dt, s.npd, s.rms_error, s.n_add, s.n_del))
        got = e.download(FIELDS)
        stats = e.slab_stats()
    gathered = [None] * world
    dist.all_gather_object(gathered, (got, its, stats))
    ok = True
    if rank == 0:
        ref = eng.Engine(eng.default_params(dim, **params), 4 * n, device=local_rank)
        ref.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], 0)
        if mesh is not None:
            ref.upload_mesh(mesh)
        if block is not None:
            ref.set_blocks([block])
        rits = []
        for _ in range(steps):
            s = ref.integrate()
            rits.append((s.iterations, s.dt, s.npd, s.rms_error, s.n_add, s.n_del))
        want = ref.download(FIELDS)
        pid = np.concatenate([g[0]["part_id"] for g in gathered])
        n_left = want["part_id"].shape[0]  # the aero mesh erases the particles that escape it
        assert len(pid) == n_left and len(np.unique(pid)) == n_left and np.array_equal(np.sort(pid), np.sort(want["part_id"])), \
            "%s: particles lost or duplicated (%d of %d)" % (name, len(np.unique(pid)), n_left)
        if n_left != n or sum(r[4] + r[5] for r in rits):
            print("%s: %d particles at the start, %d inserted and %d erased on both sides, %d at the end" % (
                name, n, sum(r[4] for r in rits), n + sum(r[4] for r in rits) - n_left, n_left))
        order = np.argsort(want["part_id"])
        pid = order[np.searchsorted(want["part_id"][order], pid)]  # rows of the reference download
        print("%s: owned per rank %s, ghosts %s, exchanges %s (beside an interior sweep: %s), redecomps %s" % (
            name, [g[2]["n_owned"] for g in gathered], [g[2]["n_ghost"] for g in gathered],
            [g[2]["exchanges"] for g in gathered], [g[2]["overlapped"] for g in gathered],
            [g[2]["redecomps"] for g in gathered]))
        for r, g in enumerate(gathered):
            for a, b in zip(g[1], rits):
                if a[0] != b[0] or abs(a[1] - b[1]) > 1e-12 * b[1] or abs(a[2] - b[2]) > 1e-10 * abs(b[2]) or a[4:] != b[4:]:
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
    # 5. an inlet on rank 0 feeding a jet that crosses into the next slab and ends at a delete plane there: insertions
    #    (update_buffer_region) on one rank, migration of PIPE / FREE particles, erasures on another, ids unique over ranks
    if world == 2:
        jet = cases.inlet_jet(n=(5, 5, 8), fixed=1, delete_x=2.5, jitter=0.03)
        dxj = jet["params"]["particle_step"]
        ok &= run("inlet jet", jet, dict(jet["params"], delta_t_min=1e-9), 14, rank, world, local_rank, block=jet["block"],
                  bounds=([-1e300, -3.5 * dxj], [-3.5 * dxj, 1e300]))
    # 6. the 2D build: x-slabs of a 2D block (rows run along y, z = 0 throughout), with migration
    case6 = cases.synthetic_block_2d((48 * world, 80), 1e-3, jitter=0.1, seed=6)
    case6["v"] = case6["v"] + np.array([20.0, 0.0])
    ok &= run("2D drifting block", case6, dict(case6["params"], delta_t_min=1e-9), 5, rank, world, local_rank, dim=2)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB PARITY %s" % ("OK" if ok else "FAILED"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        # a failed check on one rank must not leave the others waiting in a collective until the watchdog fires:
        # leave at once, without the interpreter's teardown (which joins NCCL), and let the launcher stop the rest
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
