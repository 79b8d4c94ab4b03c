"""Model behind DESIGN.md section 4, "why the sweeps do not stage neighbour records in shared memory": for a jittered
lattice (the block workload: spacing dx, support 2H = 4 dx) and a CTA that owns a box of a x b x c particles, count the
distinct particles its neighbour lists touch (the records a shared-memory tile would have to hold), the reuse per
staged record, and the union of the lists of one warp's 32 particles (what a broadcast-style sweep would evaluate).
CPU only (scipy), no engine involved:  python tools/tile_halo_model.py"""
import numpy as np
from scipy.spatial import cKDTree


def main():
    rng = np.random.default_rng(1234)
    n = (48, 40, 32)
    g = np.stack(np.meshgrid(*[np.arange(k) for k in n], indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float64)
    x = g + rng.uniform(-0.1, 0.1, g.shape)
    tree = cKDTree(x)
    lo = np.array([16, 16, 12])
    print("CTA box      particles  distinct neighbours  KB per 32-byte record  reuse per staged record")
    for box in ((32, 4, 1), (35, 8, 1), (16, 4, 4), (8, 8, 4), (8, 8, 8), (16, 8, 4)):
        sel = np.all((g >= lo) & (g < lo + np.array(box)), axis=1)
        lists = tree.query_ball_point(x[sel], 4.0)
        pairs = sum(len(l) for l in lists)
        distinct = len(set().union(*map(set, lists)))
        print("%-12s %9d  %19d  %21.0f  %23.1f" % ("x".join(map(str, box)), sel.sum(), distinct, distinct * 32 / 1024, pairs / distinct))
    print("\nwarp shape   mean list  union of the warp's lists  union / list")
    for box in ((32, 1, 1), (8, 4, 1), (4, 4, 2)):
        sel = np.all((g >= lo) & (g < lo + np.array(box)), axis=1)
        lists = tree.query_ball_point(x[sel], 4.0)
        mean = np.mean([len(l) for l in lists])
        union = len(set().union(*map(set, lists)))
        print("%-12s %9.1f  %25d  %12.2f" % ("x".join(map(str, box)), mean, union, union / mean))


if __name__ == "__main__":
    main()
