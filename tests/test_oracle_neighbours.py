"""Neighbour sets of the oracle vs brute force with the nanoflann expression, and vs scipy's KD-tree."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from fjsph_b200 import cases
from oracle import oracle as orc


def brute(xi, sr):
    n, dim = xi.shape
    out = []
    for i in range(n):
        d2 = np.zeros(n)
        for d in range(dim):  # accumulate x -> y -> z like metric_L2_Simple
            diff = xi[i, d] - xi[:, d]
            d2 = d2 + diff * diff
        j = np.nonzero(d2 < sr)[0]
        out.append((j, d2[j]))
    return out


@pytest.mark.parametrize("dim,jitter", [(3, "eps"), (3, 0.1), (2, "eps"), (2, 0.3), (3, None)])
def test_oracle_neighbours_bit_exact_vs_brute_force(dim, jitter):
    n = (9, 8, 7)[:dim] if dim == 3 else (17, 15)
    dx = 0.0015
    xi = cases.lattice(n, dx, jitter=jitter, seed=3)
    p = orc.default_params(dim, particle_step=dx)
    o = orc.Oracle(p)
    N = xi.shape[0]
    o.set_particles(xi, None, 1000.0, 0.0, 1.0, cases.FREE)
    o.update_neighbours()
    off, idx, d2 = o.neighbours()
    ref = brute(xi, p.sr)
    ties = 0
    for i in range(N):
        j, dd = ref[i]
        assert np.array_equal(idx[off[i]:off[i + 1]], j)
        assert np.array_equal(d2[off[i]:off[i + 1]], dd)  # bit-exact payload
        assert i in j  # self included
    # independent cross-check (inclusive <=, different arithmetic): sets may differ only at d ~ 2H ties
    tree = cKDTree(xi)
    r = np.sqrt(p.sr)
    for i in range(N):
        mine = set(idx[off[i]:off[i + 1]].tolist())
        theirs = set(tree.query_ball_point(xi[i], r))
        for j in mine ^ theirs:
            dist = np.linalg.norm(xi[i] - xi[j])
            assert abs(dist - r) < 1e-9 * r
            ties += 1
    if jitter in ("eps", None):
        assert ties > 0  # the lattice really exercises the d == 2H tie case


def test_neighbour_count_matches_n_full_in_the_bulk():
    dx = 0.01
    xi = cases.lattice((11, 11, 11), dx, jitter=None)
    p = orc.default_params(3, particle_step=dx)
    o = orc.Oracle(p)
    o.set_particles(xi, None, 1000.0, 0.0, 1.0, cases.FREE)
    o.update_neighbours()
    off, idx, d2 = o.neighbours()
    centre = 5 + 11 * (5 + 11 * 5)
    cnt = off[centre + 1] - off[centre]
    assert 251 <= cnt <= 257
