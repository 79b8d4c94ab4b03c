"""Inputs of the particle-tracker tests (tests/test_ipt_cpu.py, tests/test_gpu_ipt.py, tests/golden/make_ipt_vectors.py): a
sheared flow on a box mesh, hand-off records scattered over its upstream third, IPT settings.  The same function feeds the
oracle, the compiled reference and the device engine; only plain numbers and arrays leave it."""
import numpy as np

from fjsph_b200 import cases

# (dim, face kind, equation order, start pattern)
#   "scatter": anywhere in the cell, velocities in every direction -- in 3D most particles then leave a cell through the half
#              of a face MollerTrumbore does not accept (note Q9 of oracle/ipt_oracle.inc) and are failed, as in the reference
#   "lane":    in the corner of the cell cross-section nearest the faces' first vertex, moving with the stream -- trajectories
#              that cross the whole mesh
CASES = {
    "hex_tri_o1_scatter": (3, "tri", 1, "scatter"),
    "hex_tri_o2_lane": (3, "tri", 2, "lane"),
    "hex_quad_o1_lane": (3, "quad", 1, "lane"),
    "hex_quad_o2_scatter": (3, "quad", 2, "scatter"),
    "tet_o1_scatter": (3, "tet", 1, "scatter"),
    "tet_o2_scatter": (3, "tet", 2, "scatter"),
    "quad2d_o1_scatter": (2, "edge", 1, "scatter"),
    "quad2d_o2_long": (2, "edge", 2, "long"),
}
# The bound on one step (cells.maxlength, IPT.cpp:949, 1059) is set deliberately TIGHT here -- the mesh's longest edge /
# face diagonal itself, where the TAU readers leave 5 x (4 x in 2D) that -- so that the "moved too far in one step" failure
# is exercised: on the uniform 2D mesh every particle that crosses a whole cell at an angle trips it.  "long": scatter starts
# with four times the longest edge, the reference's own 2D value.
LENGTH_FACTOR = {"long": 4.0}
RECORD_CAP = 48


def build(name, n=160, seed=5):
    """-> dict(dim, mesh, start fields, settings as a plain dict)"""
    dim, kind, order, pattern = CASES[name]
    rng = np.random.default_rng(seed + sum(map(ord, name)))
    if dim == 3:
        lo, hi, cells = np.array([-0.1, -0.1, -0.1]), np.array([0.5, 0.1, 0.1]), (12, 5, 4)
        if pattern == "lane":
            vel = lambda c: np.stack([30 + 40 * c[:, 0] + 20 * c[:, 2], 0.4 * np.sin(9 * c[:, 0]) - 0.3, -0.5 + 1.5 * c[:, 1]], 1)
        else:
            vel = lambda c: np.stack([30 + 40 * c[:, 0] + 20 * c[:, 2], 5 * np.sin(9 * c[:, 0]) + 3 * c[:, 1], -4 + 10 * c[:, 1]], 1)
        rho = lambda c: 1.1 + c[:, 2] + 0.3 * c[:, 0]
        if kind == "tet":   # six tetrahedra per brick: triangles in general position, four faces per cell
            mesh, locate = cases.tet_mesh(lo, hi, cells, vel=vel, rho=rho)
        else:
            mesh = cases.hex_mesh(lo, hi, cells, vel=vel, rho=rho, triangulate=(kind == "tri"))
        grav = [0.0, 0.0, -9.81]
    else:
        lo, hi, cells = np.array([-0.1, -0.1]), np.array([0.5, 0.1]), (12, 5)
        vel = lambda c: np.stack([30 + 40 * c[:, 0] + 20 * c[:, 1], 5 * np.sin(9 * c[:, 0]) + 3 * c[:, 1]], 1)
        mesh = cases.quad_mesh(lo, hi, cells, vel=vel, rho=lambda c: 1.1 + c[:, 1] + 0.3 * c[:, 0])
        grav = [0.0, -9.81, 0.0]
    width = (hi - lo) / np.array(cells)
    ijk = np.stack([rng.integers(0, 4, size=n)] + [rng.integers(0, cells[d], size=n) for d in range(1, dim)], axis=1)
    if pattern == "lane":
        frac = np.concatenate([rng.uniform(0.05, 0.95, size=(n, 1)), rng.uniform(0.12, 0.38, size=(n, dim - 1))], axis=1)
        v = np.concatenate([rng.uniform(8.0, 14.0, size=(n, 1)), rng.normal(scale=0.05, size=(n, dim - 1))], axis=1)
    else:
        frac = rng.uniform(0.03, 0.97, size=(n, dim))
        v = rng.normal(scale=3.0, size=(n, dim)) + np.array([10.0, 0.0, 0.0])[:dim]
    x = lo + (ijk + frac) * width
    cid = ijk[:, 0] + cells[0] * ijk[:, 1] + (cells[0] * cells[1] * ijk[:, 2] if dim == 3 else 0)
    if kind == "tet":
        cid = locate(x)
    pad = lambda a: np.concatenate([a, np.zeros((n, 3 - dim))], axis=1)
    start = dict(part_id=np.arange(n, dtype=np.int64) + 100, cellID=cid.astype(np.int64), t=np.full(n, 0.25),
                 xi=pad(x), v=pad(v), cellV=pad(mesh["cVel"][cid]), cellRho=mesh["cRho"][cid].copy())
    settings = dict(eq_order=order, record=1, max_steps=4000, max_x=0.45, grav=grav)
    return dict(dim=dim, mesh=mesh, start=start, settings=settings, particle_step=1e-3,
                length_factor=LENGTH_FACTOR.get(pattern, 1.0))


def longest_edge(mesh, dim):
    """the longest edge of a triangle / longer diagonal of a quadrilateral (the edge length in 2D) over the faces of a mesh"""
    v, ptr, vtx = np.asarray(mesh["verts"], float), np.asarray(mesh["face_ptr"]), np.asarray(mesh["face_vtx"])
    d = lambda a, b: np.sqrt(((v[a] - v[b]) ** 2).sum(axis=-1))
    best = 0.0
    for k in np.unique(np.diff(ptr)):
        f = vtx[(ptr[:-1][np.diff(ptr) == k])[:, None] + np.arange(k)[None, :]]
        if dim == 2:
            e = d(f[:, 0], f[:, 1])
        elif k == 3:
            e = np.maximum(d(f[:, 0], f[:, 1]), np.maximum(d(f[:, 0], f[:, 2]), d(f[:, 1], f[:, 2])))
        else:
            e = np.maximum(d(f[:, 0], f[:, 2]), d(f[:, 1], f[:, 3]))
        best = max(best, float(e.max()))
    return best


def start_records(case, dtype, mass):
    rec = np.zeros(len(case["start"]["part_id"]), dtype=dtype)
    for k, a in case["start"].items():
        rec[k] = a
    rec["mass"] = mass
    return rec


FLOAT_FIELDS = ("t", "dt", "acc", "xi", "v", "cellV", "cellRho")
INT_FIELDS = ("part_id", "cellID", "faceID", "going", "failed")
