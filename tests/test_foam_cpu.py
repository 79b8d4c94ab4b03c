"""OpenFOAM ingestion (csrc/host_foam.cpp = FOAM::Read_FOAM, reference src/FOAMIO.cpp:346-955, ASCII cases): a small
hexahedral case written here in OpenFOAM's file layout is read back into the MESH arrays; the containment lookup of the
CPU oracle on it must give the analytic cells, and a run coupled to it must equal the run on the same mesh built
directly (fjsph_b200.cases.hex_mesh).  Host-only code: no GPU needed."""
import numpy as np
import pytest

from fjsph_b200 import _lib, cases, frontend
from oracle import oracle as orc

HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
  =========                 |
  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox
\\*---------------------------------------------------------------------------*/
FoamFile
{
    version     2.0;
    format      ascii;
    class       %s;
    location    "%s";
    object      %s;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def write_case(root, lo, hi, n, vel, p, wall_patch=False):
    """Box [lo, hi] of n = (nx, ny, nz) hexahedra as an OpenFOAM case: internal faces first (owner < neighbour), then
    the patches xmin, xmax, ymin, ymax, zmin, zmax.  Cell ids are (k*ny + j)*nx + i like cases.hex_mesh."""
    nx, ny, nz = n
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate(n)]
    vid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    cid = lambda i, j, k: (k * ny + j) * nx + i
    pts = [(xs[0][i], xs[1][j], xs[2][k]) for k in range(nz + 1) for j in range(ny + 1) for i in range(nx + 1)]
    internal, patches = [], {nm: [] for nm in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
    for k in range(nz):
        for j in range(ny):
            for i in range(nx + 1):
                q = (vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1))
                if i == 0:
                    patches["xmin"].append((q, cid(0, j, k)))
                elif i == nx:
                    patches["xmax"].append((q, cid(nx - 1, j, k)))
                else:
                    internal.append((q, cid(i - 1, j, k), cid(i, j, k)))
    for k in range(nz):
        for j in range(ny + 1):
            for i in range(nx):
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1))
                if j == 0:
                    patches["ymin"].append((q, cid(i, 0, k)))
                elif j == ny:
                    patches["ymax"].append((q, cid(i, ny - 1, k)))
                else:
                    internal.append((q, cid(i, j - 1, k), cid(i, j, k)))
    for k in range(nz + 1):
        for j in range(ny):
            for i in range(nx):
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k))
                if k == 0:
                    patches["zmin"].append((q, cid(i, j, 0)))
                elif k == nz:
                    patches["zmax"].append((q, cid(i, j, nz - 1)))
                else:
                    internal.append((q, cid(i, j, k - 1), cid(i, j, k)))
    internal.sort(key=lambda f: (f[1], f[2]))
    faces = [f[0] for f in internal]
    owner = [f[1] for f in internal]
    neigh = [f[2] for f in internal]
    blines, start = [], len(faces)
    for nm, fl in patches.items():
        kind = "wall" if (wall_patch and nm == "zmin") else "patch"
        blines.append("    %s\n    {\n        type            %s;\n        nFaces          %d;\n        startFace       %d;\n    }\n"
                      % (nm, kind, len(fl), start))
        start += len(fl)
        faces += [f[0] for f in fl]
        owner += [f[1] for f in fl]
    poly = root / "constant" / "polyMesh"
    poly.mkdir(parents=True)
    sol = root / "100"
    sol.mkdir()
    lst = lambda items: "%d\n(\n%s\n)\n" % (len(items), "\n".join(items))
    (poly / "points").write_text(HEADER % ("vectorField", "constant/polyMesh", "points")
                                 + lst(["(%.17g %.17g %.17g)" % q for q in pts]))
    (poly / "faces").write_text(HEADER % ("faceList", "constant/polyMesh", "faces")
                                + lst(["4(%d %d %d %d)" % f for f in faces]))
    (poly / "owner").write_text(HEADER % ("labelList", "constant/polyMesh", "owner") + lst(["%d" % o for o in owner]))
    (poly / "neighbour").write_text(HEADER % ("labelList", "constant/polyMesh", "neighbour") + lst(["%d" % o for o in neigh]))
    (poly / "boundary").write_text(HEADER % ("polyBoundaryMesh", "constant/polyMesh", "boundary")
                                   + "%d\n(\n%s)\n" % (len(patches), "".join(blines)))
    nc = nx * ny * nz
    centres = np.array([((xs[0][i] + xs[0][i + 1]) / 2, (xs[1][j] + xs[1][j + 1]) / 2, (xs[2][k] + xs[2][k + 1]) / 2)
                        for k in range(nz) for j in range(ny) for i in range(nx)])
    U = np.array([vel(c) for c in centres])
    P = np.array([p(c) for c in centres])
    (sol / "U").write_text(HEADER % ("volVectorField", "100", "U") + "dimensions      [0 1 -1 0 0 0 0];\n\n"
                           "internalField   nonuniform List<vector>\n" + lst(["(%.17g %.17g %.17g)" % tuple(u) for u in U])
                           + ";\n\nboundaryField\n{\n}\n")
    (sol / "p").write_text(HEADER % ("volScalarField", "100", "p") + "dimensions      [1 -1 -2 0 0 0 0];\n\n"
                           "internalField   nonuniform List<scalar>\n" + lst(["%.17g" % v for v in P])
                           + ";\n\nboundaryField\n{\n}\n")
    return nc, U, P


LO, HI, N = np.array([-0.1013, -0.1007, -0.1011]), np.array([0.1009, 0.1003, 0.1017]), (6, 7, 5)


def test_polymesh_round_trip(tmp_path):
    vel = lambda c: (1.0 + c[0], 2.0 * c[1], 3.0)
    pr = lambda c: 1.0e5 + 10.0 * c[2]
    nc, U, P = write_case(tmp_path, LO, HI, N, vel, pr, wall_patch=True)
    m = frontend.read_foam(tmp_path, "100", rho_fill=1.2262)
    nx, ny, nz = N
    assert m["verts"].shape == ((nx + 1) * (ny + 1) * (nz + 1), 3) and m["cCentre"].shape == (nc, 3)
    n_quads = (nx + 1) * ny * nz + nx * (ny + 1) * nz + nx * ny * (nz + 1)
    assert m["leftright"].shape == (2 * n_quads, 2)                     # every quad fanned into two triangles
    assert np.array_equal(m["face_ptr"], 3 * np.arange(2 * n_quads + 1))
    assert np.array_equal(np.diff(m["cell_ptr"]), np.full(nc, 12))      # 6 quads = 12 triangles per hexahedron
    assert np.array_equal(m["cVel"], U) and np.array_equal(m["cP"], P) and np.all(m["cRho"] == 1.2262)
    # boundary markers: -1 on the wall patch (zmin), -2 on every other patch, a cell id on internal faces
    lr = m["leftright"]
    assert (lr[:, 1] == -1).sum() == 2 * nx * ny and (lr[:, 1] == -2).sum() == 2 * (2 * ny * nz + 2 * nx * nz + nx * ny)
    assert (lr[:, 0] >= 0).all() and (lr[lr[:, 1] >= 0, 0] < lr[lr[:, 1] >= 0, 1]).all()
    # fan order (0, j+1, j+2): the two triangles of a quad share its first vertex and the diagonal
    tri = m["face_vtx"].reshape(-1, 3)
    assert np.array_equal(tri[0::2, 0], tri[1::2, 0]) and np.array_equal(tri[0::2, 2], tri[1::2, 1])
    # the reference's cell "centre": mean over the sorted vertex list after std::unique WITHOUT erase (FOAMIO.cpp:652-664)
    for c in (0, nc // 2, nc - 1):
        fl = m["cell_faces"][m["cell_ptr"][c]: m["cell_ptr"][c + 1]]
        v = np.sort(tri[fl].ravel())
        uniq = np.unique(v)
        after_unique = np.concatenate([uniq, v[len(uniq):]])
        assert np.allclose(m["cCentre"][c], m["verts"][after_unique].mean(axis=0), rtol=1e-14)
    true_centre = m["verts"][np.unique(tri[m["cell_faces"][:12]].ravel())].mean(axis=0)
    assert np.abs(m["cCentre"][0] - true_centre).max() < 0.3 * (HI - LO).max() / min(N)  # a seed near the cell, not in it


def test_containment_and_coupled_run_on_a_foam_mesh(tmp_path):
    write_case(tmp_path, LO, HI, N, lambda c: (0.0, 21.55, 0.0), lambda c: 100000.0)
    foam = frontend.read_foam(tmp_path, "100", rho_fill=1.1025)
    direct = cases.hex_mesh(LO, HI, N, vel=(0.0, 21.55, 0.0), p=100000.0, rho=1.1025)
    case = cases.droplet(dx=0.0125, jitter=0.05)
    # 1. containment: every FREE particle lands in its analytic cell
    o = orc.Oracle(orc.default_params(3, asource=1, **dict(case["params"], lam_cutoff=1e9)))
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
    o.set_mesh(foam)
    o.update_neighbours()
    o.prestep()
    o.aero_velocity()
    ijk = np.floor((case["xi"] - LO) / ((HI - LO) / np.array(N))).astype(int)
    assert np.array_equal(o.get("cellID"), (ijk[:, 2] * N[1] + ijk[:, 1]) * N[0] + ijk[:, 0])
    # 2. two coupled steps: the OpenFOAM-read mesh and the directly built one give the same run, bit for bit
    runs = []
    for mesh in (foam, direct):
        om = orc.Oracle(orc.default_params(3, asource=1, **dict(case["params"], delta_t_min=1e-9)))
        om.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
        om.set_mesh(mesh)
        for _ in range(2):
            om.integrate()
        runs.append({f: om.get(f) for f in ("xi", "v", "rho", "Af", "cellID")})
    for f in runs[0]:
        assert np.array_equal(runs[0][f], runs[1][f]), f
    assert np.abs(runs[0]["Af"]).max() > 1.0


def test_foam_errors(tmp_path):
    with pytest.raises(_lib.FjsphError, match="boundary"):
        frontend.read_foam(tmp_path / "nowhere")
    write_case(tmp_path, LO, HI, (2, 2, 2), lambda c: (0.0, 0.0, 0.0), lambda c: 0.0)
    pts = tmp_path / "constant" / "polyMesh" / "points"
    pts.write_text(pts.read_text().replace("ascii", "binary"))
    with pytest.raises(_lib.FjsphError, match="binary"):
        frontend.read_foam(tmp_path, "100")
    pts.write_text(pts.read_text().replace("binary", "ascii").replace("vectorField", "labelList"))
    with pytest.raises(_lib.FjsphError, match="should be"):
        frontend.read_foam(tmp_path, "100")
