"""OpenFOAM ingestion (csrc/host_foam.cpp = FOAM::Read_FOAM, reference src/FOAMIO.cpp:346-955, ASCII cases): a small
hexahedral case written here in OpenFOAM's file layout is read back into the MESH arrays; the containment lookup of the
CPU oracle on it must give the analytic cells, and a run coupled to it must equal the run on the same mesh built
directly (fjsph_b200.cases.hex_mesh).  Host-only code: no GPU needed."""
import numpy as np
import pytest

from fjsph_b200 import _lib, cases, frontend
from oracle import oracle as orc

from tests.foam_case import write_case

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
    faces = tmp_path / "constant" / "polyMesh" / "faces"
    text = faces.read_text()
    faces.write_text(text.replace("ascii", "binary"))   # a binary face file is a faceCompactList (FOAMIO.cpp:826-836)
    with pytest.raises(_lib.FjsphError, match="should be \"faceCompactList\""):
        frontend.read_foam(tmp_path, "100")
    faces.write_text(text)
    pts.write_text(pts.read_text().replace("vectorField", "labelList"))
    with pytest.raises(_lib.FjsphError, match="should be"):
        frontend.read_foam(tmp_path, "100")


@pytest.mark.parametrize("label_bits,scalar_bits", [(32, 64), (64, 64), (32, 32), (64, 32)])
def test_binary_case_reads_like_the_ascii_one(tmp_path, label_bits, scalar_bits):
    """binary::Read_*_Data (FOAMIO.cpp:113-342): `format binary` files with the label and scalar widths of their own
    `arch` entry, faces as a faceCompactList.  With 64-bit scalars the mesh arrays equal those of the same case written in
    ASCII bit for bit; with 32-bit scalars they equal the float-rounded values."""
    vel = lambda c: (1.0 + c[0], 2.0 * c[1], 3.0)
    pr = lambda c: 1.0e5 + 10.0 * c[2]
    a_dir, b_dir = tmp_path / "ascii", tmp_path / "binary"
    a_dir.mkdir()
    b_dir.mkdir()
    write_case(a_dir, LO, HI, N, vel, pr, wall_patch=True)
    nc, U, P = write_case(b_dir, LO, HI, N, vel, pr, wall_patch=True, binary=True, label_bits=label_bits, scalar_bits=scalar_bits)
    a = frontend.read_foam(a_dir, "100", rho_fill=1.2262)
    b = frontend.read_foam(b_dir, "100", rho_fill=1.2262)
    for k in ("face_ptr", "face_vtx", "leftright", "cell_ptr", "cell_faces", "cRho"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(b["cVel"], U) and np.array_equal(b["cP"], P)
    if scalar_bits == 64:
        for k in ("verts", "cCentre", "cVel", "cP"):
            assert np.array_equal(a[k], b[k]), k
    else:
        assert np.array_equal(b["verts"], a["verts"].astype(np.float32).astype(np.float64))
        assert np.abs(b["cCentre"] - a["cCentre"]).max() < 1e-7
    # a truncated binary file is an error, not a short mesh
    owner = b_dir / "constant" / "polyMesh" / "owner"
    owner.write_bytes(owner.read_bytes()[:-40])
    with pytest.raises(_lib.FjsphError, match="binary list ends early"):
        frontend.read_foam(b_dir, "100")


def test_damaged_binary_size_is_an_error(tmp_path):
    """A binary points file claiming 10^15 points: the reader reports it (allocation failure or short list), the process
    lives -- nothing thrown inside the library crosses the C boundary."""
    write_case(tmp_path, LO, HI, (2, 2, 2), lambda c: (0.0, 0.0, 0.0), lambda c: 0.0, binary=True)
    pts = tmp_path / "constant" / "polyMesh" / "points"
    raw = pts.read_bytes()
    pts.write_bytes(raw.replace(b"\n27\n(", b"\n1000000000000000\n(", 1))
    with pytest.raises(_lib.FjsphError, match="foam_read"):
        frontend.read_foam(tmp_path, "100")
