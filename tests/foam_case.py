"""Writes a small hexahedral box as an OpenFOAM case, ASCII or binary (test input for the OpenFOAM reader)."""
import numpy as np

HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
  =========                 |
  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox
\\*---------------------------------------------------------------------------*/
FoamFile
{
    version     2.0;
    format      %s;
    class       %s;
    location    "%s";
    object      %s;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def write_case(root, lo, hi, n, vel, p, wall_patch=False, binary=False, label_bits=32, scalar_bits=64):
    """Box [lo, hi] of n = (nx, ny, nz) hexahedra as an OpenFOAM case: internal faces first (owner < neighbour), then
    the patches xmin, xmax, ymin, ymax, zmin, zmax.  Cell ids are (k*ny + j)*nx + i like cases.hex_mesh.
    binary=True writes what OpenFOAM's `writeFormat binary` does: the same dictionaries with `format binary` and an
    `arch "LSB;label=..;scalar=.."` entry, raw little-endian lists between the brackets, faces as a faceCompactList
    (offsets, then the vertex labels); the boundary file stays ASCII."""
    ldt, sdt = np.dtype("<i%d" % (label_bits // 8)), np.dtype("<f%d" % (scalar_bits // 8))

    def head(cls, loc, obj, force_ascii=False):
        if binary and not force_ascii:
            return HEADER % ('binary;\n    arch        "LSB;label=%d;scalar=%d"' % (label_bits, scalar_bits), cls, loc, obj)
        return HEADER % ("ascii", cls, loc, obj)

    def raw(count, data):
        return b"%d\n(" % count + data + b")\n"
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
    (poly / "boundary").write_text(head("polyBoundaryMesh", "constant/polyMesh", "boundary", force_ascii=True)
                                   + "%d\n(\n%s)\n" % (len(patches), "".join(blines)))
    nc = nx * ny * nz
    centres = np.array([((xs[0][i] + xs[0][i + 1]) / 2, (xs[1][j] + xs[1][j + 1]) / 2, (xs[2][k] + xs[2][k + 1]) / 2)
                        for k in range(nz) for j in range(ny) for i in range(nx)])
    U = np.array([vel(c) for c in centres])
    P = np.array([p(c) for c in centres])
    dimU, dimP = "dimensions      [0 1 -1 0 0 0 0];\n\n", "dimensions      [1 -1 -2 0 0 0 0];\n\n"
    tail = ";\n\nboundaryField\n{\n}\n"
    if binary:
        enc = lambda text: text.encode("ascii")
        (poly / "points").write_bytes(enc(head("vectorField", "constant/polyMesh", "points"))
                                      + raw(len(pts), np.asarray(pts, dtype=sdt).tobytes()))
        offsets = np.arange(len(faces) + 1, dtype=ldt) * 4
        (poly / "faces").write_bytes(enc(head("faceCompactList", "constant/polyMesh", "faces"))
                                     + raw(len(offsets), offsets.tobytes()) + b"\n"
                                     + raw(4 * len(faces), np.asarray(faces, dtype=ldt).tobytes()))
        (poly / "owner").write_bytes(enc(head("labelList", "constant/polyMesh", "owner"))
                                     + raw(len(owner), np.asarray(owner, dtype=ldt).tobytes()))
        (poly / "neighbour").write_bytes(enc(head("labelList", "constant/polyMesh", "neighbour"))
                                         + raw(len(neigh), np.asarray(neigh, dtype=ldt).tobytes()))
        (sol / "U").write_bytes(enc(head("volVectorField", "100", "U") + dimU + "internalField   nonuniform List<vector> \n")
                                + raw(nc, U.astype(sdt).tobytes()) + enc(tail))
        (sol / "p").write_bytes(enc(head("volScalarField", "100", "p") + dimP + "internalField   nonuniform List<scalar> \n")
                                + raw(nc, P.astype(sdt).tobytes()) + enc(tail))
        return nc, U.astype(sdt).astype(np.float64), P.astype(sdt).astype(np.float64)
    (poly / "points").write_text(head("vectorField", "constant/polyMesh", "points")
                                 + lst(["(%.17g %.17g %.17g)" % q for q in pts]))
    (poly / "faces").write_text(head("faceList", "constant/polyMesh", "faces")
                                + lst(["4(%d %d %d %d)" % f for f in faces]))
    (poly / "owner").write_text(head("labelList", "constant/polyMesh", "owner") + lst(["%d" % o for o in owner]))
    (poly / "neighbour").write_text(head("labelList", "constant/polyMesh", "neighbour") + lst(["%d" % o for o in neigh]))
    (sol / "U").write_text(head("volVectorField", "100", "U") + dimU
                           + "internalField   nonuniform List<vector>\n" + lst(["(%.17g %.17g %.17g)" % tuple(u) for u in U])
                           + tail)
    (sol / "p").write_text(head("volScalarField", "100", "p") + dimP
                           + "internalField   nonuniform List<scalar>\n" + lst(["%.17g" % v for v in P])
                           + tail)
    return nc, U, P
