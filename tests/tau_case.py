"""Writes a small hexahedral box as a TAU face-based mesh file (the layout FJSPH's Cell2Face converter produces and
TAU::Read_tau_mesh_FACE reads, CDFIO.cpp:1228-1356) and a TAU solution file, both NetCDF-3 classic, with scipy."""
import numpy as np
from scipy.io import netcdf_file


def write_tau(root, lo, hi, n, vel, p, rho, version=2, wall_marker=-1, outer_marker=-2, float_solution=False, split="x"):
    """Box [lo, hi] of n = (nx, ny, nz) hexahedra.  The x-normal faces (split="x"; the y-normal ones with split="y") are
    split into two triangles each (so the file has triangles AND quadrilaterals; triangles come first in the face
    numbering), the others stay quadrilaterals.  Cell ids
    are (k*ny + j)*nx + i.  right_element_of_faces of a boundary face is wall_marker on zmin, outer_marker elsewhere.
    Returns (mesh path, solution path, points, list of faces as vertex tuples, left, right)."""
    nx, ny, nz = n
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate(n)]
    vid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    cid = lambda i, j, k: (k * ny + j) * nx + i
    pts = np.array([(xs[0][i], xs[1][j], xs[2][k]) for k in range(nz + 1) for j in range(ny + 1) for i in range(nx + 1)])
    tris, quads = [], []          # (vertices, left, right)
    for k in range(nz):
        for j in range(ny):
            for i in range(nx + 1):
                q = (vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1))
                if i == 0:
                    l, r = cid(0, j, k), outer_marker
                elif i == nx:
                    l, r = cid(nx - 1, j, k), outer_marker
                else:
                    l, r = cid(i - 1, j, k), cid(i, j, k)
                if split == "x":
                    tris.append(((q[0], q[1], q[2]), l, r))
                    tris.append(((q[0], q[2], q[3]), l, r))
                else:
                    quads.append((q, l, r))
    for k in range(nz):
        for j in range(ny + 1):
            for i in range(nx):
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1))
                if j == 0:
                    l, r = cid(i, 0, k), outer_marker
                elif j == ny:
                    l, r = cid(i, ny - 1, k), outer_marker
                else:
                    l, r = cid(i, j - 1, k), cid(i, j, k)
                if split == "y":
                    tris.append(((q[0], q[1], q[2]), l, r))
                    tris.append(((q[0], q[2], q[3]), l, r))
                else:
                    quads.append((q, l, r))
    for k in range(nz + 1):
        for j in range(ny):
            for i in range(nx):
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k))
                if k == 0:
                    l, r = cid(i, j, 0), wall_marker
                elif k == nz:
                    l, r = cid(i, j, nz - 1), outer_marker
                else:
                    l, r = cid(i, j, k - 1), cid(i, j, k)
                quads.append((q, l, r))
    faces = tris + quads
    left = np.array([f[1] for f in faces], dtype=np.int32)
    right = np.array([f[2] for f in faces], dtype=np.int32)
    n_surf = int((right < 0).sum())
    mesh_path, sol_path = str(root / "box.grid.faces"), str(root / "box.pval")
    with netcdf_file(mesh_path, "w", version=version) as f:
        f.history = "tests/tau_case.py"
        for name, size in (("no_of_elements", nx * ny * nz), ("no_of_faces", len(faces)), ("no_of_points", len(pts)),
                           ("no_of_surfaceelements", n_surf), ("no_of_triangles", len(tris)), ("points_per_triangle", 3),
                           ("no_of_quadrilaterals", len(quads)), ("points_per_quadrilateral", 4)):
            f.createDimension(name, size)
        v = f.createVariable("points_of_triangles", "i4", ("no_of_triangles", "points_per_triangle"))
        v[:] = np.array([t[0] for t in tris], dtype=np.int32)
        v = f.createVariable("points_of_quadrilaterals", "i4", ("no_of_quadrilaterals", "points_per_quadrilateral"))
        v[:] = np.array([q[0] for q in quads], dtype=np.int32)
        for d, name in enumerate(("points_xc", "points_yc", "points_zc")):
            v = f.createVariable(name, "f8", ("no_of_points",))
            v.units = "m"
            v[:] = pts[:, d]
        f.createVariable("left_element_of_faces", "i4", ("no_of_faces",))[:] = left
        f.createVariable("right_element_of_faces", "i4", ("no_of_faces",))[:] = right
        f.createVariable("boundarymarker_of_surfaces", "i4", ("no_of_surfaceelements",))[:] = np.arange(n_surf, dtype=np.int32) % 6 + 1
    U = np.array([vel(x) for x in pts])
    with netcdf_file(sol_path, "w", version=version) as f:
        f.createDimension("no_of_points", len(pts))
        kind = "f4" if float_solution else "f8"
        for name, data in (("density", [rho(x) for x in pts]), ("x_velocity", U[:, 0]), ("y_velocity", U[:, 1]),
                           ("z_velocity", U[:, 2]), ("pressure", [p(x) for x in pts])):
            f.createVariable(name, kind, ("no_of_points",))[:] = np.asarray(data, dtype=kind)
    return mesh_path, sol_path, pts, [f[0] for f in faces], left, right


def write_tau_edge(root, lo, hi, n, vel, p, rho, plane="xz", version=2, wall_marker=-1, outer_marker=-2, split_surface_dims=False):
    """A 2D TAU case as FJSPH's Cell2Edge leaves it: rectangle [lo, hi] of n = (nx, ny) quadrilateral cells stored by
    EDGES (points_of_element_edges, left / right_element_of_edges), the two in-plane coordinates under the names of `plane`
    ("xz": points_xc + points_zc, the y coordinate absent), and a solution file over the TWO-layer 3D point set the flow
    solver ran on (2 x the mesh's points) which `vertices_in_use` indexes.  The boundary edges on the lower side carry
    wall_marker, the others outer_marker.  split_surface_dims: the boundary-edge count as "no_of_wall_edges" +
    "no_of_farfield_edges" (the layout of the reference's own Examples/RAE2822) instead of "no_of_surfaceelements".  Returns (mesh path, solution path, points [n,2], edges, left, right, used)."""
    nx, ny = n
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate(n)]
    vid = lambda i, j: j * (nx + 1) + i
    cid = lambda i, j: j * nx + i
    pts = np.array([(xs[0][i], xs[1][j]) for j in range(ny + 1) for i in range(nx + 1)])
    edges = []
    for j in range(ny):
        for i in range(nx + 1):
            e = (vid(i, j), vid(i, j + 1))
            if i == 0:
                edges.append((e, cid(0, j), outer_marker))
            elif i == nx:
                edges.append((e, cid(nx - 1, j), outer_marker))
            else:
                edges.append((e, cid(i - 1, j), cid(i, j)))
    for j in range(ny + 1):
        for i in range(nx):
            e = (vid(i + 1, j), vid(i, j))
            if j == 0:
                edges.append((e, cid(i, 0), wall_marker))
            elif j == ny:
                edges.append((e, cid(i, ny - 1), outer_marker))
            else:
                edges.append((e, cid(i, j - 1), cid(i, j)))
    left = np.array([e[1] for e in edges], dtype=np.int32)
    right = np.array([e[2] for e in edges], dtype=np.int32)
    n_surf = int((right < 0).sum())
    npt = len(pts)
    rng = np.random.default_rng(17)
    used = rng.permutation(2 * npt)[:npt].astype(np.int32)  # where each mesh point sits in the solver's point set
    mesh_path, sol_path = str(root / "plate.grid.edges"), str(root / "plate.pval")
    names = {"x": "points_xc", "y": "points_yc", "z": "points_zc"}
    with netcdf_file(mesh_path, "w", version=version) as f:
        f.history = "tests/tau_case.py"
        n_wall = int((right == wall_marker).sum())
        surf_dims = ((("no_of_wall_edges", n_wall), ("no_of_farfield_edges", n_surf - n_wall)) if split_surface_dims
                     else (("no_of_surfaceelements", n_surf),))
        for name, size in (("no_of_elements", nx * ny), ("no_of_edges", len(edges)), ("points_per_edge", 2)) + surf_dims + (
                ("no_of_points", npt),):
            f.createDimension(name, size)
        f.createVariable("points_of_element_edges", "i4", ("no_of_edges", "points_per_edge"))[:] = np.array(
            [e[0] for e in edges], dtype=np.int32)
        f.createVariable("vertices_in_use", "i4", ("no_of_points",))[:] = used
        for d, axis in enumerate(plane):
            v = f.createVariable(names[axis], "f8", ("no_of_points",))
            v.units = "m"
            v[:] = pts[:, d]
        f.createVariable("left_element_of_edges", "i4", ("no_of_edges",))[:] = left
        f.createVariable("right_element_of_edges", "i4", ("no_of_edges",))[:] = right
        if not split_surface_dims:
            f.createVariable("boundarymarker_of_surfaces", "i4", ("no_of_surfaceelements",))[:] = np.arange(n_surf, dtype=np.int32) % 4 + 1
    # the solver's point set: the mesh points scattered over 2 n slots (the other slots: the second layer, other values)
    sol = {k: rng.normal(size=2 * npt) for k in ("density", "x_velocity", "y_velocity", "z_velocity", "pressure")}
    U = np.array([vel(x) for x in pts])
    comp = {"x": "x_velocity", "y": "y_velocity", "z": "z_velocity"}
    sol[comp[plane[0]]][used] = U[:, 0]
    sol[comp[plane[1]]][used] = U[:, 1]
    sol["density"][used] = [rho(x) for x in pts]
    sol["pressure"][used] = [p(x) for x in pts]
    with netcdf_file(sol_path, "w", version=version) as f:
        f.createDimension("no_of_points", 2 * npt)
        for name, data in sol.items():
            f.createVariable(name, "f8", ("no_of_points",))[:] = data
    return mesh_path, sol_path, pts, [e[0] for e in edges], left, right, used
