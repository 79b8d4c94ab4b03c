"""Input decks of the BASELINE.json configs as plain numpy arrays (host-side input generation only).

These restate just enough of FJSPH's block generators to reproduce the configs:
  * square/cube lattice  -- shapes/square.cpp:101-145 (grid order; one jitter value per point, all axes)
  * circle/sphere lattice -- shapes/circle.cpp:131-189 (grid order; per-axis jitter; |x-c|^2 > R^2 culled)
The reference perturbs points with std::default_random_engine in U(0, eps*dx); that stream is not
replicated bit-for-bit here (numpy PCG64 with a fixed seed is used instead) -- full-fidelity shape
generation is SURVEY 8f row N1.
"""
from __future__ import annotations

import numpy as np

BOUND, PISTON, BUFFER, BACK, PIPE, FREE, OUTLET, LOST = range(8)
EPS = float(np.finfo(np.float64).eps)


def cole_pressure(rho, rho0, c, gam=7.0, p_back=0.0):
    """FLUID::get_pressure, Cole EOS (Var.h:203-218)."""
    B = rho0 * c * c / gam
    return B * ((rho / rho0) ** gam - 1.0) + p_back


def cole_density(p, rho0, c, gam=7.0, p_back=0.0):
    """FLUID::get_density, Cole EOS (Var.h:221-236)."""
    B = rho0 * c * c / gam
    return rho0 * (((p - p_back) / B) + 1.0) ** (1.0 / gam)


def lattice(n, dx, start=(0.0, 0.0, 0.0), jitter="eps", seed=1234):
    """n = (ni, nj[, nk]) lattice, x fastest (square.cpp:107-141).  jitter: 'eps' -> U(0, eps*dx) (the
    reference's tie-stress perturbation), float f -> U(-f, f)*dx, None -> none."""
    n = tuple(int(k) for k in n)
    dim = len(n)
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(k, dtype=np.float64) for k in n[::-1]], indexing="ij")
    pts = np.stack([g.reshape(-1) for g in grids[::-1]], axis=1) * dx
    if jitter == "eps":
        pts = pts + rng.uniform(0.0, EPS * dx, size=(pts.shape[0], 1))
    elif jitter is not None:
        pts = pts + rng.uniform(-float(jitter), float(jitter), size=pts.shape) * dx
    return pts + np.asarray(start, dtype=np.float64)[:dim]


def synthetic_block(n=(500, 250, 100), dx=1e-3, jitter=0.1, seed=1234, rho0=1000.0, c=100.0, x_offset_cells=0):
    """Config C5: FREE-particle block with a smooth density and velocity field.
    x_offset_cells shifts the slab along x (rank r of a slab decomposition owns cells [r*ni,(r+1)*ni))."""
    ni, nj, nk = n
    xi = lattice(n, dx, start=(x_offset_cells * dx, 0.0, 0.0), jitter=jitter, seed=seed + x_offset_cells)
    L = np.array([ni * dx, nj * dx, nk * dx])
    rho = rho0 * (1.0 + 1e-3 * np.sin(2 * np.pi * xi[:, 0] / L[0]))
    v = 0.5 * np.stack(
        [np.sin(2 * np.pi * xi[:, 1] / L[1]), np.sin(2 * np.pi * xi[:, 2] / L[2]), np.sin(2 * np.pi * xi[:, 0] / L[0])],
        axis=1,
    )
    N = xi.shape[0]
    return dict(
        xi=xi, v=v, rho=rho, p=cole_pressure(rho, rho0, c), m=np.full(N, rho0 * dx**3),
        b=np.full(N, FREE, dtype=np.int32), bound_points=0,
        params=dict(particle_step=dx, rho_rest=rho0, speed_sound=c, mu=8.94e-4, sig=0.0708, visc_alpha=0.05,
                    dsph_delta=0.1, grav=(0.0, 0.0, -9.81)),
    )


def synthetic_block_2d(n=(40, 28), dx=1e-3, jitter=0.1, seed=21, rho0=1000.0, c=100.0):
    """The 2D cut of the C5 block: FREE particles on a jittered lattice with smooth density and velocity fields."""
    xi = lattice(n, dx, start=(0.0, 0.0), jitter=jitter, seed=seed)
    L = np.array([n[0] * dx, n[1] * dx])
    rho = rho0 * (1.0 + 1e-3 * np.sin(2 * np.pi * xi[:, 0] / L[0]))
    v = 0.5 * np.stack([np.sin(2 * np.pi * xi[:, 1] / L[1]), np.sin(2 * np.pi * xi[:, 0] / L[0])], axis=1)
    N = xi.shape[0]
    return dict(xi=xi, v=v, rho=rho, p=cole_pressure(rho, rho0, c), m=np.full(N, rho0 * dx**2),
                b=np.full(N, FREE, dtype=np.int32), bound_points=0,
                params=dict(particle_step=dx, rho_rest=rho0, speed_sound=c, mu=8.94e-4, sig=0.0708, visc_alpha=0.05,
                            dsph_delta=0.1, grav=(0.0, -9.81, 0.0)))


def synthetic_jet(nx=255, radius_cells=125, dx=1e-3, jitter=0.1, seed=1234, rho0=1000.0, c=100.0, x_offset_cells=0,
                  v_jet=30.0, v_inf=(0.0, 100.0, 0.0)):
    """Config C5, "jet" flavour: a liquid cylinder of radius radius_cells*dx along x moving at
    v = (v_jet, 0, 0) in a cross flow v_inf with the Gissler aero model on.  nx lattice columns per slab
    (nx = 255 -> 12.5 M particles at R = 125 dx); x_offset_cells shifts the slab as in synthetic_block."""
    side = 2 * int(radius_cells) + 1
    pts = lattice((nx, side, side), dx, start=(x_offset_cells * dx, -radius_cells * dx, -radius_cells * dx), jitter=None)
    keep = (pts[:, 1] ** 2 + pts[:, 2] ** 2) <= (radius_cells * dx) ** 2 * (1.0 + 1e-12)
    pts = pts[keep]
    rng = np.random.default_rng(seed + x_offset_cells)
    xi = np.ascontiguousarray(pts + rng.uniform(-float(jitter), float(jitter), size=pts.shape) * dx)
    N = xi.shape[0]
    Lx = nx * dx
    rho = rho0 * (1.0 + 1e-3 * np.sin(2 * np.pi * xi[:, 0] / Lx))
    v = np.zeros_like(xi)
    v[:, 0] = v_jet
    return dict(
        xi=xi, v=v, rho=rho, p=cole_pressure(rho, rho0, c), m=np.full(N, rho0 * dx**3),
        b=np.full(N, FREE, dtype=np.int32), bound_points=0,
        params=dict(particle_step=dx, rho_rest=rho0, speed_sound=c, mu=8.94e-4, sig=0.0708, visc_alpha=0.05,
                    dsph_delta=0.1, grav=(0.0, 0.0, -9.81), acase=1, v_inf=tuple(v_inf), p_ref=100000.0, rho_g=1.1025,
                    mu_g=1.716e-05, temp_g=298.0),
    )


def droplet(dx=0.0015, radius=0.05, seed=1234, dim=3, jitter="eps"):
    """Config C2, Examples/Droplet/{para3D,fluid_3D.bmap}: sphere R=0.05 at the origin, rho 810,
    Gissler aero with v_inf = (0, 21.55, 0).  Lattice start = centre - radius, end = centre + radius
    (circle.cpp:131-189).  dx=0.0015 -> ~155 k particles, dx=0.0008 -> ~1.02 M."""
    n1 = int(np.ceil(2 * radius / dx))
    n1 = max(n1, 1)
    pts = lattice((n1,) * dim, dx, start=(-radius,) * dim, jitter=None)
    rng = np.random.default_rng(seed)
    if jitter == "eps":  # the reference's own perturbation (circle.cpp:136-137): a tie-stress input
        pts = pts + rng.uniform(0.0, EPS * dx, size=pts.shape)
    else:  # generic positions: no neighbour sits on the support edge to the last bit
        pts = pts + rng.uniform(-float(jitter), float(jitter), size=pts.shape) * dx
    keep = (pts**2).sum(axis=1) <= radius * radius
    xi = np.ascontiguousarray(pts[keep])
    N = xi.shape[0]
    rho0 = 810.0
    v_inf = (0.0, 21.55, 0.0) if dim == 3 else (21.55, 0.0, 0.0)
    return dict(
        xi=xi, v=np.zeros_like(xi), rho=np.full(N, rho0), p=np.zeros(N), m=np.full(N, rho0 * dx**dim),
        b=np.full(N, FREE, dtype=np.int32), bound_points=0,
        params=dict(particle_step=dx, rho_rest=rho0, mu=0.000142, sig=0.0256, speed_sound=100.0, visc_alpha=0.05,
                    cfl=0.9, subits_factor=0.76, delta_t_max=1.0, delta_t_min=1e-9, frame_time_interval=1e-3,
                    acase=1, v_inf=v_inf, p_ref=100000.0, rho_g=1.1025, mu_g=1.716e-05, temp_g=298.0),
    )


def box_with_walls(n=(20, 12, 16), dx=0.01, layers=4, rho0=1000.0, c=None, hydro=True, jitter="eps", seed=7,
                   g=9.81, dim=3):
    """Open-topped tank: fluid lattice n resting on a floor and inside side walls `layers` particles thick
    (the 3D extrusion of Examples/Standing_Column, config C3).  Gravity acts along the last axis.  Walls
    come first in the particle order (Init.cpp:298-352), b = BOUND; hydrostatic initialisation follows
    Init.cpp:480-493 with the height taken along the gravity axis."""
    n = tuple(int(k) for k in n)[:dim]
    up = dim - 1
    height = n[up] * dx
    if c is None:
        c = 10.0 * np.sqrt(g * height)
    fluid = lattice(n, dx, start=(0.0,) * dim, jitter=jitter, seed=seed)
    # wall lattice: a shell around the fluid, open at the top
    nw = [k + 2 * layers for k in n]
    nw[up] = n[up] + layers + 2  # floor + a little freeboard
    shell = lattice(nw, dx, start=tuple(-layers * dx for _ in n), jitter=jitter, seed=seed + 1)
    ijk = np.rint((shell - (-layers * dx)) / dx).astype(np.int64)
    inside = np.ones(shell.shape[0], dtype=bool)
    for d in range(dim):
        if d == up:
            inside &= ijk[:, d] >= layers
        else:
            inside &= (ijk[:, d] >= layers) & (ijk[:, d] < layers + n[d])
    walls = np.ascontiguousarray(shell[~inside])
    xi = np.concatenate([walls, fluid], axis=0)
    nb, nf = walls.shape[0], fluid.shape[0]
    N = nb + nf
    rho = np.full(N, rho0)
    p = np.zeros(N)
    if hydro:
        p = np.maximum(0.0, rho0 * g * (height - xi[:, up]))
        rho = cole_density(p, rho0, c)
    b = np.concatenate([np.full(nb, BOUND, dtype=np.int32), np.full(nf, FREE, dtype=np.int32)])
    grav = [0.0, 0.0, 0.0]
    grav[up] = -g
    return dict(
        xi=xi, v=np.zeros_like(xi), rho=rho, p=p, m=np.full(N, rho0 * dx**dim), b=b, bound_points=nb,
        params=dict(particle_step=dx, rho_rest=rho0, speed_sound=float(c), mu=8.94e-4, sig=0.0, visc_alpha=0.1,
                    grav=tuple(grav), delta_t_min=1e-9, frame_time_interval=10.0),
        height=height,
    )


def dam_2d(dx=0.02):
    """Config C1, Examples/Dam_2D: 2D, fluid Square (0,0)-(2,1) -> 100x50, c=125, alpha 8.94e-4? no --
    the deck's values are kept in params below; three Pressure-Gradient walls 4 layers deep."""
    case = box_with_walls(n=(int(round(2.0 / dx)), int(round(1.0 / dx))), dx=dx, layers=4, c=125.0, hydro=True, dim=2)
    case["params"].update(dict(visc_alpha=0.1, dsph_delta=0.1, sig=0.0))
    return case


def inlet_jet(n=(5, 5, 4), dx=1e-3, v_jet=10.0, fixed=0, n_buf=4, aero_x=1.5, delete_x=None, rho0=1000.0, c=100.0,
              jitter="eps", seed=21, dim=3):
    """A square inlet block along +x (the squareCube branch of InletShape::generate_points, shapes/inlet.cpp:455-520,
    without rotation): nk PIPE layers at x = -k dx, one BACK layer at x = -nk dx and n_buf BUFFER layers behind
    it; insertion plane normal (1,0,0) with insconst = -(nk - 0.01) dx (inlet.cpp:207-210), aero entry plane at
    aero_x*dx (PIPE -> FREE, Containment.cpp:822-847), optional delete plane at delete_x*dx (Integration.cpp:127-205).
    Returns the case plus the bound_block fields of its single fluid block."""
    if dim == 2:  # n = (ni, nk): a line inlet, the 2D build's square inlet
        ni, nk = n
        nj = 1
    else:
        ni, nj, nk = n
    rng = np.random.default_rng(seed)
    pts, b = [], []
    layers = list(range(nk)) + [nk] + list(range(nk + 1, nk + 1 + n_buf))
    kinds = [PIPE] * nk + [BACK] + [BUFFER] * n_buf
    for kk, kind in zip(layers, kinds):
        for jj in range(nj):
            for ii in range(ni):
                pts.append((-kk * dx, ii * dx, jj * dx))
                b.append(kind)
    xi = np.asarray(pts, dtype=np.float64)[:, :dim]
    if jitter == "eps":
        xi = xi + rng.uniform(0.0, EPS * dx, size=xi.shape)
    elif jitter is not None:
        xi = xi + rng.uniform(-float(jitter), float(jitter), size=xi.shape) * dx
    ncol = ni * nj
    back = np.arange(nk * ncol, (nk + 1) * ncol, dtype=np.int64)
    buffer = np.stack([np.arange((nk + 1 + r) * ncol, (nk + 2 + r) * ncol, dtype=np.int64) for r in range(n_buf)], axis=1)
    N = xi.shape[0]
    v = np.zeros_like(xi)
    v[:, 0] = v_jet
    block = dict(first=0, second=N, is_fluid=1, block_type=6, fixed_vel_or_dynamic=int(fixed), insert_norm=(1.0, 0.0, 0.0),
                 insconst=-(nk - 0.01) * dx, aero_norm=(1.0, 0.0, 0.0), aeroconst=aero_x * dx, back=back, buffer=buffer)
    if delete_x is not None:
        block.update(delete_norm=(1.0, 0.0, 0.0), delconst=delete_x * dx)
    return dict(
        xi=xi, v=v, rho=np.full(N, rho0), p=np.zeros(N), m=np.full(N, rho0 * dx**dim), b=np.asarray(b, dtype=np.int32),
        bound_points=0, block=block,
        params=dict(particle_step=dx, rho_rest=rho0, speed_sound=c, mu=8.94e-4, sig=0.0708, visc_alpha=0.05,
                    delta_t_min=1e-9, frame_time_interval=1e9),
    )


def hex_mesh(lo, hi, n, vel=(0.0, 0.0, 0.0), p=100000.0, rho=1.2, outer_marker=-2, triangulate=True):
    """Uniform hexahedral mesh of the box [lo, hi] with n = (nx, ny, nz) cells, every quad face split into two
    triangles (what FOAM::Read_FOAM produces by tri-fanning, FOAMIO.cpp:604-624), in the layout of the reference's
    MESH (Var.h:396-451): verts, faces as vertex lists (CSR), leftright = (owner cell, neighbour cell or the
    boundary marker: -2 outer, -1 inner wall), cell -> faces (CSR), cell centres and the cell solution.
    vel / p / rho may be constants or callables of the cell centres [nc,3].  triangulate=False keeps the quadrilaterals
    four-cornered (what TAU::Read_tau_mesh_FACE leaves, CDFIO.cpp:1103-1227)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    nx, ny, nz = (int(k) for k in n)
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate((nx, ny, nz))]
    vid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    cid = lambda i, j, k: (k * ny + j) * nx + i
    gi, gj, gk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    verts = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    verts[vid(gi, gj, gk).ravel()] = np.stack([xs[0][gi.ravel()], xs[1][gj.ravel()], xs[2][gk.ravel()]], axis=1)
    faces, leftright, cfaces = [], [], [[] for _ in range(nx * ny * nz)]

    def add_quad(q, left, right):
        for tri in (((q[0], q[1], q[2]), (q[0], q[2], q[3])) if triangulate else (tuple(q),)):
            f = len(faces)
            faces.append(tri)
            leftright.append((left, right))
            cfaces[left].append(f)
            if right >= 0:
                cfaces[right].append(f)

    for k in range(nz):
        for j in range(ny):
            for i in range(nx + 1):   # x-normal faces
                q = (vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1))
                if i == 0:
                    add_quad(q, cid(0, j, k), outer_marker)
                elif i == nx:
                    add_quad(q, cid(nx - 1, j, k), outer_marker)
                else:
                    add_quad(q, cid(i - 1, j, k), cid(i, j, k))
    for k in range(nz):
        for j in range(ny + 1):
            for i in range(nx):       # y-normal faces
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1))
                if j == 0:
                    add_quad(q, cid(i, 0, k), outer_marker)
                elif j == ny:
                    add_quad(q, cid(i, ny - 1, k), outer_marker)
                else:
                    add_quad(q, cid(i, j - 1, k), cid(i, j, k))
    for k in range(nz + 1):
        for j in range(ny):
            for i in range(nx):       # z-normal faces
                q = (vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k))
                if k == 0:
                    add_quad(q, cid(i, j, 0), outer_marker)
                elif k == nz:
                    add_quad(q, cid(i, j, nz - 1), outer_marker)
                else:
                    add_quad(q, cid(i, j, k - 1), cid(i, j, k))
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    centre = np.zeros((nx * ny * nz, 3))
    mid = [0.5 * (x[1:] + x[:-1]) for x in xs]
    centre[cid(ci, cj, ck).ravel()] = np.stack([mid[0][ci.ravel()], mid[1][cj.ravel()], mid[2][ck.ravel()]], axis=1)
    nc = centre.shape[0]
    ev = lambda f, shape: (np.asarray(f(centre), float) if callable(f) else np.broadcast_to(np.asarray(f, float), shape)).copy()
    face_vtx = np.asarray(faces, dtype=np.int64).ravel()
    per = 3 if triangulate else 4
    return dict(
        verts=verts, face_ptr=np.arange(0, per * len(faces) + 1, per, dtype=np.int64), face_vtx=face_vtx,
        leftright=np.asarray(leftright, dtype=np.int32),
        cell_ptr=np.concatenate([[0], np.cumsum([len(c) for c in cfaces])]).astype(np.int64),
        cell_faces=np.concatenate([np.asarray(c, dtype=np.int64) for c in cfaces]),
        cCentre=centre, cVel=ev(vel, (nc, 3)), cP=ev(p, (nc,)), cRho=ev(rho, (nc,)),
    )


def tet_mesh(lo, hi, n, vel=(0.0, 0.0, 0.0), p=100000.0, rho=1.2, outer_marker=-2):
    """Tetrahedral mesh of the box [lo, hi]: every one of the n = (nx, ny, nz) bricks cut into the six tetrahedra of the Kuhn
    triangulation (the paths from its corner (0,0,0) to (1,1,1) along the axes in each of the 6 orders), which fits together
    across bricks.  Same layout as hex_mesh; every face a triangle in general position.  Also returns `locate(x)`: the
    cell holding each point (brick by floor, tetrahedron by the order of the fractional coordinates)."""
    import itertools

    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    nx, ny, nz = (int(k) for k in n)
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate((nx, ny, nz))]
    vid = lambda i, j, k: (k * (ny + 1) + j) * (nx + 1) + i
    gi, gj, gk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    verts = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    verts[vid(gi, gj, gk).ravel()] = np.stack([xs[0][gi.ravel()], xs[1][gj.ravel()], xs[2][gk.ravel()]], axis=1)
    perms = list(itertools.permutations(range(3)))
    faces, leftright, face_of, cells = [], [], {}, []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                for perm in perms:
                    at = [i, j, k]
                    tet = [vid(*at)]
                    for axis in perm:
                        at[axis] += 1
                        tet.append(vid(*at))
                    c = len(cells)
                    mine = []
                    for skip in range(4):
                        tri = tuple(v for q, v in enumerate(tet) if q != skip)
                        key = tuple(sorted(tri))
                        f = face_of.get(key)
                        if f is None:
                            f = face_of[key] = len(faces)
                            faces.append(tri)
                            leftright.append([c, outer_marker])
                        else:
                            leftright[f][1] = c
                        mine.append(f)
                    cells.append((tet, mine))
    centre = np.array([verts[list(tet)].mean(axis=0) for tet, _ in cells])
    nc = centre.shape[0]
    ev = lambda f, shape: (np.asarray(f(centre), float) if callable(f) else np.broadcast_to(np.asarray(f, float), shape)).copy()
    width = (hi - lo) / np.array([nx, ny, nz])
    rank = {perm: q for q, perm in enumerate(perms)}

    def locate(x):
        rel = (np.asarray(x, float) - lo) / width
        ijk = np.minimum(np.floor(rel).astype(int), np.array([nx, ny, nz]) - 1)
        frac = rel - ijk
        order = np.argsort(-frac, axis=1, kind="stable")      # the axis with the largest fraction is stepped along first
        which = np.array([rank[tuple(o)] for o in order])
        return ((ijk[:, 2] * ny + ijk[:, 1]) * nx + ijk[:, 0]) * 6 + which

    mesh = dict(
        verts=verts, face_ptr=np.arange(0, 3 * len(faces) + 1, 3, dtype=np.int64),
        face_vtx=np.asarray(faces, dtype=np.int64).ravel(), leftright=np.asarray(leftright, dtype=np.int32),
        cell_ptr=np.arange(0, 4 * nc + 1, 4, dtype=np.int64),
        cell_faces=np.asarray([m for _, m in cells], dtype=np.int64).ravel(),
        cCentre=centre, cVel=ev(vel, (nc, 3)), cP=ev(p, (nc,)), cRho=ev(rho, (nc,)),
    )
    return mesh, locate


def quad_mesh(lo, hi, n, vel=(0.0, 0.0), p=100000.0, rho=1.2, outer_marker=-2):
    """2D counterpart of hex_mesh: uniform mesh of the rectangle [lo, hi] with n = (nx, ny) quadrilateral cells in the
    layout of the reference's 2D MESH (what TAU::Read_tau_mesh_EDGE fills, CDFIO.cpp:1103-1227): verts [nv,2], faces =
    EDGES (two vertices each), leftright = (left cell, right cell or the boundary marker: -2 outer, -1 inner wall),
    cell -> edges, cell centres and the cell solution.  vel / p / rho may be constants or callables of the centres [nc,2]."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    nx, ny = (int(k) for k in n)
    xs = [np.linspace(lo[d], hi[d], k + 1) for d, k in enumerate((nx, ny))]
    vid = lambda i, j: j * (nx + 1) + i
    cid = lambda i, j: j * nx + i
    gi, gj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    verts = np.zeros(((nx + 1) * (ny + 1), 2))
    verts[vid(gi, gj).ravel()] = np.stack([xs[0][gi.ravel()], xs[1][gj.ravel()]], axis=1)
    faces, leftright, cfaces = [], [], [[] for _ in range(nx * ny)]

    def add_edge(e, left, right):
        f = len(faces)
        faces.append(e)
        leftright.append((left, right))
        cfaces[left].append(f)
        if right >= 0:
            cfaces[right].append(f)

    for j in range(ny):
        for i in range(nx + 1):   # x-normal edges
            e = (vid(i, j), vid(i, j + 1))
            if i == 0:
                add_edge(e, cid(0, j), outer_marker)
            elif i == nx:
                add_edge(e, cid(nx - 1, j), outer_marker)
            else:
                add_edge(e, cid(i - 1, j), cid(i, j))
    for j in range(ny + 1):
        for i in range(nx):       # y-normal edges
            e = (vid(i + 1, j), vid(i, j))
            if j == 0:
                add_edge(e, cid(i, 0), outer_marker)
            elif j == ny:
                add_edge(e, cid(i, ny - 1), outer_marker)
            else:
                add_edge(e, cid(i, j - 1), cid(i, j))
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    centre = np.zeros((nx * ny, 2))
    mid = [0.5 * (x[1:] + x[:-1]) for x in xs]
    centre[cid(ci, cj).ravel()] = np.stack([mid[0][ci.ravel()], mid[1][cj.ravel()]], axis=1)
    nc = centre.shape[0]
    ev = lambda f, shape: (np.asarray(f(centre), float) if callable(f) else np.broadcast_to(np.asarray(f, float), shape)).copy()
    return dict(
        verts=verts, face_ptr=np.arange(0, 2 * len(faces) + 1, 2, dtype=np.int64),
        face_vtx=np.asarray(faces, dtype=np.int64).ravel(), leftright=np.asarray(leftright, dtype=np.int32),
        cell_ptr=np.concatenate([[0], np.cumsum([len(c) for c in cfaces])]).astype(np.int64),
        cell_faces=np.concatenate([np.asarray(c, dtype=np.int64) for c in cfaces]),
        cCentre=centre, cVel=ev(vel, (nc, 2)), cP=ev(p, (nc,)), cRho=ev(rho, (nc,)),
    )
