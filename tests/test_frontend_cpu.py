"""Case front end (GetInput + Init_Particles restated in csrc/host_case.cpp): para + bmap decks -> particles in the
reference's order.  Host-only code behind the C ABI, so these run without a GPU.  The reference ships no fixtures for
it; the checks are the analytic lattice counts and positions the generators must produce (shapes/*.cpp), the block
layout of Init.cpp:298-475 and the consistency of the inlet tables."""
import os

import numpy as np
import pytest

from fjsph_b200 import _lib, cases, frontend

DECKS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decks")
EPS = np.finfo(np.float64).eps


def deck(name):
    return os.path.join(DECKS, name)


# a case without walls names an empty boundary file, as the reference's Examples/Droplet does
NO_WALLS = " -----------------------------------------------------\n BOUNDARY MAPPING\n -----------------------------------------------------\n"


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_sphere_block_is_the_culled_lattice():
    """circle.cpp:131-189: lattice start = centre - R, ceil(2R/dx) points per axis, points with |x - c|^2 > R^2 culled,
    x fastest; perturbation U(0, eps dx) per axis."""
    c = frontend.read_case(deck("droplet3d.para"), 3)
    ref = cases.droplet(dx=0.005, jitter="eps")
    assert c["xi"].shape == ref["xi"].shape == (4166, 3)
    assert np.abs(c["xi"] - ref["xi"]).max() <= 1e-16  # same lattice, same order, different eps stream
    assert c["bound_points"] == 0 and len(c["blocks"]) == 1
    b = c["blocks"][0]
    assert (b["first"], b["second"], b["is_fluid"], b["block_type"]) == (0, 4166, 1, 2)
    assert np.all(c["b"] == cases.FREE) and np.all(c["rho"] == 810.0)
    assert np.allclose(c["m"], 810.0 * 0.005**3, rtol=1e-15) and np.all(c["p"] == 0.0)
    assert np.array_equal(c["part_id"], np.arange(4166))
    P = c["params"]
    assert P.acase == 1 and tuple(P.v_inf) == (0.0, 21.55, 0.0) and P.solver_type == 0
    assert abs(P.H - 2.0 * 0.005) < 1e-18 and P.sr == 4 * P.H * P.H


def test_dam_2d_layout_and_hydrostatic_init():
    c = frontend.read_case(deck("dam2d.para"), 2)
    dx = 0.05
    names = [b["name"] for b in c["blocks"]]
    assert names == ["Left", "Bottom", "Water"]  # boundary blocks first, then fluid (Init.cpp:298-475)
    left, bottom, water = c["blocks"]
    # line.cpp: ni = ceil(|end - start| / dx) along the line, nk layers along (dy, -dx) of the direction
    assert (left["first"], left["second"]) == (0, 41 * 4)
    assert (bottom["first"], bottom["second"]) == (164, 164 + 81 * 4)
    assert (water["first"], water["second"]) == (488, 488 + 40 * 20)
    assert c["bound_points"] == 488
    assert left["bound_solver"] == 1 and bottom["bound_solver"] == 2  # Pressure-Gradient, Ghost (the '#' line is cut)
    assert np.all(c["b"][:488] == cases.BOUND) and np.all(c["b"][488:] == cases.FREE)
    xl = c["xi"][:164]
    # the Left wall runs downwards from (-0.05, 2): direction (0,-1), thickness normal (dy, -dx) = (-1, 0)
    assert np.allclose(xl[0], (-0.05, 2.0), atol=1e-12) and np.allclose(xl[3], (-0.05 - 3 * dx, 2.0), atol=1e-12)
    assert np.allclose(xl[4], (-0.05, 2.0 - dx), atol=1e-12)
    xw = c["xi"][488:]
    assert np.allclose(xw[0], (0, 0), atol=1e-12) and np.allclose(xw[1], (dx, 0), atol=1e-12)
    assert np.allclose(xw[40], (0, dx), atol=1e-12)
    # the perturbation is the reference's stream (square.cpp:103-104,135): std::default_random_engine = minstd_rand0
    # from seed 1, uniform_real_distribution(0, eps dx) = generate_canonical over two draws, one value for all axes
    x1 = 16807
    x2 = 16807 * x1 % 2147483647
    R = 2147483646.0
    canon = ((x1 - 1) + (x2 - 1) * R) / (R * R)
    assert xw[0, 0] == xw[0, 1] == canon * (EPS * dx)
    # Init.cpp:480-493: p = max(0, -rho0 g_y (h - y)), rho = EOS^-1(p), for EVERY particle (walls too)
    P = c["params"]
    y = c["xi"][:, 1]
    p = np.maximum(0.0, 1000.0 * 9.81 * (1.0 - y))
    assert np.allclose(c["p"], p, rtol=1e-14)
    assert np.allclose(c["rho"], cases.cole_density(p, 1000.0, 125.0), rtol=1e-14)
    assert P.dim == 2 and abs(P.W_correc - 7.0 / (4.0 * np.pi * P.H**2)) < 1e-12 * P.W_correc


def test_inlet_and_pipe_blocks():
    """inlet.cpp (Circle sub-shape, lattice disks) + cylinder.cpp (Hollow), both rotated by "Rotation angles: 0,0,90";
    Init.cpp:355-426 order: PIPE layers, the BACK row, then the BUFFER rows, with the back / buffer tables."""
    c = frontend.read_case(deck("jet3d.para"), 3)
    dx = 1e-4
    pipe, fluid = c["blocks"]
    assert pipe["name"] == "Pipe" and pipe["bound_solver"] == 2 and pipe["is_fluid"] == 0
    assert fluid["name"] == "Fluid" and fluid["block_type"] == 6 and fluid["is_fluid"] == 1
    nb = c["bound_points"]
    # hollow cylinder: ni = ceil(2 pi R / dx) around, nj = ceil(L/dx) + 1 deep, nk = 2 layers
    ni, nj, nk = int(np.ceil(2 * np.pi * 5.5e-4 / dx)), int(np.ceil(0.0008 / dx)) + 1, 2
    assert nb == ni * nj * nk
    xp = c["xi"][:nb]
    r = np.hypot(xp[:, 0], xp[:, 2])  # rotated about z by 90 degrees: the axis is y, the pipe extends to -y
    assert np.allclose(np.unique(np.round(r / dx, 6)), [5.5, 6.5])
    assert xp[:, 1].max() < 1e-12 and abs(xp[:, 1].min() + (nj - 1) * dx) < 1e-12
    # inlet: disk of the lattice points within R of the axis, nk = ceil(L/dx) PIPE layers, 1 BACK, 4 BUFFER
    b = c["b"][nb:]
    ndisk = len(fluid["back"])
    assert np.array_equal(np.bincount(b, minlength=6)[[cases.BUFFER, cases.BACK, cases.PIPE]], [4 * ndisk, ndisk, 5 * ndisk])
    assert np.all(b[: 5 * ndisk] == cases.PIPE) and np.all(b[5 * ndisk: 6 * ndisk] == cases.BACK)
    assert np.array_equal(fluid["back"], nb + 5 * ndisk + np.arange(ndisk))
    assert fluid["buffer"].shape == (ndisk, 4)
    assert np.array_equal(fluid["buffer"][:, 0], nb + 6 * ndisk + np.arange(ndisk))
    xf = c["xi"]
    for k in range(ndisk):  # buffer particles sit straight behind their back particle, one spacing apart
        col = np.concatenate([[fluid["back"][k]], fluid["buffer"][k]])
        assert np.allclose(np.diff(xf[col, 1]), -dx, atol=1e-12) and np.allclose(xf[col, 0], xf[col[0], 0], atol=1e-12)
    assert np.allclose(c["v"][nb:], (0.0, 33.63, 0.0), atol=1e-12)  # vmag * insert_norm, insert_norm = R e_x
    assert np.allclose(fluid["insert_norm"], (0, 1, 0), atol=1e-15) and fluid["insconst"] == -0.0005
    assert fluid["aero_norm"] == (0.0, 1.0, 0.0) and fluid["aeroconst"] == 0.0001
    # PIPE particles carry the boundary mass (Init.cpp:364), BACK / BUFFER the fluid mass -- equal by Set_Values
    assert np.allclose(c["m"], 997.0 * dx**3, rtol=1e-14)


def test_intersecting_particles_are_removed(tmp_path):
    """Check_Intersection (Init.cpp:61-225): fluid particles within 0.9 dx of a kept boundary particle go, and a fluid
    block loses the particles a LATER fluid block overlaps."""
    write(tmp_path, "b.bmap", """
                       Name: Floor
                      Shape: Plane
           Start coordinate: 0,0,0
           Right coordinate: 0.4,0,0
             End coordinate: 0.4,0.4,0
 Wall radial particle count: 2
 block end
""")
    write(tmp_path, "f.bmap", """
                       Name: A
                      Shape: Cube
           Start coordinate: 0,0,0
             End coordinate: 0.4,0.4,0.4
 block end
                       Name: B
                      Shape: Cube
           Start coordinate: 0.2,0,0.2
             End coordinate: 0.6,0.4,0.6
 block end
""")
    para = write(tmp_path, "para", """
 Input boundary definition filename: %s
    Input fluid definition filename: %s
               SPH initial spacing: 0.1
              SPH aerodynamic case: (none)
           SPH frame time interval: 0.1
""" % (tmp_path / "b.bmap", tmp_path / "f.bmap"))
    c = frontend.read_case(para, 3)
    floor, A, B = c["blocks"]
    # plane: 4 x 4 points, 2 layers along normalized(dj x di) = -z ... the layer at z = 0 coincides with A's bottom layer
    assert floor["second"] - floor["first"] == 32
    # A: 4^3 lattice minus its z = 0 layer (on the floor) minus the 2 x 4 x 2 points B's lattice also holds
    assert A["second"] - A["first"] == 64 - 16 - 16
    assert B["second"] - B["first"] == 64
    x = c["xi"]
    from scipy.spatial import cKDTree

    d, _ = cKDTree(x).query(x, k=2)
    assert d[:, 1].min() > 0.09  # nothing closer than 0.9 dx survives


def test_hcp_cube_spacing(tmp_path):
    """square.cpp:109-131: HCP ordering puts every nearest neighbour one spacing away."""
    write(tmp_path, "f.bmap", """
                       Name: Block
                      Shape: Cube
 Particle ordering (0=grid,1=HCP): 1
           Start coordinate: 0,0,0
             End coordinate: 1,1,1
 block end
""")
    para = write(tmp_path, "para", """
 Input boundary definition filename: %s
    Input fluid definition filename: %s
               SPH initial spacing: 0.1
              SPH aerodynamic case: (none)
           SPH frame time interval: 0.1
""" % (write(tmp_path, "none.bmap", NO_WALLS), tmp_path / "f.bmap"))
    c = frontend.read_case(para, 3)
    ni, nj, nk = 10, int(np.ceil(1 / 0.1 / np.sqrt(3) * 2)), int(np.ceil(1 / 0.1 / np.sqrt(6) * 3))
    assert c["xi"].shape[0] == ni * nj * nk
    from scipy.spatial import cKDTree

    d, _ = cKDTree(c["xi"]).query(c["xi"], k=2)
    assert np.allclose(d[:, 1], 0.1, rtol=1e-12)


def test_moving_wall_schedule_and_errors(tmp_path):
    write(tmp_path, "b.bmap", """
                       Name: Piston
                      Shape: Line
           Start coordinate: 0,0
             End coordinate: 0,1
 Wall radial particle count: 1
 Time position data
 3
 0.0 0.0 0.0
 0.5 1.0 0.0
 1.5 1.0 2.0
 block end
""")
    write(tmp_path, "f.bmap", """
                       Name: W
                      Shape: Square
           Start coordinate: 0.2,0
             End coordinate: 0.6,0.4
 block end
""")
    para = write(tmp_path, "para", """
 Input boundary definition filename: %s
    Input fluid definition filename: %s
               SPH initial spacing: 0.1
              SPH aerodynamic case: (none)
           SPH frame time interval: 0.1
""" % (tmp_path / "b.bmap", tmp_path / "f.bmap"))
    c = frontend.read_case(para, 2)
    piston = c["blocks"][0]
    # get_boundary_velocity (Init.cpp:26-38): piecewise-constant velocities between the time stamps
    assert np.array_equal(piston["times"], [0.0, 0.5, 1.5])
    # (one entry per time stamp: past the last one the wall stands still, where the reference reads out of bounds)
    assert np.allclose(piston["vels"][:, :2], [[2.0, 0.0], [0.0, 2.0], [0.0, 0.0]])
    # errors come back as FjsphError with the reference's diagnosis, never exit()
    bad = write(tmp_path, "bad.bmap", "   Name: X\n  Shape: Blob\n block end\n")
    none = " Input boundary definition filename: %s\n" % write(tmp_path, "none.bmap", NO_WALLS)
    para2 = write(tmp_path, "para2", none + " Input fluid definition filename: %s\n SPH initial spacing: 0.1\n SPH aerodynamic case: (none)\n SPH frame time interval: 1\n" % bad)
    with pytest.raises(_lib.FjsphError, match="Unrecognised boundary shape"):
        frontend.read_case(para2, 2)
    with pytest.raises(_lib.FjsphError, match="file missing"):
        frontend.read_case(write(tmp_path, "para3", none + " Input fluid definition filename: nope.bmap\n SPH initial spacing: 0.1\n SPH aerodynamic case: (none)\n"), 2)
    # the reference insists on both block files (IO.cpp:555-585)
    with pytest.raises(_lib.FjsphError, match="boundary definition filename"):
        frontend.read_case(write(tmp_path, "para5", " Input fluid definition filename: %s\n SPH initial spacing: 0.1\n SPH aerodynamic case: (none)\n" % (tmp_path / "f.bmap")), 2)
    nodx = write(tmp_path, "para4", none + " Input fluid definition filename: %s\n SPH aerodynamic case: (none)\n" % (tmp_path / "f.bmap"))
    with pytest.raises(_lib.FjsphError, match="Aerodynamic coupling model"):   # IO.cpp:606-627: the case must be named
        frontend.read_case(write(tmp_path, "para6", none + " Input fluid definition filename: %s\n SPH initial spacing: 0.1\n"
                                 .replace(" SPH aerodynamic case: (none)\n", "") % (tmp_path / "f.bmap")), 2)
    with pytest.raises(_lib.FjsphError, match="initial spacing"):
        frontend.read_case(nodx, 2)


def test_arc_and_arch_blocks():
    """arc.cpp: an Arc (2D) is nk rings inwards from the radius, each a start straight of ceil(s/dx) points, the block
    file's i-direction count of arc points dtheta = dx/R apart, and an end straight; an Arch (3D) is nj slices along the
    plane normal of nk layers OUTWARDS from the radius, ceil(arclength/dtheta) arc points each.  (The decks are compared
    particle for particle with the reference's own generator in test_frontend_vs_reference.py.)"""
    dx = 0.05
    c = frontend.read_case(deck("arc2d.para"), 2)
    names = [b["name"] for b in c["blocks"]]
    assert names == ["Bowl", "Lid", "Hook", "Sweep", "Water"]
    # the last wall block is clear of everything else, so no particle of it is culled: 2 rings of 21 points
    sweep = c["blocks"][3]
    x = c["xi"][sweep["first"]:sweep["second"]]
    assert x.shape[0] == 2 * 21
    rad = np.hypot(x[:, 0] - 7.0, x[:, 1] - 1.0)
    assert np.allclose(rad[:21], 0.8, atol=1e-12) and np.allclose(rad[21:], 0.8 - dx, atol=1e-12)
    ang = np.arctan2(x[:21, 1] - 1.0, x[:21, 0] - 7.0)
    assert np.allclose(ang, np.deg2rad(10.0) + np.arange(21) * dx / 0.8, atol=1e-12)
    # three points on an arc give the centre (5, 2) and the radius 1; the points start at the start point
    hook = c["blocks"][2]
    x = c["xi"][hook["first"]:hook["second"]]
    assert np.allclose(np.hypot(x[:20, 0] - 5.0, x[:20, 1] - 2.0), 1.0, atol=1e-12)
    assert np.allclose(x[0], (5.0, 1.0), atol=1e-12)

    c = frontend.read_case(deck("arch3d.para"), 3)
    assert [b["name"] for b in c["blocks"]] == ["Trough", "Vault", "Stub", "Water"]
    tr = c["blocks"][0]
    x = c["xi"][tr["first"]:tr["second"]]
    nrad, smax, emax, nk, nj = int(np.ceil(np.pi / (dx / 0.8))), 4, 3, 3, 12
    assert x.shape[0] == (nrad + smax + emax) * nk * nj
    ring = x.reshape(nj, nk, smax + nrad + emax, 3)
    for kk in range(nj):
        assert np.allclose(ring[kk, :, :, 1], kk * dx, atol=1e-12)            # slices along the normal (0, 1, 0)
        for jj in range(nk):
            arc = ring[kk, jj, smax:smax + nrad]
            assert np.allclose(np.hypot(arc[:, 0], arc[:, 2] - 1.0), 0.8 + jj * dx, atol=1e-12)
    # the start straight leaves the start point (-0.8, 0, 1) against the direction of travel, u x w = (0, 0, -1)
    assert np.allclose(ring[0, 0, :smax, 0], -0.8, atol=1e-12)
    assert np.allclose(ring[0, 0, :smax, 2], 1.0 + dx * np.arange(smax, 0, -1), atol=1e-12)
    # centre + start + end in 3D: the reference states a -90 degree arc whatever the end point, so only straights exist
    stub = c["blocks"][2]
    assert stub["second"] - stub["first"] == (4 + 4) * 2 * 3


def test_arch_errors_are_reported(tmp_path):
    """Where arc.cpp calls exit() (a start point off the plane of the stated normal, arc.cpp:162-166; end and start radii
    that differ, :36-42) or flags a fault (nothing that defines an arc, :455-459) the product returns the diagnosis."""
    fluid = write(tmp_path, "f.bmap", "   Name: W\n  Shape: Cube\n Start coordinate: 0,0,0\n End coordinate: 0.2,0.2,0.2\n block end\n")

    def para_for(i, block):
        wall = write(tmp_path, "b%d.bmap" % i, block)
        return write(tmp_path, "para%d" % i, " Input boundary definition filename: %s\n Input fluid definition filename: %s\n"
                     " SPH initial spacing: 0.1\n SPH aerodynamic case: (none)\n SPH frame time interval: 1\n" % (wall, fluid))

    head = "   Name: A\n  Shape: Arch\n Particle spacing: 0.1\n Wall radial particle count: 1\n j-direction count: 1\n"
    with pytest.raises(_lib.FjsphError, match="plane defined by the provided normal"):
        frontend.read_case(para_for(0, head + " Centre coordinate: 0,0,2\n Start coordinate: 1,0,2.5\n Arch normal: 0,0,1\n"
                                    " Arc length (degree): 90\n block end\n"), 3)
    with pytest.raises(_lib.FjsphError, match="ending radius differs"):
        frontend.read_case(para_for(1, head + " Centre coordinate: 0,0,2\n Start coordinate: 1,0,2\n End coordinate: 0,2,2\n block end\n"), 3)
    with pytest.raises(_lib.FjsphError, match="not been sufficiently defined"):
        frontend.read_case(para_for(2, head + " Centre coordinate: 0,0,2\n Radius: 1\n block end\n"), 3)


def test_json_block_files(tmp_path):
    """read_shapes_JSON (shapes.cpp:229-395): block files with a .json extension (any case).  The same blocks give the
    same particles as their bmap form; blocks come in the order of their names (the reference's JSON object is a
    std::map), a repeated key keeps its last value, a vector is taken only from an array of exactly `dim` numbers, and
    typed reads are as strict as nlohmann/json's -- where the reference exits, the product returns the diagnosis."""
    a = frontend.read_case(deck("jet3d.para"), 3)
    b = frontend.read_case(deck("jet3d_json.para"), 3)
    for f in ("xi", "v", "rho", "p", "m", "b", "part_id"):
        assert np.array_equal(a[f], b[f]), f
    for A, B in zip(a["blocks"], b["blocks"]):
        for k in ("name", "first", "second", "is_fluid", "bound_solver", "block_type", "fixed_vel_or_dynamic", "insconst", "aeroconst"):
            assert A[k] == B[k], k
    c = frontend.read_case(deck("tank2d_json.para"), 2)
    assert [B["name"] for B in c["blocks"]] == ["Bottom", "Bowl", "Left", "Pegs", "Right", "Drop", "Water"]   # sorted, not file order
    bottom, left, pegs = c["blocks"][0], c["blocks"][2], c["blocks"][3]
    assert np.array_equal(c["xi"][pegs["first"]:pegs["second"]], [[3.5, 1.8], [3.55, 1.8], [3.6, 1.8], [3.5, 1.85]])   # "Coordinate data"
    # the repeated "Start coordinates" of Bottom: the last one counts (and a 3-entry array in a 2D deck is ignored)
    x = c["xi"][bottom["first"]:bottom["second"]]
    assert abs(x[:, 0].min() + 0.05) < 1e-12 and x.shape[0] == 4 * 81
    assert left["no_slip"] == 1 and left["bound_solver"] == 0   # "Wall is no-slip": true, "Boundary solver": "DBC"
    water = c["blocks"][6]
    assert water["second"] - water["first"] == 40 * 20

    def para_for(i, fluid_json):
        f = write(tmp_path, "f%d.json" % i, fluid_json)
        none = write(tmp_path, "none%d.bmap" % i, NO_WALLS)
        return write(tmp_path, "para%d" % i, " Input boundary definition filename: %s\n Input fluid definition filename: %s\n"
                     " SPH initial spacing: 0.1\n SPH aerodynamic case: (none)\n SPH frame time interval: 1\n" % (none, f))

    ok = '{"W": {"Shape": "Square", "Start coordinates": [0, 0], "End coordinates": [0.4, 0.4]%s}}'
    assert frontend.read_case(para_for(0, ok % ""), 2)["xi"].shape[0] == 16
    # a dynamic inlet is named, not numbered, in JSON (shapes.cpp:104-121)
    inlet = ('{"In": {"Shape": "Inlet", "Sub-shape": "Square", "Fixed velocity or dynamic inlet BC": "Dynamic", "Start jet velocity": 1.0,'
             ' "Start coordinates": [0, 0], "End coordinates": [0, 0.4], "Length": 0.3, "Insertion plane constant": -0.3}}')
    blk = frontend.read_case(para_for(1, inlet), 2)["blocks"][0]
    assert blk["fixed_vel_or_dynamic"] == 1
    with pytest.raises(_lib.FjsphError, match="Parameter: Radius Error: type must be number"):
        frontend.read_case(para_for(2, ok % ', "Radius": "wide"'), 2)
    with pytest.raises(_lib.FjsphError, match="Parameter: Wall is no-slip Error: type must be boolean"):
        frontend.read_case(para_for(3, ok % ', "Wall is no-slip": 1'), 2)
    with pytest.raises(_lib.FjsphError, match="JSON parse error"):
        frontend.read_case(para_for(4, '{"W": {"Shape": "Square", }}'), 2)
    with pytest.raises(_lib.FjsphError, match="Unrecognised boundary shape"):
        frontend.read_case(para_for(5, '{"W": {"Shape": "Blob"}}'), 2)


def test_dam2d_example_deck_is_the_references():
    """tests/decks/dam2d_example.para states the numbers of the reference's Examples/Dam_2D (BASELINE.json configs[0]); where
    the reference tree is present the two decks give the same particles, walls and settings, bit for bit."""
    mine = frontend.read_case(os.path.join(DECKS, "dam2d_example.para"), 2)
    assert mine["xi"].shape == (8371, 2) and mine["bound_points"] == 3520
    assert [(b["name"], b["first"], b["second"]) for b in mine["blocks"]] == [
        ("Left", 0, 800), ("Right", 800, 2396), ("Bottom", 2396, 3520), ("Fluid", 3520, 8371)]
    ref_para = "/root/reference/Examples/Dam_2D/para"
    if not os.path.exists(ref_para):
        pytest.skip("reference tree absent")
    ref = frontend.read_case(ref_para, 2)
    for k in ("xi", "v", "rho", "p", "m", "b"):
        assert np.array_equal(mine[k], ref[k]), k
    from fjsph_b200.engine import params_to_dict

    pm, pr = params_to_dict(mine["params"]), params_to_dict(ref["params"])
    differing = [k for k in pm if pm[k] != pr[k] and not (isinstance(pm[k], float) and np.isnan(pm[k]) and np.isnan(pr[k]))]
    assert differing == [], differing


def test_a_vlm_deck_is_read_and_then_refused(tmp_path):
    """A 3D deck that names a VLM definition (and no mesh) takes the vortex-lattice aero source (IO.cpp:465-477): the front end
    reads it -- asource 2, the particles as for any deck -- and fjsph_create refuses it by name; it never runs on a constant
    free stream unnoticed.  A 2D build ignores the key, as the reference does."""
    import shutil

    from fjsph_b200 import _lib, engine

    for f in ("droplet3d.para", "droplet3d_fluid.bmap", "droplet3d_boundary.bmap", "dam2d.para", "dam2d_fluid.bmap", "dam2d_boundary.bmap"):
        shutil.copy(os.path.join(DECKS, f), tmp_path / f)
    for name in ("droplet3d.para", "dam2d.para"):
        with open(tmp_path / name, "a") as f:
            f.write("\n VLM definition filename: (thisfile)\n")
    c3 = frontend.read_case(str(tmp_path / "droplet3d.para"), 3)
    assert c3["params"].asource == 2 and c3["xi"].shape[0] == frontend.read_case(os.path.join(DECKS, "droplet3d.para"), 3)["xi"].shape[0]
    with pytest.raises(_lib.FjsphError, match="aero source 2 is not supported.*VLM is out of scope"):
        engine.Engine(c3["params"], 16)                # refused before a device is asked for
    assert frontend.read_case(str(tmp_path / "dam2d.para"), 2)["params"].asource == 0
