"""ctypes wrapper of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (fjsph_b200/) never does.  See oracle/fjsph_oracle.h for the contract and citations.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(_HERE, "fjsph_oracle.h")


def _struct_from_header(header: str, name: str):
    """Build a ctypes.Structure from `typedef struct <name> { ... } <name>;` (int32_t / double fields)."""
    src = open(header).read()
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (name, name), src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ctype, rest = stmt.split(None, 1)
        base = {"int32_t": C.c_int32, "double": C.c_double, "int64_t": C.c_int64}[ctype]
        for item in rest.split(","):
            item = item.strip()
            m = re.match(r"(\w+)\[(\d+)\]", item)
            if m:
                fields.append((m.group(1), base * int(m.group(2))))
            else:
                fields.append((item, base))
    return type(name, (C.Structure,), {"_fields_": fields})


OrcParams = _struct_from_header(_HEADER, "OrcParams")
OrcStepStats = _struct_from_header(_HEADER, "OrcStepStats")
OrcIptSettings = _struct_from_header(_HEADER, "OrcIptSettings")
# OrcIptStart / OrcIptPoint as numpy record types (every member 8 bytes wide or a pair of int32: no padding)
IPT_START = np.dtype([("part_id", "i8"), ("cellID", "i8"), ("t", "f8"), ("xi", "f8", 3), ("v", "f8", 3), ("mass", "f8"),
                      ("cellV", "f8", 3), ("cellRho", "f8")])
IPT_POINT = np.dtype([("part_id", "i8"), ("cellID", "i8"), ("faceID", "i8"), ("going", "i4"), ("failed", "i4"), ("t", "f8"),
                      ("dt", "f8"), ("acc", "f8"), ("xi", "f8", 3), ("v", "f8", 3), ("cellV", "f8", 3), ("cellRho", "f8")])

BOUND, PISTON, BUFFER, BACK, PIPE, FREE, OUTLET, LOST = range(8)

_LIBS = {}


def build(fast: bool = False) -> None:
    """Compile the oracle with its committed Makefile (parity builds; fast=True adds the timing builds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE] + (["fast"] if fast else []))


def build_ref() -> None:
    """Compile FJSPH's own sources from /root/reference into oracle/_ref/ (no-op where the reference is absent)."""
    subprocess.check_call(["make", "-s", "-j4", "-C", _HERE, "-f", "Makefile.ref"])


def have_ref(kind: str = "ref3d") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "liborc_%s.so" % kind))


def ref_set_values(p: "OrcParams", kind: str = "ref3d") -> "OrcParams":
    """The derived constants of Set_Values through the reference's own inline functions (ref_harness.cpp)."""
    q = OrcParams.from_buffer_copy(p)
    _load(kind).orc_ref_set_values(C.byref(q))
    return q


def _load(kind: str):
    if kind in _LIBS:
        return _LIBS[kind]
    if kind.startswith("ref"):
        # FJSPH's own sources compiled against the stand-in headers (oracle/Makefile.ref, ref_harness.cpp): the same
        # orc_* ABI, minus the parameter defaults and the small-matrix unit entry points
        path = os.path.join(_HERE, "_ref", "liborc_%s.so" % kind)
        if not os.path.exists(path):
            build_ref()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
    else:
        path = os.path.join(_HERE, "lib", "liborc%s.so" % kind)
        if not os.path.exists(path):
            build(fast="fast" in kind)
    lib = C.CDLL(path)
    P = C.POINTER
    if kind.startswith("ref"):
        lib.orc_ref_set_values.argtypes = [P(OrcParams)]
        lib.orc_ref_pressure.argtypes = [P(OrcParams), C.c_double]
        lib.orc_ref_pressure.restype = C.c_double
        lib.orc_ref_density.argtypes = [P(OrcParams), C.c_double]
        lib.orc_ref_density.restype = C.c_double
    else:
        lib.orc_default_params.argtypes = [P(OrcParams), C.c_int]
        lib.orc_set_values.argtypes = [P(OrcParams)]
        lib.orc_qr_inverse.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_min_eigenvalue.argtypes = [C.c_void_p]
        lib.orc_min_eigenvalue.restype = C.c_double
    lib.orc_create.argtypes = [P(OrcParams)]
    lib.orc_create.restype = C.c_void_p
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_get_params.argtypes = [C.c_void_p, P(OrcParams)]
    lib.orc_set_params.argtypes = [C.c_void_p, P(OrcParams)]
    lib.orc_add_block.argtypes = (
        [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_double]
        + [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    )
    lib.orc_clear_blocks.argtypes = [C.c_void_p]
    lib.orc_get_block_range.argtypes = [C.c_void_p, C.c_int, P(C.c_int64), P(C.c_int64)]
    lib.orc_set_particles.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 7
    lib.orc_count.argtypes = [C.c_void_p]
    lib.orc_count.restype = C.c_int64
    lib.orc_bound_points.argtypes = [C.c_void_p]
    lib.orc_bound_points.restype = C.c_int64
    for f in ("orc_get_f64", "orc_set_f64", "orc_get_i64", "orc_set_i64"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
    lib.orc_set_mesh.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.orc_set_mesh.argtypes += [C.c_void_p] * 6
    lib.orc_first_cell_errors.argtypes = [C.c_void_p]
    lib.orc_update_neighbours.argtypes = [C.c_void_p]
    lib.orc_neighbour_total.argtypes = [C.c_void_p]
    lib.orc_neighbour_total.restype = C.c_int64
    lib.orc_get_neighbours.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_prestep.argtypes = [C.c_void_p]
    lib.orc_prestep.restype = C.c_double
    for f in ("orc_aero_velocity", "orc_detect_surface", "orc_dissipation", "orc_particle_shift"):
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.orc_forces.argtypes = [C.c_void_p, C.c_double]
    lib.orc_nb_iter.argtypes = [C.c_void_p, C.c_double]
    lib.orc_find_timestep.argtypes = [C.c_void_p]
    lib.orc_find_timestep.restype = C.c_double
    lib.orc_integrate_no_update.argtypes = [C.c_void_p, P(OrcStepStats)]
    lib.orc_integrate_no_update.restype = C.c_double
    lib.orc_integrate.argtypes = [C.c_void_p, P(OrcStepStats)]
    lib.orc_integrate.restype = C.c_double
    lib.orc_kernel.argtypes = [C.c_double] * 3
    lib.orc_kernel.restype = C.c_double
    lib.orc_get_n_full.argtypes = [C.c_double] * 2
    lib.orc_get_n_full.restype = C.c_double
    _LIBS[kind] = lib
    return lib


def default_params(dim: int = 3, kind: str | None = None, **kw) -> "OrcParams":
    """Var.h defaults, overridden by kw, then Set_Values (IO.cpp:26-128)."""
    lib = _load(kind or ("%dd" % dim))
    p = OrcParams()
    lib.orc_default_params(C.byref(p), dim)
    set_fields(p, **kw)
    lib.orc_set_values(C.byref(p))
    return p


def ipt_settings(p: "OrcParams", **kw) -> "OrcIptSettings":
    """IPT_SETT defaults (Var.h:313-337) with ipt_diam / ipt_area as Set_Values derives them (IO.cpp:126-127) and the
    values IPT::Integrate reads from the rest of SIM (gravity, gas viscosity, rest density, max_subits)."""
    s = OrcIptSettings()
    s.eq_order, s.max_subits, s.record, s.max_steps = 2, p.max_subits, 1, 100000
    s.relax, s.n_relax, s.max_x, s.max_length = 0.6, 5.0, 9999999.0, 0.0
    s.diam = ((6.0 * p.sim_mass) / (np.pi * p.rho_rest)) ** (1.0 / 3.0)
    s.area = np.pi * s.diam * s.diam / 4.0
    s.grav[:] = list(p.grav)
    s.mu_g, s.rho_rest = p.mu_g, p.rho_rest
    for k, v in kw.items():
        if k == "grav":
            s.grav[:] = list(v)
        else:
            setattr(s, k, v)
    return s


def set_fields(p, **kw):
    for k, v in kw.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)


def params_to_dict(p) -> dict:
    out = {}
    for name, _ in p._fields_:
        v = getattr(p, name)
        out[name] = list(v) if hasattr(v, "__len__") else v
    return out


_INT_FIELDS = ("part_id", "cellID", "b", "surf", "surfzone", "internal", "ipt_n_failed")
_VEC_FIELDS = ("xi", "v", "acc", "Af", "aVisc", "cellV", "gradRho", "norm", "bNorm", "vPert")
_SCALAR_FIELDS = (
    "Rrho rho p m curve norm_curve woccl pDist deltaD cellP cellRho colourG colour lam lam_nb kernsum y".split()
)
ALL_FIELDS = _INT_FIELDS + _VEC_FIELDS + ("L",) + tuple(_SCALAR_FIELDS)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One simulation on the CPU oracle.  kind: '3d', '2d', '3d_mt' (the parity build, particle loops threaded),
    '3d_fast', '3d_fast_serialdiss' (timing builds)."""

    def __init__(self, params: "OrcParams", kind: str | None = None):
        self.dim = int(params.dim)
        self.kind = kind or ("%dd" % self.dim)
        self.lib = _load(self.kind)
        self.h = self.lib.orc_create(C.byref(params))
        if not self.h:
            raise RuntimeError("orc_create failed (dim mismatch?)")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- params
    @property
    def params(self) -> "OrcParams":
        p = OrcParams()
        self.lib.orc_get_params(self.h, C.byref(p))
        return p

    def set_params(self, **kw):
        p = self.params
        set_fields(p, **kw)
        self.lib.orc_set_params(self.h, C.byref(p))

    # -- blocks / particles
    def add_block(self, is_fluid, first, second, bound_solver=1, no_slip=0, block_type=0, fixed_vel_or_dynamic=0,
                  times=None, vels=None, insert_norm=None, insconst=9999999.0, delete_norm=None,
                  delconst=9999999.0, aero_norm=None, aeroconst=9999999.0, back=None, buffer=None):
        times_a = None if times is None else np.ascontiguousarray(times, dtype=np.float64)
        nt = 0 if times_a is None else len(times_a)
        vels_a = np.zeros((max(1, nt), 3)) if vels is None else np.ascontiguousarray(vels, dtype=np.float64).reshape(-1, 3)

        def v3(x):
            if x is None:
                return None
            a = np.zeros(3)
            a[: len(x)] = x
            return a

        ins, dele, aero = v3(insert_norm), v3(delete_norm), v3(aero_norm)
        back_a = None if back is None else np.ascontiguousarray(back, dtype=np.int64)
        buf_a = None if buffer is None else np.ascontiguousarray(buffer, dtype=np.int64)
        nback = 0 if back_a is None else len(back_a)
        nbuf = 0 if buf_a is None else buf_a.shape[1]
        r = self.lib.orc_add_block(
            self.h, int(is_fluid), int(first), int(second), int(bound_solver), int(no_slip), int(block_type),
            int(fixed_vel_or_dynamic), nt, _ptr(times_a), _ptr(vels_a), _ptr(ins), float(insconst), _ptr(dele),
            float(delconst), _ptr(aero), float(aeroconst), nback, _ptr(back_a), nbuf, _ptr(buf_a),
        )
        if r < 0:
            raise RuntimeError("orc_add_block: boundary blocks must precede fluid blocks")
        return r

    def set_particles(self, xi, v, rho, p, m, b, bound_points=0, part_id=None):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        n = xi.shape[0]
        assert xi.shape[1] == self.dim
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        rho = np.ascontiguousarray(np.broadcast_to(rho, (n,)), dtype=np.float64)
        p = np.ascontiguousarray(np.broadcast_to(p, (n,)), dtype=np.float64)
        m = np.ascontiguousarray(np.broadcast_to(m, (n,)), dtype=np.float64)
        b = np.ascontiguousarray(np.broadcast_to(b, (n,)), dtype=np.int32)
        pid = None if part_id is None else np.ascontiguousarray(part_id, dtype=np.int64)
        self.lib.orc_set_particles(self.h, n, int(bound_points), _ptr(xi), _ptr(v), _ptr(rho), _ptr(p), _ptr(m),
                                   _ptr(b), _ptr(pid))

    def set_mesh(self, mesh: dict):
        """mesh: dict with verts [nv,3], face_ptr/face_vtx (CSR), leftright [nf,2] int32, cell_ptr/cell_faces (CSR),
        cCentre [nc,3], cVel [nc,3], cP [nc], cRho [nc] -- the reference's MESH (Var.h:396-451)."""
        a = {k: np.ascontiguousarray(mesh[k], dtype=(np.int32 if k == "leftright" else
                                                      np.int64 if k in ("face_ptr", "face_vtx", "cell_ptr", "cell_faces")
                                                      else np.float64)) for k in
             ("verts", "face_ptr", "face_vtx", "leftright", "cell_ptr", "cell_faces", "cCentre", "cVel", "cP", "cRho")}
        self.lib.orc_set_mesh(self.h, a["verts"].shape[0], _ptr(a["verts"]), a["leftright"].shape[0], _ptr(a["face_ptr"]),
                              _ptr(a["face_vtx"]), _ptr(a["leftright"]), a["cCentre"].shape[0], _ptr(a["cell_ptr"]),
                              _ptr(a["cell_faces"]), _ptr(a["cCentre"]), _ptr(a["cVel"]), _ptr(a["cP"]), _ptr(a["cRho"]))

    @property
    def first_cell_errors(self) -> int:
        return int(self.lib.orc_first_cell_errors(self.h))

    @property
    def n(self) -> int:
        return int(self.lib.orc_count(self.h))

    def get(self, name: str, level: int = 1) -> np.ndarray:
        n = self.n
        if name in _INT_FIELDS:
            out = np.empty(n, dtype=np.int64)
            r = self.lib.orc_get_i64(self.h, level, name.encode(), _ptr(out))
            assert r == 1, name
            return out
        d = self.dim
        width = d if name in _VEC_FIELDS else (d * d if name == "L" else 1)
        out = np.empty((n, width), dtype=np.float64)
        r = self.lib.orc_get_f64(self.h, level, name.encode(), _ptr(out))
        assert r == width, name
        if name == "L":
            return out.reshape(n, d, d)
        return out if width > 1 else out[:, 0]

    def set(self, name: str, value, level: int = 1) -> None:
        n = self.n
        if name in _INT_FIELDS:
            a = np.ascontiguousarray(value, dtype=np.int64).reshape(n)
            r = self.lib.orc_set_i64(self.h, level, name.encode(), _ptr(a))
        else:
            a = np.ascontiguousarray(value, dtype=np.float64).reshape(n, -1)
            r = self.lib.orc_set_f64(self.h, level, name.encode(), _ptr(a))
        assert r > 0, name

    def state(self, level: int = 1) -> dict:
        return {k: self.get(k, level) for k in ALL_FIELDS}

    # -- stages
    def update_neighbours(self):
        self.lib.orc_update_neighbours(self.h)

    def neighbours(self):
        """CSR (offsets, idx, d2), ascending j within each list, self included."""
        n = self.n
        tot = int(self.lib.orc_neighbour_total(self.h))
        off = np.empty(n + 1, dtype=np.int64)
        idx = np.empty(tot, dtype=np.int64)
        d2 = np.empty(tot, dtype=np.float64)
        self.lib.orc_get_neighbours(self.h, _ptr(off), _ptr(idx), _ptr(d2))
        return off, idx, d2

    def neighbour_counts(self) -> np.ndarray:
        """List lengths, self included (outlist[i].size())."""
        if hasattr(self.lib, "orc_neighbour_counts"):
            out = np.empty(self.n, dtype=np.int64)
            self.lib.orc_neighbour_counts.argtypes = [C.c_void_p, C.c_void_p]
            self.lib.orc_neighbour_counts.restype = None
            self.lib.orc_neighbour_counts(self.h, _ptr(out))
            return out
        return np.diff(self.neighbours()[0])

    def prestep(self) -> float:
        return float(self.lib.orc_prestep(self.h))

    def aero_velocity(self):
        self.lib.orc_aero_velocity(self.h)

    def detect_surface(self):
        self.lib.orc_detect_surface(self.h)

    def dissipation(self):
        self.lib.orc_dissipation(self.h)

    def particle_shift(self):
        self.lib.orc_particle_shift(self.h)

    def forces(self, npd: float):
        self.lib.orc_forces(self.h, float(npd))

    def nb_iter(self, npd: float):
        self.lib.orc_nb_iter(self.h, float(npd))

    def find_timestep(self) -> float:
        return float(self.lib.orc_find_timestep(self.h))

    def integrate_no_update(self):
        s = OrcStepStats()
        e = self.lib.orc_integrate_no_update(self.h, C.byref(s))
        return float(e), s

    def integrate(self):
        s = OrcStepStats()
        e = self.lib.orc_integrate(self.h, C.byref(s))
        return float(e), s

    def ipt_integrate(self, settings: "OrcIptSettings", start: np.ndarray, record_cap: int = 0) -> dict:
        """IPT::Integrate (IPT.cpp:871-1107) for the IPT_START records `start` on the mesh of set_mesh.  Returns last
        (IPT_POINT per particle), n_steps, n_records, records [n, record_cap], n_success, n_failed."""
        start = np.ascontiguousarray(start, dtype=IPT_START)
        n = start.shape[0]
        last = np.zeros(n, dtype=IPT_POINT)
        n_steps, n_records = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        records = np.zeros((n, max(record_cap, 1)), dtype=IPT_POINT)
        ok, bad = C.c_int64(0), C.c_int64(0)
        self.lib.orc_ipt_integrate.argtypes = [C.c_void_p, C.POINTER(OrcIptSettings), C.c_int64] + [C.c_void_p] * 4 + [
            C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        self.lib.orc_ipt_integrate.restype = C.c_int
        self.lib.orc_ipt_integrate(self.h, C.byref(settings), n, _ptr(start), _ptr(last), _ptr(n_steps),
                                   _ptr(records) if record_cap > 0 else None, int(record_cap), _ptr(n_records), C.byref(ok), C.byref(bad))
        return dict(last=last, n_steps=n_steps, n_records=n_records, records=records[:, :record_cap], n_success=ok.value,
                    n_failed=bad.value)


# ---- FJSPH's own front end (reference libraries only): GetInput + Init_Particles, the LIMITS blocks, FOAM::Read_FOAM
def ref_read_case(para_path: str, kind: str = "ref3d") -> "Oracle":
    """The deck through the reference's GetInput (IO.cpp:305-723) and Init_Particles (Init.cpp:270-496); the result is
    a simulation handle like any other (get(), params, integrate(), ...).  A bad deck ends the process: the reference
    calls exit()."""
    lib = _load(kind)
    lib.orc_ref_read_case.restype = C.c_void_p
    lib.orc_ref_read_case.argtypes = [C.c_char_p]
    o = Oracle.__new__(Oracle)
    o.dim = 2 if "2d" in kind else 3
    o.kind, o.lib = kind, lib
    o.h = lib.orc_ref_read_case(os.fsencode(para_path))
    return o


def ref_blocks(o: "Oracle") -> list:
    """The LIMITS vector (Var.h:779-859) of a reference handle as dicts with the keys Engine.set_blocks takes."""
    lib = o.lib
    lib.orc_ref_block_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    lib.orc_ref_block_arrays.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    nb = int(lib.orc_ref_num_bound_blocks(C.c_void_p(o.h)))
    out = []
    for i in range(int(lib.orc_ref_num_blocks(C.c_void_p(o.h)))):
        rng, ints = np.zeros(2, dtype=np.int64), np.zeros(8, dtype=np.int32)
        norms, consts, name = np.zeros(9), np.zeros(3), C.create_string_buffer(256)
        assert lib.orc_ref_block_info(o.h, i, _ptr(rng), _ptr(ints), _ptr(norms), _ptr(consts), name, 256) == 0
        nt, nback, nbuf = int(ints[0]), int(ints[1]), int(ints[2])
        times, vels = np.zeros(max(nt, 1)), np.zeros((max(nt, 1) + 1, 3))
        back, buf = np.zeros(max(nback, 1), dtype=np.int64), np.zeros((max(nback, 1), max(nbuf, 1)), dtype=np.int64)
        nv = lib.orc_ref_block_arrays(o.h, i, _ptr(times), _ptr(vels), _ptr(back), _ptr(buf))
        out.append(dict(name=name.value.decode(), first=int(rng[0]), second=int(rng[1]), is_fluid=int(i >= nb), n_times=nt,
                        bound_solver=int(ints[3]), no_slip=int(ints[4]), block_type=int(ints[5]),
                        fixed_vel_or_dynamic=int(ints[6]), particle_order=int(ints[7]), times=times[:nt], vels=vels[:nv],
                        insert_norm=tuple(norms[0:3]), delete_norm=tuple(norms[3:6]), aero_norm=tuple(norms[6:9]),
                        insconst=float(consts[0]), delconst=float(consts[1]), aeroconst=float(consts[2]),
                        back=back[:nback], buffer=buf[:nback, :nbuf]))
    return out


def ref_mesh(o: "Oracle") -> dict:
    """The MESH of a reference handle (Var.h:396-451) in the layout of fjsph_b200.cases.hex_mesh."""
    lib = o.lib
    lib.orc_ref_mesh_sizes.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_ref_mesh_arrays.argtypes = [C.c_void_p] * 11
    sz = np.zeros(5, dtype=np.int64)
    lib.orc_ref_mesh_sizes(o.h, _ptr(sz))
    nv, nf, nc, nfv, ncf = (int(k) for k in sz)
    d = o.dim
    m = dict(verts=np.zeros((nv, d)), face_ptr=np.zeros(nf + 1, dtype=np.int64), face_vtx=np.zeros(nfv, dtype=np.int64),
             leftright=np.zeros((nf, 2), dtype=np.int32), cell_ptr=np.zeros(nc + 1, dtype=np.int64),
             cell_faces=np.zeros(ncf, dtype=np.int64), cCentre=np.zeros((nc, d)), cVel=np.zeros((nc, d)), cP=np.zeros(nc),
             cRho=np.zeros(nc))
    lib.orc_ref_mesh_arrays(o.h, *[_ptr(m[k]) for k in ("verts", "face_ptr", "face_vtx", "leftright", "cell_ptr",
                                                        "cell_faces", "cCentre", "cVel", "cP", "cRho")])
    return m


def ref_read_foam(o: "Oracle", foam_dir: str, foam_sol: str, buoyant: bool = False) -> dict:
    """FOAM::Read_FOAM (FOAMIO.cpp:538-955) on an ASCII case; returns the mesh it built."""
    o.lib.orc_ref_read_foam.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    o.lib.orc_ref_read_foam(o.h, os.fsencode(foam_dir), os.fsencode(foam_sol), int(buoyant))
    return ref_mesh(o)


def ref_read_bmap(o: "Oracle", bmap_file: str, alpha_deg: float, grav) -> np.ndarray:
    """TAU::Read_BMAP (CDFIO.cpp:234-315): gravity after the boundary map has been read."""
    g_in = np.ascontiguousarray(grav, dtype=np.float64)
    g_out = np.zeros_like(g_in)
    o.lib.orc_ref_read_bmap.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_void_p, C.c_void_p]
    o.lib.orc_ref_read_bmap(o.h, os.fsencode(bmap_file), float(alpha_deg), _ptr(g_in), _ptr(g_out))
    return g_out


def ref_read_tau(o: "Oracle", mesh_file: str, sol_file: str, scale: float = 1.0) -> dict:
    """TAU::Read_tau_mesh_FACE + TAU::Read_SOLUTION (CDFIO.cpp:1228-1356,655-822) on NetCDF-3 classic files; returns the mesh."""
    o.lib.orc_ref_read_tau.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]
    o.lib.orc_ref_read_tau(o.h, os.fsencode(mesh_file), os.fsencode(sol_file), float(scale))
    return ref_mesh(o)


def qr_inverse(a: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.float64)
    d = a.shape[0]
    lib = _load("%dd" % d)
    inv = np.zeros_like(a)
    ok = lib.orc_qr_inverse(_ptr(a), _ptr(inv))
    return bool(ok), inv


def min_eigenvalue(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    lib = _load("%dd" % a.shape[0])
    return float(lib.orc_min_eigenvalue(_ptr(a)))


def kernel(r, H, Wc, dim=3) -> float:
    return float(_load("%dd" % dim).orc_kernel(r, H, Wc))


def get_n_full(dx, H, dim=3) -> float:
    return float(_load("%dd" % dim).orc_get_n_full(dx, H))


def ref_read_tau_edge(o: "Oracle", mesh_file: str, sol_file: str, scale: float = 1.0, offset_axis: int = 2) -> dict:
    """TAU::Read_tau_mesh_EDGE + TAU::Read_SOLUTION of the 2D build (CDFIO.cpp:992-1097,655-822); returns the mesh."""
    o.lib.orc_ref_read_tau_edge.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_int]
    o.lib.orc_ref_read_tau_edge(o.h, os.fsencode(mesh_file), os.fsencode(sol_file), float(scale), int(offset_axis))
    return ref_mesh(o)
