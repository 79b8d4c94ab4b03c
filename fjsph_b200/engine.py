"""Host-side mirror of FJSPH's time-step interface on top of the C ABI (include/fjsph_b200.h).

Method names follow the reference functions they replace:
    update_neighbours  Neighbours.h:9        dSPH_PreStep      Shifting.h:10
    get_aero_velocity  Resid.h:43-46         Detect_Surface    Geometry.h:98-101
    dissipation_terms  Shifting.h:13-15      particle_shift    Shifting.h:18-20
    get_acc_and_Rrho   Resid.h:39-41         Do_NB_Iter        Newmark_Beta.h:11-26
    integrate / integrate_no_update          Integration.h:20-28
The arithmetic runs in hand-written sm_100a kernels; this file only moves numpy arrays across the ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FjsphBlock, FjsphMesh, FjsphParams, FjsphStateView, FjsphStepStats, check

INT64_FIELDS = ("part_id", "cellID")
INT32_FIELDS = ("b", "surf", "surfzone", "internal")
VEC_FIELDS = ("xi", "v", "acc", "Af", "aVisc", "cellV", "gradRho", "norm", "bNorm", "vPert")
SCALAR_FIELDS = tuple(
    "Rrho rho p m curve norm_curve woccl pDist deltaD cellP cellRho colourG colour lam lam_nb kernsum y".split()
)
ALL_FIELDS = INT64_FIELDS + INT32_FIELDS + VEC_FIELDS + ("L",) + SCALAR_FIELDS
# FjsphDeleted / FjsphIptPoint as numpy record types (every member 8 bytes wide or a pair of int32: no padding)
IPT_START = np.dtype([("part_id", "i8"), ("cellID", "i8"), ("t", "f8"), ("xi", "f8", 3), ("v", "f8", 3), ("mass", "f8"),
                      ("cellV", "f8", 3), ("cellRho", "f8")])
IPT_POINT = np.dtype([("part_id", "i8"), ("cellID", "i8"), ("faceID", "i8"), ("going", "i4"), ("failed", "i4"), ("t", "f8"),
                      ("dt", "f8"), ("acc", "f8"), ("xi", "f8", 3), ("v", "f8", 3), ("cellV", "f8", 3), ("cellRho", "f8")])
assert IPT_START.itemsize == C.sizeof(_lib.FjsphDeleted) and IPT_POINT.itemsize == C.sizeof(_lib.FjsphIptPoint)


def default_params(dim: int = 3, **kw) -> FjsphParams:
    """Var.h defaults, overridden by kw, then Set_Values (IO.cpp:26-128) — all in the C++ host library."""
    p = FjsphParams()
    check(_lib.lib().fjsph_default_params(C.byref(p), dim))
    set_fields(p, **kw)
    if p.frame_time_interval < 0:
        p.frame_time_interval = 1.0
    check(_lib.lib().fjsph_set_values(C.byref(p)))
    return p


def read_para(path: str, dim: int = 3, **kw):
    """GetInput + Set_Values for a FJSPH para file; returns (params, fluid_file, boundary_file)."""
    p = FjsphParams()
    L = _lib.lib()
    check(L.fjsph_default_params(C.byref(p), dim))
    fl, bd = C.create_string_buffer(512), C.create_string_buffer(512)
    check(L.fjsph_read_para(path.encode(), C.byref(p), fl, bd, 512))
    set_fields(p, **kw)
    check(L.fjsph_set_values(C.byref(p)))
    return p, fl.value.decode(), bd.value.decode()


def ipt_settings(p: FjsphParams, para: str | None = None, scale: float = 1.0, **kw):
    """IPT_SETT (Var.h:313-337) for the tracker: defaults + ipt_diam / ipt_area (IO.cpp:126-127) from `p`, then the IPT keys
    of a para file (IO.cpp:447-453) when one is given, then kw.  Returns (settings, using_ipt)."""
    s = _lib.FjsphIptSettings()
    L = _lib.lib()
    check(L.fjsph_ipt_default_settings(C.byref(p), C.byref(s)))
    use = C.c_int32(1)
    if para is not None:
        check(L.fjsph_read_para_ipt(str(para).encode(), float(scale), C.byref(use), C.byref(s)))
    set_fields(s, **kw)
    return s, int(use.value)


def mesh_max_length(mesh: dict, dim: int = 3) -> float:
    """cells.maxlength as TAU::Read_tau_mesh_FACE / _EDGE leave it (CDFIO.cpp:867-898, 1117-1183)."""
    m = FjsphMesh()
    verts = np.asarray(mesh["verts"], dtype=np.float64)
    if verts.shape[1] == 2:
        verts = np.concatenate([verts, np.zeros((verts.shape[0], 1))], axis=1)
    keep = dict(verts=np.ascontiguousarray(verts), face_ptr=np.ascontiguousarray(mesh["face_ptr"], dtype=np.int64),
                face_vtx=np.ascontiguousarray(mesh["face_vtx"], dtype=np.int64))
    for k, a in keep.items():
        setattr(m, k, a.ctypes.data)
    m.n_verts, m.n_faces = keep["verts"].shape[0], keep["face_ptr"].shape[0] - 1
    out = C.c_double(0.0)
    check(_lib.lib().fjsph_mesh_max_length(C.byref(m), int(dim), C.byref(out)))
    return float(out.value)


def set_fields(p, **kw):
    for k, v in kw.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(p, k, v)


def params_to_dict(p) -> dict:
    return {n: (list(getattr(p, n)) if hasattr(getattr(p, n), "__len__") else getattr(p, n)) for n, _ in p._fields_}


def _dtype_of(name):
    if name in INT64_FIELDS:
        return np.int64
    if name in INT32_FIELDS:
        return np.int32
    return np.float64


def _shape_of(name, n):
    if name in VEC_FIELDS:
        return (n, 3)
    if name == "L":
        return (n, 3, 3)
    return (n,)


def _pad3(name, a, n):
    """SIMDIM = 2 hosts hold [n][2] vectors and [n][2][2] matrices; the C ABI's view is [n][3] / [n][3][3] whatever the
    dimension (third components 0, L's third row and column the identity's)."""
    a = np.asarray(a, dtype=np.float64)
    if name in VEC_FIELDS and a.shape == (n, 2):
        out = np.zeros((n, 3))
        out[:, :2] = a
        return out
    if name == "L" and a.shape == (n, 2, 2):
        out = np.zeros((n, 3, 3))
        out[:, :2, :2] = a
        out[:, 2, 2] = 1.0
        return out
    return a


def make_view(arrays: dict, n: int):
    """FjsphStateView over numpy arrays (kept alive by the returned list)."""
    view = FjsphStateView()
    view.n = n
    keep = []
    for name, a in arrays.items():
        if a is None:
            continue
        if name in VEC_FIELDS or name == "L":
            a = _pad3(name, a, n)
        a = np.ascontiguousarray(a, dtype=_dtype_of(name))
        if a.shape != _shape_of(name, n):
            a = np.ascontiguousarray(np.broadcast_to(a, _shape_of(name, n)))
        keep.append(a)
        arrays[name] = a
        setattr(view, name, a.ctypes.data)
    return view, keep


class Engine:
    """One simulation resident on one B200."""

    def __init__(self, params: FjsphParams, capacity: int, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        check(self._L.fjsph_create(C.byref(params), device, int(capacity), C.byref(self._h)))
        self.capacity = int(capacity)
        self.dim = int(params.dim)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.fjsph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- settings
    @property
    def params(self) -> FjsphParams:
        p = FjsphParams()
        check(self._L.fjsph_get_params(self._h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.params
        set_fields(p, **kw)
        check(self._L.fjsph_set_params(self._h, C.byref(p)))

    def set_blocks(self, blocks):
        """blocks: list of dicts with the bound_block fields (Var.h:779-859)."""
        arr = (FjsphBlock * len(blocks))()
        keep = []
        for k, b in enumerate(blocks):
            B = arr[k]
            B.first, B.second = int(b["first"]), int(b["second"])
            B.is_fluid = int(b.get("is_fluid", 0))
            B.bound_solver = int(b.get("bound_solver", 1))
            B.no_slip = int(b.get("no_slip", 0))
            B.block_type = int(b.get("block_type", 0))
            B.fixed_vel_or_dynamic = int(b.get("fixed_vel_or_dynamic", 0))
            times = b.get("times")
            nt = 0 if times is None else len(times)
            B.n_times = nt
            if nt:
                t = np.ascontiguousarray(times, dtype=np.float64)
                keep.append(t)
                B.times = t.ctypes.data
            vels = np.zeros((max(1, nt), 3))
            if b.get("vels") is not None:
                bv = np.asarray(b["vels"], dtype=np.float64)
                bv = bv.reshape(max(1, nt), -1)
                vels[:, :bv.shape[1]] = bv
            keep.append(vels)
            B.vels = vels.ctypes.data
            for key, const in (("insert_norm", "insconst"), ("delete_norm", "delconst"), ("aero_norm", "aeroconst")):
                vec = b.get(key)
                for d in range(3):
                    getattr(B, key)[d] = 9999999.0 if vec is None else (float(vec[d]) if d < len(vec) else 0.0)
                setattr(B, const, float(b.get(const, 9999999.0)))
            back = b.get("back")
            if back is not None:
                ba = np.ascontiguousarray(back, dtype=np.int64)
                bu = np.ascontiguousarray(b["buffer"], dtype=np.int64)
                keep += [ba, bu]
                B.n_back, B.n_buf = len(ba), bu.shape[1]
                B.back, B.buffer = ba.ctypes.data, bu.ctypes.data
        check(self._L.fjsph_set_blocks(self._h, len(blocks), arr))

    def upload_mesh(self, mesh: dict):
        """The reference's MESH (Var.h:396-451) for aero source meshInfl: verts [nv,3], face_ptr/face_vtx (CSR),
        leftright [nf,2], cell_ptr/cell_faces (CSR), cCentre [nc,3], cVel [nc,3], cP [nc], cRho [nc]; in a 2D engine the
        faces are edges of two vertices and the vector arrays may be [n,2]."""
        m = FjsphMesh()
        keep = {}
        for k in ("verts", "cCentre", "cVel", "cP", "cRho"):
            a = np.asarray(mesh[k], dtype=np.float64)
            if a.ndim == 2 and a.shape[1] == 2:  # a 2D mesh (faces are edges): the ABI's [n][3] with z = 0
                a = np.concatenate([a, np.zeros((a.shape[0], 1))], axis=1)
            keep[k] = np.ascontiguousarray(a)
        for k in ("face_ptr", "face_vtx", "cell_ptr", "cell_faces"):
            keep[k] = np.ascontiguousarray(mesh[k], dtype=np.int64)
        keep["leftright"] = np.ascontiguousarray(mesh["leftright"], dtype=np.int32)
        for k, a in keep.items():
            setattr(m, k, a.ctypes.data)
        m.n_verts, m.n_faces, m.n_cells = keep["verts"].shape[0], keep["leftright"].shape[0], keep["cCentre"].shape[0]
        check(self._L.fjsph_upload_mesh(self._h, C.byref(m)))

    # -- state
    def upload_state(self, xi, v, rho, p, m, b, bound_points=0, **extra):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        n = xi.shape[0]
        arrays = dict(xi=xi, v=np.zeros_like(xi) if v is None else v, rho=rho, p=p, m=m, b=b, **extra)
        view, keep = make_view(arrays, n)
        check(self._L.fjsph_upload_state(self._h, C.byref(view), int(bound_points)))

    def upload_level(self, level: int, **fields):
        n = self.n
        view, keep = make_view(dict(fields), n)
        check(self._L.fjsph_upload_level(self._h, level, C.byref(view)))

    def download(self, fields=ALL_FIELDS, level: int = 1, out: dict | None = None) -> dict:
        n = self.n
        arrays = {}
        for f in fields:
            if out is not None and f in out and out[f].shape == _shape_of(f, n):
                arrays[f] = out[f]
            else:
                arrays[f] = np.empty(_shape_of(f, n), dtype=_dtype_of(f))
        view, keep = make_view(arrays, n)
        check(self._L.fjsph_download_state(self._h, level, C.byref(view)))
        if self.dim == 2:  # the host's shapes: [n][2] vectors, [n][2][2] matrices
            for f in fields:
                if f in VEC_FIELDS:
                    arrays[f] = np.ascontiguousarray(arrays[f][:, :2])
                elif f == "L":
                    arrays[f] = np.ascontiguousarray(arrays[f][:, :2, :2])
                if out is not None and f in out and out[f] is not arrays[f]:
                    out[f][...] = arrays[f]
                    arrays[f] = out[f]
        return arrays

    def get(self, name: str, level: int = 1) -> np.ndarray:
        return self.download((name,), level)[name]

    @property
    def n(self) -> int:
        return int(self._L.fjsph_count(self._h))

    # -- stages (reference function names)
    def update_neighbours(self):
        check(self._L.fjsph_build_neighbours(self._h))

    def neighbour_counts(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int64)
        check(self._L.fjsph_neighbour_counts(self._h, out.ctypes.data))
        return out

    def neighbours(self):
        """CSR (offsets, idx): ascending j per particle, self included — the shape of the reference's OUTL."""
        cnt = self.neighbour_counts()
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(cnt, out=off[1:])
        idx = np.empty(int(off[-1]), dtype=np.int64)
        check(self._L.fjsph_get_neighbours(self._h, off.ctypes.data, idx.ctypes.data))
        return off, idx

    def dSPH_PreStep(self) -> float:
        npd = C.c_double()
        check(self._L.fjsph_prestep(self._h, C.byref(npd)))
        return npd.value

    def get_aero_velocity(self):
        check(self._L.fjsph_aero_velocity(self._h))

    def Detect_Surface(self):
        check(self._L.fjsph_detect_surface(self._h))

    def dissipation_terms(self):
        check(self._L.fjsph_dissipation(self._h))

    def particle_shift(self):
        check(self._L.fjsph_shift(self._h))

    def get_acc_and_Rrho(self, npd: float):
        check(self._L.fjsph_forces(self._h, float(npd)))

    def Do_NB_Iter(self, npd: float) -> float:
        err = C.c_double()
        check(self._L.fjsph_nb_iter(self._h, float(npd), C.byref(err)))
        return err.value

    def find_timestep(self) -> float:
        dt = C.c_double()
        check(self._L.fjsph_find_timestep(self._h, C.byref(dt)))
        return dt.value

    def integrate_no_update(self) -> FjsphStepStats:
        s = FjsphStepStats()
        check(self._L.fjsph_integrate_no_update(self._h, C.byref(s)))
        return s

    def integrate(self) -> FjsphStepStats:
        """One Integrator::integrate (Integration.cpp:233-303) on the device-resident state."""
        s = FjsphStepStats()
        check(self._L.fjsph_step(self._h, C.byref(s)))
        return s

    def step_host(self, inputs: dict, bound_points: int, n_steps: int, out_fields=("xi", "v", "rho", "p"),
                  out: dict | None = None):
        """End-to-end call with HOST buffers: upload -> n_steps x integrate -> download.  `out` may supply
        preallocated (e.g. pinned) result arrays; they may alias the inputs."""
        n = np.asarray(inputs["xi"]).shape[0]
        vin, k1 = make_view(dict(inputs), n)
        outs = {f: (out[f] if out is not None and f in out else np.empty(_shape_of(f, n), dtype=_dtype_of(f)))
                for f in out_fields}
        vout, k2 = make_view(outs, n)
        s = FjsphStepStats()
        check(self._L.fjsph_step_host(self._h, C.byref(vin), int(bound_points), int(n_steps), C.byref(vout), C.byref(s)))
        return outs, s

    # -- instrumentation
    def timers_enable(self, on=True):
        check(self._L.fjsph_timers_enable(self._h, int(on)))

    def timers_reset(self):
        check(self._L.fjsph_timers_reset(self._h))

    def timers(self) -> dict:
        cap = 64
        names = C.create_string_buffer(cap * 32)
        ms = (C.c_double * cap)()
        launches = (C.c_int64 * cap)()
        calls = (C.c_int64 * cap)()
        n = C.c_int32()
        check(self._L.fjsph_timers_get(self._h, cap, names, ms, launches, calls, C.byref(n)))
        out = {}
        for k in range(n.value):
            nm = names.raw[k * 32:(k + 1) * 32].split(b"\0", 1)[0].decode()
            out[nm] = dict(ms=ms[k], launches=int(launches[k]), calls=int(calls[k]))
        return out

    def set_stream(self, cuda_stream: int | None):
        """Run on the given cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)."""
        check(self._L.fjsph_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def write_restart(self, path: str, frame: int = 0):
        """Checkpoint (the field set of the reference's _particles.h5, H5IO.cpp:395-538)."""
        check(self._L.fjsph_write_restart(self._h, str(path).encode(), int(frame)))

    def read_restart(self, path: str) -> int:
        """Resume from a checkpoint: pn = pnp1 from the file, settings, blocks and counters restored; returns the frame."""
        fr = C.c_int32()
        check(self._L.fjsph_read_restart(self._h, str(path).encode(), C.byref(fr)))
        return int(fr.value)

    def take_deleted(self):
        """IPT hand-off (Integration.cpp:151-169): the particles erased at a delete plane since the last call, as a dict of
        arrays (part_id, cellID, t, xi, v, mass, cellV, cellRho) in the reference's order; the engine's queue is emptied."""
        n = C.c_int64(0)
        check(self._L.fjsph_take_deleted(self._h, None, 0, C.byref(n)))
        buf = np.zeros(max(1, n.value), dtype=IPT_START)
        got = C.c_int64(0)
        check(self._L.fjsph_take_deleted(self._h, buf.ctypes.data, n.value, C.byref(got)))
        return {k: buf[k][:got.value].copy() for k in IPT_START.names}

    def ipt_integrate(self, settings, start, record_cap: int = 0) -> dict:
        """IPT::Integrate (IPT.cpp:871-1107) on the device for the hand-off records `start` -- the dict take_deleted returns,
        or an IPT_START record array -- on the mesh of upload_mesh.  Returns last (IPT_POINT per particle: pnp1 as Integrate
        leaves it), n_steps, n_records, records [n, record_cap] (the time_record handed to iptdata), n_success, n_failed."""
        if isinstance(start, dict):
            rec = np.zeros(len(start["part_id"]), dtype=IPT_START)
            for k in IPT_START.names:
                rec[k] = start[k]
            start = rec
        start = np.ascontiguousarray(start, dtype=IPT_START)
        n = start.shape[0]
        last = np.zeros(n, dtype=IPT_POINT)
        n_steps, n_records = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        records = np.zeros((n, max(record_cap, 1)), dtype=IPT_POINT)
        ok, bad = C.c_int64(0), C.c_int64(0)
        check(self._L.fjsph_ipt_integrate(self._h, C.byref(settings), n, start.ctypes.data, last.ctypes.data, n_steps.ctypes.data,
                                          records.ctypes.data if record_cap > 0 else None, int(record_cap),
                                          n_records.ctypes.data, C.byref(ok), C.byref(bad)))
        return dict(last=last, n_steps=n_steps, n_records=n_records, records=records[:, :record_cap], n_success=ok.value,
                    n_failed=bad.value)

    def set_skin(self, skin_over_dx: float):
        """Width of the neighbour superset list in units of dx (0 = cell-list sweep at every update_neighbours)."""
        check(self._L.fjsph_set_skin(self._h, float(skin_over_dx)))

    @property
    def launch_count(self) -> int:
        return int(self._L.fjsph_launch_count(self._h))
