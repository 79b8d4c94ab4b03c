"""Case front end: FJSPH para + bmap decks -> particle arrays and LIMITS blocks (GetInput + Init_Particles,
reference src/IO.cpp:305-723, src/Init.cpp:270-496, src/shapes/*.cpp).  The work is done by the C++ host library
(fjsph_b200/csrc/host_case.cpp) behind the C ABI; this file only copies the result into numpy arrays laid out like
fjsph_b200.cases builds them, so a deck can be handed to Engine / Oracle exactly like a synthetic case."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FjsphBlock, FjsphMesh, FjsphParams, FjsphStateView, check


def read_case(para_path: str, dim: int = 3) -> dict:
    """Returns dict(xi [n,dim], v [n,dim], rho, p, m, b, part_id, bound_points, params (FjsphParams after Set_Values),
    blocks (list of dicts with the bound_block fields, ready for Engine.set_blocks), dim, foam / tau (the aero mesh the deck
    names), ipt = (FjsphIptSettings, using_ipt): the tracker's settings as GetInput + Set_Values leave them, IO.cpp:29,126-127,
    447-453,666-680 -- using_ipt also needs an aero mesh, Integration.cpp:151)."""
    L = _lib.lib()
    h = C.c_void_p()
    check(L.fjsph_case_read(str(para_path).encode(), int(dim), C.byref(h)))
    try:
        n = int(L.fjsph_case_count(h))
        out = dict(xi=np.zeros((n, dim)), v=np.zeros((n, dim)), rho=np.zeros(n), p=np.zeros(n), m=np.zeros(n),
                   b=np.zeros(n, dtype=np.int32), part_id=np.zeros(n, dtype=np.int64))
        view = FjsphStateView()
        view.n = n
        for k, a in out.items():
            setattr(view, k, a.ctypes.data)
        check(L.fjsph_case_state(h, C.byref(view)))
        params = FjsphParams()
        check(L.fjsph_case_params(h, C.byref(params)))
        blocks = []
        for i in range(int(L.fjsph_case_num_blocks(h))):
            B = FjsphBlock()
            name = C.create_string_buffer(256)
            check(L.fjsph_case_block(h, i, C.byref(B), name, 256))
            nt, nb, nf = int(B.n_times), int(B.n_back), int(B.n_buf)
            nv = max(1, nt)
            blk = dict(
                name=name.value.decode(), first=int(B.first), second=int(B.second), is_fluid=int(B.is_fluid),
                bound_solver=int(B.bound_solver), no_slip=int(B.no_slip), block_type=int(B.block_type),
                fixed_vel_or_dynamic=int(B.fixed_vel_or_dynamic),
                times=None if nt == 0 else np.ctypeslib.as_array(C.cast(B.times, C.POINTER(C.c_double)), shape=(nt,)).copy(),
                vels=None if not B.vels else np.ctypeslib.as_array(C.cast(B.vels, C.POINTER(C.c_double)),
                                                                     shape=(nv, 3)).copy(),
                insert_norm=tuple(B.insert_norm), insconst=float(B.insconst), delete_norm=tuple(B.delete_norm),
                delconst=float(B.delconst), aero_norm=tuple(B.aero_norm), aeroconst=float(B.aeroconst))
            if nb:
                blk["back"] = np.ctypeslib.as_array(C.cast(B.back, C.POINTER(C.c_int64)), shape=(nb,)).copy()
                blk["buffer"] = np.ctypeslib.as_array(C.cast(B.buffer, C.POINTER(C.c_int64)), shape=(nb, nf)).copy()
            blocks.append(blk)
        fdir, fsol, buoy = C.create_string_buffer(1024), C.create_string_buffer(1024), C.c_int32()
        check(L.fjsph_case_foam(h, fdir, fsol, C.byref(buoy), 1024))
        tmesh, tsol, tscale = C.create_string_buffer(1024), C.create_string_buffer(1024), C.c_double(1.0)
        L.fjsph_case_tau.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_double), C.c_int32]
        check(L.fjsph_case_tau(h, tmesh, tsol, C.byref(tscale), 1024))
        ipt, using_ipt = _lib.FjsphIptSettings(), C.c_int32(0)
        check(L.fjsph_ipt_default_settings(C.byref(params), C.byref(ipt)))
        check(L.fjsph_read_para_ipt(str(para_path).encode(), float(tscale.value), C.byref(using_ipt), C.byref(ipt)))
        out.update(bound_points=int(L.fjsph_case_bound_points(h)), params=params, blocks=blocks, dim=dim,
                   ipt=(ipt, int(using_ipt.value) if params.asource != 0 else 0),
                   foam=(fdir.value.decode(), fsol.value.decode(), int(buoy.value)),
                   tau=(tmesh.value.decode(), tsol.value.decode(), float(tscale.value)))
        return out
    finally:
        L.fjsph_case_free(h)


def read_tau(mesh_file: str, solution_file: str | None = None, scale: float = 1.0) -> dict:
    """TAU::Read_tau_mesh_FACE + TAU::Read_SOLUTION (reference src/CDFIO.cpp:1228-1356,655-822) on NetCDF-3 classic files, read
    without the NetCDF library: the mesh dict Engine.upload_mesh and Oracle.set_mesh take."""
    L = _lib.lib()
    h = C.c_void_p()
    L.fjsph_tau_read.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.POINTER(C.c_void_p)]
    check(L.fjsph_tau_read(str(mesh_file).encode(), None if not solution_file else str(solution_file).encode(), float(scale),
                           C.byref(h)))
    return _mesh_dict(L, h)


def read_tau_edge(mesh_file: str, solution_file: str | None = None, scale: float = 1.0, offset_axis: int = 2) -> dict:
    """TAU::Read_tau_mesh_EDGE + TAU::Read_SOLUTION of the reference's 2D build (reference src/CDFIO.cpp:992-1097,655-822): an
    edge-based 2D mesh as the mesh dict a 2D Engine.upload_mesh / Oracle.set_mesh take (vector arrays [n,2])."""
    L = _lib.lib()
    h = C.c_void_p()
    L.fjsph_tau_read_edge.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_int32, C.POINTER(C.c_void_p)]
    check(L.fjsph_tau_read_edge(str(mesh_file).encode(), None if not solution_file else str(solution_file).encode(),
                                float(scale), int(offset_axis), C.byref(h)))
    m = _mesh_dict(L, h)
    for k in ("verts", "cCentre", "cVel"):
        m[k] = np.ascontiguousarray(m[k][:, :2])
    return m


def read_foam(foam_dir: str, solution_dir: str | None = None, buoyant: bool = False, rho_fill: float = 1.29251) -> dict:
    """FOAM::Read_FOAM (reference src/FOAMIO.cpp:943-953) for an OpenFOAM case (ASCII or binary): the mesh dict Engine.upload_mesh and
    Oracle.set_mesh take (verts, face_ptr/face_vtx, leftright, cell_ptr/cell_faces, cCentre, cVel, cP, cRho)."""
    L = _lib.lib()
    h = C.c_void_p()
    check(L.fjsph_foam_read(str(foam_dir).encode(), None if not solution_dir else str(solution_dir).encode(),
                            int(bool(buoyant)), float(rho_fill), C.byref(h)))
    return _mesh_dict(L, h)


def _mesh_dict(L, h) -> dict:
    """The arrays of a mesh handle (fjsph_foam_view), copied out; frees the handle."""
    try:
        m = FjsphMesh()
        check(L.fjsph_foam_view(h, C.byref(m)))
        nv, nf, nc = int(m.n_verts), int(m.n_faces), int(m.n_cells)

        def arr(ptr, ctype, shape):
            n = int(np.prod(shape))
            if n == 0:
                return np.zeros(shape, dtype=np.dtype(ctype))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape).copy()

        face_ptr = arr(m.face_ptr, C.c_int64, (nf + 1,))
        cell_ptr = arr(m.cell_ptr, C.c_int64, (nc + 1,))
        return dict(verts=arr(m.verts, C.c_double, (nv, 3)), face_ptr=face_ptr,
                    face_vtx=arr(m.face_vtx, C.c_int64, (int(face_ptr[-1]),)),
                    leftright=arr(m.leftright, C.c_int32, (nf, 2)), cell_ptr=cell_ptr,
                    cell_faces=arr(m.cell_faces, C.c_int64, (int(cell_ptr[-1]),)),
                    cCentre=arr(m.cCentre, C.c_double, (nc, 3)), cVel=arr(m.cVel, C.c_double, (nc, 3)),
                    cP=arr(m.cP, C.c_double, (nc,)), cRho=arr(m.cRho, C.c_double, (nc,)))
    finally:
        L.fjsph_foam_free(h)
