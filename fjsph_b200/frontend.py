"""Case front end: FJSPH para + bmap decks -> particle arrays and LIMITS blocks (GetInput + Init_Particles,
reference src/IO.cpp:305-723, src/Init.cpp:270-496, src/shapes/*.cpp).  The work is done by the C++ host library
(fjsph_b200/csrc/host_case.cpp) behind the C ABI; this file only copies the result into numpy arrays laid out like
fjsph_b200.cases builds them, so a deck can be handed to Engine / Oracle exactly like a synthetic case."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FjsphBlock, FjsphParams, FjsphStateView, check


def read_case(para_path: str, dim: int = 3) -> dict:
    """Returns dict(xi [n,dim], v [n,dim], rho, p, m, b, part_id, bound_points, params (FjsphParams after Set_Values),
    blocks (list of dicts with the bound_block fields, ready for Engine.set_blocks), dim)."""
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
        out.update(bound_points=int(L.fjsph_case_bound_points(h)), params=params, blocks=blocks, dim=dim)
        return out
    finally:
        L.fjsph_case_free(h)
