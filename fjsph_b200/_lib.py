"""Loader for libfjsph_b200.so (the C ABI of include/fjsph_b200.h).  There is no CPU fallback: a missing
library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(ROOT, "include", "fjsph_b200.h")
LIB_PATH = os.environ.get("FJSPH_B200_LIB") or os.path.join(_PKG, "lib", "libfjsph_b200.so")


def struct_from_header(name: str, header: str = HEADER):
    """ctypes.Structure generated from `typedef struct <name> { ... } <name>;` so the Python mirror can
    never drift from the C header."""
    src = open(header).read()
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (name, name), src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    base = {"int32_t": C.c_int32, "double": C.c_double, "int64_t": C.c_int64}
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        stmt = stmt.replace("const ", "")
        ctype, rest = stmt.split(None, 1)
        for item in rest.split(","):
            item = item.strip()
            ptr = item.startswith("*") or ctype.endswith("*")
            item = item.lstrip("*").strip()
            t = base[ctype.rstrip("*")]
            m = re.match(r"(\w+)\[(\d+)\]", item)
            if m:
                fields.append((m.group(1), t * int(m.group(2))))
            elif ptr:
                fields.append((item, C.c_void_p))
            else:
                fields.append((item, t))
    return type(name, (C.Structure,), {"_fields_": fields})


FjsphParams = struct_from_header("FjsphParams")
FjsphBlock = struct_from_header("FjsphBlock")
FjsphStateView = struct_from_header("FjsphStateView")
FjsphStepStats = struct_from_header("FjsphStepStats")
FjsphMesh = struct_from_header("FjsphMesh")
FjsphDeleted = struct_from_header("FjsphDeleted")
FjsphIptSettings = struct_from_header("FjsphIptSettings")
FjsphIptPoint = struct_from_header("FjsphIptPoint")


# FjsphCommFn, include/fjsph_b200.h
COMM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                      C.c_void_p, C.c_int64)


def declared_functions(header: str = HEADER):
    """Names of every function the header declares (used by the CPU test that checks the exports)."""
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(fjsph_\w+)\s*\(", src)))


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j8"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libfjsph_b200.so failed:\n" + out.stdout)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libfjsph_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fjsph_b200/csrc`. There is no CPU fallback." % LIB_PATH
        )
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    L.fjsph_last_error.restype = C.c_char_p
    L.fjsph_version.restype = C.c_char_p
    L.fjsph_default_params.argtypes = [P(FjsphParams), C.c_int]
    L.fjsph_read_para.argtypes = [C.c_char_p, P(FjsphParams), C.c_char_p, C.c_char_p, C.c_int]
    L.fjsph_set_values.argtypes = [P(FjsphParams)]
    L.fjsph_create.argtypes = [P(FjsphParams), C.c_int, C.c_int64, P(vp)]
    L.fjsph_destroy.argtypes = [vp]
    L.fjsph_get_params.argtypes = [vp, P(FjsphParams)]
    L.fjsph_set_params.argtypes = [vp, P(FjsphParams)]
    L.fjsph_set_blocks.argtypes = [vp, C.c_int32, P(FjsphBlock)]
    L.fjsph_upload_state.argtypes = [vp, P(FjsphStateView), C.c_int64]
    L.fjsph_upload_level.argtypes = [vp, C.c_int, P(FjsphStateView)]
    L.fjsph_download_state.argtypes = [vp, C.c_int, P(FjsphStateView)]
    L.fjsph_upload_owned.argtypes = [vp, P(FjsphStateView)]
    L.fjsph_upload_mesh.argtypes = [vp, P(FjsphMesh)]
    L.fjsph_count.argtypes = [vp]
    L.fjsph_count.restype = C.c_int64
    L.fjsph_build_neighbours.argtypes = [vp]
    L.fjsph_neighbour_counts.argtypes = [vp, vp]
    L.fjsph_get_neighbours.argtypes = [vp, vp, vp]
    L.fjsph_prestep.argtypes = [vp, P(C.c_double)]
    for f in ("fjsph_aero_velocity", "fjsph_detect_surface", "fjsph_dissipation", "fjsph_shift"):
        getattr(L, f).argtypes = [vp]
    L.fjsph_forces.argtypes = [vp, C.c_double]
    L.fjsph_nb_iter.argtypes = [vp, C.c_double, P(C.c_double)]
    L.fjsph_find_timestep.argtypes = [vp, P(C.c_double)]
    L.fjsph_integrate_no_update.argtypes = [vp, P(FjsphStepStats)]
    L.fjsph_step.argtypes = [vp, P(FjsphStepStats)]
    L.fjsph_step_host.argtypes = [vp, P(FjsphStateView), C.c_int64, C.c_int32, P(FjsphStateView), P(FjsphStepStats)]
    L.fjsph_timers_reset.argtypes = [vp]
    L.fjsph_timers_enable.argtypes = [vp, C.c_int]
    L.fjsph_timers_get.argtypes = [vp, C.c_int32, vp, vp, vp, vp, P(C.c_int32)]
    L.fjsph_set_stream.argtypes = [vp, vp]
    L.fjsph_launch_count.argtypes = [vp]
    L.fjsph_launch_count.restype = C.c_int64
    L.fjsph_set_owned.argtypes = [vp, C.c_int64]
    L.fjsph_set_skin.argtypes = [vp, C.c_double]
    L.fjsph_set_slab.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_double, COMM_FN, vp]
    L.fjsph_slab_comm_stream.argtypes = [vp, P(vp)]
    L.fjsph_slab_overlapped.argtypes = [vp, P(C.c_int64)]
    L.fjsph_get_stream.argtypes = [vp, P(vp)]
    L.fjsph_slab_device_reductions.argtypes = [vp, C.c_int32]
    L.fjsph_take_deleted.argtypes = [vp, vp, C.c_int64, P(C.c_int64)]
    L.fjsph_ipt_default_settings.argtypes = [P(FjsphParams), P(FjsphIptSettings)]
    L.fjsph_read_para_ipt.argtypes = [C.c_char_p, C.c_double, P(C.c_int32), P(FjsphIptSettings)]
    L.fjsph_mesh_max_length.argtypes = [P(FjsphMesh), C.c_int32, P(C.c_double)]
    L.fjsph_ipt_integrate.argtypes = [vp, P(FjsphIptSettings), C.c_int64, vp, vp, vp, vp, C.c_int64, vp, P(C.c_int64), P(C.c_int64)]
    L.fjsph_slab_stats.argtypes = [vp, P(C.c_int64), P(C.c_int64), P(C.c_int64), P(C.c_int64), P(C.c_int64)]
    L.fjsph_foam_read.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, P(vp)]
    L.fjsph_foam_view.argtypes = [vp, P(FjsphMesh)]
    L.fjsph_foam_free.argtypes = [vp]
    L.fjsph_foam_free.restype = None
    L.fjsph_write_restart.argtypes = [vp, C.c_char_p, C.c_int32]
    L.fjsph_read_restart.argtypes = [vp, C.c_char_p, P(C.c_int32)]
    L.fjsph_case_read.argtypes = [C.c_char_p, C.c_int, P(vp)]
    L.fjsph_case_free.argtypes = [vp]
    L.fjsph_case_free.restype = None
    for f in ("fjsph_case_count", "fjsph_case_bound_points"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_int64
    for f in ("fjsph_case_num_blocks", "fjsph_case_dim"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_int32
    L.fjsph_case_params.argtypes = [vp, P(FjsphParams)]
    L.fjsph_case_foam.argtypes = [vp, C.c_char_p, C.c_char_p, P(C.c_int32), C.c_int32]
    L.fjsph_case_io.argtypes = [vp, P(C.c_int32), P(C.c_int64), C.c_char_p, C.c_char_p, C.c_int32]
    L.fjsph_case_block.argtypes = [vp, C.c_int32, P(FjsphBlock), C.c_char_p, C.c_int32]
    L.fjsph_case_state.argtypes = [vp, P(FjsphStateView)]
    _lib = L
    return L


NCCL_LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libfjsph_b200_nccl.so")
_nccl = None


def nccl_lib():
    """libfjsph_b200_nccl.so: the slab transport natively on NCCL (include/fjsph_b200_nccl.h).  None where it is not built."""
    global _nccl
    if _nccl is not None:
        return _nccl
    if not os.path.exists(NCCL_LIB_PATH):
        return None
    lib()  # libfjsph_b200.so first: the transport library links against it
    N = C.CDLL(NCCL_LIB_PATH)
    vp, P = C.c_void_p, C.POINTER
    N.fjsph_nccl_unique_id.argtypes = [C.c_char_p]
    N.fjsph_nccl_create.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, P(vp)]
    N.fjsph_nccl_attach.argtypes = [vp, vp, C.c_double, C.c_double]
    N.fjsph_nccl_allreduce_host.argtypes = [vp, vp, C.c_int64, C.c_int32]
    N.fjsph_nccl_barrier.argtypes = [vp]
    N.fjsph_nccl_calls.argtypes = [vp, C.c_int32]
    N.fjsph_nccl_calls.restype = C.c_int64
    N.fjsph_nccl_last_error.restype = C.c_char_p
    N.fjsph_nccl_destroy.argtypes = [vp]
    _nccl = N
    return N


class FjsphError(RuntimeError):
    pass


def check(status: int):
    if status != 0:
        raise FjsphError("fjsph status %d: %s" % (status, lib().fjsph_last_error().decode(errors="replace")))
