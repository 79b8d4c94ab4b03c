"""Slab decomposition host layer: one engine per GPU on x-slabs, NCCL send/recv over NVLink for the transport.

The reference is a single shared-memory process (OpenMP only, reference src/FJSPH.cpp:62); SURVEY.md 8e maps its
step onto one process per GPU.  The engine (fjsph_b200/csrc/halo.cu) selects, packs and unpacks migrating and ghost
particles on the device and asks the host for exactly four things through one callback (include/fjsph_b200.h,
FjsphCommFn): all-reduce SUM / MAX of a few host doubles, and a send/recv pair with the two x-neighbours on device
or host buffers.  This file implements that callback on torch.distributed (backend "nccl" on GPUs; "gloo" works
for the host-buffer operations, which is what the CPU tests exercise).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, engine as eng
from ._lib import check

COMM_SUM, COMM_MAX, COMM_SENDRECV_DEV, COMM_SENDRECV_HOST, COMM_SENDRECV_DEV_ASYNC = 0, 1, 2, 3, 4
COMM_SUM_DEV, COMM_MAX_DEV = 5, 6
COMM_FN = _lib.COMM_FN


def slab_bounds(x_min: float, x_max: float, world: int):
    """Equal-width x-slabs [x_lo, x_hi) of [x_min, x_max]; the end slabs are open (-/+1e300)."""
    edges = np.linspace(x_min, x_max, world + 1)
    lo = [(-1e300 if r == 0 else float(edges[r])) for r in range(world)]
    hi = [(1e300 if r == world - 1 else float(edges[r + 1])) for r in range(world)]
    return lo, hi


def partition(xi: np.ndarray, x_lo: float, x_hi: float) -> np.ndarray:
    """Indices of the particles rank [x_lo, x_hi) owns (same predicate as k_classify in halo.cu)."""
    x = np.asarray(xi)[:, 0]
    return np.nonzero((x >= x_lo) & (x < x_hi))[0]


class _DevBuf:
    """Zero-copy view of engine-owned device memory for torch (CUDA array interface v3)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class Transport:
    """The FjsphCommFn callback on a torch.distributed process group."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu"))
        self.calls = {COMM_SUM: 0, COMM_MAX: 0, COMM_SENDRECV_DEV: 0, COMM_SENDRECV_HOST: 0, COMM_SENDRECV_DEV_ASYNC: 0}
        self.error = None
        self.comm_stream = None  # torch view of the engine's comm stream (set_comm_stream), for the ASYNC exchanges
        self.main_stream = None  # torch view of the engine's main stream (set_main_stream), for the device reductions
        self.fn = COMM_FN(self._callback)  # keep alive as long as the engine uses it

    def set_comm_stream(self, cuda_stream_ptr: int):
        """The engine's comm stream (fjsph_slab_comm_stream): FJSPH_COMM_SENDRECV_DEV_ASYNC exchanges are ordered on it,
        so that they run beside the interior sweeps the engine queues on its main stream."""
        self.comm_stream = self.torch.cuda.ExternalStream(int(cuda_stream_ptr), device=self.device)

    def set_main_stream(self, cuda_stream_ptr: int):
        """The engine's main stream (fjsph_get_stream): FJSPH_COMM_SUM_DEV / MAX_DEV all-reduce device arrays in place on it."""
        self.main_stream = self.torch.cuda.ExternalStream(int(cuda_stream_ptr), device=self.device)

    # -- pieces
    def _host_array(self, ptr, nbytes, dtype):
        n = nbytes // np.dtype(dtype).itemsize
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,))

    def allreduce(self, ptr, nbytes, op):
        a = self._host_array(ptr, nbytes, np.float64)
        t = self.torch.from_numpy(a.copy()).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == COMM_SUM else self.dist.ReduceOp.MAX, group=self.group)
        a[...] = t.cpu().numpy()

    def sendrecv(self, tensors):
        """tensors = (send_lo, send_hi, recv_lo, recv_hi) torch tensors or None."""
        dist = self.dist
        ops = []
        s_lo, s_hi, r_lo, r_hi = tensors
        if r_lo is not None:
            ops.append(dist.P2POp(dist.irecv, r_lo, self.rank - 1, self.group))
        if r_hi is not None:
            ops.append(dist.P2POp(dist.irecv, r_hi, self.rank + 1, self.group))
        if s_lo is not None:
            ops.append(dist.P2POp(dist.isend, s_lo, self.rank - 1, self.group))
        if s_hi is not None:
            ops.append(dist.P2POp(dist.isend, s_hi, self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _callback(self, user, op, a, na, b, nb, c, nc, d, nd):
        try:
            self.calls[op] = self.calls.get(op, 0) + 1
            torch = self.torch
            if op in (COMM_SUM, COMM_MAX):
                self.allreduce(a, na, op)
            elif op in (COMM_SUM_DEV, COMM_MAX_DEV):
                # in place on the engine's device array, ordered on its main stream: no host round trip
                t = torch.as_tensor(_DevBuf(a, na), device=self.device).view(torch.float64)
                with torch.cuda.stream(self.main_stream):
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == COMM_SUM_DEV else self.dist.ReduceOp.MAX,
                                         group=self.group)
            elif op in (COMM_SENDRECV_DEV, COMM_SENDRECV_DEV_ASYNC):
                t = [torch.as_tensor(_DevBuf(p, n), device=self.device) if (p and n > 0) else None
                     for p, n in ((a, na), (b, nb), (c, nc), (d, nd))]
                if op == COMM_SENDRECV_DEV_ASYNC:
                    if self.comm_stream is None:
                        raise RuntimeError("ASYNC exchange requested but set_comm_stream was never called")
                    # NCCL orders the send/recv behind the CURRENT stream and req.wait() makes that stream wait for
                    # them: neither blocks the host, so the engine goes on queueing interior work on its main stream
                    with torch.cuda.stream(self.comm_stream):
                        self.sendrecv(t)
                else:
                    self.sendrecv(t)
            elif op == COMM_SENDRECV_HOST:
                hosts = [self._host_array(p, n, np.uint8) if (p and n > 0) else None
                         for p, n in ((a, na), (b, nb), (c, nc), (d, nd))]
                t = [None if h is None else torch.from_numpy(h.copy()).to(self.device) for h in hosts]
                self.sendrecv(t)
                for h, tt in zip(hosts[2:], t[2:]):
                    if h is not None:
                        h[...] = tt.cpu().numpy()
            else:
                raise ValueError("unknown comm op %d" % op)
            return 0
        except Exception as ex:  # never let an exception cross the C boundary
            self.error = ex
            return 1


class SlabEngine(eng.Engine):
    """One rank of a slab-decomposed simulation.  `case` holds THIS rank's particles (see partition())."""

    def __init__(self, params, case, rank, world, x_lo, x_hi, device=0, stream=None, capacity=None, group=None,
                 part_id=None, transport="auto"):
        """transport: "native" = libfjsph_b200_nccl.so (ncclSend / ncclRecv / ncclAllReduce straight from C++, the NCCL id
        handed round through torch.distributed once), "torch" = the callback below on torch.distributed, "auto" = native on
        the nccl backend when the library is built."""
        n = case["xi"].shape[0]
        # room for ghosts on both faces and for migration imbalance
        super().__init__(params, int(capacity or (n * 1.25 + 400_000)), device=device)
        if stream is not None:
            self.set_stream(stream.cuda_stream)
        self.rank, self.world = rank, world
        extra = {} if part_id is None else {"part_id": np.ascontiguousarray(part_id, dtype=np.int64)}
        self.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"], **extra)
        self.transport = Transport(group=group)
        self.native = None
        import os

        transport = os.environ.get("FJSPH_B200_TRANSPORT", transport)  # "native" | "torch" | "auto"
        N = _lib.nccl_lib() if transport in ("auto", "native") and self.transport.backend == "nccl" and world > 1 else None
        if transport == "native" and N is None:
            raise RuntimeError("native NCCL transport requested but libfjsph_b200_nccl.so is not built / backend is not nccl")
        if N is not None:
            import torch
            import torch.distributed as dist

            ident = torch.zeros(256, dtype=torch.uint8, device=self.transport.device)
            if rank == 0:
                buf = C.create_string_buffer(256)
                if N.fjsph_nccl_unique_id(buf):
                    raise RuntimeError(N.fjsph_nccl_last_error().decode())
                ident.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
            dist.broadcast(ident, src=0, group=group)
            comm = C.c_void_p()
            if N.fjsph_nccl_create(bytes(ident.cpu().numpy().tobytes()), rank, world, int(device), C.byref(comm)):
                raise RuntimeError(N.fjsph_nccl_last_error().decode())
            if N.fjsph_nccl_attach(comm, self._h, float(x_lo), float(x_hi)):
                raise RuntimeError(N.fjsph_nccl_last_error().decode())
            self.native, self._N = comm, N
            return
        check(self._L.fjsph_set_slab(self._h, rank, world, float(x_lo), float(x_hi), self.transport.fn, None))
        if self.transport.backend == "nccl":
            cs = C.c_void_p()
            check(self._L.fjsph_slab_comm_stream(self._h, C.byref(cs)))
            self.transport.set_comm_stream(cs.value)
            ms = C.c_void_p()
            check(self._L.fjsph_get_stream(self._h, C.byref(ms)))
            self.transport.set_main_stream(ms.value)
            check(self._L.fjsph_slab_device_reductions(self._h, 1))

    def close(self):
        super().close()
        if getattr(self, "native", None):
            self._N.fjsph_nccl_destroy(self.native)
            self.native = None

    def _raise_transport_error(self):
        if self.transport.error is not None:
            err, self.transport.error = self.transport.error, None
            raise err

    def integrate(self):
        try:
            return super().integrate()
        except _lib.FjsphError:
            self._raise_transport_error()
            raise

    def reupload_owned(self, fields: dict):
        """Overwrite both time levels of the OWNED particles from host arrays given in the current download order
        (hosts that keep the particles on the CPU between steps); ghosts are refreshed by the next exchanges."""
        n = self.n
        view, keep = eng.make_view(dict(fields), n)
        check(self._L.fjsph_upload_owned(self._h, C.byref(view)))

    def slab_stats(self) -> dict:
        v = [C.c_int64() for _ in range(5)]
        check(self._L.fjsph_slab_stats(self._h, *[C.byref(x) for x in v]))
        out = dict(zip(("n_owned", "n_ghost", "exchanges", "redecomps", "bytes_sent"), (int(x.value) for x in v)))
        ov = C.c_int64()
        check(self._L.fjsph_slab_overlapped(self._h, C.byref(ov)))
        out["overlapped"] = int(ov.value)
        out["transport"] = "native NCCL (libfjsph_b200_nccl.so)" if self.native else "torch.distributed " + self.transport.backend
        if self.native:
            out["device_allreduces"] = int(self._N.fjsph_nccl_calls(self.native, COMM_SUM_DEV)
                                           + self._N.fjsph_nccl_calls(self.native, COMM_MAX_DEV))
            out["host_allreduces"] = int(self._N.fjsph_nccl_calls(self.native, COMM_SUM)
                                         + self._N.fjsph_nccl_calls(self.native, COMM_MAX))
        return out

    def pair_count(self) -> float:
        return float(np.sum(self.neighbour_counts() - 1))
