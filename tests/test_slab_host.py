"""Host side of the slab decomposition on CPU: the partition helpers and the FjsphCommFn callback
(fjsph_b200/slab.py) on a world_size-2 gloo group, driven through the same C function pointer the engine calls."""
import ctypes as C
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from fjsph_b200 import cases, slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_partition_covers_every_particle_once():
    case = cases.synthetic_block((30, 6, 5), 1e-3, jitter=0.1, seed=2)
    xi = case["xi"]
    dx = 1e-3
    for world in (1, 2, 3, 4):
        lo, hi = slab.slab_bounds(xi[:, 0].min() - 0.5 * dx, xi[:, 0].max() + 0.5 * dx, world)
        assert lo[0] == -1e300 and hi[-1] == 1e300
        assert all(hi[r] == lo[r + 1] for r in range(world - 1))
        owned = [slab.partition(xi, lo[r], hi[r]) for r in range(world)]
        allidx = np.concatenate(owned)
        assert len(allidx) == len(xi) and len(np.unique(allidx)) == len(xi)
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 2 * 6 * 5  # balanced to two lattice planes


WORKER = textwrap.dedent("""
    import ctypes as C, os, sys
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from fjsph_b200 import slab
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    t = slab.Transport()
    fn = t.fn                     # the C function pointer the engine would call
    vp = lambda a: C.c_void_p(a.ctypes.data)
    # all-reduce SUM and MAX of host doubles, in place
    v = np.array([1.0 + rank, 10.0 * (rank + 1), -3.0])
    assert fn(None, slab.COMM_SUM, vp(v), v.nbytes, None, 0, None, 0, None, 0) == 0
    assert np.allclose(v, [sum(1.0 + r for r in range(world)), sum(10.0 * (r + 1) for r in range(world)), -3.0 * world])
    m = np.array([float(rank), -float(rank)])
    assert fn(None, slab.COMM_MAX, vp(m), m.nbytes, None, 0, None, 0, None, 0) == 0
    assert np.allclose(m, [world - 1.0, 0.0])
    # neighbour send/recv on host buffers (the particle counts of a re-decomposition travel this way)
    s_lo, s_hi = np.array([100 + rank], dtype=np.int64), np.array([200 + rank], dtype=np.int64)
    r_lo, r_hi = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64)
    has_lo, has_hi = rank > 0, rank < world - 1
    rc = fn(None, slab.COMM_SENDRECV_HOST, vp(s_lo), 8 if has_lo else 0, vp(s_hi), 8 if has_hi else 0,
            vp(r_lo), 8 if has_lo else 0, vp(r_hi), 8 if has_hi else 0)
    assert rc == 0, t.error
    if has_lo: assert r_lo[0] == 200 + rank - 1     # what the lower neighbour sent upwards
    if has_hi: assert r_hi[0] == 100 + rank + 1     # what the upper neighbour sent downwards
    # an unknown op must be reported as a failure, not raised through the C boundary
    assert fn(None, 99, None, 0, None, 0, None, 0, None, 0) == 1 and t.error is not None
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_comm_callback_on_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
        assert "rank %d ok" % rank in out
