"""Checkpoint / resume (csrc/restart.cu: the field set of the reference's _particles.h5, H5IO.cpp:395-538): a run that
is written out after a few steps and resumed in a NEW engine must continue like the run that never stopped.  The
resumed engine rebuilds its superset neighbour list, so only the FP64 summation order may differ."""
import numpy as np
import pytest

from fjsph_b200 import cases, engine as eng
from fjsph_b200._lib import FjsphError
from tests.util import relerr

pytestmark = pytest.mark.gpu


def run(e, steps):
    out = []
    for _ in range(steps):
        s = e.integrate()
        out.append((s.iterations, s.dt, s.n_add, s.n_del, s.total_points))
    return out


def make(case, cap=None):
    e = eng.Engine(eng.default_params(3, **dict(case["params"], delta_t_min=1e-9)), cap or case["xi"].shape[0])
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    return e


@pytest.mark.parametrize("which", ["walls", "inlet"])
def test_resume_continues_the_run(which, tmp_path):
    if which == "walls":
        case = cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4, jitter=0.05)
        a = make(case)
    else:
        case = cases.inlet_jet(n=(5, 5, 4), fixed=1, delete_x=2.5, jitter=0.03)
        a = make(case, 4 * case["xi"].shape[0])
        a.set_blocks([case["block"]])
    run(a, 5)
    path = str(tmp_path / "case_particles.fjr")
    a.write_restart(path, frame=7)
    more_a = run(a, 4)
    b = eng.Engine(eng.default_params(3, particle_step=case["params"]["particle_step"]), a.capacity)
    assert b.read_restart(path) == 7
    pa, pb = a.params, b.params
    assert pb.rho_rest == pa.rho_rest and pb.speed_sound == pa.speed_sound and pb.cfl > 0
    more_b = run(b, 4)
    for sa, sb in zip(more_a, more_b):
        assert sa[0] == sb[0] and sa[2:] == sb[2:], (sa, sb)
        assert abs(sa[1] - sb[1]) <= 1e-12 * sa[1]
    ga = a.download(("part_id", "b", "xi", "v", "rho", "p", "acc", "Rrho"))
    gb = b.download(("part_id", "b", "xi", "v", "rho", "p", "acc", "Rrho"))
    assert np.array_equal(ga["part_id"], gb["part_id"]) and np.array_equal(ga["b"], gb["b"])
    for f, tol in (("xi", 1e-11), ("rho", 1e-11), ("v", 1e-8), ("p", 1e-8), ("acc", 1e-6), ("Rrho", 1e-6)):
        assert relerr(gb[f], ga[f]) <= tol, (f, relerr(gb[f], ga[f]))
    assert abs(a.params.current_time - b.params.current_time) <= 1e-12 * a.params.current_time


def test_bad_restart_files_are_errors(tmp_path):
    case = cases.synthetic_block((6, 6, 6), 1e-3, jitter=0.1)
    e = make(case)
    with pytest.raises(FjsphError, match="cannot open"):
        e.read_restart(str(tmp_path / "missing.fjr"))
    junk = tmp_path / "junk.fjr"
    junk.write_bytes(b"not a restart file at all")
    with pytest.raises(FjsphError, match="not a version"):
        e.read_restart(str(junk))
    good = tmp_path / "good.fjr"
    e.update_neighbours()
    e.write_restart(str(good))
    data = good.read_bytes()
    (tmp_path / "short.fjr").write_bytes(data[: len(data) // 2])
    with pytest.raises(FjsphError, match="truncated"):
        e.read_restart(str(tmp_path / "short.fjr"))
    small = eng.Engine(eng.default_params(3, particle_step=1e-3), 10)
    with pytest.raises(FjsphError, match="do not fit"):
        small.read_restart(str(good))
