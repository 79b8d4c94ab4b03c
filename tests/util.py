"""Shared helpers for the parity tests: run the same case through the CPU oracle and the CUDA engine."""
import numpy as np

from fjsph_b200 import engine as eng
from oracle import oracle as orc

TOL = 1e-10  # north_star: per-step density, acceleration and shifting in FP64 within 1e-10 relative


def make_pair(case, dim=3, capacity=None, **kw):
    """(oracle, engine, params) initialised with the same inputs and the same derived constants."""
    params = dict(case["params"])
    params.update(kw)
    po = orc.default_params(dim, **params)
    o = orc.Oracle(po)
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    pe = eng.default_params(dim, **params)
    n = case["xi"].shape[0]
    e = eng.Engine(pe, capacity or n)
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    return o, e, po


def relerr(a, b):
    """Normwise relative error: max|a-b| / max|b| (FP64 sums differ by summation order, so the
    yardstick is the field's scale, not each element's own magnitude)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    bad = ~(np.isfinite(a) & np.isfinite(b))
    if bad.any():
        if not np.array_equal(a[bad], b[bad], equal_nan=True):
            return np.inf
        a, b = a[~bad], b[~bad]
        if a.size == 0:
            return 0.0
    scale = np.abs(b).max()
    diff = np.abs(a - b).max()
    if scale == 0.0:
        return 0.0 if diff == 0.0 else np.inf
    return diff / scale


def assert_fields_close(e, o, fields, tol=TOL, level=1, context=""):
    got = e.download(tuple(fields), level)
    for f in fields:
        ref = o.get(f, level)
        if got[f].dtype.kind in "iu":
            neq = int((got[f] != ref).sum())
            assert neq == 0, "%s: integer field %s differs for %d of %d particles" % (context, f, neq, ref.size)
        else:
            r = relerr(got[f], ref)
            assert r <= tol, "%s: field %s relative error %.3e > %.1e" % (context, f, r, tol)
