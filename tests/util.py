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


INPUT_PARAMS = ("ale pressure_rel solver_type acase asource use_lam use_TAB_def max_subits n_stable_limit n_unstable_limit "
                "particle_step H_fac rho_rest press_pipe press_back rho_max rho_min rho_var rho_max_iter visc_alpha speed_sound "
                "mu sig gam dsph_delta grav v_inf p_ref rho_g mu_g temp_g R_g gamma_g lam_cutoff i_interp_fac tab_Cf tab_Ck "
                "tab_Cd tab_Cb cfl cfl_step cfl_max cfl_min subits_factor min_residual delta_t_max delta_t_min max_shift_vel "
                "frame_time_interval").split()


def make_pair_from_deck(case, capacity=None, kind=None, **kw):
    """(oracle, engine) for a case read by fjsph_b200.frontend.read_case: same settings, particles and LIMITS blocks.
    kind: the oracle build (None = the serial parity build, "3d_mt" = the same with its particle loops threaded)."""
    P = case["params"]
    dim = case["dim"]
    params = {k: (tuple(getattr(P, k)) if hasattr(getattr(P, k), "__len__") else getattr(P, k)) for k in INPUT_PARAMS}
    params.update(kw)
    o = orc.Oracle(orc.default_params(dim, **params), kind=kind)
    o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    o.lib.orc_clear_blocks(o.h)
    for B in case["blocks"]:
        o.add_block(B["is_fluid"], B["first"], B["second"], bound_solver=B["bound_solver"], no_slip=B["no_slip"],
                    block_type=B["block_type"], fixed_vel_or_dynamic=B["fixed_vel_or_dynamic"], times=B["times"],
                    vels=B["vels"], insert_norm=B["insert_norm"], insconst=B["insconst"], delete_norm=B["delete_norm"],
                    delconst=B["delconst"], aero_norm=B["aero_norm"], aeroconst=B["aeroconst"], back=B.get("back"),
                    buffer=B.get("buffer"))
    n = case["xi"].shape[0]
    e = eng.Engine(eng.default_params(dim, **params), capacity or 4 * n)
    e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"], case["bound_points"])
    e.set_blocks(case["blocks"])
    return o, e
