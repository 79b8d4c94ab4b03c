import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fjsph_b200 import cases, engine as eng
from oracle import oracle as orc
from tests.util import relerr
case = cases.droplet(dx=0.005)
params = dict(case["params"], delta_t_min=1e-9)
e = eng.Engine(eng.default_params(3, **params), case["xi"].shape[0], device=0)
e.upload_state(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
o = orc.Oracle(orc.default_params(3, **params))
o.set_particles(case["xi"], case["v"], case["rho"], case["p"], case["m"], case["b"])
F = ("xi","v","rho","acc","Rrho","Af","aVisc","deltaD","vPert","lam","lam_nb","norm","curve","woccl","surf","surfzone","cellID","gradRho","L","colour","kernsum")
for step in range(3):
    s = e.integrate(); _, so = o.integrate()
    print("step", step, "its", s.iterations, so.iterations, "dt", s.dt, so.dt, "npd", s.npd, so.npd)
    got = e.download(F)
    for f in F:
        ref = o.get(f)
        if got[f].dtype.kind in "iu":
            print("   %-9s neq=%d of %d" % (f, int((got[f]!=ref).sum()), ref.size))
        else:
            d = np.abs(got[f]-ref)
            if d.ndim>1: d = d.reshape(d.shape[0], -1).max(axis=1)
            print("   %-9s rel=%.3e  n(>1e-9*scale)=%d" % (f, relerr(got[f], ref), int((d > 1e-9*np.abs(ref).max()).sum())))
