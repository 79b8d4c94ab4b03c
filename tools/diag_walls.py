import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fjsph_b200 import cases
from tests.util import relerr, make_pair
F = ("xi","v","rho","p","acc","Rrho","aVisc","deltaD","vPert","lam","lam_nb","norm","curve","surf","surfzone","gradRho","L","colour","kernsum")
for jit in ("eps", 0.05):
  for solver in (0,1):
    case = cases.box_with_walls(n=(8, 7, 10), dx=0.01, layers=4, jitter=jit)
    o, e, p = make_pair(case, ale=1, solver_type=solver)
    nb = case["bound_points"]
    for step in range(2):
        err_o, so = o.integrate(); s = e.integrate()
        print("jit", jit, "solver", solver, "step", step, "its", s.iterations, so.iterations, "dt", s.dt, so.dt, "rms", s.rms_error, err_o)
        got = e.download(F)
        for f in F:
            ref = o.get(f)
            if got[f].dtype.kind in "iu":
                print("   %-9s neq=%d of %d" % (f, int((got[f]!=ref).sum()), ref.size))
            else:
                print("   %-9s rel=%.3e   walls %.3e  fluid %.3e" % (f, relerr(got[f], ref), relerr(got[f][:nb], ref[:nb]) if np.abs(ref[:nb]).max()>0 else -1, relerr(got[f][nb:], ref[nb:])))
