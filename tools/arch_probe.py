"""Arch deck on its own lattice positions and on generic ones: relative error of every field against the oracle after 4 steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_decks import _arch_pair
from tests.util import relerr
for jitter in (0.0, 0.02):
    o, e = _arch_pair(jitter)
    for step in range(4):
        _, so = o.integrate(); se = e.integrate()
        print("jitter %.2f step %d iterations %d/%d dt %.3e/%.3e" % (jitter, step, se.iterations, so.iterations, se.dt, so.dt))
    F = ("xi", "rho", "p", "v", "acc", "Rrho", "lam", "vPert", "aVisc", "deltaD")
    got = e.download(F + ("surf", "surfzone"))
    print("  " + "  ".join("%s %.2e" % (f, relerr(got[f], o.get(f))) for f in F))
    print("  flags differing: surf %d surfzone %d" % ((got["surf"] != o.get("surf")).sum(), (got["surfzone"] != o.get("surfzone")).sum()))
