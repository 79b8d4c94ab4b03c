"""Golden vectors of the particle tracker from FJSPH's own sources: IPT::Integrate (IPT.cpp), FindFace (Containment.cpp),
Cross_Plane / MollerTrumbore / RayNormalIntersection (Geometry.cpp), compiled unmodified in oracle/_ref
(oracle/Makefile.ref).  Run where /root/reference exists:  python tests/golden/make_ipt_vectors.py
Writes tests/golden/ipt_<case>.npz: the mesh, the hand-off records, the settings, and what the reference made of them
(last state, record counts, time records, success / failure tallies)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fjsph_b200 import engine as eng  # noqa: E402  (host-side helper only: the longest face edge of the mesh)
from oracle import oracle as orc  # noqa: E402
from tests import ipt_case  # noqa: E402


def main():
    orc.build_ref()
    for name in ipt_case.CASES:
        case = ipt_case.build(name)
        dim = case["dim"]
        p = orc.default_params(dim, asource=1, particle_step=case["particle_step"])
        o = orc.Oracle(p, kind="ref2d" if dim == 2 else "ref3d")
        o.set_mesh(case["mesh"])
        settings = dict(case["settings"], max_length=case["length_factor"] * ipt_case.longest_edge(case["mesh"], dim))
        S = orc.ipt_settings(p, **settings)
        start = ipt_case.start_records(case, orc.IPT_START, p.sim_mass)
        out = o.ipt_integrate(S, start, record_cap=ipt_case.RECORD_CAP)
        assert out["n_records"].max() <= ipt_case.RECORD_CAP
        meta = dict(dim=dim, particle_step=case["particle_step"], settings=settings, n_success=out["n_success"],
                    n_failed=out["n_failed"], record_cap=ipt_case.RECORD_CAP)
        arrays = {"mesh_" + k: v for k, v in case["mesh"].items()}
        arrays.update(start=start, last=out["last"], n_records=out["n_records"], records=out["records"],
                      meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        path = os.path.join(ROOT, "tests", "golden", "ipt_%s.npz" % name)
        np.savez_compressed(path, **arrays)
        print("%-22s %3d success %3d failed, longest record %2d -> %s (%d KB)" % (
            name, out["n_success"], out["n_failed"], out["n_records"].max(), os.path.relpath(path, ROOT), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
