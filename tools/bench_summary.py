"""One-line summary + per-kernel-family milliseconds of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except (OSError, ValueError, IndexError) as ex:
        print("%s: unreadable (%s)" % (path, ex))
        continue
    r = d.get("roofline") or {}
    print("%s: value %.4g  ms/step %.2f  force %.2f ms  fp64 %.2f TF  clocks %s" % (
        path, d["value"], d["ms_per_step"], r.get("ms_per_launch", 0.0), (r.get("fp64") or {}).get("achieved_tflops", 0.0),
        (d.get("clocks") or {}).get("sm_mhz")))
    print("   " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in d.get("kernels", {}).items() if v["ms_per_step"] > 0.4))
