"""Per-kernel-family milliseconds of one bench configuration (exploration helper, GPU box only)."""
import json, subprocess, sys, os
args = sys.argv[1:]
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"),
                      "--no-cpu-baseline", "--no-e2e"] + args, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
print("value %.4g p-steps/s  ms/step %.2f  fp64 %.2f TF  lib=%s" % (d["value"], d["ms_per_step"], d["roofline"]["fp64"]["achieved_tflops"], os.environ.get("FJSPH_B200_LIB", "default")))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.4))
