"""Registers, stack and spills of every kernel from the ptxas logs of the last build (fjsph_b200/lib/obj/*.ptxas.log)."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fjsph_b200", "lib", "obj")
for path in sorted(glob.glob(os.path.join(obj, "*.ptxas.log"))):
    log = open(path).read()
    pat = (r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
           r"(\d+) bytes spill loads\n.*?Used (\d+) registers")
    for m in re.finditer(pat, log):
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "")
        print("%-18s %-58s regs=%4s stack=%4s spill=%s/%s" % (os.path.basename(path)[:-10], name[:58], m.group(5), m.group(2),
                                                              m.group(3), m.group(4)))
