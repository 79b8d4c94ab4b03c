"""Executed-instruction mix of one kernel of an ncu report (source page): warp instructions by opcode, the FP64 share, and
the issue-cycle floor it implies when an FP64 instruction holds the dispatch port two cycles (half-rate pipe).

    python tools/instruction_mix.py report.ncu-rep <launch index, 0-based> [warp-steps of the launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::%d" % (idx + 1)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    seen, uniq = set(), []
    for r in data:
        if r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    ie, si = hdr.index("Instructions Executed"), hdr.index("Source")
    c = collections.Counter()
    for r in uniq:
        t = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip())
        c[t.split()[0].split(".")[0]] += int(r[ie])
    tot = sum(c.values())
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print("kernel: %s" % name[:110])
    print("warp instructions executed: %.4e   FP64 (DFMA, DMUL, DADD, DSETP): %.4e = %.1f %%" % (tot, fp64, 100.0 * fp64 / tot))
    for k, v in c.most_common(18):
        print("   %-10s %6.2f %%" % (k, 100.0 * v / tot))
    if steps:
        per, f = tot / steps, fp64 / steps
        print("per warp-step (two pairs per iteration, %.4e warp-steps): %.1f instructions, %.1f of them FP64" % (steps, per, f))
        slots = 2.0 * f + (per - f)
        print("issue cycles per warp-step with FP64 at two cycles: %.1f -> %.2f ms at 1.965 GHz over 592 schedulers" % (
            slots, slots * steps / 592.0 / 1.965e9 * 1e3))


if __name__ == "__main__":
    main()
