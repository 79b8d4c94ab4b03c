"""Where a kernel's warps wait, from the SASS page of an ncu report (warp-state samples per instruction): the share of every
stall reason over the whole kernel, by opcode of the instruction the warp was waiting to issue, and the instructions that
collect the most samples.

    python tools/stall_profile.py report.ncu-rep <launch index, 0-based> [top N]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::%d" % (idx + 1)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = next(r for r in rows if r and r[0] == "Address")
    data, seen = [], set()
    for r in rows:
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    col = {h: hdr.index(h) for h in reasons}
    si, ns = hdr.index("Source"), hdr.index("# Samples")
    num = lambda x: int(x) if x.strip() else 0
    total = sum(num(r[ns]) for r in data)
    by_reason = collections.Counter()
    by_op = collections.defaultdict(collections.Counter)
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip()).split()[0].split(".")[0]
        for h in reasons:
            v = num(r[col[h]])
            by_reason[h] += v
            by_op[op][h] += v
    print("kernel: %s" % name[:110])
    print("warp-state samples: %d" % total)
    print("by reason: " + ", ".join("%s %.1f %%" % (h[6:], 100.0 * v / total) for h, v in by_reason.most_common(8)))
    print("by the opcode waiting to issue (share of all samples; its two main reasons):")
    for op, c in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:10]:
        tot = sum(c.values())
        print("   %-8s %5.1f %%   %s" % (op, 100.0 * tot / total, ", ".join("%s %.1f" % (h[6:], 100.0 * v / total) for h, v in c.most_common(2))))
    print("instructions with the most samples:")
    for r in sorted(data, key=lambda r: -num(r[ns]))[:top]:
        c = collections.Counter({h: num(r[col[h]]) for h in reasons})
        print("   %s  %5.2f %%  %-60s %s" % (r[0][-5:], 100.0 * num(r[ns]) / total, r[si].strip()[:60],
                                             ", ".join("%s %d %%" % (h[6:], round(100.0 * v / max(num(r[ns]), 1))) for h, v in c.most_common(2))))


if __name__ == "__main__":
    main()
