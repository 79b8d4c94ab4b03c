"""Digest of an ncu report (exported raw csv + source csv): per-kernel headline metrics and the stall mix.
    python tools/ncu_digest.py report.ncu-rep [kernel-index ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "smsp__inst_executed_pipe_fp64.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def stalls(rep, idx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::%d" % (idx + 1)],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out))]
    hdr = next(r for r in rows if "Address" in r[:1] or (r and r[0] == "Address"))
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    seen, uniq = set(), []
    for r in data:  # the export lists every instruction twice
        if r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0 for i in cols}
    for r in uniq:
        for i in cols:
            try:
                tot[hdr[i]] += int(r[i])
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    mix = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7])
    si = hdr.index("Warp Stall Sampling (All Samples)")
    top = sorted(uniq, key=lambda r: -int(r[si] or 0))[:6]
    return mix, [(r[1].strip()[:60], int(r[si])) for r in top], len(uniq)


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    pick = [int(a) for a in sys.argv[2:]] or range(len(rows))
    ni = hdr.index("Kernel Name")
    for k in pick:
        r = rows[k]
        print("== launch %d: %s" % (k, r[ni][:100]))
        for w in WANT:
            if w in hdr:
                print("   %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        try:
            mix, top, n = stalls(rep, k)
            print("   stalls: " + mix)
            for t in top:
                print("      %7d  %s" % (t[1], t[0]))
        except Exception as ex:  # noqa: BLE001
            print("   (no source page: %s)" % ex)


if __name__ == "__main__":
    main()
