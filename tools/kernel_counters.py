"""profiles/kernel_counters.json from an `ncu --set full` report of one bench step: per kernel instantiation the
duration, the DRAM traffic and the FP64-pipe busy time the hardware counted, per particle where that makes sense.

    python tools/kernel_counters.py gpurun_out/x.ncu-rep <particles> profiles/kernel_counters.json "<how it was captured>"

bench.py scales these to the live run (same kernels, CUDA-event durations): `roofline.traffic` and every entry of
`roofline_kernels` say that they come from this profile, not from the timed run (ncu replays every kernel ~40 times).
"""
import csv
import io
import json
import re
import subprocess
import sys


def main():
    rep, n, out_path = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    how = sys.argv[4] if len(sys.argv) > 4 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]])

    kernels = {}
    for r in data:
        name = r[col["Kernel Name"]]
        m = re.search(r"(k_\w+)(<[^>]*>)?", name)
        key = (m.group(1) + (m.group(2) or "")).replace(" ", "") if m else name
        if key in kernels:
            kernels[key]["launches_captured"] += 1
            continue
        dur_ms = val(r, "gpu__time_duration.sum")  # ms in this export
        dram = (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
        units = rows[1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        dram = val(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]] + \
            val(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
        fp64_pct = val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
        cycles = val(r, "sm__cycles_elapsed.avg")
        kernels[key] = {
            "duration_ms": dur_ms, "dram_bytes": dram, "dram_bytes_per_particle": dram / n,
            "fp64_pipe_pct": fp64_pct, "fp64_pipe_busy_ms": fp64_pct / 100.0 * dur_ms,
            "sm_mhz": cycles / (dur_ms * 1e-3) / 1e6,
            "l1_wavefront_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
            "registers": int(val(r, "launch__registers_per_thread")),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "launches_captured": 1,
        }
    json.dump({"particles": n, "source": rep, "how": how, "kernels": kernels}, open(out_path, "w"), indent=1)
    for k, v in kernels.items():
        print("%-44s %7.2f ms  dram %6.0f B/particle  fp64 pipe %4.1f %%  L1 %4.1f %%  %3d regs" % (
            k[:44], v["duration_ms"], v["dram_bytes_per_particle"], v["fp64_pipe_pct"], v["l1_wavefront_pct"], v["registers"]))


if __name__ == "__main__":
    main()
