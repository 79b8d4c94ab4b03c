#!/bin/bash
# the default bench line with the cpu_baseline leg first; ncu --set full of the force sweep's plain and FROZEN launches
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/x_bench.json 2> $O/x_bench.err; tail -n 3 $O/x_bench.err
python tools/bench_summary.py $O/x_bench.json
python -c "
import json; d=json.loads(open('$O/x_bench.json').read().strip().splitlines()[-1]); print(d['e2e']); print({k:d['cpu_baseline'][k] for k in ('value','step_ms_min_max')})"
B="python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force -c 4 -o $O/x_force $B > $O/x_ncu.log 2>&1
python tools/ncu_digest.py $O/x_force.ncu-rep > $O/x_force_digest.txt 2>&1
grep -E "^== launch|gpu__time_duration|dram__bytes|registers|fp64.avg|lsu_wavefronts|stalls" $O/x_force_digest.txt
