#!/bin/bash
# staging the lanes' last records only (FJ_STAGE_MODE=4) against first + last (default) on the three workloads
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
for w in block jet droplet; do
  for v in default stage4; do
    L=$PWD/fjsph_b200/lib/var_$v.so; [ $v = default ] && L=$PWD/fjsph_b200/lib/libfjsph_b200.so
    FJSPH_B200_LIB=$L timeout 400 $B --workload $w > $O/r_${w}_$v.json 2> $O/r_${w}_$v.err
    python tools/bench_summary.py $O/r_${w}_$v.json
  done
done
