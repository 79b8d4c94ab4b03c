#!/bin/bash
# rows along y / z instead of x on one GPU (block workload 500 x 250 x 100): is the slab mode's transverse row axis slower by itself?
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
for a in 1 2; do
  FJSPH_B200_ROW_AXIS=$a FJSPH_B200_LIST_STATS=1 timeout 400 $B > $O/m_axis$a.json 2> $O/m_axis$a.err; grep "list:" $O/m_axis$a.err | tail -n 1
  python tools/bench_summary.py $O/m_axis$a.json
done
