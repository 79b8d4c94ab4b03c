#!/bin/bash
# how many particles one B200 holds: the block workload at 40 M and 55 M particles (2 timed steps each), memory use beside it
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 1"
for c in 800,250,200 1100,250,200; do
  n=${c//,/x}
  ( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits; sleep 2; done ) > $O/q_mem_$n.txt 2>/dev/null &
  MP=$!
  timeout 900 $B --cells $c > $O/q_big_$n.json 2> $O/q_big_$n.err || tail -n 5 $O/q_big_$n.err
  kill $MP
  python tools/bench_summary.py $O/q_big_$n.json
  echo "peak memory used (MiB): $(sort -n $O/q_mem_$n.txt | tail -n 1)"
done
