#!/bin/bash
# exact runs sized from the skin build's measured maximum: whole GPU suite, block bench, where the end-to-end step spends its time, 40 M and 50 M particles
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/u_tests.log 2>&1
echo "tests exit $?" >> $O/u_tests.log
tail -n 4 $O/u_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
timeout 400 $B > $O/u_bench.json 2> $O/u_bench.err; python tools/bench_summary.py $O/u_bench.json
timeout 300 python tools/e2e_breakdown.py > $O/u_e2e_breakdown.txt 2>&1; cat $O/u_e2e_breakdown.txt
for c in 800,250,200 1000,250,200; do
  n=${c//,/x}
  ( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits; sleep 2; done ) > $O/u_mem_$n.txt 2>/dev/null &
  MP=$!
  timeout 900 $B --steps 2 --warmup 1 --cells $c > $O/u_big_$n.json 2> $O/u_big_$n.err || tail -n 3 $O/u_big_$n.err
  kill $MP
  python tools/bench_summary.py $O/u_big_$n.json | head -n 1
  echo "peak memory used (MiB): $(sort -n $O/u_mem_$n.txt | tail -n 1)"
done
