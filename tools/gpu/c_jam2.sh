#!/bin/bash
# force sweep two neighbours at a time (branch-free walk, load-warmed L1): parity first, then the block bench at 4/8 warps per CTA
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/c_parity.log 2>&1
echo "parity exit $?" >> $O/c_parity.log


B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2"
for w in 4 8; do
  FJSPH_B200_SWEEP_WARPS=$w timeout 600 $B > $O/c_w$w.json 2> $O/c_w$w.err
done
tail -n 3 $O/c_parity.log
python tools/bench_summary.py $O/c_w4.json $O/c_w8.json
FJSPH_B200_SWEEP_WARPS=${NCU_WARPS:-8} timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1 -c 2 -o $O/c_force $B --steps 1 --warmup 1 > $O/c_ncu.log 2>&1
ncu -i $O/c_force.ncu-rep --page raw --csv > $O/c_force.csv 2>/dev/null; python tools/ncu_digest.py $O/c_force.ncu-rep 0 > $O/c_digest.txt 2>&1; tail -n 30 $O/c_digest.txt
