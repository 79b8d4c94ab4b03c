#!/bin/bash
# why the row sweeps are slow: L1 prefetch on/off x 4/8-warp CTAs, list fill statistics, ncu --set full of the force sweep
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2"
for w in 4 8; do
  FJSPH_B200_SWEEP_WARPS=$w timeout 600 $B > $O/b_pf_w$w.json 2> $O/b_pf_w$w.err
  FJSPH_B200_SWEEP_WARPS=$w FJSPH_B200_LIB=$PWD/fjsph_b200/lib/var_nopf.so timeout 600 $B > $O/b_nopf_w$w.json 2> $O/b_nopf_w$w.err
done
FJSPH_B200_LIST_STATS=1 FJSPH_B200_SWEEP_WARPS=4 timeout 600 $B --steps 1 --warmup 1 > $O/b_stats.json 2> $O/b_stats.err
grep "fjsph_b200\] list" $O/b_stats.err | tail -3
python tools/bench_summary.py $O/b_pf_w4.json $O/b_nopf_w4.json $O/b_pf_w8.json $O/b_nopf_w8.json
FJSPH_B200_SWEEP_WARPS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1 -c 3 -o $O/b_force_w4 $B --steps 1 --warmup 1 > $O/b_ncu.log 2>&1
ncu -i $O/b_force_w4.ncu-rep --page raw --csv > $O/b_force_w4.csv 2>/dev/null
ls -la $O | tail -5
