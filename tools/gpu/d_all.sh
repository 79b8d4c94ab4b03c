#!/bin/bash
# every sweep on the two-at-a-time walk: the whole GPU suite (incl. the at-size cases), block benches, ncu of every sweep kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x --durations=12 > $O/d_tests.log 2>&1
echo "tests exit $?" >> $O/d_tests.log
tail -n 25 $O/d_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
for w in 4 8; do
  FJSPH_B200_SWEEP_WARPS=$w timeout 600 $B > $O/d_w$w.json 2> $O/d_w$w.err
done
python tools/bench_summary.py $O/d_w4.json $O/d_w8.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_prestep|k_surf1_diss|k_surf23_shift|k_exact_runs|k_build_skin_runs|k_nb_update' -c 12 -o $O/d_sweeps $B --steps 1 --warmup 1 > $O/d_ncu.log 2>&1
python tools/ncu_digest.py $O/d_sweeps.ncu-rep > $O/d_digest.txt 2>&1
grep -E "^== launch|gpu__time_duration|fp64.avg|lsu_wavefronts|stalls" $O/d_digest.txt
