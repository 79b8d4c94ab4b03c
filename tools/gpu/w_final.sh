#!/bin/bash
# final library of round 2: default bench line, reference arm, ncu launch list, ncu --set full of every hot kernel (incl. the FROZEN force sweep)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/w_bench.json 2> $O/w_bench.err; tail -n 3 $O/w_bench.err
timeout 900 python bench.py --impl reference > $O/w_ref.json 2> $O/w_ref.err; tail -n 3 $O/w_ref.err
python tools/bench_summary.py $O/w_bench.json $O/w_ref.json
B="python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/w_launches.csv $B --steps 2 --warmup 1 > $O/w_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_prestep|k_surf1_diss|k_surf23_shift|k_exact_runs|k_build_skin_runs|k_nb_update' -c 14 -o $O/w_sweeps $B --steps 1 --warmup 1 > $O/w_ncu.log 2>&1
python tools/ncu_digest.py $O/w_sweeps.ncu-rep > $O/w_digest.txt 2>&1
grep -E "^== launch|gpu__time_duration|fp64.avg|lsu_wavefronts|stalls" $O/w_digest.txt | head -90
