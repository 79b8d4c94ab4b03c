#!/bin/bash
# round-2 final evidence with the current library: whole GPU suite, default bench line, reference arm, the three workloads, ncu launch list, ncu --set full of every hot kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/t_tests.log 2>&1
echo "tests exit $?" >> $O/t_tests.log
tail -n 14 $O/t_tests.log
timeout 900 python bench.py > $O/t_bench.json 2> $O/t_bench.err; tail -n 3 $O/t_bench.err
timeout 900 python bench.py --impl reference > $O/t_ref.json 2> $O/t_ref.err; tail -n 3 $O/t_ref.err
B="python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
timeout 600 $B --workload jet > $O/t_jet.json 2> $O/t_jet.err
timeout 600 $B --workload droplet > $O/t_droplet.json 2> $O/t_droplet.err
python tools/bench_summary.py $O/t_bench.json $O/t_ref.json $O/t_jet.json $O/t_droplet.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/t_launches.csv $B --steps 2 --warmup 1 > $O/t_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_prestep|k_surf1_diss|k_surf23_shift|k_exact_runs|k_build_skin_runs|k_nb_update' -c 12 -o $O/t_sweeps $B --steps 1 --warmup 1 > $O/t_ncu.log 2>&1
python tools/ncu_digest.py $O/t_sweeps.ncu-rep > $O/t_digest.txt 2>&1
grep -E "^== launch|gpu__time_duration|fp64.avg|lsu_wavefronts|stalls" $O/t_digest.txt | head -80
