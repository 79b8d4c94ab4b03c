#!/bin/bash
# final library of round 2: whole GPU suite, default bench line, reference arm, jet + droplet, ncu launch list, ncu --set full of every hot kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/y_tests.log 2>&1
echo "tests exit $?" >> $O/y_tests.log
tail -n 4 $O/y_tests.log
timeout 900 python bench.py > $O/y_bench.json 2> $O/y_bench.err; tail -n 3 $O/y_bench.err
B="python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
timeout 600 $B --workload jet > $O/y_jet.json 2> $O/y_jet.err
timeout 600 $B --workload droplet > $O/y_droplet.json 2> $O/y_droplet.err
python tools/bench_summary.py $O/y_bench.json $O/y_jet.json $O/y_droplet.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/y_launches.csv $B --steps 2 --warmup 1 > $O/y_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_prestep|k_surf1_diss|k_surf23_shift|k_exact_runs|k_build_skin_runs|k_nb_update' -c 16 -o $O/y_sweeps $B --steps 1 --warmup 1 > $O/y_ncu.log 2>&1
python tools/ncu_digest.py $O/y_sweeps.ncu-rep > $O/y_digest.txt 2>&1
grep -E "^== launch|gpu__time_duration|fp64.avg|lsu_wavefronts" $O/y_digest.txt | head -70
