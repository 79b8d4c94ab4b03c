#!/bin/bash
# first GPU pass over the row-run lists: memcheck of the smoke step, the parity suite, block benches (8- and 4-warp CTAs, round-1 library)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/a_gpu.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/a_memcheck.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/a_parity.log 2>&1
echo "parity exit $?" >> $O/a_parity.log
timeout 1200 python -m pytest tests -m gpu -q > $O/a_tests.log 2>&1
echo "tests exit $?" >> $O/a_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3"
timeout 600 $B > $O/a_bench_w8.json 2> $O/a_bench_w8.err
FJSPH_B200_SWEEP_WARPS=4 timeout 600 $B > $O/a_bench_w4.json 2> $O/a_bench_w4.err
FJSPH_B200_LIB=$PWD/fjsph_b200/lib/var_r1.so timeout 600 $B > $O/a_bench_r1.json 2> $O/a_bench_r1.err
tail -3 $O/a_memcheck.log $O/a_parity.log $O/a_tests.log
python tools/bench_summary.py $O/a_bench_w8.json $O/a_bench_w4.json $O/a_bench_r1.json
