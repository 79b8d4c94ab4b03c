#!/bin/bash
# after a change of a sweep's pair algebra: parity + golden + 2D suites, then the block bench
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_reference.py tests/test_gpu_2d.py tests/test_gpu_decks.py tests/test_gpu_inlet.py tests/test_gpu_mesh.py -m gpu -q -x 2>&1 | tail -n 6 > $O/p_tests.log; cat $O/p_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
timeout 400 $B > $O/p_bench.json 2> $O/p_bench.err || tail -n 3 $O/p_bench.err
python tools/bench_summary.py $O/p_bench.json
