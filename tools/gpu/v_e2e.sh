#!/bin/bash
# three-part upload of fjsph_step_host: the round-trip tests, the whole GPU suite, then the default bench line (e2e)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_edge_cases.py -q -x 2>&1 | tail -n 5
timeout 1500 python -m pytest tests -m gpu -q -x > $O/v_tests.log 2>&1
echo "tests exit $?" >> $O/v_tests.log
tail -n 4 $O/v_tests.log
timeout 900 python bench.py --no-cpu-baseline > $O/v_bench.json 2> $O/v_bench.err; tail -n 3 $O/v_bench.err
python tools/bench_summary.py $O/v_bench.json
python -c "
import json; d=json.loads(open('$O/v_bench.json').read().strip().splitlines()[-1]); print(d['e2e'])"
