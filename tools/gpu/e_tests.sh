#!/bin/bash
# the whole GPU suite (after the C3 / C4 test fixes); on a 2-GPU box also the slab parity and the native NCCL driver test
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > $O/e_tests.log 2>&1
echo "tests exit $?" >> $O/e_tests.log
tail -n 40 $O/e_tests.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  B="bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 $B > $O/e_bench_n2.json 2> $O/e_bench_n2.err
  tail -n 5 $O/e_bench_n2.err
  python tools/bench_summary.py $O/e_bench_n2.json
  python -c "
import json; d=json.loads(open('$O/e_bench_n2.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('slab','slab_parity','strong','workloads')}, indent=1)[:3000])"
fi
