#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_2d.py -q -s -k "other_solvers or 500_steps" --durations=5 2>&1 | grep -E "^E  |passed|failed|FAILED|steps:|Error|s call" | head -n 40 > $O/s_more.log; cat $O/s_more.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/slab_check.py > $O/s_slab_check.log 2>&1
  echo "slab_check exit $?" >> $O/s_slab_check.log
  grep -E "SLAB PARITY|exit|Error|lost|differ|2D" $O/s_slab_check.log | head -20
fi
