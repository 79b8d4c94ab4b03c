#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_check.py > $O/i_slab_check.log 2>&1
echo "slab_check exit $?" >> $O/i_slab_check.log
grep -E "SLAB PARITY|exit|Error|lost|differ" $O/i_slab_check.log | head -20
