#!/bin/bash
# the driver's own N = 2 command (slab parity check, jet and strong-scaling side lines, e2e) with the current library
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 1200 $R 29551 bench.py --gpus 2 --steps 5 --warmup 3 > $O/l_n2.json 2> $O/l_n2.err; tail -n 4 $O/l_n2.err
python tools/bench_summary.py $O/l_n2.json
python -c "
import json; d=json.loads(open('$O/l_n2.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('e2e','slab','slab_parity','strong','workloads')}, indent=1)[:3500])"
