#!/bin/bash
# the driver's command at N = 4 (middle ranks have two neighbour slabs): slab parity, jet and strong-scaling side lines, e2e
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port"
timeout 900 $R 29561 bench.py --gpus $NG --steps 5 --warmup 3 > $O/o_n$NG.json 2> $O/o_n$NG.err; tail -n 4 $O/o_n$NG.err
python tools/bench_summary.py $O/o_n$NG.json
python -c "
import json; d=json.loads(open('$O/o_n$NG.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('e2e','slab','slab_parity','strong','workloads')}, indent=1)[:3500])"
