#!/bin/bash
# (1 GPU part) MUFU accuracy probe, the arch deck against the accuracy variants, the C4 test; (2 GPU part) slab bench: native vs torch transport, overlap on / off
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
tools/bin/mufu_probe > $O/f_mufu.txt 2>&1; cat $O/f_mufu.txt
for v in default rcp2 polish; do
  L=$PWD/fjsph_b200/lib/var_$v.so; [ $v = default ] && L=$PWD/fjsph_b200/lib/libfjsph_b200.so
  echo "== arch deck with $v"; FJSPH_B200_LIB=$L timeout 300 python -m pytest tests/test_gpu_decks.py -q -k arch 2>&1 | grep -E "relative error|passed|failed" | head -3
done
timeout 600 python -m pytest tests/test_gpu_at_size.py -q -k c4 2>&1 | tail -n 5
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  B="bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-check --no-extras"
  R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
  timeout 600 $R 29511 $B > $O/f_n2_native.json 2> $O/f_n2_native.err
  FJSPH_B200_TRANSPORT=torch timeout 600 $R 29512 $B > $O/f_n2_torch.json 2> $O/f_n2_torch.err
  FJSPH_SLAB_OVERLAP=0 timeout 600 $R 29513 $B > $O/f_n2_nooverlap.json 2> $O/f_n2_nooverlap.err
  python tools/bench_summary.py $O/f_n2_native.json $O/f_n2_torch.json $O/f_n2_nooverlap.json
  timeout 900 $R 29514 tools/slab_check.py > $O/f_slab_check.log 2>&1; echo "slab_check exit $?"; tail -n 12 $O/f_slab_check.log
fi
