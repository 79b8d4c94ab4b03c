#!/bin/bash
# A/B of compile-time variants of the pair sweeps (tools/build_variant.sh): block workload, 12.5 M particles, 3 timed steps;
# then the parity suite on the variant named by PARITY_VARIANT
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
for v in "$@"; do
  L=$PWD/fjsph_b200/lib/var_$v.so; [ $v = default ] && L=$PWD/fjsph_b200/lib/libfjsph_b200.so
  FJSPH_B200_LIB=$L timeout 400 $B > $O/k_$v.json 2> $O/k_$v.err || tail -n 3 $O/k_$v.err
  python tools/bench_summary.py $O/k_$v.json
done
if [ -n "$PARITY_VARIANT" ]; then
  FJSPH_B200_LIB=$PWD/fjsph_b200/lib/var_$PARITY_VARIANT.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -n 3
fi
