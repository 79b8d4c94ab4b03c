#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/memcheck_paths.py > $O/z_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/z_memcheck.log
tail -n 6 $O/z_memcheck.log
