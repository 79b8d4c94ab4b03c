#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_decks.py tests/test_gpu_at_size.py -q -k "arch or c4" 2>&1 | grep -v "^\s*$" | grep -E "^E |passed|failed|FAILED|Error|assert" | head -n 60 > $O/h_tests.log; cat $O/h_tests.log
timeout 300 python tools/arch_probe.py > $O/h_arch_probe.txt 2>&1; cat $O/h_arch_probe.txt
