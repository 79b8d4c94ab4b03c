#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python tools/second_engine_probe.py > $O/g_second_engine.txt 2>&1; cat $O/g_second_engine.txt | tail -n 8
timeout 600 python -m pytest tests/test_gpu_decks.py tests/test_gpu_at_size.py -q -k "arch or c4" 2>&1 | tail -n 30 > $O/g_tests.log; cat $O/g_tests.log
