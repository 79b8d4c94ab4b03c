#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
FJSPH_GOLDEN_REPORT=1 timeout 900 python -m pytest tests/test_gpu_2d.py tests/test_gpu_mesh.py tests/test_gpu_driver.py "tests/test_golden_reference.py::test_engine_reproduces_reference_vectors" -q -s -k "2d or 2D or mesh" 2>&1 | grep -E "^E  |passed|failed|FAILED|Dam_2D|Error|assert" | head -n 70 > $O/n_2d.log; cat $O/n_2d.log
