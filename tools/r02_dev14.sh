#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev14.txt
( timeout 600 python -m pytest tests/test_gpu_exprflow.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -20 ) > $O
cat $O
