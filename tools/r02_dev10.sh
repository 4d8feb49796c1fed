#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev10.txt
( timeout 600 python -m pytest tests/test_gpu_mqg.py tests/test_golden.py tests/test_gpu_parity.py tests/test_gpu_fused.py -q -m gpu -k "dealias or mqg or golden or example or coupled" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -20 ) > $O
cat $O
