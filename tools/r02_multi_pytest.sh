#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab2d.py -q -m gpu --durations=8 2>&1 | tail -16 ) > gpurun_out/r02_multi_pytest.log
cat gpurun_out/r02_multi_pytest.log
