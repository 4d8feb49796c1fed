#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
cat gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_smoke.log
