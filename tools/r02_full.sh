#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1400 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 ) > $O/r02_full_pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/r02_full_smoke.log 2>&1
tail -30 $O/r02_full_pytest_gpu.log; tail -4 $O/r02_full_smoke.log
