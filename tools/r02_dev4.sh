#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused3d.py -x -q 2>&1 | tail -8 ) > $O/r02_dev4_fused3d.log
( timeout 600 python -m pytest tests/test_gpu_mqg.py -x -q 2>&1 | tail -8 ) > $O/r02_dev4_mqg.log
timeout 900 python bench.py > $O/r02_dev4_bench.json 2> $O/r02_dev4_bench.err
tail -3 $O/r02_dev4_fused3d.log; tail -3 $O/r02_dev4_mqg.log; tail -3 $O/r02_dev4_bench.err; cut -c1-3000 $O/r02_dev4_bench.json
