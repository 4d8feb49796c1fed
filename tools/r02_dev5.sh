#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused3d.py tests/test_gpu_fused.py -x -q 2>&1 | tail -8 ) > $O/r02_dev5_fused.log
timeout 300 python bench.py --no-cpu-baseline --no-partitioned > $O/r02_dev5_bench.json 2> $O/r02_dev5_bench.err
tail -3 $O/r02_dev5_fused.log; python -c "
import json; d=json.load(open('$O/r02_dev5_bench.json')); print(d['ms_per_step']/25, {k:(round(v['ms'],4), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items()})"
