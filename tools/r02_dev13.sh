#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev13.txt
: > $O
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused3d.py -x -q -m gpu -k "fft_core or all_steppers or rectangular or 4096_one" 2>&1 | tail -2 ) >> $O
timeout 200 python bench.py --no-cpu-baseline --no-partitioned --steps 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('4096^2', round(d['ms_per_step']/25,4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})" >> $O
timeout 300 python tools/profile3d.py 512 2 arrays time >> $O 2>&1
timeout 200 python bench.py --workload ensemble --members 32 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ens32', d['ms_per_step'], d['step_roofline']['frac'])" >> $O
cat $O
