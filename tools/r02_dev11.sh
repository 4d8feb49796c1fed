#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev11.txt
( timeout 600 python -m pytest tests/test_gpu_fused3d.py -q -m gpu -k "separable or agrees" 2>&1 | tail -2 ) > $O
timeout 300 python bench.py --workload slab3d --n3 1024 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1024^3', d['ms_per_step'], d['step_roofline']['frac'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['clocks'])" >> $O
timeout 300 python bench.py --workload slab3d --n3 512 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('512^3', d['ms_per_step'], d['step_roofline']['frac'], {k:round(v['ms'],2) for k,v in d['kernels'].items()})" >> $O
cat $O
