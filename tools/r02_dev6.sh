#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev6.txt
: > $O
b2() { echo "== 2-D 4096 $*" >> $O; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-partitioned --steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']/25,4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})" >> $O; }
b3() { echo "== 3-D $*" >> $O; env "$@" timeout 200 python tools/profile3d.py 512 2 arrays time >> $O 2>&1; }
b2 PTF_X_DIRECT=0
b2 PTF_X_DIRECT=1
b3 PTF_X_DIRECT=0
b3 PTF_X_DIRECT=1
PTF_X_DIRECT=1 timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused3d.py -x -q -k "all_steppers or rectangular or ensemble or layered_512 or callback" 2>&1 | tail -2 >> $O
cat $O
