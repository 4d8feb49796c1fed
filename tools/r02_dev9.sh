#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev9.txt
: > $O
b2() { echo "== 2-D 4096 $*" >> $O; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-partitioned --steps 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']/25,4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})" >> $O; }
b2 A=1
b2 PTF_LIB_PATH=$PWD/passivetracerflows.jl_b200/libptf_b200_x12.so
b2 A=2
b2 PTF_LIB_PATH=$PWD/passivetracerflows.jl_b200/libptf_b200_x12.so
( timeout 600 python -m pytest tests/test_gpu_mqg.py tests/test_gpu_parity.py -x -q -k "dealias or mqg or golden or example" 2>&1 | tail -2 ) >> $O
( timeout 300 python -m pytest tests/test_golden.py -x -q -m gpu 2>&1 | tail -2 ) >> $O
cat $O
