#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev7.txt
: > $O
b2() { echo "== 2-D 4096 $*" >> $O; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-partitioned --steps 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']/25,4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})" >> $O; }
b3() { echo "== 3-D $*" >> $O; env "$@" timeout 200 python tools/profile3d.py 512 2 arrays time >> $O 2>&1; }
( timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused3d.py -x -q 2>&1 | tail -2 ) >> $O
b2 A=1
b2 PTF_LIB_PATH=$PWD/passivetracerflows.jl_b200/libptf_b200_nb8.so
b2 PTF_PF_STATE=1
b3 A=1
echo "== ensemble 32x1024^2" >> $O; timeout 200 python bench.py --workload ensemble --members 32 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['step_roofline']['frac'])" >> $O
echo "== 1024^3" >> $O; timeout 300 python tools/profile3d.py 1024 1 time >> $O 2>&1
cat $O
