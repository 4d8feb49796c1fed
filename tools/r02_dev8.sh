#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_dev8.txt
: > $O
b3() { echo "== 3-D $*" >> $O; env "$@" timeout 300 python tools/profile3d.py $N 2 arrays time >> $O 2>&1; }
( PTF_X_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_fused3d.py -x -q 2>&1 | tail -2 ) >> $O
N=512 b3 PTF_X_DIRECT=0
N=512 b3 PTF_X_DIRECT=1
N=256 b3 PTF_X_DIRECT=1
echo "== 1024^3 direct" >> $O; PTF_X_DIRECT=1 timeout 300 python tools/profile3d.py 1024 1 time >> $O 2>&1
cat $O
