#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_slab_check.py > gpurun_out/r02_mgpar${N}.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r02_mgpar${N}.log | head -60
