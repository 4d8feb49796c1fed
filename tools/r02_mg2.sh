#!/bin/bash
# 2-GPU run: multi-rank parity (3-D fused slab + cuFFT slab, 2-D slab) and the default bench line at N = 2
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29511 tests/mgpu_slab_check.py 2>&1 | tail -25 ) > $O/r02_mg2_slab3d_parity.log
( timeout 300 $TR --master-port 29512 tests/mgpu_slab2d_check.py 2>&1 | tail -12 ) > $O/r02_mg2_slab2d_parity.log
( timeout 900 $TR --master-port 29513 bench.py --gpus 2 --no-cpu-baseline > $O/r02_mg2_bench.json ) 2> $O/r02_mg2_bench.err
tail -14 $O/r02_mg2_slab3d_parity.log; tail -3 $O/r02_mg2_slab2d_parity.log; tail -5 $O/r02_mg2_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_mg2_bench.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'])
    print(json.dumps(d['partitioned'])[:3000])
except Exception as e:
    print("bench parse failed", e)
PY
