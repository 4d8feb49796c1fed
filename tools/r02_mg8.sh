#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_slab_check.py 2>&1 | tail -16 ) > $O/r02_mg${N}_slab3d_parity.log
( timeout 200 $TR --master-port 29513 bench.py --gpus $N --workload slab3d --n3 1024 --steps 5 --warmup 3 2>$O/r02_mg${N}_p2p.err | tail -1 > $O/r02_mg${N}_p2p.json )
( PTF_F3_P2P=0 timeout 200 $TR --master-port 29514 bench.py --gpus $N --workload slab3d --n3 1024 --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/r02_mg${N}_nccl.json )
tail -8 $O/r02_mg${N}_slab3d_parity.log; tail -3 $O/r02_mg${N}_p2p.err | cut -c1-400
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("p2p","nccl"):
    try:
        d=json.loads(open(f'gpurun_out/r02_mg{N}_{f}.json').read())
        print(f, "ms/step", d['ms_per_step'], "compute", d.get('compute_ms_per_step'), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, d['clocks'])
        print("   ", json.dumps(d.get('exchange'))[:900])
    except Exception as e:
        print(f, "FAILED", e)
PY
