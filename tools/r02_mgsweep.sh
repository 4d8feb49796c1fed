#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tests/mgpu_slab_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -16 ) > $O/r02_mg${N}e_parity.log
run() { label=$1; shift
  ( env "$@" timeout 200 $TR --master-port 29513 bench.py --gpus $N --workload slab3d --n3 1024 --steps 5 --warmup 3 2>/tmp/err.txt | tail -1 > /tmp/line.json )
  python - "$label" <<'PY' >> gpurun_out/r02_mge_sweep.txt
import json,sys
try:
    d=json.loads(open('/tmp/line.json').read())
    ex=d.get('exchange') or {}
    print(sys.argv[1], "n_gpus", d['n_gpus'], "ms/step %.2f"%d['ms_per_step'], "compute %.1f"%(d.get('compute_ms_per_step') or -1), {k:round(v['ms'],2) for k,v in d['kernels'].items()}, "hidden", ex.get('hidden_frac'))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open('/tmp/err.txt').read()[-400:])
PY
}
: > gpurun_out/r02_mge_sweep.txt
shift
for v in "$@"; do run "$v" $v; done
tail -3 $O/r02_mg${N}e_parity.log; cat gpurun_out/r02_mge_sweep.txt
