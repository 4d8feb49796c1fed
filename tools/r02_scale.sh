#!/bin/bash
# default bench line at N GPUs (headline replicas + partitioned 1024^3 leg with parity), as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 2> gpurun_out/r02_scale_n${N}.err | tail -1 > gpurun_out/r02_scale_n${N}.json )
( timeout 300 $TR --master-port 29532 bench.py --gpus $N --workload ensemble --members 256 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_ensemble256_n${N}.json )
( timeout 200 $TR --master-port 29533 tests/mgpu_batch_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -5 ) > gpurun_out/r02_batch_parity_n${N}.log
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_scale_n{N}.json').read())
print("headline", d['value'], d['ms_per_step'])
p=d['partitioned']; print(json.dumps({k:p.get(k) for k in ('ms_per_step','efficiency_vs_n1','n1_ms_per_step','alltoall_gbs_achieved','nvlink_frac_of_900','parity_rel_l2','compute_ms_per_step','error')}))
print(json.dumps(p.get('exchange'))[:600])
e=json.loads(open(f'gpurun_out/r02_ensemble256_n{N}.json').read()); print("ensemble256", e['ms_per_step'], e['value'], e['step_roofline']['frac'])
print(open(f'gpurun_out/r02_batch_parity_n{N}.log').read())
PY
