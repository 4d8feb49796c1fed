#!/bin/bash
# 2-GPU sweep of the pipelined exchange: chunk count, NCCL CTA limits, copy-engine P2P
mkdir -p gpurun_out
O=gpurun_out/r02_mg2b_sweep.txt
: > $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() {  # label, env...
  label=$1; shift
  ( env "$@" timeout 300 $TR --master-port 29520 bench.py --gpus 2 --workload slab3d --n3 1024 --steps 3 --warmup 2 2>/dev/null | tail -1 > /tmp/line.json )
  python - "$label" <<'PY' >> gpurun_out/r02_mg2b_sweep.txt
import json,sys
try:
    d=json.loads(open('/tmp/line.json').read())
    ex=d.get('exchange') or {}
    print(sys.argv[1], "ms/step %.1f"%d['ms_per_step'], "compute %.1f"%(d.get('compute_ms_per_step') or -1), "exch alone/field %.2f GB/s %.0f"%(ex.get('ms_per_field_alone',-1), ex.get('alltoall_gbs_achieved',-1)), "hidden %.2f"%ex.get('hidden_frac',-1))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run "chunks=1" PTF_F3_CHUNKS=1
run "chunks=2" PTF_F3_CHUNKS=2
run "chunks=4" PTF_F3_CHUNKS=4
run "chunks=8" PTF_F3_CHUNKS=8
run "chunks=4,maxctas=8" PTF_F3_CHUNKS=4 NCCL_MAX_CTAS=8
run "chunks=4,maxctas=4" PTF_F3_CHUNKS=4 NCCL_MAX_CTAS=4
run "chunks=4,memcpy" PTF_F3_CHUNKS=4 NCCL_P2P_USE_CUDA_MEMCPY=1
run "chunks=8,memcpy" PTF_F3_CHUNKS=8 NCCL_P2P_USE_CUDA_MEMCPY=1
run "chunks=4,nograph" PTF_F3_CHUNKS=4 PTF_NO_GRAPH=1
cat $O
