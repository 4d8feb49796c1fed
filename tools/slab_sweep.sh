#!/bin/bash
# usage: tools/slab_sweep.sh NGPU N3 "ENV1=.. ENV2=.." ...   -> prints ms/step per setting
NG=$1; N3=$2; shift 2
for spec in "$@"; do
  env $spec timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --workload slab3d --n3 $N3 --steps 5 --warmup 3 > /tmp/slab_sweep.out 2> /tmp/slab_sweep.err
  grep '^{"metric"' /tmp/slab_sweep.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$spec', 'n_gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'])" || tail -5 /tmp/slab_sweep.err
done
