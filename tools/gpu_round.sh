#!/bin/bash
# One gpurun call that refreshes every piece of round evidence: GPU parity tests, smoke, the default bench line,
# the ncu launch list of the bench command, one `--set full` capture of the two fused kernels, per-config throughput.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python tools/bench_configs.py > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 16 -c 8 -f \
    -o $O/${TAG}_fused_4096 python tools/profile_step.py 4096 3 > $O/${TAG}_ncu_full.log 2>&1
timeout 300 python tools/bench_mqg.py > $O/${TAG}_mqg_coupled.jsonl 2>&1          # BASELINE configs[2], flow + tracer on the device
timeout 300 python bench.py --workload slab2d --n2 16384 --steps 5 --warmup 3 > $O/${TAG}_slab2d_n1.json 2>/dev/null
tail -3 $O/${TAG}_pytest_gpu.log
cat $O/${TAG}_smoke.log | tail -3
cat $O/${TAG}_bench.json
