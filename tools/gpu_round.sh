#!/bin/bash
# One gpurun call that refreshes every piece of single-GPU round evidence: GPU parity tests, smoke, the default bench
# line (headline + partitioned leg), the ncu launch list of the bench command, `--set full` captures of the fused 2-D
# and 3-D kernels (condensed on the box: the .ncu-rep files are too large to bring back), per-config throughput.
# usage: gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r02}
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> /dev/null
timeout 600 python tools/bench_configs.py > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-partitioned > $O/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 16 -c 8 -f \
    -o $O/${TAG}_fused_4096 python tools/profile_step.py 4096 3 > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $O/${TAG}_fused_4096.ncu-rep $O/${TAG}_ncu_fused_4096 > /dev/null 2>&1
# (source-line attribution needs the object files, which do not travel: run tools/ncu_source_lines.py here afterwards)
timeout 600 ncu --set full --clock-control none -k regex:"k_fused|k_y" -s 20 -c 4 -f \
    -o /tmp/${TAG}_fused3d_512 python tools/profile3d.py 512 3 arrays > $O/${TAG}_ncu3d.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_fused3d_512.ncu-rep $O/${TAG}_ncu_fused3d_512 > /dev/null 2>&1
timeout 300 python tools/bench_mqg.py > $O/${TAG}_mqg_coupled.jsonl 2>&1          # BASELINE configs[2], flow + tracer on the device
timeout 300 python bench.py --workload ensemble --members 32 --steps 10 --warmup 3 > $O/${TAG}_ensemble32_n1.json 2>/dev/null
tail -3 $O/${TAG}_pytest_gpu.log
cat $O/${TAG}_smoke.log | tail -4
cut -c1-1500 $O/${TAG}_bench.json
