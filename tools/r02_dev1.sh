#!/bin/bash
# dev run 1 of round 2: new FFT lengths, fused 3-D engine, 2-D regression + timing
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused3d.py -x -q 2>&1 | tail -30 ) > $O/r02_dev1_fused3d.log
( timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -k "fft_core or all_steppers or rectangular or 4096_one_step or config3" 2>&1 | tail -15 ) > $O/r02_dev1_fused2d.log
( timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "auto" 2>&1 | tail -15 ) > $O/r02_dev1_parity_auto.log
timeout 300 python bench.py --no-cpu-baseline > $O/r02_dev1_bench.json 2> $O/r02_dev1_bench.err
timeout 300 python bench.py --workload slab3d --n3 256 --steps 10 --warmup 3 > $O/r02_dev1_3d_256.json 2> $O/r02_dev1_3d_256.err
timeout 300 python bench.py --workload slab3d --n3 512 --steps 5 --warmup 3 > $O/r02_dev1_3d_512.json 2> $O/r02_dev1_3d_512.err
timeout 300 python bench.py --workload slab3d --n3 1024 --steps 3 --warmup 3 > $O/r02_dev1_3d_1024.json 2> $O/r02_dev1_3d_1024.err
tail -5 $O/r02_dev1_fused3d.log; tail -3 $O/r02_dev1_fused2d.log; tail -3 $O/r02_dev1_parity_auto.log
cat $O/r02_dev1_bench.json | cut -c1-600; cat $O/r02_dev1_3d_*.json | cut -c1-700
