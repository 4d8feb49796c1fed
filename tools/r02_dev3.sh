#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused3d.py -x -q 2>&1 | tail -8 ) > $O/r02_dev3_fused3d.log
for n in 256 512; do timeout 200 python tools/profile3d.py $n 2 time > $O/r02_dev3_k_$n.json 2>&1; done
timeout 300 python tools/profile3d.py 1024 1 time > $O/r02_dev3_k_1024.json 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_fused|k_y" -s 20 -c 4 -f \
    -o /tmp/ncu3d_512 python tools/profile3d.py 512 3 > $O/r02_dev3_ncu.log 2>&1
ncu -i /tmp/ncu3d_512.ncu-rep --page raw --csv > $O/r02_dev3_ncu3d_512_raw.csv 2>/dev/null
tail -3 $O/r02_dev3_fused3d.log; cat $O/r02_dev3_k_*.json
