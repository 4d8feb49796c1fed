#!/bin/bash
# sanitizer pass over the round's new kernels + one A/B of an experimental build
mkdir -p gpurun_out
O=gpurun_out/r02_sanitizer.txt
: > $O
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run3.py" >> $O
  ( timeout 400 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_run3.py 2>&1 | grep -E "^3-D|^2-D|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -30 ) >> $O
done
cat $O
if [ -f passivetracerflows.jl_b200/libptf_b200_pipe.so ]; then
  for v in default pipe; do
    if [ $v = pipe ]; then export PTF_LIB_PATH=$PWD/passivetracerflows.jl_b200/libptf_b200_pipe.so; fi
    timeout 200 python bench.py --no-cpu-baseline --no-partitioned --steps 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v 4096^2', round(d['ms_per_step']/25,4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
  done
fi
