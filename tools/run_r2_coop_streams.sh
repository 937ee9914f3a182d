#!/bin/bash
O=gpurun_out/coops
mkdir -p $O; rm -f $O/*
for cfg in $CFGS; do
  v=${cfg%%:*}; c=${cfg##*:}
  PYCHEM_B200_STREAMS=$c PYCHEM_B200_LIB=pychem_b200/variants/lib_$v.so timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_${v}_$c.json 2> $O/bench_${v}_$c.err
  python - <<PY
import json
d=json.load(open('$O/bench_${v}_$c.json')); print('$v streams $c', 'ms_per_step', round(d['ms_per_step'],3), 'serialised/step', round(d['roofline']['serialised_kernel_ms_over_step_ms'],3))
PY
done
