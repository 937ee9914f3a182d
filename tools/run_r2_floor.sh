#!/bin/bash
# Which classes does the concurrent Fock build rest on?  Step time with only the classes lo <= L <= hi launched
# (PYCHEM_B200_DEBUG_LRANGE; results incomplete by construction, timing only).
O=gpurun_out/floor
mkdir -p $O; rm -f $O/*
for r in $RANGES; do
  PYCHEM_B200_DEBUG_LRANGE=$r timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_$r.json 2> $O/bench_$r.err
  python - <<PY
import json
d=json.load(open('$O/bench_$r.json')); print('L in [$r]', 'ms_per_step', round(d['ms_per_step'],3))
PY
done
