#!/bin/bash
# Round 2, GPU call 6: stored-tensor kernels with the per-CTA rotated streaming order.
O=gpurun_out/r2c6
mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests -m gpu -q -x -k "golden or stored or dropin or mp2 or noci" > $O/tests.log 2>&1; tail -2 $O/tests.log
for mode in tma old; do
  if [ $mode = old ]; then export PYCHEM_B200_STORED_NO_TMA=1; fi
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sweep 32 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$mode stored', d['stored_mode']['jk_ms'], d['stored_mode']['roofline']['frac'], 'tensor build', d['stored_mode']['tensor_build_ms'])"
done
