#!/bin/bash
# Fock-build time and parity (checksums + sampled oracle rows) against the primitive-pair cut-off
O=gpurun_out/primeps
mkdir -p $O; rm -f $O/*
for e in $EPS; do
  PYCHEM_B200_PRIM_EPS=$e timeout 600 python bench.py --steps 40 --warmup 5 --no-stored --sweep 32 > $O/bench_$e.json 2> $O/bench_$e.err
  python - <<PY
import json
d=json.load(open('$O/bench_$e.json')); c=d['checks']; o=c.get('oracle') or {}
print('eps $e ms', round(d['ms_per_step'],3), 'J_fro %.15g Xa_fro %.15g' % (c['J_fro'], c['Xa_fro']), 'oracle dJ', o.get('max_abs_diff_J'), 'dXa', o.get('max_abs_diff_Xa'), 'gflop', round(d['roofline']['algorithmic_gflop_per_step'],2))
PY
done
