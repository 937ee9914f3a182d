#!/bin/bash
# Round 2, GPU call 3: new digestion (shared-memory segmented reduction), bra records + prefetch,
# TMA-staged stored J/K with Dt in the pipeline: tests, per-class A/B against the round-1 kernels.
O=gpurun_out/r2c3
mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
V=pychem_b200/variants
timeout 1200 python tools/ab_classes.py --reps 3 --check base=$V/lib_base.so new=pychem_b200/libpychem_b200.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2c3/ab.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d['name'], d['error'][-300:]); continue
    print(d['name'], 'wall', d['wall_ms_best'], 'jk', d['jk_total_ms'], 'gen', d['gen_total_ms'], 'dJ', d.get('max_dJ'), 'dX', d.get('max_dX'))
    print('   jk ', ' '.join('%s=%.3f'%(k,v) for k,v in d['jk_ms'].items()))
    print('   gen', ' '.join('%s=%.3f'%(k,v) for k,v in d['gen_ms'].items()))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sweep 32 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c3/bench.json'))
    print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'stored', d['stored_mode']['jk_ms'], d['stored_mode']['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
PY
ls -la $O
