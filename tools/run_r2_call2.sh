#!/bin/bash
# Round 2, GPU call 2: GPU suite on the current tree (f-shell tests at tight convergence, TMA-staged
# stored J/K), the ablation A/B of call 1 that did not find its libraries, stored-mode leg.
O=gpurun_out/r2c2
mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
V=pychem_b200/variants
timeout 1200 python tools/ab_classes.py --reps 3 --check base=$V/lib_base.so ldg256=$V/lib_ldg256.so fakeboys=$V/lib_fakeboys.so nored=$V/lib_nored.so noshfl=$V/lib_noshfl.so nodload=$V/lib_nodload.so noall=$V/lib_noall.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2c2/ab.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d['name'], d['error'][-300:]); continue
    print(d['name'], 'wall', d['wall_ms_best'], 'jk', d['jk_total_ms'], 'gen', d['gen_total_ms'])
    print('   jk ', ' '.join('%s=%.3f'%(k,v) for k,v in d['jk_ms'].items()))
    print('   gen', ' '.join('%s=%.3f'%(k,v) for k,v in d['gen_ms'].items()))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sweep 32 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c2/bench.json'))
    print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'stored', d['stored_mode'])
except Exception as e: print('bench parse failed', e)
PY
PYCHEM_B200_STORED_NO_TMA=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sweep 32 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('no-tma stored', d['stored_mode'])"
ls -la $O
