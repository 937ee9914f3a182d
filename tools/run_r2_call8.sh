#!/bin/bash
# Round 2, GPU call 8: occupancy-floor variants vs the current build; bench.
O=gpurun_out/r2c8
mkdir -p $O; rm -f $O/*
V=pychem_b200/variants
timeout 1500 python tools/ab_classes.py --reps 3 --check new=$V/lib_new.so rl1_8=$V/lib_rl1_8.so rl1_7=$V/lib_rl1_7.so rl23=$V/lib_rl23.so sl23=$V/lib_sl23.so sl3_3=$V/lib_sl3_3.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2c8/ab.jsonl')]
rows=[r for r in rows if 'error' not in r]
names=[r['name'] for r in rows]
print('variant   wall    jk_total gen_total  dJ dX')
for r in rows: print('%-8s %7.3f %8.3f %8.3f  %.1e %.1e'%(r['name'], r['wall_ms_best'], r['jk_total_ms'], r['gen_total_ms'], r.get('max_dJ',0), r.get('max_dX',0)))
classes=sorted(rows[0]['jk_ms'], key=lambda c:-rows[0]['jk_ms'][c])
print('jk   '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['jk_ms'].get(c,0) for r in rows))
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 2> $O/bench.err | python -c "import json,sys; d=json.load(sys.stdin); print('bench ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'setup', d['setup_seconds'])"
