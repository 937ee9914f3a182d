#!/bin/bash
O=gpurun_out/r2c26
mkdir -p $O; rm -f $O/*
V=pychem_b200/variants
timeout 1500 python tools/ab_classes.py --reps 3 --check cur=$V/lib_cur.so M5=$V/lib_M5.so M6=$V/lib_M6.so D8=$V/lib_D8.so D9=$V/lib_D9.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2c26/ab.jsonl')]
rows=[r for r in rows if 'error' not in r]
names=[r['name'] for r in rows]
print('variant   wall    jk_total gen_total  dJ dX')
for r in rows: print('%-8s %7.3f %8.3f %8.3f  %.1e %.1e'%(r['name'], r['wall_ms_best'], r['jk_total_ms'], r['gen_total_ms'], r.get('max_dJ',0), r.get('max_dX',0)))
classes=sorted(rows[0]['jk_ms'], key=lambda c:-rows[0]['jk_ms'][c])
print('jk   '+' '.join('%7s'%n for n in names))
for c in classes[:21]: print('%-5s'%c+' '.join('%7.3f'%r['jk_ms'].get(c,0) for r in rows))
print('gen  '+' '.join('%7s'%n for n in names))
for c in classes[:21]: print('%-5s'%c+' '.join('%7.3f'%r['gen_ms'].get(c,0) for r in rows))
PY
