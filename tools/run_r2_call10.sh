#!/bin/bash
# Round 2, GPU call 10: L1 carve-out preference and lower occupancy floors for the run kernels.
O=gpurun_out/r2c10
mkdir -p $O; rm -f $O/*
for c in -1 0 25 50; do
  PYCHEM_B200_CARVEOUT=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('carveout $c ms', round(d['ms_per_step'],3), 'gen', round(d['eri_generation_only']['ms_per_pass'],3))"
done
V=pychem_b200/variants
timeout 1500 python tools/ab_classes.py --reps 3 --check cur=pychem_b200/libpychem_b200.so rl1_5=$V/lib_rl1_5.so rl1_4=$V/lib_rl1_4.so rl0_6=$V/lib_rl0_6.so rl2_4=$V/lib_rl2_4.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2c10/ab.jsonl')]
rows=[r for r in rows if 'error' not in r]
names=[r['name'] for r in rows]
print('variant   wall    jk_total gen_total  dJ dX')
for r in rows: print('%-8s %7.3f %8.3f %8.3f  %.1e %.1e'%(r['name'], r['wall_ms_best'], r['jk_total_ms'], r['gen_total_ms'], r.get('max_dJ',0), r.get('max_dX',0)))
classes=sorted(rows[0]['jk_ms'], key=lambda c:-rows[0]['jk_ms'][c])
print('jk   '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['jk_ms'].get(c,0) for r in rows))
PY
