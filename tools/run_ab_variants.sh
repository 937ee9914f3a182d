#!/bin/bash
# A/B of library variants on the GPU box (one gpurun call):
#   python tools/build_variant.py NAME PC_GEN_...=...      (here; the .so files travel with the snapshot)
#   gpurun -- 'VARIANTS="cur=pychem_b200/variants/lib_cur.so x=pychem_b200/variants/lib_x.so" bash tools/run_ab_variants.sh'
# per-class device times of every variant, one class launched alone (JK_RHF and generation only), J/X checked against
# the first variant; with STEP=1 also the whole-build time of bench.py for each of them.
O=gpurun_out/ab
mkdir -p $O; rm -f $O/*
timeout 1500 python tools/ab_classes.py --reps 3 ${CHECK---check} $VARIANTS > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/ab/ab.jsonl')]
rows=[r for r in rows if 'error' not in r]
names=[r['name'] for r in rows]
print('variant   wall    jk_total gen_total  dJ dX')
for r in rows: print('%-8s %7.3f %8.3f %8.3f  %.1e %.1e'%(r['name'], r['wall_ms_best'], r['jk_total_ms'], r['gen_total_ms'], r.get('max_dJ',0), r.get('max_dX',0)))
classes=sorted(rows[0]['jk_ms'], key=lambda c:-rows[0]['jk_ms'][c])
print('jk   '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['jk_ms'].get(c,0) for r in rows))
print('gen  '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['gen_ms'].get(c,0) for r in rows))
PY
if [ -n "$STEP" ]; then
  for kv in $VARIANTS; do
    PYCHEM_B200_LIB=${kv#*=} python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-stored --sweep 32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${kv%%=*} step', round(d['ms_per_step'],3), d['checks']['J_fro'], d['checks']['Xa_fro'])"
  done
fi
