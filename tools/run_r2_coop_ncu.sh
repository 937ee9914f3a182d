#!/bin/bash
# A/B of the cooperative variants + ncu stall samples of the cooperative (dp|pp) kernel
O=gpurun_out/coop
mkdir -p $O; rm -f $O/*
V=pychem_b200/variants
timeout 900 python tools/ab_classes.py --reps 3 --check $VARIANTS > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/coop/ab.jsonl')]
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
if [ -n "$NCU_LIB" ]; then
for cls in ${NCU_CLASSES:-dppp}; do
  PYCHEM_B200_LIB=$NCU_LIB timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_coop_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/ncu_${cls}c.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}c.ncu-rep --page raw --csv > $O/${cls}c_raw.csv 2>> $O/ncu_${cls}c.log
  python tools/ncu_source_dump.py /tmp/${cls}c.ncu-rep "eri_${cls}_coop_kernel" >> $O/ncu_${cls}c.log 2>&1
  mv gpurun_out/src_eri_${cls}_coop_kernel.csv.gz $O/src_${cls}_coop.csv.gz 2>/dev/null
  python tools/ncu_raw_summary.py $O/${cls}c_raw.csv
done
fi
