#!/bin/bash
# whole-step time (bench.py, concurrent class launches) of library variants + optional ncu of one cooperative kernel
O=gpurun_out/coopb
mkdir -p $O; rm -f $O/*
for v in $LIBS; do
  PYCHEM_B200_LIB=pychem_b200/variants/lib_$v.so timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
d=json.load(open('$O/bench_$v.json')); print('$v', 'ms_per_step', round(d['ms_per_step'],3), 'sustained', round(d['sustained']['ms_per_step'],3), 'J_fro', d['checks']['J_fro'], 'Xa_fro', d['checks']['Xa_fro'])
PY
done
if [ -n "$NCU_LIB" ]; then
for cls in ${NCU_CLASSES:-dpdp}; do
  PYCHEM_B200_LIB=$NCU_LIB timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_coop_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/ncu_${cls}c.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}c.ncu-rep --page raw --csv > $O/${cls}c_raw.csv 2>> $O/ncu_${cls}c.log
  python tools/ncu_source_dump.py /tmp/${cls}c.ncu-rep "eri_${cls}_coop_kernel" >> $O/ncu_${cls}c.log 2>&1
  mv gpurun_out/src_eri_${cls}_coop_kernel.csv.gz $O/src_${cls}_coop.csv.gz 2>/dev/null
  python tools/ncu_raw_summary.py $O/${cls}c_raw.csv
  python tools/ncu_top_stalls.py $O/src_${cls}_coop.csv.gz | head -4
done
fi
