#!/bin/bash
# End-of-round evidence on the FINAL tree (one gpurun call):
#   1. the whole GPU suite
#   2. the default bench line (cpu baseline, oracle check, sweep, stored leg) -> gpurun_out/final/bench.json
#   3. ncu launch list of `bench.py --steps 1 --warmup 1` (--metrics gpu__time_duration.sum --clock-control none)
#   4. ncu --set full of the three busiest class kernels (+ SASS stall samples) and the stored kernels
O=gpurun_out/final
mkdir -p $O; rm -f $O/*
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 900 python bench.py --profile-classes > $O/bench.json 2> $O/bench_classes.txt; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
for cls in psss psps ppps; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/ncu_${cls}2.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}2.ncu-rep --page raw --csv > $O/${cls}2_raw.csv 2>> $O/ncu_${cls}2.log
  python tools/ncu_source_dump.py /tmp/${cls}2.ncu-rep "eri_${cls}_kernel" >> $O/ncu_${cls}2.log 2>&1
  mv gpurun_out/src_eri_${cls}_kernel.csv.gz $O/src_${cls}_mode2.csv.gz 2>/dev/null
done
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/final/bench.json'))
    print('ms', d['ms_per_step'], 'sustained', d['sustained'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pageable_inputs']['ms_per_step'], 'frac', d['roofline']['frac'])
    print('checks', d['checks']); print('sweep', [(s['waters'], round(s['fock_build_ms'],3), round(s['roofline_frac'],3)) for s in d['sweep']]); print('stored', d['stored_mode']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind']); print('clocks', d['clocks'])
except Exception as e: print('bench parse failed', e)
PY
ls -la $O
