#!/bin/bash
# Round 2, GPU call 1: the whole GPU suite (not -x), the f-shell bisect, sanitizer runs, the
# digestion/Boys ablation variants (tools/build_variant.py) and ncu captures of three class
# kernels of the current build.  Outputs: gpurun_out/r2c1/.
O=gpurun_out/r2c1
mkdir -p $O; rm -f $O/*
nvidia-smi -L > $O/host.txt; nproc >> $O/host.txt; free -g | head -2 >> $O/host.txt
timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -4 $O/tests.log
timeout 300 python oracle/diag_f_shell.py > $O/diag.log 2>&1; echo "diag rc=$?"; grep -v "^DIAG" $O/diag.log | tail -22
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python oracle/diag_f_shell.py > $O/memcheck_diag.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|Error" $O/memcheck_diag.log | head -5
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|smoke ok" $O/racecheck_smoke.log | head -5
V=pychem_b200/variants
timeout 900 python tools/ab_classes.py --reps 3 --check base=$V/lib_base.so ldg256=$V/lib_ldg256.so fakeboys=$V/lib_fakeboys.so nored=$V/lib_nored.so noshfl=$V/lib_noshfl.so nodload=$V/lib_nodload.so noall=$V/lib_noall.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2c1/ab.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d); continue
    print(d['name'], 'wall', d['wall_ms_best'], 'jk', d['jk_total_ms'], 'gen', d['gen_total_ms'], 'psss', d['jk_ms'].get('psss'), d['gen_ms'].get('psss'), 'ppps', d['jk_ms'].get('ppps'), d['gen_ms'].get('ppps'), 'dpps', d['jk_ms'].get('dpps'), d['gen_ms'].get('dpps'))
PY
for cls in psss ppps dpps; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_${cls}2.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}2.ncu-rep --page raw --csv > $O/${cls}2_raw.csv 2>> $O/ncu_${cls}2.log
  python tools/ncu_source_dump.py /tmp/${cls}2.ncu-rep "eri_${cls}_kernel" >> $O/ncu_${cls}2.log 2>&1
  mv gpurun_out/src_eri_${cls}_kernel.csv.gz $O/src_${cls}_mode2.csv.gz 2>/dev/null
done
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 --profile-classes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c1/bench.json'))
    print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pageable_inputs']['ms_per_step'], 'frac', d['roofline']['frac'])
    print('checks', d['checks']); print('sweep', d['sweep']); print('stored', d['stored_mode'])
except Exception as e: print('bench parse failed', e)
PY
ls -la $O
