#!/bin/bash
# GPU-box driver: parity tests, A/B of the default library against pychem_b200/variants/lib_prev.so
# (if present), then -- with the faster one -- the default bench line, the ncu launch list of one
# build and an `ncu --set full` capture (+ SASS stall samples) of the dominant class kernel in
# digestion mode.  Outputs: gpurun_out/final/.
O=gpurun_out/final
mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
Q="python bench.py --no-cpu-baseline --steps 30 --warmup 5 --profile-classes"
timeout 300 $Q > $O/ab_new.json 2> $O/ab_new.err
BEST=""
if [ -f pychem_b200/variants/lib_prev.so ]; then
  PYCHEM_B200_LIB=$PWD/pychem_b200/variants/lib_prev.so timeout 300 $Q > $O/ab_prev.json 2> $O/ab_prev.err
  BEST=$(python - <<PY
import json
a = json.load(open("$O/ab_new.json"))["ms_per_step"]; b = json.load(open("$O/ab_prev.json"))["ms_per_step"]
print("new %.3f prev %.3f" % (a, b), file=open("$O/ab.txt", "w"))
print("" if a <= b else "$PWD/pychem_b200/variants/lib_prev.so")
PY
)
  cat $O/ab.txt
fi
[ -n "$BEST" ] && export PYCHEM_B200_LIB=$BEST && echo "profiling lib_prev" > $O/which.txt
timeout 600 python bench.py --profile-classes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['eri_generation_only'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"eri_psss_kernel<\(int\)2>" -c 2 -f -o /tmp/psss2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full2.log 2>&1; echo "full rc=$?"
ncu -i /tmp/psss2.ncu-rep --page raw --csv > $O/psss2_raw.csv 2>> $O/ncu_full2.log
python tools/ncu_source_dump.py /tmp/psss2.ncu-rep "eri_psss_kernel" >> $O/ncu_full2.log 2>&1
mv gpurun_out/src_eri_psss_kernel.csv.gz $O/src_psss_mode2.csv.gz 2>/dev/null
ls $O
