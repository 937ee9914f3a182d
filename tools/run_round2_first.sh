#!/bin/bash
# GPU-box driver for the FIRST call of the next round (one `gpurun --timeout 1500 -- bash tools/run_round2_first.sh`):
#   1. the whole GPU suite (the batched-NOCI and f-shell files ran only on the CPU emulation so far),
#   2. the default bench line with per-class times,
#   3. `ncu --set full` (+ SASS stall samples) of the classes the static analysis
#      (profiles/r1d_sass_instruction_mix.txt) puts furthest from their FP64-instruction bound:
#      (dp|pp) and (dp|dp), digestion (<2>) and generation only (<5>), and of (ps|ss) for reference.
# Outputs: gpurun_out/r2first/.
O=gpurun_out/r2first
mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
timeout 600 python bench.py --profile-classes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['eri_generation_only'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['setup_seconds'])"
for cls in dppp dpdp psss; do
  for mode in 2 5; do
    timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:"eri_${cls}_kernel<\(int\)$mode>" -c 1 -f -o /tmp/$cls$mode python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_${cls}$mode.log 2>&1; echo "ncu $cls $mode rc=$?"
    ncu -i /tmp/$cls$mode.ncu-rep --page raw --csv > $O/${cls}${mode}_raw.csv 2>> $O/ncu_${cls}$mode.log
    python tools/ncu_source_dump.py /tmp/$cls$mode.ncu-rep "eri_${cls}_kernel" >> $O/ncu_${cls}$mode.log 2>&1
    mv gpurun_out/src_eri_${cls}_kernel.csv.gz $O/src_${cls}_mode$mode.csv.gz 2>/dev/null
  done
done
ls -la $O
