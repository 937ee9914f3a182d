#!/bin/bash
# GPU-box driver: parity tests, the default bench line, the ncu launch list of one build and an
# `ncu --set full` capture (+ SASS stall samples) of the dominant class kernel in digestion (<2>)
# and generation-only (<5>) mode.  Outputs: gpurun_out/prof/.
O=gpurun_out/prof
mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 600 python bench.py --profile-classes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-400 $O/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"eri_psss_kernel<(2|5)>" -c 6 -f -o /tmp/psss python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "full rc=$?"
ncu -i /tmp/psss.ncu-rep --page raw --csv > $O/psss_raw.csv 2>> $O/ncu_full.log
python tools/ncu_source_dump.py /tmp/psss.ncu-rep "eri_psss_kernel<2>" "eri_psss_kernel<5>" >> $O/ncu_full.log 2>&1
mv gpurun_out/src_eri_psss_kernel_*.csv.gz $O/ 2>/dev/null
ls -la $O
