#!/bin/bash
# GPU-box driver: `ncu --set full` capture (+ SASS stall samples) of the dominant class kernel in
# digestion (<2>) and generation-only (<5>) mode.  Outputs: gpurun_out/prof/.
O=gpurun_out/prof
mkdir -p $O
for mode in 2 5; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_psss_kernel<\(int\)$mode>" -c 2 -f -o /tmp/psss$mode python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full$mode.log 2>&1; echo "full $mode rc=$?"
  ncu -i /tmp/psss$mode.ncu-rep --page raw --csv > $O/psss${mode}_raw.csv 2>> $O/ncu_full$mode.log
  python tools/ncu_source_dump.py /tmp/psss$mode.ncu-rep "eri_psss_kernel" >> $O/ncu_full$mode.log 2>&1
  mv gpurun_out/src_eri_psss_kernel.csv.gz $O/src_psss_mode$mode.csv.gz 2>/dev/null
done
ls -la $O
