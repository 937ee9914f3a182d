#!/bin/bash
# ncu --set full of the high-L class kernels (local-memory behaviour): raw pages + SASS stall samples
O=gpurun_out/hiL
mkdir -p $O; rm -f $O/*
for cls in ${CLASSES:-dppp dpdp ddpp}; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/ncu_${cls}2.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}2.ncu-rep --page raw --csv > $O/${cls}2_raw.csv 2>> $O/ncu_${cls}2.log
  python tools/ncu_source_dump.py /tmp/${cls}2.ncu-rep "eri_${cls}_kernel" >> $O/ncu_${cls}2.log 2>&1
  mv gpurun_out/src_eri_${cls}_kernel.csv.gz $O/src_${cls}_mode2.csv.gz 2>/dev/null
done
ls -la $O
