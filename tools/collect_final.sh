#!/bin/bash
# gpurun_out/final (tools/run_final.sh) -> the committed evidence under profiles/ (round-2 names)
F=gpurun_out/final; P=profiles
cp $F/bench.json $P/r2_final_bench.json
cp $F/bench_reference.json $P/r2_final_bench_reference_arm.json
cp $F/launches.csv $P/r2_final_launches.csv
cp $F/bench_classes.txt $P/r2_final_per_class.txt
: > $P/r2_final_ncu_full_psss_psps_ppps_jk.txt
for c in psss psps ppps; do
  python tools/ncu_raw_summary.py $F/${c}2_raw.csv >> $P/r2_final_ncu_full_psss_psps_ppps_jk.txt
  python tools/ncu_regions.py $F/src_${c}_mode2.csv.gz > $P/r2_final_stalls_${c}_jk.txt 2>/dev/null || python tools/ncu_top_stalls.py $F/src_${c}_mode2.csv.gz > $P/r2_final_stalls_${c}_jk.txt
done
extra=""; for f in gpurun_out/hiL/dppp2_raw.csv gpurun_out/hiL/dpdp2_raw.csv gpurun_out/hiL/ddpp2_raw.csv; do [ -f $f ] && extra="$extra $f"; done
python tools/ncu_traffic_json.py $P/ncu_traffic.json "ncu --set full --clock-control none, round 2" $F/psss2_raw.csv $F/psps2_raw.csv $F/ppps2_raw.csv $extra
tail -2 $F/tests.log > $P/r2_final_gpu_tests.txt
ls -la $P/r2_final_*
