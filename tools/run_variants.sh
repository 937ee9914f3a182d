#!/bin/bash
# GPU-box experiment driver: parity tests on the default library, then bench.py per-class timings
# for library variants (generator options) and run lengths.  Results go to gpurun_out/variants/.
mkdir -p gpurun_out/variants
O=gpurun_out/variants
rm -f $O/*
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -5 $O/tests.log
B="python bench.py --no-cpu-baseline --steps 20 --warmup 3 --profile-classes"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['ms_per_step'], d.get('eri_generation_only',{}).get('ms_per_pass'), d['e2e']['value'])" 2>&1 | tail -1)"; }
run new_default X=1
run new_run2 PYCHEM_B200_RUN=2
run new_run3 PYCHEM_B200_RUN=3
run new_run6 PYCHEM_B200_RUN=6
run nofuse PYCHEM_B200_LIB=$PWD/pychem_b200/variants/lib_nofuse.so
run noouter PYCHEM_B200_LIB=$PWD/pychem_b200/variants/lib_noouter.so
run occ PYCHEM_B200_LIB=$PWD/pychem_b200/variants/lib_occ.so
