#!/bin/bash
# GPU-box driver: parity tests + the default bench line with per-class times -> gpurun_out/check/
O=gpurun_out/check
mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 600 python bench.py --profile-classes "$@" > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['eri_generation_only'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])"
