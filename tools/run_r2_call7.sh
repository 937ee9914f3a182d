#!/bin/bash
# Round 2, GPU call 7: suite (benzene config 3, MP2 on the new GEMM, pc_jk_direct_auto), MP2 bench.
O=gpurun_out/r2c7
mkdir -p $O; rm -f $O/*
timeout 1200 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -12 $O/tests.log
timeout 600 python tools/bench_mp2.py 8 > $O/mp2.json 2> $O/mp2.err; echo "mp2 rc=$?"; cat $O/mp2.json; tail -3 $O/mp2.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sweep 32 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c7/bench.json'))
    print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pageable_inputs']['ms_per_step'], 'stored', d['stored_mode']['kernels_ms'])
except Exception as e: print('bench parse failed', e)
PY
