#!/bin/bash
# Sharded host traffic of the N>1 path: correctness (tools/check_share_ngpu.py) and bench e2e with / without it
N=${NGPU:-2}
O=gpurun_out/share$N
mkdir -p $O; rm -f $O/*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
[ -n "$SKIP_CHECK" ] || timeout 600 $TR tools/check_share_ngpu.py > $O/check.json 2> $O/check.err; echo "check rc=$?"; cat $O/check.json; tail -5 $O/check.err
k=0
for sh in ${ORDER:-1 0}; do
  k=$((k+1))
  PYCHEM_B200_SHARED_RESULTS=$sh timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_shared${sh}_$k.json 2> $O/bench_shared${sh}_$k.err
  echo "bench shared=$sh rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_shared${sh}_$k.json') if l.startswith('{')][-1])
    e=d['e2e']
    print('shared=$sh', 'ms', round(d['ms_per_step'],3), 'e2e', round(e['ms_per_step'],3), 'pageable', round(e['pageable_inputs']['ms_per_step'],3), e['host_traffic'], 'h2d/rank', e['h2d_bytes_per_step_per_rank'], 'diff', d['checks']['e2e_max_abs_diff_vs_device_path'], 'J_fro', d['checks']['J_fro'])
except Exception as ex: print('parse failed', ex); print(open('$O/bench_shared${sh}_$k.err').read()[-2500:])
PY
done
