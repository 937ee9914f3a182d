#!/bin/bash
# Round 2: scaling run on one 8-GPU box (bench.py under torchrun, N = 2, 4, 8) + benzene drop-in test.
O=gpurun_out/r2scale
mkdir -p $O; rm -f $O/*
nvidia-smi -L | wc -l
[ -z "$WITH_TESTS" ] || { timeout 600 python -m pytest tests/test_gpu_mp2.py -m gpu -q -x > $O/tests.log 2>&1; tail -3 $O/tests.log; }
# N > 1: host traffic sharded over the ranks (default); PYCHEM_B200_SHARED_RESULTS=0 for whole matrices per rank
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
[ -n "$SKIP_SHARE_CHECK" ] || timeout 600 $TR8 tools/check_share_ngpu.py > $O/check_share_8.json 2> $O/check_share_8.err; echo "share check rc=$?"; cat $O/check_share_8.json
[ -n "$SKIP_SHARE_CHECK" ] || PYCHEM_B200_SHARED_RESULTS=0 timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_8_unshared.json 2> $O/bench_8_unshared.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_8_unshared.json') if l.startswith('{')][-1]); print('8 unshared: ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pageable', round(d['e2e']['pageable_inputs']['ms_per_step'],3))
except Exception as e: print('parse failed', e)
PY
for n in 8 4 2 1; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_$n.json 2> $O/bench_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 > $O/bench_$n.json 2> $O/bench_$n.err
  fi
  echo "bench $n rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$n.json') if l.startswith('{')][-1])
    print($n, 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pageable', round(d['e2e']['pageable_inputs']['ms_per_step'],3), 'checks', d['checks']['J_fro'], d['checks']['Xa_fro'], d['checks']['J_00'], d['checks']['J_trace'])
except Exception as e: print('parse failed', e); print(open('$O/bench_$n.err').read()[-1500:])
PY
done
