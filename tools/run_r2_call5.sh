#!/bin/bash
# Round 2, GPU call 5: suite + bench on the bra-record build, ncu of three class kernels and of the
# two stored-tensor kernels.
O=gpurun_out/r2c5
mkdir -p $O; rm -f $O/*
timeout 900 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 5 --profile-classes > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -22 $O/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c5/bench.json'))
    print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['pageable_inputs']['ms_per_step'], 'frac', d['roofline']['frac'], 'gen', d['eri_generation_only']['ms_per_pass'])
    print('checks', d['checks']); print('sweep', [(s['waters'], round(s['fock_build_ms'],3), round(s['roofline_frac'],3)) for s in d['sweep']]); print('stored', d['stored_mode']['jk_ms'], d['stored_mode']['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
PY
for cls in psss ppps dppp; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"eri_${cls}_kernel<\(int\)2>" -c 1 -f -o /tmp/${cls}2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stored --sweep 32 > $O/ncu_${cls}2.log 2>&1; echo "ncu $cls rc=$?"
  ncu -i /tmp/${cls}2.ncu-rep --page raw --csv > $O/${cls}2_raw.csv 2>> $O/ncu_${cls}2.log
  python tools/ncu_source_dump.py /tmp/${cls}2.ncu-rep "eri_${cls}_kernel" >> $O/ncu_${cls}2.log 2>&1
  mv gpurun_out/src_eri_${cls}_kernel.csv.gz $O/src_${cls}_mode2.csv.gz 2>/dev/null
done
cat > /tmp/stored_once.py <<'PY'
import numpy as np, torch, os, sys
sys.path.insert(0, os.getcwd())
from pychem_b200 import engine, structures as S
db = engine.DeviceBasis(S.Molecule(S.water_cluster(8), "6-31G**"))
db.schwarz()
G_dev, _ = db.eri_tensor(1e-8, to_host=False)
N = db.nbf
X = np.random.default_rng(1).uniform(-1, 1, (N, N)); Da = torch.from_numpy(0.5 * (X + X.T)).cuda()
for _ in range(3): db.jk_stored(G_dev, 2 * Da, Da, Da)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"jk_stored" -c 3 -f -o /tmp/stored_tma python /tmp/stored_once.py > $O/ncu_stored_tma.log 2>&1; echo "ncu stored tma rc=$?"
ncu -i /tmp/stored_tma.ncu-rep --page raw --csv > $O/stored_tma_raw.csv 2>> $O/ncu_stored_tma.log
python tools/ncu_source_dump.py /tmp/stored_tma.ncu-rep "jk_stored_tma_kernel" >> $O/ncu_stored_tma.log 2>&1
mv gpurun_out/src_jk_stored_tma_kernel.csv.gz $O/src_stored_tma.csv.gz 2>/dev/null
PYCHEM_B200_STORED_NO_TMA=1 timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"jk_stored" -c 3 -f -o /tmp/stored_old python /tmp/stored_once.py > $O/ncu_stored_old.log 2>&1; echo "ncu stored old rc=$?"
ncu -i /tmp/stored_old.ncu-rep --page raw --csv > $O/stored_old_raw.csv 2>> $O/ncu_stored_old.log
ls -la $O
