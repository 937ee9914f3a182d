#!/bin/bash
# Round 2, GPU call 9: how much of the Fock build is overlap between the class launches (side streams)
for s in 1 2 4 8 12 21; do
  PYCHEM_B200_STREAMS=$s timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stored --sweep 32 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('streams $s ms', round(d['ms_per_step'],3), 'gen', round(d['eri_generation_only']['ms_per_pass'],3))"
done
