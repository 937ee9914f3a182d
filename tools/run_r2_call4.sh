#!/bin/bash
# Round 2, GPU call 4: A/B of the digestion / bra-record variants (tools/build_variant.py a..f) and a
# calibration of the pure-read HBM bandwidth (what a streaming reduction can reach on this box).
O=gpurun_out/r2c4
mkdir -p $O; rm -f $O/*
V=pychem_b200/variants
timeout 1500 python tools/ab_classes.py --reps 3 --check base=$V/lib_base.so a=$V/lib_a.so b=$V/lib_b.so c=$V/lib_c.so d=$V/lib_d.so e=$V/lib_e.so f=$V/lib_f.so > $O/ab.jsonl 2> $O/ab.err; echo "ab rc=$?"
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2c4/ab.jsonl')]
rows=[r for r in rows if 'error' not in r]
names=[r['name'] for r in rows]
print('variant   wall    jk_total gen_total  dJ dX')
for r in rows: print('%-8s %7.3f %8.3f %8.3f  %.1e %.1e'%(r['name'], r['wall_ms_best'], r['jk_total_ms'], r['gen_total_ms'], r.get('max_dJ',0), r.get('max_dX',0)))
classes=sorted(rows[0]['jk_ms'], key=lambda c:-rows[0]['jk_ms'][c])
print('jk   '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['jk_ms'].get(c,0) for r in rows))
print('gen  '+' '.join('%7s'%n for n in names))
for c in classes: print('%-5s'%c+' '.join('%7.3f'%r['gen_ms'].get(c,0) for r in rows))
PY
python - <<'PY'
import torch, time
x = torch.empty(192**4, dtype=torch.float64, device='cuda').normal_()
torch.cuda.synchronize()
for name, fn in (('sum', lambda: x.sum()), ('max', lambda: x.max()), ('dot', lambda: torch.dot(x, x))):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('torch.%s over 10.9 GB: %.3f ms = %.0f GB/s' % (name, ms, x.numel() * 8 / ms / 1e6))
y = torch.empty_like(x)
for _ in range(2): y.copy_(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print('copy 10.9 GB: %.3f ms = %.0f GB/s (read+write)' % (ms, 2 * x.numel() * 8 / ms / 1e6))
PY
ls -la $O
