#!/usr/bin/env python
"""Static instruction mix of the primitive-quartet loop of every class kernel, from the SASS of the
built objects (pychem_b200/build/eri_*.o), next to the flop model the roofline uses.

For each class and mode (NULL = generation only, JK_RHF) the innermost loop that contains the
fundamentals' reciprocal square roots (MUFU.RSQ64H; one per Boys branch, two per copy of the loop
body) is located by its backward branch; per copy of the body the script counts all instructions,
those that occupy the FP64 pipe (DFMA, DMUL, DADD, DSETP, MUFU.RSQ64H) and the global loads.
B200 has 64 FP64 lanes per SM (the measured peak of 36.6 TFLOP/s counts 2 flop per lane and
cycle) and issues at most 128 thread-instructions per SM and cycle, so a loop body costs at least
max(fp64, instr / 2) FP64-slot equivalents per primitive quartet and the largest share of the FP64
peak the SURVEY flop model can show for it is
    ceiling = model_flop / (2 * max(fp64, instr / 2))
(static counts: both Boys branches of the fundamentals are counted although a uniform warp runs one
of them, so the low classes are a little better than listed; classes marked `rolled` keep inner
rolled loops inside the primitive loop, their per-body numbers are not per primitive quartet).
With --per-class FILE (bench.py --profile-classes output) the per-class ceilings are weighted with
the measured model flops into a whole-step ceiling.

Usage: python tools/sass_mix.py [> profiles/r1d_sass_instruction_mix.txt]
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "pychem_b200", "build")
MODEL = json.load(open(os.path.join(ROOT, "pychem_b200", "data", "flop_model.json")))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RSQ64H", "DMNMX")
INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")


def functions(obj):
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name, cur = {}, None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = cur
            name, cur = m.group(1), []
            continue
        m = INS.match(line)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        out[name] = cur
    return out


def prim_loop(instrs):
    """Smallest backward-branch interval that contains a MUFU.RSQ64H."""
    rsq = [a for a, t in instrs if "MUFU.RSQ64H" in t]
    if not rsq:
        return None
    best = None
    for a, t in instrs:
        m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and any(tgt <= r <= a for r in rsq):
            if best is None or a - tgt < best[1] - best[0]:
                best = (tgt, a)
    return best


def analyse(instrs):
    loop = prim_loop(instrs)
    if loop is None:
        return None
    body = [t for a, t in instrs if loop[0] <= a <= loop[1]]
    copies = max(1, sum("MUFU.RSQ64H" in t for t in body) // 2)
    op = lambda t: t.split()[1] if t.startswith("@") else t.split()[0]     # noqa: E731
    n = len(body)
    f = sum(op(t).startswith(FP64) for t in body)
    ld = sum(op(t).startswith(("LDG", "LDL", "LD.")) for t in body)
    st = sum(op(t).startswith(("STL", "STG")) for t in body)
    return dict(instr=n / copies, fp64=f / copies, loads=ld / copies, local_st=st / copies, copies=copies)


ROLLED = ("ddpp", "dddp", "dddd")


def main():
    classes = sorted(MODEL, key=lambda c: (MODEL[c]["L"], c))
    ceil = {}
    print("# static SASS mix of the primitive-quartet loop (per primitive quartet), sm_100a build")
    print("# class  L  model flop | mode: instr  fp64  ld  st(local) | ceiling of the model's share of the FP64 peak")
    for c in classes:
        obj = os.path.join(BUILD, "eri_%s.o" % c)
        if not os.path.exists(obj):
            continue
        fns = functions(obj)
        row = "%-5s L=%d model=%5d |" % (c, MODEL[c]["L"], MODEL[c]["flop_prim"])
        for mode, label in ((5, "gen"), (2, "jk")):
            key = [k for k in fns if ("kernelILi%dEE" % mode) in k]
            a = analyse(fns[key[0]]) if key else None
            if a is None:
                row += " %s: (no primitive loop found)" % label
                continue
            if c in ROLLED:
                row += " %s: %6.0f %6.0f %4.0f %4.0f | rolled |" % (label, a["instr"], a["fp64"], a["loads"], a["local_st"])
                continue
            ceiling = MODEL[c]["flop_prim"] / (2.0 * max(a["fp64"], a["instr"] / 2.0))
            ceil[c] = ceiling
            row += " %s: %6.0f %6.0f %4.0f %4.0f | %5.1f%% |" % (label, a["instr"], a["fp64"], a["loads"], a["local_st"], 100 * ceiling)
        print(row)
    if "--per-class" in sys.argv:
        path = sys.argv[sys.argv.index("--per-class") + 1]
        peak = 36.6e12
        tot_f = tot_t = tot_ms = tot_gen = 0.0
        print("# class  model GFLOP  ceiling  ms at ceiling | measured ms (generation only)")
        for line in open(path):
            m = re.match(r"class (\w+):\s+([0-9.]+) ms \(generation only\s+([0-9.]+) ms\)\s+(\d+) quartets\s+([0-9.]+) TFLOP/s", line)
            if not m:
                continue
            c, ms, gen, tf = m.group(1), float(m.group(2)), float(m.group(3)), float(m.group(5))
            flops = tf * 1e12 * ms * 1e-3
            ce = min(ceil.get(c, 1.0), 1.0)
            t = flops / (peak * ce) * 1e3
            print("%-5s %8.2f %6.1f%% %8.3f | %8.3f (%.3f)" % (c, flops / 1e9, 100 * ce, t, ms, gen))
            tot_f += flops; tot_t += t; tot_ms += ms; tot_gen += gen
        print("# all classes: %.1f GFLOP (model); FP64-instruction bound %.2f ms = %.1f%% of the peak under the model; "
              "measured, serialised: %.2f ms with digestion, %.2f ms generation only" %
              (tot_f / 1e9, tot_t, 100 * tot_f / (peak * tot_t * 1e-3), tot_ms, tot_gen))


if __name__ == "__main__":
    main()
