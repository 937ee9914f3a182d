#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY: build tests/emu/libpychem_b200_emu.so, a host emulation of the CUDA
library made from the SAME kernel sources (pychem_b200/csrc/*.cu, *.cuh and the generated
csrc/gen/eri_*.cu), compiled by g++ against the stand-in tests/emu/include/cuda_runtime.h.

Used by tests/test_emu_cpu.py to execute the real device code (decode, recursions, warp
reductions, digestion) without a GPU and check it against the oracle.  The product never loads
this library: pychem_b200/_lib.py only knows pychem_b200/libpychem_b200.so.

The sources are copied into tests/emu/build/src with three purely textual substitutions that
g++ cannot take as they are:
  kernel<<<grid, block, smem, stream>>>(args);  ->  pcemu::launch(grid, block, [&] { kernel(args); });
  the inline-PTX reciprocal-square-root seed    ->  pcemu::rsqrt_seed(x)
  the inline-PTX L1 prefetch                    ->  nothing
pc_mp2.cu (inline-PTX mma.sync) is not emulated; its three entry points report an error.
"""
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "pychem_b200")
CSRC = os.path.join(PKG, "csrc")
GEN = os.path.join(CSRC, "gen")
OUT = os.path.join(HERE, "build")
SRC = os.path.join(OUT, "src")
LIB = os.path.join(HERE, "libpychem_b200_emu.so")

CXX_FLAGS = ["-std=c++17", "-O1", "-fPIC", "-mfma", "-DPC_HOST_EMU=1", "-w",
             "-I", os.path.join(HERE, "include")]

LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)
RSQRT_ASM = 'asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));'
PREFETCH_ASM = 'asm volatile("prefetch.global.L1 [%0];" ::"l"(p));'

MP2_STUB = r'''
// pc_mp2.cu is not emulated (inline-PTX mma.sync): the entry points exist and fail loudly
#include "../../include/pychem_b200.h"
extern "C" {
const char* pc_mp2_last_error(void) { return "pc_mp2: not available in the host emulation"; }
int pc_mp2_energy(int, int, const double*, const double*, const double*, const double*, const double*, int, int,
                  int, double*, double*, double*) { return 1; }
int pc_dgemm_dmma(int, int, int, int, const double*, const double*, double*) { return 1; }
int pc_mp2_release(void) { return 0; }
unsigned long long pcemu_launches(void) { return pcemu::ctx().launches; }
unsigned long long pcemu_switches(void) { return pcemu::ctx().switches; }
}
'''


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(text, name):
    def repl(m):
        kern, cfg, args = m.group(1), _split_args(m.group(2)), m.group(3)
        return "pcemu::launch((unsigned)(%s), (unsigned)(%s), [&] { %s(%s); });" % (cfg[0], cfg[1], kern, args)
    # real-PTX regions that have a functional model under PC_HOST_EMU (pc_async.cuh)
    text = re.sub(r"// PC_EMU_SKIP_BEGIN.*?// PC_EMU_SKIP_END", "", text, flags=re.S)
    text, n = LAUNCH.subn(repl, text)
    if name == "pc_common.cuh":
        if RSQRT_ASM not in text or PREFETCH_ASM not in text:
            raise RuntimeError("build_emu: the inline-PTX statements of pc_common.cuh changed; update build_emu.py")
        text = text.replace(RSQRT_ASM, "y = pcemu::rsqrt_seed(x);").replace(PREFETCH_ASM, "(void)p;")
    if "<<<" in text or "asm(" in text.replace(" ", "") or "asmvolatile(" in text.replace(" ", ""):
        raise RuntimeError("build_emu: %s still holds CUDA-only syntax after the substitutions" % name)
    return text


def _write_if_changed(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path) and open(path).read() == text:
        return
    with open(path, "w") as fh:
        fh.write(text)


def _compile(src, deps_key):
    obj = os.path.join(OUT, os.path.basename(src) + ".o")
    stamp = obj + ".sha1"
    h = hashlib.sha1((deps_key + " ".join(CXX_FLAGS)).encode())
    h.update(open(src, "rb").read())
    key = h.hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    subprocess.check_call(["g++"] + CXX_FLAGS + ["-c", src, "-o", obj])
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj, True


def build(jobs=None, regenerate=True):
    if regenerate or not os.path.isdir(GEN):
        sys.path.insert(0, os.path.join(PKG, "codegen"))
        import contextlib
        import io
        import gen_eri
        with contextlib.redirect_stdout(io.StringIO()):
            gen_eri.main(GEN)
    csrc_out = os.path.join(SRC, "pychem_b200", "csrc")
    headers = ["pc_common.cuh", "pc_one_electron.cuh", "pc_generic.cuh", "pc_generic_class.h", "pc_jk_kernels.cuh",
               "pc_async.cuh", "pc_boys_table.h"]
    dep = hashlib.sha1()
    for hname in headers:
        t = transform(open(os.path.join(CSRC, hname)).read(), hname)
        dep.update(t.encode())
        _write_if_changed(os.path.join(csrc_out, hname), t)
    inc = open(os.path.join(ROOT, "include", "pychem_b200.h")).read()
    dep.update(inc.encode())
    dep.update(open(os.path.join(HERE, "include", "cuda_runtime.h"), "rb").read())
    _write_if_changed(os.path.join(SRC, "include", "pychem_b200.h"), inc)
    units = []
    t = transform(open(os.path.join(CSRC, "pc_api.cu")).read(), "pc_api.cu")
    p = os.path.join(csrc_out, "pc_api.cpp")
    _write_if_changed(p, t)
    units.append(p)
    t = transform(open(os.path.join(CSRC, "pc_generic.cu")).read(), "pc_generic.cu")
    p = os.path.join(csrc_out, "pc_generic.cpp")
    _write_if_changed(p, t)
    units.append(p)
    p = os.path.join(csrc_out, "pc_mp2_stub.cpp")
    _write_if_changed(p, '#include <cuda_runtime.h>\n' + MP2_STUB)
    units.append(p)
    for f in sorted(os.listdir(GEN)):
        if f.endswith(".cu"):
            t = transform(open(os.path.join(GEN, f)).read(), f)
            p = os.path.join(csrc_out, "gen", f[:-3] + ".cpp")
            _write_if_changed(p, t)
            units.append(p)
    units.sort(key=lambda q: -os.path.getsize(q))
    jobs = jobs or min(8, os.cpu_count() or 1)
    with ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(lambda s: _compile(s, dep.hexdigest()), units))
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB] + [o for o, _ in res])
    return LIB


if __name__ == "__main__":
    print(build())
