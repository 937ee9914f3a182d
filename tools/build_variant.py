#!/usr/bin/env python
"""Build a VARIANT of the library for A/B measurements: the generator runs with the given
PC_GEN_* switches into a scratch directory and the result is linked into
pychem_b200/variants/lib_<name>.so (git-ignored, shipped with the gpurun snapshot).  The default
library and pychem_b200/csrc/gen are not touched.

  python tools/build_variant.py bf PC_GEN_FUND=bf
  python tools/build_variant.py bf_ilp2 PC_GEN_FUND=2phase PC_GEN_ILP2_MAXL=2
  python tools/build_variant.py --emu bf PC_GEN_FUND=bf      # host emulation of the variant
                                                              # -> tests/emu/variants/lib_<name>_emu.so
Measure with tools/ab_classes.py name=pychem_b200/variants/lib_<name>.so.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pychem_b200")


def main():
    args = sys.argv[1:]
    emu = False
    if args and args[0] == "--emu":
        emu = True
        args = args[1:]
    name, env = args[0], dict(a.split("=", 1) for a in args[1:])
    nvcc_extra = env.pop("NVCC", "").split()          # e.g. NVCC="-DPC_ABL_NO_RED -DPC_BOYS_COMPACT_MINL=2"
    if not emu:
        env.setdefault("PC_GEN_SKIP_CART", "1")       # timing variants: spherical kernels only
        env.setdefault("PC_GEN_MODES", "0,2,5")       # ... in the modes the A/B harness runs
    os.environ.update(env)
    work = os.path.join("/tmp", "pychem_b200_variant_%s%s" % (name, "_emu" if emu else ""))
    csrc = os.path.join(work, "pychem_b200", "csrc")
    os.makedirs(os.path.join(csrc, "gen"), exist_ok=True)
    os.makedirs(os.path.join(work, "include"), exist_ok=True)
    for f in os.listdir(os.path.join(PKG, "csrc")):
        if f.endswith((".cu", ".cuh", ".h")):
            shutil.copy(os.path.join(PKG, "csrc", f), os.path.join(csrc, f))
    shutil.copy(os.path.join(ROOT, "include", "pychem_b200.h"), os.path.join(work, "include", "pychem_b200.h"))
    sys.path.insert(0, os.path.join(PKG, "codegen"))
    import contextlib
    import io
    import gen_eri                         # reads the PC_GEN_* switches at import
    model = os.path.join(PKG, "data", "flop_model.json")
    saved = open(model).read()
    with contextlib.redirect_stdout(io.StringIO()):
        gen_eri.main(os.path.join(csrc, "gen"))
    with open(model, "w") as fh:           # the flop model belongs to the default generator run
        fh.write(saved)
    if emu:
        sys.path.insert(0, ROOT)
        from tests.emu import build_emu
        build_emu.CSRC, build_emu.GEN = csrc, os.path.join(csrc, "gen")
        build_emu.OUT = os.path.join(work, "emu_build")
        build_emu.SRC = os.path.join(build_emu.OUT, "src")
        outdir = os.path.join(ROOT, "tests", "emu", "variants")
        os.makedirs(outdir, exist_ok=True)
        build_emu.LIB = os.path.join(outdir, "lib_%s_emu.so" % name)
        print(build_emu.build(regenerate=False))
        return
    objdir = os.path.join(work, "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
             "-diag-suppress", "177", "-diag-suppress", "550"] + nvcc_extra
    srcs = [os.path.join(csrc, "pc_api.cu"), os.path.join(csrc, "pc_mp2.cu"), os.path.join(csrc, "pc_generic.cu")] + sorted(
        os.path.join(csrc, "gen", f) for f in os.listdir(os.path.join(csrc, "gen")) if f.endswith(".cu"))
    srcs.sort(key=lambda p: -os.path.getsize(p))

    import hashlib
    hdr = hashlib.sha1(" ".join(flags).encode())
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cuh", ".h")):
            hdr.update(open(os.path.join(csrc, f), "rb").read())
    cache = os.path.join("/tmp", "pychem_b200_variant_objcache")
    os.makedirs(cache, exist_ok=True)

    def cc(src):
        # objects are shared between variants when source, headers and flags agree
        key = hashlib.sha1(hdr.hexdigest().encode() + open(src, "rb").read()).hexdigest()
        obj = os.path.join(cache, os.path.basename(src)[:-3] + "." + key + ".o")
        if not os.path.exists(obj):
            subprocess.check_call(["nvcc"] + flags + ["-c", src, "-o", obj + ".tmp"])
            os.replace(obj + ".tmp", obj)
        return obj
    with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, srcs))
    outdir = os.path.join(PKG, "variants")
    os.makedirs(outdir, exist_ok=True)
    lib = os.path.join(outdir, "lib_%s.so" % name)
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
