#!/usr/bin/env python
"""Build libpychem_b200.so in-tree with nvcc for sm_100a.

Same idea as the reference's Methods/setup.py (an explicit source list compiled in place next to
the Python that loads it, `install.sh:1-3`), with nvcc instead of distutils:
  1. pychem_b200/codegen/gen_eri.py writes the per-class kernels into csrc/gen/
  2. every .cu is compiled to an object with
       nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
     (objects are cached by source hash, compiled in parallel)
  3. objects are linked into pychem_b200/libpychem_b200.so (git-ignored, shipped with the snapshot)
  4. pychem_b200/setup_c_ints.py (setuptools Extension) builds the host-side `_c_ints` module
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpychem_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177", "-diag-suppress", "550"]
NVCC_FLAGS += os.environ.get("PYCHEM_B200_NVCC_EXTRA", "").split()     # experiments only


def _hash(paths, extra=""):
    h = hashlib.sha1(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src, deps, verbose):
    name = os.path.splitext(os.path.basename(src))[0]
    obj = os.path.join(OBJ, name + ".o")
    stamp = obj + ".sha1"
    key = _hash([src] + deps, " ".join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = ["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj, True


def build(verbose=False, jobs=None):
    sys.path.insert(0, os.path.join(HERE, "codegen"))
    import gen_eri
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        gen_eri.main(GEN)
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, "pc_common.cuh")]
    gen_deps = deps + [os.path.join(CSRC, "pc_generic.cuh"), os.path.join(CSRC, "pc_generic_class.h")]
    srcs = [os.path.join(CSRC, "pc_api.cu"), os.path.join(CSRC, "pc_mp2.cu"), os.path.join(CSRC, "pc_generic.cu")] + sorted(
        os.path.join(GEN, f) for f in os.listdir(GEN) if f.endswith(".cu"))
    api_deps = deps + [os.path.join(HERE, "..", "include", "pychem_b200.h"), os.path.join(CSRC, "pc_one_electron.cuh"),
                       os.path.join(CSRC, "pc_generic_class.h"), os.path.join(CSRC, "pc_jk_kernels.cuh")]
    # biggest files first so the pool stays busy
    srcs.sort(key=lambda p: -os.path.getsize(p))
    jobs = jobs or min(8, os.cpu_count() or 1)
    with ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(lambda s: _compile(s, api_deps if (s.endswith("pc_api.cu") or s.endswith("pc_mp2.cu")) else
                                              (gen_deps if s.endswith("pc_generic.cu") else deps), verbose), srcs))
    # the host-side `_c_ints` module (legacy entry points), built like the reference's own extension
    sys.path.insert(0, HERE)
    import setup_c_ints
    setup_c_ints.build_inplace()
    objs = [o for o, _ in res]
    if any(changed for _, changed in res) or not os.path.exists(LIB):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd[:8]), "...", flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
