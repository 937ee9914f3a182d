#!/usr/bin/env python
"""A/B harness for library variants (GPU box): per-class device time of one (H2O)n 6-31G** RHF
Fock build for every shared library given on the command line, without importing torch
(start-up of a few seconds per library, so several variants fit into one short gpurun call).

  python tools/ab_classes.py [--waters 32] [--reps 3] [--check] name=path/to/lib.so ...

For each library (run in a fresh subprocess, the C ABI is loaded with ctypes):
  * JK:   pc_jk_direct (RHF variant) with profiling on -> per-class ms (serialised launches)
  * GEN:  the same schedule with the integrals discarded (PC_ERI_ONLY)
  * wall: best wall-clock of `reps` un-profiled builds (graph replay, host buffers, includes
          the 14 MB copies each way) -- the end-to-end figure
  * --check: max |J - J_first_library| and max |X - X_first_library| (same seeded density)
Writes one JSON line per library to stdout (and gpurun_out/ab_classes.jsonl when that exists).
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(lib_path, waters, reps, dump):
    os.environ["PYCHEM_B200_LIB"] = lib_path
    sys.path.insert(0, ROOT)
    import numpy as np
    from pychem_b200 import _lib, structures as S
    from pychem_b200.basis_table import BasisTable
    t0 = time.time()
    lib = _lib.load()
    tb = BasisTable(S.Molecule(S.water_cluster(waters), "6-31G**"))
    h = ctypes.c_void_p()
    ip = lambda a: a.ctypes.data_as(_lib.c_ip)      # noqa: E731
    dp = lambda a: a.ctypes.data_as(_lib.c_dp)      # noqa: E731
    _lib.check(lib.pc_basis_create(0, tb.nshell, ip(tb.l), ip(tb.K), ip(tb.is_cart), ip(tb.first_fn),
                                   dp(tb.centres), dp(tb.exps), dp(tb.scc), ctypes.byref(h)))
    _lib.check(lib.pc_schwarz(h, None, None))
    v = [ctypes.c_longlong() for _ in range(4)]
    _lib.check(lib.pc_plan(h, 1.0e-8, 0, 1, *[ctypes.byref(x) for x in v]))
    N = tb.nbf
    rng = np.random.default_rng(1234)
    X = rng.uniform(-1, 1, (N, N))
    Da = np.ascontiguousarray(0.5 * (X + X.T))
    Dt = np.ascontiguousarray(2.0 * Da)
    J, Xa, Xb = (np.empty((N, N)) for _ in range(3))
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731
    setup = time.time() - t0

    def build(variant):
        _lib.check(lib.pc_jk_direct(h, variant, vp(Dt), vp(Da), vp(Da), vp(J), vp(Xa), vp(Xb)))

    def per_class():
        n = ctypes.c_int()
        _lib.check(lib.pc_plan_items(h, 0, ctypes.byref(n), None, None, None, None, None))
        m = n.value
        cls = np.zeros((m, 4), dtype=np.int32)
        ms = np.zeros(m, dtype=np.float32)
        _lib.check(lib.pc_plan_items(h, m, ctypes.byref(n), ip(cls), None, None,
                                     ms.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), None))
        out = {}
        for c, t in zip(cls, ms):
            name = "".join("spdf"[int(x)] for x in c)
            out[name] = out.get(name, 0.0) + float(t)
        return out

    res = {"lib": lib_path, "setup_s": round(setup, 2), "quartets": v[2].value}
    build(2)                                            # warm-up (captures the graph)
    wall = []
    for _ in range(reps):
        t = time.perf_counter()
        build(2)
        wall.append((time.perf_counter() - t) * 1e3)
    res["wall_ms_best"] = round(min(wall), 3)
    _lib.check(lib.pc_set_profiling(h, 1))
    for tag, variant in (("jk", 2), ("gen", 5)):
        best = None
        for _ in range(reps):
            build(variant)
            pc = per_class()
            if best is None or sum(pc.values()) < sum(best.values()):
                best = pc
        res[tag + "_ms"] = {k: round(x, 4) for k, x in sorted(best.items(), key=lambda kv: -kv[1])}
        res[tag + "_total_ms"] = round(sum(best.values()), 3)
    _lib.check(lib.pc_set_profiling(h, 0))
    build(2)
    if dump:
        np.savez(dump, J=J, Xa=Xa)
    lib.pc_basis_destroy(h)
    print("ABRESULT " + json.dumps(res), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        child(args[1], int(args[2]), int(args[3]), args[4] if len(args) > 4 and args[4] != "-" else None)
        return
    waters, reps, check, libs = 32, 3, False, []
    it = iter(args)
    for a in it:
        if a == "--waters":
            waters = int(next(it))
        elif a == "--reps":
            reps = int(next(it))
        elif a == "--check":
            check = True
        else:
            name, _, path = a.partition("=")
            libs.append((name, os.path.abspath(path or name)))
    out_path = os.path.join(ROOT, "gpurun_out", "ab_classes.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    first = None
    with open(out_path, "a") as fh:
        for name, path in libs:
            dump = "/tmp/ab_%s.npz" % name if check else "-"
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, str(waters), str(reps), dump],
                               capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("ABRESULT ")]
            if not line:
                rec = {"name": name, "error": (p.stderr or p.stdout)[-400:]}
            else:
                rec = json.loads(line[0][9:])
                rec["name"] = name
                if check:
                    import numpy as np
                    cur = np.load(dump)
                    if first is None:
                        first = cur
                    rec["max_dJ"] = float(abs(cur["J"] - first["J"]).max())
                    rec["max_dX"] = float(abs(cur["Xa"] - first["Xa"]).max())
            s = json.dumps(rec)
            print(s, flush=True)
            fh.write(s + "\n")


if __name__ == "__main__":
    main()
