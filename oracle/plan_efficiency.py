"""CPU-side analysis of the run plan: SIMD efficiency of the run kernels for (H2O)n 6-31G**.

Schwarz maxima come from the CPU oracle (analysis tool, not a product path); buckets, groups and
segments are built exactly as pc_schwarz / pc_plan build them (the segment builder is the
library's own host function pc_plan_segments_host).  For every class it reports
    quartets / (32 * sum over warps of the longest run in the warp)
i.e. the fraction of lane-iterations of the run loop that hold a quartet.
"""
import ctypes
import sys
from collections import defaultdict

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle                       # noqa: E402
from pychem_b200 import _lib, structures as S    # noqa: E402
from pychem_b200.basis_table import BasisTable   # noqa: E402

LN = "spd"


def buckets(tb, pmax):
    ns = tb.nshell
    kinds = defaultdict(list)
    p = 0
    for a in range(ns):
        for b in range(a, ns):
            la, lb = int(tb.l[a]), int(tb.l[b])
            x, y = (a, b) if la >= lb else (b, a)
            kinds[(int(tb.l[x]), int(tb.l[y]), int(tb.K[x]) * int(tb.K[y]))].append((x, pmax[p], p))
            p += 1
    out = {}
    for key, lst in kinds.items():
        gmax = defaultdict(float)
        for x, pm, _ in lst:
            gmax[x] = max(gmax[x], pm)
        lst.sort(key=lambda e: (-gmax[e[0]], e[0], -e[1]))
        pm = np.array([e[1] for e in lst])
        xs = np.array([e[0] for e in lst])
        gstart = np.concatenate([[0], np.nonzero(np.diff(xs))[0] + 1, [len(lst)]]).astype(np.int32)
        out[key] = (np.ascontiguousarray(pm), gstart)
    return out


def segments(lib, B, K, same, run, thresh=1e-8):
    pmB, gB = B
    pmK, gK = K
    cap = 64
    while True:
        n = ctypes.c_int()
        off = np.zeros(cap + 1, dtype=np.int64)
        ij = np.zeros(2 * cap, dtype=np.int32)
        q = np.zeros(cap + 1, dtype=np.int64)
        rc = lib.pc_plan_segments_host(len(pmB), pmB.ctypes.data_as(_lib.c_dp), len(gB) - 1,
                                       gB.ctypes.data_as(_lib.c_ip), len(pmK), pmK.ctypes.data_as(_lib.c_dp),
                                       len(gK) - 1, gK.ctypes.data_as(_lib.c_ip), int(same), int(run), thresh, cap,
                                       ctypes.byref(n), off.ctypes.data_as(_lib.c_llp),
                                       ij.ctypes.data_as(_lib.c_ip), q.ctypes.data_as(_lib.c_llp))
        if rc == 0:
            return n.value, off[:n.value + 1], ij[:2 * n.value].reshape(-1, 2), q[:n.value + 1]
        cap = max(cap * 4, n.value + 1)


def main():
    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    runs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
    tb = BasisTable(S.Molecule(S.water_cluster(nw), "6-31G**"))
    _, pmax = oracle.OracleBasis(tb).schwarz()
    lib = _lib.load()
    bk = buckets(tb, pmax)
    keys = sorted(bk)
    pc = lambda k: k[0] * (k[0] + 1) // 2 + k[1]     # noqa: E731
    stats = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, 0]))   # class -> run -> [quartets, lane-iters, tasks, segs]
    for ia, ka in enumerate(keys):
        for kb in keys[ia:]:
            kB, kK = (ka, kb) if pc(ka) >= pc(kb) else (kb, ka)
            cls = LN[kB[0]] + LN[kB[1]] + LN[kK[0]] + LN[kK[1]]
            for run in runs:
                n, off, ij, q = segments(lib, bk[kB], bk[kK], ka == kb, run)
                if n == 0:
                    continue
                r = (ij[:, 0] >> 24) & 15
                total = int(off[-1])
                nwarp = (total + 31) // 32
                # longest run among the segments overlapping every warp
                seg_of_task_start = np.searchsorted(off, np.arange(nwarp) * 32, side="right") - 1
                seg_of_task_end = np.searchsorted(off, np.minimum(np.arange(nwarp) * 32 + 31, total - 1), side="right") - 1
                rmax = np.array([r[a:b + 1].max() for a, b in zip(seg_of_task_start, seg_of_task_end)])
                st = stats[cls][run]
                st[0] += int(q[-1]); st[1] += int(32 * rmax.sum()); st[2] += total; st[3] += n
    print("class      " + "".join("   R=%-2d eff  lanes/seg" % r for r in runs))
    for cls in sorted(stats, key=lambda c: -stats[c][runs[0]][0]):
        line = "%-6s %9.2e" % (cls, stats[cls][runs[0]][0])
        for run in runs:
            qn, li, tasks, segs = stats[cls][run]
            line += "   %5.1f%%  %6.1f   " % (100.0 * qn / max(li, 1), tasks / max(segs, 1))
        print(line)


if __name__ == "__main__":
    main()
