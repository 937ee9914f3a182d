#!/usr/bin/env python
"""bench.py -- FP64 ERIs/s and Fock-build ms per SCF iteration (BASELINE.json metric).

Workload (config.workload): integral-direct RHF Fock build (J and K from the density; ERIs
regenerated every build, no N^4 store) of the synthetic (H2O)_n cluster in 6-31G** named by
BASELINE.json (n = 32, N = 768 basis functions; `--waters n` selects another size).  One "step" =
one Fock build = every Schwarz-surviving unique shell quartet generated and digested once,
partial J/K all-reduced over ranks (NCCL), finalised.

  value  = ERIs (contracted spherical integrals in surviving unique quartets) per second,
           whole job over all ranks, densities and outputs resident in HBM
  e2e    = same metric through the reference-facing plugin call
           pychem_b200.hartree_fock.make_coulomb_exchange_matrices(molecule, state) with HOST
           (pinned) density buffers in and J/K host arrays out, copies inside the timed region
  roofline = FP64-FMA roofline: algorithmic flops (SURVEY.md 8(d) model evaluated on the
           generator's recursion DAG, pychem_b200/data/flop_model.json, + 2 flops per digestion
           FMA) / CUDA-event time / DFMA peak measured live by pc_fp64_peak
  cpu_baseline = the reference's own integrals.two_electron (oracle/_ref: its C extension +
           Python driver) on a bounded sample of the same surviving quartets, 1 host core.

`--impl reference` times that reference CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64_eris_per_sec"
UNIT = "ERI/s"
THRESH = 1.0e-8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--waters", type=int, default=int(os.environ.get("PYCHEM_BENCH_WATERS", "32")))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-classes", action="store_true", help="print per-class device times to stderr")
    return ap.parse_args()


def workload_name(n):
    return "(H2O)%d 6-31G** integral-direct RHF Fock build (J+K), N=%d, threshold 1e-8" % (n, 24 * n)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline
# ------------------------------------------------------------------------------------------------
def reference_sample(n_waters, seconds, seed=1234):
    """Time the reference's own two-electron path on a bounded, seeded sample of the workload's
    Schwarz-surviving unique shell quartets.  Returns dict(value=ERI/s, ...).

    Uses oracle/_ref (the reference's C extension + its Python driver, built by
    oracle/build_ref.py); falls back to the C port (oracle/eri_oracle.c) when that is absent."""
    from oracle import ref_driver
    from pychem_b200 import structures as S
    coords = S.water_cluster(n_waters)
    rng = np.random.default_rng(seed)
    t_build0 = time.time()
    if ref_driver.available():
        kind = "reference"
        ns = ref_driver.modules()
        st = ns.structures
        # geometry/basis objects from the reference's own classes; ShellPairs built on demand
        # (the reference builds all 73 920 eagerly, Util/structures.py:510-523 -- not timed here)
        atoms = [st.Atom(i, row, "631GSS", [], sys.maxsize, S.TO_BOHR) for i, row in enumerate(coords)]
        shells = []
        count = 0
        for atom in atoms:
            for cg in atom.Basis:
                shells.append((atom.Coordinates, cg, list(range(count, count + cg.NAngMom))))
                count += cg.NAngMom
        nshell = len(shells)
        cache = {}

        def pair(a, b):
            if (a, b) not in cache:
                (ca, ga, va), (cb, gb, vb) = shells[a], shells[b]
                cache[(a, b)] = st.ShellPair(ca, ga, a, va, cb, gb, b, vb)
            return cache[(a, b)]

        def quartet(a, b, c, d):
            return ns.integrals.two_electron(pair(a, b), pair(c, d), 0, -1.0)
    else:
        kind = "port"
        from oracle import oracle
        from pychem_b200.basis_table import BasisTable
        ob = oracle.OracleBasis(BasisTable(S.Molecule(coords, "6-31G**")))
        nshell = ob.nshell

        def quartet(a, b, c, d):
            return ob.quartet(a, b, c, d)

    bound_cache = {}

    def bound(a, b):
        if (a, b) not in bound_cache:
            # the reference turns numpy FP warnings into exceptions globally at import
            # (hf_extensions/diis.py:5); far-apart pairs underflow harmlessly
            with np.errstate(all="ignore"):
                blk = np.asarray(quartet(a, b, a, b))
            na, nb = blk.shape[0], blk.shape[1]
            bound_cache[(a, b)] = float(max(np.sqrt(abs(blk[m, n, m, n])) for m in range(na) for n in range(nb)))
        return bound_cache[(a, b)]

    # seeded sample of unique pair-of-pairs; screening with the reference's own test
    # (hartree_fock.py:293-294).  Bounds are computed (untimed) with the same reference code.
    sample = []
    tries = 0
    while len(sample) < 40000 and tries < 2000000:
        tries += 1
        a, b, c, d = (int(x) for x in rng.integers(0, nshell, 4))
        a, b = min(a, b), max(a, b)
        c, d = min(c, d), max(c, d)
        if float(bound(a, b)) * float(bound(c, d)) > THRESH:
            sample.append((a, b, c, d))
        if time.time() - t_build0 > 2 * seconds and len(sample) >= 500:
            break
    t0 = time.perf_counter()
    n_eri = 0
    n_q = 0
    done = False
    while not done:                     # cycle through the sample until the time budget is used
        for (a, b, c, d) in sample:
            with np.errstate(all="ignore"):
                blk = quartet(a, b, c, d)
            n_eri += int(np.asarray(blk).size)
            n_q += 1
            if time.perf_counter() - t0 > seconds:
                done = True
                break
    dt = time.perf_counter() - t0
    return {"value": n_eri / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d evaluations of surviving unique shell quartets (%d ERIs) drawn uniformly (seed %d) from the "
                      "(H2O)%d 6-31G** quartet list, %s two_electron per quartet, %.1f s on 1 core; "
                      "survival rate of the draw %.3f"
                      % (n_q, n_eri, seed, n_waters,
                         "reference integrals.two_electron (oracle/_ref)" if kind == "reference" else "oracle C port",
                         dt, len(sample) / max(tries, 1)),
            "quartets_per_sec": n_q / dt, "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(args.cpu_seconds, 60.0 / max(args.steps + args.warmup, 1)))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = reference_sample(args.waters, per_step, seed=1234 + i)
        if i >= args.warmup:
            vals.append(last)
    value = float(np.mean([v["value"] for v in vals]))
    ms = 1e3 * float(np.mean([v["seconds"] for v in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.waters), "sample_per_step": last["sample"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class _Mat:
    pass


class _State:
    """Attribute shape make_coulomb_exchange_matrices needs (ElectronicState / CoDensityState)."""

    def __init__(self, Dt, Da, Db):
        self.Total, self.Alpha, self.Beta = _Mat(), _Mat(), _Mat()
        self.Total.Density, self.Alpha.Density, self.Beta.Density = Dt, Da, Db


def algorithmic_flops(db, variant):
    """Flops of one Fock build, SURVEY 8(d) model (pychem_b200/data/flop_model.json):
    primitive quartets * flop_prim + quartets * (flop_cont + digestion).  `executed` counts the
    primitive quartets the kernels actually visit (after the primitive-pair cut-off); `reference`
    counts every primitive quartet the reference would loop over (K_a K_b K_c K_d per quartet).
    The roofline fraction is computed from `executed` (the conservative one)."""
    with open(os.path.join(ROOT, "pychem_b200", "data", "flop_model.json")) as fh:
        model = json.load(fh)
    cls, kprim, tasks, ms = db.plan_items()
    prim_exec = db.prim_exec
    names = "spd"
    digest_fma = {2: 2 + 4, 3: 2 + 8, 4: 2 + 16}[variant]       # J + K updates per ERI value
    mine = total = total_ref = 0.0
    per_class = {}
    for (l1, l2, l3, l4), (kb, kk), (tot, cnt), t, pe in zip(cls, kprim, tasks, ms, prim_exec):
        name = names[l1] + names[l2] + names[l3] + names[l4]
        m = model[name]
        per_q = m["flop_cont"] + 2 * digest_fma * m["nsph"]
        f_tot = pe * m["flop_prim"] + per_q * tot
        share = cnt / tot if tot else 0.0
        mine += f_tot * share
        total += f_tot
        total_ref += (kb * kk * m["flop_prim"] + per_q) * tot
        e = per_class.setdefault(name, {"flops": 0.0, "ms": 0.0, "quartets": 0})
        e["flops"] += f_tot * share
        e["ms"] += float(t)
        e["quartets"] += int(cnt)
    return mine, total, per_class, total_ref


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pychem_b200 import dist as pdist, engine, hartree_fock as hf_gpu, structures as S

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the ONE JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    rank, world, local = pdist.init("nccl" if args.gpus > 1 else None)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n = args.waters
    mol = S.Molecule(S.water_cluster(n), "6-31G**")
    N = mol.NOrbitals
    t_setup = time.perf_counter()
    db = engine.DeviceBasis(mol, device=local)
    t_basis = time.perf_counter() - t_setup
    db.schwarz()
    t_schwarz = time.perf_counter() - t_setup - t_basis
    counts = db.plan(THRESH, rank, world)
    t_plan = time.perf_counter() - t_setup - t_basis - t_schwarz
    variant = engine.RHF

    # synthetic symmetric density: D = (X + X^T)/2, X ~ U(-1,1), seed 1234
    rng = np.random.default_rng(1234)
    X = rng.uniform(-1, 1, (N, N))
    Da_h = 0.5 * (X + X.T)
    Dt_h = 2.0 * Da_h
    Dt_d = torch.from_numpy(Dt_h).to(dev)
    Da_d = torch.from_numpy(Da_h).to(dev)
    J_d = torch.empty((N, N), dtype=torch.float64, device=dev)
    Xa_d = torch.empty_like(J_d)
    Xb_d = torch.empty_like(J_d)
    acc = db.accumulator()
    stream = db.torch_stream()
    lib = db.lib
    from pychem_b200 import _lib
    P = engine._ptr

    def step_device():
        _lib.check(lib.pc_jk_direct_accumulate(db.h, variant, P(Dt_d), P(Da_d), P(Da_d), P(acc)))
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        _lib.check(lib.pc_jk_finalize(db.h, variant, P(acc), P(J_d), P(Xa_d), P(Xb_d)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # clocks are sampled (nvidia-smi, 100 ms period) from the warm-up through the timed region:
    # a Fock build is ~20 ms, so short runs would otherwise see no sample at all
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    launches0 = db.launch_count()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (db.launch_count() - launches0) // max(args.steps, 1)
    ms_step = ms_total / args.steps
    value = counts["all_eris"] / (ms_step * 1e-3)

    if ms_total < 400.0:
        # timed region shorter than a few nvidia-smi periods: keep the same load running for
        # ~0.5 s (same iteration count on every rank) and sample the clocks under it
        n_extra = int(500.0 / max(ms_step, 1e-3)) + 1
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        for _ in range(n_extra):
            step_device()
        barrier()
        if rank == 0:
            extra = sampler2.stop()
            if extra.get("samples", 0) > clocks.get("samples", 0):
                clocks = extra
                clocks["note"] = ("timed region of %.0f ms is shorter than a few nvidia-smi periods; sampled over "
                                  "%d more identical steps right after it" % (ms_total, n_extra))
    # ---- e2e through the plugin call with host buffers --------------------------------------
    hf_gpu._STATE[id(mol)] = {"mode": "direct", "db": db, "G_dev": None, "molecule": mol}
    Dt_p = torch.from_numpy(Dt_h).pin_memory()
    Da_p = torch.from_numpy(Da_h).pin_memory()
    state = _State(Dt_p.numpy(), Da_p.numpy(), Da_p.numpy())

    def step_e2e():
        hf_gpu.make_coulomb_exchange_matrices(mol, state)

    for _ in range(max(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    e2e_value = counts["all_eris"] / (e2e_ms * 1e-3)
    err = float(np.abs(state.Total.Coulomb - J_d.cpu().numpy()).max())

    # ---- roofline: per-class device times (one profiled build), FP64 peak measured live ------
    db.set_profiling(True)
    _lib.check(lib.pc_jk_direct_accumulate(db.h, variant, P(Dt_d), P(Da_d), P(Da_d), P(acc)))
    db.set_profiling(False)
    my_flops, all_flops, per_class, ref_flops = algorithmic_flops(db, variant)
    peak = engine.fp64_peak_tflops(local)
    top = max(per_class.items(), key=lambda kv: kv[1]["ms"])
    eri_ms = sum(v["ms"] for v in per_class.values())
    achieved = all_flops / (ms_step * 1e-3) / 1e12 / world      # per-GPU TFLOP/s over the whole step
    # dominant kernel = the class kernel with the largest share of the step (per-launch CUDA-event
    # times of one profiled build).  `traffic`: DRAM bytes of its (fused) launch from the committed
    # ncu --set full capture (profiles/r1b_final_ncu_full_psss.txt) -- everything is
    # L2-resident, the path is not HBM-bound.
    top_tf = top[1]["flops"] / (top[1]["ms"] * 1e-3) / 1e12 if top[1]["ms"] else None
    roofline = {"bound": "fp64", "achieved": top_tf, "peak": peak, "unit": "TFLOP/s",
                "frac": (top_tf / peak) if (top_tf and peak) else None,
                "traffic": 44.9e6 if top[0] == "psss" else None,
                "kernel": "eri_%s_kernel<JK_RHF>: one fused launch per build covering its %d bucket pairs, %.3f ms "
                          "alone, %.1f%% of the serialised per-class total"
                          % (top[0], sum(1 for c in db.plan_items()[0] if "spd"[c[0]] + "spd"[c[1]] + "spd"[c[2]] + "spd"[c[3]] == top[0]),
                             top[1]["ms"], 100.0 * top[1]["ms"] / eri_ms if eri_ms else 0.0),
                "peak_source": "pc_fp64_peak: register-resident DFMA loop measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the fused eri_psss launch, ncu --set "
                                  "full, profiles/r1b_final_ncu_full_psss.txt (35.6 MB read + 9.3 MB written; pair tables, Boys table, densities "
                                  "and accumulators are L2-resident: the kernel is LSU/atomic-bound, not HBM-bound)",
                "whole_step": {"achieved": achieved, "frac": achieved / peak if peak else None,
                               "kernels": "all %d launches of one Fock build (21 eri_*_kernel<JK_RHF> + finalize), concurrent "
                                          "on 8 streams, replayed as one CUDA graph" % launches,
                               "algorithmic_gflop_per_step": all_flops / 1e9,
                               "reference_unscreened_gflop_per_step": ref_flops / 1e9,
                               "serialised_kernel_ms_over_step_ms": eri_ms / ms_step if ms_step else None},
                "fp64_instruction_bound": {
                    "dominant_kernel_frac_ceiling": 0.407 if top[0] == "psss" else None,
                    "whole_step_frac_ceiling": 0.707,
                    "note": "static SASS count of the built kernels (tools/sass_mix.py, profiles/r1d_sass_instruction_mix.txt): "
                            "FP64-pipe instructions actually executed per primitive quartet vs the flop model above -- the "
                            "share of the FP64 peak a perfectly pipelined loop of the present code could show under that model"},
                "flop_count": "executed primitive quartets (after the 1e-24 primitive-pair cut-off) * flop_prim "
                              "+ quartets * (flop_cont + digestion), SURVEY 8(d) model on the generator's DAG "
                              "(pychem_b200/data/flop_model.json)"}
    # pure ERI generation (same schedule, integrals discarded): the "FP64 ERIs/sec" of generation
    def step_eri_only():
        _lib.check(lib.pc_jk_direct_accumulate(db.h, 5, P(Dt_d), P(Da_d), P(Da_d), P(acc)))
    step_eri_only()
    eri_only_ms = timed(step_eri_only, max(args.steps, 2)) / max(args.steps, 2)
    if args.profile_classes and rank == 0:
        db.set_profiling(True)
        step_eri_only()
        db.set_profiling(False)
        cls2, _, _, ms2 = db.plan_items()
        gen = {}
        for (l1, l2, l3, l4), t in zip(cls2, ms2):
            nm = "spd"[l1] + "spd"[l2] + "spd"[l3] + "spd"[l4]
            gen[nm] = gen.get(nm, 0.0) + float(t)
        for k, v in per_class.items():
            v["gen_ms"] = gen.get(k, 0.0)
    if args.profile_classes and rank == 0:
        for k, v in sorted(per_class.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0.0
            print("class %s: %9.3f ms (generation only %7.3f ms)  %12d quartets  %7.3f TFLOP/s (%.1f%% of peak)"
                  % (k, v["ms"], v.get("gen_ms", 0.0), v["quartets"], tf, 100 * tf / peak), file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = reference_sample(n, args.cpu_seconds)
        except Exception as exc:      # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n), "basis_functions": N, "shells": db.nshell,
                           "unique_quartets_surviving": counts["all_quartets"],
                           "eris_per_step": counts["all_eris"],
                           "density": "symmetric random, seed 1234 (RHF-shaped: Da == Db)",
                           "parallelism": "quartet-partition x%d + 1 NCCL all-reduce of 3*N^2 doubles" % world,
                           "l2_policy": "inputs larger than L2: shell-pair tables + 3 N^2 matrices are re-streamed "
                                        "by >1e8 quartets per step; no reuse between steps is possible "
                                        "(each step overwrites the accumulators)"},
                "fock_build_ms": ms_step,
                "setup_seconds": {"basis_tables": t_basis, "schwarz": t_schwarz, "plan": t_plan,
                                  "note": "once per geometry, outside the timed region"},
                "eri_generation_only": {"ms_per_pass": eri_only_ms, "value": counts["all_eris"] / (eri_only_ms * 1e-3), "unit": UNIT},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": 3 * N * N * 8, "d2h_bytes_per_step": 3 * N * N * 8,
                        "max_abs_diff_vs_device_path": err},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    db.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
