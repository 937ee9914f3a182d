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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--waters", type=int, default=int(os.environ.get("PYCHEM_BENCH_WATERS", "32")))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-classes", action="store_true", help="print per-class device times to stderr")
    ap.add_argument("--sweep", default=os.environ.get("PYCHEM_BENCH_SWEEP", "8,16,32"),
                    help="(H2O)n sizes of the BASELINE metric's sweep reported in the `sweep` array")
    ap.add_argument("--no-stored", action="store_true", help="skip the stored-tensor leg ((H2O)8, 10.9 GB)")
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
        e = per_class.setdefault(name, {"flops": 0.0, "ms": 0.0, "quartets": 0, "items": 0})
        e["flops"] += f_tot * share
        e["ms"] += float(t)
        e["quartets"] += int(cnt)
        e["items"] += 1
    return mine, total, per_class, total_ref


def committed_traffic(kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's launch from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json, written from the .ncu-rep by
    tools/ncu_raw_summary.py); None when the capture holds no launch of that class."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tab = json.load(fh)
        e = tab.get("eri_%s_kernel<JK_RHF>" % kernel_class)
        return (e["dram_bytes"], e["source"]) if e else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def sample_targets(n):
    """Outputs compared with the oracle (same choice as tests/helpers.water_cluster_samples)."""
    w = lambda k, s: 12 * (k % n) + s          # noqa: E731
    j_pairs = [(w(0, 0), w(0, 0)), (w(0, 5), w(0, 5)), (w(0, 2), w(1, 4)), (w(0, 5), w(1, 7)),
               (w(0, 3), w(5, 5)), (w(3, 8), w(3, 11)), (w(0, 4), w(n - 1, 4)), (w(7, 1), w(9, 5))]
    return [(min(a, b), max(a, b)) for a, b in j_pairs], [w(0, 5), w(0, 4), w(n // 2, 0), w(n - 1, 8)]


def oracle_check(table, n, Dt, Da, J, Xa):
    """Part of the cpu_baseline leg (rank 0, N = 1): eight J blocks and four X rows of the bench
    workload from the CPU oracle (oracle/eri_oracle.c, orc_jk_sample) against the device result."""
    from oracle import oracle
    t0 = time.perf_counter()
    ob = oracle.OracleBasis(table)
    _, pm = ob.schwarz()
    jp, ks = sample_targets(n)
    J0, Xa0, _, mJ, mX, nq = oracle.jk_sample(ob, Dt, Da, Da, jp, ks, thresh=THRESH, pmax=pm)
    return {"max_abs_diff_J": float(np.abs(J - J0)[mJ].max()), "max_abs_diff_Xa": float(np.abs(Xa - Xa0)[mX].max()),
            "entries_J": int(mJ.sum()), "entries_Xa": int(mX.sum()), "oracle_quartets": int(nq),
            "seconds": time.perf_counter() - t0,
            "what": "8 Coulomb blocks + 4 exchange rows of this workload from oracle/eri_oracle.c (orc_jk_sample: "
                    "reference screening hartree_fock.py:276-295, contraction :345-347) vs the device-timed result"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from pychem_b200 import _lib, dist as pdist, engine, hartree_fock as hf_gpu, integrals as ints_gpu, structures as S

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the ONE JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    rank, world, local = pdist.init("nccl" if args.gpus > 1 else None)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["PYCHEM_B200_MODE"] = "direct"
    P = engine._ptr
    variant = engine.RHF

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Case:
        """One (H2O)n molecule set up through the plugin (evaluate_2e_ints) + device-resident buffers."""

        def __init__(self, n):
            self.n = n
            self.mol = S.Molecule(S.water_cluster(n), "6-31G**")
            self.N = self.mol.NOrbitals
            t0 = time.perf_counter()
            self.db = ints_gpu.device_basis(self.mol)
            self.t_basis = time.perf_counter() - t0
            self.db.schwarz()
            self.t_schwarz = time.perf_counter() - t0 - self.t_basis
            hf_gpu.evaluate_2e_ints(self.mol)                   # Bounds + plan for this rank's slice
            self.t_plan = time.perf_counter() - t0 - self.t_basis - self.t_schwarz
            self.counts = self.db.counts
            assert self.counts["rank"] == rank and self.counts["nranks"] == world
            # synthetic symmetric density: D = (X + X^T)/2, X ~ U(-1,1), seed 1234
            X = np.random.default_rng(1234).uniform(-1, 1, (self.N, self.N))
            self.Da_h = 0.5 * (X + X.T)
            self.Dt_h = 2.0 * self.Da_h
            self.Dt_d = torch.from_numpy(self.Dt_h).to(dev)
            self.Da_d = torch.from_numpy(self.Da_h).to(dev)
            self.J_d = torch.empty((self.N, self.N), dtype=torch.float64, device=dev)
            self.Xa_d = torch.empty_like(self.J_d)
            self.Xb_d = torch.empty_like(self.J_d)
            self.acc = self.db.accumulator()
            self.stream = self.db.torch_stream()
            torch.cuda.synchronize()

        def step_device(self, v=None):
            db, lib = self.db, self.db.lib
            _lib.check(lib.pc_jk_direct_accumulate(db.h, variant if v is None else v, P(self.Dt_d), P(self.Da_d),
                                                   P(self.Da_d), P(self.acc)))
            if v is not None:
                return
            if world > 1:
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
            _lib.check(lib.pc_jk_finalize(db.h, variant, P(self.acc), P(self.J_d), P(self.Xa_d), P(self.Xb_d)))

        def timed(self, fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
            e1.synchronize()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

    main = Case(args.waters)
    n, N, db, counts = main.n, main.N, main.db, main.counts
    lib = db.lib

    # clocks are sampled (nvidia-smi, 100 ms period) from the warm-up through the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        main.step_device()
    launches0 = db.launch_count()
    ms_total = main.timed(main.step_device, args.steps)
    launches = (db.launch_count() - launches0) // max(args.steps, 1)
    ms_step = ms_total / args.steps
    value = counts["all_eris"] / (ms_step * 1e-3)
    # a timed region of >= 2 s beside the K steps the caller asked for (same steps, same timing)
    sustained = None
    if ms_total < 2000.0:
        n_more = int(2000.0 / max(ms_step, 1e-3)) + 1
        ms_more = main.timed(main.step_device, n_more)
        sustained = {"steps": n_more, "ms_per_step": ms_more / n_more, "seconds": ms_more * 1e-3,
                     "value": counts["all_eris"] / (ms_more / n_more * 1e-3)}
    clocks = sampler.stop() if rank == 0 else None
    J_dev_host = main.J_d.cpu().numpy()
    Xa_dev_host = main.Xa_d.cpu().numpy()

    # ---- e2e through the plugin call with host buffers --------------------------------------
    def e2e_run(pinned):
        if pinned:
            Dt_p, Da_p = torch.from_numpy(main.Dt_h).pin_memory(), torch.from_numpy(main.Da_h).pin_memory()
            Db_p = torch.from_numpy(main.Da_h).pin_memory()
            state = _State(Dt_p.numpy(), Da_p.numpy(), Db_p.numpy())        # three distinct pinned arrays
            state._keep = (Dt_p, Da_p, Db_p)
        else:
            # what the reference's SCF driver hands over: three distinct pageable numpy arrays
            state = _State(main.Dt_h.copy(), main.Da_h.copy(), main.Da_h.copy())
        for _ in range(max(args.warmup, 3)):
            hf_gpu.make_coulomb_exchange_matrices(main.mol, state)
        barrier()
        prof = None
        if os.environ.get("PYCHEM_B200_BENCH_PROFILE") and rank == 0:      # diagnostic: host profile of the e2e loop
            import cProfile
            prof = cProfile.Profile()
            prof.enable()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            hf_gpu.make_coulomb_exchange_matrices(main.mol, state)
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device=dev)
        if prof is not None:
            import pstats
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(30)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), state

    e2e_ms, state = e2e_run(pinned=True)
    # (results of the N > 1 path are views of the shared host buffer, valid until the next call: compare now)
    err = max(float(np.abs(state.Total.Coulomb - J_dev_host).max()), float(np.abs(state.Alpha.Exchange - Xa_dev_host).max()))
    e2e_page_ms, state_p = e2e_run(pinned=False)
    err = max(err, float(np.abs(state_p.Total.Coulomb - J_dev_host).max()))
    e2e_value = counts["all_eris"] / (e2e_ms * 1e-3)
    shared = bool(getattr(db, "_share", None) and db._share[1] is not None)

    # ---- roofline: per-class device times (one profiled build), FP64 peak measured live ------
    db.set_profiling(True)
    main.step_device(v=variant)
    db.set_profiling(False)
    my_flops, all_flops, per_class, ref_flops = algorithmic_flops(db, variant)
    peak = engine.fp64_peak_tflops(local)
    top = max(per_class.items(), key=lambda kv: kv[1]["ms"])
    eri_ms = sum(v["ms"] for v in per_class.values())
    achieved = all_flops / (ms_step * 1e-3) / 1e12 / world      # per-GPU TFLOP/s over the whole step
    top_tf = top[1]["flops"] / (top[1]["ms"] * 1e-3) / 1e12 if top[1]["ms"] else None
    traffic, traffic_src = committed_traffic(top[0])
    # `frac` = the WHOLE STEP (all class kernels + finalize of one Fock build): no single kernel
    # holds more than ~1/6 of the step, so the whole-step fraction is the honest headline; the
    # dominant kernel is reported beside it
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "kernels": "all %d launches of one Fock build (eri_*_kernel<JK_RHF> per class + finalize), concurrent on "
                           "8 streams, replayed as one CUDA graph" % launches,
                "algorithmic_gflop_per_step": all_flops / 1e9,
                "reference_unscreened_gflop_per_step": ref_flops / 1e9,
                "serialised_kernel_ms_over_step_ms": eri_ms / ms_step if ms_step else None,
                "dominant_kernel": {"name": "eri_%s_kernel<JK_RHF>" % top[0], "ms": top[1]["ms"],
                                    "share_of_serialised_step": top[1]["ms"] / eri_ms if eri_ms else None,
                                    "achieved": top_tf, "frac": (top_tf / peak) if (top_tf and peak) else None,
                                    "bucket_pairs_in_launch": top[1]["items"], "traffic": traffic,
                                    "traffic_source": traffic_src},
                "peak_source": "pc_fp64_peak: register-resident DFMA loop measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "flop_count": "executed primitive quartets (after the 1e-24 primitive-pair cut-off) * flop_prim "
                              "+ quartets * (flop_cont + digestion), SURVEY 8(d) model on the generator's DAG "
                              "(pychem_b200/data/flop_model.json)"}

    # pure ERI generation (same schedule, integrals discarded): the "FP64 ERIs/sec" of generation
    main.step_device(v=5)
    eri_only_ms = main.timed(lambda: main.step_device(v=5), max(args.steps, 2)) / max(args.steps, 2)
    if args.profile_classes and rank == 0:
        db.set_profiling(True)
        main.step_device(v=5)
        db.set_profiling(False)
        cls2, _, _, ms2 = db.plan_items()
        gen = {}
        for (l1, l2, l3, l4), t in zip(cls2, ms2):
            nm = "spd"[l1] + "spd"[l2] + "spd"[l3] + "spd"[l4]
            gen[nm] = gen.get(nm, 0.0) + float(t)
        for k, v in sorted(per_class.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0.0
            print("class %s: %9.3f ms (generation only %7.3f ms)  %12d quartets  %7.3f TFLOP/s (%.1f%% of peak)"
                  % (k, v["ms"], gen.get(k, 0.0), v["quartets"], tf, 100 * tf / peak), file=sys.stderr)

    # ---- checks: checksums of the result (comparable across N) ---------------------------------
    main.step_device()
    barrier()
    J_h, Xa_h = main.J_d.cpu().numpy(), main.Xa_d.cpu().numpy()
    checks = {"J_fro": float(np.linalg.norm(J_h)), "Xa_fro": float(np.linalg.norm(Xa_h)), "J_00": float(J_h[0, 0]),
              "Xa_00": float(Xa_h[0, 0]), "J_trace": float(np.trace(J_h)),
              "e2e_max_abs_diff_vs_device_path": err}

    # ---- the BASELINE metric's sweep: (H2O)n, n = 8..32 ------------------------------------------
    sweep = []
    for nw in [int(x) for x in args.sweep.split(",") if x]:
        if nw == n:
            sweep.append({"waters": n, "basis_functions": N, "fock_build_ms": ms_step, "value": value, "unit": UNIT,
                          "quartets": counts["all_quartets"], "eris": counts["all_eris"],
                          "roofline_frac": roofline["frac"]})
            continue
        c = Case(nw)
        for _ in range(3):
            c.step_device()
        k_steps = max(args.steps, 10)
        t = c.timed(c.step_device, k_steps) / k_steps
        _, fl, _, _ = algorithmic_flops(c.db, variant)
        sweep.append({"waters": nw, "basis_functions": c.N, "fock_build_ms": t,
                      "value": c.counts["all_eris"] / (t * 1e-3), "unit": UNIT,
                      "quartets": c.counts["all_quartets"], "eris": c.counts["all_eris"],
                      "roofline_frac": fl / (t * 1e-3) / 1e12 / world / peak if peak else None})
        hf_gpu.release(c.mol)
        ints_gpu.release(c.mol)
        del c
    sweep.sort(key=lambda e: e["waters"])

    # ---- stored-tensor mode at (H2O)8 (N = 192, 10.9 GB): HBM roofline of the streaming J/K ----
    stored = None
    if rank == 0 and world == 1 and not args.no_stored:
        try:
            stored = stored_mode_leg(torch, engine, S, dev, args)
        except Exception as exc:
            stored = {"error": repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = reference_sample(n, args.cpu_seconds)
        except Exception as exc:      # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}
        try:
            checks["oracle"] = oracle_check(db.table, n, main.Dt_h, main.Da_h, J_h, Xa_h)
        except Exception as exc:
            checks["oracle"] = {"error": repr(exc)}

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
                "sustained": sustained,
                "setup_seconds": {"basis_tables": main.t_basis, "schwarz": main.t_schwarz, "plan": main.t_plan,
                                  "note": "once per geometry, outside the timed region"},
                "eri_generation_only": {"ms_per_pass": eri_only_ms, "value": counts["all_eris"] / (eri_only_ms * 1e-3), "unit": UNIT},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                        # bytes moved over PCIe by ALL ranks per step.  N > 1 on one node: every rank uploads
                        # 1/N of the rows of each density (all-gather over NVLink) and downloads 1/N of the
                        # rows of J and X into the host buffer the ranks share (pychem_b200/dist.py NodeShare)
                        "h2d_bytes_per_step": 3 * N * N * 8 * (1 if shared or world == 1 else world),
                        "d2h_bytes_per_step": 2 * N * N * 8 * (1 if shared or world == 1 else world),
                        "h2d_bytes_per_step_per_rank": 3 * N * N * 8 // (world if shared else 1),
                        "d2h_bytes_per_step_per_rank": 2 * N * N * 8 // (world if shared else 1),
                        "host_traffic": ("row slices per rank, results in a host buffer shared by the ranks" if shared
                                         else "whole matrices per rank"),
                        "call": "pychem_b200.hartree_fock.make_coulomb_exchange_matrices(molecule, state), three pinned host "
                                "densities in, J / X host arrays out (closed-shell result: X_beta is the X_alpha array, 2 N^2 doubles back)",
                        "pageable_inputs": {"ms_per_step": e2e_page_ms, "value": counts["all_eris"] / (e2e_page_ms * 1e-3),
                                            "note": "three distinct pageable numpy arrays, as the reference's SCF driver passes"}},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "checks": checks,
                "sweep": sweep,
                "stored_mode": stored,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    hf_gpu.release()
    ints_gpu.release()
    if world > 1:
        dist.destroy_process_group()


def stored_mode_leg(torch, engine, S, dev, args):
    """(H2O)8 6-31G**: dense tensor build (8-fold scatter) and the streaming J/K pass over it."""
    mol = S.Molecule(S.water_cluster(8), "6-31G**")
    db = engine.DeviceBasis(mol, device=dev.index)
    N = db.nbf
    db.schwarz()
    stream = db.torch_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    G_dev, _ = db.eri_tensor(THRESH, to_host=False)            # warm-up (plan + first touch)
    e0.record(stream)
    from pychem_b200 import _lib
    _lib.check(db.lib.pc_eri_tensor(db.h, engine._ptr(G_dev), None))
    e1.record(stream)
    e1.synchronize()
    tensor_ms = e0.elapsed_time(e1)
    X = np.random.default_rng(1234).uniform(-1, 1, (N, N))
    Da = torch.from_numpy(0.5 * (X + X.T)).to(dev)
    Dt = 2.0 * Da
    torch.cuda.synchronize()
    reps = 20
    kernels = {}
    for name, flag in (("tma_pipeline", "1"), ("ldg128", "0")):       # the default (ldg128) last
        os.environ["PYCHEM_B200_STORED_TMA"] = flag
        for _ in range(3):
            db.jk_stored(G_dev, Dt, Da, Da)
        e0.record(stream)
        for _ in range(reps):
            db.jk_stored(G_dev, Dt, Da, Da)
        e1.record(stream)
        e1.synchronize()
        kernels[name] = e0.elapsed_time(e1) / reps
    os.environ.pop("PYCHEM_B200_STORED_TMA", None)
    jk_ms = kernels["ldg128"]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, ValueError, KeyError):
        peak, src = 6550.0, "fallback (tools guide): 6.55 TB/s"
    gbs = 8.0 * N ** 4 / (jk_ms * 1e-3) / 1e9
    out = {"workload": "(H2O)8 6-31G** stored-tensor mode, N=%d, tensor %.1f GB" % (N, 8.0 * N ** 4 / 1e9),
           "tensor_build_ms": tensor_ms, "jk_ms": jk_ms,
           "kernels_ms": {"jk_stored_kernel (16-byte loads, default)": kernels["ldg128"],
                          "jk_stored_tma_kernel (cp.async.bulk + mbarrier pipeline, PYCHEM_B200_STORED_TMA=1)": kernels["tma_pipeline"]},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "algorithmic_bytes": 8.0 * N ** 4, "peak_source": src,
                        "note": "one streaming pass over the tensor per Fock build (the reference's three einsum passes read 24 N^4 bytes); "
                                "tensor larger than L2, timed over %d consecutive passes" % reps}}
    del G_dev
    db.close()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
