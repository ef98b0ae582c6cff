#!/usr/bin/env python
"""bench.py -- tour fitness evaluations/s of the `allhic optimize` hot path on B200.

The workload is BASELINE.json configs[2] ("synthetic 2,000-contig group, population 8k, ordering +
EvaluateQ orientation phase, island GA at 1/2/4/8 B200").  One "step" = G generations of CLM.GARun
(evaluate.go:258) on a population of P tours per GPU -- at N > 1 as the island model whose exchange
step runs inside the GA kernels over NVLink peer memory (include/ahx.h "Islands"), so the collective
is inside the timed region.
  value  Tour.Evaluate calls per second made by that GA, all GPUs together, population resident in
         HBM, timed with CUDA events around the GA kernel launches (max over ranks)
  e2e    the same through the call a host makes (CLM.GARun -> ahx_ga_begin / advance / best, at N > 1
         with the island wiring): initial tour from a pinned host buffer in, best tour + statistics
         out, wall clock, every step a complete run
  evaluate_only  (extra) batched Tour.Evaluate of P tours per GPU: device-resident kernel rate and
         roofline, end to end through ahx_eval_m_u16 with pinned and with pageable host buffers
`--impl reference` times the CPU restatement of the reference's own implementation (oracle/, C, all
host threads) on the same workload: the same GA (eaopt emulation), a bounded number of generations per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config3] [--impl ours|reference]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LIMIT = 10_000_000
HBM_PEAK = 6544.3   # GB/s, fallback when MEASURED_PEAKS.json is absent (it is driver-written per pod)
HBM_PEAK_SRC = "fallback 6544.3 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as _f:
        HBM_PEAK = float(json.load(_f)["hbm_gbs"])
        HBM_PEAK_SRC = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    pass
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
NCU_DRAM_BYTES_PER_LAUNCH = {("eval_m_res_kernel", "config3", 8000): 64222976 + 442880}      # profiles/r2_ncu_full_summary.txt
# ga_persist_kernel: one 30-generation launch at config3 read 15 527 936 B and wrote 398 512 384 B (profiles/r2_ncu_ga_persist_summary.txt)
NCU_DRAM_BYTES_PER_GENERATION = {("ga_persist_kernel", "config3", 8000): (15527936 + 398512384) / 30.0}
METRIC = "tour fitness evals/sec"
UNIT = "evals/s"
GA_GENS_PER_STEP = 500          # evaluate.go:294: the reference logs every 500 generations; one step = one such period
MUT_PROB = 0.2                  # base.go:71

_JSON_FD = None


def emit(obj):
    """The bench's one JSON line, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


DESCS = {"config2": "synthetic 500-contig linkage group, population 1k, ordering phase (BASELINE configs[1])",
         "config3": "synthetic 2,000-contig linkage group, population 8k per GPU, GA ordering phase as an island model (BASELINE configs[2], the north_star target size)",
         "config5": "synthetic 10,000-contig linkage group, population 8k per GPU, GA ordering phase as an island model (BASELINE configs[4])"}


def workload(name):
    from allhic_b200 import synth
    cfg = synth.CONFIGS[name]
    ids, clm, _ = synth.materialize(name)
    return ids, clm, cfg["n"], cfg["npop"], DESCS[name]


def gens_per_step(n):
    return GA_GENS_PER_STEP if n <= 2000 else 100


def config_dict(desc, n, P, E, world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": desc, "n_contigs": n, "population_per_gpu": P, "contig_pairs_E": E, "mut_prob": MUT_PROB, "tournament": 3,
            "step": "%d generations of CLM.GARun (a bounded sample of them per step on the CPU arm)" % gens_per_step(n),
            "l2": "GA state per GPU (2 x population x (tour + coordinate row) = %.0f MB) is larger than L2; evaluate_only flushes L2 between steps" % (16.0 * P * n / 1e6),
            "parallelism": ("island model: one population per GPU, %d GPUs, 2 elites to every other island every 50 generations, exchange inside the GA kernels over NVLink" % world)
            if world > 1 else "single GPU, single population"}


def make_tours(n, P, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)


def work_per_tour(sizes, tables, tours, sample=128):
    """Mean pair-terms per tour: E_act (contig-space kernel: stored pairs within LIMIT) and W
    (position-space window pairs of the reference loop)."""
    idx = np.linspace(0, len(tours) - 1, min(sample, len(tours))).astype(int)
    e_act, w = [], []
    for i in idx:
        t = tours[i]
        s = sizes[t].astype(np.float64)
        mid = np.cumsum(s) - s / 2
        x = np.full(len(sizes), np.nan); x[t] = mid
        d = np.abs(x[tables["pa"]] - x[tables["pb"]])
        e_act.append(int((d <= LIMIT).sum()))
        hi = np.searchsorted(mid, mid + LIMIT, side="right")
        w.append(int((hi - np.arange(len(t)) - 1).sum()))
    return float(np.mean(e_act)), float(np.mean(w))


def evalq_terms(sizes, tables, tour, signs):
    """E_w of one EvaluateQ call: stored pairs that have a GArray for the current signs
    (orientation.go:143) and lie within LIMIT (dist = cumsize[j-1] - cumsize[i], orientation.go:181)."""
    n = len(sizes)
    pos = np.full(n, -1, np.int64); pos[tour] = np.arange(len(tour))
    pa, pb = tables["pa"], tables["pb"]
    act = (pos[pa] >= 0) & (pos[pb] >= 0)
    a_first = pos[pa] < pos[pb]
    i = np.minimum(pos[pa], pos[pb]); j = np.maximum(pos[pa], pos[pb])
    sa = (signs[pa] == ord('-')).astype(np.int64); sb = (signs[pb] == ord('-')).astype(np.int64)
    slot = np.where(a_first, 2 * sa + sb, 4 + 2 * sb + sa)
    has = tables["slots"][np.arange(len(pa)), slot] >= 0
    cum = np.concatenate([[0], np.cumsum(sizes[tour])])[:-1]
    dist = cum[np.maximum(j - 1, 0)] - cum[np.maximum(i, 0)]
    return int((act & has & (dist <= LIMIT)).sum())


class ClockSampler:
    """Samples SM clock + throttle reasons DURING the timed region (NVML from a
    thread; ctypes calls release the GIL so sampling continues while kernels run)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu):
        import threading
        self.gpu, self.samples, self.mask, self.max_mhz, self.err = gpu, [], 0, None, None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical GPUs; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self._stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    try:
                        self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                self.samples.append((float(sm), pw))
                self._stop.wait(0.01)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        self._t.start()

    def peak_mhz(self):
        """SM clock under load so far (median of the upper half by power), else the NVML max, else 1965."""
        if self.samples:
            by_power = sorted(self.samples, key=lambda x: x[1])
            return float(np.median([s for s, _ in by_power[len(by_power) // 2:]]))
        return float(self.max_mhz or 1965.0)

    def stop(self):
        self._stop.set(); self._t.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples: %s" % self.err]}
        # the GPU idles between phases (host-side setup): keep the samples taken under load (upper half by power)
        by_power = sorted(self.samples, key=lambda x: x[1])
        busy = [s for s, _ in by_power[len(by_power) // 2:]]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for bit, n in self.REASONS.items() if self.mask & bit), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_ga_sample(ids, clm, init, P, gens, threads, repeats=1):
    """`repeats` runs of the CPU restatement's GARun (oracle, eaopt emulation with ParallelEval over `threads`
    host threads) for `gens` generations each from `init`; returns (evaluations, generations, seconds, best)."""
    from oracle_lib import OracleCLM
    orc = OracleCLM(clm, ids)
    ev = gn = 0
    t0 = time.perf_counter()
    best = None
    for _ in range(repeats):
        orc.set_tour(np.ascontiguousarray(init, np.int64))
        r = orc.ga_run(npop=P, ngen=1 << 30, mutprob=MUT_PROB, seed=42, max_gens=gens, nthreads=threads)
        ev += r.evaluations; gn += r.generations; best = r.best_score
    return ev, gn, time.perf_counter() - t0, best


def cpu_population_evals(ids, clm, tours, min_seconds, threads):
    """The CPU restatement's ParallelEval over the same population (oracle, kind=port)."""
    from oracle_lib import OracleCLM
    orc = OracleCLM(clm, ids)
    t64 = np.ascontiguousarray(tours, np.int64)
    orc.population_evaluate(t64[: max(threads, 8)], nthreads=threads)       # warm-up: builds M
    t0 = time.perf_counter(); reps = 0
    while True:
        orc.population_evaluate(t64, nthreads=threads); reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return reps * len(t64) / dt, reps, dt


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path.  The Go binary cannot be built in
    this image (no Go toolchain), so this is the C restatement in oracle/ (kind=port) on all host threads;
    rank 0 only.  One step = a bounded sample of the workload's step: GARun from the same initial tour for
    `gens` generations (population initialisation included, as in the reference: eaopt evaluates all P clones)."""
    if rank != 0:
        return
    from oracle_lib import OracleCLM
    ids, clm, n, P, desc = workload(args.workload)
    _o = OracleCLM(clm, ids)
    E = int(_o.l.orc_ncontacts(_o.h))          # contig pairs with a CLM record (`extract` writes every pair once, a < b)
    del _o
    threads = os.cpu_count() or 1
    gens = 2 if n >= 2000 else 20
    init = make_tours(n, 1, 1000)[0]
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_ga_sample(ids, clm, init, P, 1, threads)
    ev, gn, dt, best = cpu_ga_sample(ids, clm, init, P, gens, threads, repeats=args.steps)
    v = ev / dt
    sample = ("%d steps x GARun for %d generations of the %d-generation step (npop %d, %d Tour.Evaluate calls per step incl. the initial population), "
              "oracle/allhic_oracle.c orc_ga_run, %d threads" % (args.steps, gens, gens_per_step(n), P, ev // max(1, args.steps), threads))
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(desc, n, P, E, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "ga_generations_per_s": gn / dt},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ------------------------------------------------------------------------------------------- GPU arm
def evaluate_only(A, eng, clocks, args, world, barrier, max_over_ranks, tours, sizes, tables, n, P, E, wname):
    """Batched Tour.Evaluate of P tours per GPU: device-resident kernel rate + roofline, and end to end through the
    C ABI with pinned and with pageable host buffers."""
    L = n
    e_act, w_dense = work_per_tour(sizes, tables, tours)
    use16 = n <= 65535
    pin_t = A.PinnedArray((P, L), np.uint16 if use16 else np.int32)
    pin_o = A.PinnedArray((P,), np.float64)
    pin_t.array[:] = tours
    page_t = np.ascontiguousarray(tours.astype(np.uint16 if use16 else np.int32))    # ordinary (pageable) numpy memory: what a Go slice is
    page_o = np.empty(P, np.float64)
    h2d, d2h = pin_t.nbytes, pin_o.nbytes
    eng.pop_set(tours)
    W = max(3, args.warmup)
    for _ in range(W):
        eng.flush_l2(); eng.pop_eval_m("sparse", 1)
    for _ in range(W):
        eng.eval_m(pin_t.array, out=pin_o.array); eng.eval_m(page_t, out=page_o)
    barrier()
    lc0 = eng.launch_count()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.flush_l2()                                    # not timed: evict the tours / pair table from L2
        dev_ms += eng.pop_eval_m("sparse", 1)             # CUDA events around the kernel on its own stream
    launches = (eng.launch_count() - lc0) - args.steps    # minus the flush kernels
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    scores_dev = eng.pop_scores()
    e2e = {}
    for tag, tin, tout in (("pinned", pin_t.array, pin_o.array), ("pageable", page_t, page_o)):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.eval_m(tin, out=tout)
        e2e[tag] = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert np.allclose(scores_dev, tout, rtol=1e-13, atol=0), "resident and e2e paths disagree"
    kern_s = dev_ms * 1e-3 / args.steps
    sm_mhz = clocks.peak_mhz()
    smem_peak = 128.0 * eng.sm_count() * sm_mhz * 1e6 / 1e9                      # GB/s: 128 B/clk/SM
    kind, tpb = eng.eval_m_kernel(P, L)
    flop_per_pair = 10.0 if kind >= 2 else 2.0 * e_act / max(E, 1)               # the streaming fallbacks divide in-window pairs only
    kname = {3: "eval_m_res_kernel (table streamed)", 2: "eval_m_res_kernel", 1: "eval_m_x32_kernel", 0: "eval_m_sparse_kernel"}[kind]
    legs = {"smem": (P * 8.0 * E / kern_s / 1e9, smem_peak, "GB/s"),
            "fp64": (P * E * flop_per_pair / kern_s / 1e12, args.dfma_tflops, "TFLOP/s"),
            "hbm": (P * (4.0 * L + 8.0) / kern_s / 1e9, HBM_PEAK, "GB/s")}
    bound = max(legs, key=lambda k: legs[k][0] / legs[k][1])
    roof = {"kernel": kname, "bound": bound, "achieved": legs[bound][0], "peak": legs[bound][1], "unit": legs[bound][2],
            "frac": legs[bound][0] / legs[bound][1],
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(("eval_m_res_kernel", wname, P)),
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of this kernel at this size (profiles/), bytes per launch; null = no capture at this size",
            "tours_per_batch": tpb, "pair_table_entries": eng.pair_table_entries(tpb) if kind in (2, 3) else None,
            "algorithmic_per_tour": {"stored_pairs_E": E, "fp64_flop": flop_per_pair * E, "smem_gather_bytes": 8 * E, "in_window_pairs_E_act": e_act, "hbm_bytes": 4 * L + 8},
            "legs": {k: {"achieved": v[0], "peak": v[1], "unit": v[2], "frac": v[0] / v[1]} for k, v in legs.items()},
            "dense_equivalent": {"window_pairs_per_tour": w_dense, "bytes_per_tour": 4 * w_dense + 12 * L,
                                 "algorithmic_GBs": P * (4 * w_dense + 12 * L) / kern_s / 1e9, "hbm_peak_GBs": HBM_PEAK,
                                 "note": "north_star's dense M-gather bytes at this eval rate: the algorithmic speed-up of the contig-space formulation, not a roofline fraction"}}
    out = {"value": world * P * args.steps / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / args.steps, "gpu_launches": int(launches),
           "note": "one launch scores the P tours of a GPU; L2 flushed (write of 2 x L2) between launches; at N > 1 the GPUs do not communicate",
           "e2e_pinned": {"value": world * P * args.steps / e2e["pinned"], "unit": UNIT, "ms_per_step": e2e["pinned"] / args.steps * 1e3,
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "call": "ahx_eval_m_u16" if use16 else "ahx_eval_m", "host_buffers": "ahx_host_alloc (pinned)"},
           "e2e_pageable": {"value": world * P * args.steps / e2e["pageable"], "unit": UNIT, "ms_per_step": e2e["pageable"] / args.steps * 1e3,
                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_buffers": "ordinary heap memory (what a GC-managed Go slice passed through cgo is)"},
           "roofline": roof}
    pin_t.close(); pin_o.close()
    return out


def ga_run_once(A, eng, init_pin, prm_kw, world, gens):
    """One complete CLM.GARun through the public API: initial tour from a pinned host buffer in, best tour and
    statistics out.  world > 1: IslandGA (begin with island parameters, all-gather of the mailbox handles, advance)."""
    prm = eng.ga_params(**prm_kw)
    if world > 1:
        from allhic_b200.islands import IslandGA
        ga = IslandGA(eng, init_pin, prm, migrate_every=50, n_elite=2)
        tour, best, st = ga.run(gens)
        mode = ga.mode
        ga.close()
        return st, best, mode
    best_t, st = eng.ga_run(init_pin, prm)
    return st, st.best_score, "single"


def orientation_legs(A, eng, tours, sizes, tables, n, args):
    """EvaluateQ and flipOne: device-timed kernels against SURVEY 8(d)'s per-call figures."""
    rng = np.random.default_rng(7)
    nq = 1024
    L = n
    signs = np.where(rng.random((nq, n)) < 0.5, ord('+'), ord('-')).astype(np.uint8)
    ew = float(np.mean([evalq_terms(sizes, tables, tours[0], signs[k]) for k in range(8)]))
    flop_call = 24.0 * ew                   # 12 x (int add, cvt, fma) per pair; + 12 logs per pair, listed next to it
    bytes_call = 48.0 * ew + 12.0 * L + n
    out = {"tour_length": int(L), "E_w_mean": ew, "algorithmic_per_call": {"fp64_flop": flop_call, "logs": 12 * ew, "bytes": bytes_call}}
    # (1) one tour, many sign vectors: pair_scores_kernel (12 logs per stored pair, once) + eval_q_lookup_kernel
    eng.eval_q(tours[0], signs[:8])
    w_t, d_t = [], []
    for _ in range(9):
        t0 = time.perf_counter(); eng.eval_q(tours[0], signs); w_t.append(time.perf_counter() - t0); d_t.append(eng.last_device_ms() * 1e-3)
    dk = float(np.median(d_t))
    out["shared_tour"] = {"kernel": "pair_scores_kernel + eval_q_lookup_kernel", "batch": nq, "eval_q_per_s_device": nq / dk, "eval_q_per_s_e2e": nq / float(np.median(w_t)),
                          "roofline": {"bound": "hbm", "achieved": nq * bytes_call / dk / 1e9, "peak": HBM_PEAK, "unit": "GB/s", "frac": nq * bytes_call / dk / 1e9 / HBM_PEAK,
                                       "fp64": {"achieved": nq * flop_call / dk / 1e12, "peak": args.dfma_tflops, "unit": "TFLOP/s", "frac": nq * flop_call / dk / 1e12 / args.dfma_tflops},
                                       "note": "SURVEY 8(d) per-call bytes / flops x calls; the 12 logs per pair are computed once per batch (S4 table), so the achieved figures count work the kernel shares between calls"}}
    # (2) per-individual tours: eval_q_kernel, one CTA per (tour, signs) individual
    ni = 512
    eng.eval_q(tours[:8], signs[:8])
    w_t, d_t = [], []
    for _ in range(9):
        t0 = time.perf_counter(); eng.eval_q(tours[:ni], signs[:ni]); w_t.append(time.perf_counter() - t0); d_t.append(eng.last_device_ms() * 1e-3)
    dk = float(np.median(d_t))
    out["per_individual"] = {"kernel": "eval_q_kernel", "batch": ni, "eval_q_per_s_device": ni / dk, "eval_q_per_s_e2e": ni / float(np.median(w_t)),
                             "roofline": {"bound": "fp64", "achieved": ni * flop_call / dk / 1e12, "peak": args.dfma_tflops, "unit": "TFLOP/s", "frac": ni * flop_call / dk / 1e12 / args.dfma_tflops,
                                          "logs_per_s": ni * 12 * ew / dk,
                                          "hbm": {"achieved": ni * bytes_call / dk / 1e9, "peak": HBM_PEAK, "unit": "GB/s", "frac": ni * bytes_call / dk / 1e9 / HBM_PEAK}}}
    # (3) one greedy flipOne pass (1 + L EvaluateQ calls in the reference, orientation.go:87-120)
    w_t, d_t = [], []
    for _ in range(7):
        t0 = time.perf_counter(); eng.flip_one(tours[0], signs[0]); w_t.append(time.perf_counter() - t0); d_t.append(eng.last_device_ms() * 1e-3)
    dk = float(np.median(d_t))
    out["flip_one"] = {"kernel": "pair_scores_kernel + flip_records_kernel + flip_one_stream_kernel", "pass_ms_device": dk * 1e3, "pass_ms_e2e": float(np.median(w_t)) * 1e3,
                       "reference_equivalent_eval_q_per_s": (1 + L) / dk,
                       "note": "the reference's pass is 1 + L full EvaluateQ calls; here each step is an O(degree) update of the sum"}
    # flipAll: sparse Lanczos on the host
    t0 = time.perf_counter(); eng.flip_all(tours[0], np.full(n, ord('+'), np.uint8)); out["flip_all_ms_e2e"] = (time.perf_counter() - t0) * 1e3
    return out


def config1_cli(args):
    """BASELINE configs[0]: the reference's fixture through the CLI with and without --skipGA (functional-tests.sh:10-14),
    beside the CPU restatement of the same run (GA phases only: npop 100, ngen 5000)."""
    from allhic_b200 import farm
    gold = os.path.join(ROOT, "tests", "golden")
    out = {}
    d = tempfile.mkdtemp(prefix="ahx_config1_")
    try:
        for tag, flags in (("optimize", []), ("optimize_skipGA", ["--skipGA"])):
            shutil.copy(os.path.join(gold, "test.ids"), d); shutil.copy(os.path.join(gold, "test.clm"), d)
            t0 = time.perf_counter()
            r = subprocess.run([farm.BINARY, "optimize", os.path.join(d, "test.ids"), os.path.join(d, "test.clm"), *flags], capture_output=True, text=True, timeout=600)
            dt = time.perf_counter() - t0
            scores = re.findall(r"max_score=([0-9.eE+-]+)", r.stdout)
            gens = re.findall(r"GA(\d)-(\d+):", r.stdout)
            out[tag] = {"wall_s": dt, "success": r.returncode == 0 and "Success" in r.stderr, "final_max_score": float(scores[-1]) if scores else None,
                        "last_logged_generation": {"GA" + p: int(g) for p, g in gens}}
        if args.cpu_seconds > 0:
            from oracle_lib import OracleCLM
            orc = OracleCLM(os.path.join(gold, "test.clm"), os.path.join(gold, "test.ids"))
            orc.activate(shuffle=True, seed=42, flip_all=False)
            t0 = time.perf_counter(); res = None; gens = 0
            for phase in (1, 2):
                res = orc.ga_run(npop=100, ngen=5000, mutprob=0.2, seed=42 + phase, max_gens=1000000, nthreads=os.cpu_count() or 1)
                gens += res.generations
            out["cpu_restatement"] = {"ga_phases_wall_s": time.perf_counter() - t0, "final_max_score": res.best_score, "generations": gens, "cores": os.cpu_count() or 1,
                                      "note": "the two GA phases only (the flips of the reference are 1 + L dense-Q EvaluateQ calls each pass)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def small_leg(A, name, args, dfma):
    """A short evaluate + GA measurement of another BASELINE config on this GPU (value, roofline fraction, CPU sample)."""
    ids, clmf, n, P, desc = workload(name)
    clm = A.CLM(clmf, ids)
    tables, sizes = clm.tables(), clm.sizes()
    E = len(tables["pa"])
    eng = A.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    eng.load(clm)
    tours = make_tours(n, P, 1000)
    eng.pop_set(tours)
    for _ in range(3):
        eng.flush_l2(); eng.pop_eval_m("sparse", 1)
    ms = 0.0
    reps = 10
    for _ in range(reps):
        eng.flush_l2(); ms += eng.pop_eval_m("sparse", 1)
    ks = ms * 1e-3 / reps
    kind, tpb = eng.eval_m_kernel(P, n)
    eng.ga_begin(tours[0], eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=42))
    gk = eng.ga_kernel()
    eng.ga_advance(50)
    gens = 1000 if n <= 2000 else 200
    st0 = eng.ga_advance(0)
    st = eng.ga_advance(gens)
    ga_s = (st.device_ms - st0.device_ms) * 1e-3
    ga_ev = st.evaluations - st0.evaluations
    eng.ga_end()
    # orientation phase at this size: flipAll (sparse Lanczos on the host + two EvaluateQ sums) and one flipOne pass
    rng = np.random.default_rng(7)
    sg = np.where(rng.random(n) < 0.5, ord('+'), ord('-')).astype(np.uint8)
    eng.flip_one(tours[0], sg)
    t0 = time.perf_counter(); eng.flip_one(tours[0], sg); fo_wall = time.perf_counter() - t0; fo_dev = eng.last_device_ms()
    t0 = time.perf_counter(); eng.flip_all(tours[0], np.full(n, ord('+'), np.uint8)); fa_wall = time.perf_counter() - t0
    info = eng.flip_all_info()
    out = {"workload": desc, "n_contigs": n, "population": P, "contig_pairs_E": E,
           "orientation": {"flip_one_pass_ms_device": fo_dev, "flip_one_pass_ms_e2e": fo_wall * 1e3, "flip_all_ms_e2e": fa_wall * 1e3,
                           "flip_all_lanczos": {"converged": bool(info["converged"]), "dense": bool(info["dense"]), "residual": info["residual"], "matvecs": info["matvecs"]}},
           "evaluate_only": {"evals_per_s": P / ks, "ms_per_launch": ks * 1e3, "kernel": {3: "eval_m_res_kernel (table streamed)", 2: "eval_m_res_kernel", 1: "eval_m_x32_kernel", 0: "eval_m_sparse_kernel"}[kind],
                             "tours_per_batch": tpb,
                             "roofline": {"bound": "fp64", "achieved": P * E * 10.0 / ks / 1e12, "peak": dfma, "unit": "TFLOP/s", "frac": P * E * 10.0 / ks / 1e12 / dfma}},
           "ga": {"evals_per_s": ga_ev / ga_s, "generations_per_s": gens / ga_s, "kernel": {3: "ga_persist_kernel (table streamed)", 2: "ga_persist_kernel", 1: "ga_plan_kernel + ga_apply_kernel", 0: "ga_generation_kernel"}[gk],
                  "roofline": {"bound": "fp64", "achieved": ga_ev * E * 10.0 / ga_s / 1e12, "peak": dfma, "unit": "TFLOP/s", "frac": ga_ev * E * 10.0 / ga_s / 1e12 / dfma}}}
    if args.cpu_seconds > 0:
        threads = os.cpu_count() or 1
        v, reps, dt = cpu_population_evals(ids, clmf, tours[: min(P, 2000)], min(3.0, args.cpu_seconds), threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "%d tours x %d passes (%.1f s) of Tour.Evaluate, C restatement + eaopt ParallelEval chunking" % (min(P, 2000), reps, dt)}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config5"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true", help="only the scored GA step, the evaluate-only block and the CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: native libraries (NCCL prints its version banner
    # to fd 1) and stray prints are sent to stderr; the line itself is written to the saved descriptor.
    sys.stdout.flush()
    global _JSON_FD
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import allhic_b200 as A
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if rank == 0:
        workload(args.workload)
    barrier()
    ids, clmf, n, P, desc = workload(args.workload)       # other ranks reuse the files rank 0 wrote
    clm = A.CLM(clmf, ids)
    tables, sizes = clm.tables(), clm.sizes()
    E = len(tables["pa"])
    eng = A.Engine(local)
    eng.load(clm)
    tours = make_tours(n, P, 1000 + rank)
    L = n
    G = gens_per_step(n)
    W = max(3, args.warmup)
    clocks = ClockSampler(local); clocks.start()
    dfma_tflops, ddiv_g = eng.probe_fp64()
    args.dfma_tflops = dfma_tflops
    l2_gbs = eng.probe_l2(64 << 20)

    # ------------------------------------------------------------------ the scored step: G generations of the (island) GA
    init_pin = A.PinnedArray((L,), np.int32)
    init_pin.array[:] = make_tours(n, 1, 1000)[0]          # every island starts from the same tour (MakeTour clones r.Tour, evaluate.go:259)
    prm_kw = dict(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=42, mut_prob=MUT_PROB)
    if world > 1:
        from allhic_b200.islands import IslandGA
        ga = IslandGA(eng, init_pin.array, eng.ga_params(**prm_kw), migrate_every=50, n_elite=2)
        island_mode = ga.mode
    else:
        eng.ga_begin(init_pin.array, eng.ga_params(**prm_kw))
        island_mode = "single"
    ga_kernel = eng.ga_kernel()
    # island_mode "host" (groups the persistent kernel cannot hold): the exchange is IslandGA's host-mediated one
    step = (lambda: ga.run(G)[2]) if island_mode == "host" else (lambda: eng.ga_advance(G))
    for _ in range(W):
        step()
    barrier()
    lc0 = eng.launch_count()
    st0 = eng.ga_advance(0)
    t0 = time.perf_counter()
    st = st0
    for _ in range(args.steps):
        st = step()                                         # CUDA events around the GA kernel launches on the context's stream
    wall_ga = max_over_ranks(time.perf_counter() - t0)
    barrier()
    ga_launches = eng.launch_count() - lc0
    dev_s = max_over_ranks((st.device_ms - st0.device_ms) * 1e-3)
    my_evals = float(st.evaluations - st0.evaluations)
    evals = sum_over_ranks(my_evals)
    gens_done = int(st.generations - st0.generations)
    epochs = eng.ga_island_epochs() if world > 1 else 0
    best_after = st.best_score
    if world > 1:
        ga.close()
    else:
        eng.ga_end()
    value = evals / dev_s
    # single-population control on the same GPUs (no exchange): what the island step costs relative to it
    control = None
    if world > 1:
        eng.ga_begin(init_pin.array, eng.ga_params(island=rank, **prm_kw))
        for _ in range(W):
            eng.ga_advance(G)
        barrier()
        c0 = eng.ga_advance(0)
        c1 = c0
        for _ in range(args.steps):
            c1 = eng.ga_advance(G)
        barrier()
        c_s = max_over_ranks((c1.device_ms - c0.device_ms) * 1e-3)
        c_ev = sum_over_ranks(float(c1.evaluations - c0.evaluations))
        eng.ga_end()
        control = {"evals_per_s_no_exchange": c_ev / c_s, "generations_per_s_no_exchange": args.steps * G / c_s,
                   "island_step_relative_to_no_exchange": (evals / dev_s) / (c_ev / c_s),
                   "exchange_ms_each": (dev_s - c_s) / max(1, args.steps * G // 50) * 1e3,
                   "note": "the same populations and Philox streams advanced without islands on the same GPUs in the same run; the difference is the exchange step (incl. waiting for the slowest island)"}

    # ------------------------------------------------------------------ e2e: complete GARun calls through the public API
    for _ in range(2):
        ga_run_once(A, eng, init_pin.array, dict(prm_kw, max_gens=G), world, G)
    barrier()
    e_ev = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st_e, _, _ = ga_run_once(A, eng, init_pin.array, dict(prm_kw, max_gens=G), world, G)
        e_ev += float(st_e.evaluations)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e_ev = sum_over_ranks(e_ev)
    barrier()
    h2d = 4 * L + 64 + (64 * world if world > 1 else 0)    # initial tour + ahx_ga_params (+ the islands' mailbox handles)
    d2h = 4 * L + 56 + 32                                  # best tour + ahx_ga_stats + the state read-back

    # roofline of the dominant kernel of the step (ga_persist_kernel: selection, clone-by-reference, mutation and the
    # evaluate loop of eval_m_res_kernel in one persistent cooperative kernel).  Per Tour.Evaluate call the algorithmic
    # work is E stored pairs x 5 fp64-pipe instructions (10 flop); HBM: the parent's and the child's tour + coordinate rows.
    sm_mhz = clocks.peak_mhz()
    ev_per_gpu_s = my_evals / ((st.device_ms - st0.device_ms) * 1e-3)
    legs = {"fp64": (ev_per_gpu_s * E * 10.0 / 1e12, dfma_tflops, "TFLOP/s"),
            "smem": (ev_per_gpu_s * 8.0 * E / 1e9, 128.0 * eng.sm_count() * sm_mhz * 1e6 / 1e9, "GB/s"),
            "hbm": (ev_per_gpu_s * 16.0 * L / 1e9, HBM_PEAK, "GB/s")}
    bound = max(legs, key=lambda k: legs[k][0] / legs[k][1])
    kname = {3: "ga_persist_kernel (pair table streamed)", 2: "ga_persist_kernel", 1: "ga_plan_kernel + ga_apply_kernel", 0: "ga_generation_kernel"}[ga_kernel]
    roof = {"kernel": kname, "bound": bound, "achieved": legs[bound][0], "peak": legs[bound][1], "unit": legs[bound][2], "frac": legs[bound][0] / legs[bound][1],
            "traffic": (NCU_DRAM_BYTES_PER_GENERATION[("ga_persist_kernel", args.workload, P)] * gens_done / max(1, ga_launches))
            if ga_kernel == 2 and ("ga_persist_kernel", args.workload, P) in NCU_DRAM_BYTES_PER_GENERATION else None,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of a 30-generation launch of this kernel at this size (profiles/r2_ncu_ga_persist_summary.txt), scaled to the generations of one launch here; null = no capture at this size",
            "peak_source": {"fp64": "DFMA microbenchmark measured live by this run (ahx_probe_fp64; MEASURED_PEAKS.json has no fp64 figure)",
                            "smem": "128 B/clk/SM x %d SMs x %.0f MHz" % (eng.sm_count(), sm_mhz), "hbm": HBM_PEAK_SRC}[bound],
            "algorithmic_per_eval": {"stored_pairs_E": E, "fp64_flop": 10.0 * E, "smem_gather_bytes": 8 * E, "hbm_bytes": 16 * L},
            "evals_per_launch": my_evals / max(1, ga_launches), "launch_ms": (st.device_ms - st0.device_ms) / max(1, ga_launches),
            "legs": {k: {"achieved": v[0], "peak": v[1], "unit": v[2], "frac": v[0] / v[1]} for k, v in legs.items()},
            "note": "per-GPU figures (rank 0); the evaluate loop inside this kernel is eval_m_res_kernel's (evaluate_only.roofline is that kernel alone, back to back, no selection / barriers)"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
           "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(desc, n, P, E, world),
           "e2e": {"value": e_ev / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                   "call": ("IslandGA over ahx_ga_begin + ahx_ga_island_export/connect + ahx_ga_advance + ahx_ga_best" if world > 1 else "ahx_ga_run") + ", %d generations per call" % G},
           "gpu_launches": int(ga_launches), "roofline": roof,
           "ga": {"generations_per_s": args.steps * G / dev_s, "generations_per_step": G, "evals_per_generation_per_gpu": my_evals / max(1, gens_done),
                  "population_per_gpu": P, "best_score_after": best_after, "wall_s": wall_ga, "island_mode": island_mode, "exchanges": int(epochs), "control": control},
           "probes": {"dfma_tflops": dfma_tflops, "ddiv_add_gops": ddiv_g, "l2_read_GBs_64MB": l2_gbs}}

    # ------------------------------------------------------------------ evaluate-only block (round 1's headline, kept for continuity)
    out["evaluate_only"] = evaluate_only(A, eng, clocks, args, world, barrier, max_over_ranks, tours, sizes, tables, n, P, E, args.workload)
    out["ga"]["rate_inside_ga_relative_to_evaluate_only"] = value / out["evaluate_only"]["value"]

    # ------------------------------------------------------------------ strong scaling: the same total population split over the GPUs
    if world > 1 and not args.no_extras and P % world == 0:
        Ps = P // world
        gens = 2000
        kw = dict(prm_kw, npop=Ps)
        from allhic_b200.islands import IslandGA
        ga = IslandGA(eng, init_pin.array, eng.ga_params(**kw), migrate_every=50, n_elite=2)
        barrier()
        s0 = eng.ga_advance(0)
        tour_s, best_s, st_s = ga.run(gens)
        s_s = max_over_ranks((st_s.device_ms - s0.device_ms) * 1e-3)
        s_ev = sum_over_ranks(float(st_s.evaluations - s0.evaluations))
        ga.close()
        single = None
        if rank == 0:                                    # the 1-GPU run it is compared with: one population of P, the same number of generations
            eng.ga_begin(init_pin.array, eng.ga_params(**prm_kw))
            a0 = eng.ga_advance(0); a1 = eng.ga_advance(gens)
            single = {"best_score": a1.best_score, "evals": int(a1.evaluations - a0.evaluations), "seconds": (a1.device_ms - a0.device_ms) * 1e-3}
            eng.ga_end()
        barrier()
        out["strong_scaling"] = {"population_total": P, "population_per_gpu": Ps, "generations": gens, "evals": int(s_ev), "seconds": s_s, "evals_per_s": s_ev / s_s,
                                 "best_score_islands": best_s, "single_gpu_same_total_population": single,
                                 "speedup_vs_single_gpu": (single["seconds"] / s_s) if single else None}

    if rank == 0 and not args.no_extras:
        # ---- dense (north_star literal) kernel, for the record
        sub = min(P, 2048)
        e_act, w_dense = work_per_tour(sizes, tables, tours[:sub], sample=32)
        eng.pop_set(tours[:sub])
        eng.pop_eval_m("dense", 1)
        ms = 0.0
        for _ in range(5):
            eng.flush_l2(); ms += eng.pop_eval_m("dense", 1)
        dk = ms * 1e-3 / 5
        out["dense_kernel"] = {"evals_per_s": sub / dk, "population": sub,
                               "roofline": {"bound": "hbm", "achieved": sub * (4 * w_dense + 12 * L) / dk / 1e9, "peak": HBM_PEAK, "unit": "GB/s",
                                            "frac": sub * (4 * w_dense + 12 * L) / dk / 1e9 / HBM_PEAK,
                                            "note": "M (%.0f MB int32) is %s" % (4.0 * n * n / 1e6, "L2-resident: bytes come from L2, not HBM" if 4 * n * n < 100e6 else "larger than L2")}}
        # ---- orientation phase
        out["orientation"] = orientation_legs(A, eng, tours, sizes, tables, n, args)
        if world == 1 and args.cpu_seconds > 0:
            from oracle_lib import OracleCLM
            orc = OracleCLM(clmf, ids)
            sg = np.where(np.random.default_rng(7).random(n) < 0.5, ord('+'), ord('-')).astype(np.uint8)
            orc.set_tour(tours[0].astype(np.int64)); orc.set_signs(sg)
            t0 = time.perf_counter(); orc.evaluate_q(literal=False); lk = time.perf_counter() - t0
            lit = None
            if 96.0 * n * n < 2e9:
                t0 = time.perf_counter(); orc.evaluate_q(literal=True); lit = time.perf_counter() - t0
            out["orientation"]["cpu_restatement"] = {"eval_q_lookup_per_s": 1.0 / lk, "eval_q_literal_per_s": (1.0 / lit) if lit else None,
                                                      "flip_one_pass_ms_estimate": (1 + L) * (lit or lk) * 1e3, "cores": 1,
                                                      "note": "one call each; flipOne estimated as (1 + L) x the literal EvaluateQ (lookup variant where dense Q is infeasible)"}

    out["clocks"] = clocks.stop()
    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N == 1 only)
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        if args.cpu_seconds > 0:
            gens = 2 if n >= 2000 else 20
            reps = 1
            ev, gn, dt, best = cpu_ga_sample(ids, clmf, init_pin.array, P, gens, threads)
            if dt < 0.5 * args.cpu_seconds:
                reps = max(1, int(args.cpu_seconds / max(dt, 1e-3)) - 1)
                ev2, gn2, dt2, best = cpu_ga_sample(ids, clmf, init_pin.array, P, gens, threads, repeats=reps)
                ev, gn, dt = ev + ev2, gn + gn2, dt + dt2
                reps += 1
            out["cpu_baseline"] = {"value": ev / dt, "unit": UNIT, "cores": threads, "kind": "port", "ga_generations_per_s": gn / dt,
                                   "sample": "%d x GARun for %d generations (npop %d, %d Tour.Evaluate calls, %.1f s): C restatement of evaluate.go:258-315 on an eaopt emulation, ParallelEval over %d threads"
                                             % (reps, gens, P, ev, dt, threads)}
            v, reps, dt = cpu_population_evals(ids, clmf, tours, min(4.0, args.cpu_seconds), threads)
            out["evaluate_only"]["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                                    "sample": "the same %d-tour population x %d passes (%.1f s), C restatement of Tour.Evaluate + eaopt ParallelEval chunking" % (P, reps, dt)}
        else:
            out["cpu_baseline"] = None
        if not args.no_extras:
            eng.close(); eng = None
            out["config1_cli"] = config1_cli(args)
            out["other_configs"] = {name: small_leg(A, name, args, dfma_tflops) for name in ("config2", "config5") if name != args.workload}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        emit(out)
    init_pin.close()
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
