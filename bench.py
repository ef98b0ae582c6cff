#!/usr/bin/env python
"""bench.py -- tour fitness evaluations/s of the `allhic optimize` hot path.

One "step" = one pass of batched Tour.Evaluate ("EvaluateM", evaluate.go:116)
over a population of P synthetic tours of one linkage group.
  value  evals/s with the population resident in HBM (CUDA-event time of the
         evaluate kernel, L2 flushed between steps)
  e2e    evals/s through the C ABI call a host makes (ahx_eval_m_u16): pinned
         host tours -> H2D -> kernel -> D2H scores, all inside the timed region
  ga     (extra) generations/s and evals/s of the device-resident GA, as an
         island model over NCCL when launched on N > 1 GPUs
`--impl reference` times the CPU restatement of the reference's own
implementation (oracle/, C, all host threads) on the same workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config3] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LIMIT = 10_000_000
HBM_PEAK = 6544.3   # GB/s, fallback when MEASURED_PEAKS.json is absent (it is driver-written per pod)
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as _f:
        HBM_PEAK = float(json.load(_f)["hbm_gbs"])
except Exception:
    pass
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the evaluate kernel, from the committed
# ncu --set full capture of (workload, P) -- profiles/r1r_ncu_full_summary.txt
NCU_DRAM_BYTES_PER_LAUNCH = {("config3", 8000): 64223232 + 508160}
METRIC = "tour fitness evals/sec"
UNIT = "evals/s"


_JSON_FD = None


def emit(obj):
    """The bench's one JSON line, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def workload(name):
    from allhic_b200 import synth
    cfg = synth.CONFIGS[name]
    ids, clm, _ = synth.materialize(name)
    desc = {"config2": "synthetic 500-contig linkage group, population 1k, EvaluateM (BASELINE configs[1])",
            "config3": "synthetic 2,000-contig linkage group, population 8k, EvaluateM (BASELINE configs[2], the north_star target size)",
            "config5": "synthetic 10,000-contig linkage group, population 8k, EvaluateM (BASELINE configs[4])"}[name]
    return ids, clm, cfg["n"], cfg["npop"], desc


def make_tours(n, P, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)


def work_per_tour(sizes, tables, tours, sample=128):
    """Mean pair-terms per tour: E_act (contig-space kernel: stored pairs within
    LIMIT) and W (position-space window pairs of the reference loop)."""
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
    """Reference arm: the reference's own CPU implementation of the path.  The Go
    binary cannot be built in this image (no Go toolchain), so this is the C
    restatement in oracle/ (kind=port) on all host threads; rank 0 only."""
    if rank != 0:
        return
    ids, clm, n, P, desc = workload(args.workload)
    threads = os.cpu_count() or 1
    from oracle_lib import OracleCLM
    orc = OracleCLM(clm, ids)
    S = P                                       # the whole population per step (0.3-2 s of CPU work per step)
    tours = np.ascontiguousarray(make_tours(n, S, 1234), np.int64)
    for _ in range(max(1, args.warmup)):
        orc.population_evaluate(tours[: threads * 2], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.population_evaluate(tours, nthreads=threads)
    dt = time.perf_counter() - t0
    v = args.steps * S / dt
    sample = "%d of the %d tours per step x %d steps, oracle/allhic_oracle.c orc_population_evaluate, %d threads" % (S, P, args.steps, threads)
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": desc, "n_contigs": n, "population": P, "sample_per_step": S},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config5"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ga-gens", type=int, default=2000)
    ap.add_argument("--no-ga", action="store_true")
    ap.add_argument("--no-dense", action="store_true")
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

    if rank == 0:
        ids, clmf, n, P, desc = workload(args.workload)
    barrier()
    ids, clmf, n, P, desc = workload(args.workload)       # other ranks reuse the files rank 0 wrote
    clm = A.CLM(clmf, ids)
    tables, sizes = clm.tables(), clm.sizes()
    E = len(tables["pa"])
    eng = A.Engine(local)
    eng.load(clm)
    tours = make_tours(n, P, 1000 + rank)                  # weak scaling: P tours per GPU
    L = n
    e_act, w_dense = work_per_tour(sizes, tables, tours)

    # pinned host buffers for the e2e call
    use16 = n <= 65535
    pin_t = A.PinnedArray((P, L), np.uint16 if use16 else np.int32)
    pin_o = A.PinnedArray((P,), np.float64)
    pin_t.array[:] = tours
    h2d, d2h = pin_t.nbytes, pin_o.nbytes

    clocks = ClockSampler(local); clocks.start()
    eng.pop_set(tours)
    dfma_tflops, ddiv_g = eng.probe_fp64()
    l2_gbs = eng.probe_l2(64 << 20)
    for _ in range(max(3, args.warmup)):
        eng.flush_l2(); eng.pop_eval_m("sparse", 1)
    for _ in range(max(3, args.warmup)):
        eng.eval_m(pin_t.array, out=pin_o.array)

    # ---- device-resident steps (value)
    barrier()
    lc0 = eng.launch_count()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.flush_l2()                                    # not timed: evict the tours/COO from L2
        dev_ms += eng.pop_eval_m("sparse", 1)             # CUDA events around the kernel on its own stream
    eval_launches = (eng.launch_count() - lc0) - args.steps   # minus the flush kernels
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    scores_dev = eng.pop_scores()
    # ---- end-to-end steps (e2e): host buffers through the C ABI
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.eval_m(pin_t.array, out=pin_o.array)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert np.allclose(scores_dev, pin_o.array, rtol=1e-13, atol=0), "resident and e2e paths disagree"

    value = world * P * args.steps / (dev_ms * 1e-3)
    e2e_v = world * P * args.steps / e2e_s
    kern_s = dev_ms * 1e-3 / args.steps
    terms_per_s = P * e_act / kern_s                       # per GPU (roofline is a per-kernel figure)
    # Roofline of the dominant kernel.  eval_m_res_kernel evaluates EVERY stored pair of every
    # tour branch-free (the window test only masks the reciprocal seed), so per tour the algorithmic work is
    #   fp64: E terms x 5 fp64-pipe instructions (exact u32->f64, 3 FMA of the cubic reciprocal step, 1 FMA to
    #         accumulate) = 10 flop per stored pair, against the measured DFMA peak;
    #   smem: E x two 4-byte coordinate gathers from shared memory, against 128 B/clk/SM;
    #   hbm : 4*L + 8 bytes (the tour in, the score out), against MEASURED_PEAKS.json.
    # The slowest leg at its peak is the bound.
    sm_mhz = clocks.peak_mhz()
    smem_peak = 128.0 * eng.sm_count() * sm_mhz * 1e6 / 1e9                      # GB/s: 128 B/clk/SM
    kind, tpb = eng.eval_m_kernel(P, L)
    flop_per_pair = 10.0 if kind >= 2 else 2.0 * e_act / max(E, 1)               # the streaming kernels divide in-window pairs only
    legs = {"smem": (P * 8.0 * E / kern_s / 1e9, smem_peak, "GB/s"),
            "fp64": (P * E * flop_per_pair / kern_s / 1e12, dfma_tflops, "TFLOP/s"),
            "hbm": (P * (4.0 * L + 8.0) / kern_s / 1e9, HBM_PEAK, "GB/s")}
    bound = max(legs, key=lambda k: legs[k][0] / legs[k][1])
    roof = {"bound": bound, "achieved": legs[bound][0], "peak": legs[bound][1], "unit": legs[bound][2],
            "frac": legs[bound][0] / legs[bound][1],
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((args.workload, P)),
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel (profiles/), bytes per launch",
            "peak_source": {"smem": "128 B/clk/SM x %d SMs x %.0f MHz (SM clock sampled during the timed region)" % (eng.sm_count(), sm_mhz),
                            "fp64": "DFMA microbenchmark measured live by this run (ahx_probe_fp64; MEASURED_PEAKS.json has no fp64 figure)",
                            "hbm": "MEASURED_PEAKS.json hbm_gbs"}[bound],
            "kernel": {3: "eval_m_res_kernel (table streamed)", 2: "eval_m_res_kernel", 1: "eval_m_x32_kernel", 0: "eval_m_sparse_kernel"}[kind], "tours_per_batch": tpb,
            "pair_table_entries": eng.pair_table_entries(tpb) if kind in (2, 3) else None,
            "algorithmic_per_tour": {"stored_pairs_E": E, "fp64_flop": flop_per_pair * E, "smem_gather_bytes": 8 * E,
                                     "in_window_pairs_E_act": e_act, "hbm_bytes": 4 * L + 8},
            "issue_model": {"issue_cycles_per_pair_term": 16.06, "note": "SASS of the consumer loop: 5 fp64 (2 issue cycles each, term_probe.cu) + 6.06 other instructions per (pair, tour); "
                            "frac = terms/s x cycles / (SMs x 4 schedulers x clock)",
                            "frac": P * E * 16.06 / 32.0 / kern_s / (eng.sm_count() * 4 * sm_mhz * 1e6)},
            "measured_div_add_peak_per_s": ddiv_g * 1e9,
            "legs": {k: {"achieved": v[0], "peak": v[1], "unit": v[2], "frac": v[0] / v[1]} for k, v in legs.items()},
            "dense_equivalent": {"window_pairs_per_tour": w_dense, "bytes_per_tour": 4 * w_dense + 12 * L,
                                 "algorithmic_GBs": P * (4 * w_dense + 12 * L) / kern_s / 1e9, "hbm_peak_GBs": HBM_PEAK,
                                 "note": "north_star's dense M-gather bytes at this eval rate: the algorithmic speed-up of the contig-space formulation, not a roofline fraction"}}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": desc, "n_contigs": n, "population_per_gpu": P, "contig_pairs_E": E, "l2": "flushed between device-timed steps (write of 2x L2)",
                      "parallelism": "independent sub-populations, one per GPU" if world > 1 else "single GPU"},
           "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "call": "ahx_eval_m_u16" if use16 else "ahx_eval_m", "ms_per_step": e2e_s / args.steps * 1e3},
           "gpu_launches": int(eval_launches), "roofline": roof,
           "probes": {"dfma_tflops": dfma_tflops, "ddiv_add_gops": ddiv_g, "l2_read_GBs_64MB": l2_gbs}}

    # ---- dense (north_star literal) kernel, for the record
    if not args.no_dense and rank == 0:
        sub = min(P, 2048)
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

    # ---- GA (generations/s), islands over NCCL when world > 1
    if not args.no_ga:
        prm = eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=42)
        init = tours[0]
        if world > 1:
            from allhic_b200.islands import IslandGA
            ga = IslandGA(eng, init, prm, migrate_every=50, n_elite=2)
            ga.run(50)                                    # warm-up
            barrier(); t0 = time.perf_counter()
            _, best, st = ga.run(args.ga_gens)
            torch.cuda.synchronize(); ga_s = max_over_ranks(time.perf_counter() - t0)
            ga.close()
        else:
            eng.ga_begin(init, prm)
            st0 = eng.ga_advance(50)
            barrier(); t0 = time.perf_counter()
            st = eng.ga_advance(args.ga_gens)
            ga_s = time.perf_counter() - t0
            best = st.best_score
            eng.ga_end()
            st = type("S", (), dict(evaluations=st.evaluations - st0.evaluations))
        ev = torch.tensor([float(st.evaluations)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ev)
        out["ga"] = {"generations_per_s": args.ga_gens / ga_s, "population_per_gpu": P, "generations": args.ga_gens,
                     "evals_per_s_inside_ga": float(ev.item()) / ga_s if world == 1 else None, "best_score": best,
                     "model": "islands, all-gather of 2 elites every 50 generations over NCCL" if world > 1 else "single population"}

    # ---- orientation phase (CLM.EvaluateQ batched over sign vectors; one flipOne pass), for the record
    if not args.no_ga and rank == 0:
        rng = np.random.default_rng(7)
        nq = 1024
        signs = np.where(rng.random((nq, n)) < 0.5, ord('+'), ord('-')).astype(np.uint8)
        eng.eval_q(tours[0], signs[:8])
        q_t, f_t = [], []
        for _ in range(15):                       # medians: these calls are short and host-timed
            t0 = time.perf_counter()
            eng.eval_q(tours[0], signs)
            q_t.append(time.perf_counter() - t0)
        for _ in range(7):
            t0 = time.perf_counter()
            s2, *_ = eng.flip_one(tours[0], signs[0])
            f_t.append(time.perf_counter() - t0)
        q_s, f_s = float(np.median(q_t)), float(np.median(f_t))
        out["orientation"] = {"eval_q_per_s_e2e": nq / q_s, "batch": nq, "flip_one_pass_ms_e2e": f_s * 1e3, "tour_length": int(L),
                              "note": "host buffers through ahx_eval_q / ahx_flip_one, medians of 15 / 7 calls; the reference does 1 + L full EvaluateQ per flipOne pass"}
        if world == 1 and args.cpu_seconds > 0:
            # the CPU restatement beside it: EvaluateQ as the reference does it (dense 96*N^2-byte Q() rebuilt per call,
            # orientation.go:137-193) and with the map lookup that avoids Q(); a flipOne pass = (1 + L) such calls
            from oracle_lib import OracleCLM
            orc = OracleCLM(clmf, ids)
            orc.set_tour(tours[0].astype(np.int64)); orc.set_signs(signs[0])
            t0 = time.perf_counter(); orc.evaluate_q(literal=False); lk = time.perf_counter() - t0
            lit = None
            if 96.0 * n * n < 2e9:
                t0 = time.perf_counter(); orc.evaluate_q(literal=True); lit = time.perf_counter() - t0
            out["orientation"]["cpu_restatement"] = {"eval_q_lookup_per_s": 1.0 / lk, "eval_q_literal_per_s": (1.0 / lit) if lit else None,
                                                      "flip_one_pass_ms_estimate": (1 + L) * (lit or lk) * 1e3, "cores": 1,
                                                      "note": "one call each; flipOne estimated as (1 + L) x the literal EvaluateQ (lookup variant where dense Q is infeasible)"}

    out["clocks"] = clocks.stop()
    # ---- CPU baseline beside it (rank 0, N == 1 only)
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        v, reps, dt = cpu_population_evals(ids, clmf, tours, args.cpu_seconds, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "the same %d-tour population x %d passes (%.1f s), C restatement of Tour.Evaluate + eaopt ParallelEval chunking, %d threads" % (P, reps, dt, threads)}
        if not args.no_ga:
            # the same GA on the CPU restatement (eaopt emulation, ParallelEval over all threads), bounded to a few seconds
            from oracle_lib import OracleCLM
            orc = OracleCLM(clmf, ids)
            orc.set_tour(tours[0].astype(np.int64))
            t0 = time.perf_counter()
            r = orc.ga_run(npop=P, ngen=1 << 30, mutprob=0.2, seed=42, max_gens=1 << 30, max_seconds=min(6.0, max(0.5, args.cpu_seconds)), nthreads=threads)   # 0 would mean no time limit
            dt = time.perf_counter() - t0
            out["cpu_baseline"]["ga_generations_per_s"] = r.generations / dt
            out["cpu_baseline"]["ga_sample"] = "%d generations of the same GA (npop %d) in %.1f s" % (r.generations, P, dt)
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        emit(out)
    pin_t.close(); pin_o.close(); eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
