#!/usr/bin/env python
"""Benchmark of the pyKriging fit hot path on B200 (BASELINE.json metric: FP64 NegLnLike evals/sec
at n=1024, k=10).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, candidates sharded (weak scaling)

One "step" = one GA/PSO generation: the concentrated negative ln-likelihood of a population of C
(theta, p) candidates for the n x k sample set -- what ``kriging.fittingObjective`` computes
(pyKriging/krige.py:429-452).  Inputs are synthetic: a seeded random Latin hypercube
(samplingplan.rlh recipe), y = Styblinski-Tang, candidates half uniform (what generate_population
draws) and half log-uniform in theta (late-generation-like).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SAMPLES, K_DIMS, POP = 1024, 10, 1024
# BASELINE.json's metric, verbatim; `value` is its first half (likelihood evaluations/s), the second half
# (predictions/s) is reported in the same line under "predictions"
METRIC = "FP64 NegLnLike evals/sec at n=1024,k=10 and predictions/sec, 1/2/4/8 B200"
UNIT = "evals/s"
PSI_FP64_PER_ENTRY = 103  # FP64 instructions of psi_pop_kernel<10,1> per Psi entry (SASS count: 72 DFMA + 15 DMUL + 14 DADD + 2 DSETP)


# ---- synthetic workload (no oracle import: this is the measured path's input) ---------------------------
def make_samples(n, k, seed=1234):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, k))
    for d in range(k):
        X[:, d] = rng.permutation(np.arange(1, n + 1, 1))
    X = (X - 0.5) / n
    Z = X * 10 - 5
    y = np.sum(Z ** 4 - 16 * Z ** 2 + 5 * Z, axis=1) / 2.0
    return X, y


def make_candidates(C, k, seed):
    rng = np.random.default_rng(seed)
    half = C // 2
    th_u = rng.uniform(1e-5, 100.0, (half, k))
    th_l = 10.0 ** rng.uniform(-5.0, 2.0, (C - half, k))
    theta = np.concatenate([th_u, th_l])
    p = rng.uniform(1.0, 2.0, (C, k))
    return np.ascontiguousarray(theta), np.ascontiguousarray(p)


PRED_N, PRED_K, PRED_GRID = 512, 2, 4096  # C5: EI over a 4096 x 4096 = 16 777 216-point grid, n = 512 model


def make_branin_samples(n, seed=5):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 2))
    for d in range(2):
        X[:, d] = rng.permutation(np.arange(1, n + 1, 1))
    X = (X - 0.5) / n
    x1, x2 = X[:, 0] * 15 - 5, X[:, 1] * 15
    y = (x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5 / np.pi * x1 - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10
    return X, y


def run_predictions(torch, dev, local_rank, rank, world, dist):
    """predictions/s (the second half of BASELINE.json's metric): yhat, s and EI for this rank's rows of
    the 16 Mi-point query grid on a fitted n = 512 model, query points resident in HBM."""
    from pykriging_b200.krige import kriging

    X, y = make_branin_samples(PRED_N)
    model = kriging(X, y, device=local_rank)
    model.update([5.0, 5.0, 1.8, 1.8])
    model.neglikelihood()
    h = model._pk_handle()
    m_total = PRED_GRID * PRED_GRID
    per = -(-m_total // world)
    lo, hi = min(rank * per, m_total), min((rank + 1) * per, m_total)
    m = hi - lo
    idx = torch.arange(lo, hi, device=dev, dtype=torch.int64)
    ax = torch.linspace(0.0, 1.0, PRED_GRID, dtype=torch.float64, device=dev)
    Xq = torch.stack([ax[idx // PRED_GRID], ax[idx % PRED_GRID]], dim=1).contiguous()
    del idx
    out = torch.empty(3, m, dtype=torch.float64, device=dev)
    theta = np.array([5.0, 5.0])
    pl = np.array([1.8, 1.8])
    y_min = float(np.min(model.y))
    stream = torch.cuda.ExternalStream(h.stream, device=dev)

    def go():
        h.predict_batch_raw(m, Xq, theta, pl, float(model.mu), float(model.SigmaSqr), y_min, -1.0, 0.0, out[0], out[1],
                            out[2])

    go()
    h.synchronize()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(reps):
            go()
        ev1.record(stream)
    h.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    best = float(out[2].max().item())
    # end to end through the C ABI with HOST buffers (pinned): upload of the query points, kernels and download of
    # yhat / s / EI inside the timed region (chunks of 1 Mi points, copies overlapped with the kernels)
    Xq_h = Xq.cpu().pin_memory()
    out_h = torch.empty(3, m, dtype=torch.float64).pin_memory()
    del Xq, out

    def go_host():
        h.predict_batch_raw(m, Xq_h, theta, pl, float(model.mu), float(model.SigmaSqr), y_min, -1.0, 0.0, out_h[0], out_h[1],
                            out_h[2])

    go_host()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    go_host()
    h.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    assert float(out_h[2].max().item()) == best
    e2e = {"value": m_total / (e2e_ms * 1e-3), "unit": "predictions/s", "ms": e2e_ms,
           "h2d_bytes": int(m_total * PRED_K * 8), "d2h_bytes": int(m_total * 3 * 8),
           "api": "pk_predict_batch (C ABI) with pinned host buffers"}
    del Xq_h, out_h
    return {"value": m_total / (ms * 1e-3), "unit": "predictions/s", "ms": ms, "e2e": e2e, "points": m_total, "n": PRED_N,
            "k": PRED_K, "outputs": "yhat, s, EI per point", "scaling": "strong (grid rows sharded over ranks)",
            "flops_per_point_algorithmic": PRED_N ** 2,
            "tflops_algorithmic": m_total * PRED_N ** 2 / (ms * 1e-3) * 1e-12, "max_ei": best}


def load_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed `ncu --set full` capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def config_dict(n_gpus):
    return {"workload": "C3: 10-D synthetic (rlh seed 1234, Styblinski-Tang), n=1024 samples, population of "
                        "1024 (theta,p) candidates per GPU per step (512 uniform + 512 log-uniform theta)",
            "n": N_SAMPLES, "k": K_DIMS, "candidates_per_gpu_per_step": POP,
            "parallelism": "candidates sharded, dp%d" % n_gpus,
            "l2": "per-step working set 8.5 GiB of Psi/L matrices per GPU >> 126 MB L2 (no explicit flush needed)"}


# ---- clocks sampler ------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---- reference arm / CPU baseline (the only place bench.py touches oracle/) -------------------------------
CPU_BASELINE_SAMPLE = 48  # candidates timed for cpu_baseline: ~10 s at ~4 evals/s on the box's 16 host cores


def cpu_reference_evals(n_cands, seeds):
    """Times the reference's CPU algorithm (NumPy restatement in reference order: np.power over the
    (n,n,k) distance tensor, dpotrf, six dgesv solves -- matrixops.py:24-56) on `n_cands` candidates of
    the same workload per seed.  Returns list of (seconds, evals)."""
    from oracle import oracle_np as onp

    X, y = make_samples(N_SAMPLES, K_DIMS)
    model = onp.OracleKriging(X, y)  # builds the distance tensor once, like matrixops.updateData
    out = []
    for s in seeds:
        theta, p = make_candidates(POP, K_DIMS, 1000 + s)
        idx = np.linspace(0, POP - 1, n_cands).astype(int)  # spans both candidate families
        cands = np.concatenate([theta[idx], p[idx]], axis=1).tolist()
        t0 = time.perf_counter()
        model.fittingObjective(cands)
        out.append((time.perf_counter() - t0, n_cands))
    return out


def cpu_reference_predictions(n_points):
    """predictions/s of the reference's per-point path (predict + predict_var + expimp per query point: the Python
    loop over the samples and the dgesv solves of matrixops.py:68-95, krige.py:201-234) on `n_points` points of the
    C5 grid, n = 512 model, oracle port in reference order."""
    from oracle import oracle_np as onp

    X, y = make_branin_samples(PRED_N)
    model = onp.OracleKriging(X, y)
    model.update([5.0, 5.0, 1.8, 1.8])
    model.neglikelihood()
    ax = np.linspace(0.0, 1.0, PRED_GRID)
    idx = np.linspace(0, PRED_GRID * PRED_GRID - 1, n_points).astype(np.int64)
    pts = np.stack([ax[idx // PRED_GRID], ax[idx % PRED_GRID]], axis=1)
    t0 = time.perf_counter()
    for x in pts:
        model.predict(x)
        model.predict_var(x)
        model.expimp(x)
    secs = time.perf_counter() - t0
    return {"value": n_points / secs, "unit": "predictions/s", "points": int(n_points), "n": PRED_N, "k": PRED_K,
            "outputs": "yhat, s, EI per point",
            "sample": "%d points of the 4096 x 4096 grid, one at a time as the reference does" % n_points}


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 4
    res = cpu_reference_evals(sample, list(range(args.warmup + args.steps)))
    timed = res[args.warmup:]
    secs = sum(t for t, _ in timed)
    evals = sum(e for _, e in timed)
    value = evals / secs
    cores = os.cpu_count()
    pred = cpu_reference_predictions(24)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the %d candidates of each step (evenly spaced over both candidate "
                                       "families), NumPy/OpenBLAS oracle port of matrixops.py:24-56 with all host "
                                       "threads" % (sample, POP)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "predictions": pred,
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- this repo's arm -----------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    from pykriging_b200.krige import kriging

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    X, y = make_samples(N_SAMPLES, K_DIMS)
    model = kriging(X, y, device=local_rank)   # normalises (krige.py:117-133) and uploads X, y
    h = model._pk_handle()
    h.set_cholesky_variant(args.chol_variant)
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    n_steps = args.warmup + args.steps
    C, k = POP, K_DIMS
    th_host, p_host = [], []
    for s in range(n_steps):
        th, p = make_candidates(C, k, 1000 + s + 7919 * rank)
        th_host.append(torch.from_numpy(th).pin_memory())
        p_host.append(torch.from_numpy(p).pin_memory())
    th_dev = [t.to(dev) for t in th_host]
    p_dev = [t.to(dev) for t in p_host]
    out_dev = torch.empty(4, C, dtype=torch.float64, device=dev)
    info_dev = torch.empty(C, dtype=torch.int32, device=dev)
    out_host = torch.empty(4, C, dtype=torch.float64).pin_memory()
    info_host = torch.empty(C, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def step_device(s):
        h.nll_batch_raw(C, th_dev[s], p_dev[s], None, 0, out_dev[0], out_dev[1], out_dev[2], out_dev[3], info_dev)

    def step_host(s):
        h.nll_batch_raw(C, th_host[s].numpy(), p_host[s].numpy(), None, 0, out_host[0].numpy(), out_host[1].numpy(),
                        out_host[2].numpy(), out_host[3].numpy(), info_host.numpy())

    peak = h.fp64_peak()
    for s in range(args.warmup):
        step_device(s)
    h.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        h.synchronize()

    # ---- timed region 1: inputs resident in HBM, device events on the launching stream ------------------
    # (the library's own pipelining is on: Psi of the second part of the population overlaps the first
    # part's factorisation on a second stream; everything is ordered behind the handle's main stream)
    sampler = ClockSampler(local_rank)
    launches0 = h.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for s in range(args.warmup, n_steps):
            step_device(s)
        ev1.record(stream)
    h.synchronize()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = h.launch_count - launches0
    fails = int((info_dev != 0).sum().item())

    # ---- per-kernel times for the roofline: the SAME steps once more with the phases serialised (profiling
    # mode brackets every kernel with CUDA events on its stream and disables the Psi/Cholesky overlap, so a
    # kernel's duration is its own) ------------------------------------------------------------------------
    h.set_profiling(True)
    h.phase_times()
    for s in range(args.warmup, n_steps):
        step_device(s)
    h.synchronize()
    phases = h.phase_times()
    h.set_profiling(False)
    barrier()

    # ---- timed region 2: end to end through the C ABI with HOST buffers (H2D + compute + D2H) ---------------
    for s in range(min(args.warmup, 2)):
        step_host(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, n_steps):
        step_host(s)
    h.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()
    evals = world * C * args.steps
    value = evals / (ms_max * 1e-3)
    e2e_value = evals / (e2e_ms_max * 1e-3)

    pred = None
    if not args.no_predictions:
        pred = run_predictions(torch, dev, local_rank, rank, world, dist)
    traffic = load_traffic()
    if rank == 0:
        if pred is not None:
            pred["kernels"] = ("psi_rows_kernel<2> (psi rows + yhat, FP64 ALU, HBM scratch) + tn::trinorm_kernel "
                               "(|L^-1 psi|^2 on the DMMA pipe, TMA-fed, + s + EI)")
            pred["frac_of_dmma_peak"] = pred["tflops_algorithmic"] / (peak["dmma_tflops"] * world)  # whole-job TFLOP/s over N GPUs
        n = N_SAMPLES
        flops_chol = (n ** 3) / 3.0 * C * args.steps
        chol_tf = flops_chol / (phases["cholesky"] * 1e-3) * 1e-12 if phases["cholesky"] > 0 else None
        nt = n // 64
        psi_bytes = 8.0 * (nt * (nt + 1) // 2) * 64 * 64 * C * args.steps
        hbm_peak, hbm_src = 6650.0, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        psi_gbs = psi_bytes / (phases["psi"] * 1e-3) * 1e-9 if phases["psi"] > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * C * k * 8,
                    "d2h_bytes_per_step": 4 * C * 8 + 4 * C, "api": "pk_nll_batch (C ABI) with pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "cta::chol_cta_kernel (persistent blocked Cholesky, one CTA per candidate, "
                                   "DMMA 8x8x4, TMA-fed stages)" if args.chol_variant != 1 else "chol_first_kernel + chol_step_kernel",
                         "bound": "tensor", "achieved": chol_tf, "peak": peak["dmma_tflops"], "unit": "TFLOP/s",
                         "frac": (chol_tf / peak["dmma_tflops"]) if chol_tf else None,
                         "traffic": (traffic or {}).get("chol_cta_kernel_dram_bytes_per_launch"),
                         "traffic_source": (traffic or {}).get("source"),
                         "algorithmic": "n^3/3 flop per eval = %.4g flop x %d evals per step" % (n ** 3 / 3.0, C),
                         "peak_source": "FP64 DMMA peak measured in this process (pk_fp64_peak; "
                                        "MEASURED_PEAKS.json has no FP64 entry)",
                         "timing": "CUDA events around every launch of the kernel on its stream, over the same %d "
                                   "steps as the timed region, run once more with the Psi/Cholesky overlap "
                                   "switched off (pk_set_profiling) so that the duration is the kernel's own"
                                   % args.steps,
                         "phase_ms_per_step": {kk: v / args.steps for kk, v in phases.items()}},
            "roofline_psi": {"kernel": "psi_pop_kernel<10,1>", "bound": "hbm", "achieved": psi_gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": (psi_gbs / hbm_peak) if psi_gbs else None,
                             "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)",
                             "algorithmic": "8 B per lower-triangle tile entry written; FP64-ALU bound for general p "
                                            "(10 table-driven 2^x + 1 e^-S per entry = 103 FP64 instructions, see fp64_pipe_frac); "
                                            "issue bound: an FP64 instruction takes two dispatch cycles",
                             "fp64_instr_per_entry": PSI_FP64_PER_ENTRY,
                             "fp64_pipe_frac": (PSI_FP64_PER_ENTRY * (n * (n - 1) / 2) * C * args.steps * 2.0 /
                                                (phases["psi"] * 1e-3) * 1e-12 / peak["dfma_tflops"])
                             if phases["psi"] > 0 else None},
            "predictions": pred,
            "fp64_peak": peak, "failed_factorisations_last_step": fails,
        }
        if world == 1 and not args.no_cpu_baseline:
            if pred is not None:
                pred["cpu_baseline"] = cpu_reference_predictions(24)
            res = cpu_reference_evals(CPU_BASELINE_SAMPLE, [0])
            secs, ev = res[0]
            line["cpu_baseline"] = {"value": ev / secs, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "%d candidates of step 0 (evenly spaced over both families) through "
                                              "the NumPy/OpenBLAS oracle port of matrixops.py:24-56, all host "
                                              "threads (about 10 s of CPU work)" % CPU_BASELINE_SAMPLE}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predictions", action="store_true")
    ap.add_argument("--chol-variant", type=int, default=0, help="0 auto, 1 per-block-column launches, 2 persistent CTA")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
