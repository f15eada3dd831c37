#!/usr/bin/env python
"""Benchmark of the EM-GMM hot path (BASELINE.json: EM iter/s & point*comp evals/s, N=1e8 d=16 K=64 full cov).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm on the host cores

A step is one EM iteration (ExpStep + MaxStep, src/unsupervised/mlf_emgmm.f90:147-149) over the whole synthetic
data set (sharded over the ranks).  `value` is timed with inputs resident in HBM; `e2e` goes through the same C-ABI
with HOST buffers: every step re-uploads the rank's X from pinned memory, runs mlf_step(1) and reads the
parameters and sumLL back.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "EM point*component evaluations/s (N=1e8, d=16, K=64 full covariance); em_iter_per_s alongside"
UNIT = "point*comp/s"
SIGMA_C, SIGMA_X, SEED = 8.0, 1.0, 3


def flops_e(d):  # E = d^2 + 4d + 6 flops + one exp (40 flops) per (point, component)
    return d * d + 4 * d + 6 + 40


def flops_m(d):  # M = d^2 + 3d + 4
    return d * d + 3 * d + 4


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.3):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index: int) -> str:
    """Pin this rank's host threads (and therefore its pinned staging buffer, placed at first touch / pin time) to the CPU
    cores NVML reports as local to its GPU.  Without it the eight ranks' host buffers tend to land on one NUMA node and the
    per-step H2D of the e2e path shares that node's memory and PCIe root."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} cores local to GPU {gpu_index}"
        return "no local cores reported"
    except Exception as exc:  # noqa: BLE001 -- affinity is an optimisation, never a failure
        return f"not bound ({type(exc).__name__})"


def cpu_oracle_rate(d, K, n_cpu, min_proba, steps, warmup, threads=None, sigma_c=SIGMA_C, sigma_x=SIGMA_X):
    """Times the CPU restatement of the reference algorithm (oracle/, kind "port") on a bounded sample."""
    from mlfortran_b200 import synth
    from oracle.emgmm_oracle import get

    orc = get()
    nthr = threads or min(K, os.cpu_count() or 1)
    orc.set_threads(nthr)
    data = synth.make_mixture(d, K, max(n_cpu // K, 1), sigma_c, sigma_x, SEED)
    X = data["X"]
    Mu, Cov, lam = synth.injected_init(data["centres"], SEED)
    if warmup > 0:
        Mu, Cov, lam, _ = orc.em_iterations(X, Mu, Cov, lam, min_proba, warmup)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        Mu, Cov, lam, _ = orc.em_iterations(X, Mu, Cov, lam, min_proba, 1)
        per.append(time.perf_counter() - t0)
    n = X.shape[0]
    tot = float(np.sum(per))
    return {"value": n * K * steps / tot, "sec_per_step": tot / steps, "n": int(n), "threads": int(nthr), "lapack": bool(orc.has_lapack)}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU algorithm for this path on the box's host cores.  The reference
    itself (Fortran + LAPACK + HDF5) cannot be compiled in this image, so this times the oracle port."""
    if rank != 0:
        return
    d, K = args.dims, args.comps
    n_cpu = args.cpu_points
    r = cpu_oracle_rate(d, K, n_cpu, args.min_proba, args.steps, args.warmup, sigma_c=args.sigma_c, sigma_x=args.sigma_x)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3 * (args.points / r["n"]), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "em_iter_per_s": r["value"] / (args.points * K),
        "config": {"workload": f"synthetic GMM N={args.points:.0e} d={d} K={K} full covariance (BASELINE configs[2])", "N": args.points, "d": d,
                   "K": K, "minProba": args.min_proba, "note": "ms_per_step and em_iter_per_s are linear extrapolations from the CPU sample to N"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": f"N_cpu={r['n']} points of the same d={d}, K={K} mixture, {args.steps} EM iterations, OpenMP over components, "
                                   f"LAPACK={'OpenBLAS(scipy)' if r['lapack'] else 'built-in Jacobi'}; reference layout needs 8*N*K bytes so N=1e8 cannot run on the host"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_infer(args):
    """`--mode infer`: getClass / getProba on caller-owned HOST data through the C-ABI, pinned and
    pageable, in points/s, next to the PCIe bound measured in the same run with plain pinned cudaMemcpy probes."""
    import ctypes as C

    import torch

    from mlfortran_b200 import _lib, mlf_algo_emgmm, synth

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, K = args.dims, args.comps
    nq = int(args.infer_points)
    ntrain = min(nq, 4_000_000)
    X, centres, _ = synth.make_mixture_device(d, K, nq, 0, nq, args.sigma_c, args.sigma_x, SEED, dev)
    Mu0, Cov0, lam0 = synth.injected_init(centres, SEED)
    em = mlf_algo_emgmm()
    if em.init(X[:ntrain].contiguous(), K, args.min_proba) != 0:
        raise SystemExit("bench.py: init failed")
    em.inject(Mu0, Cov0, lam0)
    em.step(2)
    lib = _lib.load()
    Xpin = torch.empty((nq, d), dtype=torch.float64, pin_memory=True)
    Xpin.copy_(X)
    torch.cuda.synchronize(dev)
    Xpage = Xpin.numpy().copy()
    cl_pin = torch.zeros(nq, dtype=torch.int32, pin_memory=True)
    cl_page = np.zeros(nq, dtype=np.int32)
    npb = min(nq, int(args.infer_proba_points))
    P_pin = torch.empty((K, npb), dtype=torch.float64, pin_memory=True)
    P_page = np.empty((K, npb))

    def timed(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            rc = fn()
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
            if rc != 0:
                raise SystemExit(f"bench.py: inference call failed: {em.last_error()}")
        return best

    # PCIe probes: one pinned cudaMemcpy each way, same bytes as the calls move
    dbuf = torch.empty((nq, d), dtype=torch.float64, device=dev)
    dP = torch.empty((K, npb), dtype=torch.float64, device=dev)
    t_h2d = timed(lambda: (dbuf.copy_(Xpin, non_blocking=True), 0)[1])
    t_d2h = timed(lambda: (P_pin.copy_(dP, non_blocking=True), 0)[1])
    h2d_gbs, d2h_gbs = 8 * nq * d / t_h2d * 1e-9, 8 * K * npb / t_d2h * 1e-9
    res = {}
    res["getClass_pinned"] = nq / timed(lambda: lib.mlf_getClass(em._h, Xpin.data_ptr(), cl_pin.data_ptr(), d, nq))
    res["getClass_pageable"] = nq / timed(lambda: lib.mlf_getClass(em._h, Xpage.ctypes.data, cl_page.ctypes.data, d, nq))
    res["getProba_pinned"] = npb / timed(lambda: lib.mlf_getProba(em._h, Xpin.data_ptr(), P_pin.data_ptr(), d, npb, K))
    res["getProba_pageable"] = npb / timed(lambda: lib.mlf_getProba(em._h, Xpage.ctypes.data, P_page.ctypes.data, d, npb, K))
    assert np.array_equal(cl_pin.numpy(), cl_page)
    bound_class = 1.0 / (8 * d / (h2d_gbs * 1e9))                       # H2D of X bounds getClass (labels are 4 B/point)
    bound_proba = 1.0 / max(8 * d / (h2d_gbs * 1e9), 8 * K / (d2h_gbs * 1e9))   # full duplex: the larger direction bounds getProba
    line = {"metric": "inference on host data (points/s)", "mode": "infer", "unit": "points/s", "n_gpus": 1, "dtype": "f64",
            "config": {"workload": f"getClass on {nq:.0e} / getProba on {npb:.0e} caller-owned host points, d={d} K={K}", "d": d, "K": K},
            "value": res["getClass_pinned"], "results": res,
            "pcie": {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs, "bound_getClass": bound_class, "bound_getProba": bound_proba},
            "frac_of_pcie_bound": {"getClass_pinned": res["getClass_pinned"] / bound_class, "getClass_pageable": res["getClass_pageable"] / bound_class,
                                   "getProba_pinned": res["getProba_pinned"] / bound_proba, "getProba_pageable": res["getProba_pageable"] / bound_proba}}
    print(json.dumps(line), flush=True)
    em.finalize()


def parity_spot_check(X, d, K, Mu0, Cov0, lam0, min_proba, n_check=1_000_000, iters=2):
    """the first n_check points of the bench's own data set copied back to the host and run through the oracle
    from the bench's initial parameters; the CUDA path (through the C-ABI, a second small object) does the same iterations on the
    same points.  Outside every timed region; part of the cpu_baseline leg (the one place bench.py may execute oracle/ as checker)."""
    import numpy as np

    from mlfortran_b200 import mlf_algo_emgmm
    from oracle.emgmm_oracle import get

    ns = int(min(n_check, X.shape[0]))
    Xs = np.ascontiguousarray(X[:ns] if isinstance(X, np.ndarray) else X[:ns].cpu().numpy())
    o = get()
    oMu, oCov, olam, otr = o.em_iterations(Xs, Mu0, Cov0, lam0, min_proba, iters)
    em = mlf_algo_emgmm()
    if em.init(Xs, K, min_proba) != 0:
        return {"error": "init failed"}
    em.inject(Mu0, Cov0, lam0)
    info, _ = em.step(iters)
    if info < 0:
        return {"error": em.last_error()}
    tr = em.sumLL_trace()
    nq = min(20000, ns)
    _, P = em.getProba(Xs[:nq])
    _, oP, _ = o.exp_step(Xs[:nq], em.Mu, em.Cov, em.lambda_)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    out = {"sample": f"first {ns} points of rank 0's shard, {iters} EM iterations from the bench's initial parameters: CUDA path through the C-ABI vs the oracle; "
                     f"E-step (getProba) on the first {nq} of them with the final parameters",
           "sumLL_rel": float(np.abs((tr - otr) / otr).max()), "Mu_rel": rel(em.Mu, oMu), "Cov_rel": rel(em.Cov, oCov),
           "lambda_abs": float(np.abs(em.lambda_ - olam).max()), "proba_abs": float(np.abs(P - oP).max()),
           "logits": em.qf_stats(), "tolerances": {"sumLL_rel": 1e-10, "Mu_rel": 1e-9, "Cov_rel": 1e-9, "lambda_abs": 1e-9, "proba_abs": 1e-9}}
    out["pass"] = bool(out["sumLL_rel"] <= 1e-10 and out["Mu_rel"] <= 1e-9 and out["Cov_rel"] <= 1e-9 and out["lambda_abs"] <= 1e-9 and out["proba_abs"] <= 1e-9)
    em.finalize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=float, default=1e8, help="total points N (all ranks)")
    ap.add_argument("--dims", type=int, default=16)
    ap.add_argument("--comps", type=int, default=64)
    ap.add_argument("--min-proba", type=float, default=1.0, help="minProba as in tests/test_emgmm.f90:125")
    ap.add_argument("--cov-type", default="full", choices=["full", "diag"], help="diag = extension (BASELINE configs[4]); full = reference")
    ap.add_argument("--sigma-c", type=float, default=SIGMA_C, help="spread of the cluster centres (tests/test_emgmm.f90 sigmaC); 8 = headline, "
                    "<= 2 gives heavily overlapping clusters (worst case of the threshold protocol)")
    ap.add_argument("--sigma-x", type=float, default=SIGMA_X, help="cluster width scale (tests/test_emgmm.f90 sigmaX)")
    ap.add_argument("--cpu-points", type=int, default=1000000, help="bounded CPU sample for cpu_baseline / --impl reference (SURVEY 8d: N_cpu = 1e6)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer: getClass / getProba throughput on host data (not the contract line)")
    ap.add_argument("--infer-points", type=float, default=2e7)
    ap.add_argument("--infer-proba-points", type=float, default=5e6)
    ap.add_argument("--host-buffer", default=os.environ.get("MLF_BENCH_HOST_BUFFER", "pinned"), choices=["pinned", "wc"],
                    help="e2e leg: the caller's host array is page-locked (torch pin_memory) or page-locked write-combined (mlf_b200_host_alloc)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.points = int(args.points)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "infer":
        if rank == 0:
            run_infer(args)
        return
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)  # the driver's default is >= 3; tiny dev runs may pass less

    import torch
    import torch.distributed as dist

    from mlfortran_b200 import mlf_algo_emgmm, parallel, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mlfortran_b200 has no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # the single rank is bound too (its pinned staging buffer should sit on the GPU's NUMA node); the CPU-baseline leg
    # gets the full affinity mask back
    full_affinity = os.sched_getaffinity(0)
    numa_note = bind_to_gpu_numa_node(local_rank) if (world > 1 or os.environ.get("MLF_BENCH_NO_BIND") != "1") else "not bound (MLF_BENCH_NO_BIND=1)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d, K, N = args.dims, args.comps, args.points
    lo, hi = parallel.shard_bounds(N, rank, world)
    n_loc = hi - lo
    # the library's own device generator (mlf_b200_synth_gmm, recipe of tests/test_emgmm.f90:105-124): every point is a pure
    # function of (seed, global index), so the N-point data set is the same whatever the number of ranks
    X, centres, _ = synth.make_mixture_device(d, K, n_loc, lo, N, args.sigma_c, args.sigma_x, SEED, dev)
    Mu0, Cov0, lam0 = synth.injected_init(centres, SEED)

    em = mlf_algo_emgmm()
    if em.init(X, K, args.min_proba, cov_type=args.cov_type) != 0:
        raise SystemExit("bench.py: init failed")
    def attach(obj):
        if world > 1:
            if os.environ.get("MLF_BENCH_ALLREDUCE", "nccl") == "callback":   # A/B: Python callback over torch.distributed
                obj.set_allreduce(parallel.make_allreduce(None, dev), rank, world)
            else:                                                           # native: ncclAllReduce on the engine stream, no Python per iteration
                parallel.attach_nccl(obj, rank, world, device=dev)

    attach(em)
    em.inject(Mu0, Cov0, lam0)
    info, _ = em.step(args.warmup)
    if info < 0:
        raise SystemExit(f"bench.py: warm-up failed: {em.last_error()}")

    est = torch.cuda.ExternalStream(em.stream(), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    em.set_profile(True)
    l0 = em.launch_count()
    barrier()
    tw0 = time.perf_counter()
    ev0.record(est)
    info, _ = em.step(args.steps)
    ev1.record(est)
    torch.cuda.synchronize(dev)
    tw1 = time.perf_counter()
    barrier()
    if info < 0:
        raise SystemExit(f"bench.py: timed step failed: {em.last_error()}")
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(tw0, tw1) if sampler else None
    launches = em.launch_count() - l0
    times = em.times()
    counters = em.counters()
    qf = em.qf_stats()
    plan = em.kernel_plan()
    em.set_profile(False)
    sumll = em.sumLL
    value = N * K * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (E-pass), CUDA events on the engine's stream inside the timed region ------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
    except Exception:
        pass
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    hbm_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback 6650 GB/s (B200_PROFILING.md)"
    hbm_peak = hbm_peak or 6650.0
    it = max(times["iters"], 1)
    fused = times.get("sweeps", 0) > 0
    sweeps = max(times.get("sweeps", 0), 1) if fused else it
    # ev1..ev2 brackets the sweep(s) + the statistics all-reduce; with the fused path it holds `sweeps` launches of k_em
    e_ms, m_ms = times["estep_ms"] / sweeps, times["mstep_ms"] / it
    peak_tf = peaks.get("dmma_tflops", 37.16)
    # algorithmic flops per (point, component): E = d^2+4d+6 (+40 for the exp); the fused sweep also does the dense
    # part of M (sum g, sum g^2, sum g x: 3d+4).  The d^2 scatter term is only owed for pairs that pass the
    # reference's W >= minW test, a data-dependent handful per point, and is left out (conservative).
    fl_pair = flops_e(d) + ((3 * d + 4) if fused else 0)
    diag_cc = "k_emd" in plan
    if diag_cc:
        # diagonal-covariance extension on the fp64 CUDA cores: the whitening is 2 multiply-adds per dimension (4d flops)
        # instead of a d x d contraction, and the bound is the DFMA rate, not the tensor datapath
        fl_pair = (4 * d + 6 + 40) + (3 * d + 4)
        peak_tf = peaks.get("dfma_tflops", 34.16)
    ach_tf = fl_pair * n_loc * K / (e_ms * 1e-3) * 1e-12 if e_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "estep_traffic.json")))
        traffic = tj["dram_bytes_per_point"] * n_loc if (tj.get("d") == d and tj.get("K") == K) else None
    except Exception:
        pass
    roofline = {
        "kernel": plan + (" (diagonal whitening DFMA + log-sum-exp + pass-A statistics + thresholded variance scatter)" if diag_cc else
                          " (logits DMMA + log-sum-exp + pass-A statistics + thresholded covariance scatter)" if fused
                          else " (roofline line = E-pass: whitening DMMA + log-sum-exp + pass-A statistics)"), "bound": "tensor",
        "sweeps_per_iteration": sweeps / it if fused else 1.0, "side_list_pairs": times.get("side_list_pairs", 0),
        "kept_pairs_per_point": (counters["kept_pairs"] / it / max(n_loc * world, 1)) if fused else None,
        # how the sweep formed the log-densities: expanded quadratic form (guarded by the magnitude bound of its terms) or whitening
        "logits": {"quadratic_form_sweeps": qf["qf_sweeps"], "whitening_sweeps": int(sweeps) - qf["qf_sweeps"] if fused else None,
                   "scatter_passes_from_stored_responsibilities": qf["scatter_passes"], "term_magnitude_bound": qf["bound"], "guard": "4 u S_k <= 5e-11 (MLF_B200_QF_TOL)"} if ("k_em16" in plan or "k_emd" in plan) else None,
        "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None, "traffic": traffic,
        "peak_source": ("fp64 DFMA on the CUDA cores (same fp64 datapath as DMMA; 'tensor' is kept as the bound label of the contract)"
                        if diag_cc else "fp64 DMMA (mma.sync m8n8k4.f64)") +
                       " measured on this pool by bench/micro/fp64_peaks.cu -> profiles/fp64_peaks_r01.json; "
                       "MEASURED_PEAKS.json has no fp64 figure (bf16 tensor peak does not apply to an fp64 path)",
        "algorithmic_flops_per_launch": fl_pair * n_loc * K, "flops_per_pair": fl_pair, "ms_per_launch": e_ms,
        "hbm": {"achieved_gbs": (8.0 * n_loc * d + (0.0 if fused else 8.0 * n_loc * K)) / (e_ms * 1e-3) * 1e-9 if e_ms > 0 else 0.0, "peak_gbs": hbm_peak,
                "peak_source": hbm_src, "note": ("fused sweep reads X once and writes nothing per point" if fused else
                                                 "E-pass reads X once and writes the responsibilities once") + "; fp64-bound, not HBM-bound"},
        "mstep": (None if fused else {"ms_per_launch": m_ms, "algorithmic_bytes": 8.0 * n_loc * K + 8.0 * n_loc * d,
                                      "achieved_gbs": (8.0 * n_loc * K + 8.0 * n_loc * d) / (m_ms * 1e-3) * 1e-9 if m_ms > 0 else 0.0}),
        "per_iteration_ms": {k: v / it for k, v in times.items() if k != "iters"},
    }

    # ---- e2e: same C-ABI, HOST buffers; H2D of X (pinned) + mlf_step(1) + D2H of parameters every step --------------
    e2e = None
    if not args.no_e2e:
        e_steps = args.e2e_steps or min(args.steps, 5)
        wc_ptr = None
        if args.host_buffer == "wc":
            import ctypes
            wc_ptr = em._lib.mlf_b200_host_alloc(max(8 * n_loc * d, 8), 1)
            if not wc_ptr:
                raise SystemExit("bench.py: mlf_b200_host_alloc failed")
            Xn = np.ctypeslib.as_array((ctypes.c_double * (n_loc * d)).from_address(wc_ptr)).reshape(n_loc, d)
            Xh = torch.from_numpy(Xn)
        else:
            Xh = torch.empty((n_loc, d), dtype=torch.float64, pin_memory=True)
            Xn = Xh.numpy()
        Xh.copy_(X)
        torch.cuda.synchronize(dev)
        # ceiling of the upload: the rank's whole shard as ONE plain cudaMemcpyAsync from the same host buffer, all ranks at once
        probe_ms = []
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            barrier()
            pe0.record()
            X.copy_(Xh, non_blocking=True)
            pe1.record()
            torch.cuda.synchronize(dev)
            probe_ms.append(max_over_ranks(pe0.elapsed_time(pe1)))
        probe_gbs = 8.0 * n_loc * d / (min(probe_ms) * 1e-3) * 1e-9 if n_loc else 0.0
        em.finalize()
        del X
        torch.cuda.empty_cache()
        em = mlf_algo_emgmm()
        if em.init(Xn, K, args.min_proba, device=local_rank, cov_type=args.cov_type) != 0:
            raise SystemExit("bench.py: e2e init failed")
        attach(em)
        em.inject(Mu0, Cov0, lam0)
        em.step(1)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            if em._lib.mlf_b200_refresh_x(em._h, Xn.ctypes.data) != 0:
                raise SystemExit("bench.py: refresh_x failed")
            info, _ = em.step(1)
            if info < 0:
                raise SystemExit(f"bench.py: e2e step failed: {em.last_error()}")
        torch.cuda.synchronize(dev)
        el = max_over_ranks(time.perf_counter() - t0)
        barrier()
        pbytes = 8 * (K * d + K * d * d + K)
        e2e = {"value": N * K * e_steps / el, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_loc * d + pbytes), "d2h_bytes_per_step": int(pbytes + 8),
               "steps": e_steps, "ms_per_step": el / e_steps * 1e3, "em_iter_per_s": e_steps / el, "host_affinity": numa_note,
               "note": "per rank and step: mlf_b200_refresh_x(pinned host X) + mlf_step(1) + parameters/sumLL back; the X upload is "
                       "chunked underneath the step's sweep, so the step is PCIe-bound (h2d_bytes / ms_per_step = achieved H2D GB/s)",
               "h2d_gbs": (8 * n_loc * d + pbytes) / (el / e_steps) * 1e-9,
               "host_buffer": "page-locked" + (", write-combined" if args.host_buffer == "wc" else ""),
               "h2d_probe_gbs": probe_gbs, "h2d_probe_note": "per rank: the shard as one plain cudaMemcpyAsync from the same host buffer with every rank "
                                                              "copying at once (slowest rank, best of 3) -- the ceiling the streamed step is held against"}
    em.finalize()
    if e2e is not None and args.host_buffer == "wc":
        del Xh
        if not (rank == 0 and world == 1 and not args.no_cpu):
            del Xn
            em._lib.mlf_b200_host_free(wc_ptr)

    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, full_affinity)
        if args.cov_type == "full":
            parity = parity_spot_check(X if args.no_e2e else Xn, d, K, Mu0, Cov0, lam0, args.min_proba)   # the device copy was released for the e2e leg
        r = cpu_oracle_rate(d, K, args.cpu_points, args.min_proba, 3, 1, sigma_c=args.sigma_c, sigma_x=args.sigma_x)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"oracle (C restatement of the reference, OpenMP over components) on N_cpu={r['n']} points of the same d={d}, K={K} mixture, "
                         f"3 EM iterations after 1 warm-up; {r['sec_per_step']:.2f} s/iteration"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "em_iter_per_s": args.steps / (ms * 1e-3), "sumLL_last": sumll,
            "config": {"workload": f"synthetic GMM N={N:.0e} d={d} K={K} {'full' if args.cov_type == 'full' else 'diagonal'} covariance, points sharded over {world} GPU(s)"
                                   + (" (BASELINE configs[2])" if (d, K, args.cov_type) == (16, 64, "full") else ""),
                       "N": N, "d": d, "K": K, "minProba": args.min_proba, "sigmaC": args.sigma_c, "sigmaX": args.sigma_x, "points_per_gpu": n_loc, "parallelism": f"points-sharded x{world}",
                       "l2": "inputs larger than L2 (X is 12.8 GB per 1e8 points >> 126 MB); no explicit flush",
                       "timing": "CUDA events on the engine stream, max over ranks"},
            "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
