#!/usr/bin/env python
"""bench.py -- batched FP64 LU on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): dgesv_batched n=16, nrhs=1, batch=1,000,000 per GPU
(the batch is sharded by matrix index, every rank owns 1e6 matrices: weak scaling, no collective
on the data path). A step = one magma_dgesv_batched call over one resident batch. K distinct
dlarnv batches are generated in HBM before the timed region (2.2 GB each, > L2), so no step
re-reads cached data and no step works on already-factored input.

One JSON line on stdout (rank 0). Extra keys beyond the contract: `sweep` (the other BASELINE
configs, each with GFLOP/s and its roofline fraction), `peaks`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched dgetrf/dgesv GFLOP/s (FLOPS_DGETRF + FLOPS_DGETRS, testing/flops.h)"
UNIT = "GFLOP/s"
N_HEAD, NRHS_HEAD, BATCH_HEAD = 16, 1, 1_000_000
KERNEL_HEAD = "lu_sqs_kernel<16,8,1>"  # the one kernel a headline step launches (magma_b200/csrc/lu_small_sq.cu)


def advance_seed(iseed, ndraws):
    """LAPACK dlarnv seed after ndraws draws (x <- a x mod 2^48, a = 33952834046453; base-4096 digits)."""
    import numpy as np
    x = (int(iseed[0]) << 36) | (int(iseed[1]) << 24) | (int(iseed[2]) << 12) | int(iseed[3])
    x = (x * pow(33952834046453, int(ndraws), 1 << 48)) & ((1 << 48) - 1)
    return np.array([(x >> 36) & 4095, (x >> 24) & 4095, (x >> 12) & 4095, x & 4095], dtype=np.int32)


def workload_str(batch):
    """config.workload: the SAME string in both arms (the driver compares them)."""
    return f"dgesv_batched n={N_HEAD} nrhs={NRHS_HEAD} batch={batch} per GPU (BASELINE configs[1])"


def flops_getrf(m, n):
    if m < n:
        mul = 0.5 * m * (m * (n - m / 3.0 - 1.0) + n) + 2.0 * m / 3.0
        add = 0.5 * m * (m * (n - m / 3.0) - n) + m / 6.0
    else:
        mul = 0.5 * n * (n * (m - n / 3.0 - 1.0) + m) + 2.0 * n / 3.0
        add = 0.5 * n * (n * (m - n / 3.0) - m) + n / 6.0
    return mul + add


def flops_getrs(n, nrhs):
    return nrhs * (2.0 * n * n - n)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks: poll NVML from a thread during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nv is None:
            return
        self._stop.clear()
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        import statistics
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference testers' LAPACK + OpenMP loop (oracle/lapack_loop.c)
# ------------------------------------------------------------------------------------------------
def cpu_lapack_gesv(n, nrhs, batch_total, budget_s=12.0):
    """Times dgesv_ over a bounded sample of the workload. Returns (gflops, cores, sample, seconds)."""
    import numpy as np

    import oracle
    L = oracle.lapack()
    try:  # all the host threads this process may use (torchrun pins OMP_NUM_THREADS=1 in its workers)
        L.lapack_loop_set_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    cores = L.lapack_loop_threads()
    fl = flops_getrf(n, n) + flops_getrs(n, nrhs)

    def run(cnt):
        seed = np.array([0, 0, 0, 1], dtype=np.int32)
        A = np.empty(cnt * n * n)
        B = np.empty(cnt * n * nrhs)
        L.lapack_dlarnv(1, seed, A.size, A)
        L.lapack_dlarnv(1, seed, B.size, B)
        ipiv = np.zeros(cnt * n, dtype=np.int32)
        info = np.zeros(cnt, dtype=np.int32)
        return L.lapack_dgesv_loop(n, nrhs, A, n, n * n, ipiv, n, B, n, n * nrhs, info, cnt)

    probe = min(batch_total, 50_000)
    run(min(probe, 5000))  # warm the library / threads
    t = run(probe)
    cnt = int(min(batch_total, max(probe, probe * budget_s / max(t, 1e-6) / 3)))
    t = min(run(cnt), run(cnt))
    return fl * cnt / t / 1e9, cores, cnt, t, L.lapack_loop_describe().decode()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    vals, secs = [], []
    cores = sample = desc = None
    for i in range(args.warmup + args.steps):
        g, cores, sample, t, desc = cpu_lapack_gesv(N_HEAD, NRHS_HEAD, args.batch, budget_s=4.0)
        if i >= args.warmup:
            vals.append(g)
            secs.append(t)
    import statistics
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_str(args.batch),
                   "inputs": "dlarnv(1,{0,0,0,1})", "step": f"LAPACK dgesv loop over a {sample}-matrix sample"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{sample} of {args.batch} matrices per step; omp parallel for schedule(dynamic) "
                                   f"over dgesv_, BLAS threads = 1 ({desc})"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref", action="store_true", help="skip the same-box reference-GPU / cuBLAS rows")
    ap.add_argument("--batch", type=int, default=BATCH_HEAD)
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    from magma_b200 import batched as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the batched LU path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert mb.magma_init() == 0
    q = mb.Queue.from_torch(local)
    stream = torch.cuda.current_stream(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    hbm_peak, peak_src = measured_peaks()
    clocks = ClockSampler(local)

    # ---- headline: gesv n=16 nrhs=1, `batch` matrices per rank --------------------------------
    n, nrhs, batch = N_HEAD, NRHS_HEAD, args.batch
    K, W = args.steps, args.warmup
    nbuf = min(K, 32)
    bufs = [mb.DeviceBatch(batch, n, n, nrhs=nrhs, device=local, queue=q) for _ in range(nbuf)]
    # rank r owns matrices [r*batch, (r+1)*batch) of the global dlarnv stream of each step's batch
    def fill(i):
        seed = np.array([(i * 7 + 3) % 4096, (rank * 13 + 1) % 4096, 0, 1], dtype=np.int32)
        mb.dlarnv_uniform(seed, batch * n * n, bufs[i].A, q)
        mb.dlarnv_uniform(seed, batch * n * nrhs, bufs[i].B, q)

    for i in range(nbuf):
        fill(i)
    for i in range(W):
        assert bufs[i % nbuf].gesv() == 0
    for i in range(min(W, nbuf)):
        fill(i)
    barrier()
    launches0 = mb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    clocks.start()
    barrier()
    ev[0].record(stream)
    for i in range(K):
        rc = bufs[i % nbuf].gesv()
        ev[i + 1].record(stream)
    barrier()
    clocks.stop()
    assert rc == 0
    gpu_launches = mb.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    kern_ms = max_over_ranks(sum(step_ms) / K)  # one launch per step: launch duration == step time
    info_bad = int(bufs[0].info.abs().max().item())
    assert info_bad == 0, "singular matrix in the synthetic batch?"
    fl_mat = flops_getrf(n, n) + flops_getrs(n, nrhs)
    by_mat = 2 * 8 * n * n + 2 * 8 * n * nrhs  # read A,B + write LU,X (pivots/info excluded)
    value = fl_mat * batch * world * K / (total_ms * 1e-3) / 1e9
    achieved = by_mat * batch / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": KERNEL_HEAD, "peak_source": peak_src,
                "alg_bytes_per_launch": by_mat * batch, "launch_ms": kern_ms}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(KERNEL_HEAD)
        except Exception:
            pass
    del bufs
    torch.cuda.empty_cache()

    # ---- e2e: host buffers through the C ABI (pinned), copies inside the timed region ---------
    e2e_steps = max(1, min(K, 5))
    # pinned staging on the NUMA node next to this rank's GPU: first-touch placement follows the CPUs the thread may
    # run on, so the rank pins itself to the GPU's local CPU set (NVML) while it allocates and while it drives copies
    numa = {"cpus_bound": None}
    affinity0 = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa["cpus_bound"] = len(cpus)
    except Exception as ex:  # noqa: BLE001
        numa["error"] = str(ex)[:120]
    hA = torch.empty((batch, n, n), dtype=torch.float64).pin_memory()
    hB = torch.empty((batch, nrhs, n), dtype=torch.float64).pin_memory()
    hip = torch.empty((batch, n), dtype=torch.int32).pin_memory()
    hinfo = torch.empty((batch,), dtype=torch.int32).pin_memory()
    src = mb.DeviceBatch(batch, n, n, nrhs=nrhs, device=local, queue=q)
    hA0 = torch.empty_like(hA).pin_memory()
    hB0 = torch.empty_like(hB).pin_memory()
    seed = np.array([5, rank, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, batch * n * n, src.A, q)
    mb.dlarnv_uniform(seed, batch * n * nrhs, src.B, q)
    q.sync()
    hA0.copy_(src.A)
    hB0.copy_(src.B)
    del src
    torch.cuda.empty_cache()
    e2e_t = []
    for i in range(1 + e2e_steps):
        hA.copy_(hA0)
        hB.copy_(hB0)
        barrier()
        if i == 1:
            clocks.start()
        t0 = time.perf_counter()
        rc = mb.dgesv_batched_host(n, nrhs, hA, n, hip, hB, n, hinfo, batch, q)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert rc == 0
        if i >= 1:
            e2e_t.append(max_over_ranks(dt))
    clocks.stop()
    assert int(hinfo.abs().max()) == 0
    # the box's host<->device ceiling for the same bytes: H2D and D2H of one step's payload on two streams, no compute
    # (what the host side -- pinned memory bandwidth and the PCIe root complexes the ranks share -- allows at this N)
    dscr = torch.empty((batch, n + nrhs, n), dtype=torch.float64, device=dev)
    s_in, s_out = torch.cuda.Stream(local), torch.cuda.Stream(local)
    copy_t = []
    for i in range(3):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            dscr[:, :n].copy_(hA, non_blocking=True)
            dscr[:, n:].copy_(hB, non_blocking=True)
        with torch.cuda.stream(s_out):
            hA.copy_(dscr[:, :n], non_blocking=True)
            hB.copy_(dscr[:, n:], non_blocking=True)
        torch.cuda.synchronize()
        copy_t.append(max_over_ranks(time.perf_counter() - t0))
    del dscr
    e2e_val = fl_mat * batch * world / (sum(e2e_t) / len(e2e_t)) / 1e9
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": (n * n + n * nrhs) * 8 * batch,
           "d2h_bytes_per_step": ((n * n + n * nrhs) * 8 + n * 4 + 4) * batch,
           "ms_per_step": 1e3 * sum(e2e_t) / len(e2e_t),
           "pcie_GBs_per_rank": ((n * n + n * nrhs) * 16 + n * 4 + 4) * batch / (sum(e2e_t) / len(e2e_t)) / 1e9,
           "copy_only_ms": 1e3 * min(copy_t), "numa": numa,
           "copy_only_note": "same bytes both ways on two streams with no kernel: the host-side ceiling at this rank count",
           "call": "magma_b200_dgesv_batched_host (pinned host A,B in; LU,X,ipiv,info out)"}
    del hA, hB, hA0, hB0, hip, hinfo
    os.sched_setaffinity(0, affinity0)  # the CPU baseline below may use every core again

    # ---- peaks for the compute-bound rows ------------------------------------------------------
    peaks = {"hbm_gbs": hbm_peak, "hbm_source": peak_src,
             "fp64_dfma_tflops": mb.fp64_peak_tflops(0, q), "fp64_dmma_tflops": mb.fp64_peak_tflops(1, q),
             "hbm_copy_gbs_live": mb.hbm_copy_gbs(1 << 30, q)}
    fp64_peak = max(peaks["fp64_dfma_tflops"], peaks["fp64_dmma_tflops"]) * 1e3  # GFLOP/s

    # ---- sweep over the other BASELINE configs ---------------------------------------------------
    sweep = []
    if not args.no_sweep:
        sweep = run_sweep(mb, torch, np, q, local, rank, world, barrier, max_over_ranks, sum_over_ranks, hbm_peak,
                          fp64_peak)

    strong = None
    if world > 1 and not args.no_sweep:
        strong = run_strong(mb, torch, np, q, local, rank, world, barrier, max_over_ranks, dist, dev)
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_sweep and not args.no_ref:
        ref_gpu = run_ref_gpu(sweep, e2e=None, head_ms=total_ms / K)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        g, cores, cnt, t, desc = cpu_lapack_gesv(n, nrhs, batch)
        cpu_baseline = {"value": g, "unit": UNIT, "cores": cores, "kind": "reference",
                        "sample": f"{cnt} of {batch} matrices, {t:.2f} s; omp parallel for schedule(dynamic) over "
                                  f"dgesv_, BLAS threads = 1 ({desc})"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_str(batch),
                       "inputs": f"dlarnv U(0,1) generated in HBM; {nbuf} distinct batches of "
                                 f"{(by_mat // 2) * batch / 1e9:.2f} GB (> L2) rotated, ldda=lddb=n",
                       "parallelism": f"batch sharded by matrix index over {world} GPU(s), no collective",
                       "l2": "inputs larger than L2"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "clocks": clocks.summary(), "peaks": peaks, "metric_sizes": metric_sizes(sweep), "sweep": sweep, "strong": strong,
            "ref_gpu": ref_gpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def metric_sizes(sweep):
    """BASELINE.json's metric names dgetrf at n = 32 / 128 / 512: their rows of the sweep, pulled up beside the headline
    (which is BASELINE configs[1], gesv n = 16) so that no reader has to dig for them."""
    out = {}
    for key, tag in (("n32_batch1e4", "C1 "), ("n32_batch1e6", "C1b "), ("n128_batch5e4", "C3 "), ("n512_batch4e3", "C5 ")):
        for r in sweep:
            if r["config"].startswith(tag):
                out[key] = {"ms": r["ms"], "gflops_per_gpu": r.get("gflops_per_gpu"), "roofline_gflops": r.get("roofline_gflops"),
                            "frac_of_roofline": r.get("frac_of_roofline"), "bound": r.get("bound", "min(hbm, fp64)")}
    return out


def run_sweep(mb, torch, np, q, local, rank, world, barrier, max_over_ranks, sum_over_ranks, hbm_peak, fp64_peak):
    """The other BASELINE configs; per-rank shard = the config's batch (weak scaling)."""
    out = []
    stream = torch.cuda.current_stream(local)

    def timed(fn, restore, reps):
        ts = []
        for _ in range(reps):
            restore()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn()
            e1.record(stream)
            barrier()
            assert rc == 0
            ts.append(max_over_ranks(e0.elapsed_time(e1)))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def fixed(name, n, batch, nrhs=0, solve_after=False, reps=5):
        db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, device=local, queue=q)
        seed = np.array([11, rank, 0, 1], dtype=np.int32)
        mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
        if nrhs:
            mb.dlarnv_uniform(seed, batch * n * nrhs, db.B, q)
        q.sync()
        A0 = db.A.clone()
        B0 = db.B.clone() if nrhs else None

        def restore():
            db.A.copy_(A0)
            if nrhs:
                db.B.copy_(B0)

        l0 = mb.launch_count()
        med, best = timed(db.getrf, restore, reps)
        launches = (mb.launch_count() - l0) // reps
        fl = flops_getrf(n, n) * batch * world
        gf = fl / (med * 1e-3) / 1e9
        hbm_roof = flops_getrf(n, n) / (16.0 * n * n) * hbm_peak  # GFLOP/s at HBM speed
        roof = min(hbm_roof, fp64_peak)
        row = {"config": name, "n": n, "batch_per_gpu": batch, "ms": med, "ms_best": best, "gflops": gf,
               "gflops_per_gpu": gf / world, "roofline_gflops": roof, "bound": "hbm" if hbm_roof < fp64_peak else "fp64",
               "frac_of_roofline": gf / world / roof, "alg_GBs_per_gpu": 16.0 * n * n * batch / (med * 1e-3) / 1e9,
               "launches_per_call": int(launches), "info_max": int(db.info.abs().max().item())}
        if solve_after:
            # getrs on the factors just computed
            def restore_b():
                db.B.copy_(B0)
            med2, best2 = timed(db.getrs, restore_b, reps)
            fl2 = flops_getrs(n, nrhs) * batch * world
            by2 = (8.0 * n * n + 16.0 * n * nrhs)
            roof2 = min(flops_getrs(n, nrhs) / by2 * hbm_peak, fp64_peak)
            row["getrs"] = {"nrhs": nrhs, "ms": med2, "gflops": fl2 / (med2 * 1e-3) / 1e9,
                            "frac_of_roofline": fl2 / world / (med2 * 1e-3) / 1e9 / roof2,
                            "alg_GBs_per_gpu": by2 * batch / (med2 * 1e-3) / 1e9}
        out.append(row)
        del db, A0, B0
        torch.cuda.empty_cache()

    # C1 is an 80 us call: 25 timed calls over four rotated 164 MB batches (each restored outside the timed region;
    # together 650 MB > L2), median -- one launch per call, so launch latency is part of the number, as for a user
    def c1(name, n, batch, nbuf=4, reps=25):
        dbs = [mb.DeviceBatch(batch, n, n, device=local, queue=q) for _ in range(nbuf)]
        A0s = []
        for i, db in enumerate(dbs):
            seed = np.array([11 + i, rank, 0, 1], dtype=np.int32)
            mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
        q.sync()
        A0s = [db.A.clone() for db in dbs]
        ts = []
        for r in range(reps + 3):
            i = r % nbuf
            dbs[i].A.copy_(A0s[i])
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = dbs[i].getrf()
            e1.record(stream)
            barrier()
            assert rc == 0
            if r >= 3:
                ts.append(max_over_ranks(e0.elapsed_time(e1)))
        ts.sort()
        med, best = ts[len(ts) // 2], ts[0]
        gf = flops_getrf(n, n) * batch * world / (med * 1e-3) / 1e9
        roof = min(flops_getrf(n, n) / (16.0 * n * n) * hbm_peak, fp64_peak)
        out.append({"config": name, "n": n, "batch_per_gpu": batch, "ms": med, "ms_best": best, "gflops": gf,
                    "gflops_per_gpu": gf / world, "roofline_gflops": roof, "bound": "hbm", "frac_of_roofline": gf / world / roof,
                    "alg_GBs_per_gpu": 16.0 * n * n * batch / (med * 1e-3) / 1e9, "launches_per_call": 1,
                    "timing": f"median of {reps} calls over {nbuf} rotated batches", "info_max": int(dbs[0].info.abs().max().item())})
        del dbs, A0s
        torch.cuda.empty_cache()

    c1("C1 dgetrf_batched n=32 batch=10000", 32, 10_000)
    fixed("C1b dgetrf_batched n=32 batch=1000000", 32, 1_000_000)
    fixed("C3 dgetrf_batched n=128 batch=50000", 128, 50_000)
    fixed("C5 dgetrf_batched n=512 batch=4000 (+dgetrs nrhs=16)", 512, 4_000, nrhs=16, solve_after=True, reps=3)
    fixed("X1 dgetrf_batched n=256 batch=16000 (left-looking slab driver)", 256, 16_000, reps=3)

    # F1: out-of-place inverse from the factors, n = 64: one launch for n <= 64 (csrc/getri.cu); identity fill + the getrs path above that.
    # FLOPs: testing/flops.h:87-88,279 (FLOPS_DGETRI); bytes: read LU, write inv(A).
    def getri(name, n, batch, reps=5):
        db = mb.DeviceBatch(batch, n, n, nrhs=n, device=local, queue=q)
        seed = np.array([31, rank, 0, 1], dtype=np.int32)
        mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
        q.sync()
        assert db.getrf() == 0
        fn = lambda: mb.magma_dgetri_outofplace_batched(n, db.dA_array, db.ldda, db.dipiv_array, db.dB_array,  # noqa: E731
                                                        db.lddb, db.info, batch, q)
        med, best = timed(fn, lambda: None, reps)
        fl = n * ((5. / 6.) + n * ((2. / 3.) * n + 0.5)) + n * ((5. / 6.) + n * ((2. / 3.) * n - 1.5))
        gf = fl * batch * world / (med * 1e-3) / 1e9
        roof = min(fl / (16.0 * n * n) * hbm_peak, fp64_peak)
        out.append({"config": name, "n": n, "batch_per_gpu": batch, "ms": med, "ms_best": best, "gflops": gf,
                    "gflops_per_gpu": gf / world, "roofline_gflops": roof, "frac_of_roofline": gf / world / roof,
                    "alg_GBs_per_gpu": 16.0 * n * n * batch / (med * 1e-3) / 1e9})
        del db
        torch.cuda.empty_cache()

    getri("F1 dgetri_outofplace_batched n=64 batch=100000", 64, 100_000)
    getri("F1b dgetri_outofplace_batched n=32 batch=400000", 32, 400_000)

    # F2: LU without pivoting, n = 128 (diagonally dominant inputs: dlarnv + n on the diagonal)
    def nopiv(name, n, batch, reps=5):
        db = mb.DeviceBatch(batch, n, n, device=local, queue=q)
        seed = np.array([41, rank, 0, 1], dtype=np.int32)
        mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
        q.sync()
        db.A.diagonal(dim1=1, dim2=2).add_(float(n))
        A0 = db.A.clone()
        fn = lambda: mb.magma_dgetrf_nopiv_batched(n, n, db.dA_array, db.ldda, db.info, batch, q)  # noqa: E731
        med, best = timed(fn, lambda: db.A.copy_(A0), reps)
        gf = flops_getrf(n, n) * batch * world / (med * 1e-3) / 1e9
        roof = min(flops_getrf(n, n) / (16.0 * n * n) * hbm_peak, fp64_peak)
        out.append({"config": name, "n": n, "batch_per_gpu": batch, "ms": med, "ms_best": best, "gflops": gf,
                    "gflops_per_gpu": gf / world, "roofline_gflops": roof, "frac_of_roofline": gf / world / roof,
                    "alg_GBs_per_gpu": 16.0 * n * n * batch / (med * 1e-3) / 1e9,
                    "info_max": int(db.info.abs().max().item())})
        del db, A0
        torch.cuda.empty_cache()

    nopiv("F2 dgetrf_nopiv_batched n=128 batch=50000", 128, 50_000)

    # G1-G3: the standalone batched BLAS-3 behind the reference names and the butterfly solver.
    # flops: 2mnk (gemm), m^2 n (trsm, testing/flops.h FLOPS_DTRSM left); bytes: every operand once in, the result once out.
    def blas3(reps=5):
        dev = torch.device("cuda", local)
        g = torch.Generator(device=dev)
        g.manual_seed(61 + rank)
        idx = lambda cnt: torch.arange(cnt, dtype=torch.int64, device=dev)  # noqa: E731
        # G1: C <- C - A B (the LU update's alpha = -1, beta = 1), 128 x 128 x 128
        n, batch = 128, 20_000
        A, B, Cm = (torch.rand((batch, n, n), dtype=torch.float64, device=dev, generator=g) for _ in range(3))
        pa, pb, pc = (idx(batch) * (n * n * 8) + t.data_ptr() for t in (A, B, Cm))
        fn = lambda: (mb.magma_dgemm_batched_core(111, 111, n, n, n, -1.0, pa, 0, 0, n, pb, 0, 0, n, 1.0, pc, 0, 0, n, batch, q), 0)[1]  # noqa: E731
        med, best = timed(fn, lambda: None, reps)
        fl = 2.0 * n * n * n
        roof = min(fl / (4 * 8.0 * n * n) * hbm_peak, fp64_peak)
        out.append({"config": "G1 dgemm_batched NN n=128 batch=20000 (alpha=-1, beta=1)", "n": n, "batch_per_gpu": batch, "ms": med,
                    "ms_best": best, "gflops": fl * batch * world / (med * 1e-3) / 1e9, "gflops_per_gpu": fl * batch / (med * 1e-3) / 1e9,
                    "roofline_gflops": roof, "frac_of_roofline": fl * batch / (med * 1e-3) / 1e9 / roof})
        # G2: B <- L^-1 B, unit lower, left, 128 x 128 triangle, 64 right-hand sides
        m, nr = 128, 64
        Bm = torch.rand((batch, nr, m), dtype=torch.float64, device=dev, generator=g)
        B0 = Bm.clone()
        pb2 = idx(batch) * (nr * m * 8) + Bm.data_ptr()
        fn = lambda: (mb.magmablas_dtrsm_batched(141, 122, 111, 132, m, nr, 1.0, pa, n, pb2, m, batch, q), 0)[1]  # noqa: E731
        med, best = timed(fn, lambda: Bm.copy_(B0), reps)
        fl = float(m) * m * nr
        roof = min(fl / (8.0 * (m * m / 2 + 2 * m * nr)) * hbm_peak, fp64_peak)
        out.append({"config": "G2 dtrsm_batched L,Lower,NoTrans,Unit m=128 n=64 batch=20000", "n": m, "batch_per_gpu": batch, "ms": med,
                    "ms_best": best, "gflops": fl * batch * world / (med * 1e-3) / 1e9, "gflops_per_gpu": fl * batch / (med * 1e-3) / 1e9,
                    "roofline_gflops": roof, "frac_of_roofline": fl * batch / (med * 1e-3) / 1e9 / roof})
        del A, B, Cm, Bm, B0
        torch.cuda.empty_cache()

    blas3()

    # P1-P3: the s / c / z entry points (csrc/lu_scz.cu: coverage kernels, not tuned to the roofline).
    # bytes: read + write the matrix in its own element size; flops: real = FLOPS_DGETRF, complex = 4x (testing/flops.h:24-27)
    def prec(name, p, n, batch, reps=5):
        from magma_b200 import _lib
        dev = torch.device("cuda", local)
        rdt, dt = {"s": (torch.float32, torch.float32), "c": (torch.float32, torch.complex64),
                   "z": (torch.float64, torch.complex128)}[p]
        g = torch.Generator(device=dev)
        g.manual_seed(51 + rank)
        A = torch.rand((batch, n, n), dtype=rdt, device=dev, generator=g)
        if p != "s":
            A = torch.complex(A, torch.rand((batch, n, n), dtype=rdt, device=dev, generator=g))
        A0 = A.clone()
        esz = A.element_size()
        idx = torch.arange(batch, dtype=torch.int64, device=dev)
        ip = torch.zeros((batch, n), dtype=torch.int32, device=dev)
        info = torch.zeros(batch, dtype=torch.int32, device=dev)
        pA, pP = idx * (n * n * esz) + A.data_ptr(), idx * (n * 4) + ip.data_ptr()
        f = getattr(_lib.load(), f"magma_{p}getrf_batched")
        fn = lambda: f(n, n, pA.data_ptr(), n, pP.data_ptr(), info.data_ptr(), batch, q.handle)  # noqa: E731
        med, best = timed(fn, lambda: A.copy_(A0), reps)
        fl = flops_getrf(n, n) * (1.0 if p == "s" else 4.0)
        out.append({"config": name, "n": n, "batch_per_gpu": batch, "ms": med, "ms_best": best,
                    "gflops": fl * batch * world / (med * 1e-3) / 1e9, "dtype": {"s": "f32", "c": "c64", "z": "c128"}[p],
                    "alg_GBs_per_gpu": 2.0 * esz * n * n * batch / (med * 1e-3) / 1e9,
                    "frac_of_hbm": 2.0 * esz * n * n * batch / (med * 1e-3) / 1e9 / hbm_peak,
                    "info_max": int(info.abs().max().item())})
        del A, A0, ip, info
        torch.cuda.empty_cache()

    prec("P1 sgetrf_batched n=32 batch=1000000", "s", 32, 1_000_000)
    prec("P2 zgetrf_batched n=32 batch=250000", "z", 32, 250_000)
    prec("P3 cgetrf_batched n=64 batch=50000", "c", 64, 50_000, reps=3)

    # C4: vbatched, sizes 16 + (lcg mod 497), square, ldda = n
    batch = 20_000
    x = 1234 + rank
    ns = []
    for _ in range(batch):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        ns.append(16 + (x >> 8) % 497)
    ns = np.array(ns, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(ns * ns)])
    poffs = np.concatenate([[0], np.cumsum(ns)])
    dev = torch.device("cuda", local)
    dA = torch.empty(int(offs[-1]), dtype=torch.float64, device=dev)
    seed = np.array([21, rank, 0, 1], dtype=np.int32)
    mb.dlarnv_uniform(seed, int(offs[-1]), dA, q)
    q.sync()
    A0 = dA.clone()
    dip = torch.zeros(int(poffs[-1]), dtype=torch.int32, device=dev)
    dinfo = torch.zeros(batch, dtype=torch.int32, device=dev)
    pA = (torch.from_numpy(offs[:-1] * 8).to(dev) + dA.data_ptr())
    pP = (torch.from_numpy(poffs[:-1] * 4).to(dev) + dip.data_ptr())
    dn = torch.from_numpy(ns.astype(np.int32)).to(dev)
    fn = lambda: mb.magma_dgetrf_vbatched(dn, dn, pA, dn, pP, dinfo, batch, q)  # noqa: E731
    med, best = timed(fn, lambda: dA.copy_(A0), 3)
    fl = float(sum(flops_getrf(int(k), int(k)) for k in ns))
    fl_all = sum_over_ranks(fl)
    out.append({"config": "C4 dgetrf_vbatched n~U[16,512] batch=20000", "batch_per_gpu": batch, "ms": med,
                "ms_best": best, "gflops": fl_all / (med * 1e-3) / 1e9, "sum_flops_per_gpu": fl,
                "alg_GBs_per_gpu": 16.0 * float((ns * ns).sum()) / (med * 1e-3) / 1e9,
                "info_max": int(dinfo.abs().max().item())})
    return out


def run_strong(mb, torch, np, q, local, rank, world, barrier, max_over_ranks, dist, dev):
    """STRONG scaling (north_star: "the batch is sharded across the GPUs of one box by matrix index, with per-GPU
    queues and no NCCL collective"): ONE batch of the BASELINE size, rank r owns mgpu.shard_range(batch, world, r).
    t_full = the same call over the whole batch on one GPU (every rank runs it at once on its own GPU; max taken),
    efficiency = t_full / (world * t_shard). C4: ONE 20,000-matrix vbatched batch, LPT partition by LU cost."""
    from magma_b200 import mgpu
    stream = torch.cuda.current_stream(local)
    rows = []

    def timed(fn, restore, reps=3):
        ts = []
        for _ in range(reps):
            restore()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rc = fn()
            e1.record(stream)
            barrier()
            assert rc == 0
            ts.append(e0.elapsed_time(e1))
        return min(ts)

    def fixed(name, n, batch, nrhs=0):
        lo, hi = mgpu.shard_range(batch, world, rank)
        res = {}
        for tag, cnt in (("full", batch), ("shard", hi - lo)):
            db = mb.DeviceBatch(cnt, n, n, nrhs=nrhs, device=local, queue=q)
            seed = np.array([17, 0, 0, 1], dtype=np.int32)  # the SAME global batch on every rank ...
            if tag == "shard" and lo > 0:                   # ... of which this rank takes [lo, hi): skip lo matrices
                seed = advance_seed(seed, lo * n * n)
            mb.dlarnv_uniform(seed, cnt * n * n, db.A, q)
            if nrhs:
                mb.dlarnv_uniform(np.array([19, rank, 0, 1], dtype=np.int32), cnt * n * nrhs, db.B, q)
            q.sync()
            A0 = db.A.clone()
            B0 = db.B.clone() if nrhs else None

            def restore():
                db.A.copy_(A0)
                if nrhs:
                    db.B.copy_(B0)
            t = timed(db.gesv if nrhs else db.getrf, restore)
            res[tag] = max_over_ranks(t)
            del db, A0, B0
            torch.cuda.empty_cache()
        fl = (flops_getrf(n, n) + (flops_getrs(n, nrhs) if nrhs else 0)) * batch
        rows.append({"config": name, "total_batch": batch, "ms_one_gpu": res["full"], "ms_sharded": res["shard"],
                     "gflops_sharded": fl / (res["shard"] * 1e-3) / 1e9, "speedup": res["full"] / res["shard"],
                     "strong_efficiency": res["full"] / (world * res["shard"])})

    fixed("C2 dgesv_batched n=16 nrhs=1, ONE batch of 1000000", N_HEAD, BATCH_HEAD, nrhs=NRHS_HEAD)
    fixed("C3 dgetrf_batched n=128, ONE batch of 50000", 128, 50_000)
    fixed("C5 dgetrf_batched n=512, ONE batch of 4000", 512, 4_000)

    # C4: one batch for the whole box, partitioned by LU cost (mgpu.lpt_partition)
    batch = 20_000
    x, ns = 1234, []
    for _ in range(batch):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        ns.append(16 + (x >> 8) % 497)
    ns = np.array(ns, dtype=np.int64)
    for scheme in ("lpt", "contiguous"):
        parts = mgpu.lpt_partition(ns, ns, world) if scheme == "lpt" else \
            [np.arange(*mgpu.shard_range(batch, world, r)) for r in range(world)]
        mine = ns[parts[rank]]
        offs = np.concatenate([[0], np.cumsum(mine * mine)])
        poffs = np.concatenate([[0], np.cumsum(mine)])
        dA = torch.empty(int(offs[-1]), dtype=torch.float64, device=dev)
        mb.dlarnv_uniform(np.array([21, rank, 0, 1], dtype=np.int32), int(offs[-1]), dA, q)
        q.sync()
        A0 = dA.clone()
        dip = torch.zeros(int(poffs[-1]), dtype=torch.int32, device=dev)
        dinfo = torch.zeros(len(mine), dtype=torch.int32, device=dev)
        pA = torch.from_numpy(offs[:-1] * 8).to(dev) + dA.data_ptr()
        pP = torch.from_numpy(poffs[:-1] * 4).to(dev) + dip.data_ptr()
        dn = torch.from_numpy(mine.astype(np.int32)).to(dev)
        t = timed(lambda: mb.magma_dgetrf_vbatched(dn, dn, pA, dn, pP, dinfo, len(mine), q), lambda: dA.copy_(A0))
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        allt = [float(v.item()) for v in allt]
        fl = float(sum(flops_getrf(int(k), int(k)) for k in ns))
        rows.append({"config": f"C4 dgetrf_vbatched n~U[16,512], ONE batch of 20000, {scheme} partition",
                     "ms_per_rank": allt, "ms": max(allt), "gflops": fl / (max(allt) * 1e-3) / 1e9,
                     "time_imbalance_max_over_mean": max(allt) / (sum(allt) / world),
                     "cost_model_imbalance": mgpu.imbalance(ns, ns, parts)})
        del dA, A0, dip, dinfo
        torch.cuda.empty_cache()
    return rows


def run_ref_gpu(sweep, e2e, head_ms):
    """Same-box rows for the reference's own GPU path (oracle/_ref/libmagma_ref.so: MAGMA 2.10.0 kernels + cuBLAS,
    compiled by oracle/build_ref.py) and cublasDgetrfBatched, each in a child process (the
    reference exports the same symbol names as this library). Timed the way the reference testers time: wall clock
    around a queue sync, best of 3 (testing/testing_zgetrf_batched.cpp:203-206,233-249); ours: CUDA events."""
    import subprocess
    so = os.path.join(ROOT, "oracle", "_ref", "libmagma_ref.so")
    tool = os.path.join(ROOT, "tools", "ref_run.py")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libmagma_ref.so not built (python oracle/build_ref.py needs /root/reference)"}
    ours = {r["config"].split()[0]: r for r in sweep}
    rows = []
    for tag, n, batch, nrhs, our_ms in (("C1", 32, 10_000, 0, ours.get("C1", {}).get("ms")),
                                        ("C2", 16, 1_000_000, 1, head_ms),
                                        ("C3", 128, 50_000, 0, ours.get("C3", {}).get("ms")),
                                        ("C5", 512, 4_000, 0, ours.get("C5", {}).get("ms"))):
        row = {"config": tag, "n": n, "batch": batch, "nrhs": nrhs, "ours_ms": our_ms}
        for impl in ("magma", "cublas"):
            if impl == "cublas" and nrhs:
                continue
            try:
                r = subprocess.run([sys.executable, tool, str(n), str(batch), str(nrhs), "-", "3", "tile", impl],
                                   capture_output=True, text=True, timeout=600)
                js = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                ms = json.loads(js[-1])["ms_best"] if js else None
            except Exception as ex:  # noqa: BLE001
                ms = None
                row[impl + "_error"] = str(ex)[:200]
            row[impl + "_ms"] = ms
            if ms and our_ms:
                row["ours_over_" + impl] = ms / our_ms
        rows.append(row)
    return rows


if __name__ == "__main__":
    main()
