"""Quick on-GPU probe: FP64/HBM peaks and raw kernel timings for the BASELINE shapes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb

def flops_getrf(n):
    m = n
    mul = 0.5 * n * (n * (m - n / 3.0 - 1.0) + m) + 2.0 * n / 3.0
    add = 0.5 * n * (n * (m - n / 3.0) - m) + n / 6.0
    return mul + add

import os
CONFIGS=[(32, 10000, 0), (32, 400000, 0), (16, 1000000, 1), (16, 1000000, 0), (8, 2000000, 0), (24, 400000, 0), (64, 100000, 0), (128, 50000, 0), (256, 8000, 0), (512, 4000, 0)]
if os.environ.get('PROBE_SMALL'): CONFIGS=[c for c in CONFIGS if c[0]<=32]
if os.environ.get('PROBE_MID'): CONFIGS=[(48,100000,0),(64,100000,0),(96,50000,0),(128,50000,0)]
def main():
    torch.cuda.set_device(0)
    mb.magma_init()
    q = mb.Queue.from_torch(0)
    if os.environ.get('SMALL_ROWS'): mb.set_small_rows(int(os.environ['SMALL_ROWS']))
    out = {}
    out["dfma_tflops"] = mb.fp64_peak_tflops(0, q)
    out["dmma_tflops"] = mb.fp64_peak_tflops(1, q)
    out["hbm_copy_gbs"] = mb.hbm_copy_gbs(1 << 30, q)
    print(json.dumps(out), flush=True)
    def timeit(fn, reset, reps=5):
        ts = []
        for _ in range(reps):
            reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts), float(np.median(ts))
    for (n, batch, nrhs) in CONFIGS:
        db = mb.DeviceBatch(batch, n, n, nrhs=max(nrhs, 0) or 0, queue=q) if nrhs else mb.DeviceBatch(batch, n, n, queue=q)
        seed = np.array([0, 0, 0, 1], dtype=np.int32)
        mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
        if nrhs: mb.dlarnv_uniform(seed, batch * n * nrhs, db.B, q)
        q.sync()
        A0 = db.A.clone(); B0 = db.B.clone() if nrhs else None
        def reset():
            db.A.copy_(A0)
            if nrhs: db.B.copy_(B0)
        fn = db.gesv if nrhs else db.getrf
        best, med = timeit(fn, reset)
        fl = flops_getrf(n) + (nrhs * (2 * n * n - n) if nrhs else 0)
        by = 16.0 * n * n + (16.0 * n * nrhs if nrhs else 0)
        print(json.dumps({"n": n, "batch": batch, "nrhs": nrhs, "ms_best": best, "ms_med": med,
                          "gflops": fl * batch / best / 1e6, "alg_GBs": by * batch / best / 1e6}), flush=True)
        if n == 512 and not nrhs:
            dbs = mb.DeviceBatch(batch, n, n, nrhs=16, queue=q)
            dbs.A.copy_(db.A); dbs.ipiv.copy_(db.ipiv)
            mb.dlarnv_uniform(seed, batch * n * 16, dbs.B, q)
            B0 = dbs.B.clone()
            best, med = timeit(dbs.getrs, lambda: dbs.B.copy_(B0))
            print(json.dumps({"getrs_n": n, "nrhs": 16, "ms_best": best, "gflops": 16 * (2 * n * n - n) * batch / best / 1e6,
                              "alg_GBs": (8.0 * n * n + 16 * n * 16) * batch / best / 1e6}), flush=True)
            del dbs
        del db, A0
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
