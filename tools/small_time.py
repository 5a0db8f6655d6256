"""python tools/small_time.py n batch rows [reps]  -- time dgetrf_batched in the register tier with magma_b200_set_small_rows(rows)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from magma_b200 import batched as mb
n, batch, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
mb.set_small_rows(rows)
# parity on a small sample first
A0, _ = oracle.random_batch(37, n, n)
db = mb.DeviceBatch(37, n, n, queue=q); db.upload(A0); db.getrf(); LU, ipiv, info = db.download()
ref = A0.copy(); ipr, infr = oracle.getrf_batched(ref, n)
ok = np.array_equal(LU, ref) and np.array_equal(ipiv, ipr) and np.array_equal(info, infr)
db = mb.DeviceBatch(batch, n, n, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q); q.sync(); A0 = db.A.clone()
ts = []
for _ in range(reps):
    db.A.copy_(A0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts)
print(f"n={n} batch={batch} rows={rows}: {t:.3f} ms  {2*8*n*n*batch/t/1e6:.0f} GB/s  parity={'ok' if ok else 'BAD'}", flush=True)
