"""python tools/fused_time.py n batch [reps]  -- time dgetrf_batched (fused tier on)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
n, batch = int(sys.argv[1]), int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
if os.environ.get("FUSED_MAX"): mb.set_fused_max(int(os.environ["FUSED_MAX"]))
if os.environ.get("CHAIN_PANEL"): mb.set_chain_panel(int(os.environ["CHAIN_PANEL"]))
db = mb.DeviceBatch(batch, n, n, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q); q.sync(); A0 = db.A.clone()
fl = 0.5*n*(n*(n-n/3.0-1.0)+n)+2.0*n/3.0 + 0.5*n*(n*(n-n/3.0)-n)+n/6.0
for _ in range(reps):
    db.A.copy_(A0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"n={n} batch={batch}: {t:.3f} ms  {fl*batch/t/1e6:.0f} GF/s", flush=True)
