"""python tools/getri_time.py n batch  -- time magma_dgetri_outofplace_batched, single-launch kernel vs identity + getrs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
n, batch = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
db = mb.DeviceBatch(batch, n, n, nrhs=n, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q); q.sync()
db.getrf(); q.sync()
for mode in (1, 0, 1, 0):
    mb.set_getri_fused(mode); ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mb.magma_dgetri_outofplace_batched(n, db.dA_array, db.ldda, db.dipiv_array, db.dB_array, db.lddb, db.info, batch, q)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"n={n} batch={batch} fused={mode}: {t:.3f} ms  {2*8*n*n*batch/t/1e6:.0f} GB/s  {2.0*n**3*batch/t/1e9:.2f} TF/s", flush=True)
mb.set_getri_fused(1)
