"""Run one fixed-size config a few times (for ncu):  python tools/run_config.py n batch [nrhs] [reps] [getrs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
n, batch = int(sys.argv[1]), int(sys.argv[2])
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
do_getrs = len(sys.argv) > 5
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
if nrhs: mb.dlarnv_uniform(seed, batch * n * nrhs, db.B, q)
q.sync(); A0 = db.A.clone(); B0 = db.B.clone() if nrhs else None
for _ in range(reps):
    db.A.copy_(A0)
    if nrhs: db.B.copy_(B0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if do_getrs:
        db.getrf(); db.getrs()
    elif nrhs: db.gesv()
    else: db.getrf()
    e1.record(); torch.cuda.synchronize()
    print(n, batch, nrhs, "ms", e0.elapsed_time(e1), flush=True)
