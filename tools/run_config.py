"""Run one fixed-size config a few times (for ncu / quick timing):
   python tools/run_config.py n batch [nrhs] [reps] [getrs]      env: TIER=0|1|2  SMALL_ROWS=0|1|2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
n, batch = int(sys.argv[1]), int(sys.argv[2])
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
do_getrs = len(sys.argv) > 5
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
if os.environ.get("TIER"): mb.set_tier(int(os.environ["TIER"]))
if os.environ.get("SMALL_ROWS"): mb.set_small_rows(int(os.environ["SMALL_ROWS"]))
if os.environ.get("MID_MAX"): mb.set_mid_max(int(os.environ["MID_MAX"]))
db = mb.DeviceBatch(batch, n, n, nrhs=nrhs, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q)
if nrhs: mb.dlarnv_uniform(seed, batch * n * nrhs, db.B, q)
q.sync(); A0 = db.A.clone(); B0 = db.B.clone() if nrhs else None
fl = 0.5*n*(n*(n-n/3.0-1.0)+n)+2.0*n/3.0 + 0.5*n*(n*(n-n/3.0)-n)+n/6.0
for _ in range(reps):
    db.A.copy_(A0)
    if nrhs: db.B.copy_(B0)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    if do_getrs:
        db.getrf(); e1.record(); db.getrs()
    elif nrhs: db.gesv(); e1.record()
    else: db.getrf(); e1.record()
    e2.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"n={n} batch={batch} nrhs={nrhs} getrf/gesv ms {t:.3f}  {fl*batch/t/1e6:.0f} GF/s" + (f"  getrs ms {e1.elapsed_time(e2):.3f} {nrhs*(2*n*n-n)*batch/e1.elapsed_time(e2)/1e6:.0f} GF/s" if do_getrs else ""), flush=True)
