"""Column timeline of lu_mid_kernel for CTA 0 (needs tools/libtrace.so built with -DMB200_MID_TRACE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from magma_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libtrace.so")
from magma_b200 import batched as mb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 4
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
db = mb.DeviceBatch(batch, n, n, queue=q)
seed = np.array([0, 0, 0, 1], dtype=np.int32)
mb.dlarnv_uniform(seed, batch * n * n, db.A, q); q.sync()
A0 = db.A.clone()
for _ in range(2):
    db.A.copy_(A0); torch.cuda.synchronize(); db.getrf(); torch.cuda.synchronize()
L = _lib.load()
out = (C.c_longlong * 1024)()
L.magma_b200_mid_trace(out)
t = np.array(out[:], dtype=np.int64)
t0 = t[1000]
cols = t[:n] - t0
print("start->col0", cols[0], " end(w0)", t[1001] - t0, " end(w1)", t[1002] - t0)
d = np.diff(cols)
print("per-column deltas by block:")
for b in range(0, n, 8):
    print(b, cols[b], d[b:b+8] if b + 8 <= n - 1 else d[b:])
print("mean in-block", np.mean([d[i] for i in range(len(d)) if (i + 1) % 8 != 0]), "mean handoff", np.mean([d[i] for i in range(len(d)) if (i + 1) % 8 == 0]))
