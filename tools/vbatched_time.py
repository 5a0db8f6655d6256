"""python tools/vbatched_time.py [reps]  -- BASELINE config 4 alone: magma_dgetrf_vbatched, 20000 square matrices n ~ U[16,512]
(the generator of bench.py's C4 row); for timing and for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
batch, x, ns = 20_000, 1234, []
for _ in range(batch):
    x = (x * 1103515245 + 12345) & 0x7FFFFFFF
    ns.append(16 + (x >> 8) % 497)
ns = np.array(ns, dtype=np.int64)
if os.environ.get("EVEN"):  # every size rounded up to even: shows what the aligned (TMA) paths are worth
    ns = (ns + 1) // 2 * 2
offs = np.concatenate([[0], np.cumsum(ns * ns)]); poffs = np.concatenate([[0], np.cumsum(ns)])
dev = torch.device("cuda", 0)
dA = torch.empty(int(offs[-1]), dtype=torch.float64, device=dev)
mb.dlarnv_uniform(np.array([21, 0, 0, 1], dtype=np.int32), int(offs[-1]), dA, q); q.sync()
A0 = dA.clone()
dip = torch.zeros(int(poffs[-1]), dtype=torch.int32, device=dev)
dinfo = torch.zeros(batch, dtype=torch.int32, device=dev)
pA = torch.from_numpy(offs[:-1] * 8).to(dev) + dA.data_ptr()
pP = torch.from_numpy(poffs[:-1] * 4).to(dev) + dip.data_ptr()
dn = torch.from_numpy(ns.astype(np.int32)).to(dev)
fl = float(sum(0.5*n*(n*(n-n/3.0-1.0)+n)+2.0*n/3.0 + 0.5*n*(n*(n-n/3.0)-n)+n/6.0 for n in ns))
for _ in range(reps):
    dA.copy_(A0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = mb.magma_dgetrf_vbatched(dn, dn, pA, dn, pP, dinfo, batch, q); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"C4 vbatched: {t:.3f} ms  {fl/t/1e6:.0f} GF/s rc={rc}", flush=True)
