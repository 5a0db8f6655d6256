"""e2e (host pinned -> C ABI -> host) timing of the headline workload for a few staging chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magma_b200 import batched as mb
n, nrhs, batch = 16, 1, 1_000_000
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
hA0 = torch.rand((batch, n, n), dtype=torch.float64).pin_memory(); hB0 = torch.rand((batch, nrhs, n), dtype=torch.float64).pin_memory()
hA = torch.empty_like(hA0).pin_memory(); hB = torch.empty_like(hB0).pin_memory()
hip = torch.empty((batch, n), dtype=torch.int32).pin_memory(); hinfo = torch.empty((batch,), dtype=torch.int32).pin_memory()
for mbs in (256, 64, 32, 16, 8):
    os.environ["MB200_HOST_CHUNK_MB"] = str(mbs)
    ts = []
    for i in range(4):
        hA.copy_(hA0); hB.copy_(hB0); torch.cuda.synchronize()
        t0 = time.perf_counter(); rc = mb.dgesv_batched_host(n, nrhs, hA, n, hip, hB, n, hinfo, batch, q); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0); assert rc == 0
    print(mbs, "MB chunks:", [round(t * 1e3, 2) for t in ts], flush=True)
