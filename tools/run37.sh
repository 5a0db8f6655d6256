#!/bin/bash
mkdir -p gpurun_out
SMALL=6 timeout 300 python - <<'PY' > gpurun_out/r37_tests.log 2>&1
import numpy as np, oracle
from magma_b200 import batched as mb
import torch
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
mb.set_small_rows(6)
ok = True
for (m, n, batch) in [(33,33,21),(40,40,9),(44,44,9),(64,64,9),(48,64,5),(64,40,5),(35,60,4),(33,64,3)]:
    A0, _ = oracle.random_batch(batch, m, n)
    db = mb.DeviceBatch(batch, m, n, queue=q); db.upload(A0); assert db.getrf() == 0
    LU, ipiv, info = db.download()
    ref = A0.copy(); ipr, infr = oracle.getrf_batched(ref, m)
    e = np.array_equal(LU, ref) and np.array_equal(ipiv, ipr) and np.array_equal(info, infr)
    print(m, n, batch, "OK" if e else "MISMATCH"); ok &= e
print("ALL OK" if ok else "FAILED")
PY
tail -3 gpurun_out/r37_tests.log
for n in 36 40 44 48 64; do
  b=$((50000*128*128/n/n))
  SMALL_ROWS=7 timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
  SMALL_ROWS=6 timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
