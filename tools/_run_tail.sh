timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "fused_tail or chain_panel" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
cat > /tmp/tt.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from magma_b200 import batched as mb
n, batch, lvl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
mb.set_fused_tail(lvl)
db = mb.DeviceBatch(batch, n, n, queue=q)
mb.dlarnv_uniform(np.array([0,0,0,1],dtype=np.int32), batch*n*n, db.A, q); q.sync(); A0 = db.A.clone()
ts=[]
for _ in range(4):
    db.A.copy_(A0); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"n={n} batch={batch} fused_tail={lvl}: {min(ts):.3f} ms", flush=True)
PY
for cfg in "128 50000" "96 50000" "64 100000" "48 100000" "112 50000"; do
 for lvl in 0 1 2 0 1 2; do python /tmp/tt.py $cfg $lvl; done
done
