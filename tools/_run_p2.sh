cat > /tmp/tp.py <<PY
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from magma_b200 import batched as mb
n, batch, on = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
mb.set_tall_panel(on)
db = mb.DeviceBatch(batch, n, n, queue=q)
mb.dlarnv_uniform(np.array([0,0,0,1],dtype=np.int32), batch*n*n, db.A, q); q.sync(); A0 = db.A.clone()
ts=[]
for _ in range(4):
    db.A.copy_(A0); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"n={n} batch={batch} tall_panel2={on}: {min(ts):.3f} ms", flush=True)
PY
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for cfg in "512 4000" "256 16000"; do for on in 0 1; do python /tmp/tp.py $cfg $on; done; done
python tools/vbatched_time.py 3 | tail -1
