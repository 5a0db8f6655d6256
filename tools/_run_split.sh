cat > /tmp/ts.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from magma_b200 import batched as mb
n, batch = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
db = mb.DeviceBatch(batch, n, n, queue=q)
mb.dlarnv_uniform(np.array([0,0,0,1],dtype=np.int32), batch*n*n, db.A, q); q.sync(); A0 = db.A.clone()
for parts in (1, 2, 3, 4):
    mb.set_split(parts); ts=[]
    for _ in range(4):
        db.A.copy_(A0); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); db.getrf(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"prio={os.environ.get('MB200_SPLIT_PRIO')} n={n} batch={batch} split={parts}: {min(ts):.3f} ms", flush=True)
PY
for cfg in "128 50000" "512 4000" "256 16000"; do MB200_SPLIT_PRIO=1 python /tmp/ts.py $cfg; done
