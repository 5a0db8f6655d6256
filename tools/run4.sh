python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, bench
from magma_b200 import batched as mb
torch.cuda.set_device(0); mb.magma_init(); q = mb.Queue.from_torch(0)
def barrier(): torch.cuda.synchronize()
out = bench.run_sweep(mb, torch, np, q, 0, 0, 1, barrier, lambda x: x, lambda x: x, 6546.6, 37000.0)
for r in out: print(r['config'], round(r['ms'],3), round(r['gflops']), r.get('getrs'))
PY
