python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import os,sys
sys.path.insert(0,'.')
os.environ['PROBE_MID']='1'
from magma_b200 import batched as mb
mb.set_mid_max(128)
import runpy
runpy.run_path('tools/gpu_probe.py', run_name='__main__')
PY
