run() { python bench.py --no-sweep --no-cpu --no-ref --steps 6 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'e2e ms', round(d['e2e']['ms_per_step'],2), 'copy-only', round(d['e2e']['copy_only_ms'],2), 'kernel', round(d['ms_per_step'],4))"; }
for nb in 2 3 4 2 3; do MB200_HOST_NBUF=$nb run "nbuf=$nb chunk=64"; done
for mbs in 32 128; do MB200_HOST_NBUF=3 MB200_HOST_CHUNK_MB=$mbs run "nbuf=3 chunk=$mbs"; done
MB200_HOST_NBUF=4 MB200_HOST_CHUNK_MB=32 run "nbuf=4 chunk=32"
