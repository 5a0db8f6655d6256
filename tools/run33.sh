#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 60 python tools/run_config.py 32 1000000 0 3 | tail -1
TIER=2 timeout 60 python tools/run_config.py 32 1000000 0 3 | tail -1
timeout 60 python tools/run_config.py 24 1000000 0 3 | tail -1
TIER=2 timeout 60 python tools/run_config.py 24 1000000 0 3 | tail -1
