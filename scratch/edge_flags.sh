#!/bin/bash
for f in 0 1 2 4 6 8 16 31; do
  echo "== GNB_EDGE_FLAGS=$f"
  GNB_EDGE_FLAGS=$f timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ms/step %.3f  edge_core %.3f ms/launch' % (d['ms_per_step'], d['kernels']['tc_edge_core']['ms_per_step']/4))
    elif 'Error' in l or 'error' in l: print(l.strip()[:200])
"
done
