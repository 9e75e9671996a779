#!/bin/bash
for m in 1 0; do
  echo "== GNB_EDGE_CTA_PAIR=$m"
  GNB_EDGE_CTA_PAIR=$m timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ms/step %.3f  edge_core %.3f ms/launch node_core %.3f' % (d['ms_per_step'], d['kernels']['tc_edge_core']['ms_per_step']/4, d['kernels']['tc_node_core']['ms_per_step']/4))
    elif 'rror' in l: print(l.strip()[:200])
"
done
