#!/bin/bash
# GPU box: end-to-end step (packed host rows in, PRG strings out) against the number of ranges a build is cut into
for r in 1 2 3 4 6; do
  MPRG_DEV_RANGES=$r python bench.py --steps 20 --warmup 5 --no-files --no-big 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ranges $r: resident', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['step_ms']['e2e'], 'from_text', round(d['e2e']['from_text']['ms_per_step'],3))"
done
