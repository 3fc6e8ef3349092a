#!/bin/bash
# e2e ms/step against the number of groups the host call pipelines: tools/e2e_chunks.sh
for wl in cfg4 cfg2; do
  for k in 1 2 4 8 16; do
    python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu --no-sub --no-parity --e2e-chunks $k 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$wl chunks $k: e2e', round(d['e2e']['ms_per_step'],2), 'ms', round(d['e2e']['value']/1e6,3), 'M planes/s', round(d['e2e']['bones_per_sec']), 'bones/s; f32', round(d['e2e_f32']['ms_per_step'],2) if d.get('e2e_f32') else None)"
  done
done
