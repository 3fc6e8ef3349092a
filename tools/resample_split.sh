#!/bin/bash
# resample stage ms with and without the two-class launch (SHB_DEBUG_RESAMPLE_SPLIT=1000000 disables it): tools/resample_split.sh
run() { wl=$1; shift; env "$@" python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu --no-sub --no-parity 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('$wl', '$*', 'ms/step', round(d['ms_per_step'],4), 'stitch', round(s['stitch'],4), 'resample', round(s['resample'],4), 'launches', d['gpu_launches'])"; }
for wl in cfg3 cfg3l3 cfg2 cfg4; do
  run $wl SHB_DEBUG_RESAMPLE_SPLIT=1000000
  run $wl A=1
done
run cfg3 SHB_DEBUG_RESAMPLE_SPLIT=600
run cfg3 SHB_DEBUG_RESAMPLE_SPLIT=900
run cfg3l3 SHB_DEBUG_RESAMPLE_SPLIT=1250
run cfg3l3 SHB_DEBUG_RESAMPLE_SPLIT=1800
