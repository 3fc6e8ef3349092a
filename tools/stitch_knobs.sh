#!/bin/bash
# stitch stage ms on config 3 under the launch knobs of shb_launch_stitch: tools/stitch_knobs.sh
run() { wl=$1; shift; env "$@" python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu --no-sub --no-parity 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('$wl', '$*', 'ms/step', round(d['ms_per_step'],4), 'stitch', round(s['stitch'],4), 'resample', round(s['resample'],4))"; }
run cfg3l3 A=1
run cfg3l3 SHB_DEBUG_STITCH_NW=2047
run cfg3l3 SHB_DEBUG_STITCH_NW=2047 SHB_DEBUG_STITCH_ARENA=110000
run cfg3l3 SHB_DEBUG_STITCH_NW=2047 SHB_DEBUG_STITCH_ARENA=150000
run cfg3l3 SHB_DEBUG_STITCH_ARENA=112000
run cfg3 A=1
run cfg3 SHB_DEBUG_STITCH_ARENA=72000
run cfg3 SHB_DEBUG_STITCH_ARENA=90000
run cfg3 SHB_DEBUG_STITCH_ARENA=140000
