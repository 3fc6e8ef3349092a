#!/bin/bash
# per-kernel durations (ncu, serialised) of the probe workload under an environment: tools/ktimes.sh TAG WL "ENV=..."
tag=$1; wl=$2; shift 2
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_stitch|k_merge|k_resample|k_intersect" -c 60 --csv \
    --log-file gpurun_out/kt_${tag}.csv python tools/stitch_probe.py $wl > gpurun_out/kt_${tag}.log 2>&1
python - <<PY
import csv, collections
tot = collections.defaultdict(list)
for r in csv.reader(open("gpurun_out/kt_${tag}.csv")):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        tot[r[4].split("(")[0]].append(float(r[14]) / 1e3)
print("$tag", "$@", {k[:28]: round(sorted(v)[len(v) // 2], 1) for k, v in tot.items()})
PY
