#!/bin/bash
# launch list + full capture of the stitch / resample kernels on the short bench command (never a bench number)
tag=${1:-r2x}; kern=${2:-"k_stitch_group|k_stitch_list|k_merge_vertices|k_resample|k_intersect"}; wl=${3:-cfg2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity --workload $wl"
$CMD > gpurun_out/${tag}_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
python - <<PY
import csv, collections
tot, cnt = collections.Counter(), collections.Counter()
for r in csv.reader(open("gpurun_out/${tag}_launches.csv")):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        name = r[4].split("(")[0]; tot[name] += float(r[14]) / 1e3; cnt[name] += 1
for k, v in tot.most_common(14): print(f"{k:60s} n={cnt[k]:3d} avg_us={v / cnt[k]:9.1f}")
PY
$CMD > gpurun_out/${tag}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on \
    -k regex:"$kern" -s 10 -c 5 -f -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -n 2 gpurun_out/${tag}_ncu2.log
