#!/bin/bash
# dynamic instruction counts + durations of the two big kernels (ncu, two metrics; not a bench number)
tag=${1:-x}
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_stitch|k_resample|k_intersect' -s 8 -c 4 --csv --log-file gpurun_out/inst_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/inst_$tag.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")])
PY
