#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- tools/profile_round.sh TAG): the default bench, then the two ncu passes of the
# profiling recipe on the short form of the same command (each after a plain run of it exited 0).  Outputs -> gpurun_out/.
tag=${1:-rX}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity"
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 300 gpurun_out/${tag}_bench.err
$CMD > gpurun_out/${tag}_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on \
    -k regex:"k_stitch_group|k_stitch_list|k_resample|k_intersect" -s 12 -c 6 -f -o gpurun_out/${tag}_full $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -n 2 gpurun_out/${tag}_ncu2.log
CMD3="python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-parity --workload cfg3"
$CMD3 > gpurun_out/${tag}_cfg3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/${tag}_cfg3_launches.csv $CMD3 > gpurun_out/${tag}_cfg3_ncu1.log 2>&1
