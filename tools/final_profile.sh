set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_f.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_f1.log 2>&1
$CMD > gpurun_out/plain_f2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_stitch|k_resample|k_intersect" -s 9 -c 4 -o gpurun_out/prof_final $CMD > gpurun_out/ncu_f2.log 2>&1
tail -n 2 gpurun_out/ncu_f2.log
