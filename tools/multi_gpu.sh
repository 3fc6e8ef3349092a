#!/bin/bash
# tools/multi_gpu.sh N TAG : default bench (headline + cfg3 / cfg4 sub-records) and cfg4 at 128 bones per GPU on N GPUs of one box
N=$1; tag=$2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 10 --warmup 3 --no-parity > gpurun_out/${tag}_n${N}.json 2> gpurun_out/${tag}_n${N}.err
tail -c 300 gpurun_out/${tag}_n${N}.err
run 29512 --steps 5 --warmup 3 --workload cfg4 --bones 128 --no-parity --no-sub > gpurun_out/${tag}_cfg4x128_n${N}.json 2> gpurun_out/${tag}_cfg4x128_n${N}.err
tail -c 300 gpurun_out/${tag}_cfg4x128_n${N}.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_n${N}.json", "gpurun_out/${tag}_cfg4x128_n${N}.json"):
    try:
        r = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(r["value"] / 1e6, 2), "M planes/s", round(r["ms_per_step"], 3), "ms; e2e", round(r["e2e"]["value"] / 1e6, 2), "M planes/s", round(r["e2e"]["ms_per_step"], 2), "ms; bones/s", round(r["bones_per_sec"]), round(r["e2e"]["bones_per_sec"]))
    for k, v in (r.get("configs") or {}).items():
        print("   ", k, v.get("error") or ((round(v["value"] / 1e6, 2), round(v["ms_per_step"], 3), "e2e", round(v["e2e"]["ms_per_step"], 2), "bones/s", round(v["bones_per_sec"]), round(v["e2e"]["bones_per_sec"])) if "e2e" in v else v))
PY
