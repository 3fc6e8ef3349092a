#!/bin/bash
# Round-2 GPU check: parity tests, default bench, then stitch-kernel parameter sweeps (stage times only).
tag=${1:-r2x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.txt
cat gpurun_out/${tag}_pytest.txt
python bench.py --steps 10 --warmup 3 --no-cpu --no-sub --no-parity > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
r = json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", round(r["ms_per_step"], 4), {k: round(v, 4) for k, v in r["roofline"]["stage_ms_per_step"].items()}, "e2e", round(r["e2e"]["ms_per_step"], 2))
PY
