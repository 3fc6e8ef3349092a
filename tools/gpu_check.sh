#!/bin/bash
# GPU check used between kernel changes: parity tests, then the default bench; prints the stage times.
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err
python - <<PY
import json
r = json.load(open("gpurun_out/b_$tag.json"))
print("ms/step", round(r["ms_per_step"], 4), {k: round(v, 4) for k, v in r["roofline"]["stage_ms_per_step"].items()}, "frac", round(r["roofline"]["frac"], 4), "e2e", round(r["e2e"]["ms_per_step"], 2))
PY
