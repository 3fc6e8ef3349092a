"""Stage times of one workload under several settings of the stitch / resample launch knobs (env hooks read per run).
   python tools/stitch_probe.py cfg2|cfg3|cfg4 [bones]"""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench
from shoulder_b200 import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
bones = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if wl.startswith("cfg3") else 32)
_lib.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); _lib.set_stream(stream.cuda_stream)
meshes, sweeps = bench.make_bones(wl, bones, 0, 8192 if wl.startswith("cfg3") else 2048, 360)
packed = list(_lib._pack(meshes, sweeps))
batch = _lib.SweepBatch(None, None, packed=packed)
mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
settings = [{}] + [json.loads(a) for a in sys.argv[3:]]
for env in settings:
    for k, v in env.items():
        os.environ[k] = str(v)
    _lib.profile_enable(True)
    for _ in range(3):
        batch.run(mask, 360).close()
    torch.cuda.synchronize()
    _lib.profile_read(reset=True)
    n = 10
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.fill_(1); a.record(stream); r = batch.run(mask, 360); b.record(stream); r.close()
    torch.cuda.synchronize()
    st = {k: round(v[0] / n, 4) for k, v in _lib.profile_read(reset=True).items()}
    print(wl, env, f"{sum(a.elapsed_time(b) for a, b in ev) / n:.4f} ms/step", st, flush=True)
    for k in env:
        del os.environ[k]
