"""Default bench workload, device-resident: ms/step with the per-stage event timers on and off (their cost is inside the
bench's timed region because the roofline line needs them)."""
import sys, argparse
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import bench
from shoulder_b200 import _lib

_lib.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); _lib.set_stream(stream.cuda_stream)
meshes, sweeps = bench.make_bones("cfg2", 32, 0, 2048, 360)
packed = list(_lib._pack(meshes, sweeps))
batch = _lib.SweepBatch(None, None, packed=packed)
mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prof in (False, True, False, True):
    _lib.profile_enable(prof)
    for _ in range(3):
        batch.run(mask, 360).close()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        flush.fill_(1); a.record(stream); r = batch.run(mask, 360); b.record(stream); r.close()
    torch.cuda.synchronize()
    print(f"stage timers {'on ' if prof else 'off'}: {sum(a.elapsed_time(b) for a, b in ev) / len(ev):.4f} ms/step")
    _lib.profile_read(reset=True)
