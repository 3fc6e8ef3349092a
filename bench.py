#!/usr/bin/env python
"""Benchmark of the multiplane slicing hot path (BASELINE.json metric: cross-section planes/sec;
bones/sec and % of HBM roofline ride along in the same JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]

A "step" is one pass of the whole hot path (bucket -> intersect -> stitch -> resample/unroll)
over one batch of synthetic bones.  Default workload = BASELINE.json configs[1]: the reference's
``humerus_left`` test bone under the config-4 jitter, 2,048 planes along the shaft axis per bone,
outlines resampled to 360 points + the 360-ray radius image, ``--bones`` bones per step per GPU.

  value  planes/s with the batch already resident in HBM (shb_batch_run only), CUDA events
  e2e    planes/s through the public host call (shb_sweep_batch): pinned host inputs -> H2D ->
         kernels -> D2H of the arrays the Slices API serves, wall clock
  roofline      dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the numpy restatement of the trimesh path (oracle/) on a bounded sample, 1 core

``--impl reference`` times that CPU restatement with every host core (trimesh itself is not
installable here) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HBM_FALLBACK_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def make_bones(workload: str, n_bones: int, first_id: int, planes: int, interp: int):
    """Returns (meshes, sweeps) in the packing order of shoulder_b200._lib._pack."""
    from shoulder_b200 import meshio
    base = meshio.load_mesh(ROOT / "tests" / "golden" / "bones" / "humerus_left.npz")
    names = ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"]
    bases = {n: meshio.load_mesh(ROOT / "tests" / "golden" / "bones" / f"{n}.npz") for n in names} if workload == "cfg4" else None
    meshes, sweeps = [], []
    for i in range(n_bones):
        bid = first_id + i
        if workload == "cfg3":
            obb = meshio.PcaObb(base)
            v, f = meshio.loop_subdivide(obb.mesh.vertices, obb.mesh.faces, 2)        # 519,040 triangles
            m = meshio.Mesh(v, f)
        else:
            src = bases[names[bid % 4]] if workload == "cfg4" else base
            m = meshio.PcaObb(meshio.synthetic_bone(src, bid)).mesh
        z = m.vertices[:, 2]
        k = len(meshes)
        meshes.append((m.vertices, m.faces))
        if workload == "cfg4":      # the three default sweeps of bone.Humerus (bone.py:116-121)
            full = np.linspace(0.99 * z.max(), 0.99 * z.min(), 200)
            dist = np.linspace(0.99 * z.min(), 0.0, 200)
            prox = np.linspace(0.99 * z.max(), 0.55 * z.max(), 600)     # neck_z stand-in (ruptures absent)
            for zs, n in ((full, 100), (dist, 500), (prox, 512)):
                sweeps.append((k, float(zs.mean()), zs - zs.mean(), n))
        else:
            zs = np.linspace(0.99 * z.max(), 0.99 * z.min(), planes)
            sweeps.append((k, float(zs.mean()), zs - zs.mean(), interp))
    return meshes, sweeps


# ------------------------------------------------------------------------------------------
# CPU restatement (oracle) timing — the reference arm and the cpu_baseline leg
# ------------------------------------------------------------------------------------------
def _cpu_one(args):
    import oracle
    from oracle.slice_arrays import radial_image
    v, f, zs, n, angles = args
    s = oracle.OracleSlices(v, f, zs, n, merge="hash", version="4")
    _ = (s.centroids, s.areas1, s.ixy, s.itr_start, s.itr_centered_start)
    if angles:
        radial_image(s.paths, angles)
    return len(zs)


def cpu_jobs(meshes, sweeps, angles, max_planes=None):
    jobs = []
    for (k, zo, h, n) in sweeps:
        zs = np.asarray(h) + zo
        if max_planes is not None and len(zs) > max_planes:
            zs = zs[:: len(zs) // max_planes][:max_planes]
        jobs.append((meshes[k][0], meshes[k][1], zs, n, angles))
    return jobs


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from before the warm-up to the end of
    the run; only samples whose arrival time falls inside a timed region are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        inside = [r for (t, r) in self.rows if len(r) >= 9 and any(a - 0.06 <= t <= b + 0.06 for a, b in self.windows)]
        if not inside:                       # region shorter than the sampling period: take every sample of the run
            inside = [r for (_, r) in self.rows if len(r) >= 9]
        num = lambda x: float(x) if x.replace(".", "", 1).isdigit() else None
        sm = [num(r[1]) for r in inside if num(r[1]) is not None]
        mx = [num(r[2]) for r in inside if num(r[2]) is not None]
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


def stage_alg_bytes(stage: str, c: dict) -> float:
    """Compulsory HBM bytes of one launch of each stage: every input it must
    read once and every output it must write once; hash tables, sort passes and re-gathers of the
    L2-resident mesh are NOT credited."""
    V, T, S, P, C, PN, items, A = c["V"], c["T"], c["S"], c["P"], c["C"], c["PN"], c["items"], c["PA"]
    if stage == "bucket":
        return 16 * T + 8 * V + 8 * P + 8 * items
    if stage == "scatter":
        return 8 * items + 16 * c["M"]
    if stage == "intersect":    # bucketed triangles + mesh z in; 16-byte hit records out (one pass)
        return 16 * c["M"] + 16 * T + 8 * V + 8 * P + 16 * S
    if stage == "stitch":       # hit records + mesh once in; closed contours + plane records out (+ face_index, segments when requested)
        return 16 * S + 16 * T + 32 * V + (36 * S if c.get("full") else 0) + 16 * (S + C) + 20 * C + 200 * P
    if stage == "resample":     # chosen outlines in; k profile arrays + radius image out
        return 16 * (S + C) + 8 * c["k"] * 2 * PN + 8 * A + 88 * P
    return 8 * P * 6


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_bones = min(args.bones, cores) if args.workload != "cfg3" else 1
    meshes, sweeps = make_bones(args.workload, n_bones, 0, args.planes, args.interp)
    jobs = cpu_jobs(meshes, sweeps, args.angles, max_planes=args.ref_planes)
    planes_per_step = sum(len(j[2]) for j in jobs)
    with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
        for _ in range(args.warmup_ref):
            pool.map(_cpu_one, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps_ref):
            pool.map(_cpu_one, jobs)
        dt = time.perf_counter() - t0
    value = planes_per_step * args.steps_ref / dt
    used = min(cores, len(jobs))
    sample = (f"{n_bones} bone(s) x {len(jobs) // max(n_bones, 1)} sweep(s), {planes_per_step} planes per step "
              f"(at most {args.ref_planes} planes per sweep), "
              f"numpy restatement of the trimesh path, multiprocessing over sweeps")
    line = {
        "impl": "reference", "metric": "planes_per_sec", "value": value, "unit": "planes/s", "n_gpus": args.gpus,
        "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": 1e3 * dt / args.steps_ref, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.bones if args.workload != "cfg3" else 1),
        "bones_per_sec": n_bones * args.steps_ref / dt,
        "cpu_baseline": {"value": value, "unit": "planes/s", "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "trimesh is not installable in this image; this is the oracle/ restatement of its algorithm (parity unpinned)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, bones):
    desc = {"cfg2": "BASELINE configs[1]: humerus_left (config-4 jitter per bone), dense multiplane sweep along the shaft axis + radial unroll",
            "cfg3": "BASELINE configs[2]: Loop-subdivided humerus_left, 519,040 triangles, single mesh",
            "cfg4": "BASELINE configs[3]: jittered test bones, the three default sweeps of bone.Humerus (200x100, 200x500, 600x512)"}
    return {"workload": desc[args.workload], "bones_per_step_per_gpu": bones, "planes_per_bone": args.planes if args.workload != "cfg4" else 1000,
            "interp_num": args.interp if args.workload != "cfg4" else "100/500/512", "radial_angles": args.angles,
            "triangles_per_bone": 32440 if args.workload != "cfg3" else 519040,
            "l2": "256 MiB buffer written between timed steps (L2 flush)",
            "sharding": "by contiguous plane range of the one mesh, no collective" if args.workload == "cfg3" else "by bone, no data-path collective"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    from shoulder_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: shoulder_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries the ONE JSON line and nothing else: NCCL prints its version banner there when the first
        # communicator comes up, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.init(local_rank)
    # an explicit (non-NULL) stream: the library enqueues on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    _lib.set_stream(stream.cuda_stream)

    bones = args.bones if args.workload != "cfg3" else 1
    meshes, sweeps = make_bones(args.workload, bones, rank * bones if args.workload != "cfg3" else 0, args.planes, args.interp)
    if args.workload == "cfg3" and world > 1:
        # one large mesh: replicate it, shard the sweep by contiguous plane range (z_orig stays the full-list mean)
        from shoulder_b200 import sharding
        k, zo, h, n = sweeps[0]
        lo, hi = sharding.shard_planes(len(h), rank, world)
        sweeps = [(k, zo, h[lo:hi], n)]
    packed = list(_lib._pack(meshes, sweeps))
    # pinned host copies: the e2e leg copies from pinned memory
    pinned = []
    for a in packed:
        t = torch.from_numpy(a).pin_memory()
        pinned.append(t)
    packed_pinned = tuple(t.numpy() for t in pinned)
    h2d = int(sum(a.nbytes for a in packed_pinned))
    planes_per_step = int(packed[7][-1])
    mask_dev = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | (_lib.OUT_RADIAL if args.angles else 0)
    k_prof = 3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg -------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    batch = _lib.SweepBatch(None, None, packed=packed_pinned)
    counts = None
    for _ in range(args.warmup):
        r = batch.run(mask_dev, args.angles)
        if counts is None:
            tot = r.totals()
            counts = tot
        r.close()
    torch.cuda.synchronize()
    # every stage timed once, OUTSIDE the timed region, for the stage table; inside the timed region only the three
    # heavy stages carry event pairs (an event pair costs ~3 us of device time, all seven ~2.5 % of this step)
    _lib.profile_enable(True)
    _lib.profile_read(reset=True)
    n_tab = max(2, min(args.steps, 5))
    for _ in range(n_tab):
        flush.fill_(1)
        batch.run(mask_dev, args.angles).close()
    torch.cuda.synchronize()
    stage_table = {n: v[0] / n_tab for n, v in _lib.profile_read(reset=True).items()}
    heavy = ("intersect", "stitch", "resample")
    _lib.profile_enable(True, stages=heavy)
    barrier()
    t_dev0 = time.perf_counter()
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)                      # L2 flush, outside the timed events
        a.record(stream)
        r = batch.run(mask_dev, args.angles)
        b.record(stream)
        r.close()
    barrier()
    sampler.window(t_dev0, time.perf_counter())
    launches = _lib.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    stages = {n: v for n, v in _lib.profile_read(reset=True).items() if n in heavy}      # measured live, inside the timed region
    _lib.profile_enable(False)

    # ---------------- end-to-end leg (public host call) -----------------------------------
    # the batch is handed over as `e2e_chunks` groups of whole bones (pinned host arrays); the library overlaps
    # each group's device->host copy with the next group's upload + kernels.  Every step moves every input byte
    # host->device and every output byte device->host.
    mask_e2e = mask_dev
    chunks, first = _lib.split_packed(packed_pinned, args.e2e_chunks)
    chunks = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in c) for c in chunks]

    def e2e_call(mask):
        return _lib.sweep_batch_pipelined(chunks, first, mask, args.angles)

    for _ in range(max(2, args.warmup)):
        r = e2e_call(mask_e2e)
        r.close()
    barrier()
    d2h = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        r = e2e_call(mask_e2e)
        if i == 0:
            for s in range(r.n_sweep):
                for w in (_lib.ARR_N_SEG, _lib.ARR_N_ENT, _lib.ARR_STATUS, _lib.ARR_SEL, _lib.ARR_BOUNDS, _lib.ARR_CENTROID,
                          _lib.ARR_AREA1, _lib.ARR_IXY, _lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED_START):
                    d2h += r.array(w, s).nbytes
                d2h += 4 * (len(r.array(_lib.ARR_SEG_OFF, s)))
                if args.angles:
                    d2h += r.array(_lib.ARR_RADIAL, s).nbytes
        r.close()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.window(t0, t0 + e2e_s)
    # same call with float32 profile / radius outputs (SHB_OUT_F32), reported beside the float64 headline
    for _ in range(2):
        e2e_call(mask_e2e | _lib.OUT_F32).close()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for i in range(args.steps):
        e2e_call(mask_e2e | _lib.OUT_F32).close()
    torch.cuda.synchronize()
    e2e32_s = time.perf_counter() - t1
    clocks = sampler.stop()

    # ---------------- reduce over ranks ----------------------------------------------------
    tvals = torch.tensor([ms, e2e_s * 1e3, e2e32_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tvals, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e32_ms_max = float(tvals[0]), float(tvals[1]), float(tvals[2])
    ptot = torch.tensor([planes_per_step], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ptot, op=dist.ReduceOp.SUM)
    total_planes = int(ptot.item())
    value = total_planes * args.steps / (ms_max * 1e-3)
    e2e_value = total_planes * args.steps / (e2e_ms_max * 1e-3)

    # ---------------- roofline of the dominant kernel (rank 0, its own GPU) ---------------
    V = int(packed[1][-1]); T = int(packed[3][-1])
    S = counts["segments"]; Cn = counts["contours"]; P = planes_per_step
    PN = int(sum(len(s[2]) * s[3] for s in sweeps)); PA = P * args.angles
    items = int(sum(len(meshes[s[0]][1]) for s in sweeps))
    c = {"V": V, "T": T, "S": S, "P": P, "C": Cn, "PN": PN, "items": items, "M": items, "PA": PA, "k": k_prof}
    dom = max(stages, key=lambda n: stages[n][0])
    dom_ms = stages[dom][0] / args.steps
    peak, peak_kind = hbm_peak()
    alg = stage_alg_bytes(dom, c)
    achieved = alg / (dom_ms * 1e-3) / 1e9
    # whole step: inputs once + the outputs this run delivers (intermediates such as hit lists and contour points are not credited)
    pipeline_alg = 24 * V + 12 * T + 8 * k_prof * 2 * PN + 8 * PA + 76 * P
    traffic = None
    tfile = ROOT / "profiles" / "r1g_traffic.json"
    if tfile.exists() and args.workload == "cfg2" and bones == 32 and args.planes == 2048 and args.interp == 360 and args.angles == 360:
        for name, rec in json.loads(tfile.read_text())["kernels"].items():     # ncu --set full capture of this very command
            if name.startswith("k_" + dom) and not name.endswith("<0>"):
                traffic = rec["traffic_bytes"]
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "alg_bytes_per_launch": alg, "ms_per_launch": dom_ms,
                "stage_ms_per_step": {n: (stages[n][0] / args.steps if n in stages else stage_table[n]) for n in stage_table},
                "pipeline": {"alg_bytes_per_step": pipeline_alg, "achieved": pipeline_alg / (ms_max / args.steps * 1e-3) / 1e9,
                             "frac": pipeline_alg / (ms_max / args.steps * 1e-3) / 1e9 / peak}}

    if rank != 0:
        return
    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only) -----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        n_pl, nb, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0 and nb < len(meshes):          # >= ~12 s of CPU work, whole bones
            for j in cpu_jobs(meshes, [s for s in sweeps if s[0] == nb], args.angles, max_planes=args.cpu_planes):
                n_pl += _cpu_one(j)
            nb += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n_pl / dt, "unit": "planes/s", "cores": 1, "kind": "port",
               "sample": f"first {nb} bone(s) of the batch, {n_pl} planes, {dt:.1f} s of single-thread numpy (oracle/ restatement of the trimesh path)"}
    line = {
        "metric": "planes_per_sec", "value": value, "unit": "planes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "cfg3" else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, bones),
        "bones_per_sec": bones * world * args.steps / (ms_max * 1e-3),
        "segments_per_step_per_gpu": S, "contours_per_step_per_gpu": Cn,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "planes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms_max / args.steps, "bones_per_sec": bones * world * args.steps / (e2e_ms_max * 1e-3),
                "outputs": "plane records + ixy + itr_start + itr_centered_start (+ radius image), float64, every plane",
                "call": f"shoulder_b200._lib.sweep_batch_pipelined, {len(chunks)} groups of bones"},
        "e2e_f32": {"value": total_planes * args.steps / (e2e32_ms_max * 1e-3), "unit": "planes/s",
                    "ms_per_step": e2e32_ms_max / args.steps, "note": "same call with SHB_OUT_F32 (float32 profile arrays)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--bones", type=int, default=32, help="bones per step per GPU")
    ap.add_argument("--planes", type=int, default=None)
    ap.add_argument("--interp", type=int, default=360)
    ap.add_argument("--angles", type=int, default=360)
    ap.add_argument("--cpu-planes", type=int, default=None, help="planes per sweep in the cpu_baseline sample")
    ap.add_argument("--ref-planes", type=int, default=None, help="planes per sweep in one reference-arm step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="groups of bones the e2e call pipelines (D2H of one overlaps compute of the next)")
    args = ap.parse_args()
    if args.planes is None:
        args.planes = 8192 if args.workload == "cfg3" else 2048
    if args.workload == "cfg3" and args.cpu_planes is None:
        args.cpu_planes = 512
    if args.workload == "cfg3" and args.ref_planes is None:
        args.ref_planes = 16
    args.warmup = max(args.warmup, 3)
    # the reference arm keeps the driver's K and W; its per-step sample is bounded so the run ends in minutes
    args.steps_ref, args.warmup_ref = args.steps, args.warmup
    if args.ref_planes is None:
        budget_s, per_plane_s = 150.0, 0.005        # ~5 ms per plane per core for the numpy restatement
        args.ref_planes = max(32, int(budget_s / ((args.steps + args.warmup) * per_plane_s)))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
