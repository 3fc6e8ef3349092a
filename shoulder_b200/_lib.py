"""ctypes binding of ``libshoulder_b200.so`` (C ABI: ``include/shoulder_b200.h``).

The library is the product; there is no Python or CPU fallback.  Loading fails loudly when the
shared object is missing, and ``init()`` fails loudly when there is no sm_100 device.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os as _os
LIB_PATH = Path(_os.environ.get("SHB_LIB") or Path(__file__).resolve().parent / "libshoulder_b200.so")      # SHB_LIB: experiments only

# --- constants mirrored from include/shoulder_b200.h ---------------------------------------
ABI_VERSION = 4
OUT_PLANE, OUT_SEGMENTS, OUT_CONTOURS = 0x001, 0x002, 0x004
OUT_IXY, OUT_IXY_CENTERED, OUT_ITR, OUT_ITR_START = 0x008, 0x010, 0x020, 0x040
OUT_ITR_CENTERED, OUT_ITR_CENTERED_START, OUT_RADIAL = 0x080, 0x100, 0x200
OUT_ALL_PROFILES = 0x1F8
OUT_F32 = 0x400
(ARR_N_SEG, ARR_SEG_OFF, ARR_N_ENT, ARR_STATUS, ARR_BOUNDS, ARR_CENTROID, ARR_AREA1, ARR_SEL, ARR_FACE_INDEX,
 ARR_SEGMENTS, ARR_CONTOUR_OFF, ARR_CONTOUR_PT_OFF, ARR_CONTOUR_AREA, ARR_POINTS, ARR_IXY, ARR_IXY_CENTERED, ARR_ITR,
 ARR_ITR_START, ARR_ITR_CENTERED, ARR_ITR_CENTERED_START, ARR_RADIAL, ARR_COUNT) = range(22)
STL_FRAME = 0x1
ST_EMPTY, ST_OPEN, ST_NONMANIFOLD, ST_RANK_TIE, ST_SPLIT_COPY, ST_GENERAL, ST_MERGED = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40
N_WINDOWED = 7                       # six profile arrays + the radius image (shb_sweep_request)
WINDOWED_BITS = (OUT_IXY, OUT_IXY_CENTERED, OUT_ITR, OUT_ITR_START, OUT_ITR_CENTERED, OUT_ITR_CENTERED_START, OUT_RADIAL)
N_STAGES = 7
STAGE_NAMES = ("bucket", "scan", "scatter", "intersect", "scan2", "stitch", "resample")
_DTYPES = {1: np.int32, 2: np.int64, 3: np.uint32, 4: np.float64, 5: np.float32}

EXPORTS = (
    "shb_init", "shb_set_stream", "shb_batch_create", "shb_batch_free", "shb_batch_run", "shb_batch_run_req", "shb_result_window",
    "shb_sweep_batch",
    "shb_result_fetch", "shb_result_fetch_async", "shb_result_array", "shb_result_totals", "shb_result_free", "shb_profile_enable",
    "shb_profile_read", "shb_trim", "shb_launch_count", "shb_last_error", "shb_abi_version",
    "shb_mesh_create", "shb_mesh_free", "shb_mesh_transform", "shb_batch_create_on", "shb_section", "shb_ray_cast",
    "shb_groove_features", "shb_groove_points", "shb_neck_image", "shb_forest_create", "shb_forest_predict", "shb_forest_free",
    "shb_mesh_from_stl", "shb_mesh_read", "shb_groove_theta", "shb_host_alloc", "shb_host_free", "shb_landmark_front",
    "shb_landmark_wait",
)


class BackendError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library (no device needed for this step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise BackendError(
            f"{LIB_PATH} is missing: build it with `python -m shoulder_b200.build` (needs nvcc). "
            "shoulder_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    p, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    pp = C.POINTER(C.c_void_p)
    lib.shb_init.argtypes = [C.c_int]
    lib.shb_set_stream.argtypes = [p]
    lib.shb_batch_create.argtypes = [i32, p, p, p, p, i32, p, p, p, p, p, pp]
    lib.shb_batch_free.argtypes = [p]
    lib.shb_batch_run.argtypes = [p, u32, i32, pp]
    lib.shb_batch_run_req.argtypes = [p, p, u32, i32, pp]
    lib.shb_result_window.argtypes = [p, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.shb_sweep_batch.argtypes = [i32, p, p, p, p, i32, p, p, p, p, p, u32, i32, pp]
    lib.shb_result_fetch.argtypes = [p, u32]
    lib.shb_result_fetch_async.argtypes = [p, u32]
    lib.shb_result_array.argtypes = [p, i32, i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.shb_result_array.restype = C.c_void_p
    lib.shb_result_totals.argtypes = [p, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.shb_result_free.argtypes = [p]
    lib.shb_mesh_create.argtypes = [p, i64, p, i64, pp]
    lib.shb_mesh_free.argtypes = [p]
    lib.shb_mesh_transform.argtypes = [p, p, pp]
    lib.shb_batch_create_on.argtypes = [p, i32, p, p, p, p, pp]
    lib.shb_section.argtypes = [p, p, p, u32, p, pp]
    lib.shb_ray_cast.argtypes = [p, i32, p, p, i32, p, p, p, p, C.POINTER(i32)]
    lib.shb_groove_features.argtypes = [p, i32, p, p, p, p, p, p, p]
    lib.shb_groove_points.argtypes = [p, i32, p, p, i32, p, p, p]
    lib.shb_neck_image.argtypes = [p, i32, p, p, p, p, p]
    lib.shb_forest_create.argtypes = [i32, i32, i32, p, p, p, p, p, p, pp]
    lib.shb_forest_predict.argtypes = [p, p, i32, p]
    lib.shb_forest_free.argtypes = [p]
    lib.shb_landmark_front.argtypes = [p, p]
    lib.shb_landmark_wait.argtypes = [p]
    lib.shb_profile_enable.argtypes = [C.c_int]
    lib.shb_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(i64), C.c_int]
    lib.shb_launch_count.restype = i64
    lib.shb_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("shb_result_array", "shb_launch_count", "shb_last_error"):
            getattr(lib, name).restype = C.c_int
    if lib.shb_abi_version() != ABI_VERSION:
        raise BackendError("libshoulder_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class _PinnedBlock:
    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(load().shb_host_alloc(int(nbytes), C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            if self.ptr:
                load().shb_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """Uninitialised numpy array in page-locked memory of the library's cache (``shb_host_alloc``): the output buffers of
    the feature / mesh calls, so that their device->host copies run at PCIe speed.  Returned to the cache when the array
    (and every view of it) is gone."""
    init(_inited if _inited is not None else 0)
    dt = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    blk = _PinnedBlock(max(n, 16))
    buf = (C.c_char * max(n, 1)).from_address(blk.ptr)
    buf._blk = blk                                  # the numpy array keeps `buf` as its base, `buf` keeps the block
    return np.frombuffer(buf, dtype=dt, count=n // dt.itemsize).reshape(shape)


def check(rc: int) -> None:
    if rc != 0:
        raise BackendError(f"shoulder_b200 error {rc}: {load().shb_last_error().decode()}")


_inited = None


def init(device: int = 0) -> None:
    """Bring the device up.  Raises if there is no sm_100 GPU — nothing runs on the CPU."""
    global _inited
    if _inited == device:
        return
    check(load().shb_init(int(device)))
    _inited = device


def set_stream(stream_ptr: int | None) -> None:
    check(load().shb_set_stream(C.c_void_p(stream_ptr or 0)))


def profile_enable(on, stages=None) -> None:
    """on: bool.  stages: optional iterable of stage names (STAGE_NAMES) to time; default every stage."""
    if not on:
        code = 0
    elif stages is None:
        code = 1
    else:
        code = sum(2 << STAGE_NAMES.index(s) for s in stages)
    check(load().shb_profile_enable(code))


def profile_read(reset: bool = True):
    ms = (C.c_double * N_STAGES)()
    n = (C.c_int64 * N_STAGES)()
    check(load().shb_profile_read(ms, n, 1 if reset else 0))
    return {STAGE_NAMES[i]: (ms[i], n[i]) for i in range(N_STAGES)}


def trim() -> None:
    """Give cached device memory and idle pinned buffers back (``shb_trim``)."""
    check(load().shb_trim())


def launch_count() -> int:
    return int(load().shb_launch_count())


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class SweepRequest(C.Structure):
    """``shb_sweep_request``: the windowed outputs one sweep wants and the plane rows they cover."""
    _fields_ = [("outputs_mask", C.c_uint32), ("row_lo", C.c_int32 * N_WINDOWED), ("row_hi", C.c_int32 * N_WINDOWED)]


def make_requests(specs):
    """specs: per sweep either an int mask (all rows) or a dict {OUT_* bit: (row_lo, row_hi)}."""
    arr = (SweepRequest * len(specs))()
    for k, spec in enumerate(specs):
        if isinstance(spec, dict):
            arr[k].outputs_mask = 0
            for a, bit in enumerate(WINDOWED_BITS):
                arr[k].row_lo[a], arr[k].row_hi[a] = 0, -1
                if bit in spec:
                    lo, hi = spec[bit]
                    arr[k].outputs_mask |= bit
                    arr[k].row_lo[a], arr[k].row_hi[a] = int(lo), int(hi)
        else:
            arr[k].outputs_mask = int(spec)
            for a in range(N_WINDOWED):
                arr[k].row_lo[a], arr[k].row_hi[a] = 0, -1
    return arr


class SweepResult:
    """Owner of one ``shb_result``.  Arrays are numpy views of backend-owned pinned memory and
    stay valid while this object is alive."""

    def __init__(self, handle, n_sweep: int, keepalive=None):
        self._h = handle
        self.n_sweep = n_sweep
        self._keep = keepalive

    def array(self, which: int, sweep: int = 0) -> np.ndarray:
        shape = (C.c_int64 * 4)()
        ndim, dt = C.c_int32(), C.c_int32()
        ptr = load().shb_result_array(self._h, which, sweep, shape, C.byref(ndim), C.byref(dt))
        if not ptr:
            raise BackendError(load().shb_last_error().decode())
        shp = tuple(int(shape[i]) for i in range(ndim.value))
        dtype = np.dtype(_DTYPES[dt.value])
        n = int(np.prod(shp)) if shp else 1
        if n == 0:
            return np.zeros(shp, dtype=dtype)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
        buf._shb_owner = self            # arr.base -> buf -> self keeps the pinned memory alive
        return np.frombuffer(buf, dtype=dtype).reshape(shp)   # a view, like the reference's cached arrays

    def array_shape(self, which: int, sweep: int = 0):
        return self.array(which, sweep).shape

    def window(self, which: int, sweep: int = 0):
        """(row_lo, row_hi) of the sweep's planes that array ``which`` covers in this result."""
        lo, hi = C.c_int32(), C.c_int32()
        check(load().shb_result_window(self._h, which, sweep, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def fetch(self, mask: int) -> None:
        check(load().shb_result_fetch(self._h, mask))

    def fetch_async(self, mask: int) -> None:
        """Enqueue the device->host copies and return; any later access waits for them."""
        check(load().shb_result_fetch_async(self._h, mask & ~OUT_CONTOURS))

    def totals(self):
        v = [C.c_int64() for _ in range(4)]
        check(load().shb_result_totals(self._h, *[C.byref(x) for x in v]))
        return {"planes": v[0].value, "segments": v[1].value, "contours": v[2].value, "points": v[3].value}

    def close(self):
        if self._h:
            load().shb_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pack(meshes, sweeps):
    """meshes: list of (vertices (V,3) f64, faces (T,3) i64); sweeps: list of
    (mesh_index, z_orig, heights, interp_num)."""
    verts = np.ascontiguousarray(np.concatenate([np.asarray(m[0], dtype=np.float64).reshape(-1, 3) for m in meshes]))
    faces = np.ascontiguousarray(np.concatenate([np.asarray(m[1], dtype=np.int64).reshape(-1, 3) for m in meshes]))
    vert_off = np.zeros(len(meshes) + 1, dtype=np.int64)
    face_off = np.zeros(len(meshes) + 1, dtype=np.int64)
    vert_off[1:] = np.cumsum([len(m[0]) for m in meshes])
    face_off[1:] = np.cumsum([len(m[1]) for m in meshes])
    sweep_mesh = np.array([s[0] for s in sweeps], dtype=np.int32)
    z_orig = np.array([s[1] for s in sweeps], dtype=np.float64)
    hs = [np.asarray(s[2], dtype=np.float64).reshape(-1) for s in sweeps]
    heights = np.ascontiguousarray(np.concatenate(hs)) if hs else np.zeros(0)
    height_off = np.zeros(len(sweeps) + 1, dtype=np.int64)
    height_off[1:] = np.cumsum([len(h) for h in hs])
    interp = np.array([s[3] for s in sweeps], dtype=np.int32)
    return verts, vert_off, faces, face_off, sweep_mesh, z_orig, heights, height_off, interp


class SweepBatch:
    """HBM-resident meshes + sweeps (``shb_batch_create``); ``run`` enqueues the hot path."""

    def __init__(self, meshes, sweeps, packed=None):
        init(_inited if _inited is not None else 0)
        a = packed if packed is not None else _pack(meshes, sweeps)
        self.n_sweep = len(a[4])
        self._inputs = a                  # the upload is asynchronous: keep the host arrays alive
        h = C.c_void_p()
        check(load().shb_batch_create(len(a[1]) - 1, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), self.n_sweep,
                                      _ptr(a[4]), _ptr(a[5]), _ptr(a[6]), _ptr(a[7]), _ptr(a[8]), C.byref(h)))
        self._h = h

    def run(self, outputs_mask: int, n_angles: int = 0, requests=None) -> SweepResult:
        """``requests``: optional per-sweep list (see :func:`make_requests`) -> ``shb_batch_run_req``."""
        r = C.c_void_p()
        if requests is None:
            check(load().shb_batch_run(self._h, outputs_mask, n_angles, C.byref(r)))
        else:
            assert len(requests) == self.n_sweep
            req = requests if isinstance(requests, C.Array) else make_requests(requests)
            check(load().shb_batch_run_req(self._h, C.cast(req, C.c_void_p), outputs_mask, n_angles, C.byref(r)))
        return SweepResult(r, self.n_sweep, keepalive=self)

    def close(self):
        if self._h:
            load().shb_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sweep_batch(meshes, sweeps, outputs_mask: int, n_angles: int = 0, packed=None, lazy: bool = False, requests=None) -> SweepResult:
    """One-call host-to-host form (``shb_sweep_batch``).  ``lazy=True`` computes everything in
    ``outputs_mask`` on the device but copies an array to the host only when it is first asked for
    (``shb_batch_create`` + ``shb_batch_run``; the inputs are released right after the kernels are enqueued)."""
    init(_inited if _inited is not None else 0)
    a = packed if packed is not None else _pack(meshes, sweeps)
    if lazy or requests is not None:
        batch = SweepBatch(None, None, packed=a)
        res = batch.run(outputs_mask, n_angles, requests)
        res._keep = None
        batch.close()                     # stream ordered: the enqueued kernels still see the inputs
        if not lazy:
            res.fetch(outputs_mask | OUT_PLANE)
        return res
    r = C.c_void_p()
    check(load().shb_sweep_batch(len(a[1]) - 1, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), len(a[4]), _ptr(a[4]),
                                 _ptr(a[5]), _ptr(a[6]), _ptr(a[7]), _ptr(a[8]), outputs_mask, n_angles, C.byref(r)))
    return SweepResult(r, len(a[4]))


class PipelinedResult:
    """Results of :func:`sweep_batch_pipelined`: one ``SweepResult`` per chunk, addressed by global sweep index."""

    def __init__(self, parts, first_sweep):
        self.parts, self._first = parts, first_sweep
        self.n_sweep = first_sweep[-1]

    def _locate(self, sweep: int):
        c = int(np.searchsorted(self._first, sweep, side="right")) - 1
        return self.parts[c], sweep - int(self._first[c])

    def array(self, which: int, sweep: int = 0) -> np.ndarray:
        part, k = self._locate(sweep)
        return part.array(which, k)

    def totals(self):
        out = {}
        for p in self.parts:
            for k, v in p.totals().items():
                out[k] = out.get(k, 0) + v
        return out

    def close(self):
        for p in self.parts:
            p.close()


def split_packed(packed, n_chunks: int):
    """Cuts a packed batch (see :func:`_pack`) into ``n_chunks`` batches of whole meshes.  Sweeps must be grouped
    by mesh in ascending mesh order (what :func:`_pack` produces from per-bone lists)."""
    verts, voff, faces, foff, smesh, zo, hts, hoff, interp = packed
    n_mesh = len(voff) - 1
    n_chunks = max(1, min(n_chunks, n_mesh))
    assert (np.diff(smesh) >= 0).all(), "sweeps must be grouped by mesh"
    cuts = np.linspace(0, n_mesh, n_chunks + 1).astype(np.int64)
    chunks, first = [], [0]
    for a, b in zip(cuts[:-1], cuts[1:]):
        s0, s1 = int(np.searchsorted(smesh, a, side="left")), int(np.searchsorted(smesh, b, side="left"))
        chunks.append((verts[voff[a]:voff[b]], voff[a:b + 1] - voff[a], faces[foff[a]:foff[b]], foff[a:b + 1] - foff[a],
                       (smesh[s0:s1] - a).astype(np.int32), zo[s0:s1], hts[hoff[s0]:hoff[s1]], hoff[s0:s1 + 1] - hoff[s0],
                       interp[s0:s1]))
        first.append(s1)
    return chunks, np.array(first, dtype=np.int64)


def sweep_batch_pipelined(chunks, first_sweep, outputs_mask: int, n_angles: int = 0, requests=None) -> PipelinedResult:
    """Host-to-host call for a batch pre-cut into chunks (:func:`split_packed`): chunk i+1 is uploaded and
    computed while chunk i's outputs travel back over PCIe (the library copies on its own stream).  The
    device->host transfer dominates a large batch (16 bytes per sample), so hiding everything else behind
    it is the whole gain."""
    init(_inited if _inited is not None else 0)
    parts = []
    for k, c in enumerate(chunks):
        batch = SweepBatch(None, None, packed=c)
        req = None if requests is None else requests[int(first_sweep[k]):int(first_sweep[k + 1])]
        res = batch.run(outputs_mask, n_angles, req)
        res._keep = None
        batch.close()
        res.fetch_async(outputs_mask)        # copy stream: starts as soon as this group's kernels finish
        parts.append(res)
    for res in parts:
        res.fetch(outputs_mask)              # wait (and fetch what the async form does not cover)
    return PipelinedResult(parts, first_sweep)
