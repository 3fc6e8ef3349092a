// Host side of the C ABI declared in include/shoulder_b200.h: device bring-up, batch upload,
// the launch sequence of the hot path, lazy device->host fetch of results.
// Replaces the host loop of reference src/shoulder/humerus/slice.py:21-147 (one Python call into
// trimesh + per-plane Python loops) with one enqueue of K1..K4 per batch.
#include "shb_common.cuh"
#include "../../include/shoulder_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <unistd.h>
#include <cstring>
#include <memory>
#include <type_traits>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(SHB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct PinnedBuf { void* p; size_t size; bool used; };

struct Ctx {
    std::recursive_mutex mu;     // every entry point takes it: one caller at a time per process
    bool inited = false;
    int device = 0, n_sm = 148;
    size_t smem_optin = 0;
    cudaStream_t own = nullptr, stream = nullptr;
    cudaStream_t copy = nullptr;           // device->host copies run here so they overlap the next batch's kernels
    cudaStream_t aux = nullptr;            // the declined planes of the first part of a stitch launch run here, beside the bulk
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;
    cudaEvent_t sized = nullptr;           // recorded behind the size publication of a run
    int64_t launches = 0;
    uint32_t profile = 0;                  // bit s: stage s is timed
    double stage_ms[SHB_N_STAGES] = {};
    int64_t stage_launches[SHB_N_STAGES] = {};
    struct Pending { cudaEvent_t a, b; int stage; int n; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> ev_free;
    std::vector<PinnedBuf> pinned;
    uint32_t* h_totals = nullptr;          // pinned readback slot
    unsigned long long* h_totals64 = nullptr;
    double2* angle_tab = nullptr; int angle_n = 0;
    size_t pinned_limit = (size_t)4 << 30;    // bytes of pinned host cache kept across calls: min(16 GiB, RAM / 4) at init, SHB_PINNED_LIMIT_MB
    cudaMemPool_t pool = nullptr;             // the library's OWN stream-ordered pool (the device's default pool is left alone)
} g;

// CUDA's current device is per host thread: every entry point switches to the library's device for its duration, so a
// call from a thread that never saw shb_init (or that works on another GPU) still launches where the streams live
struct DeviceGuard {
    int prev = -1; bool switched = false;
    DeviceGuard() {
        if (!g.inited) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != g.device) switched = cudaSetDevice(g.device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define SHB_ENTER std::lock_guard<std::recursive_mutex> lk(g.mu); DeviceGuard dg_

void* pinned_get(size_t bytes) {
    if (bytes == 0) bytes = 16;
    PinnedBuf* best = nullptr;
    for (auto& b : g.pinned)
        if (!b.used && b.size >= bytes && (!best || b.size < best->size)) best = &b;
    if (best && best->size <= 2 * bytes + (1 << 20)) { best->used = true; return best->p; }
    // keep the cache bounded: beyond the limit, idle buffers are returned to the OS before a new one is pinned
    size_t total = 0;
    for (auto& b : g.pinned) total += b.size;
    if (total + bytes > g.pinned_limit) {
        for (size_t i = 0; i < g.pinned.size();) {
            if (!g.pinned[i].used) { cudaFreeHost(g.pinned[i].p); g.pinned[i] = g.pinned.back(); g.pinned.pop_back(); }
            else ++i;
        }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    g.pinned.push_back({p, bytes, true});
    return p;
}
void pinned_put(void* p) {
    for (auto& b : g.pinned) if (b.p == p) { b.used = false; return; }
}

// host->device copy of a library-owned (pageable) array through a pinned staging buffer: a pageable
// cudaMemcpyAsync serialises with transfers in flight on other streams, which would stall the pipelined call
struct Staging {
    std::vector<void*> bufs;
    cudaError_t copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
        if (bytes == 0) return cudaSuccess;
        void* p = pinned_get(bytes);
        if (!p) return cudaErrorMemoryAllocation;
        std::memcpy(p, src, bytes);
        bufs.push_back(p);
        return cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, st);
    }
    void release() { for (void* p : bufs) pinned_put(p); bufs.clear(); }      // call after the stream has been synchronised
    ~Staging() { release(); }
};

template <class T> cudaError_t dalloc(T** p, size_t n, cudaStream_t st) {
    return cudaMallocFromPoolAsync(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T), g.pool, st);
}
cudaError_t dalloc_bytes(void** p, size_t bytes, cudaStream_t st) {
    return cudaMallocFromPoolAsync(p, std::max<size_t>(bytes, 16), g.pool, st);
}
template <class T> void dfree(T*& p, cudaStream_t st) { if (p) { cudaFreeAsync((void*)p, st); p = nullptr; } }

struct StageTimer {
    int stage; cudaEvent_t a = nullptr, b = nullptr; bool on; cudaStream_t st;
    StageTimer(int s, cudaStream_t stream) : stage(s), on((g.profile >> s) & 1u), st(stream) {
        if (!on) return;
        auto get = [] { cudaEvent_t e; if (!g.ev_free.empty()) { e = g.ev_free.back(); g.ev_free.pop_back(); } else cudaEventCreate(&e); return e; };
        a = get(); b = get();
        cudaEventRecord(a, st);
    }
    void stop(int n_launch) {
        g.launches += n_launch;
        if (!on) return;
        cudaEventRecord(b, st);
        g.pending.push_back({a, b, stage, n_launch});
    }
};

}  // namespace

// a mesh that stays in HBM between calls (shb_mesh_create): K0 form of the vertices and faces + the face adjacency.
// Batches made from it (shb_batch_create_on) borrow the arrays; the device memory goes when the last user is gone.
struct shb_mesh {
    double4* vert = nullptr; double* vz = nullptr; int4* face = nullptr; uint32_t* adj = nullptr; uint32_t* d_bad = nullptr;
    int64_t nv = 0, nf = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr, last_use = nullptr;
    int refs = 1;                          // the handle itself + every batch that borrows it
    double frame[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // to_3D of a mesh made by shb_mesh_transform (frame -> source mesh)
};

namespace { void mesh_unref(shb_mesh* m); }

struct shb_batch {
    shb_mesh* shared = nullptr;            // vert / vz / face / adj / d_bad belong to this mesh, not to the batch
    int32_t n_mesh = 0, n_sweep = 0;
    int64_t n_vert = 0, n_face = 0;
    uint32_t G = 0, n_item = 0, max_interp = 0, max_faces = 0;
    uint32_t* adj = nullptr;               // [T][4] face adjacency (built once per batch by K0b)
    std::vector<ShbSweep> sweeps;          // host copy
    double4* vert = nullptr; double* vz = nullptr; int4* face = nullptr;
    ShbSweep* d_sweep = nullptr; uint32_t* d_item_off = nullptr;
    double* h_sorted = nullptr; double* h_orig = nullptr; double* oz = nullptr;
    uint32_t *plane_out = nullptr, *plane_in = nullptr, *plane_sweep = nullptr;
    uint32_t* stitch_order = nullptr;      // planes nearest the ends of their sweep first (see shb_batch_create)
    uint32_t* d_bad = nullptr;             // set by K0 when a face names a vertex outside its mesh; checked at the first run
    Staging* stage = nullptr;              // pinned copies of the library-built arrays, held until the upload has executed
    bool checked = false;
    // Stream ownership: the batch lives on the stream it was created on.  A run on another stream (shb_set_stream in
    // between) waits for `uploaded`; the batch's memory is freed on its own stream behind `last_use` (the last run).
    cudaStream_t stream = nullptr;
    cudaEvent_t uploaded = nullptr, last_use = nullptr;
    // compact launch list of the resample kernel for the last request that excluded planes (per-sweep windows)
    uint32_t* rs_order = nullptr; uint32_t n_rs = 0; uint64_t rs_key = 0;
};

struct shb_result {
    const shb_batch* batch = nullptr;
    std::vector<ShbSweep> sweeps;
    uint32_t G = 0, mask = 0, n_angles = 0;
    uint64_t arr_total[SHB_N_ARR] = {};     // elements of each windowed output array (profiles: rows x 2 x N summed over sweeps)
    cudaStream_t stream = nullptr;          // the stream the result was computed on; its device memory is freed there
    uint32_t W = 0;                         // capacity of the per-segment arrays
    ShbDev d = {};                          // device pointers owned by the result
    // device memory of a run comes in four blocks carved by offset (not ~45 pool calls): scratch sized by the batch and
    // scratch sized by the segment count (both released when the run is enqueued), and the two blocks the result keeps
    void *blk_keep_a = nullptr, *blk_keep_b = nullptr, *blk_tmp_a = nullptr, *blk_tmp_b = nullptr;
    cudaEvent_t done = nullptr;             // recorded on the compute stream when the result's kernels are enqueued
    uint32_t* d_ct_off = nullptr; uint32_t* d_pt_off = nullptr;
    double* d_pts_c = nullptr; int64_t* d_ctpt_c = nullptr; double* d_ctarea_c = nullptr;
    // host (pinned) copies, filled lazily
    bool have_plane = false, have_seg = false, have_cont = false;
    bool pending = false;                   // copies enqueued on the copy stream, not yet waited for
    uint32_t S = 0, n_cont = 0, n_pts = 0;
    int32_t *h_nseg = nullptr, *h_nent = nullptr, *h_sel = nullptr, *h_face_index = nullptr;
    uint32_t *h_status = nullptr, *h_seg_off = nullptr, *h_ct_off = nullptr, *h_pt_off = nullptr;
    double *h_bounds = nullptr, *h_centroid = nullptr, *h_area1 = nullptr, *h_segments = nullptr;
    double *h_pts = nullptr, *h_ctarea = nullptr; int64_t* h_ctpt = nullptr;
    void* h_arr[SHB_N_ARR] = {};            // six profile arrays + the radius image
    size_t esz = 8;                         // bytes per profile / radius element
    std::vector<std::vector<int64_t>> rel;  // per-sweep relative offset arrays handed out
    std::vector<void*> lf_staged;           // pinned staging buffers of shb_landmark_front(SHB_LF_NO_WAIT) calls in flight
};

namespace {
void mesh_unref(shb_mesh* m) {
    if (--m->refs > 0) return;
    cudaStream_t st = m->stream ? m->stream : g.stream;
    if (m->last_use) { cudaStreamWaitEvent(st, m->last_use, 0); cudaEventDestroy(m->last_use); }
    if (m->ready) cudaEventDestroy(m->ready);
    dfree(m->vert, st); dfree(m->vz, st); dfree(m->face, st); dfree(m->adj, st); dfree(m->d_bad, st);
    delete m;
}
}  // namespace

extern "C" {

static int batch_run_impl(shb_batch* b, const shb_sweep_request* req, uint32_t outputs_mask, int32_t n_angles, uint32_t flags, shb_result** out);

SHB_API const char* shb_last_error(void) { return g_err.c_str(); }
SHB_API int shb_abi_version(void) { return SHB_ABI_VERSION; }
SHB_API int64_t shb_launch_count(void) { return g.launches; }

SHB_API int shb_init(int device) {
    SHB_ENTER;
    if (g.inited) {
        if (device != g.device) return fail(SHB_E_STATE, "already initialised on device %d", g.device);
        return SHB_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SHB_E_CUDA, "no CUDA device (%s); this backend has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SHB_E_INVALID, "device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SHB_E_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    g.device = device;
    g.n_sm = prop.multiProcessorCount;
    g.smem_optin = prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&g.own, cudaStreamNonBlocking));
    g.stream = g.own;
    CK(cudaStreamCreateWithFlags(&g.copy, cudaStreamNonBlocking));
    {   // highest priority: its few CTAs (declined planes) must get SM slots while the bulk launch still has thousands queued
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&g.aux, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_mid, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.sized, cudaEventDisableTiming));
    // a private pool: its "never release on its own" threshold must not leak into other users of cudaMallocAsync
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CK(cudaMemPoolCreate(&g.pool, &props));
    uint64_t thr = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(g.pool, cudaMemPoolAttrReleaseThreshold, &thr));
    {   // page-locked host cache: a quarter of the host's RAM, at most 16 GiB
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        if (pages > 0 && psz > 0) g.pinned_limit = std::min<size_t>((size_t)16 << 30, (size_t)pages * (size_t)psz / 4);
    }
    if (const char* lim = getenv("SHB_PINNED_LIMIT_MB")) g.pinned_limit = (size_t)atoll(lim) << 20;
    CK(cudaHostAlloc(&g.h_totals, 8 * sizeof(uint32_t), cudaHostAllocMapped));       // written by k_publish, read after a stream sync
    CK(cudaHostAlloc(&g.h_totals64, 2 * sizeof(unsigned long long), cudaHostAllocMapped));
    g.inited = true;
    return SHB_OK;
}

SHB_API int shb_set_stream(void* cuda_stream) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own;
    return SHB_OK;
}

SHB_API int shb_trim(void) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaStreamSynchronize(g.copy));
    CK(cudaMemPoolTrimTo(g.pool, 0));
    for (size_t i = 0; i < g.pinned.size();) {
        if (!g.pinned[i].used) { cudaFreeHost(g.pinned[i].p); g.pinned[i] = g.pinned.back(); g.pinned.pop_back(); }
        else ++i;
    }
    return SHB_OK;
}

SHB_API int shb_host_alloc(int64_t bytes, void** out) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!out || bytes < 0) return fail(SHB_E_INVALID, "bad argument");
    *out = pinned_get((size_t)bytes);
    if (!*out) return fail(SHB_E_CUDA, "page-locked allocation of %lld bytes failed", (long long)bytes);
    return SHB_OK;
}
SHB_API int shb_host_free(void* p) {
    SHB_ENTER;
    if (p) pinned_put(p);
    return SHB_OK;
}

SHB_API int shb_profile_enable(int on) {
    SHB_ENTER;
    g.profile = on == 0 ? 0u : (on == 1 ? (1u << SHB_N_STAGES) - 1u : ((uint32_t)on >> 1) & ((1u << SHB_N_STAGES) - 1u));
    return SHB_OK;
}

SHB_API int shb_profile_read(double stage_ms[SHB_N_STAGES], int64_t stage_launches[SHB_N_STAGES], int reset) {
    SHB_ENTER;
    for (auto& p : g.pending) {
        CK(cudaEventSynchronize(p.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, p.a, p.b));
        g.stage_ms[p.stage] += ms;
        g.stage_launches[p.stage] += p.n;
        g.ev_free.push_back(p.a); g.ev_free.push_back(p.b);
    }
    g.pending.clear();
    for (int i = 0; i < SHB_N_STAGES; ++i) {
        if (stage_ms) stage_ms[i] = g.stage_ms[i];
        if (stage_launches) stage_launches[i] = g.stage_launches[i];
        if (reset) { g.stage_ms[i] = 0; g.stage_launches[i] = 0; }
    }
    return SHB_OK;
}

SHB_API int shb_batch_free(shb_batch* b) {
    SHB_ENTER;
    if (!b) return SHB_OK;
    cudaStream_t st = b->stream ? b->stream : g.stream;
    if (b->stage) { cudaStreamSynchronize(st); delete b->stage; b->stage = nullptr; }
    if (b->last_use) { cudaStreamWaitEvent(st, b->last_use, 0); cudaEventDestroy(b->last_use); }     // runs on other streams still read the batch
    if (b->uploaded) cudaEventDestroy(b->uploaded);
    if (b->shared) {
        shb_mesh* m = b->shared;
        if (b->last_use || true) {          // later frees of the mesh must wait for this batch's last run
            if (!m->last_use) cudaEventCreateWithFlags(&m->last_use, cudaEventDisableTiming);
            cudaEventRecord(m->last_use, st);
        }
        b->vert = nullptr; b->vz = nullptr; b->face = nullptr; b->adj = nullptr; b->d_bad = nullptr;
        mesh_unref(m);
    }
    dfree(b->d_bad, st);
    dfree(b->adj, st);
    dfree(b->vert, st); dfree(b->vz, st); dfree(b->face, st); dfree(b->d_sweep, st); dfree(b->d_item_off, st);
    dfree(b->h_sorted, st); dfree(b->h_orig, st); dfree(b->oz, st); dfree(b->plane_out, st); dfree(b->plane_in, st); dfree(b->plane_sweep, st); dfree(b->stitch_order, st);
    dfree(b->rs_order, st);
    delete b;
    return SHB_OK;
}

// K0b: face adjacency of a set of faces (one hash table over its undirected edges; temporary)
static int build_adjacency(const int4* face, int64_t nf, uint32_t** adj_out, cudaStream_t st) {
    CK(dalloc(adj_out, 4 * (size_t)std::max<int64_t>(nf, 1), st));
    if (nf) {
        uint32_t hs = 1024;
        while ((uint64_t)hs < 3ull * (uint64_t)nf) hs <<= 1;              // 1.5 T edges -> load <= 0.5
        unsigned long long* keys = nullptr; uint32_t *cnt = nullptr, *own = nullptr, *hslot = nullptr;
        CK(dalloc(&keys, hs, st)); CK(dalloc(&cnt, hs, st)); CK(dalloc(&own, 2 * (size_t)hs, st)); CK(dalloc(&hslot, 3 * (size_t)nf, st));
        CK(cudaMemsetAsync(keys, 0xFF, (size_t)hs * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(cnt, 0, (size_t)hs * sizeof(uint32_t), st));
        g.launches += shb_launch_adjacency(face, nf, keys, cnt, own, hslot, hs, *adj_out, st);
        CK(cudaGetLastError());
        dfree(keys, st); dfree(cnt, st); dfree(own, st); dfree(hslot, st);
    }
    return SHB_OK;
}

static int batch_create_impl(int32_t n_mesh, const double* verts, const int64_t* vert_off, const int64_t* faces,
                             const int64_t* face_off, int32_t n_sweep, const int32_t* sweep_mesh, const double* z_orig,
                             const double* heights, const int64_t* height_off, const int32_t* interp_num, shb_mesh* shared, shb_batch** out) {
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!out) return fail(SHB_E_INVALID, "out is null");
    *out = nullptr;
    if (n_mesh <= 0 || n_sweep <= 0 || (!shared && (!verts || !faces)) || !vert_off || !face_off || !sweep_mesh || !z_orig || !heights ||
        !height_off || !interp_num)
        return fail(SHB_E_INVALID, "null or empty input");
    if (vert_off[0] != 0 || face_off[0] != 0 || height_off[0] != 0) return fail(SHB_E_INVALID, "offset arrays must start at 0");
    for (int m = 0; m < n_mesh; ++m)
        if (vert_off[m + 1] < vert_off[m] || face_off[m + 1] < face_off[m]) return fail(SHB_E_INVALID, "offsets of mesh %d decrease", m);
    const int64_t nv = vert_off[n_mesh], nf = face_off[n_mesh], G64 = height_off[n_sweep];
    if (nv >= (int64_t)1 << 31 || nf >= (int64_t)1 << 29) return fail(SHB_E_CAPACITY, "too many vertices/faces in one batch");
    if (G64 <= 0 || G64 >= (int64_t)1 << 31) return fail(SHB_E_CAPACITY, "plane count %lld out of range", (long long)G64);

    struct BatchDel { void operator()(shb_batch* p) const { shb_batch_free(p); } };
    std::unique_ptr<shb_batch, BatchDel> b(new shb_batch);
    b->n_mesh = n_mesh; b->n_sweep = n_sweep; b->n_vert = nv; b->n_face = nf; b->G = (uint32_t)G64;
    b->sweeps.resize(n_sweep);
    std::vector<uint32_t> item_off(n_sweep + 1, 0);
    std::vector<double> hs(G64), ho(heights, heights + G64), oz(G64);
    std::vector<uint32_t> pout(G64), pin(G64), psw(G64);
    uint64_t items = 0;
    std::vector<uint32_t> idx;
    for (int s = 0; s < n_sweep; ++s) {
        const int m = sweep_mesh[s];
        const int64_t p0 = height_off[s], P = height_off[s + 1] - p0;
        if (m < 0 || m >= n_mesh) return fail(SHB_E_INVALID, "sweep %d names mesh %d", s, m);
        if (P < 0) return fail(SHB_E_INVALID, "height offsets of sweep %d decrease", s);
        if (interp_num[s] < 2) return fail(SHB_E_INVALID, "interp_num of sweep %d must be >= 2", s);
        ShbSweep& sw = b->sweeps[s];
        sw = ShbSweep{};
        sw.z_orig = z_orig[s];
        sw.plane_off = (uint32_t)p0; sw.n_plane = (uint32_t)P;
        sw.face_off = (uint32_t)face_off[m]; sw.n_face = (uint32_t)(face_off[m + 1] - face_off[m]);
        sw.item_off = (uint32_t)items; sw.interp_num = (uint32_t)interp_num[s]; sw.mesh = (uint32_t)m; sw.pad = 0;
        item_off[s] = (uint32_t)items;
        items += sw.n_face;
        b->max_interp = std::max<uint32_t>(b->max_interp, (uint32_t)interp_num[s]);
        // planes ascending in height for the range search; linspace inputs are already monotone
        const double* h = heights + p0;
        bool asc = true, desc = true;
        for (int64_t i = 1; i < P; ++i) { asc &= h[i] >= h[i - 1]; desc &= h[i] <= h[i - 1]; }
        idx.resize(P);
        if (asc) std::iota(idx.begin(), idx.end(), 0u);
        else if (desc) for (int64_t i = 0; i < P; ++i) idx[i] = (uint32_t)(P - 1 - i);
        else {
            std::iota(idx.begin(), idx.end(), 0u);
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t c) { return h[a] < h[c]; });
        }
        for (int64_t i = 0; i < P; ++i) oz[p0 + i] = z_orig[s] + h[i];        // one rounding, like numpy's origin + normal * height
        for (int64_t i = 0; i < P; ++i) {
            if (!(h[idx[i]] == h[idx[i]])) return fail(SHB_E_INVALID, "NaN height in sweep %d", s);
            hs[p0 + i] = h[idx[i]];
            pout[p0 + i] = (uint32_t)(p0 + idx[i]);
            pin[p0 + idx[i]] = (uint32_t)(p0 + i);
            psw[p0 + i] = (uint32_t)s;
        }
    }
    if (items >= (uint64_t)1 << 32) return fail(SHB_E_CAPACITY, "sum of faces over sweeps = %llu needs a split", (unsigned long long)items);
    item_off[n_sweep] = (uint32_t)items;
    // Launch order of the stitcher: the planes nearest the two ends of their sweep first.  Sections with several
    // contours (several times the work of a plain one) sit at the ends of a bone; started first, they are hidden behind
    // the bulk of the launch instead of extending its tail.
    std::vector<uint32_t> order(G64);
    {
        std::vector<double> endness(G64);
        for (int s = 0; s < n_sweep; ++s) {
            const int64_t p0 = height_off[s], P = height_off[s + 1] - p0;
            for (int64_t i = 0; i < P; ++i) {
                const int64_t r = (int64_t)pin[p0 + i] - p0;           // rank of the plane along the sweep axis
                endness[p0 + i] = (double)std::min(r, P - 1 - r) / (double)std::max<int64_t>(P, 1);
            }
        }
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return endness[a] < endness[c]; });
    }
    b->n_item = (uint32_t)items;

    cudaStream_t st = g.stream;
    b->stream = st;
    const bool dbg = getenv("SHB_DEBUG_TIMING") != nullptr;
    auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    double T0 = now();
    double* raw_v = nullptr; int64_t* raw_f = nullptr; int64_t *d_voff = nullptr, *d_foff = nullptr;
    if (!shared) {
        CK(dalloc(&raw_v, 3 * (size_t)nv, st)); CK(dalloc(&raw_f, 3 * (size_t)nf, st));
        CK(dalloc(&d_voff, n_mesh + 1, st)); CK(dalloc(&d_foff, n_mesh + 1, st)); CK(dalloc(&b->d_bad, 1, st));
        CK(dalloc(&b->vert, nv, st)); CK(dalloc(&b->vz, nv, st)); CK(dalloc(&b->face, nf, st));
    }
    CK(dalloc(&b->d_sweep, n_sweep, st)); CK(dalloc(&b->d_item_off, n_sweep + 1, st));
    CK(dalloc(&b->h_sorted, G64, st)); CK(dalloc(&b->h_orig, G64, st)); CK(dalloc(&b->oz, G64, st));
    CK(dalloc(&b->plane_out, G64, st)); CK(dalloc(&b->plane_in, G64, st)); CK(dalloc(&b->plane_sweep, G64, st));
    CK(dalloc(&b->stitch_order, G64, st));
    double T1 = now();
    if (!shared) {
        CK(cudaMemcpyAsync(raw_v, verts, 3 * (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(raw_f, faces, 3 * (size_t)nf * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_voff, vert_off, (n_mesh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_foff, face_off, (n_mesh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(b->d_bad, 0, sizeof(uint32_t), st));
    }
    b->stage = new Staging;
    Staging& stage = *b->stage;
    CK(stage.copy(b->d_sweep, b->sweeps.data(), n_sweep * sizeof(ShbSweep), st));
    CK(stage.copy(b->d_item_off, item_off.data(), (n_sweep + 1) * sizeof(uint32_t), st));
    CK(stage.copy(b->h_sorted, hs.data(), G64 * sizeof(double), st));
    CK(stage.copy(b->h_orig, ho.data(), G64 * sizeof(double), st));
    CK(stage.copy(b->oz, oz.data(), G64 * sizeof(double), st));
    CK(stage.copy(b->plane_out, pout.data(), G64 * sizeof(uint32_t), st));
    CK(stage.copy(b->plane_in, pin.data(), G64 * sizeof(uint32_t), st));
    CK(stage.copy(b->plane_sweep, psw.data(), G64 * sizeof(uint32_t), st));
    CK(stage.copy(b->stitch_order, order.data(), G64 * sizeof(uint32_t), st));
    double T2 = now();
    for (int m = 0; m < n_mesh; ++m) b->max_faces = std::max<uint32_t>(b->max_faces, (uint32_t)(face_off[m + 1] - face_off[m]));
    if (shared) {      // a resident mesh: borrow its arrays, behind the event of its upload
        if (shared->stream != st && shared->ready) CK(cudaStreamWaitEvent(st, shared->ready, 0));
        b->vert = shared->vert; b->vz = shared->vz; b->face = shared->face; b->adj = shared->adj; b->d_bad = shared->d_bad;
        b->shared = shared; ++shared->refs;
    } else {
    g.launches += shb_launch_prep_mesh(raw_v, raw_f, d_voff, d_foff, n_mesh, nv, nf, b->vert, b->vz, b->face, b->d_bad, st);
    CK(cudaGetLastError());
    dfree(raw_v, st); dfree(raw_f, st); dfree(d_voff, st); dfree(d_foff, st);
    { int rc_adj = build_adjacency(b->face, nf, &b->adj, st); if (rc_adj) return rc_adj; }
    }
    // no host synchronisation here: the upload and K0 are only enqueued.  verts / faces must stay valid until the
    // first shb_batch_run on this batch returns (it synchronises); the face-index range check is reported there.
    CK(cudaEventCreateWithFlags(&b->uploaded, cudaEventDisableTiming));
    CK(cudaEventRecord(b->uploaded, st));
    if (dbg) fprintf(stderr, "[shb] create: alloc %.3f  h2d-enqueue %.3f  launch+free %.3f ms\n", T1 - T0, T2 - T1, now() - T2);
    *out = b.release();
    return SHB_OK;
}

SHB_API int shb_batch_create(int32_t n_mesh, const double* verts, const int64_t* vert_off, const int64_t* faces,
                     const int64_t* face_off, int32_t n_sweep, const int32_t* sweep_mesh, const double* z_orig,
                     const double* heights, const int64_t* height_off, const int32_t* interp_num, shb_batch** out) {
    SHB_ENTER;
    return batch_create_impl(n_mesh, verts, vert_off, faces, face_off, n_sweep, sweep_mesh, z_orig, heights, height_off, interp_num, nullptr, out);
}

SHB_API int shb_mesh_create(const double* verts, int64_t n_vert, const int64_t* faces, int64_t n_face, shb_mesh** out) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!out || !verts || !faces || n_vert <= 0 || n_face < 0) return fail(SHB_E_INVALID, "bad argument");
    *out = nullptr;
    // the upload path of a one-mesh batch with a dummy sweep; its arrays move into the mesh handle
    const int64_t voff[2] = {0, n_vert}, foff[2] = {0, n_face}, hoff[2] = {0, 1};
    const int32_t smesh = 0, interp = 2; const double zo = 0.0, h = 0.0;
    shb_batch* b = nullptr;
    int rc = batch_create_impl(1, verts, voff, faces, foff, 1, &smesh, &zo, &h, hoff, &interp, nullptr, &b);
    if (rc) return rc;
    shb_mesh* m = new shb_mesh;
    m->vert = b->vert; m->vz = b->vz; m->face = b->face; m->adj = b->adj; m->d_bad = b->d_bad; m->nv = n_vert; m->nf = n_face;
    m->stream = b->stream;
    b->vert = nullptr; b->vz = nullptr; b->face = nullptr; b->adj = nullptr; b->d_bad = nullptr;
    CK(cudaStreamSynchronize(m->stream));              // verts / faces are free to go when this call returns
    uint32_t bad = 0;
    CK(cudaMemcpy(&bad, m->d_bad, sizeof bad, cudaMemcpyDeviceToHost));
    CK(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming));
    CK(cudaEventRecord(m->ready, m->stream));
    shb_batch_free(b);
    if (bad) { mesh_unref(m); return fail(SHB_E_INVALID, "face index out of range for its mesh"); }
    *out = m;
    return SHB_OK;
}

SHB_API int shb_mesh_free(shb_mesh* m) {
    SHB_ENTER;
    if (m) mesh_unref(m);
    return SHB_OK;
}

// ---- scope row f2: STL bytes -> welded mesh (-> frame) on the device (kernels: shb_meshio.cu) ---------------------------
extern "C" {
size_t shb_frame_state_bytes(void);
size_t shb_frame_transform_offset(void);
size_t shb_frame_zb_offset(void);
size_t shb_frame_resid_offset(void);
size_t shb_frame_flip_offset(void);
int shb_launch_weld_count(const unsigned char* stl, uint32_t n_corner, uint32_t* table, uint32_t* first, uint32_t mask, uint32_t* slot_of,
                          uint32_t* tile_sum, uint32_t* n_vert_out, cudaStream_t st);
int shb_launch_weld_emit(const unsigned char* stl, uint32_t n_corner, const uint32_t* first, const uint32_t* slot_of, const uint32_t* tile_off,
                         uint32_t* vid, double4* vert, double* vz, int4* face, cudaStream_t st);
int shb_launch_frame(double4* vert, double* vz, uint32_t V, void* state, double* part, int n_sm, cudaStream_t st);
int shb_launch_mesh_unpack(const double4* vert, int64_t nv, const int4* face, int64_t nf, double* v_out, int64_t* f_out, cudaStream_t st);
}

SHB_API int shb_mesh_from_stl(const void* stl, int64_t n_bytes, uint32_t flags, shb_mesh** out, int64_t* n_vert, int64_t* n_face,
                              double* frame_out) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!out || !stl) return fail(SHB_E_INVALID, "bad argument");
    *out = nullptr;
    if (n_bytes < 84) return fail(SHB_E_INVALID, "not a binary STL: %lld bytes", (long long)n_bytes);
    uint32_t T = 0;
    std::memcpy(&T, static_cast<const unsigned char*>(stl) + 80, 4);
    if (84 + 50 * (int64_t)T != n_bytes)
        return fail(SHB_E_INVALID, "not a binary STL (header says %u triangles, file has %lld bytes; ASCII STL is not read on the device)", T,
                    (long long)n_bytes);
    if (T == 0) return fail(SHB_E_INVALID, "STL without triangles");
    if (T >= (1u << 28)) return fail(SHB_E_CAPACITY, "too many triangles in one mesh");      // 6 T table slots must fit 32 bits
    const uint32_t nc = 3u * T;
    cudaStream_t st = g.stream;
    unsigned char* d_stl = nullptr; uint32_t *table = nullptr, *first = nullptr, *slot_of = nullptr, *vid = nullptr, *tile_sum = nullptr, *d_nv = nullptr;
    uint32_t hs = 1024;
    while ((uint64_t)hs < 2ull * nc) hs <<= 1;
    const uint32_t tiles = (nc + 2047u) / 2048u;
    CK(dalloc(&d_stl, (size_t)n_bytes + 16, st)); CK(dalloc(&table, 2 * (size_t)hs, st)); first = table + hs;
    CK(dalloc(&slot_of, 2 * (size_t)nc, st)); vid = slot_of + nc;
    CK(dalloc(&tile_sum, (size_t)tiles + 1, st)); d_nv = tile_sum + tiles;
    CK(cudaMemcpyAsync(d_stl, stl, (size_t)n_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(table, 0xFF, 2 * (size_t)hs * sizeof(uint32_t), st));
    g.launches += shb_launch_weld_count(d_stl, nc, table, first, hs - 1, slot_of, tile_sum, d_nv, st);
    CK(cudaGetLastError());
    uint32_t* h_nv = static_cast<uint32_t*>(pinned_get(16));
    if (!h_nv) return fail(SHB_E_CUDA, "pinned allocation failed");
    CK(cudaMemcpyAsync(h_nv, d_nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));                             // the vertex count sizes the mesh's own arrays
    const uint32_t V = *h_nv;
    pinned_put(h_nv);
    shb_mesh* m = new shb_mesh;
    m->nv = V; m->nf = T; m->stream = st;
    CK(dalloc(&m->vert, V, st)); CK(dalloc(&m->vz, V, st)); CK(dalloc(&m->face, T, st)); CK(dalloc(&m->d_bad, 1, st));
    CK(cudaMemsetAsync(m->d_bad, 0, sizeof(uint32_t), st));
    g.launches += shb_launch_weld_emit(d_stl, nc, first, slot_of, tile_sum, vid, m->vert, m->vz, m->face, st);
    CK(cudaGetLastError());
    double* h_state = nullptr; unsigned char* d_state = nullptr; double* part = nullptr;
    const size_t sb = shb_frame_state_bytes();
    if (flags & SHB_STL_FRAME) {
        CK(dalloc(&d_state, sb, st)); CK(dalloc(&part, (size_t)g.n_sm * 20, st));
        CK(cudaMemsetAsync(d_state, 0, sb, st));
        g.launches += shb_launch_frame(m->vert, m->vz, V, d_state, part, g.n_sm, st);
        CK(cudaGetLastError());
        h_state = static_cast<double*>(pinned_get(sb));
        if (!h_state) return fail(SHB_E_CUDA, "pinned allocation failed");
        CK(cudaMemcpyAsync(h_state, d_state, sb, cudaMemcpyDeviceToHost, st));
    }
    { int rc_adj = build_adjacency(m->face, T, &m->adj, st); if (rc_adj) return rc_adj; }
    CK(cudaStreamSynchronize(st));
    if (frame_out) {
        for (int i = 0; i < 22; ++i) frame_out[i] = 0.0;
        frame_out[0] = frame_out[5] = frame_out[10] = frame_out[15] = 1.0; frame_out[19] = 1.0;
        if (h_state) {
            const unsigned char* hb = reinterpret_cast<const unsigned char*>(h_state);
            std::memcpy(frame_out, hb + shb_frame_transform_offset(), 16 * sizeof(double));
            std::memcpy(frame_out + 16, hb + shb_frame_zb_offset(), 2 * sizeof(double));
            frame_out[18] = std::fabs(frame_out[16]) + std::fabs(frame_out[17]);
            std::memcpy(frame_out + 19, hb + shb_frame_flip_offset(), sizeof(double));
            std::memcpy(frame_out + 20, hb + shb_frame_resid_offset(), 2 * sizeof(double));
        }
    }
    if (h_state) pinned_put(h_state);
    dfree(d_stl, st); dfree(table, st); dfree(slot_of, st); dfree(tile_sum, st); dfree(d_state, st); dfree(part, st);
    CK(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming));
    CK(cudaEventRecord(m->ready, st));
    if (n_vert) *n_vert = V;
    if (n_face) *n_face = T;
    *out = m;
    return SHB_OK;
}

SHB_API int shb_mesh_read(shb_mesh* mesh, double* verts, int64_t* faces) {
    SHB_ENTER;
    if (!mesh || !verts || !faces) return fail(SHB_E_INVALID, "bad argument");
    cudaStream_t st = g.stream;
    if (mesh->stream != st && mesh->ready) CK(cudaStreamWaitEvent(st, mesh->ready, 0));
    double* dv = nullptr; int64_t* df = nullptr;
    CK(dalloc(&dv, 3 * (size_t)mesh->nv, st)); CK(dalloc(&df, 3 * (size_t)mesh->nf, st));
    g.launches += shb_launch_mesh_unpack(mesh->vert, mesh->nv, mesh->face, mesh->nf, dv, df, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(verts, dv, 3 * (size_t)mesh->nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(faces, df, 3 * (size_t)mesh->nf * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(dv, st); dfree(df, st);
    return SHB_OK;
}

extern "C" int shb_launch_transform_verts(const double4* in, int64_t n, const double* m16_dev, double4* out, double* vz, cudaStream_t st);

/* a second resident mesh = `src` under the 4x4 row-major matrix `to_2d` (vertices only; faces and adjacency are copied on
 * the device).  x' = ((m00 x + m01 y) + m02 z) + m03, each product and sum rounded on its own (no FMA). */
SHB_API int shb_mesh_transform(shb_mesh* src, const double* to_2d, shb_mesh** out) {
    SHB_ENTER;
    if (!src || !to_2d || !out) return fail(SHB_E_INVALID, "bad argument");
    cudaStream_t st = g.stream;
    if (src->stream != st && src->ready) CK(cudaStreamWaitEvent(st, src->ready, 0));
    shb_mesh* m = new shb_mesh;
    m->nv = src->nv; m->nf = src->nf; m->stream = st;
    double* d_m = nullptr;
    CK(dalloc(&m->vert, m->nv, st)); CK(dalloc(&m->vz, m->nv, st)); CK(dalloc(&m->face, std::max<int64_t>(m->nf, 1), st));
    CK(dalloc(&m->adj, 4 * (size_t)std::max<int64_t>(m->nf, 1), st)); CK(dalloc(&m->d_bad, 1, st)); CK(dalloc(&d_m, 16, st));
    Staging stage;
    CK(stage.copy(d_m, to_2d, 16 * sizeof(double), st));
    g.launches += shb_launch_transform_verts(src->vert, m->nv, d_m, m->vert, m->vz, st);
    CK(cudaMemcpyAsync(m->face, src->face, (size_t)m->nf * sizeof(int4), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(m->adj, src->adj, 4 * (size_t)m->nf * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(m->d_bad, 0, sizeof(uint32_t), st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    stage.release();
    dfree(d_m, st);
    if (!src->last_use) CK(cudaEventCreateWithFlags(&src->last_use, cudaEventDisableTiming));
    CK(cudaEventRecord(src->last_use, st));
    CK(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming));
    CK(cudaEventRecord(m->ready, st));
    *out = m;
    return SHB_OK;
}

/* sweeps on ONE resident mesh: shb_batch_create without the upload */
SHB_API int shb_batch_create_on(shb_mesh* mesh, int32_t n_sweep, const double* z_orig, const double* heights, const int64_t* height_off,
                                const int32_t* interp_num, shb_batch** out) {
    SHB_ENTER;
    if (!mesh) return fail(SHB_E_INVALID, "null mesh");
    const int64_t voff[2] = {0, mesh->nv}, foff[2] = {0, mesh->nf};
    std::vector<int32_t> smesh((size_t)std::max(n_sweep, 0), 0);
    return batch_create_impl(1, nullptr, voff, nullptr, foff, n_sweep, smesh.data(), z_orig, heights, height_off, interp_num, mesh, out);
}

/* Trimesh.section(plane_normal, plane_origin) on a resident mesh (reference call sites: mesh.py:95-99,158-161,
 * surgical_neck.py:37-50, anatomic_neck.py:160-165, arthroplasty.py:71): ONE plane, any normal.  A +z normal runs on the
 * mesh as it is; any other normal first moves the vertices into a frame whose z axis is the normal and whose origin is
 * plane_origin (on the device).  to_3d (16 doubles, row-major) receives the matrix that takes the result's 2-D
 * coordinates (x, y, 0) back to the mesh frame.  The result carries plane records and contours; fetch as usual. */
SHB_API int shb_section(shb_mesh* mesh, const double* plane_normal, const double* plane_origin, uint32_t outputs_mask, double* to_3d,
                        shb_result** out) {
    SHB_ENTER;
    if (!mesh || !plane_normal || !plane_origin || !out) return fail(SHB_E_INVALID, "bad argument");
    *out = nullptr;
    const double nx = plane_normal[0], ny = plane_normal[1], nz = plane_normal[2];
    const double nn = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (!(nn > 0.0)) return fail(SHB_E_INVALID, "zero plane normal");
    const double z0 = 0.0, h0 = 0.0; const int64_t hoff[2] = {0, 1}; const int32_t interp = 2;
    double t3[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    shb_batch* b = nullptr;
    shb_mesh* tilted = nullptr;
    int rc;
    if (nx == 0.0 && ny == 0.0 && nz > 0.0) {
        const double zo = plane_origin[2];
        t3[11] = zo;
        rc = shb_batch_create_on(mesh, 1, &zo, &h0, hoff, &interp, &b);
    } else {
        // right-handed frame (ex, ey, n): ex = the coordinate axis least aligned with n, made orthogonal to it
        const double n[3] = {nx / nn, ny / nn, nz / nn};
        int k = std::fabs(n[0]) <= std::fabs(n[1]) ? (std::fabs(n[0]) <= std::fabs(n[2]) ? 0 : 2) : (std::fabs(n[1]) <= std::fabs(n[2]) ? 1 : 2);
        double ex[3] = {0, 0, 0}; ex[k] = 1.0;
        const double dp = ex[0] * n[0] + ex[1] * n[1] + ex[2] * n[2];
        for (int i = 0; i < 3; ++i) ex[i] -= dp * n[i];
        const double en = std::sqrt(ex[0] * ex[0] + ex[1] * ex[1] + ex[2] * ex[2]);
        for (int i = 0; i < 3; ++i) ex[i] /= en;
        const double ey[3] = {n[1] * ex[2] - n[2] * ex[1], n[2] * ex[0] - n[0] * ex[2], n[0] * ex[1] - n[1] * ex[0]};
        const double* R[3] = {ex, ey, n};
        double t2[16] = {0};
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) { t2[4 * r + c] = R[r][c]; t3[4 * c + r] = R[r][c]; }
            t2[4 * r + 3] = -(R[r][0] * plane_origin[0] + R[r][1] * plane_origin[1] + R[r][2] * plane_origin[2]);
            t3[4 * r + 3] = plane_origin[r];
        }
        t2[15] = 1.0;
        rc = shb_mesh_transform(mesh, t2, &tilted);
        if (rc) return rc;
        rc = shb_batch_create_on(tilted, 1, &z0, &h0, hoff, &interp, &b);
    }
    if (rc) { if (tilted) mesh_unref(tilted); return rc; }
    // flag 8: the Path3D route of Trimesh.section (row hashes over three columns never pack; no CCW normalisation)
    rc = batch_run_impl(b, nullptr, outputs_mask | SHB_OUT_PLANE | SHB_OUT_CONTOURS, 0, 8u, out);
    if (!rc) { (*out)->batch = nullptr; }
    shb_batch_free(b);
    if (tilted) mesh_unref(tilted);
    if (rc) return rc;
    if (to_3d) std::memcpy(to_3d, t3, sizeof t3);
    return SHB_OK;
}

SHB_API int shb_result_free(shb_result* r) {
    SHB_ENTER;
    if (!r) return SHB_OK;
    if (r->pending) { cudaStreamSynchronize(g.copy); r->pending = false; }
    if (!r->lf_staged.empty()) {                             // a no-wait front end that was never waited for: its inputs may still be read
        cudaStreamSynchronize(r->stream ? r->stream : g.stream);
        for (void* p : r->lf_staged) pinned_put(p);
        r->lf_staged.clear();
    }
    cudaStream_t st = r->stream ? r->stream : g.stream;      // the stream that computed it: frees are ordered behind its kernels
    ShbDev& d = r->d;
    dfree(r->blk_keep_a, st); dfree(r->blk_keep_b, st); dfree(r->blk_tmp_a, st); dfree(r->blk_tmp_b, st);
    for (int a = 0; a < 6; ++a) if (d.prof[a]) { cudaFreeAsync(d.prof[a], st); d.prof[a] = nullptr; }
    if (d.radial) { cudaFreeAsync(d.radial, st); d.radial = nullptr; }
    dfree(d.scratch, st);
    if (d.sweep) { cudaFreeAsync(const_cast<ShbSweep*>(d.sweep), st); d.sweep = nullptr; }
    dfree(r->d_ct_off, st); dfree(r->d_pt_off, st); dfree(r->d_pts_c, st); dfree(r->d_ctpt_c, st); dfree(r->d_ctarea_c, st);
    void* hp[] = {r->h_nseg, r->h_nent, r->h_sel, r->h_face_index, r->h_status, r->h_seg_off, r->h_ct_off, r->h_pt_off,
                  r->h_bounds, r->h_centroid, r->h_area1, r->h_segments, r->h_pts, r->h_ctarea, r->h_ctpt,
                  r->h_arr[0], r->h_arr[1], r->h_arr[2], r->h_arr[3], r->h_arr[4], r->h_arr[5], r->h_arr[6]};
    for (void* p : hp) if (p) pinned_put(p);
    if (r->done) cudaEventDestroy(r->done);
    delete r;
    return SHB_OK;
}

static const uint32_t kArrBit[SHB_N_ARR] = {SHB_OUT_IXY, SHB_OUT_IXY_CENTERED, SHB_OUT_ITR, SHB_OUT_ITR_START, SHB_OUT_ITR_CENTERED,
                                            SHB_OUT_ITR_CENTERED_START, SHB_OUT_RADIAL};

SHB_API int shb_batch_run(shb_batch* b, uint32_t outputs_mask, int32_t n_angles, shb_result** out) {
    return shb_batch_run_req(b, nullptr, outputs_mask, n_angles, out);
}

SHB_API int shb_batch_run_req(shb_batch* b, const shb_sweep_request* req, uint32_t outputs_mask, int32_t n_angles, shb_result** out) {
    SHB_ENTER;
    return batch_run_impl(b, req, outputs_mask, n_angles, 0u, out);
}

static int batch_run_impl(shb_batch* b, const shb_sweep_request* req, uint32_t outputs_mask, int32_t n_angles, uint32_t flags, shb_result** out) {
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!b || !out) return fail(SHB_E_INVALID, "null batch/out");
    *out = nullptr;
    outputs_mask |= SHB_OUT_PLANE;
    cudaStream_t st = g.stream;
    const uint32_t G = b->G;
    struct ResultDel { void operator()(shb_result* p) const { shb_result_free(p); } };
    std::unique_ptr<shb_result, ResultDel> r(new shb_result);
    r->batch = b; r->sweeps = b->sweeps; r->G = G; r->stream = st;
    r->rel.resize((size_t)b->n_sweep * 3);
    r->esz = (outputs_mask & SHB_OUT_F32) ? 4 : 8;
    // ---- per-sweep requests: which windowed arrays, over which rows (slice.py:157-164 windows of the consumers)
    uint32_t any_mask = 0;
    uint64_t excluded = 0, key = 1469598103934665603ull;          // FNV over the request: identifies the cached launch list
    for (int s = 0; s < b->n_sweep; ++s) {
        ShbSweep& sw = r->sweeps[s];
        const uint32_t m = req ? req[s].outputs_mask : outputs_mask;
        uint32_t lo_any = sw.n_plane, hi_any = 0;
        for (int a = 0; a < SHB_N_ARR; ++a) {
            uint32_t lo = 0, hi = 0;
            if (m & kArrBit[a]) {
                int64_t l = req ? req[s].row_lo[a] : 0, h = req ? req[s].row_hi[a] : -1;
                if (h < 0) h = sw.n_plane;
                if (l < 0 || l > h || h > (int64_t)sw.n_plane)
                    return fail(SHB_E_INVALID, "sweep %d: window [%lld, %lld) of array %d is outside its %u planes", s, (long long)l, (long long)h, a, sw.n_plane);
                lo = (uint32_t)l; hi = (uint32_t)h;
            }
            sw.win_lo[a] = lo; sw.win_hi[a] = hi;
            const uint64_t per = a == SHB_A_RADIAL ? (uint64_t)std::max(n_angles, 0) : 2ull * sw.interp_num;
            sw.arr_off[a] = r->arr_total[a];
            r->arr_total[a] += (uint64_t)(hi - lo) * per;
            if (hi > lo) { any_mask |= kArrBit[a]; lo_any = std::min(lo_any, lo); hi_any = std::max(hi_any, hi); }
            key = (key ^ (((uint64_t)lo << 32) | hi)) * 1099511628211ull;
        }
        excluded += hi_any > lo_any ? sw.n_plane - (hi_any - lo_any) : sw.n_plane;
    }
    if ((any_mask & SHB_OUT_RADIAL) && n_angles < 1) return fail(SHB_E_INVALID, "n_angles must be >= 1 for the radial image");
    if (!(any_mask & SHB_OUT_RADIAL)) n_angles = 0;
    outputs_mask = (outputs_mask & ~(SHB_OUT_ALL_PROFILES | SHB_OUT_RADIAL)) | any_mask;
    r->mask = outputs_mask; r->n_angles = (uint32_t)n_angles;
    ShbDev& d = r->d;
    d.vert = b->vert; d.vz = b->vz; d.face = b->face; d.adj = b->adj; d.item_off = b->d_item_off;
    d.h_sorted = b->h_sorted; d.h_orig = b->h_orig; d.oz = b->oz; d.plane_out = b->plane_out; d.plane_in = b->plane_in; d.plane_sweep = b->plane_sweep;
    d.n_sweep = (uint32_t)b->n_sweep; d.n_plane = G; d.n_item = b->n_item;
    d.n_angles = (uint32_t)n_angles; d.outputs_mask = outputs_mask;
    if (b->stream != st && b->uploaded) CK(cudaStreamWaitEvent(st, b->uploaded, 0));     // the batch was uploaded on another stream
    // sweep descriptors carry the windows and output offsets of this run
    ShbSweep* d_sw = nullptr;
    CK(dalloc(&d_sw, b->n_sweep, st));
    Staging stage;
    struct StageSync { Staging& s; cudaStream_t st; bool armed = true;      // an early return must not hand the pinned buffers back
        ~StageSync() { if (armed && !s.bufs.empty()) { cudaStreamSynchronize(st); s.release(); } } } stage_sync{stage, st};
    CK(stage.copy(d_sw, r->sweeps.data(), b->n_sweep * sizeof(ShbSweep), st));
    d.sweep = d_sw;
    // launch list of the resample kernel: only the planes some windowed output wants (built once per request)
    d.resample_order = nullptr; d.n_resample = 0;
    if (excluded && any_mask) {
        if (b->rs_key != key || !b->rs_order) {
            std::vector<uint32_t> so(G), list;
            // the host copy of the stitch order is not kept: rebuild "sweep ends first" for the wanted planes only
            std::vector<std::pair<double, uint32_t>> cand;
            for (int s = 0; s < b->n_sweep; ++s) {
                const ShbSweep& sw = r->sweeps[s];
                for (uint32_t lp = 0; lp < sw.n_plane; ++lp) {
                    bool want = false;
                    for (int a = 0; a < SHB_N_ARR; ++a) want |= lp >= sw.win_lo[a] && lp < sw.win_hi[a];
                    if (!want) continue;
                    const uint32_t e = std::min(lp, sw.n_plane - 1 - lp);
                    cand.push_back({(double)e / (double)std::max<uint32_t>(sw.n_plane, 1u), sw.plane_off + lp});
                }
            }
            std::stable_sort(cand.begin(), cand.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            list.reserve(cand.size());
            for (auto& c : cand) list.push_back(c.second);
            if (b->rs_order) dfree(b->rs_order, b->stream);
            CK(dalloc(&b->rs_order, list.size(), st));
            CK(stage.copy(b->rs_order, list.data(), list.size() * sizeof(uint32_t), st));
            b->n_rs = (uint32_t)list.size(); b->rs_key = key;
        }
        d.resample_order = b->rs_order; d.n_resample = b->n_rs;
    }

    const size_t n_tiles = ((size_t)G + 4095) / 4096;
    {
        // one block of scratch and one block the result keeps, carved by offset; the counters that must start at zero
        // lie together at the front of the scratch block and are cleared by ONE memset
        struct Carve { size_t off = 0; size_t add(size_t bytes) { off = (off + 255) & ~(size_t)255; const size_t o = off; off += std::max<size_t>(bytes, 16); return o; } };
        Carve t, k;
        const size_t o_inc = t.add(G * 4), o_cur = t.add(G * 4), o_dec = t.add(((size_t)G + 1) * 4), o_scan = t.add((4 * n_tiles + 4) * 8);
        const size_t zero_end = t.off;
        const size_t o_lo = t.add((size_t)d.n_item * 4), o_span = t.add((size_t)d.n_item * 4), o_rec = t.add((size_t)d.n_item * 16),
                     o_soff = t.add(((size_t)G + 1) * 4), o_cnt = t.add(G * 4), o_cap = t.add(((size_t)G + 1) * 4), o_caps = t.add(G * 4),
                     o_big = t.add(G * 4), o_decl = t.add(2 * (size_t)G * 4), o_dup = t.add(G * 4);
        const size_t k_tot = k.add(16 * 4), k_tot64 = k.add(2 * 8), k_seg = k.add(((size_t)G + 1) * 4), k_meta = k.add((size_t)G * sizeof(ShbPlaneMeta)),
                     k_nseg = k.add(G * 4), k_nent = k.add(G * 4), k_stat = k.add(G * 4), k_bnd = k.add((size_t)G * 32), k_cen = k.add((size_t)G * 16),
                     k_area = k.add((size_t)G * 8), k_sel = k.add((size_t)G * 8);
        CK(dalloc_bytes(&r->blk_tmp_a, t.off, st)); CK(dalloc_bytes(&r->blk_keep_a, k.off, st));
        unsigned char* T = (unsigned char*)r->blk_tmp_a; unsigned char* K = (unsigned char*)r->blk_keep_a;
        d.inc = (uint32_t*)(T + o_inc); d.sort_cur = (uint32_t*)(T + o_cur); d.dec = (uint32_t*)(T + o_dec); d.scan_state = (unsigned long long*)(T + o_scan);
        d.item_lo = (uint32_t*)(T + o_lo); d.item_span = (uint32_t*)(T + o_span); d.rec = (uint4*)(T + o_rec); d.sort_off = (uint32_t*)(T + o_soff);
        d.cnt = (uint32_t*)(T + o_cnt); d.cap_off = (uint32_t*)(T + o_cap); d.cap_sorted = (uint32_t*)(T + o_caps); d.big_list = (uint32_t*)(T + o_big);
        d.decl_list = (uint32_t*)(T + o_decl); d.dup_list = (uint32_t*)(T + o_dup);
        d.totals = (uint32_t*)(K + k_tot); d.totals64 = (unsigned long long*)(K + k_tot64); d.seg_off = (uint32_t*)(K + k_seg);
        d.meta = (ShbPlaneMeta*)(K + k_meta); d.o_nseg = (int32_t*)(K + k_nseg); d.o_nent = (int32_t*)(K + k_nent); d.o_status = (uint32_t*)(K + k_stat);
        d.o_bounds = (double*)(K + k_bnd); d.o_centroid = (double*)(K + k_cen); d.o_area1 = (double*)(K + k_area); d.o_sel = (int32_t*)(K + k_sel);
        CK(cudaMemsetAsync(T, 0, zero_end, st));
        CK(cudaMemsetAsync(d.totals, 0, 16 * sizeof(uint32_t), st));
    }
    CK(cudaMemcpyAsync(d.totals + SHB_T_BAD, b->d_bad, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));   // slot 1: bad-face flag

    // shared-memory capacities (leave headroom for static shared memory)
    const size_t budget = (g.smem_optin > 8192 ? g.smem_optin - 4096 : 40960);
    uint32_t cap = 1;
    for (uint32_t step = 1u << 20; step; step >>= 1)
        if (shb_stitch_ws_bytes(cap + step) <= budget) cap += step;
    d.stitch_cap = cap;
    uint32_t pcap = 2;
    for (uint32_t step = 1u << 20; step; step >>= 1)
        if (shb_resample_ws_bytes(pcap + step, b->max_interp, (uint32_t)n_angles) <= budget) pcap += step;
    d.resample_cap = pcap;
    if (const char* f = getenv("SHB_DEBUG_SMEM_CAP")) {      // test hook: push planes onto the global-workspace path
        uint32_t v = (uint32_t)atoi(f);
        if (v >= 1) { d.stitch_cap = std::min(d.stitch_cap, v); d.resample_cap = std::min(d.resample_cap, v + 2); }
    }
    cap = d.stitch_cap; pcap = d.resample_cap;
    d.debug = flags;
    if (getenv("SHB_DEBUG_RADIAL_GENERAL")) d.debug |= 1u;       // radius image by the all-candidates path on every plane
    if (getenv("SHB_DEBUG_MINRANK_ORDER")) d.debug |= 2u;        // contour order by minimum rank (no CPython-set emulation)
    if (getenv("SHB_DEBUG_NO_WARP_STITCH")) d.debug |= 4u;       // every plane through the CTA stitcher
    if (getenv("SHB_DEBUG_ALLPAIRS_RANK")) d.debug |= 16u;       // node ranks of several-contour planes by the all-pairs loop
    d.stitch_order = getenv("SHB_DEBUG_NO_PERMUTE") ? nullptr : b->stitch_order;
    // K1 sizes everything downstream: the candidate triangles per plane (an upper bound of the hits that is exact
    // except on planes through vertices) are published as soon as the bucket histograms are scanned, and the host
    // waits for them while the device goes on with the counting sort
    { StageTimer t(0, st); t.stop(shb_launch_bucket(d, st)); }
    { StageTimer t(1, st); t.stop(shb_launch_scan_candidates(d, st)); }
    g.launches += shb_launch_publish(d.totals, 8, d.totals64, g.h_totals, g.h_totals64, st);
    CK(cudaEventRecord(g.sized, st));
    { StageTimer t(1, st); t.stop(shb_launch_scan_planes(d, st)); }
    { StageTimer t(2, st); t.stop(shb_launch_scatter(d, st)); }
    CK(cudaMemsetAsync(d.sort_cur, 0, G * sizeof(uint32_t), st));      // reused as the hit-list cursors
    const bool dbg_t = getenv("SHB_DEBUG_TIMING") != nullptr;
    auto now_ms = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    const double tw0 = now_ms();
    CK(cudaEventSynchronize(g.sized));      // the one host wait of a run
    const double tw1 = now_ms();
    stage.release();
    if (b->stage && b->stream == st) { delete b->stage; b->stage = nullptr; }      // the batch upload has executed too
    if (g.h_totals[SHB_T_BAD]) return fail(SHB_E_INVALID, "face index out of range for its mesh");
    const uint32_t S = g.h_totals[SHB_T_CAP], maxcand = g.h_totals[SHB_T_MAXN];      // S: capacity (candidates), >= segments
    if (g.h_totals64[0] >= (1ull << 31)) return fail(SHB_E_CAPACITY, "%llu candidate segments in one batch; split it", g.h_totals64[0]);
    r->W = S;
    const uint32_t avgn = (uint32_t)(S / std::max<uint32_t>(G, 1u));     // mean segments per plane picks the CTA size
    const bool want_seg = (outputs_mask & SHB_OUT_SEGMENTS) != 0;
    {
        struct Carve { size_t off = 0; size_t add(size_t bytes) { off = (off + 255) & ~(size_t)255; const size_t o = off; off += std::max<size_t>(bytes, 16); return o; } };
        Carve k;
        const size_t k_fi = k.add((want_seg ? (size_t)S : 1) * 4), k_segm = k.add(32 * (size_t)S), k_pts = k.add(32 * (size_t)S + 64),
                     k_cs = k.add((size_t)S * 4), k_cl = k.add((size_t)S * 4), k_ca = k.add((size_t)S * 8);
        CK(dalloc_bytes(&r->blk_tmp_b, 16 * ((size_t)S + 1), st));      // 16-byte hit records: every plane's list is a TMA-aligned run
        CK(dalloc_bytes(&r->blk_keep_b, k.off, st));
        unsigned char* K = (unsigned char*)r->blk_keep_b;
        d.hits = (uint4*)r->blk_tmp_b;
        d.face_index = (int32_t*)(K + k_fi); d.segments = (double*)(K + k_segm); d.pts = (double*)(K + k_pts);
        d.ct_start = (uint32_t*)(K + k_cs); d.ct_len = (uint32_t*)(K + k_cl); d.ct_area = (double*)(K + k_ca);
    }
    bool any_prof = false;
    for (int a = 0; a < 6; ++a)
        if (r->arr_total[a]) { CK(dalloc_bytes(&d.prof[a], r->arr_total[a] * r->esz, st)); any_prof = true; }
    if (r->arr_total[SHB_A_RADIAL]) {
        CK(dalloc_bytes(&d.radial, r->arr_total[SHB_A_RADIAL] * r->esz, st)); any_prof = true;
        // ray directions from the host's libm (what numpy evaluates): theta_k = -pi + (2*pi/A)*k; cached per A
        if (g.angle_n != n_angles) {
            std::vector<double> tab(2 * (size_t)n_angles);
            const double pi = 3.141592653589793, dA = 2.0 * pi / (double)n_angles;
            for (int k = 0; k < n_angles; ++k) { double th = -pi + dA * (double)k; tab[2 * k] = std::cos(th); tab[2 * k + 1] = std::sin(th); }
            if (g.angle_tab) { CK(cudaDeviceSynchronize()); cudaFree(g.angle_tab); g.angle_tab = nullptr; }
            CK(cudaMalloc(&g.angle_tab, (size_t)n_angles * sizeof(double2)));
            CK(cudaMemcpy(g.angle_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
            g.angle_n = n_angles;
        }
        d.angle_cs = g.angle_tab;
    }
    const bool need_scratch = maxcand > cap || (size_t)maxcand + 1 > pcap;
    if (need_scratch) {
        size_t s1 = shb_stitch_ws_bytes(maxcand), s2 = shb_resample_ws_bytes(maxcand + 1, b->max_interp, (uint32_t)n_angles);
        d.scratch_stride = (std::max(s1, s2) + 255) & ~(size_t)255;
        CK(dalloc(&d.scratch, d.scratch_stride * (size_t)g.n_sm, st));
    }

    { StageTimer t(3, st); t.stop(shb_launch_intersect(d, st)); }
    { StageTimer t(4, st); t.stop(shb_launch_scan_counts(d, st)); }
    { StageTimer t(5, st); t.stop(shb_launch_stitch(d, maxcand, avgn, b->max_faces, budget, g.n_sm, st, g.aux, g.ev_fork, g.ev_join, g.ev_mid, b->n_sweep > b->n_mesh)); }
    if (any_prof) { StageTimer t(6, st); t.stop(shb_launch_resample(d, maxcand, avgn, b->max_interp, g.n_sm, st)); }
    CK(cudaGetLastError());
    // stage scratch is dead once the kernels above are enqueued (stream ordered)
    dfree(r->blk_tmp_a, st); dfree(r->blk_tmp_b, st); dfree(d.scratch, st);
    d.item_lo = d.item_span = d.inc = d.sort_off = d.sort_cur = d.cnt = d.dec = d.cap_off = d.cap_sorted = d.big_list = d.decl_list = d.dup_list = nullptr;
    d.scan_state = nullptr; d.rec = nullptr; d.hits = nullptr;
    cudaFreeAsync(d_sw, st); d.sweep = nullptr;
    d.resample_order = nullptr;
    CK(cudaEventCreateWithFlags(&r->done, cudaEventDisableTiming));
    CK(cudaEventRecord(r->done, st));
    if (!b->last_use) CK(cudaEventCreateWithFlags(&b->last_use, cudaEventDisableTiming));
    CK(cudaEventRecord(b->last_use, st));
    stage_sync.armed = false;
    if (dbg_t) fprintf(stderr, "[shb] run: host waited %.3f ms for the sizes, then enqueued the rest in %.3f ms\n", tw1 - tw0, now_ms() - tw1);
    *out = r.release();
    return SHB_OK;
}

static int fetch_plane(shb_result* r) {
    if (r->have_plane) return SHB_OK;
    cudaStream_t st = g.copy;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    const size_t G = r->G;
    auto grab = [&](auto** hp, const void* dp, size_t bytes) -> int {
        *hp = reinterpret_cast<std::remove_pointer_t<decltype(hp)>>(pinned_get(bytes));
        if (!*hp) return fail(SHB_E_NOMEM, "pinned host allocation of %zu bytes failed", bytes);
        CK(cudaMemcpyAsync(*hp, dp, bytes, cudaMemcpyDeviceToHost, st));
        return SHB_OK;
    };
    int rc;
    if ((rc = grab(&r->h_nseg, r->d.o_nseg, G * 4))) return rc;
    if ((rc = grab(&r->h_nent, r->d.o_nent, G * 4))) return rc;
    if ((rc = grab(&r->h_status, r->d.o_status, G * 4))) return rc;
    if ((rc = grab(&r->h_sel, r->d.o_sel, G * 8))) return rc;
    if ((rc = grab(&r->h_bounds, r->d.o_bounds, G * 32))) return rc;
    if ((rc = grab(&r->h_centroid, r->d.o_centroid, G * 16))) return rc;
    if ((rc = grab(&r->h_area1, r->d.o_area1, G * 8))) return rc;
    if ((rc = grab(&r->h_seg_off, r->d.seg_off, (G + 1) * 4))) return rc;
    r->have_plane = true; r->pending = true;
    return SHB_OK;
}

static int finish_fetch(shb_result* r) {
    if (r->pending) { CK(cudaStreamSynchronize(g.copy)); r->pending = false; }
    return SHB_OK;
}

static int enqueue_fetch(shb_result* r, uint32_t mask);

SHB_API int shb_result_fetch(shb_result* r, uint32_t mask) {
    SHB_ENTER;
    if (!r) return fail(SHB_E_INVALID, "null result");
    int rc = enqueue_fetch(r, mask);
    return rc ? rc : finish_fetch(r);
}

SHB_API int shb_result_fetch_async(shb_result* r, uint32_t mask) {
    SHB_ENTER;
    if (!r) return fail(SHB_E_INVALID, "null result");
    if (mask & SHB_OUT_CONTOURS) return fail(SHB_E_INVALID, "contours need a size readback; fetch them with shb_result_fetch");
    return enqueue_fetch(r, mask);
}

static int enqueue_fetch(shb_result* r, uint32_t mask) {
    cudaStream_t st = g.copy, cst = r->stream ? r->stream : g.stream;
    int rc = fetch_plane(r);
    if (rc) return rc;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    if ((mask & SHB_OUT_SEGMENTS) && !r->have_seg) {
        if (!(r->mask & SHB_OUT_SEGMENTS)) return fail(SHB_E_STATE, "segments / face_index were not in the outputs_mask of the run");
        const size_t S = r->W;                  // capacity: the exact count is seg_off[G], known once the plane arrays are in
        r->h_face_index = (int32_t*)pinned_get(S * 4); r->h_segments = (double*)pinned_get(S * 32);
        if (!r->h_face_index || !r->h_segments) return fail(SHB_E_NOMEM, "pinned host allocation failed");
        CK(cudaMemcpyAsync(r->h_face_index, r->d.face_index, S * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(r->h_segments, r->d.segments, S * 32, cudaMemcpyDeviceToHost, st));
        r->have_seg = true; r->pending = true;
    }
    if ((mask & SHB_OUT_CONTOURS) && !r->have_cont) {
        const size_t G = r->G;
        ShbDev d = r->d;
        CK(dalloc(&r->d_ct_off, G + 1, cst)); CK(dalloc(&r->d_pt_off, G + 1, cst));
        g.launches += shb_launch_scan_contours(d, r->d_ct_off, r->d_pt_off, cst);
        g.launches += shb_launch_publish(d.totals, 8, nullptr, g.h_totals, nullptr, cst);
        CK(cudaStreamSynchronize(cst));
        r->n_cont = g.h_totals[SHB_T_NCONT]; r->n_pts = g.h_totals[SHB_T_NPTS];
        CK(dalloc(&r->d_pts_c, 2 * (size_t)r->n_pts, cst)); CK(dalloc(&r->d_ctpt_c, (size_t)r->n_cont + 1, cst));
        CK(dalloc(&r->d_ctarea_c, r->n_cont, cst));
        g.launches += shb_launch_compact(d, r->d_ct_off, r->d_pt_off, r->d_pts_c, r->d_ctpt_c, r->d_ctarea_c, cst);
        CK(cudaGetLastError());
        CK(cudaEventRecord(r->done, cst));
        CK(cudaStreamWaitEvent(st, r->done, 0));
        r->h_ct_off = (uint32_t*)pinned_get((G + 1) * 4); r->h_pt_off = (uint32_t*)pinned_get((G + 1) * 4);
        r->h_pts = (double*)pinned_get((size_t)r->n_pts * 16); r->h_ctpt = (int64_t*)pinned_get(((size_t)r->n_cont + 1) * 8);
        r->h_ctarea = (double*)pinned_get((size_t)r->n_cont * 8);
        if (!r->h_ct_off || !r->h_pt_off || !r->h_pts || !r->h_ctpt || !r->h_ctarea) return fail(SHB_E_NOMEM, "pinned host allocation failed");
        CK(cudaMemcpyAsync(r->h_ct_off, r->d_ct_off, (G + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(r->h_pt_off, r->d_pt_off, (G + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(r->h_pts, r->d_pts_c, (size_t)r->n_pts * 16, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(r->h_ctpt, r->d_ctpt_c, ((size_t)r->n_cont + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(r->h_ctarea, r->d_ctarea_c, (size_t)r->n_cont * 8, cudaMemcpyDeviceToHost, st));
        r->have_cont = true; r->pending = true;
    }
    for (int a = 0; a < SHB_N_ARR; ++a)
        if ((mask & kArrBit[a]) && !r->h_arr[a]) {
            void* src = a == SHB_A_RADIAL ? r->d.radial : r->d.prof[a];
            if (!src) {
                if (mask & r->mask & kArrBit[a]) continue;              // requested, but no sweep window held a row
                return fail(SHB_E_STATE, "output array %d was not in the outputs_mask of the run", a);
            }
            r->h_arr[a] = pinned_get(r->arr_total[a] * r->esz);
            if (!r->h_arr[a]) return fail(SHB_E_NOMEM, "pinned host allocation failed");
            CK(cudaMemcpyAsync(r->h_arr[a], src, r->arr_total[a] * r->esz, cudaMemcpyDeviceToHost, st));
            r->pending = true;
        }
    return SHB_OK;
}

SHB_API int shb_result_totals(const shb_result* r_, int64_t* n_plane, int64_t* n_seg, int64_t* n_contour, int64_t* n_point) {
    SHB_ENTER;
    shb_result* r = const_cast<shb_result*>(r_);
    if (!r) return fail(SHB_E_INVALID, "null result");
    int rc = fetch_plane(r);
    if (rc) return rc;
    if ((rc = finish_fetch(r))) return rc;
    if (n_plane) *n_plane = r->G;
    if (n_seg) *n_seg = r->h_seg_off[r->G];
    if (n_contour || n_point) {
        int64_t c = 0, p = 0;
        if (r->have_cont) { c = r->n_cont; p = r->n_pts; }
        else for (uint32_t i = 0; i < r->G; ++i) c += r->h_nent[i];
        if (n_contour) *n_contour = c;
        if (n_point) *n_point = r->have_cont ? p : -1;
    }
    return SHB_OK;
}

SHB_API const void* shb_result_array(shb_result* r, int32_t which, int32_t sweep, int64_t shape[4], int32_t* ndim, int32_t* dtype) {
    SHB_ENTER;
    if (!r || !shape || !ndim || !dtype) { fail(SHB_E_INVALID, "null argument"); return nullptr; }
    if (sweep < 0 || sweep >= (int32_t)r->sweeps.size()) { fail(SHB_E_INVALID, "sweep %d out of range", sweep); return nullptr; }
    uint32_t need = SHB_OUT_PLANE;
    switch (which) {
        case SHB_ARR_FACE_INDEX: case SHB_ARR_SEGMENTS: need = SHB_OUT_SEGMENTS; break;
        case SHB_ARR_CONTOUR_OFF: case SHB_ARR_CONTOUR_PT_OFF: case SHB_ARR_CONTOUR_AREA: case SHB_ARR_POINTS: need = SHB_OUT_CONTOURS; break;
        case SHB_ARR_IXY: need = SHB_OUT_IXY; break;
        case SHB_ARR_IXY_CENTERED: need = SHB_OUT_IXY_CENTERED; break;
        case SHB_ARR_ITR: need = SHB_OUT_ITR; break;
        case SHB_ARR_ITR_START: need = SHB_OUT_ITR_START; break;
        case SHB_ARR_ITR_CENTERED: need = SHB_OUT_ITR_CENTERED; break;
        case SHB_ARR_ITR_CENTERED_START: need = SHB_OUT_ITR_CENTERED_START; break;
        case SHB_ARR_RADIAL: need = SHB_OUT_RADIAL; break;
        default: break;
    }
    if (which < 0 || which >= SHB_ARR_COUNT) { fail(SHB_E_INVALID, "unknown array id %d", which); return nullptr; }
    if (shb_result_fetch(r, need) != SHB_OK) return nullptr;
    const ShbSweep& sw = r->sweeps[sweep];
    const size_t p0 = sw.plane_off, P = sw.n_plane, N = sw.interp_num;
    auto rel = [&](int slot, const uint32_t* src, size_t first, size_t count) -> const void* {
        std::vector<int64_t>& v = r->rel[(size_t)sweep * 3 + slot];
        v.resize(count);
        for (size_t i = 0; i < count; ++i) v[i] = (int64_t)src[first + i] - (int64_t)src[first];
        return v.data();
    };
    *ndim = 1; shape[0] = (int64_t)P; shape[1] = shape[2] = shape[3] = 0;
    switch (which) {
        case SHB_ARR_N_SEG: *dtype = SHB_DT_I32; return r->h_nseg + p0;
        case SHB_ARR_N_ENT: *dtype = SHB_DT_I32; return r->h_nent + p0;
        case SHB_ARR_STATUS: *dtype = SHB_DT_U32; return r->h_status + p0;
        case SHB_ARR_AREA1: *dtype = SHB_DT_F64; return r->h_area1 + p0;
        case SHB_ARR_SEL: *dtype = SHB_DT_I32; *ndim = 2; shape[1] = 2; return r->h_sel + 2 * p0;
        case SHB_ARR_BOUNDS: *dtype = SHB_DT_F64; *ndim = 3; shape[1] = 2; shape[2] = 2; return r->h_bounds + 4 * p0;
        case SHB_ARR_CENTROID: *dtype = SHB_DT_F64; *ndim = 2; shape[1] = 2; return r->h_centroid + 2 * p0;
        case SHB_ARR_SEG_OFF: *dtype = SHB_DT_I64; shape[0] = (int64_t)P + 1; return rel(0, r->h_seg_off, p0, P + 1);
        case SHB_ARR_FACE_INDEX: *dtype = SHB_DT_I32; shape[0] = (int64_t)r->h_seg_off[p0 + P] - r->h_seg_off[p0];
            return r->h_face_index + r->h_seg_off[p0];
        case SHB_ARR_SEGMENTS: *dtype = SHB_DT_F64; *ndim = 3; shape[0] = (int64_t)r->h_seg_off[p0 + P] - r->h_seg_off[p0];
            shape[1] = 2; shape[2] = 2; return r->h_segments + 4 * (size_t)r->h_seg_off[p0];
        case SHB_ARR_CONTOUR_OFF: *dtype = SHB_DT_I64; shape[0] = (int64_t)P + 1; return rel(1, r->h_ct_off, p0, P + 1);
        case SHB_ARR_CONTOUR_AREA: *dtype = SHB_DT_F64; shape[0] = (int64_t)r->h_ct_off[p0 + P] - r->h_ct_off[p0];
            return r->h_ctarea + r->h_ct_off[p0];
        case SHB_ARR_CONTOUR_PT_OFF: {
            *dtype = SHB_DT_I64;
            const size_t c0 = r->h_ct_off[p0], c1 = r->h_ct_off[p0 + P];
            std::vector<int64_t>& v = r->rel[(size_t)sweep * 3 + 2];
            v.resize(c1 - c0 + 1);
            const int64_t base = r->h_pt_off[p0];
            for (size_t c = c0; c < c1; ++c) v[c - c0] = r->h_ctpt[c] - base;
            v[c1 - c0] = (int64_t)r->h_pt_off[p0 + P] - base;
            shape[0] = (int64_t)(c1 - c0 + 1);
            return v.data();
        }
        case SHB_ARR_POINTS: *dtype = SHB_DT_F64; *ndim = 2; shape[0] = (int64_t)r->h_pt_off[p0 + P] - r->h_pt_off[p0]; shape[1] = 2;
            return r->h_pts + 2 * (size_t)r->h_pt_off[p0];
        case SHB_ARR_RADIAL: *dtype = r->esz == 4 ? SHB_DT_F32 : SHB_DT_F64; *ndim = 2; shape[1] = r->n_angles;
            shape[0] = (int64_t)sw.win_hi[SHB_A_RADIAL] - sw.win_lo[SHB_A_RADIAL];         // the rows of the sweep's window
            if (shape[0] == 0 && (sw.n_plane > 0 || !r->h_arr[SHB_A_RADIAL])) { fail(SHB_E_STATE, "sweep %d did not request the radial image", sweep); return nullptr; }
            return (const char*)r->h_arr[SHB_A_RADIAL] + sw.arr_off[SHB_A_RADIAL] * r->esz;
        default: {
            const int a = which - SHB_ARR_IXY;
            *dtype = r->esz == 4 ? SHB_DT_F32 : SHB_DT_F64; *ndim = 3; shape[1] = 2; shape[2] = (int64_t)N;
            shape[0] = (int64_t)sw.win_hi[a] - sw.win_lo[a];
            if (shape[0] == 0 && (sw.n_plane > 0 || !r->h_arr[a])) { fail(SHB_E_STATE, "sweep %d did not request profile array %d", sweep, a); return nullptr; }
            return (const char*)r->h_arr[a] + sw.arr_off[a] * r->esz;
        }
    }
}

SHB_API int shb_result_window(const shb_result* r, int32_t which, int32_t sweep, int32_t* row_lo, int32_t* row_hi) {
    SHB_ENTER;
    if (!r || sweep < 0 || sweep >= (int32_t)r->sweeps.size()) return fail(SHB_E_INVALID, "bad result / sweep");
    const int a = which == SHB_ARR_RADIAL ? SHB_A_RADIAL : which - SHB_ARR_IXY;
    if (a < 0 || a >= SHB_N_ARR) return fail(SHB_E_INVALID, "array %d has no row window", which);
    if (row_lo) *row_lo = (int32_t)r->sweeps[sweep].win_lo[a];
    if (row_hi) *row_hi = (int32_t)r->sweeps[sweep].win_hi[a];
    return SHB_OK;
}

SHB_API int shb_sweep_batch(int32_t n_mesh, const double* verts, const int64_t* vert_off, const int64_t* faces,
                    const int64_t* face_off, int32_t n_sweep, const int32_t* sweep_mesh, const double* z_orig,
                    const double* heights, const int64_t* height_off, const int32_t* interp_num, uint32_t outputs_mask,
                    int32_t n_angles, shb_result** out) {
    SHB_ENTER;
    shb_batch* b = nullptr;
    int rc = shb_batch_create(n_mesh, verts, vert_off, faces, face_off, n_sweep, sweep_mesh, z_orig, heights, height_off, interp_num, &b);
    if (rc) return rc;
    rc = shb_batch_run(b, outputs_mask, n_angles, out);
    if (rc) { shb_batch_free(b); return rc; }
    rc = shb_result_fetch(*out, outputs_mask | SHB_OUT_PLANE);
    (*out)->batch = nullptr;
    shb_batch_free(b);                    // stream ordered: the result no longer reads the inputs
    if (rc) { shb_result_free(*out); *out = nullptr; }
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// f3: feature extraction on polar stacks that are still on the device (shb_features.cu)
// ------------------------------------------------------------------------------------------
extern "C" {
int shb_launch_groove_features(const ShbRowSrc* src, int n_src, uint32_t rows, uint32_t maxN, const double* zs, double* feat, double* theta,
                               int32_t* idx, int32_t* cnt, cudaStream_t st);
int shb_launch_groove_points(const ShbRowSrc* src, int n_src, uint32_t rows, const double* zs, const double* bg, int ivar,
                             const double* centroid, double* pts, double* local_theta, cudaStream_t st);
int shb_launch_neck_image(const ShbRowSrc* src, int n_src, uint32_t rows, uint32_t N, const double* bg, double* vals, double* shft,
                          unsigned long long* mm, float* image, double* mm_out, cudaStream_t st);
int shb_launch_forest(const float* X, uint32_t n, uint32_t n_feat, uint32_t n_trees, const uint32_t* root, const int32_t* feature,
                      const float* value, const uint32_t* tchild, const uint32_t* fchild, const float* weight, float* score, cudaStream_t st);
int shb_launch_canal_axes(const ShbCanalJob* jobs, int n_jobs, const double* centroid, const double* z, ShbRowSrc* src, double* axes, cudaStream_t st);
int shb_launch_groove_scale(const ShbRowSrc* src, int n_src, uint32_t max_rows, size_t smem_limit, const double* feat, const int32_t* cnt, float* X,
                            double* stats, cudaStream_t st);
int shb_launch_forest_slots(const float* X, const int32_t* cnt, uint32_t n_rows, uint32_t n_feat, uint32_t n_trees, const uint32_t* root,
                            const int32_t* feature, const float* value, const uint32_t* tchild, const uint32_t* fchild, const float* weight,
                            float* score, cudaStream_t st);
int shb_launch_groove_theta_slots(const ShbRowSrc* src, int n_src, uint32_t max_rows, const double* theta, const int32_t* cnt, const float* score,
                                  float threshold, double* bg, double* dens_max, cudaStream_t st);
}

namespace {
// rows of profile array `a` that the listed sweeps hold in the result, as device row sources; returns total rows or < 0
int64_t row_sources(shb_result* r, int a, int32_t n_sw, const int32_t* sweeps, const double* zs, const double* canal_axes,
                    std::vector<ShbRowSrc>& out, uint32_t& maxN) {
    if (r->esz != 8) return fail(SHB_E_STATE, "feature extraction needs float64 profile arrays (run without SHB_OUT_F32)");
    if (!r->d.prof[a]) return fail(SHB_E_STATE, "profile array %d was not computed by this run", a);
    int64_t rows = 0;
    maxN = 0;
    for (int k = 0; k < n_sw; ++k) {
        const int s = sweeps[k];
        if (s < 0 || s >= (int)r->sweeps.size()) return fail(SHB_E_INVALID, "sweep %d out of range", s);
        const ShbSweep& sw = r->sweeps[s];
        ShbRowSrc q = {};
        q.rows = sw.win_hi[a] - sw.win_lo[a];
        if (q.rows == 0) return fail(SHB_E_STATE, "sweep %d holds no row of profile array %d", s, a);
        q.N = sw.interp_num; q.out_row0 = (uint32_t)rows; q.plane0 = sw.plane_off + sw.win_lo[a];
        q.base = reinterpret_cast<const double*>(r->d.prof[a]) + sw.arr_off[a];
        if (zs) {      // sklearn MinMaxScaler over the rows' z (bicipital_groove.py:92): z * scale_ + min_
            double lo = zs[rows], hi = zs[rows];
            for (uint32_t i = 0; i < q.rows; ++i) { lo = std::min(lo, zs[rows + i]); hi = std::max(hi, zs[rows + i]); }
            double range = hi - lo;
            if (range == 0.0) range = 1.0;
            q.z_scale = 1.0 / range; q.z_min = 0.0 - lo * q.z_scale;
        }
        if (canal_axes) {   // utils.unit_vector(axis[0], axis[1])
            const double* ax = canal_axes + 6 * (size_t)k;
            const double vx = ax[0] - ax[3], vy = ax[1] - ax[4], vz = ax[2] - ax[5];
            const double nrm = std::sqrt(vx * vx + vy * vy + vz * vz);
            q.cu[0] = vx / nrm; q.cu[1] = vy / nrm; q.cu[2] = vz / nrm;
        }
        maxN = std::max(maxN, q.N);
        rows += q.rows;
        out.push_back(q);
    }
    return rows;
}
}  // namespace

struct shb_forest {
    uint32_t n_nodes = 0, n_trees = 0, n_feat = 0;
    uint32_t *root = nullptr, *tchild = nullptr, *fchild = nullptr; int32_t* feature = nullptr; float *value = nullptr, *weight = nullptr;
};

extern "C" {

SHB_API int shb_groove_features(shb_result* r, int32_t n_sw, const int32_t* sweeps, const double* zs, const double* canal_axes,
                                double* feat, double* theta, int32_t* peak_index, int32_t* n_peaks) {
    SHB_ENTER;
    if (!r || !sweeps || !zs || !canal_axes || !feat || !theta || !peak_index || !n_peaks || n_sw <= 0) return fail(SHB_E_INVALID, "null argument");
    std::vector<ShbRowSrc> src; uint32_t maxN = 0;
    const int64_t rows = row_sources(r, 5, n_sw, sweeps, zs, canal_axes, src, maxN);
    if (rows < 0) return (int)rows;
    if (maxN > 1024) return fail(SHB_E_CAPACITY, "interp_num %u > 1024", maxN);
    cudaStream_t st = r->stream ? r->stream : g.stream;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    ShbRowSrc* d_src = nullptr; double *d_zs = nullptr, *d_feat = nullptr, *d_th = nullptr; int32_t *d_idx = nullptr, *d_cnt = nullptr;
    CK(dalloc(&d_src, n_sw, st)); CK(dalloc(&d_zs, rows, st)); CK(dalloc(&d_feat, (size_t)rows * 63, st)); CK(dalloc(&d_th, (size_t)rows * 7, st));
    CK(dalloc(&d_idx, (size_t)rows * 7, st)); CK(dalloc(&d_cnt, rows, st));
    CK(cudaMemcpyAsync(d_src, src.data(), n_sw * sizeof(ShbRowSrc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_zs, zs, rows * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_feat, 0, (size_t)rows * 63 * sizeof(double), st)); CK(cudaMemsetAsync(d_th, 0, (size_t)rows * 7 * sizeof(double), st));
    CK(cudaMemsetAsync(d_idx, 0xFF, (size_t)rows * 7 * sizeof(int32_t), st));
    g.launches += shb_launch_groove_features(d_src, n_sw, (uint32_t)rows, maxN, d_zs, d_feat, d_th, d_idx, d_cnt, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(feat, d_feat, (size_t)rows * 63 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(theta, d_th, (size_t)rows * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(peak_index, d_idx, (size_t)rows * 7 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_peaks, d_cnt, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(d_src, st); dfree(d_zs, st); dfree(d_feat, st); dfree(d_th, st); dfree(d_idx, st); dfree(d_cnt, st);
    return SHB_OK;
}

SHB_API int shb_groove_points(shb_result* r, int32_t n_sw, const int32_t* sweeps, const double* bg_theta, int32_t ivar, const double* zs,
                              double* points, double* local_theta) {
    SHB_ENTER;
    if (!r || !sweeps || !bg_theta || !zs || !points || !local_theta || n_sw <= 0 || ivar < 1) return fail(SHB_E_INVALID, "bad argument");
    std::vector<ShbRowSrc> src; uint32_t maxN = 0;
    const int64_t rows = row_sources(r, 5, n_sw, sweeps, nullptr, nullptr, src, maxN);
    if (rows < 0) return (int)rows;
    cudaStream_t st = r->stream ? r->stream : g.stream;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    ShbRowSrc* d_src = nullptr; double *d_zs = nullptr, *d_bg = nullptr, *d_pts = nullptr, *d_lt = nullptr;
    CK(dalloc(&d_src, n_sw, st)); CK(dalloc(&d_zs, rows, st)); CK(dalloc(&d_bg, n_sw, st)); CK(dalloc(&d_pts, (size_t)rows * 3, st)); CK(dalloc(&d_lt, rows, st));
    CK(cudaMemcpyAsync(d_src, src.data(), n_sw * sizeof(ShbRowSrc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_zs, zs, rows * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_bg, bg_theta, n_sw * sizeof(double), cudaMemcpyHostToDevice, st));
    g.launches += shb_launch_groove_points(d_src, n_sw, (uint32_t)rows, d_zs, d_bg, ivar, r->d.o_centroid, d_pts, d_lt, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(points, d_pts, (size_t)rows * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(local_theta, d_lt, rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(d_src, st); dfree(d_zs, st); dfree(d_bg, st); dfree(d_pts, st); dfree(d_lt, st);
    return SHB_OK;
}

SHB_API int shb_neck_image(shb_result* r, int32_t n_sw, const int32_t* sweeps, const double* bg_theta, float* image, double* itr_shft, double* minmax) {
    SHB_ENTER;
    if (!r || !sweeps || !bg_theta || !image || n_sw <= 0) return fail(SHB_E_INVALID, "bad argument");
    std::vector<ShbRowSrc> src; uint32_t maxN = 0;
    const int64_t rows = row_sources(r, 3, n_sw, sweeps, nullptr, nullptr, src, maxN);
    if (rows < 0) return (int)rows;
    for (auto& q : src) if (q.N != maxN || q.N < 3 || q.N > 1024) return fail(SHB_E_INVALID, "the sweeps of one image call must share interp_num (3..1024)");
    cudaStream_t st = r->stream ? r->stream : g.stream;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    const size_t tot = (size_t)rows * maxN;
    ShbRowSrc* d_src = nullptr; double *d_bg = nullptr, *d_vals = nullptr, *d_shft = nullptr, *d_mmo = nullptr; unsigned long long* d_mm = nullptr; float* d_img = nullptr;
    CK(dalloc(&d_src, n_sw, st)); CK(dalloc(&d_bg, n_sw, st)); CK(dalloc(&d_vals, tot, st)); CK(dalloc(&d_mm, 2 * (size_t)n_sw, st));
    CK(dalloc(&d_mmo, 2 * (size_t)n_sw, st)); CK(dalloc(&d_img, tot, st));
    if (itr_shft) CK(dalloc(&d_shft, 2 * tot, st));
    CK(cudaMemcpyAsync(d_src, src.data(), n_sw * sizeof(ShbRowSrc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_bg, bg_theta, n_sw * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<unsigned long long> mm0(2 * (size_t)n_sw);
    for (int k = 0; k < n_sw; ++k) { mm0[2 * k] = ~0ull; mm0[2 * k + 1] = 0ull; }
    CK(cudaMemcpyAsync(d_mm, mm0.data(), mm0.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    g.launches += shb_launch_neck_image(d_src, n_sw, (uint32_t)rows, maxN, d_bg, d_vals, d_shft, d_mm, d_img, d_mmo, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(image, d_img, tot * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (itr_shft) CK(cudaMemcpyAsync(itr_shft, d_shft, 2 * tot * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (minmax) CK(cudaMemcpyAsync(minmax, d_mmo, 2 * (size_t)n_sw * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(d_src, st); dfree(d_bg, st); dfree(d_vals, st); dfree(d_mm, st); dfree(d_mmo, st); dfree(d_img, st); dfree(d_shft, st);
    return SHB_OK;
}

extern "C" int shb_launch_ray_cast(const double4* vert, const int4* face, int64_t n_face, const double* org, const double* dir, int n_ray, int max_hits,
                                   int32_t* hit_ray, int32_t* hit_tri, double* hit_loc, double* hit_dist, uint32_t* n_hits, cudaStream_t st);

SHB_API int shb_ray_cast(shb_mesh* mesh, int32_t n_ray, const double* origins, const double* directions, int32_t max_hits,
                         int32_t* hit_ray, int32_t* hit_tri, double* hit_loc, double* hit_dist, int32_t* n_hits) {
    SHB_ENTER;
    if (!mesh || !origins || !directions || !hit_ray || !hit_tri || !hit_loc || !n_hits || n_ray <= 0 || n_ray > 65535 || max_hits <= 0)
        return fail(SHB_E_INVALID, "bad argument");
    cudaStream_t st = g.stream;
    if (mesh->stream != st && mesh->ready) CK(cudaStreamWaitEvent(st, mesh->ready, 0));
    double *d_o = nullptr, *d_d = nullptr, *d_loc = nullptr, *d_dist = nullptr; int32_t *d_r = nullptr, *d_t = nullptr; uint32_t* d_n = nullptr;
    CK(dalloc(&d_o, 3 * (size_t)n_ray, st)); CK(dalloc(&d_d, 3 * (size_t)n_ray, st)); CK(dalloc(&d_loc, 3 * (size_t)max_hits, st)); CK(dalloc(&d_dist, max_hits, st));
    CK(dalloc(&d_r, max_hits, st)); CK(dalloc(&d_t, max_hits, st)); CK(dalloc(&d_n, 1, st));
    CK(cudaMemcpyAsync(d_o, origins, 3 * (size_t)n_ray * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_d, directions, 3 * (size_t)n_ray * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_n, 0, sizeof(uint32_t), st));
    g.launches += shb_launch_ray_cast(mesh->vert, mesh->face, mesh->nf, d_o, d_d, n_ray, max_hits, d_r, d_t, d_loc, d_dist, d_n, st);
    CK(cudaGetLastError());
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t keep = std::min<uint32_t>(n, (uint32_t)max_hits);
    CK(cudaMemcpyAsync(hit_ray, d_r, keep * 4, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(hit_tri, d_t, keep * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hit_loc, d_loc, 3 * (size_t)keep * 8, cudaMemcpyDeviceToHost, st));
    if (hit_dist) CK(cudaMemcpyAsync(hit_dist, d_dist, (size_t)keep * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!mesh->last_use) CK(cudaEventCreateWithFlags(&mesh->last_use, cudaEventDisableTiming));
    CK(cudaEventRecord(mesh->last_use, st));
    dfree(d_o, st); dfree(d_d, st); dfree(d_loc, st); dfree(d_dist, st); dfree(d_r, st); dfree(d_t, st); dfree(d_n, st);
    *n_hits = (int32_t)n;
    return n > (uint32_t)max_hits ? fail(SHB_E_CAPACITY, "%u hits, room for %d", n, max_hits) : SHB_OK;
}

SHB_API int shb_forest_create(int32_t n_nodes, int32_t n_trees, int32_t n_features, const uint32_t* root, const int32_t* feature,
                              const float* value, const uint32_t* true_child, const uint32_t* false_child, const float* weight,
                              shb_forest** out) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (!out || n_nodes <= 0 || n_trees <= 0 || n_features <= 0 || !root || !feature || !value || !true_child || !false_child || !weight)
        return fail(SHB_E_INVALID, "bad argument");
    for (int i = 0; i < n_nodes; ++i) {
        if (feature[i] >= n_features) return fail(SHB_E_INVALID, "node %d tests feature %d of %d", i, feature[i], n_features);
        if (feature[i] >= 0 && (true_child[i] >= (uint32_t)n_nodes || false_child[i] >= (uint32_t)n_nodes || true_child[i] <= (uint32_t)i ||
                                false_child[i] <= (uint32_t)i))
            return fail(SHB_E_INVALID, "node %d: children must follow their parent", i);       // guarantees that every walk ends
    }
    for (int t = 0; t < n_trees; ++t) if (root[t] >= (uint32_t)n_nodes) return fail(SHB_E_INVALID, "bad root");
    cudaStream_t st = g.stream;
    shb_forest* f = new shb_forest;
    f->n_nodes = n_nodes; f->n_trees = n_trees; f->n_feat = n_features;
    CK(dalloc(&f->root, n_trees, st)); CK(dalloc(&f->feature, n_nodes, st)); CK(dalloc(&f->value, n_nodes, st));
    CK(dalloc(&f->tchild, n_nodes, st)); CK(dalloc(&f->fchild, n_nodes, st)); CK(dalloc(&f->weight, n_nodes, st));
    CK(cudaMemcpyAsync(f->root, root, n_trees * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(f->feature, feature, n_nodes * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f->value, value, n_nodes * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(f->tchild, true_child, n_nodes * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f->fchild, false_child, n_nodes * 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(f->weight, weight, n_nodes * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    *out = f;
    return SHB_OK;
}

SHB_API int shb_forest_predict(shb_forest* f, const float* X, int32_t n, float* score) {
    SHB_ENTER;
    if (!f || !X || !score || n < 0) return fail(SHB_E_INVALID, "bad argument");
    if (n == 0) return SHB_OK;
    cudaStream_t st = g.stream;
    float *dX = nullptr, *ds = nullptr;
    CK(dalloc(&dX, (size_t)n * f->n_feat, st)); CK(dalloc(&ds, n, st));
    CK(cudaMemcpyAsync(dX, X, (size_t)n * f->n_feat * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ds, 0, (size_t)n * 4, st));
    g.launches += shb_launch_forest(dX, n, f->n_feat, f->n_trees, f->root, f->feature, f->value, f->tchild, f->fchild, f->weight, ds, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(score, ds, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(dX, st); dfree(ds, st);
    return SHB_OK;
}

extern "C" int shb_launch_groove_theta(const long long* off, int n_set, uint32_t max_peaks, const double* peak_theta, const float* proba1,
                                       float threshold, double* bg, double* dens_max, cudaStream_t st);

SHB_API int shb_groove_theta(int32_t n_set, const int64_t* off, const double* peak_theta, const float* proba1, float threshold,
                             double* bg_theta, double* density_max) {
    SHB_ENTER;
    if (!g.inited) return fail(SHB_E_STATE, "shb_init not called");
    if (n_set < 0 || !off || !bg_theta || (off[n_set] > 0 && (!peak_theta || !proba1))) return fail(SHB_E_INVALID, "bad argument");
    if (n_set == 0) return SHB_OK;
    const int64_t n = off[n_set];
    int64_t mx = 0;
    for (int i = 0; i < n_set; ++i) { if (off[i + 1] < off[i]) return fail(SHB_E_INVALID, "offsets decrease"); mx = std::max(mx, off[i + 1] - off[i]); }
    if ((size_t)mx * 8 > g.smem_optin - 1024) return fail(SHB_E_CAPACITY, "%lld peaks in one set", (long long)mx);
    cudaStream_t st = g.stream;
    long long* d_off = nullptr; double *d_th = nullptr, *d_bg = nullptr; float* d_pr = nullptr;
    CK(dalloc(&d_off, n_set + 1, st)); CK(dalloc(&d_th, n, st)); CK(dalloc(&d_pr, n, st)); CK(dalloc(&d_bg, 2 * (size_t)n_set, st));
    CK(cudaMemcpyAsync(d_off, off, (n_set + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (n) {
        CK(cudaMemcpyAsync(d_th, peak_theta, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_pr, proba1, n * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    g.launches += shb_launch_groove_theta(d_off, n_set, (uint32_t)mx, d_th, d_pr, threshold, d_bg, d_bg + n_set, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(bg_theta, d_bg, n_set * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (density_max) CK(cudaMemcpyAsync(density_max, d_bg + n_set, n_set * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dfree(d_off, st); dfree(d_th, st); dfree(d_pr, st); dfree(d_bg, st);
    return SHB_OK;
}

SHB_API int shb_landmark_front(shb_result* r, const shb_landmark_args* a) {
    SHB_ENTER;
    if (!r || !a || a->n_bones <= 0 || !a->full_sweeps || !a->prox_sweeps || !a->canal_z || !a->canal_half || !a->groove_zs || !a->forest ||
        a->canal_hi <= a->canal_lo || a->canal_lo < 0 || a->ivar < 1)
        return fail(SHB_E_INVALID, "bad argument");
    const int nb = a->n_bones;
    const shb_forest* f = a->forest;
    if (f->n_feat != 9) return fail(SHB_E_INVALID, "the groove forest must read 9 features, this one reads %u", f->n_feat);
    std::vector<ShbRowSrc> gsrc, isrc; uint32_t gN = 0, iN = 0;
    const int64_t grows = row_sources(r, 5, nb, a->prox_sweeps, a->groove_zs, nullptr, gsrc, gN);
    if (grows < 0) return (int)grows;
    const int64_t irows = row_sources(r, 3, nb, a->prox_sweeps, nullptr, nullptr, isrc, iN);
    if (irows < 0) return (int)irows;
    if (gN > 1024) return fail(SHB_E_CAPACITY, "interp_num %u > 1024", gN);
    for (auto& q : isrc) if (q.N != iN || q.N < 3 || q.N > 1024) return fail(SHB_E_INVALID, "the sweeps of one image call must share interp_num (3..1024)");
    const uint32_t crow = (uint32_t)(a->canal_hi - a->canal_lo);
    std::vector<ShbCanalJob> jobs(nb);
    uint32_t max_rows = 0;
    for (int k = 0; k < nb; ++k) {
        const int s = a->full_sweeps[k];
        if (s < 0 || s >= (int)r->sweeps.size()) return fail(SHB_E_INVALID, "sweep %d out of range", s);
        const ShbSweep& sw = r->sweeps[s];
        if ((uint32_t)a->canal_hi > sw.n_plane) return fail(SHB_E_INVALID, "canal rows %d..%d of a sweep of %u planes", a->canal_lo, a->canal_hi, sw.n_plane);
        jobs[k] = {sw.plane_off + (uint32_t)a->canal_lo, crow, (uint32_t)k * crow, 0u, a->canal_half[k]};
        max_rows = std::max(max_rows, gsrc[k].rows);
    }
    if ((size_t)max_rows * 7 * 8 + 1024 > g.smem_optin) return fail(SHB_E_CAPACITY, "%u rows in one groove window", max_rows);
    cudaStream_t st = r->stream ? r->stream : g.stream;
    if (r->done) CK(cudaStreamWaitEvent(st, r->done, 0));
    const size_t G = (size_t)grows, tot = (size_t)irows * iN;
    // one block of device memory, carved
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_gsrc = carve(nb * sizeof(ShbRowSrc)), o_isrc = carve(nb * sizeof(ShbRowSrc)), o_jobs = carve(nb * sizeof(ShbCanalJob)),
                 o_cz = carve((size_t)nb * crow * 8), o_zs = carve(G * 8), o_axes = carve((size_t)nb * 48), o_feat = carve(G * 63 * 8),
                 o_th = carve(G * 7 * 8), o_idx = carve(G * 7 * 4), o_cnt = carve(G * 4), o_X = carve(G * 63 * 4), o_score = carve(G * 7 * 4),
                 o_stats = carve((size_t)nb * 18 * 8), o_bg = carve((size_t)nb * 16), o_pts = carve(G * 24), o_lt = carve(G * 8),
                 o_vals = carve(tot * 8), o_mm = carve((size_t)nb * 16), o_mmo = carve((size_t)nb * 16), o_img = carve(tot * 4);
    unsigned char* D = nullptr;
    CK(dalloc_bytes(reinterpret_cast<void**>(&D), off, st));
    // small inputs: one pinned staging block, one copy each (they are a few KB)
    Staging stg;
    struct Bail { Staging& s; unsigned char*& dev; cudaStream_t st; bool armed = true;      // an early return: wait, then give back
        ~Bail() { if (!armed) return; if (!s.bufs.empty()) { cudaStreamSynchronize(st); s.release(); } if (dev) { cudaFreeAsync(dev, st); dev = nullptr; } } } bail{stg, D, st};
    CK(stg.copy(D + o_gsrc, gsrc.data(), nb * sizeof(ShbRowSrc), st)); CK(stg.copy(D + o_isrc, isrc.data(), nb * sizeof(ShbRowSrc), st));
    CK(stg.copy(D + o_jobs, jobs.data(), nb * sizeof(ShbCanalJob), st)); CK(stg.copy(D + o_cz, a->canal_z, (size_t)nb * crow * 8, st));
    CK(stg.copy(D + o_zs, a->groove_zs, G * 8, st));
    std::vector<unsigned long long> mm0(2 * (size_t)nb);
    for (int k = 0; k < nb; ++k) { mm0[2 * k] = ~0ull; mm0[2 * k + 1] = 0ull; }
    CK(stg.copy(D + o_mm, mm0.data(), mm0.size() * 8, st));
    CK(cudaMemsetAsync(D + o_feat, 0, o_idx - o_feat, st));                      // feat, theta
    CK(cudaMemsetAsync(D + o_idx, 0xFF, G * 7 * 4, st));
    CK(cudaMemsetAsync(D + o_score, 0, G * 7 * 4, st));
    ShbRowSrc* d_gsrc = reinterpret_cast<ShbRowSrc*>(D + o_gsrc); ShbRowSrc* d_isrc = reinterpret_cast<ShbRowSrc*>(D + o_isrc);
    double* d_bg = reinterpret_cast<double*>(D + o_bg);
    g.launches += shb_launch_canal_axes(reinterpret_cast<ShbCanalJob*>(D + o_jobs), nb, r->d.o_centroid, reinterpret_cast<double*>(D + o_cz), d_gsrc,
                                        reinterpret_cast<double*>(D + o_axes), st);
    g.launches += shb_launch_groove_features(d_gsrc, nb, (uint32_t)G, gN, reinterpret_cast<double*>(D + o_zs), reinterpret_cast<double*>(D + o_feat),
                                             reinterpret_cast<double*>(D + o_th), reinterpret_cast<int32_t*>(D + o_idx), reinterpret_cast<int32_t*>(D + o_cnt), st);
    g.launches += shb_launch_groove_scale(d_gsrc, nb, max_rows, g.smem_optin, reinterpret_cast<double*>(D + o_feat), reinterpret_cast<int32_t*>(D + o_cnt),
                                          reinterpret_cast<float*>(D + o_X), reinterpret_cast<double*>(D + o_stats), st);
    g.launches += shb_launch_forest_slots(reinterpret_cast<float*>(D + o_X), reinterpret_cast<int32_t*>(D + o_cnt), (uint32_t)G, f->n_feat, f->n_trees, f->root,
                                          f->feature, f->value, f->tchild, f->fchild, f->weight, reinterpret_cast<float*>(D + o_score), st);
    g.launches += shb_launch_groove_theta_slots(d_gsrc, nb, max_rows, reinterpret_cast<double*>(D + o_th), reinterpret_cast<int32_t*>(D + o_cnt),
                                                reinterpret_cast<float*>(D + o_score), a->threshold, d_bg, d_bg + nb, st);
    g.launches += shb_launch_groove_points(d_gsrc, nb, (uint32_t)G, reinterpret_cast<double*>(D + o_zs), d_bg, a->ivar, r->d.o_centroid,
                                           reinterpret_cast<double*>(D + o_pts), reinterpret_cast<double*>(D + o_lt), st);
    g.launches += shb_launch_neck_image(d_isrc, nb, (uint32_t)irows, iN, d_bg, reinterpret_cast<double*>(D + o_vals), nullptr,
                                        reinterpret_cast<unsigned long long*>(D + o_mm), reinterpret_cast<float*>(D + o_img), reinterpret_cast<double*>(D + o_mmo), st);
    CK(cudaGetLastError());
    // the copies back: behind the kernels on the compute stream, or (SHB_LF_NO_WAIT) on the copy stream behind an event, so that
    // whatever the caller enqueues next on the compute stream runs beside them
    const bool nowait = (a->flags & SHB_LF_NO_WAIT) != 0;
    cudaStream_t cs = st;
    if (nowait) {
        cudaEvent_t ev = nullptr;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventRecord(ev, st));
        CK(cudaStreamWaitEvent(g.copy, ev, 0));
        cudaEventDestroy(ev);                                       // released once it has completed
        cs = g.copy;
    }
    auto back = [&](void* dst, size_t o, size_t bytes) -> cudaError_t { return dst && bytes ? cudaMemcpyAsync(dst, D + o, bytes, cudaMemcpyDeviceToHost, cs) : cudaSuccess; };
    CK(back(a->canal_axes, o_axes, (size_t)nb * 48)); CK(back(a->feat, o_feat, G * 63 * 8)); CK(back(a->peak_theta, o_th, G * 7 * 8));
    CK(back(a->peak_index, o_idx, G * 7 * 4)); CK(back(a->n_peaks, o_cnt, G * 4)); CK(back(a->X, o_X, G * 63 * 4)); CK(back(a->proba1, o_score, G * 7 * 4));
    CK(back(a->scaler, o_stats, (size_t)nb * 18 * 8)); CK(back(a->bg_theta, o_bg, (size_t)nb * 8)); CK(back(a->points, o_pts, G * 24));
    CK(back(a->local_theta, o_lt, G * 8)); CK(back(a->image, o_img, tot * 4)); CK(back(a->minmax, o_mmo, (size_t)nb * 16));
    bail.armed = false;
    if (nowait) {
        dfree(D, cs);                                               // stream ordered: behind the copies
        r->lf_staged.insert(r->lf_staged.end(), stg.bufs.begin(), stg.bufs.end());      // the staged inputs live until the wait
        stg.bufs.clear();
        r->pending = true;
        return SHB_OK;
    }
    CK(cudaStreamSynchronize(st));
    stg.release();
    dfree(D, st);
    return SHB_OK;
}

SHB_API int shb_landmark_wait(shb_result* r) {
    SHB_ENTER;
    if (!r) return fail(SHB_E_INVALID, "null result");
    if (r->pending) { CK(cudaStreamSynchronize(g.copy)); r->pending = false; }
    for (void* p : r->lf_staged) pinned_put(p);
    r->lf_staged.clear();
    return SHB_OK;
}

SHB_API int shb_forest_free(shb_forest* f) {
    SHB_ENTER;
    if (!f) return SHB_OK;
    cudaStream_t st = g.stream;
    dfree(f->root, st); dfree(f->feature, st); dfree(f->value, st); dfree(f->tchild, st); dfree(f->fchild, st); dfree(f->weight, st);
    delete f;
    return SHB_OK;
}

}  // extern "C"
