// Shared declarations of the sm_100a slicing backend (host + device).
// Path replaced: reference src/shoulder/humerus/slice.py:21-147 and the trimesh calls under it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SHB_TOL_MERGE 1e-8     // trimesh.constants.tol.merge  (sign classification band)
#define SHB_HIT_FACE   0x1FFFFFFFu   // face bits of a hit record's x word

// sweep descriptor, one per (mesh, height list) pair — device resident
#define SHB_N_ARR 7          // windowed per-plane outputs of the resample kernel: the six profile arrays + the radius image
#define SHB_A_RADIAL 6
struct ShbSweep {
    double   z_orig;      // plane origin z (slice.py:18)
    uint64_t arr_off[SHB_N_ARR];   // element offset of this sweep's block in output array a: (rows,2,N) profiles, (rows,A) radius image
    uint32_t win_lo[SHB_N_ARR];    // rows [win_lo, win_hi) of the sweep are computed and delivered for array a (the consumers'
    uint32_t win_hi[SHB_N_ARR];    // _cutoff windows, slice.py:157-164); lo == hi: the sweep does not ask for array a
    uint32_t plane_off;   // first global plane index
    uint32_t n_plane;
    uint32_t face_off;    // first global face index of the sweep's mesh
    uint32_t n_face;
    uint32_t item_off;    // first (sweep, triangle) work item
    uint32_t interp_num;  // N
    uint32_t mesh;
    uint32_t pad;
};

// per-plane record written by the stitch kernel (original plane order)
struct ShbPlaneMeta {
    double   bounds[4];   // minx, miny, maxx, maxy   (Path2D.bounds)
    double   centroid[2]; // bounds midpoint           (Path2D.centroid, slice.py:38)
    double   area1;       // slice.py:49-60
    uint32_t n_seg;
    uint32_t n_ent;       // closed contours of the plane (entities that form a path)
    uint32_t status;      // SHB_ST_*
    uint32_t sel_contour; // contour the outline is taken from (slice.py:70-76)
    uint32_t sel_start;   // its first point, relative to the plane's point region
    uint32_t sel_len;     // its point count including the closing duplicate
    uint32_t n_pts;       // points written for the plane (closing duplicates included)
    uint32_t n_open;      // open chains: entities without a contour; len(Path2D.entities) = n_ent + n_open
    // what the resample kernel needs of the plane's sweep, so that it starts from ONE record instead of the
    // plane -> sorted plane -> sweep -> descriptor chain of dependent loads
    uint32_t interp_num;  // N of the plane's sweep
    uint32_t arr_mask;    // bit a: output array a is wanted for this plane (requested by its sweep and inside the window)
    uint64_t arr_row[SHB_N_ARR];   // element offset of the plane's (2, N) block in profile array a / of its A rays in the radius image
    uint64_t sel_pt;      // index (in points) of the outline's first point in ShbDev::pts
};

struct ShbDev {
    // ---- batch-static inputs
    const double4*  vert;        // (x, y, z, 0) per global vertex
    const double*   vz;          // z only (dense, for the bucket / intersect kernels)
    const int4*     face;        // global vertex ids (a, b, c, 0) per global face
    const uint32_t* adj;         // [T][4] global ids of the faces across edges (v0 v1), (v1 v2), (v2 v0); SHB_NIL = none
    const ShbSweep* sweep;
    const uint32_t* item_off;    // [n_sweep+1] prefix of faces per sweep
    const double*   h_sorted;    // [G] heights ascending within each sweep
    const double*   h_orig;      // [G] heights in caller order (indexed by original plane)
    const double*   oz;          // [G] plane z = fl(z_orig + height), caller order (trimesh: new_origin = origin + normal * height)
    const uint32_t* plane_out;   // [G] sorted plane -> original global plane
    const uint32_t* plane_in;    // [G] original global plane -> sorted plane
    const uint32_t* plane_sweep; // [G] sweep of sorted plane
    const uint32_t* stitch_order;// [G] CTA b of the stitch launch takes plane stitch_order[b]: sweep ends first (nullptr: identity)
    const uint32_t* resample_order; // [n_resample] planes the resample launch covers (nullptr: stitch_order, all planes)
    uint32_t n_resample;
    uint32_t n_sweep, n_plane /*G*/, n_item;
    // ---- per-run scratch
    uint32_t* item_lo;    // [n_item] first sorted plane of the triangle's range
    uint32_t* item_span;  // [n_item]
    uint32_t* inc;        // [G]   #triangles whose range starts at (sorted) plane = counting-sort histogram
    uint32_t* sort_off;   // [G+1] bucket offsets
    uint32_t* sort_cur;   // [G]   bucket cursors (scatter), then hit-list cursors (intersect) = exact hits per sorted plane
    uint32_t* dec;        // [G+1] #triangles whose range ends before (sorted) plane
    uint32_t* cnt;        // [G]   candidate triangles per sorted plane (ranges covering it) >= exact hits
    uint32_t* cap_off;    // [G+1] hit-list offsets by candidate capacity, original plane order
    uint32_t* cap_sorted; // [G]   the same offsets indexed by sorted plane (what the intersect kernel has at hand)
    unsigned long long* scan_state;   // [4][ceil(G/4096)] tile aggregates of the four scans of a run + 4 tile tickets, zeroed per run
    uint32_t* totals;     // [16]  M, bad, cap, S, maxn, nbig, ncont, npts, ndecl, ndup
    uint4*    rec;        // [n_item] bucketed triangles (face, lo, span, sweep); first M are live
    uint4*    hits;       // [S]   per-plane hit records, caller plane order: x = mesh-local face id | tag << 29 | (lone vertex above) << 31,
                          //       basic crossings (tag = position of the lone vertex u): y = mesh-local id of the face the contour
                          //       continues in (across the END edge), z = u, w = the other vertex of the END edge;
                          //       tag 3 = a vertex on the plane (the stitcher classifies such faces itself)
    uint32_t* seg_off;    // [G+1] exact segment offsets, original plane order
    uint32_t* big_list;   // [G]   planes too large for shared memory
    uint32_t* decl_list;  // [2 G] planes the warp stitcher declined (several contours, on-plane vertices, open / non-manifold
                          //       nodes, inconsistent winding, too large): the CTA stitcher takes them; count in totals[SHB_T_NDECL]
    uint32_t* dup_list;   // [G]   planes with two consecutive contour nodes closer than Path.merge_vertices' grid: the
                          //       merge pass takes them; count in totals[SHB_T_NDUP]
    // ---- outputs (device)
    ShbPlaneMeta* meta;   // [G]  (kernel-internal AoS; the arrays below are what the host reads)
    int32_t*  o_nseg;     // [G]
    int32_t*  o_nent;     // [G]
    uint32_t* o_status;   // [G]
    double*   o_bounds;   // [G][4]
    double*   o_centroid; // [G][2]
    double*   o_area1;    // [G]
    int32_t*  o_sel;      // [G][2]
    unsigned long long* totals64;  // [2] 64-bit candidate total (overflow guard)
    int32_t*  face_index; // [S]
    double*   segments;   // [S][2][2]
    double*   pts;        // [2S][2] capacity layout: plane p owns [2*seg_off[p], 2*seg_off[p+1])
    uint32_t* ct_start;   // [S] capacity layout: plane p owns [seg_off[p], seg_off[p+1])
    uint32_t* ct_len;     // [S]
    double*   ct_area;    // [S]
    void*     prof[6];    // ixy, ixy_centered, itr, itr_start, itr_centered, itr_centered_start (f64, or f32 with SHB_OUT_F32)
    void*     radial;
    const double2* angle_cs;     // [n_angles] (cos, sin) of theta_k = -pi + 2*pi*k/n_angles
    unsigned char* scratch;      // global workspaces for oversized planes
    size_t    scratch_stride;
    uint32_t  n_angles;
    uint32_t  outputs_mask;
    uint32_t  stitch_cap;        // largest n handled in shared memory
    uint32_t  resample_cap;      // largest point count handled in shared memory
    uint32_t  debug;             // test hooks: bit 0 = radius image by the all-candidates path on every plane
};

// one sweep's window of a profile array, as the feature kernels (shb_features.cu) see it
struct ShbRowSrc {
    const double* base;        // first row: (rows, 2, N) float64
    uint32_t rows, N;
    uint32_t out_row0;         // first output row of this sweep in the concatenated outputs
    uint32_t plane0;           // global plane index of the window's first row (centroids)
    double z_scale, z_min;     // MinMaxScaler of the rows' z: z * z_scale + z_min
    double cu[3];              // unit vector of the canal axis as Canal.axis() returns it (bicipital_groove.py:70)
};

// one bone of shb_landmark_front's canal fit: rows [plane0, plane0 + rows) of the plane records, their z at z[z_off ...]
struct ShbCanalJob { uint32_t plane0, rows, z_off, pad; double half_len; };

#ifdef __CUDACC__
// eigenvectors of a symmetric 3x3 matrix by cyclic Jacobi rotations: evec columns, eval diagonal
__device__ inline void shb_jacobi3(double a[3][3], double v[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
            }
    }
}
#endif

enum { SHB_T_M = 0, SHB_T_BAD = 1, SHB_T_CAP = 2, SHB_T_S = 3, SHB_T_MAXN = 4, SHB_T_NBIG = 5,
       SHB_T_NCONT = 6, SHB_T_NPTS = 7, SHB_T_NDECL = 8, SHB_T_NDUP = 9, SHB_T_NDECL2 = 10 };

// bytes of workspace the stitch kernel needs for a plane with n segments
__host__ __device__ inline uint32_t shb_pow2_ge(uint32_t x) {
#ifdef __CUDA_ARCH__
    return x <= 1u ? 1u : 1u << (32 - __clz((int)(x - 1u)));
#else
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
#endif
}
__host__ __device__ inline uint32_t shb_hash_size(uint32_t n) { return shb_pow2_ge(2 * n + n / 2 + 1); }
// fast path of the stitcher: staged hit records, later the start-node coordinates [n x 16 B] | node keys [E x 8] |
// mate [E x 4] | record head words [n x 4] | hash table, later next / prev / jump words [max(4H, 12n)]
__host__ __device__ inline size_t shb_stitch_fast_table_bytes(uint32_t n) {
    size_t t = 4 * (size_t)shb_hash_size(n), p = 12 * (size_t)n;
    return ((t > p ? t : p) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t shb_stitch_fast_bytes(uint32_t n) {
    return 16 * (size_t)n + 16 * (size_t)n + 8 * (size_t)n + 4 * (size_t)n + shb_stitch_fast_table_bytes(n) + 16;
}
__host__ __device__ inline size_t shb_stitch_ws_bytes(uint32_t n) {
    size_t E = 2 * (size_t)n;
    size_t c1 = 4 * (size_t)shb_pow2_ge(n) + 4 * (size_t)shb_hash_size(n);   // sort keys + hash table
    size_t c2 = 12 * E;                                                       // jump pairs + heads
    size_t c = c1 > c2 ? c1 : c2;
    size_t g = 4 * E /*mate*/ + 8 * E /*node key | rank key | area acc*/ + c + 4 * E /*contour list + starts*/ + n /*lone-vertex signs*/ + 96;
    size_t f = shb_stitch_fast_bytes(n);
    return g > f ? g : f;
}
// outline (16 B) + chord lengths / vertex angles (8 B) per point; region X = per-edge slopes, later theta / r;
// region Y = resampled x / y, later the ray accumulators (8 B) + ray owners (4 B); sort keys only when a theta-sorted array is requested
__host__ __device__ inline size_t shb_resample_x_bytes(uint32_t npts, uint32_t N) { return 16 * (size_t)(npts > N ? npts : N); }
__host__ __device__ inline size_t shb_resample_y_bytes(uint32_t N, uint32_t A) { return 16 * (size_t)N > 12 * (size_t)A ? 16 * (size_t)N : 12 * (size_t)A; }
__host__ __device__ inline size_t shb_resample_ws_bytes(uint32_t npts, uint32_t N, uint32_t A, bool sorted = true) {
    return 24 * ((size_t)npts + 1) + 16 + shb_resample_x_bytes(npts, N) + shb_resample_y_bytes(N, A) +
           (sorted ? 12 * (size_t)shb_pow2_ge(N) : 0) + 64;
}
// byte offsets of the resample workspace's arrays, fixed per LAUNCH (capacity npts points, N samples, A rays): they
// reach the kernel as parameters, so no thread derives an address from its plane's own point count
struct ShbRsLayout { uint32_t dd, x, rr, y, sy, own, skeys, svals; };
inline ShbRsLayout shb_resample_layout(uint32_t npts, uint32_t N, uint32_t A) {
    ShbRsLayout L;
    L.dd = 16u * (npts + 1);
    L.x = (24u * (npts + 1) + 15u) & ~15u;
    L.rr = L.x + 8u * N;
    L.y = L.x + (uint32_t)shb_resample_x_bytes(npts, N);
    L.sy = L.y + 8u * N;
    L.own = L.y + 8u * A;
    L.skeys = L.y + (uint32_t)shb_resample_y_bytes(N, A);
    L.svals = L.skeys + 8u * shb_pow2_ge(N);
    return L;
}

#ifdef __cplusplus
extern "C" {
#endif
// launch wrappers (shb_kernels.cu); every one returns the number of kernels it enqueued
int shb_launch_prep_mesh(const double* verts_in, const int64_t* faces_in, const int64_t* vert_off,
                         const int64_t* face_off, int n_mesh, int64_t n_vert, int64_t n_face,
                         double4* vert, double* vz, int4* face, uint32_t* bad, cudaStream_t st);
int shb_launch_adjacency(const int4* face, int64_t n_face, unsigned long long* keys, uint32_t* cnt, uint32_t* own, uint32_t* hslot,
                         uint32_t hsize, uint32_t* adj, cudaStream_t st);
int shb_launch_bucket(const ShbDev& d, cudaStream_t st);
int shb_launch_scan_planes(const ShbDev& d, cudaStream_t st);
int shb_launch_scatter(const ShbDev& d, cudaStream_t st);
int shb_launch_intersect(const ShbDev& d, cudaStream_t st);
int shb_launch_scan_candidates(const ShbDev& d, cudaStream_t st);
int shb_launch_scan_counts(const ShbDev& d, cudaStream_t st);
int shb_launch_stitch(const ShbDev& d, uint32_t maxcand, uint32_t avgn, uint32_t max_faces, size_t smem_budget, int n_sm, cudaStream_t st,
                      cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join, cudaEvent_t ev_mid, bool partial_sweeps);
int shb_launch_resample(const ShbDev& d, uint32_t maxcand, uint32_t avgn, uint32_t maxN, int n_sm, cudaStream_t st);
int shb_launch_compact(const ShbDev& d, const uint32_t* ct_off, const uint32_t* pt_off,
                       double* pts_out, int64_t* ctpt_out, double* ctarea_out, cudaStream_t st);
int shb_launch_scan_contours(const ShbDev& d, uint32_t* ct_off, uint32_t* pt_off, cudaStream_t st);
int shb_launch_publish(const uint32_t* src, int n, const unsigned long long* src64, uint32_t* dst_host,
                       unsigned long long* dst64_host, cudaStream_t st);
#ifdef __cplusplus
}
#endif
