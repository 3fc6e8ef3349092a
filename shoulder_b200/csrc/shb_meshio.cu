// Mesh input on the device (scope row f2): binary STL bytes -> welded mesh -> oriented frame, without a host pass over
// the triangles.  What the reference does on the CPU before the slicing path starts:
//
//   mesh.py:24   trimesh.load_mesh(stl)   = exchange.stl.load_stl (80-byte header, uint32 count, 50-byte records: normal,
//                three float32 corners, attribute word; faces = arange(3T).reshape(-1, 3)) followed by
//                Trimesh(process=True) -> merge_vertices: corners whose coordinates round equal at 1e-8
//                (round(v * 1e8).astype(int64), tol.merge) become one vertex, numbered by first occurrence, carrying the
//                coordinates of that first occurrence; the face order is kept.
//   mesh.py:82   mesh.apply_obb()         qhull's minimum-volume box — not restatable here (frame note):
//                the frame is the PCA stand-in every config of this repo uses (shoulder_b200/meshio.py PcaObb: principal
//                axes, smallest variance -> x, largest -> z, deterministic signs, AABB centred on the origin), followed by
//                the reference's own end test (mesh.py:91-117: the rounder end is the humeral head and goes to +z), judged
//                on a thin vertex band at 0.95 of each z limit with an algebraic circle fit.
//
//   k_weld_insert   corner-parallel open-addressing hash on the three rounded coordinates; every cell keeps the smallest
//                   corner index that fell into it (first occurrence)
//   k_weld_count / k_weld_tiles / k_weld_number   exclusive scan of the first-occurrence flags = vertex numbering
//   k_weld_emit     faces (int4) and vertices (double4 + dense z) in the layout the slicing kernels read
//   k_frame_acc<S> / k_frame_step<S>   five reductions over the vertices (sums, centred second moments, projected
//                   bounds, circle-fit normal equations of the two end bands, their residuals), each followed by a
//                   one-thread step (mean; 3x3 Jacobi eigenvectors; centring; 3x3 solve; end choice and final matrix)
//   k_frame_apply   vertices into the frame, in place
//
// Reductions use a fixed grid and a fixed combination order, so a mesh always gets the same frame bits.
#include "shb_common.cuh"
#include "../../include/shoulder_b200.h"
#include <math_constants.h>

#define SHB_NILU 0xFFFFFFFFu
#define SHB_STL_HEADER 84u
#define SHB_STL_RECORD 50u
#define SHB_WELD_TILE 2048u        // corners per CTA of the numbering scan (256 threads x 8)
#define SHB_FRAME_K 20             // accumulators per thread of the widest reduction

// corner c of a binary STL: record c / 3, corner c % 3; the floats sit at even byte offsets (84 + 50 t + 12 (1 + k)), so
// they are read as 16-bit halves
__device__ __forceinline__ float3 shb_stl_corner(const unsigned char* __restrict__ stl, uint32_t c) {
    const unsigned short* h = reinterpret_cast<const unsigned short*>(stl + SHB_STL_HEADER + (size_t)SHB_STL_RECORD * (c / 3u) + 12u + 12u * (c % 3u));
    uint32_t w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = (uint32_t)__ldg(h + 2 * i) | ((uint32_t)__ldg(h + 2 * i + 1) << 16);
    return make_float3(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]));
}
// merge_vertices' cell of a coordinate: round(v * 10^8) as int64 (numpy rounds half to even, like cvt.rni)
__device__ __forceinline__ long long shb_weld_cell(float v) { return __double2ll_rn(__dmul_rn((double)v, 1.0e8)); }
__device__ __forceinline__ uint32_t shb_weld_hash(long long a, long long b, long long c) {
    uint64_t h = (uint64_t)a * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h += (uint64_t)b; h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32; h += (uint64_t)c; h *= 0x94D049BB133111EBull;
    h ^= h >> 31;
    return (uint32_t)h;
}

__global__ void __launch_bounds__(256) k_weld_insert(const unsigned char* __restrict__ stl, uint32_t n_corner, uint32_t* __restrict__ table,
                                                     uint32_t* __restrict__ first, uint32_t mask, uint32_t* __restrict__ slot_of) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c >= n_corner) return;
    const float3 p = shb_stl_corner(stl, c);
    const long long q0 = shb_weld_cell(p.x), q1 = shb_weld_cell(p.y), q2 = shb_weld_cell(p.z);
    uint32_t h = shb_weld_hash(q0, q1, q2) & mask;
    for (;;) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t*>(table + h);
        if (cur == SHB_NILU) {
            cur = atomicCAS(table + h, SHB_NILU, c);
            if (cur == SHB_NILU) break;                           // this corner names the cell
        }
        const float3 r = shb_stl_corner(stl, cur);
        if (shb_weld_cell(r.x) == q0 && shb_weld_cell(r.y) == q1 && shb_weld_cell(r.z) == q2) break;
        h = (h + 1u) & mask;
    }
    slot_of[c] = h;
    atomicMin(first + h, c);
}

__device__ __forceinline__ uint32_t shb_block_exscan_u32(uint32_t v, uint32_t* total, uint32_t* sh /*[9]*/) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += y; }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < 8; ++i) { const uint32_t t = sh[i]; sh[i] = run; run += t; } sh[8] = run; }
    __syncthreads();
    *total = sh[8];
    return sh[w] + x - v;
}

// pass 1: first occurrences per tile
__global__ void __launch_bounds__(256) k_weld_count(const uint32_t* __restrict__ first, const uint32_t* __restrict__ slot_of, uint32_t n_corner,
                                                    uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t sh[9];
    const uint32_t base = blockIdx.x * SHB_WELD_TILE + threadIdx.x * 8u;
    uint32_t n = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8u; ++i) { const uint32_t c = base + i; if (c < n_corner && first[slot_of[c]] == c) ++n; }
    uint32_t tot;
    shb_block_exscan_u32(n, &tot, sh);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
// pass 2: exclusive scan of the tile sums (one CTA walks them in rounds of 256), total -> n_vert_out
__global__ void __launch_bounds__(256) k_weld_tiles(uint32_t* __restrict__ tile_sum, uint32_t n_tile, uint32_t* __restrict__ n_vert_out) {
    __shared__ uint32_t sh[9];
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < n_tile; t0 += 256u) {
        const uint32_t t = t0 + threadIdx.x;
        const uint32_t v = t < n_tile ? tile_sum[t] : 0u;
        uint32_t tot;
        const uint32_t ex = shb_block_exscan_u32(v, &tot, sh);
        if (t < n_tile) tile_sum[t] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_vert_out = carry;
}
// pass 3: vertex id of every first occurrence
__global__ void __launch_bounds__(256) k_weld_number(const uint32_t* __restrict__ first, const uint32_t* __restrict__ slot_of, uint32_t n_corner,
                                                     const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ vid) {
    __shared__ uint32_t sh[9];
    const uint32_t base = blockIdx.x * SHB_WELD_TILE + threadIdx.x * 8u;
    uint32_t flags = 0, n = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8u; ++i) { const uint32_t c = base + i; if (c < n_corner && first[slot_of[c]] == c) { flags |= 1u << i; ++n; } }
    uint32_t tot;
    uint32_t id = tile_off[blockIdx.x] + shb_block_exscan_u32(n, &tot, sh);
#pragma unroll
    for (uint32_t i = 0; i < 8u; ++i) if ((flags >> i) & 1u) vid[base + i] = id++;
}
// faces and vertices in the layout of the slicing kernels
__global__ void __launch_bounds__(256) k_weld_emit(const unsigned char* __restrict__ stl, uint32_t n_corner, const uint32_t* __restrict__ first,
                                                   const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ vid,
                                                   double4* __restrict__ vert, double* __restrict__ vz, int* __restrict__ face /*[T][4]*/) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c >= n_corner) return;
    const uint32_t f = first[slot_of[c]], id = vid[f];
    face[4u * (c / 3u) + (c % 3u)] = (int)id;
    if (c % 3u == 0u) face[4u * (c / 3u) + 3u] = 0;
    if (f == c) {
        const float3 p = shb_stl_corner(stl, c);
        vert[id] = make_double4((double)p.x, (double)p.y, (double)p.z, 0.0);
        vz[id] = (double)p.z;
    }
}

// ------------------------------------------------------------------------------------------
// frame
// ------------------------------------------------------------------------------------------
struct ShbFrame {
    double mean[3];
    double axes[9];        // rows: x, y, z axis of the frame in source coordinates
    double shift[3];       // - mid point of the projected bounds
    double zb[2];          // z bounds in the frame before the end flip (mesh.py:88: kept as they were, like the reference)
    double zlen, band, zslice[2];
    double circ[2][3];     // cx, cy, r of the algebraic circle fit of the two end bands
    double resid[2];
    double nband[2];
    double flip;           // -1: x and z are negated (mesh.py:112-117), +1: not
    double transform[16];  // flip @ rot, row-major: source -> frame
};

__device__ __forceinline__ void shb_frame_point(const ShbFrame& F, const double4 v, double& x, double& y, double& z) {
    x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(F.axes[0], v.x), __dmul_rn(F.axes[1], v.y)), __dmul_rn(F.axes[2], v.z)), F.shift[0]);
    y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(F.axes[3], v.x), __dmul_rn(F.axes[4], v.y)), __dmul_rn(F.axes[5], v.z)), F.shift[1]);
    z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(F.axes[6], v.x), __dmul_rn(F.axes[7], v.y)), __dmul_rn(F.axes[8], v.z)), F.shift[2]);
}

// STAGE 1: sums of x, y, z.  2: centred second moments.  3: bounds of the projection on the axes (min x3, max x3).
// 4: normal equations of the circle fit of both end bands.  5: their residual sums.
template <int STAGE>
__global__ void __launch_bounds__(256) k_frame_acc(const double4* __restrict__ vert, uint32_t V, const ShbFrame* __restrict__ Fp,
                                                   double* __restrict__ part /*[grid][SHB_FRAME_K]*/) {
    constexpr int K = STAGE == 1 ? 3 : (STAGE == 2 ? 6 : (STAGE == 3 ? 6 : (STAGE == 4 ? 18 : 2)));
    constexpr bool MINMAX = STAGE == 3;
    __shared__ double sh[8][SHB_FRAME_K];
    __shared__ ShbFrame F;
    if (STAGE > 1) {
        for (uint32_t i = threadIdx.x; i < sizeof(ShbFrame) / 8; i += 256) reinterpret_cast<double*>(&F)[i] = reinterpret_cast<const double*>(Fp)[i];
        __syncthreads();
    }
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = MINMAX ? (k < 3 ? CUDART_INF : -CUDART_INF) : 0.0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < V; i += gridDim.x * 256u) {
        const double4 v = vert[i];
        if (STAGE == 1) { acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; }
        if (STAGE == 2) {
            const double dx = v.x - F.mean[0], dy = v.y - F.mean[1], dz = v.z - F.mean[2];
            acc[0] += dx * dx; acc[1] += dx * dy; acc[2] += dx * dz; acc[3] += dy * dy; acc[4] += dy * dz; acc[5] += dz * dz;
        }
        if (STAGE == 3) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double p = __dadd_rn(__dadd_rn(__dmul_rn(F.axes[3 * r], v.x), __dmul_rn(F.axes[3 * r + 1], v.y)), __dmul_rn(F.axes[3 * r + 2], v.z));
                acc[r] = fmin(acc[r], p); acc[3 + r] = fmax(acc[3 + r], p);
            }
        }
        if (STAGE == 4 || STAGE == 5) {
            double x, y, z;
            shb_frame_point(F, v, x, y, z);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (!(fabs(z - F.zslice[e]) < F.band)) continue;
                if (STAGE == 4) {
                    const double b = x * x + y * y, ax = 2.0 * x, ay = 2.0 * y;
                    double* a = acc + 9 * e;
                    a[0] += ax * ax; a[1] += ax * ay; a[2] += ax; a[3] += ay * ay; a[4] += ay; a[5] += 1.0;
                    a[6] += ax * b; a[7] += ay * b; a[8] += b;
                } else {
                    const double dx = x - F.circ[e][0], dy = y - F.circ[e][1];
                    const double t = sqrt(dx * dx + dy * dy) - F.circ[e][2];
                    acc[e] += t * t;
                }
            }
        }
    }
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double b = __shfl_xor_sync(0xffffffffu, a, o);
            a = MINMAX ? (k < 3 ? fmin(a, b) : fmax(a, b)) : a + b;
        }
        if (lane == 0) sh[w][k] = a;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        double a = sh[0][k];
        for (int i = 1; i < 8; ++i) a = MINMAX ? (k < 3 ? fmin(a, sh[i][k]) : fmax(a, sh[i][k])) : a + sh[i][k];
        part[(size_t)blockIdx.x * SHB_FRAME_K + k] = a;
    }
}

template <int STAGE>
__global__ void k_frame_step(ShbFrame* __restrict__ Fp, const double* __restrict__ part, uint32_t n_part, uint32_t V) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ShbFrame& F = *Fp;
    double s[SHB_FRAME_K];
    constexpr bool MINMAX = STAGE == 3;
    for (int k = 0; k < SHB_FRAME_K; ++k) s[k] = MINMAX ? (k < 3 ? CUDART_INF : -CUDART_INF) : 0.0;
    constexpr int K = STAGE == 1 ? 3 : (STAGE == 2 ? 6 : (STAGE == 3 ? 6 : (STAGE == 4 ? 18 : 2)));
    for (uint32_t b = 0; b < n_part; ++b)
        for (int k = 0; k < K; ++k) {
            const double p = part[(size_t)b * SHB_FRAME_K + k];
            s[k] = MINMAX ? (k < 3 ? fmin(s[k], p) : fmax(s[k], p)) : s[k] + p;
        }
    if (STAGE == 1) { for (int k = 0; k < 3; ++k) F.mean[k] = s[k] / (double)V; }
    if (STAGE == 2) {
        const double inv = 1.0 / (double)(V > 1 ? V - 1 : 1);
        double a[3][3] = {{s[0] * inv, s[1] * inv, s[2] * inv}, {s[1] * inv, s[3] * inv, s[4] * inv}, {s[2] * inv, s[4] * inv, s[5] * inv}};
        double v[3][3];
        shb_jacobi3(a, v);
        int ord[3] = {0, 1, 2};                                      // ascending eigenvalue: x, y, z (stable)
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j) if (a[ord[j]][ord[j]] > a[ord[j + 1]][ord[j + 1]]) { const int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
        for (int r = 0; r < 3; ++r) {
            double e[3] = {v[0][ord[r]], v[1][ord[r]], v[2][ord[r]]};
            int m = 0;                                               // sign: the component of largest magnitude is positive
            if (fabs(e[1]) > fabs(e[m])) m = 1;
            if (fabs(e[2]) > fabs(e[m])) m = 2;
            const double sg = e[m] < 0.0 ? -1.0 : 1.0;
            for (int k = 0; k < 3; ++k) F.axes[3 * r + k] = sg * e[k];
        }
        const double* A = F.axes;
        const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
        if (det < 0.0) for (int k = 0; k < 3; ++k) F.axes[k] = -F.axes[k];
    }
    if (STAGE == 3) {
        for (int r = 0; r < 3; ++r) F.shift[r] = -(0.5 * (s[r] + s[3 + r]));
        F.zb[0] = s[2] + F.shift[2]; F.zb[1] = s[5] + F.shift[2];
        F.zlen = fabs(F.zb[0]) + fabs(F.zb[1]);
        F.band = 0.01 * F.zlen;
        F.zslice[0] = 0.95 * F.zb[0]; F.zslice[1] = 0.95 * F.zb[1];
    }
    if (STAGE == 4) {
        for (int e = 0; e < 2; ++e) {
            const double* a = s + 9 * e;
            F.nband[e] = a[5];
            // normal equations of  [2x 2y 1] sol = x^2 + y^2 : symmetric 3x3, Gaussian elimination with partial pivoting
            double M[3][4] = {{a[0], a[1], a[2], a[6]}, {a[1], a[3], a[4], a[7]}, {a[2], a[4], a[5], a[8]}};
            bool ok = a[5] >= 8.0;
            for (int c = 0; c < 3 && ok; ++c) {
                int p = c;
                for (int r = c + 1; r < 3; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
                if (M[p][c] == 0.0) { ok = false; break; }
                if (p != c) for (int k = 0; k < 4; ++k) { const double t = M[c][k]; M[c][k] = M[p][k]; M[p][k] = t; }
                for (int r = c + 1; r < 3; ++r) { const double f = M[r][c] / M[c][c]; for (int k = c; k < 4; ++k) M[r][k] -= f * M[c][k]; }
            }
            double sol[3] = {0.0, 0.0, 0.0};
            if (ok) for (int r = 2; r >= 0; --r) { double t = M[r][3]; for (int k = r + 1; k < 3; ++k) t -= M[r][k] * sol[k]; sol[r] = t / M[r][r]; }
            F.circ[e][0] = sol[0]; F.circ[e][1] = sol[1];
            F.circ[e][2] = ok ? sqrt(sol[2] + sol[0] * sol[0] + sol[1] * sol[1]) : 0.0;
            if (!ok) F.nband[e] = 0.0;
        }
    }
    if (STAGE == 5) {
        double best = CUDART_INF, humeral_end = 0.0;
        for (int e = 0; e < 2; ++e) {
            F.resid[e] = F.nband[e] >= 8.0 ? s[e] / F.nband[e] : CUDART_INF;
            if (F.resid[e] < best) { best = F.resid[e]; humeral_end = F.zb[e]; }
        }
        F.flip = humeral_end < 0.0 ? -1.0 : 1.0;
        for (int r = 0; r < 3; ++r) {
            const double sg = (r == 1) ? 1.0 : F.flip;
            for (int k = 0; k < 3; ++k) F.transform[4 * r + k] = sg * F.axes[3 * r + k];
            F.transform[4 * r + 3] = sg * F.shift[r];
        }
        F.transform[12] = F.transform[13] = F.transform[14] = 0.0; F.transform[15] = 1.0;
    }
}

__global__ void __launch_bounds__(256) k_frame_apply(double4* __restrict__ vert, double* __restrict__ vz, uint32_t V, const ShbFrame* __restrict__ Fp) {
    __shared__ ShbFrame F;
    for (uint32_t i = threadIdx.x; i < sizeof(ShbFrame) / 8; i += 256) reinterpret_cast<double*>(&F)[i] = reinterpret_cast<const double*>(Fp)[i];
    __syncthreads();
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= V) return;
    double x, y, z;
    shb_frame_point(F, vert[i], x, y, z);
    if (F.flip < 0.0) { x = -x; z = -z; }
    vert[i] = make_double4(x, y, z, 0.0);
    vz[i] = z;
}

// (V,3) float64 + (T,3) int64 copies of a resident mesh for the host
__global__ void __launch_bounds__(256) k_mesh_unpack(const double4* __restrict__ vert, int64_t nv, const int4* __restrict__ face, int64_t nf,
                                                     double* __restrict__ v_out, int64_t* __restrict__ f_out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < nv) { const double4 v = vert[i]; v_out[3 * i] = v.x; v_out[3 * i + 1] = v.y; v_out[3 * i + 2] = v.z; }
    if (i < nf) { const int4 f = face[i]; f_out[3 * i] = f.x; f_out[3 * i + 1] = f.y; f_out[3 * i + 2] = f.z; }
}

extern "C" {

size_t shb_frame_state_bytes(void) { return sizeof(ShbFrame); }
size_t shb_frame_transform_offset(void) { return offsetof(ShbFrame, transform); }
size_t shb_frame_zb_offset(void) { return offsetof(ShbFrame, zb); }
size_t shb_frame_resid_offset(void) { return offsetof(ShbFrame, resid); }
size_t shb_frame_flip_offset(void) { return offsetof(ShbFrame, flip); }

// weld: slot_of / vid [n_corner], table / first [mask + 1] (table and first preset to 0xFF), tile_sum [ceil(n_corner / 2048)]
int shb_launch_weld_count(const unsigned char* stl, uint32_t n_corner, uint32_t* table, uint32_t* first, uint32_t mask, uint32_t* slot_of,
                          uint32_t* tile_sum, uint32_t* n_vert_out, cudaStream_t st) {
    if (n_corner == 0) return 0;
    const uint32_t tiles = (n_corner + SHB_WELD_TILE - 1) / SHB_WELD_TILE;
    k_weld_insert<<<(n_corner + 255) / 256, 256, 0, st>>>(stl, n_corner, table, first, mask, slot_of);
    k_weld_count<<<tiles, 256, 0, st>>>(first, slot_of, n_corner, tile_sum);
    k_weld_tiles<<<1, 256, 0, st>>>(tile_sum, tiles, n_vert_out);
    return 3;
}
int shb_launch_weld_emit(const unsigned char* stl, uint32_t n_corner, const uint32_t* first, const uint32_t* slot_of, const uint32_t* tile_off,
                         uint32_t* vid, double4* vert, double* vz, int4* face, cudaStream_t st) {
    if (n_corner == 0) return 0;
    const uint32_t tiles = (n_corner + SHB_WELD_TILE - 1) / SHB_WELD_TILE;
    k_weld_number<<<tiles, 256, 0, st>>>(first, slot_of, n_corner, tile_off, vid);
    k_weld_emit<<<(n_corner + 255) / 256, 256, 0, st>>>(stl, n_corner, first, slot_of, vid, vert, vz, reinterpret_cast<int*>(face));
    return 2;
}
// frame: part [148][SHB_FRAME_K] doubles, state = one ShbFrame
int shb_launch_frame(double4* vert, double* vz, uint32_t V, void* state, double* part, int n_sm, cudaStream_t st) {
    if (V == 0) return 0;
    ShbFrame* F = reinterpret_cast<ShbFrame*>(state);
    uint32_t grid = (V + 255) / 256;
    if (grid > (uint32_t)n_sm) grid = (uint32_t)n_sm;
    k_frame_acc<1><<<grid, 256, 0, st>>>(vert, V, F, part); k_frame_step<1><<<1, 32, 0, st>>>(F, part, grid, V);
    k_frame_acc<2><<<grid, 256, 0, st>>>(vert, V, F, part); k_frame_step<2><<<1, 32, 0, st>>>(F, part, grid, V);
    k_frame_acc<3><<<grid, 256, 0, st>>>(vert, V, F, part); k_frame_step<3><<<1, 32, 0, st>>>(F, part, grid, V);
    k_frame_acc<4><<<grid, 256, 0, st>>>(vert, V, F, part); k_frame_step<4><<<1, 32, 0, st>>>(F, part, grid, V);
    k_frame_acc<5><<<grid, 256, 0, st>>>(vert, V, F, part); k_frame_step<5><<<1, 32, 0, st>>>(F, part, grid, V);
    k_frame_apply<<<(V + 255) / 256, 256, 0, st>>>(vert, vz, V, F);
    return 11;
}
int shb_launch_mesh_unpack(const double4* vert, int64_t nv, const int4* face, int64_t nf, double* v_out, int64_t* f_out, cudaStream_t st) {
    const int64_t n = nv > nf ? nv : nf;
    if (n == 0) return 0;
    k_mesh_unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vert, nv, face, nf, v_out, f_out);
    return 1;
}

}  // extern "C"
