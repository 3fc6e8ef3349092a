// Feature extraction on the polar stacks while they are still in HBM (scope row f3): what the reference's landmark
// code does in per-row Python loops right after slice.py hands over `itr_centered_start` / `itr_start`:
//
//   k_groove_features   bicipital_groove.py:94-156   per row: zero-mean, negate, Savitzky-Golay(window 10, order 1, mode
//                       'interp'), roll to the arg-min, scipy.signal.find_peaks(height=-10, prominence=0.6, width=0.1)
//                       (local maxima with plateaus, prominences, widths at half prominence), the 7 most prominent peaks,
//                       9 features per peak (the StandardScaler over all peaks of a bone is two numbers per column: host)
//   k_groove_points     bicipital_groove.py:190-238  per row: searchsorted(theta, bg_theta), arg-min of the zero-mean
//                       radius within +-ivar samples (with the reference's wrap-around slice), back to x, y, z
//   k_neck_rows / k_neck_scale   anatomic_neck.py:38-58   per row: theta resampled evenly (np.linspace + np.interp),
//                       rolled to the sample nearest bg_theta; MinMaxScaler over the whole image -> float32
//
// Inputs are windows of profile arrays a result still holds on the device (float64); outputs are small (<= 7 peaks x
// 9 features per row, 3 numbers per row, one float32 image), so the 4.2 + 2.7 MB of polar stacks per bone never cross
// PCIe for these consumers.  One warp per row; find_peaks' walks are done one lane per peak.
#include "shb_common.cuh"
#include "../../include/shoulder_b200.h"
#include <math_constants.h>

#define SHB_F_TOP 7            // bicipital_groove.py:122  n = 7
#define SHB_F_NFEAT 9
#define SHB_F_MAXPK 256        // local maxima kept per row before the filters (a 512-sample row has at most 255)


__device__ __forceinline__ double shb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// which sweep an output row belongs to: last s with src[s].out_row0 <= row (binary search: a batch lists one sweep per bone)
__device__ __forceinline__ int shb_find_src(const ShbRowSrc* src, int n_src, uint32_t row) {
    int lo = 0, hi = n_src;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (src[mid].out_row0 <= row) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(128) k_groove_features(const ShbRowSrc* __restrict__ src, int n_src, uint32_t total_rows,
                                                         const double* __restrict__ zs /*[total_rows]*/, double* __restrict__ feat,
                                                         double* __restrict__ theta_out, int32_t* __restrict__ idx_out,
                                                         int32_t* __restrict__ cnt_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, FULLM = 0xffffffffu;
    const uint32_t row = blockIdx.x * 4u + w;
    if (row >= total_rows) return;
    const int si = shb_find_src(src, n_src, row);
    const ShbRowSrc S = src[si];
    const uint32_t N = S.N, lr = row - S.out_row0;
    const double* th = S.base + (size_t)lr * 2 * N;
    const double* rr = th + N;
    double* xs = reinterpret_cast<double*>(smem) + (size_t)w * (2 * (size_t)N + 64);       // negated zero-mean radius, later the rolled signal
    double* ys = xs + N;                                                                    // smoothed
    __shared__ uint16_t s_pk[4][SHB_F_MAXPK];
    __shared__ double s_prom[4][64], s_wid[4][64], s_hgt[4][64];
    __shared__ uint16_t s_kp[4][64];
    // ---- zero mean, negate (bicipital_groove.py:98-103,166-168)
    double sum = 0.0;
    for (uint32_t i = lane; i < N; i += 32) sum += rr[i];
    const double mean = shb_warp_sum(sum) / (double)N;
    for (uint32_t i = lane; i < N; i += 32) xs[i] = -1.0 * (rr[i] - mean);
    __syncwarp();
    // ---- scipy.signal.savgol_filter(x, 10, 1): interior = convolution with ten equal weights (a degree-1 fit evaluated at
    //      the window centre is the mean), the first / last five samples from a straight-line fit to the first / last ten
    //      (mode='interp')
    for (uint32_t i = lane; i < N; i += 32) {
        double y;
        if (N < 10) y = xs[i];
        else if (i >= 5 && i < N - 5) {                             // convolve1d with an even window: samples i - 4 .. i + 5
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 10; ++j) acc += 0.1 * xs[i - 4 + j];
            y = acc;
        } else {
            const uint32_t base = i < 5 ? 0u : N - 10u;
            double sx = 0.0, sxy = 0.0;
#pragma unroll
            for (int t = 0; t < 10; ++t) { const double v = xs[base + t]; sx += v; sxy += ((double)t - 4.5) * v; }
            const double slope = sxy / 82.5, icpt = sx / 10.0 - slope * 4.5;
            y = icpt + slope * (double)(i - base);
        }
        ys[i] = y;
    }
    __syncwarp();
    // ---- roll so that the arg-min comes first (:106-108)
    double bv = CUDART_INF; uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < N; i += 32) { const double v = ys[i]; if (v < bv) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULLM, bv, o); const uint32_t oi = __shfl_xor_sync(FULLM, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const uint32_t am = bi;
    __syncwarp();
    for (uint32_t k = lane; k < N; k += 32) { uint32_t q = k + am; if (q >= N) q -= N; xs[k] = ys[q]; }       // z = np.roll(y, -am)
    __syncwarp();
    const double* z = xs;
    // ---- scipy _local_maxima_1d: rising edge at i, plateau to i_ahead - 1, falling edge behind it -> midpoint
    uint32_t npk = 0;
    for (uint32_t i0 = 1; i0 + 1 < N; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t mid = 0; bool is = false;
        if (i + 1 < N && z[i - 1] < z[i]) {
            uint32_t a = i + 1;
            while (a < N - 1 && z[a] == z[i]) ++a;
            if (z[a] < z[i]) { is = true; mid = (i + a - 1) / 2; }
        }
        const uint32_t b = __ballot_sync(FULLM, is);
        if (is) { const uint32_t q = npk + __popc(b & ((1u << lane) - 1u)); if (q < SHB_F_MAXPK) s_pk[w][q] = (uint16_t)mid; }
        npk += __popc(b);
    }
    if (npk > SHB_F_MAXPK) npk = SHB_F_MAXPK;
    __syncwarp();
    // ---- per peak: height >= -10, prominence >= 0.6 (wlen = None), width at half prominence >= 0.1
    uint32_t nk = 0;
    for (uint32_t p0 = 0; p0 < npk; p0 += 32) {
        const uint32_t q = p0 + lane;
        bool keep = false; double prom = 0.0, wid = 0.0, hgt = 0.0; uint32_t p = 0;
        if (q < npk) {
            p = s_pk[w][q];
            const double zp = z[p];
            int i = (int)p, lb = (int)p, rb = (int)p;
            double lmin = zp, rmin = zp;
            while (0 <= i && z[i] <= zp) { if (z[i] < lmin) { lmin = z[i]; lb = i; } --i; }
            i = (int)p;
            while (i <= (int)N - 1 && z[i] <= zp) { if (z[i] < rmin) { rmin = z[i]; rb = i; } ++i; }
            prom = zp - fmax(lmin, rmin);
            hgt = zp - prom * 0.5;
            i = (int)p;
            while (lb < i && hgt < z[i]) --i;
            double lip = (double)i;
            if (z[i] < hgt) lip += (hgt - z[i]) / (z[i + 1] - z[i]);
            i = (int)p;
            while (i < rb && hgt < z[i]) ++i;
            double rip = (double)i;
            if (z[i] < hgt) rip -= (hgt - z[i]) / (z[i - 1] - z[i]);
            wid = rip - lip;
            keep = zp >= -10.0 && prom >= 0.6 && wid >= 0.1;
        }
        const uint32_t b = __ballot_sync(FULLM, keep);
        if (keep) {
            const uint32_t o = nk + __popc(b & ((1u << lane) - 1u));
            if (o < 64) { s_kp[w][o] = (uint16_t)p; s_prom[w][o] = prom; s_wid[w][o] = wid; s_hgt[w][o] = hgt; }
        }
        nk += __popc(b);
    }
    if (nk > 64) nk = 64;
    __syncwarp();
    // ---- the SHB_F_TOP most prominent (np.argpartition: any order; here ascending position), their features
    uint32_t nsel = 0;
    uint32_t selq[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    for (uint32_t t = 0; t < 2; ++t) {
        const uint32_t q = 32u * t + lane;
        bool sel = false;
        if (q < nk) {
            uint32_t better = 0;
            for (uint32_t o = 0; o < nk; ++o) better += s_prom[w][o] > s_prom[w][q] || (s_prom[w][o] == s_prom[w][q] && o > q);
            sel = nk <= SHB_F_TOP || better < SHB_F_TOP;
        }
        const uint32_t b = __ballot_sync(FULLM, sel);
        if (sel) selq[t] = nsel + __popc(b & ((1u << lane) - 1u));
        nsel += __popc(b);
    }
    __shared__ uint16_t s_sel[4][SHB_F_TOP + 1];
    for (uint32_t t = 0; t < 2; ++t) if (selq[t] < SHB_F_TOP) s_sel[w][selq[t]] = (uint16_t)(32u * t + lane);
    __syncwarp();
    if (nsel > SHB_F_TOP) nsel = SHB_F_TOP;
    if (lane == 0) cnt_out[row] = (int32_t)nsel;
    if (lane < nsel) {
        const uint32_t q = s_sel[w][lane];
        uint32_t io = s_kp[w][q] + am; if (io >= N) io -= N;                 // (peaks - rmin) % interp_num
        const double t = th[io], r = rr[io];
        // closest_angles / peak_nearest / peak_next_nearest (:31-65): |atan2(sin(d), cos(d))| to every peak of the row, those that
        // round to 0.00 dropped, sorted
        double a0 = CUDART_INF, a1 = CUDART_INF;
        for (uint32_t o = 0; o < nsel; ++o) {
            uint32_t jo = s_kp[w][s_sel[w][o]] + am; if (jo >= N) jo -= N;
            const double dlt = t - th[jo];
            const double a = fabs(atan2(sin(dlt), cos(dlt)));
            if (rint(a * 100.0) == 0.0) continue;
            if (a < a0) { a1 = a0; a0 = a; } else if (a < a1) a1 = a;
        }
        const double near0 = nsel == 1 || a0 == CUDART_INF ? 0.0 : a0;
        const double near1 = nsel <= 2 || a1 == CUDART_INF ? 0.0 : a1;
        const double zrow = zs[row];
        const double px = r * cos(t) - S.cu[0] * zrow, py = r * sin(t) - S.cu[1] * zrow;        // canal_dist (:67-82)
        double* f = feat + ((size_t)row * SHB_F_TOP + lane) * SHB_F_NFEAT;
        f[0] = r; f[1] = near0; f[2] = near1; f[3] = zrow * S.z_scale + S.z_min; f[4] = s_prom[w][q]; f[5] = s_wid[w][q];
        f[6] = s_hgt[w][q]; f[7] = sqrt(px * px + py * py); f[8] = (double)nsel / (double)SHB_F_TOP;
        theta_out[(size_t)row * SHB_F_TOP + lane] = t;
        idx_out[(size_t)row * SHB_F_TOP + lane] = (int32_t)io;
    }
}

__global__ void __launch_bounds__(128) k_groove_points(const ShbRowSrc* __restrict__ src, int n_src, uint32_t total_rows,
                                                       const double* __restrict__ zs, const double* __restrict__ bg_theta /*[n_src]*/,
                                                       int ivar, const double* __restrict__ centroid /*[G][2]*/,
                                                       double* __restrict__ pts /*[rows][3]*/, double* __restrict__ local_theta) {
    const uint32_t row = blockIdx.x * 128u + threadIdx.x;
    if (row >= total_rows) return;
    const int si = shb_find_src(src, n_src, row);
    const ShbRowSrc S = src[si];
    const int N = (int)S.N;
    const uint32_t lr = row - S.out_row0;
    const double* th = S.base + (size_t)lr * 2 * N;
    const double* rr = th + N;
    const double key = bg_theta[si];
    int lo = 0, hi = N;                                            // np.searchsorted(theta, bg_theta, side='left')
    while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (th[mid] < key) lo = mid + 1; else hi = mid; }
    int est = lo == N ? N - 1 : lo;
    // candidates in the order of the reference's slice (concatenated around the end when ivar > est, :207-221)
    double best = CUDART_INF; int loc = 0;
    auto consider = [&](int orig, int j) { const double v = rr[orig]; if (v < best) { best = v; loc = j + est - ivar; } };
    if (ivar > est) {
        int j = 0;
        for (int o = N + (est - ivar); o < N; ++o, ++j) consider(o, j);
        for (int o = 0; o < est + ivar && o < N; ++o, ++j) consider(o, j);
    } else {
        int j = 0;
        for (int o = est - ivar; o < est + ivar && o < N; ++o, ++j) consider(o, j);
    }
    const int io = loc < 0 ? loc + N : loc;                        // python's negative index
    const double t = th[io], r = rr[io];
    const uint32_t plane = S.plane0 + lr;
    local_theta[row] = t;
    pts[3 * (size_t)row + 0] = r * cos(t) + centroid[2 * (size_t)plane];
    pts[3 * (size_t)row + 1] = r * sin(t) + centroid[2 * (size_t)plane + 1];
    pts[3 * (size_t)row + 2] = zs[row];
}

__device__ __forceinline__ unsigned long long shb_sortable(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}

// one warp per row: r on even theta, rolled to the groove; the image's min / max by two atomics per row
__global__ void __launch_bounds__(128) k_neck_rows(const ShbRowSrc* __restrict__ src, int n_src, uint32_t total_rows,
                                                   const double* __restrict__ bg_theta, double* __restrict__ vals /*[rows][N]*/,
                                                   double* __restrict__ shft /*[rows][2][N] or null*/,
                                                   unsigned long long* __restrict__ mm /*[n_src][2] sortable min, max*/) {
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5, FULLM = 0xffffffffu;
    const uint32_t row = blockIdx.x * 4u + w;
    if (row >= total_rows) return;
    const int si = shb_find_src(src, n_src, row);
    const ShbRowSrc S = src[si];
    const uint32_t N = S.N, lr = row - S.out_row0, M = N - 1;      // interp over the first N - 1 samples (anatomic_neck.py:43-44)
    const double* th = S.base + (size_t)lr * 2 * N;
    const double* rr = th + N;
    const double start = th[0], stop = th[N - 2];
    const double step = (stop - start) / (double)(N - 1);           // np.linspace(start, stop, N)
    const double key = bg_theta[si];
    double bd = CUDART_INF; uint32_t bk = 0xFFFFFFFFu;
    double* orow = vals + (size_t)row * N;
    for (uint32_t k = lane; k < N; k += 32) {
        const double x = k == N - 1 ? stop : (double)k * step + start;
        // np.interp(x, xp = th[:M], fp = rr[:M])
        double y;
        if (x >= th[M - 1]) y = rr[M - 1];
        else if (x < th[0]) y = rr[0];
        else {
            uint32_t lo = 0, hi = M;                                // last j with xp[j] <= x
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (th[mid] <= x) lo = mid; else hi = mid; }
            const double slope = (rr[lo + 1] - rr[lo]) / (th[lo + 1] - th[lo]);
            y = slope * (x - th[lo]) + rr[lo];
        }
        orow[k] = y;                                                // unrolled position; rolled below
        const double dk = fabs(x - key);
        if (dk < bd) { bd = dk; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULLM, bd, o); const uint32_t ok = __shfl_xor_sync(FULLM, bk, o);
        if (ov < bd || (ov == bd && ok < bk)) { bd = ov; bk = ok; }
    }
    __syncwarp();
    // roll in place through registers: N / 32 values per lane
    double mn = CUDART_INF, mx = -CUDART_INF;
    double buf[32];                                                 // N <= 1024
    const uint32_t per = (N + 31u) / 32u;
    for (uint32_t q = 0; q < per && q < 32; ++q) { const uint32_t j = q * 32u + lane; uint32_t s = j + bk; if (s >= N) s -= N; buf[q] = j < N ? orow[s] : 0.0; }
    __syncwarp();
    for (uint32_t q = 0; q < per && q < 32; ++q) {
        const uint32_t j = q * 32u + lane;
        if (j < N) {
            orow[j] = buf[q]; mn = fmin(mn, buf[q]); mx = fmax(mx, buf[q]);
            if (shft) {
                uint32_t s = j + bk; if (s >= N) s -= N;
                shft[(size_t)row * 2 * N + j] = s == N - 1 ? stop : (double)s * step + start;
                shft[(size_t)row * 2 * N + N + j] = buf[q];
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(FULLM, mn, o)); mx = fmax(mx, __shfl_xor_sync(FULLM, mx, o)); }
    if (lane == 0) { atomicMin(mm + 2 * si, shb_sortable(mn)); atomicMax(mm + 2 * si + 1, shb_sortable(mx)); }
}

__device__ __forceinline__ double shb_unsortable(unsigned long long k) {
    const unsigned long long u = (k & 0x8000000000000000ULL) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
    return __longlong_as_double((long long)u);
}
// sklearn MinMaxScaler(feature_range=(0, 1)): X * scale_ + min_ with scale_ = 1 / (max - min), min_ = -min * scale_; float32 out
__global__ void __launch_bounds__(256) k_neck_scale(const ShbRowSrc* __restrict__ src, int n_src, uint32_t total_rows, uint32_t N,
                                                    const double* __restrict__ vals, const unsigned long long* __restrict__ mm,
                                                    float* __restrict__ image, double* __restrict__ mm_out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)total_rows * N) return;
    const uint32_t row = (uint32_t)(i / N);
    const int si = shb_find_src(src, n_src, row);
    const double lo = shb_unsortable(mm[2 * si]), hi = shb_unsortable(mm[2 * si + 1]);
    double range = hi - lo;
    if (range == 0.0) range = 1.0;                                   // sklearn _handle_zeros_in_scale
    const double scale = 1.0 / range, mn = 0.0 - lo * scale;
    image[i] = (float)(vals[i] * scale + mn);
    if (i == (size_t)src[si].out_row0 * N) { mm_out[2 * si] = lo; mm_out[2 * si + 1] = hi; }
}

// random forest (onnx TreeEnsembleClassifier, BRANCH_LEQ): one thread per (sample, tree), class-1 score summed per sample
__global__ void __launch_bounds__(256) k_forest(const float* __restrict__ X, uint32_t n, uint32_t n_feat, uint32_t n_trees,
                                                const uint32_t* __restrict__ root, const int32_t* __restrict__ feature,
                                                const float* __restrict__ value, const uint32_t* __restrict__ tchild,
                                                const uint32_t* __restrict__ fchild, const float* __restrict__ weight,
                                                float* __restrict__ score /*[n], zeroed*/) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)n * n_trees) return;
    const uint32_t s = (uint32_t)(i / n_trees), t = (uint32_t)(i % n_trees);
    uint32_t cur = root[t];
    while (feature[cur] >= 0) cur = X[(size_t)s * n_feat + feature[cur]] <= value[cur] ? tchild[cur] : fchild[cur];
    atomicAdd(score + s, weight[cur]);
}

// ------------------------------------------------------------------------------------------
// ray - mesh queries (scope row f4; anatomic_neck.py:184-191,217-224: mesh.ray.intersects_location, four rays per bone).
// trimesh's numpy backend (ray_triangle_id): the ray meets the triangle's plane (face normal = unit cross product of the
// first two edges; skipped when |direction . normal| <= 1e-5), the point is inside when its barycentric coordinates
// (Cramer's rule, points_to_barycentric) lie in (-tol.zero, 1 + tol.zero), and only points at distance > -1e-6 along the
// ray count.  trimesh prefilters candidates with an R-tree; the exact test is the same, so testing every triangle gives
// the same set of hits.  One thread per (ray, triangle); hits are appended and sorted by (ray, triangle) on the host.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 shb_ldv4(const double4* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double shb_dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));       // (a * b).sum(axis=1)
}
__global__ void __launch_bounds__(256) k_ray_cast(const double4* __restrict__ vert, const int4* __restrict__ face, int64_t n_face,
                                                  const double* __restrict__ org, const double* __restrict__ dir, int n_ray,
                                                  int max_hits, int32_t* __restrict__ hit_ray, int32_t* __restrict__ hit_tri,
                                                  double* __restrict__ hit_loc, double* __restrict__ hit_dist, uint32_t* __restrict__ n_hits) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int ray = blockIdx.y;
    if (t >= n_face || ray >= n_ray) return;
    const int4 f = __ldg(face + t);
    const double4 A = shb_ldv4(vert + f.x), B = shb_ldv4(vert + f.y), C = shb_ldv4(vert + f.z);
    const double e0x = __dsub_rn(B.x, A.x), e0y = __dsub_rn(B.y, A.y), e0z = __dsub_rn(B.z, A.z);
    const double e1x = __dsub_rn(C.x, A.x), e1y = __dsub_rn(C.y, A.y), e1z = __dsub_rn(C.z, A.z);
    // face normal: unitize(cross(e0, e1)); degenerate faces have none
    double nx = __dsub_rn(__dmul_rn(e0y, e1z), __dmul_rn(e0z, e1y)), ny = __dsub_rn(__dmul_rn(e0z, e1x), __dmul_rn(e0x, e1z)),
           nz = __dsub_rn(__dmul_rn(e0x, e1y), __dmul_rn(e0y, e1x));
    const double nn = __dsqrt_rn(shb_dot3(nx, ny, nz, nx, ny, nz));
    if (!(nn > 1e-13)) return;
    nx = __ddiv_rn(nx, nn); ny = __ddiv_rn(ny, nn); nz = __ddiv_rn(nz, nn);
    const double ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const double dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    const double p_ori = shb_dot3(__dsub_rn(A.x, ox), __dsub_rn(A.y, oy), __dsub_rn(A.z, oz), nx, ny, nz);
    const double p_dir = shb_dot3(dx, dy, dz, nx, ny, nz);
    if (!(fabs(p_dir) > 1e-5)) return;
    const double dist = __ddiv_rn(p_ori, p_dir);
    const double px = __dadd_rn(__dmul_rn(dx, dist), ox), py = __dadd_rn(__dmul_rn(dy, dist), oy), pz = __dadd_rn(__dmul_rn(dz, dist), oz);
    const double wx = __dsub_rn(px, A.x), wy = __dsub_rn(py, A.y), wz = __dsub_rn(pz, A.z);
    const double d00 = shb_dot3(e0x, e0y, e0z, e0x, e0y, e0z), d01 = shb_dot3(e0x, e0y, e0z, e1x, e1y, e1z), d02 = shb_dot3(e0x, e0y, e0z, wx, wy, wz),
                 d11 = shb_dot3(e1x, e1y, e1z, e1x, e1y, e1z), d12 = shb_dot3(e1x, e1y, e1z, wx, wy, wz);
    const double inv = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(d00, d11), __dmul_rn(d01, d01)));
    const double b2 = __dmul_rn(__dsub_rn(__dmul_rn(d00, d12), __dmul_rn(d01, d02)), inv);
    const double b1 = __dmul_rn(__dsub_rn(__dmul_rn(d11, d02), __dmul_rn(d01, d12)), inv);
    const double b0 = __dsub_rn(__dsub_rn(1.0, b1), b2);
    const double tz = 1e-13;                                        // trimesh tol.zero
    if (!(b0 > -tz && b1 > -tz && b2 > -tz && b0 < 1.0 + tz && b1 < 1.0 + tz && b2 < 1.0 + tz)) return;
    const double fwd = shb_dot3(__dsub_rn(px, ox), __dsub_rn(py, oy), __dsub_rn(pz, oz), dx, dy, dz);
    if (!(fwd > -1e-6)) return;
    const uint32_t k = atomicAdd(n_hits, 1u);
    if ((int)k < max_hits) {
        hit_ray[k] = ray; hit_tri[k] = (int32_t)t;
        hit_loc[3 * (size_t)k] = px; hit_loc[3 * (size_t)k + 1] = py; hit_loc[3 * (size_t)k + 2] = pz;
        hit_dist[k] = fwd;
    }
}

// bicipital_groove.py:184-188: the groove angle of a bone = arg-max over np.linspace(-pi, pi, 1024) of the linear-kernel
// density (bandwidth 1: sum of max(0, 1 - |t - theta_i|)) of the peaks the forest accepted (probability > threshold).
// One CTA per bone, one thread per grid angle, the accepted peaks staged in shared memory in peak order; arg-max with
// np.argmax's first-occurrence rule.  (sklearn's KernelDensity returns the log of the same sum over a constant: same arg-max.)
__global__ void __launch_bounds__(1024) k_groove_theta(const long long* __restrict__ off, const double* __restrict__ peak_theta,
                                                       const float* __restrict__ proba1, float threshold, double* __restrict__ bg,
                                                       double* __restrict__ dens_max) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* pts = reinterpret_cast<double*>(smem);
    __shared__ uint32_t n_acc;
    __shared__ double red_v[32];
    __shared__ uint32_t red_i[32];
    const uint32_t b = blockIdx.x, k = threadIdx.x, lane = k & 31u, w = k >> 5;
    const long long p0 = off[b], p1 = off[b + 1];
    if (k == 0) {                                                   // compaction in peak order (a few thousand peaks)
        uint32_t m = 0;
        for (long long i = p0; i < p1; ++i) if (proba1[i] > threshold) pts[m++] = peak_theta[i];
        n_acc = m;
    }
    __syncthreads();
    const double pi = 3.141592653589793;
    const double step = __ddiv_rn(__dsub_rn(pi, -pi), 1023.0);
    const double t = k == 1023u ? pi : __dadd_rn(__dmul_rn((double)k, step), -pi);       // np.linspace
    double dens = 0.0;
    const uint32_t m = n_acc;
    for (uint32_t i = 0; i < m; ++i) { const double v = __dsub_rn(1.0, fabs(__dsub_rn(t, pts[i]))); dens = __dadd_rn(dens, v > 0.0 ? v : 0.0); }
    double bv = dens; uint32_t bi = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[w] = bv; red_i[w] = bi; }
    __syncthreads();
    if (k == 0) {
        for (int q = 1; q < 32; ++q) if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) { bv = red_v[q]; bi = red_i[q]; }
        bg[b] = bi == 1023u ? pi : __dadd_rn(__dmul_rn((double)bi, step), -pi);
        if (dens_max) dens_max[b] = bv;
    }
}


// ------------------------------------------------------------------------------------------
// The landmark front end in ONE enqueue (shb_landmark_front): what the step-by-step calls above leave to the host between
// them, on the device, so that a batch of bones goes from resident polar stacks to landmark-model inputs with one
// host wait: canal axis (canal.py:40-85), StandardScaler (bicipital_groove.py:171-172), forest and density arg-max over
// the slot layout (rows, 7) the feature kernel writes — no compaction, the peaks are visited in row-major slot order, which
// IS the order of the reference's per-row appends.
// ------------------------------------------------------------------------------------------

// canal.py:40-85 + scikit-spatial Line.best_fit: centroid = points.mean(axis=0) (numpy adds the rows in order), direction =
// first right singular vector of the centred points = eigenvector of the largest eigenvalue of their 3x3 scatter matrix
// (cyclic Jacobi), pointed proximally (direction[-1] >= 0), axis = centroid +- direction * half.  One warp per bone.
__global__ void __launch_bounds__(32) k_canal_axes(const ShbCanalJob* __restrict__ jobs, int n_jobs, const double* __restrict__ centroid /*[plane][2]*/,
                                                   const double* __restrict__ z, ShbRowSrc* __restrict__ src /*cu[] written*/,
                                                   double* __restrict__ axes /*[n][2][3]*/) {
    const int b = blockIdx.x;
    if (b >= n_jobs) return;
    const ShbCanalJob J = jobs[b];
    const uint32_t lane = threadIdx.x, n = J.rows;
    const double* c = centroid + 2 * (size_t)J.plane0;
    const double* zz = z + J.z_off;
    double mean = 0.0;
    if (lane < 3) {
        double acc = 0.0;
        for (uint32_t i = 0; i < n; ++i) acc = __dadd_rn(acc, lane < 2 ? c[2 * (size_t)i + lane] : zz[i]);
        mean = __ddiv_rn(acc, (double)n);
    }
    const double mx = __shfl_sync(0xffffffffu, mean, 0), my = __shfl_sync(0xffffffffu, mean, 1), mz = __shfl_sync(0xffffffffu, mean, 2);
    double sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const double dx = c[2 * (size_t)i] - mx, dy = c[2 * (size_t)i + 1] - my, dz = zz[i] - mz;
        sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
    sxx = shb_warp_sum(sxx); sxy = shb_warp_sum(sxy); sxz = shb_warp_sum(sxz); syy = shb_warp_sum(syy); syz = shb_warp_sum(syz); szz = shb_warp_sum(szz);
    if (lane != 0) return;
    double a[3][3] = {{sxx, sxy, sxz}, {sxy, syy, syz}, {sxz, syz, szz}}, v[3][3];
    shb_jacobi3(a, v);
    int m = 0;
    if (a[1][1] > a[m][m]) m = 1;
    if (a[2][2] > a[m][m]) m = 2;
    double d[3] = {v[0][m], v[1][m], v[2][m]};
    const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double sg = d[2] < 0.0 ? -1.0 : 1.0;
    for (int k = 0; k < 3; ++k) d[k] = sg * d[k] / nrm;
    const double mid[3] = {mx, my, mz};
    double ax[6];
    for (int k = 0; k < 3; ++k) { ax[k] = mid[k] + d[k] * J.half_len; ax[3 + k] = mid[k] - d[k] * J.half_len; }
    for (int k = 0; k < 6; ++k) axes[6 * (size_t)b + k] = ax[k];
    // utils.unit_vector(axis[0], axis[1]) as the host path computes it for the feature kernel (row_sources)
    const double vx = ax[0] - ax[3], vy = ax[1] - ax[4], vz = ax[2] - ax[5];
    const double vn = sqrt(vx * vx + vy * vy + vz * vz);
    src[b].cu[0] = vx / vn; src[b].cu[1] = vy / vn; src[b].cu[2] = vz / vn;
}

// sklearn StandardScaler over all peaks of one bone: mean_ = X.mean(axis=0), scale_ = X.std(axis=0) with zeros -> 1, X' = (X - mean_)
// / scale_, cast to float32 for the forest.  numpy reduces a C-ordered (n, 9) array over axis 0 by adding the rows in order,
// so one thread per column adding in row-major slot order reproduces both statistics bit for bit.  One CTA per bone; the bone's
// feature rows are staged in shared memory when they fit (330 rows x 504 B).
template <bool STAGED>
__global__ void __launch_bounds__(256) k_groove_scale(const ShbRowSrc* __restrict__ src, const double* __restrict__ feat,
                                                      const int32_t* __restrict__ cnt, float* __restrict__ X /*[rows][7][9]*/,
                                                      double* __restrict__ stats /*[n_src][2][9] or null*/) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double s_mean[SHB_F_NFEAT], s_std[SHB_F_NFEAT];
    const ShbRowSrc S = src[blockIdx.x];
    const uint32_t rows = S.rows, tid = threadIdx.x;
    constexpr uint32_t RW = SHB_F_TOP * SHB_F_NFEAT;
    const double* g = feat + (size_t)S.out_row0 * RW;
    const int32_t* c = cnt + S.out_row0;
    const double* f = g;
    if (STAGED) {
        double* sf = reinterpret_cast<double*>(smem);
        for (uint32_t i = tid; i < rows * RW; i += 256) sf[i] = g[i];
        f = sf;
    }
    __syncthreads();
    if (tid < SHB_F_NFEAT) {
        double acc = 0.0; uint32_t n = 0;
        for (uint32_t r = 0; r < rows; ++r) {
            const uint32_t k = (uint32_t)c[r];
            for (uint32_t q = 0; q < k; ++q) acc = __dadd_rn(acc, f[r * RW + q * SHB_F_NFEAT + tid]);
            n += k;
        }
        const double mean = __ddiv_rn(acc, (double)n);
        acc = 0.0;
        for (uint32_t r = 0; r < rows; ++r) {
            const uint32_t k = (uint32_t)c[r];
            for (uint32_t q = 0; q < k; ++q) { const double dlt = __dsub_rn(f[r * RW + q * SHB_F_NFEAT + tid], mean); acc = __dadd_rn(acc, __dmul_rn(dlt, dlt)); }
        }
        double sd = __dsqrt_rn(__ddiv_rn(acc, (double)n));
        if (sd == 0.0) sd = 1.0;                                      // sklearn _handle_zeros_in_scale
        s_mean[tid] = mean; s_std[tid] = sd;
        if (stats) { stats[(size_t)blockIdx.x * 2 * SHB_F_NFEAT + tid] = mean; stats[(size_t)blockIdx.x * 2 * SHB_F_NFEAT + SHB_F_NFEAT + tid] = sd; }
    }
    __syncthreads();
    for (uint32_t i = tid; i < rows * RW; i += 256) {
        const uint32_t r = i / RW, q = (i % RW) / SHB_F_NFEAT, j = i % SHB_F_NFEAT;
        X[(size_t)S.out_row0 * RW + i] = q < (uint32_t)c[r] ? (float)__ddiv_rn(__dsub_rn(f[i], s_mean[j]), s_std[j]) : 0.f;
    }
}

// k_forest over the slot layout: sample = row * 7 + slot, slots at or beyond the row's peak count are skipped
__global__ void __launch_bounds__(256) k_forest_slots(const float* __restrict__ X, const int32_t* __restrict__ cnt, uint32_t n_rows, uint32_t n_feat,
                                                      uint32_t n_trees, const uint32_t* __restrict__ root, const int32_t* __restrict__ feature,
                                                      const float* __restrict__ value, const uint32_t* __restrict__ tchild,
                                                      const uint32_t* __restrict__ fchild, const float* __restrict__ weight,
                                                      float* __restrict__ score /*[n_rows * 7], zeroed*/) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)n_rows * SHB_F_TOP * n_trees) return;
    const uint32_t s = (uint32_t)(i / n_trees), t = (uint32_t)(i % n_trees);
    if ((int32_t)(s % SHB_F_TOP) >= cnt[s / SHB_F_TOP]) return;
    uint32_t cur = root[t];
    while (feature[cur] >= 0) cur = X[(size_t)s * n_feat + feature[cur]] <= value[cur] ? tchild[cur] : fchild[cur];
    atomicAdd(score + s, weight[cur]);
}

// k_groove_theta over the slot layout: one CTA per bone; warp 0 compacts the accepted peaks in row-major slot order
__global__ void __launch_bounds__(1024) k_groove_theta_slots(const ShbRowSrc* __restrict__ src, const double* __restrict__ theta /*[rows][7]*/,
                                                             const int32_t* __restrict__ cnt, const float* __restrict__ score, float threshold,
                                                             double* __restrict__ bg, double* __restrict__ dens_max) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* pts = reinterpret_cast<double*>(smem);
    __shared__ uint32_t n_acc;
    __shared__ double red_v[32];
    __shared__ uint32_t red_i[32];
    const uint32_t b = blockIdx.x, k = threadIdx.x, lane = k & 31u, w = k >> 5;
    const ShbRowSrc S = src[b];
    if (w == 0) {
        uint32_t m = 0;
        const uint32_t tot = S.rows * SHB_F_TOP;
        for (uint32_t e0 = 0; e0 < tot; e0 += 32) {
            const uint32_t e = e0 + lane;
            bool ok = false; double t = 0.0;
            if (e < tot) {
                const uint32_t row = S.out_row0 + e / SHB_F_TOP, slot = e % SHB_F_TOP;
                if ((int32_t)slot < cnt[row] && score[(size_t)row * SHB_F_TOP + slot] > threshold) { ok = true; t = theta[(size_t)row * SHB_F_TOP + slot]; }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, ok);
            if (ok) pts[m + __popc(bal & ((1u << lane) - 1u))] = t;
            m += __popc(bal);
        }
        if (lane == 0) n_acc = m;
    }
    __syncthreads();
    const double pi = 3.141592653589793;
    const double step = __ddiv_rn(__dsub_rn(pi, -pi), 1023.0);
    const double t = k == 1023u ? pi : __dadd_rn(__dmul_rn((double)k, step), -pi);       // np.linspace
    double dens = 0.0;
    const uint32_t m = n_acc;
    for (uint32_t i = 0; i < m; ++i) { const double v = __dsub_rn(1.0, fabs(__dsub_rn(t, pts[i]))); dens = __dadd_rn(dens, v > 0.0 ? v : 0.0); }
    double bv = dens; uint32_t bi = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[w] = bv; red_i[w] = bi; }
    __syncthreads();
    if (k == 0) {
        for (int q = 1; q < 32; ++q) if (red_v[q] > bv || (red_v[q] == bv && red_i[q] < bi)) { bv = red_v[q]; bi = red_i[q]; }
        bg[b] = bi == 1023u ? pi : __dadd_rn(__dmul_rn((double)bi, step), -pi);
        if (dens_max) dens_max[b] = bv;
    }
}

extern "C" {
int shb_launch_ray_cast(const double4* vert, const int4* face, int64_t n_face, const double* org, const double* dir, int n_ray, int max_hits,
                        int32_t* hit_ray, int32_t* hit_tri, double* hit_loc, double* hit_dist, uint32_t* n_hits, cudaStream_t st) {
    if (!n_face || !n_ray) return 0;
    dim3 grid((unsigned)((n_face + 255) / 256), (unsigned)n_ray);
    k_ray_cast<<<grid, 256, 0, st>>>(vert, face, n_face, org, dir, n_ray, max_hits, hit_ray, hit_tri, hit_loc, hit_dist, n_hits);
    return 1;
}
size_t shb_groove_smem_bytes(uint32_t N) { return 4 * (2 * (size_t)N + 64) * sizeof(double); }
int shb_launch_groove_features(const ShbRowSrc* src, int n_src, uint32_t rows, uint32_t maxN, const double* zs, double* feat, double* theta,
                               int32_t* idx, int32_t* cnt, cudaStream_t st) {
    if (!rows) return 0;
    const size_t smem = shb_groove_smem_bytes(maxN);
    cudaFuncSetAttribute(k_groove_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_groove_features<<<(rows + 3) / 4, 128, smem, st>>>(src, n_src, rows, zs, feat, theta, idx, cnt);
    return 1;
}
int shb_launch_groove_points(const ShbRowSrc* src, int n_src, uint32_t rows, const double* zs, const double* bg, int ivar,
                             const double* centroid, double* pts, double* local_theta, cudaStream_t st) {
    if (!rows) return 0;
    k_groove_points<<<(rows + 127) / 128, 128, 0, st>>>(src, n_src, rows, zs, bg, ivar, centroid, pts, local_theta);
    return 1;
}
int shb_launch_neck_image(const ShbRowSrc* src, int n_src, uint32_t rows, uint32_t N, const double* bg, double* vals, double* shft,
                          unsigned long long* mm, float* image, double* mm_out, cudaStream_t st) {
    if (!rows) return 0;
    k_neck_rows<<<(rows + 3) / 4, 128, 0, st>>>(src, n_src, rows, bg, vals, shft, mm);
    const size_t tot = (size_t)rows * N;
    k_neck_scale<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, n_src, rows, N, vals, mm, image, mm_out);
    return 2;
}
int shb_launch_forest(const float* X, uint32_t n, uint32_t n_feat, uint32_t n_trees, const uint32_t* root, const int32_t* feature,
                      const float* value, const uint32_t* tchild, const uint32_t* fchild, const float* weight, float* score, cudaStream_t st) {
    if (!n) return 0;
    const size_t tot = (size_t)n * n_trees;
    k_forest<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(X, n, n_feat, n_trees, root, feature, value, tchild, fchild, weight, score);
    return 1;
}
int shb_launch_groove_theta(const long long* off, int n_set, uint32_t max_peaks, const double* peak_theta, const float* proba1, float threshold,
                            double* bg, double* dens_max, cudaStream_t st) {
    if (!n_set) return 0;
    const size_t smem = 8 * (size_t)(max_peaks ? max_peaks : 1);
    cudaFuncSetAttribute(k_groove_theta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_groove_theta<<<n_set, 1024, smem, st>>>(off, peak_theta, proba1, threshold, bg, dens_max);
    return 1;
}
int shb_launch_canal_axes(const ShbCanalJob* jobs, int n_jobs, const double* centroid, const double* z, ShbRowSrc* src, double* axes, cudaStream_t st) {
    if (n_jobs <= 0) return 0;
    k_canal_axes<<<n_jobs, 32, 0, st>>>(jobs, n_jobs, centroid, z, src, axes);
    return 1;
}
int shb_launch_groove_scale(const ShbRowSrc* src, int n_src, uint32_t max_rows, size_t smem_limit, const double* feat, const int32_t* cnt, float* X,
                            double* stats, cudaStream_t st) {
    if (n_src <= 0) return 0;
    const size_t need = (size_t)max_rows * SHB_F_TOP * SHB_F_NFEAT * sizeof(double);
    if (need + 1024 <= smem_limit) {
        cudaFuncSetAttribute(k_groove_scale<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        k_groove_scale<true><<<n_src, 256, need, st>>>(src, feat, cnt, X, stats);
    } else {
        k_groove_scale<false><<<n_src, 256, 0, st>>>(src, feat, cnt, X, stats);
    }
    return 1;
}
int shb_launch_forest_slots(const float* X, const int32_t* cnt, uint32_t n_rows, uint32_t n_feat, uint32_t n_trees, const uint32_t* root,
                            const int32_t* feature, const float* value, const uint32_t* tchild, const uint32_t* fchild, const float* weight,
                            float* score, cudaStream_t st) {
    if (!n_rows) return 0;
    const size_t tot = (size_t)n_rows * SHB_F_TOP * n_trees;
    k_forest_slots<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(X, cnt, n_rows, n_feat, n_trees, root, feature, value, tchild, fchild, weight, score);
    return 1;
}
int shb_launch_groove_theta_slots(const ShbRowSrc* src, int n_src, uint32_t max_rows, const double* theta, const int32_t* cnt, const float* score,
                                  float threshold, double* bg, double* dens_max, cudaStream_t st) {
    if (n_src <= 0) return 0;
    const size_t smem = (size_t)max_rows * SHB_F_TOP * sizeof(double) + 16;
    cudaFuncSetAttribute(k_groove_theta_slots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_groove_theta_slots<<<n_src, 1024, smem, st>>>(src, theta, cnt, score, threshold, bg, dens_max);
    return 1;
}
}
