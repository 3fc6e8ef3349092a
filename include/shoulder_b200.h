/*
 * shoulder_b200.h — C ABI of the B200 (sm_100a) multiplane slicing / unrolling backend.
 *
 * Drop-in boundary for ONE path of gregspangenberg/shoulder: the `Slices` provider
 * (reference src/shoulder/humerus/slice.py:9-207).  The reference has no FFI of its own for
 * this path — it makes one Python call into trimesh,
 *
 *     self.obb.mesh.section_multiplane(plane_origin=[0,0,z_orig], plane_normal=[0,0,1],
 *                                      heights=z_incrs)            (slice.py:24-28)
 *
 * and then loops over the returned Path2D objects (slice.py:34-147).  The entry points below
 * are what a ctypes binding for that path needs; each cites the reference code it replaces.
 * Plain pointers and sizes only; no torch / numpy types.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SHB_E_* code; shb_last_error() gives
 *     the thread-local message.  Nothing throws across the boundary.
 *   - inputs are caller-owned and only borrowed for the duration of the call.
 *   - results are backend-owned; arrays returned by shb_result_array() are pinned host
 *     buffers valid until shb_result_free().
 *   - per-plane anomalies (no intersection, open contour, non-manifold node) are DATA
 *     (SHB_ARR_STATUS bits), not errors — trimesh returns None / odd paths there too.
 *   - there is no CPU fallback: without a CUDA device every call fails with SHB_E_CUDA.
 */
#ifndef SHOULDER_B200_H
#define SHOULDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHB_ABI_VERSION 4

#if defined(__GNUC__)
#define SHB_API __attribute__((visibility("default")))
#else
#define SHB_API
#endif

/* error codes */
#define SHB_OK            0
#define SHB_E_INVALID    -1   /* bad argument (null pointer, negative size, index out of range) */
#define SHB_E_CUDA       -2   /* CUDA runtime error, or no device */
#define SHB_E_NOMEM      -3
#define SHB_E_CAPACITY   -4   /* a 32-bit internal index would overflow; split the batch */
#define SHB_E_STATE      -5   /* array not computed (not in outputs_mask) / bad handle */

/* outputs_mask bits: which arrays shb_batch_run computes / shb_sweep_batch brings to the host */
#define SHB_OUT_PLANE              0x001u  /* always on: n_seg, n_entities, status, bounds, centroid, area1 */
#define SHB_OUT_SEGMENTS           0x002u  /* mesh_multiplane's lines_2D + face_index          */
#define SHB_OUT_CONTOURS           0x004u  /* Path2D.discrete of every closed entity + areas   */
#define SHB_OUT_IXY                0x008u  /* slice.py:65-80   */
#define SHB_OUT_IXY_CENTERED       0x010u  /* slice.py:85-87   */
#define SHB_OUT_ITR                0x020u  /* slice.py:92-97   (theta-sorted polar)            */
#define SHB_OUT_ITR_START          0x040u  /* slice.py:102-108 (polar rolled to argmin theta)  */
#define SHB_OUT_ITR_CENTERED       0x080u  /* slice.py:124-134 */
#define SHB_OUT_ITR_CENTERED_START 0x100u  /* slice.py:136-144 */
#define SHB_OUT_RADIAL             0x200u  /* extra product: ray-cast radius image, n_angles per plane */
#define SHB_OUT_ALL_PROFILES       0x1F8u
#define SHB_OUT_F32                0x400u  /* deliver the profile arrays and the radius image as float32 (computed in fp64,
                                              rounded on store); halves their HBM and PCIe bytes.  north_star tolerance: 1e-5 */

/* array ids for shb_result_array */
enum shb_array {
    SHB_ARR_N_SEG = 0,        /* int32  [P]        segments on the plane                                */
    SHB_ARR_SEG_OFF,          /* int64  [P+1]      offset of the plane's segments, sweep-relative       */
    SHB_ARR_N_ENT,            /* int32  [P]        len(Path2D.entities): closed contours + open chains  slice.py:53,70 */
    SHB_ARR_STATUS,           /* uint32 [P]        SHB_ST_* bits                                         */
    SHB_ARR_BOUNDS,           /* f64    [P,2,2]    Path2D.bounds                                         */
    SHB_ARR_CENTROID,         /* f64    [P,2]      Path2D.centroid                         slice.py:34-39 */
    SHB_ARR_AREA1,            /* f64    [P]        slice.py:49-60 for planes with one shell (see DESIGN) */
    SHB_ARR_SEL,              /* int32  [P,2]      (contour id, point count) of the outline ixy uses     */
    SHB_ARR_FACE_INDEX,       /* int32  [S]        metadata['face_index'], basic|vertex|edge, ascending  */
    SHB_ARR_SEGMENTS,         /* f64    [S,2,2]    lines_2D                                              */
    SHB_ARR_CONTOUR_OFF,      /* int64  [P+1]      first contour of each plane, sweep-relative.  Contours of a plane come
                                                  in the reference's entity order and start at the reference's start node
                                                  (trimesh graph.traversals: CPython set pop order) */
    SHB_ARR_CONTOUR_PT_OFF,   /* int64  [C+1]      first point of each contour in POINTS                 */
    SHB_ARR_CONTOUR_AREA,     /* f64    [C]        |area| of each closed polygon            slice.py:55-57 */
    SHB_ARR_POINTS,           /* f64    [Npts,2]   Path2D.discrete, CCW, closed (first == last)          */
    SHB_ARR_IXY,              /* f64    [P,2,N] */
    SHB_ARR_IXY_CENTERED,     /* f64    [P,2,N] */
    SHB_ARR_ITR,              /* f64    [P,2,N] */
    SHB_ARR_ITR_START,        /* f64    [P,2,N] */
    SHB_ARR_ITR_CENTERED,     /* f64    [P,2,N] */
    SHB_ARR_ITR_CENTERED_START,/* f64   [P,2,N] */
    SHB_ARR_RADIAL,           /* f64    [P,A]   */
    SHB_ARR_COUNT
};

/* dtype codes written by shb_result_array */
#define SHB_DT_I32 1
#define SHB_DT_I64 2
#define SHB_DT_U32 3
#define SHB_DT_F64 4
#define SHB_DT_F32 5

/* per-plane status bits */
#define SHB_ST_EMPTY        0x01u  /* no face crosses the plane: section_multiplane yields None   */
#define SHB_ST_OPEN         0x02u  /* open chain(s) on the plane (mesh not watertight there): counted in N_ENT, no
                                      contour; closed contours of the same plane are still assembled              */
#define SHB_ST_NONMANIFOLD  0x04u  /* a node has more than two incident segments                   */
#define SHB_ST_RANK_TIE     0x08u  /* two distinct nodes share a rounded-coordinate hash (H4-i)    */
#define SHB_ST_SPLIT_COPY   0x10u  /* two copies of one node round differently (H4-ii)             */
#define SHB_ST_GENERAL      0x20u  /* internal consistency check of the contour ranking failed; contours of the plane
                                      are incomplete (never observed; kept as a guard instead of an out-of-bounds write) */
#define SHB_ST_MERGED       0x40u  /* trimesh Path.__init__ -> merge_vertices fused contour nodes of this plane: consecutive nodes
                                      whose coordinates round equal at digits = |int(log10(1e-8 * scale))| (closer than ~1e-6 mm
                                      on a 10..100 mm section); the delivered contour has fewer points than the plane has segments */

typedef struct shb_batch  shb_batch;
typedef struct shb_result shb_result;

/* Per-sweep request of shb_batch_run_req: which windowed outputs a sweep wants and over which of its planes.
 * Index a = 0..5: the profile arrays in SHB_ARR_IXY .. SHB_ARR_ITR_CENTERED_START order; a = 6: the radius image.
 * Rows [row_lo[a], row_hi[a]) of the sweep are computed and delivered for array a (row_hi < 0: to the last plane) — the
 * consumers' fractional windows (slice.py:157-164: anatomic_neck.py:34 reads 512 of the 600 proximal rows of itr_start,
 * bicipital_groove.py:161 reads 330 of itr_centered_start, canal.py / surgical_neck.py read no profile at all).  Plane
 * records (SHB_OUT_PLANE) always cover every plane.  Replaces nothing in the reference, which computes every row. */
#define SHB_N_WINDOWED 7
typedef struct shb_sweep_request {
    uint32_t outputs_mask;                 /* SHB_OUT_IXY .. SHB_OUT_ITR_CENTERED_START, SHB_OUT_RADIAL bits of this sweep */
    int32_t  row_lo[SHB_N_WINDOWED];
    int32_t  row_hi[SHB_N_WINDOWED];
} shb_sweep_request;

/* Library / device bring-up.  device = CUDA ordinal for this process (one process per GPU).
 * Idempotent.  Replaces nothing in the reference (it is CPU-only). */
SHB_API int shb_init(int device);

/* Kernels are enqueued on this stream (a cudaStream_t; NULL = the library's own stream).  A batch and a result belong to
 * the stream that was current when they were made: later runs on another stream wait for the batch's upload, and their
 * device memory is released on their own stream behind the last work that used it. */
SHB_API int shb_set_stream(void* cuda_stream);

/* Upload a batch of meshes and the sweeps to run on them; inputs become HBM-resident.
 *   verts  (sum V,3) f64 OBB-frame vertices, meshes concatenated; vert_off [n_mesh+1]
 *   faces  (sum T,3) i64 mesh-local vertex ids (as trimesh holds them); face_off [n_mesh+1]
 *   sweep k slices mesh sweep_mesh[k] with the planes  z = z_orig[k] + heights[height_off[k] .. height_off[k+1])
 *   and resamples each outline to interp_num[k] points.
 * The upload is only enqueued: verts / faces must stay valid until the first shb_batch_run on the batch returns
 * (or the batch is freed); an out-of-range face index is reported by that run (SHB_E_INVALID).
 * Limits per batch (SHB_E_CAPACITY beyond them; split the batch): < 2^31 vertices, < 2^29 faces, < 2^31 planes,
 * < 2^32 (sweep, face) pairs, < 2^31 candidate segments (reported by the run).
 * Replaces the argument set of slice.py:24-28 plus Slices.__init__ (slice.py:10-19). */
SHB_API int shb_batch_create(int32_t n_mesh,
                     const double* verts, const int64_t* vert_off,
                     const int64_t* faces, const int64_t* face_off,
                     int32_t n_sweep, const int32_t* sweep_mesh,
                     const double* z_orig,
                     const double* heights, const int64_t* height_off,
                     const int32_t* interp_num,
                     shb_batch** out);
SHB_API int shb_batch_free(shb_batch* batch);

/* Run the whole hot path on the device for every sweep of the batch: bucket -> intersect ->
 * stitch -> resample/unroll.  Results stay on the device until fetched.
 * Replaces slice.py:21-29 (_slices) and the loops of slice.py:34-147. */
SHB_API int shb_batch_run(shb_batch* batch, uint32_t outputs_mask, int32_t n_angles, shb_result** out);

/* Same with one request per sweep (req[n_sweep]; NULL = every sweep takes outputs_mask over all its planes).  The
 * SEGMENTS / CONTOURS / F32 bits of outputs_mask apply to the whole batch.  shb_result_array then hands back only the
 * rows of each sweep's window (shape[0] = row_hi - row_lo); shb_result_window tells which. */
SHB_API int shb_batch_run_req(shb_batch* batch, const shb_sweep_request* req, uint32_t outputs_mask, int32_t n_angles,
                              shb_result** out);
SHB_API int shb_result_window(const shb_result* result, int32_t which, int32_t sweep, int32_t* row_lo, int32_t* row_hi);

/* One-call form with host buffers in and out (create + run + fetch of outputs_mask + free). */
SHB_API int shb_sweep_batch(int32_t n_mesh,
                    const double* verts, const int64_t* vert_off,
                    const int64_t* faces, const int64_t* face_off,
                    int32_t n_sweep, const int32_t* sweep_mesh,
                    const double* z_orig,
                    const double* heights, const int64_t* height_off,
                    const int32_t* interp_num,
                    uint32_t outputs_mask, int32_t n_angles,
                    shb_result** out);

/* Bring the arrays selected by `mask` to pinned host memory (device -> host copy + sync). */
SHB_API int shb_result_fetch(shb_result* result, uint32_t mask);

/* Same, without waiting: the copies are enqueued on the library's copy stream (behind the result's kernels) and the
 * call returns; the next shb_result_fetch / shb_result_array / shb_result_totals / shb_result_free on this result
 * waits for them.  Lets the transfer of one batch overlap the upload and kernels of the next.  Not for contours. */
SHB_API int shb_result_fetch_async(shb_result* result, uint32_t mask);

/* Borrow one array of one sweep.  shape[0..ndim) is filled, *dtype gets a SHB_DT_* code.
 * Returns NULL (and sets the error) if the array was not computed; fetches it if needed. */
SHB_API const void* shb_result_array(shb_result* result, int32_t which, int32_t sweep,
                             int64_t shape[4], int32_t* ndim, int32_t* dtype);

/* Totals of a result: planes, segments, contours, points; device milliseconds per stage of the
 * last run when profiling is enabled (stage order: bucket, scan, scatter, intersect, scan2,
 * stitch, resample).  Any pointer may be NULL. */
SHB_API int shb_result_totals(const shb_result* result, int64_t* n_plane, int64_t* n_seg,
                      int64_t* n_contour, int64_t* n_point);
SHB_API int shb_result_free(shb_result* result);

/* The library caches device memory (stream-ordered pool, never trimmed on its own) and pinned host buffers between
 * calls.  shb_trim() waits for outstanding work and hands everything that is idle back to the driver / OS. */
SHB_API int shb_trim(void);

/* Page-locked host memory from the library's cache, for the caller-owned output buffers of the calls below (shb_neck_image,
 * shb_groove_features, shb_ray_cast, shb_mesh_read ...): a device->host copy into it runs at PCIe speed, into pageable memory at
 * a fraction of it.  shb_host_free hands the block back to the cache (shb_trim releases idle blocks to the OS). */
SHB_API int shb_host_alloc(int64_t bytes, void** out);
SHB_API int shb_host_free(void* p);

/* Per-stage CUDA-event timing of shb_batch_run (one event pair per timed stage, ~3 us of device time each).
 * on = 0: off; 1: every stage; otherwise a mask, bit (s + 1) = time stage s (e.g. 2 << 5 | 2 << 6: stitch and resample). */
#define SHB_N_STAGES 7
SHB_API int shb_profile_enable(int on);
SHB_API int shb_profile_read(double stage_ms[SHB_N_STAGES], int64_t stage_launches[SHB_N_STAGES], int reset);

/* Number of kernels this library has launched since shb_init (bench.py's gpu_launches). */
SHB_API int64_t shb_launch_count(void);

/* ---- resident meshes and single-plane sections ("next" row f1) ---------------------------------------------------------
 * shb_mesh_create uploads one mesh (K0 conversion + face adjacency) and keeps it in HBM; verts / faces may be released
 * when it returns.  shb_batch_create_on builds sweeps on it without an upload; shb_mesh_transform makes a second resident
 * mesh under a 4x4 matrix on the device.  shb_section replaces Trimesh.section(plane_normal, plane_origin) — reference call
 * sites mesh.py:95-99,158-161, surgical_neck.py:37-50, anatomic_neck.py:160-165, arthroplasty.py:71 — for ONE plane with
 * any normal: the result holds plane records and the closed contours in the plane's own 2-D frame, and to_3d (16 doubles,
 * row-major, may be NULL) maps (x, y, 0, 1) back to the mesh frame.  A mesh handle may be freed while batches / results
 * made from it are alive. */
typedef struct shb_mesh shb_mesh;
SHB_API int shb_mesh_create(const double* verts, int64_t n_vert, const int64_t* faces, int64_t n_face, shb_mesh** out);
SHB_API int shb_mesh_free(shb_mesh* mesh);
SHB_API int shb_mesh_transform(shb_mesh* src, const double* to_2d, shb_mesh** out);
SHB_API int shb_batch_create_on(shb_mesh* mesh, int32_t n_sweep, const double* z_orig, const double* heights,
                                const int64_t* height_off, const int32_t* interp_num, shb_batch** out);
SHB_API int shb_section(shb_mesh* mesh, const double* plane_normal, const double* plane_origin, uint32_t outputs_mask,
                        double* to_3d, shb_result** out);

/* ---- mesh input on the device ("next" row f2) -------------------------------------------------------------------------------
 * shb_mesh_from_stl: the bytes of a binary STL file -> a resident mesh.  Replaces MeshLoader._mesh_ct (mesh.py:24,
 * trimesh.load_mesh = load_stl + Trimesh(process=True) -> merge_vertices: corners whose coordinates round equal at 1e-8 are
 * one vertex, numbered by first occurrence, face order kept) and, with SHB_STL_FRAME, the frame step of FullObb._obb
 * (mesh.py:82-117): trimesh's apply_obb (qhull) is not restatable, so the frame is this repo's PCA stand-in (principal axes,
 * smallest variance -> x, largest -> z, AABB centred on the origin) followed by the reference's end test (the rounder end,
 * judged at 0.95 of each z limit, goes to +z).  ASCII STL is rejected (SHB_E_INVALID).
 *   n_vert / n_face   sizes of the welded mesh (may be NULL)
 *   frame_out [22]    may be NULL: [0..15] 4x4 row-major matrix source -> frame (identity without SHB_STL_FRAME),
 *                     [16..17] z bounds in the frame before the end flip (what mesh.py:88 keeps), [18] z length,
 *                     [19] -1 when x and z were negated by the end test else +1, [20..21] circle-fit residuals of the two ends
 * shb_mesh_read copies a resident mesh to the host as (n_vert,3) float64 + (n_face,3) int64 (what trimesh holds). */
#define SHB_STL_FRAME 0x1u
SHB_API int shb_mesh_from_stl(const void* stl_bytes, int64_t n_bytes, uint32_t flags, shb_mesh** out, int64_t* n_vert, int64_t* n_face,
                              double* frame_out);
SHB_API int shb_mesh_read(shb_mesh* mesh, double* verts, int64_t* faces);

/* Ray - mesh queries on a resident mesh ("next" row f4): trimesh mesh.ray.intersects_location as anatomic_neck.py:184-191,
 * 217-224 calls it (four rays per bone).  Every triangle is tested with trimesh's plane / barycentric test (numpy backend);
 * hits come in no particular order: (ray index, triangle index, location [3], distance along the ray).  n_hits receives the
 * number of hits found; SHB_E_CAPACITY when it exceeds max_hits (the first max_hits are delivered). */
SHB_API int shb_ray_cast(shb_mesh* mesh, int32_t n_ray, const double* origins, const double* directions, int32_t max_hits,
                         int32_t* hit_ray, int32_t* hit_tri, double* hit_loc, double* hit_dist, int32_t* n_hits);

/* ---- feature extraction on the polar stacks while they are in HBM ("next" row f3 of the scope table) --------------------
 * The landmark code of the reference loops over the rows of two windows of the proximal sweep right after slice.py hands
 * them over.  These calls run those loops on the device on the rows a result still holds (float64 runs only), for a list
 * of sweeps at once; the outputs are small, so the stacks themselves need not be fetched.  Outputs are caller-owned host
 * buffers, filled when the call returns.  rows = sum over the listed sweeps of their window of the array named below.
 *
 * shb_groove_features: bicipital_groove.py:94-156 on the itr_centered_start window (SHB_OUT_ITR_CENTERED_START): per row
 *   Savitzky-Golay(10, 1), scipy.signal.find_peaks(height=-10, prominence=0.6, width=0.1), the 7 most prominent peaks and
 *   their 9 features BEFORE the StandardScaler (radius, nearest and next-nearest angular distance to another peak, scaled z,
 *   prominence, width, width height, distance to the canal axis, peak count / 7).
 *     zs [rows] z of every row; canal_axes [n_sweep][2][3] Canal.axis() per sweep;
 *     feat [rows][7][9], theta [rows][7], peak_index [rows][7] (sample index in the row, -1 = none), n_peaks [rows].
 * shb_groove_points: bicipital_groove.py:190-238: per row the local minimum of the radius within +-ivar samples of
 *   bg_theta[k] (k = position of the row's sweep in the list), as OBB-frame x, y, z (centroid added) + its theta.
 *     points [rows][3], local_theta [rows].
 * shb_neck_image: anatomic_neck.py:38-58 on the itr_start window (SHB_OUT_ITR_START): rows re-sampled on even theta,
 *   rolled to bg_theta[k], MinMax-scaled per sweep -> image [rows][N] float32 (the UNet input); itr_shft (may be NULL)
 *   [rows][2][N] float64 theta / r as rolled; minmax (may be NULL) [n_sweep][2]. */
SHB_API int shb_groove_features(shb_result* result, int32_t n_sweep, const int32_t* sweeps, const double* zs, const double* canal_axes,
                                double* feat, double* theta, int32_t* peak_index, int32_t* n_peaks);
SHB_API int shb_groove_points(shb_result* result, int32_t n_sweep, const int32_t* sweeps, const double* bg_theta, int32_t ivar,
                              const double* zs, double* points, double* local_theta);
SHB_API int shb_neck_image(shb_result* result, int32_t n_sweep, const int32_t* sweeps, const double* bg_theta, float* image,
                           double* itr_shft, double* minmax);

/* Random forest of the groove detector on the device (bicipital_groove.py:174-181 opens rfc_bg3.onnx, an onnx-ml
 * TreeEnsembleClassifier, through onnxruntime's CPU provider).  Nodes in one flat array, children after their parent;
 * feature < 0 marks a leaf whose `weight` is added to the sample's score (BRANCH_LEQ: x[feature] <= value -> true child).
 * score [n] = sum over trees = probability of class 1 for the binary forests skl2onnx writes. */
typedef struct shb_forest shb_forest;
SHB_API int shb_forest_create(int32_t n_nodes, int32_t n_trees, int32_t n_features, const uint32_t* root, const int32_t* feature,
                              const float* value, const uint32_t* true_child, const uint32_t* false_child, const float* weight,
                              shb_forest** out);
SHB_API int shb_forest_predict(shb_forest* forest, const float* X, int32_t n, float* score);
SHB_API int shb_forest_free(shb_forest* forest);

/* The groove angle of each of n_set bones (bicipital_groove.py:184-188): peaks off[k] .. off[k+1] belong to bone k; those with
 * proba1 > threshold (the reference uses 0.4) enter a linear-kernel density of bandwidth 1 evaluated on np.linspace(-pi, pi, 1024);
 * bg_theta [n_set] receives the arg-max angle (first maximum, like np.argmax), density_max [n_set] (may be NULL) its value. */
SHB_API int shb_groove_theta(int32_t n_set, const int64_t* off, const double* peak_theta, const float* proba1, float threshold,
                             double* bg_theta, double* density_max);

/* The landmark front end of a batch in ONE enqueue and ONE host wait (BASELINE configs[4]): what bone.Humerus runs between
 * slice.py's arrays and its landmark models, for n_bones bones whose sweeps are resident in `result`:
 *   canal axis        canal.py:40-85 (centroids((0.35, 0.75)) of the Full sweep + z -> Line.best_fit, pointed proximally,
 *                     end points at +- half along it) from the plane records on the device;
 *   groove features   as shb_groove_features, the canal axis taken from the step above;
 *   StandardScaler    bicipital_groove.py:171-172, per bone over all its peaks (numpy's summation order) -> float32 X;
 *   forest            as shb_forest_predict, over the slot layout [rows][7] (slots >= n_peaks[row] are skipped);
 *   groove angle      as shb_groove_theta; groove points as shb_groove_points; neck image as shb_neck_image.
 * The polar stacks never leave the device; every output pointer may be NULL (not copied).  Peaks keep the slot layout of
 * shb_groove_features: row-major order over (row, slot < n_peaks[row]) is the order of the reference's per-row appends. */
typedef struct shb_landmark_args {
    int32_t n_bones;
    int32_t canal_lo, canal_hi;       /* rows of the Full sweep used by the canal fit: Slices._cutoff((0.35, 0.75)) */
    const int32_t* full_sweeps;       /* [n_bones] sweep whose plane records feed the canal fit */
    const int32_t* prox_sweeps;       /* [n_bones] sweep holding the itr_start / itr_centered_start windows */
    const double* canal_z;            /* [n_bones][canal_hi - canal_lo] z of those rows (OBB frame) */
    const double* canal_half;         /* [n_bones] obb.z_length * mean(cutoff_pcts) / 2 (canal.py:73-75) */
    const double* groove_zs;          /* z of every row of the itr_centered_start windows, bones concatenated */
    shb_forest* forest;
    float threshold;                  /* bicipital_groove.py:183: 0.4 */
    int32_t ivar;                     /* half window of the local-minimum search, in samples (bicipital_groove.py:196) */
    double* canal_axes;               /* [n_bones][2][3] */
    double* feat;                     /* [rows][7][9] */
    double* peak_theta;               /* [rows][7] */
    int32_t* peak_index;              /* [rows][7] */
    int32_t* n_peaks;                 /* [rows] */
    float* X;                         /* [rows][7][9] scaled features (0 in unused slots) */
    float* proba1;                    /* [rows][7] class-1 score */
    double* scaler;                   /* [n_bones][2][9] mean_, scale_ */
    double* bg_theta;                 /* [n_bones] */
    double* points;                   /* [rows][3] */
    double* local_theta;              /* [rows] */
    float* image;                     /* [image rows][N] */
    double* minmax;                   /* [n_bones][2] */
    uint32_t flags;                   /* SHB_LF_NO_WAIT: return once everything is enqueued (outputs must be page-locked) */
} shb_landmark_args;
#define SHB_LF_NO_WAIT 0x1u
SHB_API int shb_landmark_front(shb_result* result, const shb_landmark_args* args);
/* Waits for the outputs of shb_landmark_front(..., SHB_LF_NO_WAIT) calls on this result (their device->host copies run on the
 * library's copy stream, so the kernels of the next batch overlap them). */
SHB_API int shb_landmark_wait(shb_result* result);

SHB_API const char* shb_last_error(void);
SHB_API int shb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SHOULDER_B200_H */
