/* icp_b200.h -- C ABI of libicp_b200.so, the B200 (sm_100a) ICP alignment engine.
 *
 * Drop-in boundary for ONE path of patmo141/object_alignment: the iterative-closest-point
 * loop  OBJECT_OT_icp_align.execute -> make_pairs -> affine_matrix_from_points
 * (reference: operators/icp_align.py:47-161, functions/general.py:257-329 and :105-217).
 * The reference is pure Python; a maintainer binds these entry points with ctypes
 * (INTEGRATION.md shows the stub).  Plain pointers and sizes only -- no torch, no C++ types.
 *
 * Conventions
 *   - every function returns an icp_status (0 = OK); icp_last_error() gives the message
 *   - all matrices are 4x4 ROW-major (numpy C order; mathutils `M[row][col]`,
 *     operators/icp_align.py:117-119)
 *   - caller owns every host buffer; the library owns all device memory (freed by icp_destroy)
 *   - a handle is bound to one CUDA device and one stream; it is not thread-safe
 *   - multi-GPU = one process (one handle) per GPU; the base vertex set is sharded across ranks
 *     and the per-query nearest neighbour is resolved by an NCCL all-reduce MIN over uint64
 *     keys (float_bits(d2) << 32 | global base index)
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP_B200_VERSION 100 /* 0.1.0 */

typedef enum icp_status {
    ICP_OK = 0,
    ICP_ERR_BAD_ARG = 1,       /* null pointer, negative size, thresh <= 0 (general.py:277), ... */
    ICP_ERR_TOO_FEW_PAIRS = 2, /* fewer than 3 pairs survived: ValueError at general.py:149-157 */
    ICP_ERR_CUDA = 3,
    ICP_ERR_NCCL = 4,
    ICP_ERR_OOM = 5,
    ICP_ERR_STATE = 6,         /* call order: base / moving / matrices not set yet */
    ICP_ERR_SINGULAR = 7       /* matrix_world not invertible (Matrix.inverted() raises) */
} icp_status;

typedef enum icp_nn_mode {
    ICP_NN_BRUTEFORCE = 0, /* O(M N) tiled min-search, the north-star path (general.py:297, D2) */
    ICP_NN_BVH = 1         /* exact search through an implicit BVH over base vertices sorted along a space-filling curve
                              (the "KD-tree equivalent"); bit-identical results */
    , ICP_NN_SURFACE = 2   /* closest point on the TRIANGULATED SURFACE of the base mesh -- the reference's
                              real BVHTree.find_nearest semantics; needs
                              icp_set_base_triangles instead of icp_set_base */
} icp_nn_mode;

typedef enum icp_shard_mode {
    ICP_SHARD_BASE = 0,    /* base vertices split across ranks, queries replicated; per iteration
                              ncclAllReduce MIN over the uint64 keys + SUM over 20 doubles (north star) */
    ICP_SHARD_QUERIES = 1  /* base replicated, every rank owns a slice of the queries (pass only that
                              slice to icp_set_moving); the only exchange is the 20-double SUM.  The
                              right split for the O(log N) searches (BVH / surface)              */
} icp_shard_mode;

/* Parameter block; mirrors ObjectAlignmentPreferences (lib/preferences.py:31-72) as read by
 * execute (operators/icp_align.py:48-49,82-89).  Set struct_size = sizeof(icp_params). */
typedef struct icp_params {
    int32_t struct_size;
    int32_t device;          /* CUDA device ordinal */
    double  thresh;          /* settings.min_start   (icp_align.py:83);  must be > 0            */
    double  target_d;        /* settings.target_d    (icp_align.py:86)                          */
    double  reject_fraction; /* NEW: drop this fraction of largest-d2 pairs; 0 = off */
    int32_t stride;          /* factor = round(1/sample_fraction) (icp_align.py:89), >= 0;
                                values <= 1 mean every vertex (general.py:274)                  */
    int32_t use_target;      /* settings.use_target  (icp_align.py:87): stats + convergence test */
    int32_t align_meth;      /* 0 RIGID, 1 ROT_LOC_SCALE (icp_align.py:106-109)                  */
    int32_t nn_mode;         /* icp_nn_mode                                                      */
    int32_t rank;            /* this process's rank among the base shards (0 when single GPU)   */
    int32_t world_size;      /* number of base shards / GPUs (1 when single GPU)                 */
    int32_t nn_variant;      /* kernel variant for tuning / A-B runs; 0 = library default.  Every variant
                                returns the same keys.  Brute force: 1-6 direct metric, 20-24 expansion
                                filter, 30-36 cylinder filter, 40-45 slab filter (round-1 default), 50-57 /
                                60-73 16-bit quantised slab test (alu-only / split fma+alu loops; default for
                                seeded passes: 71, slots walked in class order with the hard classes served by
                                the expansion-filter kernel; 77 = 71 in plain slot order, 78 = 71 in class order
                                with every class in the q16 kernel, 58 = without the global pruning bound, 79 = greedy
                                tile-tree seeds for the cold pass instead of the exact tree search, 59 =
                                cold pass by the expansion filter instead of tile-tree seeds).
                                Tree (ICP_NN_BVH): default = persistent warp walk over a 32-ary tree of Hilbert-ordered
                                tiles, the whole iteration in one launch; 13 = its first generation (one CTA per
                                128 slots), 12 = binary LBVH packet walk, 10 per-thread walk, 11 packet walk
                                without the fused epilogue.  Surface: 10 per-thread walk, 11 packet walk with
                                eager triangle tests, 14 deferred tests with separate reduce + solve launches
                                (default: deferred tests, the whole iteration in one launch)            */
    int32_t shard_mode;      /* icp_shard_mode when world_size > 1                               */
} icp_params;

/* One record per executed iteration (what execute prints/uses at icp_align.py:129-149). */
typedef struct icp_iter_stats {
    int64_t pairs;     /* K, pairs kept                                   (general.py:302)      */
    double  mean_dist; /* d_stats[0] = mean world distance of kept pairs  (general.py:324)      */
    double  std_dist;  /* d_stats[1] = population std                     (general.py:325)      */
    double  trans;     /* |translation(new_mat)|                          (icp_align.py:138)    */
    double  angle;     /* |rotation angle(new_mat)|, reporting only       (icp_align.py:139)    */
    double  M[16];     /* the fp64 4x4 returned by affine_matrix_from_points (icp_align.py:107) */
} icp_iter_stats;

/* Environment variables read by the library (tuning and diagnostics only; results never depend on them):
 *   ICP_Q16_CLASS_MIN   smallest number of query slots for which the brute-force pass walks the slots in
 *                       class order and leaves the hard classes to the gated expansion-filter launch
 *                       (default 262144; 0 = always)
 *   ICP_Q16_HARD_CLASS  first class that goes to the gated launch when the cloud has no outliers (default 3;
 *                       16 = every class stays in the q16 kernel)
 *   ICP_TRACE_BUILD     when set, host wall-clock time of every one-time index build goes to stderr */

typedef struct icp_handle icp_handle;

int  icp_version(void);
/* Message of the last failing call on this handle (or of the last failing icp_create when h == NULL). */
const char *icp_last_error(const icp_handle *h);
/* Number of visible CUDA devices, or a negative icp_status. */
int  icp_device_count(void);

/* ---- lifetime ------------------------------------------------------------------------- */
int  icp_create(const icp_params *params, icp_handle **out);
int  icp_destroy(icp_handle *h);
int  icp_set_params(icp_handle *h, const icp_params *params); /* device/rank/world are fixed   */

/* ---- inputs ----------------------------------------------------------------------------
 * icp_set_base replaces BVHTree.FromObject(base_obj, depsgraph) (icp_align.py:53): uploads THIS
 * rank's shard of the base vertices (base-local xyz, fp32, n_local x 3).  global_offset is the
 * index of the shard's first vertex in the whole base, n_global the size of the whole base. */
int  icp_set_base(icp_handle *h, const float *xyz, int64_t n_local, int64_t global_offset,
                  int64_t n_global);
/* Surface mode: THIS rank's shard of the base mesh's triangles, as corner coordinates
 * (n_local x 9 floats: a.xyz, b.xyz, c.xyz, base-local space; Blender's loop triangles of the
 * evaluated mesh).  Indices reported by icp_nn are then triangle indices. */
int  icp_set_base_triangles(icp_handle *h, const float *tri_xyz, int64_t n_local, int64_t global_offset,
                            int64_t n_global);
/* Moving (align) object's vertices, align-local xyz fp32 (mesh.vertices.foreach_get("co")), and
 * the vertex list produced by the include/exclude rule (icp_align.py:56-80); vlist == NULL
 * means all vertices (icp_align.py:80).  Indices must be in [0, n_vertices). */
int  icp_set_moving(icp_handle *h, const float *xyz, int64_t n_vertices, const int32_t *vlist,
                    int64_t n_vlist);
/* Collective form for a multi-GPU run in which every rank holds the same host array (world_size > 1,
 * after icp_comm_init; every rank must call it): each rank sends only its 1/world_size slice of the
 * coordinates over its host link and the slices are exchanged with ncclAllGather over NVLink.  Same
 * result as icp_set_moving on every rank. */
int  icp_set_moving_collective(icp_handle *h, const float *xyz, int64_t n_vertices, const int32_t *vlist,
                               int64_t n_vlist);
/* align_obj.matrix_world and base_obj.matrix_world (general.py:262-263), fp32 row-major. */
int  icp_set_matrices(icp_handle *h, const float mw_align[16], const float mw_base[16]);

/* ---- multi-GPU -------------------------------------------------------------------------
 * Rank 0 calls icp_comm_unique_id and ships the 128 bytes to the other ranks by any means;
 * then every rank calls icp_comm_init (collective).  Not needed when world_size == 1. */
int  icp_comm_unique_id(void *id128);
int  icp_comm_init(icp_handle *h, const void *id128);

/* ---- the hot path ----------------------------------------------------------------------
 * icp_nn  ("calc_target"): transform the sampled moving vertices into base-local space
 * (general.py:287) and find each one's nearest base vertex (general.py:297, vertex variant).
 * idx[i] = global base index or -1, d2[i] = fp32 squared base-local distance; both sized
 * icp_query_count().  Either pointer may be NULL. */
int  icp_query_count(const icp_handle *h, int64_t *n_queries);
int  icp_nn(icp_handle *h, int32_t *idx, float *d2);

/* make_pairs (general.py:257-329): A = kept moving vertices, B = their targets brought into
 * align-local space, both 3 x K float64 row-major; d_stats = {mean, std} of kept world
 * distances when calc_stats != 0.  Call with A == B == NULL to learn K first (pairs_out).
 * capacity = number of columns available in A and B. */
int  icp_make_pairs(icp_handle *h, int calc_stats, double *A, double *B, int64_t capacity,
                    int64_t *pairs_out, double d_stats[2]);

/* affine_matrix_from_points(v0, v1, shear=False, scale=<scale>, usesvd=True)
 * (general.py:105-217; the only branch ICP reaches).  v0, v1: 3 x K float64 row-major.
 * Returns ICP_ERR_TOO_FEW_PAIRS when K < 3 (ValueError at general.py:149-157). */
int  icp_solve_pairs(icp_handle *h, const double *v0, const double *v1, int64_t K, int scale,
                     double M_out[16]);

/* One loop body of execute (icp_align.py:98-149): pairs -> M -> matrix_world @= fp32(M) ->
 * convergence ring.  M_out (nullable) receives the fp64 M, stats (nullable) the record. */
int  icp_step(icp_handle *h, double M_out[16], icp_iter_stats *stats);

/* The whole `while n < iters and not converged` loop (icp_align.py:96-151), device-resident:
 * nothing crosses PCIe until it ends.  per_iter (nullable) has room for max_iters records. */
int  icp_run(icp_handle *h, int32_t max_iters, int32_t *iters_done, int32_t *converged,
             icp_iter_stats *per_iter);
/* Forget iteration count / convergence ring (a fresh click of the button, icp_align.py:91-94). */
int  icp_reset_convergence(icp_handle *h);

/* ---- outputs --------------------------------------------------------------------------- */
int  icp_get_matrix_world(icp_handle *h, float mw_align_out[16]);

/* ---- measurement (used by bench.py; CUDA events on the handle's stream) ----------------- */
/* Device time of the last icp_run / icp_step / icp_nn in milliseconds. */
int  icp_last_elapsed_ms(const icp_handle *h, double *ms);
/* With profiling on, every launch of the dominant kernel is bracketed by events.
 * which: 0 = nn search, 1 = filter+moments, 2 = solve, 3 = transform, 4 = collectives. */
int  icp_set_profiling(icp_handle *h, int enabled);
int  icp_kernel_time_ms(const icp_handle *h, int which, double *total_ms, int64_t *launches);
int  icp_launch_count(const icp_handle *h, int64_t *launches); /* kernels launched so far */
int  icp_synchronize(icp_handle *h);
/* Sustained fp32 FMA rate of this device in TFLOP/s (2 flop per lane-FMA), for the roofline
 * of the O(M N) search, which is fp32-issue bound and not HBM bound. */
int  icp_measure_fp32_peak(icp_handle *h, double *tflops);
/* Device-to-device copy bandwidth in GB/s (read + write bytes), for the streaming kernels. */
int  icp_measure_copy_bandwidth(icp_handle *h, double *gbs);
/* Overwrite a > L2-sized scratch buffer (flushes L2 between timed iterations). */
int  icp_flush_l2(icp_handle *h);

/* Debug counters of the BVH search (only counted when nn_variant == 99):
 * out = {inner-node visits, leaf visits, stack pushes, queries}. */
int  icp_debug_counters(icp_handle *h, uint64_t out[4], int reset);

/* ---- pinned host staging buffers (so uploads run at full PCIe rate) ----------------------- */
int  icp_alloc_pinned(size_t bytes, void **out);
int  icp_free_pinned(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
