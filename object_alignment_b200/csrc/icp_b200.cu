// icp_b200.cu -- libicp_b200.so: sm_100a kernels + the C ABI declared in include/icp_b200.h.
//
// One handle = one B200 + one stream.  An ICP iteration (the loop body of
// OBJECT_OT_icp_align.execute, reference operators/icp_align.py:96-151) is this launch train,
// with nothing crossing PCIe and no host synchronisation in between:
//
//   brute force (nn_mode ICP_NN_BRUTEFORCE, the north-star path; base sharded across ranks)
//   k_prepare_pose       imx1 = inv(mw_align), imx2 = inv(mw_base) (only after icp_set_matrices) general.py:262-266
//   k_transform_queries  q = imx2 @ (mx1 @ v.co) for the sampled vlist, seeds, global bound       general.py:274-275,287
//   [k_q16_class_keys + 4-bit radix sort: the slots in class order, from 2^18 slots on]          nn_bruteforce_q16.cuh
//   k_nn_bruteforce_q16  nearest base vertex per query -> uint64 keys                            general.py:297 (D2)
//   [k_nn_bruteforce_filter<GATED>: the hard classes (outliers)]
//   [ncclAllReduce MIN over the keys when the base is sharded across ranks]
//   [radix select of the reject_fraction percentile when enabled]                                NEW
//   k_filter_moments     world distance, thresh, targets, fp64 moment partials                   general.py:299-328,160-181
//   k_reduce_solve       fixed-order sum of the partials, Jacobi SVD -> R, scale, t -> M; matrix_world @= fp32(M);
//                        convergence ring  (multi-GPU: k_reduce_moments, ncclAllReduce SUM over 20 doubles, k_solve_update)
//                                                                                                general.py:181-217, icp_align.py:116-149
//   tree / surface search (ICP_NN_BVH, ICP_NN_SURFACE; queries sharded across ranks)
//   k_nn_tiletree_p<FUSED> / k_nn_surface_defer<FUSED, TICKET>: ALL of the above in one launch (nn_tiletree.cuh)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <vector>

#include "../../include/icp_b200.h"
#include "device_math.cuh"
#include "icp_state.cuh"
#include "nn_bruteforce.cuh"
#include "nn_bruteforce_cyl.cuh"
#include "nn_bruteforce_slab.cuh"
#include "nn_bruteforce_q16.cuh"
#include "nn_bvh.cuh"
#include "nn_tiletree.cuh"
#include "nn_surface.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace icp {

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
__global__ void k_prepare_pose(DevState* st)
{
    if (st->converged) return;
    if (!s2_invert(st->mw_align, st->imx1) || !s2_invert(st->mw_base, st->imx2)) {
        st->error = ICP_ERR_SINGULAR;
        st->converged = 1;
    }
}

// base upload: raw xyz -> tile-blocked SoA (+inf padded) and float4 AoS for gathers
__global__ void k_build_base(const float* __restrict__ xyz, long long n, long long n_padded,
                             float* __restrict__ tiles, float4* __restrict__ base4)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_padded) return;
    const float inf = __int_as_float(0x7f800000);
    float x = inf, y = inf, z = inf;
    if (j < n) {
        x = xyz[3 * j]; y = xyz[3 * j + 1]; z = xyz[3 * j + 2];
        base4[j] = make_float4(x, y, z, 0.0f);
    }
    float* t = tiles + (j / NN_TILE) * (NN_TILE * 3) + (j % NN_TILE);
    t[0] = x; t[NN_TILE] = y; t[2 * NN_TILE] = z;
}

// K1: sampled vertex -> base-local query (two S1 products), SoA out, one query per SLOT.
// slot_pos (nullable = identity) maps a slot to its query position: slots are the Morton order
// of the moving vertices, so that neighbouring threads of the tree search walk the same nodes.
// Keys are reset to "none" -- or, with `warm_src`, to the previous pass's LOCAL winner re-evaluated
// under the new pose: a real base vertex, hence a valid candidate and a tight initial bound
// (tree searches prune with it; the brute-force filter kernel starts its running minimum from it).
__global__ void k_transform_queries(const DevState* __restrict__ st, const float* __restrict__ mov,
                                    const int* __restrict__ vlist, int stride, long long n_q,
                                    const uint32_t* __restrict__ slot_pos,
                                    float* __restrict__ qx, float* __restrict__ qy,
                                    float* __restrict__ qz, unsigned long long* keys,
                                    const unsigned long long* warm_src, const float4* __restrict__ base4,
                                    const float4* __restrict__ tris4, long long base_off,
                                    long long n_local, const unsigned long long* prev_global = nullptr,
                                    float* __restrict__ bound_out = nullptr, unsigned int* __restrict__ cls_scratch = nullptr,
                                    int cls_pass = 0)
{
    // cls_scratch (nullable; launched with 256 threads then): the q16 kernel's class machinery -- this pass's
    // exponent histogram of the pruning radii is accumulated here, where the seeds are made (nn_bruteforce_q16.cuh)
    if (st->converged) return;
    __shared__ unsigned int s_hist[256];
    if (cls_scratch != nullptr) {
        s_hist[threadIdx.x & 255] = 0;
        if (blockIdx.x == 0 && threadIdx.x < (1 << Q16_CLASS_BITS)) cls_scratch[Q16_CLS_COUNT + threadIdx.x] = 0;
        __syncthreads();
    }
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_q) {
    const long long p = slot_pos ? (long long)slot_pos[t] : t;
    const long long vi = vlist ? (long long)vlist[p * stride] : p * stride;
    const float ax = mov[3 * vi], ay = mov[3 * vi + 1], az = mov[3 * vi + 2];
    float wx, wy, wz, x, y, z;
    s1_matvec(st->mw_align, ax, ay, az, wx, wy, wz);
    s1_matvec(st->imx2, wx, wy, wz, x, y, z);
    if (bound_out != nullptr) {
        // Base sharded across ranks: this rank's own previous winner can be far from the query (the
        // global winner lives on another rank), but every rank holds the GLOBAL winner's distance d of
        // the previous pass (prev_global, after the all-reduce) and that pass's query q_old of this slot
        // (still in qx/qy/qz).  The same vertex is at most sqrt(d) + |q - q_old| from the new query
        // (triangle inequality), so nothing farther than that can win or tie: a pruning radius that
        // needs no coordinate exchange.  (1 + 4e-6) covers the fp32 roundings of d, of this
        // expression and of the S4 distance the new winner will be compared with.
        float bd = __int_as_float(0x7f800000);
        const unsigned long long g = prev_global[t];      // may alias keys (same element, read first)
        if (g != KEY_NONE) {
            const float ex = __fsub_rn(x, qx[t]), ey = __fsub_rn(y, qy[t]), ez = __fsub_rn(z, qz[t]);
            const float dq = sqrtf(__fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex))));
            const float sr = __fadd_rn(sqrtf(__uint_as_float((uint32_t)(g >> 32))), dq);
            bd = __fmul_rn(__fmul_rn(sr, sr), 1.000004f);
        }
        bound_out[t] = bd;
    }
    qx[t] = x; qy[t] = y; qz[t] = z;
    unsigned long long key = KEY_NONE;
    if (warm_src) {                                        // may alias keys (same element only)
        const unsigned long long old = warm_src[t];
        const long long j = (long long)(uint32_t)(old & 0xFFFFFFFFull) - base_off;
        if (old != KEY_NONE && j >= 0 && j < n_local) {
            float d;
            if (tris4) {                                   // surface mode: previous winning triangle
                const float4 a = tris4[3 * j], b = tris4[3 * j + 1], c = tris4[3 * j + 2];
                float rx, ry, rz;
                closest_on_tri(x, y, z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, rx, ry, rz);
                d = s10_d2(x, y, z, rx, ry, rz);
            } else {
                const float4 b = base4[j];
                d = s4_d2(x, y, z, b.x, b.y, b.z);
            }
            if (d < __int_as_float(0x7f800000))
                key = ((unsigned long long)__float_as_uint(d) << 32) | (old & 0xFFFFFFFFull);
        }
    }
    keys[t] = key;
    if (cls_scratch != nullptr)
        atomicAdd(&s_hist[q16_exponent_of(key, bound_out != nullptr ? bound_out[t] : __int_as_float(0x7f800000))], 1u);
    }
    if (cls_scratch != nullptr) {
        __syncthreads();
        const unsigned int c = s_hist[threadIdx.x & 255];
        if (c) atomicAdd(&cls_scratch[Q16_CLS_EXP + 256 * (cls_pass & 1) + (threadIdx.x & 255)], c);
    }
}

// Morton code of every sampled moving vertex (align-local space; a rigid pose keeps the order useful)
__global__ void k_query_bbox_partial(const float* __restrict__ mov, const int* __restrict__ vlist, int stride,
                                     long long n_q, float* __restrict__ part)
{
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_q;
         p += (long long)gridDim.x * blockDim.x) {
        const long long vi = vlist ? (long long)vlist[p * stride] : p * stride;
        for (int k = 0; k < 3; ++k) { const float c = mov[3 * vi + k]; lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c); }
    }
    __shared__ float sh[6][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
        for (int o = 16; o > 0; o >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if (lane == 0) { sh[k][warp] = a; sh[3 + k][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        float v = sh[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = (k < 3) ? fminf(v, sh[k][w]) : fmaxf(v, sh[k][w]);
        part[blockIdx.x * 6 + k] = v;
    }
}
__global__ void k_query_codes(const float* __restrict__ mov, const int* __restrict__ vlist, int stride,
                              long long n_q, const float* __restrict__ bbox,
                              unsigned long long* __restrict__ codes, uint32_t* __restrict__ pos)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_q) return;
    const long long vi = vlist ? (long long)vlist[p * stride] : p * stride;
    uint32_t g[3];
    for (int k = 0; k < 3; ++k) {
        const float ext = bbox[3 + k] - bbox[k];
        float t = ext > 0.0f ? (mov[3 * vi + k] - bbox[k]) / ext : 0.0f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        g[k] = (uint32_t)fminf(t * 2097152.0f, 2097151.0f);
    }
    codes[p] = curve_code21(g[0], g[1], g[2]);
    pos[p] = (uint32_t)p;
}
__global__ void k_gather_mov_slots(const float* __restrict__ mov, const int* __restrict__ vlist, int stride,
                                   const uint32_t* __restrict__ slot_pos, long long n_q, float4* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_q) return;
    const long long p = slot_pos[t];
    const long long vi = vlist ? (long long)vlist[p * stride] : p * stride;
    out[t] = make_float4(mov[3 * vi], mov[3 * vi + 1], mov[3 * vi + 2], __uint_as_float((uint32_t)p));
}
__global__ void k_invert_permutation(const uint32_t* __restrict__ slot_pos, long long n, uint32_t* __restrict__ pos_slot)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) pos_slot[slot_pos[t]] = (uint32_t)t;
}

// ---- percentile (reject_fraction) : exact radix select of the K-th smallest composite key
//      composite(p) = d2_bits(p) << 32 | p          (NEW semantics)
__global__ void k_select_init(DevState* st, long long rank, unsigned int* hist)
{
    const int t = threadIdx.x;
    if (t < 256) hist[t] = 0;
    if (t == 0) {
        st->sel_prefix = 0; st->sel_mask = 0; st->sel_rank = rank; st->sel_threshold = 0;
        st->sel_valid = rank >= 0 ? 1 : 0;
    }
}
__global__ void k_select_hist(const DevState* __restrict__ st,
                              const unsigned long long* __restrict__ keys,
                              const uint32_t* __restrict__ slot_pos, long long n, int byte,
                              unsigned int* __restrict__ hist)
{
    if (st->converged) return;
    __shared__ unsigned int sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned long long prefix = st->sel_prefix, mask = st->sel_mask;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        const unsigned long long p = slot_pos ? (unsigned long long)slot_pos[t] : (unsigned long long)t;
        const unsigned long long c = (keys[t] & 0xFFFFFFFF00000000ull) | p;
        if ((c & mask) == prefix) atomicAdd(&sh[(c >> (8 * byte)) & 0xFF], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void k_select_pick(DevState* st, int byte, unsigned int* hist)
{
    if (st->converged) return;
    if (threadIdx.x == 0 && st->sel_valid) {
        long long r = st->sel_rank;
        int d = 0;
        for (; d < 256; ++d) {
            const long long c = hist[d];
            if (r < c) break;
            r -= c;
        }
        if (d == 256) { d = 255; st->sel_valid = 0; }
        st->sel_rank = r;
        st->sel_prefix |= (unsigned long long)d << (8 * byte);
        st->sel_mask |= 0xFFull << (8 * byte);
        if (byte == 0) st->sel_threshold = st->sel_prefix;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
}

// ---- K3 + K4 : filter and fp64 moments ------------------------------------------------------
struct PairEval {
    bool keep;
    float ax, ay, az, bx, by, bz;
    double dist;
};

// everything the reference does for one vertex after find_nearest (general.py:299-306)
__device__ __forceinline__ PairEval eval_pair(const DevState* __restrict__ st,
                                              const float* __restrict__ mov,
                                              const int* __restrict__ vlist, int stride,
                                              const float* __restrict__ qx,
                                              const float* __restrict__ qy,
                                              const float* __restrict__ qz,
                                              const unsigned long long* __restrict__ keys,
                                              const float4* __restrict__ base4,
                                              const float4* __restrict__ tris4, long long base_off,
                                              long long n_local, double thresh, int use_select,
                                              long long t, long long p)
{
    // t = slot (where q and the key live), p = query position (what vlist and the percentile use)
    PairEval e;
    e.keep = false;
    const unsigned long long key = keys[t];
    if (key == KEY_NONE) return e;
    const long long j = (long long)(uint32_t)(key & 0xFFFFFFFFull) - base_off;
    if (j < 0 || j >= n_local) return e;                        // another rank owns this target
    if (use_select) {
        const unsigned long long c = (key & 0xFFFFFFFF00000000ull) | (unsigned long long)p;
        if (!st->sel_valid || c > st->sel_threshold) return e;
    }
    float4 co1;
    if (tris4) {       // surface mode: the target is the closest point on the winning triangle (S9)
        const float4 a = tris4[3 * j], b = tris4[3 * j + 1], c = tris4[3 * j + 2];
        closest_on_tri(qx[t], qy[t], qz[t], a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, co1.x, co1.y, co1.z);
    } else {
        co1 = base4[j];
    }
    float wqx, wqy, wqz, w1x, w1y, w1z;
    s1_matvec(st->mw_base, qx[t], qy[t], qz[t], wqx, wqy, wqz);
    s1_matvec(st->mw_base, co1.x, co1.y, co1.z, w1x, w1y, w1z);
    e.dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
    if (!(e.dist < thresh)) return e;                            // strict, general.py:302
    const long long vi = vlist ? (long long)vlist[p * stride] : p * stride;
    e.ax = mov[3 * vi]; e.ay = mov[3 * vi + 1]; e.az = mov[3 * vi + 2];
    s1_matvec(st->imx1, w1x, w1y, w1z, e.bx, e.by, e.bz);        // general.py:304
    e.keep = true;
    return e;
}

__global__ void __launch_bounds__(MOM_THREADS)
k_filter_moments(const DevState* __restrict__ st, const float* __restrict__ mov,
                 const int* __restrict__ vlist, int stride, long long n_q,
                 const float* __restrict__ qx, const float* __restrict__ qy,
                 const float* __restrict__ qz, const unsigned long long* __restrict__ keys,
                 const float4* __restrict__ base4, const float4* __restrict__ tris4, long long base_off,
                 long long n_local, double thresh, int use_select, const uint32_t* __restrict__ slot_pos,
                 const float4* __restrict__ mov_slots, double* __restrict__ partials)
{
    // mov_slots (nullable; present whenever slot_pos is): float4(moving vertex, query position bits) per
    // slot, so the moving coordinates stream in slot order instead of being gathered through slot_pos.
    if (st->converged) return;
    double acc[MOM];
#pragma unroll
    for (int i = 0; i < MOM; ++i) acc[i] = 0.0;
    const double s0 = st->shift[0], s1 = st->shift[1], s2 = st->shift[2];
    // Each CTA owns one contiguous, 32-aligned range of slots (balanced to within a warp), and each
    // thread keeps MOM_BATCH queries in flight: the key -> winner-gather chain is two dependent DRAM
    // round trips, so the loads of a batch are issued together, stage by stage, before any arithmetic.
    const long long per = ((n_q + gridDim.x - 1) / gridDim.x + 31) & ~31ll;
    const long long lo = (long long)blockIdx.x * per;
    const long long hi = lo + per < n_q ? lo + per : n_q;
    if (tris4 != nullptr) {             // surface mode: three gathers + region walk per pair, unbatched
        for (long long t = lo + threadIdx.x; t < hi; t += MOM_THREADS) {
            const long long p = slot_pos ? (long long)slot_pos[t] : t;
            const PairEval e = eval_pair(st, mov, vlist, stride, qx, qy, qz, keys, base4, tris4, base_off,
                                         n_local, thresh, use_select, t, p);
            if (!e.keep) continue;
            moments_add(acc, e.ax, e.ay, e.az, e.bx, e.by, e.bz, e.dist, s0, s1, s2);
        }
        block_reduce_moments<MOM_THREADS>(acc, partials);
        return;
    }
    const bool sel_ok = use_select ? (st->sel_valid != 0) : true;
    const unsigned long long sel_thr = use_select ? st->sel_threshold : ~0ull;
    for (long long t0 = lo + threadIdx.x; t0 < hi; t0 += (long long)MOM_THREADS * MOM_BATCH) {
        unsigned long long key[MOM_BATCH];
        long long p[MOM_BATCH];
        float x[MOM_BATCH], y[MOM_BATCH], z[MOM_BATCH];
        float ax[MOM_BATCH], ay[MOM_BATCH], az[MOM_BATCH];
        bool ok[MOM_BATCH];
#pragma unroll
        for (int k = 0; k < MOM_BATCH; ++k) {                 // stage 1: streaming loads
            const long long t = t0 + (long long)k * MOM_THREADS;
            ok[k] = t < hi;
            key[k] = KEY_NONE; p[k] = t; x[k] = y[k] = z[k] = 0.0f;
            ax[k] = ay[k] = az[k] = 0.0f;
            if (ok[k]) {
                key[k] = keys[t];
                if (mov_slots) {
                    const float4 a = mov_slots[t];
                    ax[k] = a.x; ay[k] = a.y; az[k] = a.z;
                    p[k] = (long long)__float_as_uint(a.w);
                } else if (slot_pos) {
                    p[k] = (long long)slot_pos[t];
                }
                x[k] = qx[t]; y[k] = qy[t]; z[k] = qz[t];
            }
        }
        float4 co1[MOM_BATCH];
#pragma unroll
        for (int k = 0; k < MOM_BATCH; ++k) {                 // stage 2: gathers
            const long long j = (long long)(uint32_t)(key[k] & 0xFFFFFFFFull) - base_off;
            ok[k] = ok[k] && key[k] != KEY_NONE && j >= 0 && j < n_local;   // else another rank owns the target
            if (use_select) {
                const unsigned long long c = (key[k] & 0xFFFFFFFF00000000ull) | (unsigned long long)p[k];
                ok[k] = ok[k] && sel_ok && c <= sel_thr;
            }
            co1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[k]) {
                co1[k] = base4[j];
                if (!mov_slots) {
                    const long long vi = vlist ? (long long)vlist[p[k] * stride] : p[k] * stride;
                    ax[k] = mov[3 * vi]; ay[k] = mov[3 * vi + 1]; az[k] = mov[3 * vi + 2];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < MOM_BATCH; ++k) {                 // stage 3: general.py:299-306, in slot order
            if (!ok[k]) continue;
            float wqx, wqy, wqz, w1x, w1y, w1z, bx, by, bz;
            s1_matvec(st->mw_base, x[k], y[k], z[k], wqx, wqy, wqz);
            s1_matvec(st->mw_base, co1[k].x, co1[k].y, co1[k].z, w1x, w1y, w1z);
            const double dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
            if (!(dist < thresh)) continue;                   // strict, general.py:302
            s1_matvec(st->imx1, w1x, w1y, w1z, bx, by, bz);   // general.py:304
            moments_add(acc, ax[k], ay[k], az[k], bx, by, bz, dist, s0, s1, s2);
        }
    }
    block_reduce_moments<MOM_THREADS>(acc, partials);
}

// moments of explicit pair arrays (affine_matrix_from_points entry; v0, v1 are 3 x K row-major)
__global__ void __launch_bounds__(MOM_THREADS)
k_pairs_moments(const double* __restrict__ v0, const double* __restrict__ v1, long long K,
                const DevState* __restrict__ st, double* __restrict__ partials)
{
    double acc[MOM];
#pragma unroll
    for (int i = 0; i < MOM; ++i) acc[i] = 0.0;
    const double s0 = st->shift[0], s1 = st->shift[1], s2 = st->shift[2];
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K;
         k += (long long)gridDim.x * blockDim.x) {
        const double a0 = v0[k] - s0, a1 = v0[K + k] - s1, a2 = v0[2 * K + k] - s2;
        const double b0 = v1[k] - s0, b1 = v1[K + k] - s1, b2 = v1[2 * K + k] - s2;
        acc[0] += 1.0;
        acc[1] += a0; acc[2] += a1; acc[3] += a2;
        acc[4] += b0; acc[5] += b1; acc[6] += b2;
        acc[7] += b0 * a0; acc[8] += b0 * a1; acc[9] += b0 * a2;
        acc[10] += b1 * a0; acc[11] += b1 * a1; acc[12] += b1 * a2;
        acc[13] += b2 * a0; acc[14] += b2 * a1; acc[15] += b2 * a2;
        acc[16] += a0 * a0 + a1 * a1 + a2 * a2;
        acc[17] += b0 * b0 + b1 * b1 + b2 * b2;
    }
    __shared__ double sh[MOM_THREADS / 32][MOM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MOM; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) sh[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < MOM) {
        double v = 0.0;
        for (int w = 0; w < MOM_THREADS / 32; ++w) v += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * MOM + threadIdx.x] = v;
    }
}

// fixed-order sum of the per-CTA partials -> moments[MOM]   (CTAs of MOM warps; CTA b sums the slice
// [b * per_cta, (b + 1) * per_cta) into out[b][MOM]; the last stage is one CTA over everything left)
__global__ void k_reduce_moments(const DevState* __restrict__ st, const double* __restrict__ partials,
                                 int n_partials, double* __restrict__ moments, int per_cta = 0x7fffffff)
{
    if (st != nullptr && st->converged) return;
    const int lane = threadIdx.x & 31, comp = threadIdx.x >> 5;
    if (comp >= MOM) return;
    const long long lo = (long long)blockIdx.x * per_cta;
    const long long hi = lo + per_cta < n_partials ? lo + per_cta : n_partials;
    double v = 0.0;
    for (long long i = lo + lane; i < hi; i += 32) v += partials[(size_t)i * MOM + comp];
    v = warp_sum(v);
    if (lane == 0) moments[(size_t)blockIdx.x * MOM + comp] = v;
}

// Kabsch/Umeyama step from the moments (general.py:160-217 for shear=False, usesvd=True)
__device__ inline bool solve_from_moments(const double* mom, const double* shift, int scale_mode,
                                          double* M /*4x4*/, double* mean_d, double* std_d)
{
    const double K = mom[0];
    if (!(K >= 3.0)) return false;
    double ca[3], cb[3];
    for (int i = 0; i < 3; ++i) { ca[i] = mom[1 + i] / K; cb[i] = mom[4 + i] / K; }
    double h[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) h[i * 3 + j] = mom[7 + i * 3 + j] - K * cb[i] * ca[j];
    double rot[9];
    rotation_from_covariance(h, rot);
    if (scale_mode) {
        const double va = mom[16] - K * (ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2]);
        const double vb = mom[17] - K * (cb[0] * cb[0] + cb[1] * cb[1] + cb[2] * cb[2]);
        const double s = sqrt(vb / va);                          // general.py:208-212
        for (int i = 0; i < 9; ++i) rot[i] *= s;
    }
    // un-shift the centroids, then t = cb - R ca                 (general.py:215-217)
    for (int i = 0; i < 3; ++i) { ca[i] += shift[i]; cb[i] += shift[i]; }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) M[i * 4 + j] = rot[i * 3 + j];
        M[i * 4 + 3] = cb[i] - (rot[i * 3] * ca[0] + rot[i * 3 + 1] * ca[1] + rot[i * 3 + 2] * ca[2]);
    }
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
    const double m = mom[18] / K;
    double var = mom[19] / K - m * m;
    *mean_d = m;
    *std_d = sqrt(var > 0 ? var : 0.0);
    return true;
}

__device__ inline double rotation_angle_of(const float* nm)
{
    double r[9], tr = 0;
    for (int j = 0; j < 3; ++j) {
        double n = 0;
        for (int i = 0; i < 3; ++i) n += (double)nm[i * 4 + j] * (double)nm[i * 4 + j];
        n = sqrt(n);
        if (n == 0) n = 1;
        for (int i = 0; i < 3; ++i) r[i * 3 + j] = (double)nm[i * 4 + j] / n;
    }
    tr = r[0] + r[4] + r[8];
    double c = (tr - 1.0) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    return acos(c);
}

// K5: solve, pose update (icp_align.py:116-121), record, convergence ring (icp_align.py:136-142)
__device__ void solve_update_body(DevState* st, const double* __restrict__ moments, int scale_mode,
                                  int use_target, double target_d, icp_iter_stats* __restrict__ stats,
                                  int stats_capacity);

__global__ void k_solve_update(DevState* st, const double* __restrict__ moments, int scale_mode,
                               int use_target, double target_d, icp_iter_stats* __restrict__ stats,
                               int stats_capacity)
{
    if (threadIdx.x != 0 || st->converged) return;
    solve_update_body(st, moments, scale_mode, use_target, target_d, stats, stats_capacity);
}

__device__ __noinline__ void solve_update_body(DevState* st, const double* __restrict__ moments, int scale_mode,
                                               int use_target, double target_d, icp_iter_stats* __restrict__ stats,
                                               int stats_capacity)
{
    double M[16], mean_d, std_d;
    if (!solve_from_moments(moments, st->shift, scale_mode, M, &mean_d, &std_d)) {
        st->error = ICP_ERR_TOO_FEW_PAIRS;
        st->converged = 1;
        return;
    }
    float nm[16];
    for (int i = 0; i < 16; ++i) { st->M[i] = M[i]; nm[i] = __double2float_rn(M[i]); }
    float mw[16];
    s3_matmul(st->mw_align, nm, mw);
    for (int i = 0; i < 16; ++i) st->mw_align[i] = mw[i];
    const double tlen = s5_length(nm[3], nm[7], nm[11]);
    const int n = st->n_iter;
    if (stats != nullptr && n < stats_capacity) {
        icp_iter_stats& s = stats[n];
        s.pairs = (int64_t)moments[0];
        s.mean_dist = mean_d; s.std_dist = std_d; s.trans = tlen;
        s.angle = rotation_angle_of(nm);
        for (int i = 0; i < 16; ++i) s.M[i] = M[i];
    }
    if (use_target) {
        st->ring[n % 5] = tlen;
        bool all = true;
        for (int i = 0; i < 5; ++i) all = all && (st->ring[i] < target_d);
        if (all) st->converged = 1;
    }
    st->n_iter = n + 1;
    // the next iteration's mx1.inverted() (general.py:265): mw_base has not changed, so imx2 stands
    if (!s2_invert(st->mw_align, st->imx1)) {
        st->error = ICP_ERR_SINGULAR;
        st->converged = 1;
    }
}

// Single-GPU tail of an iteration in ONE launch: fixed-order sum of the per-CTA moment partials (one
// warp per component), then the solve + pose update by thread 0.
__global__ void __launch_bounds__(MOM * 32)
k_reduce_solve(DevState* st, const double* __restrict__ partials, int n_partials, double* __restrict__ moments,
               int scale_mode, int use_target, double target_d, icp_iter_stats* __restrict__ stats, int stats_capacity);

__global__ void __launch_bounds__(MOM * 32)
k_reduce_solve(DevState* st, const double* __restrict__ partials, int n_partials, double* __restrict__ moments,
               int scale_mode, int use_target, double target_d, icp_iter_stats* __restrict__ stats, int stats_capacity)
{
    if (st->converged) return;
    const int lane = threadIdx.x & 31, comp = threadIdx.x >> 5;
    double v = 0.0;
    for (int i = lane; i < n_partials; i += 32) v += partials[(size_t)i * MOM + comp];
    v = warp_sum(v);
    if (lane == 0) moments[comp] = v;
    __syncthreads();
    if (threadIdx.x == 0) solve_update_body(st, moments, scale_mode, use_target, target_d, stats, stats_capacity);
}

// solve only (affine_matrix_from_points entry point): no pose update
__global__ void k_solve_only(DevState* st, const double* __restrict__ moments, int scale_mode)
{
    if (threadIdx.x != 0) return;
    double M[16], mean_d, std_d;
    if (!solve_from_moments(moments, st->shift, scale_mode, M, &mean_d, &std_d)) {
        st->error = ICP_ERR_TOO_FEW_PAIRS;
        return;
    }
    for (int i = 0; i < 16; ++i) st->M[i] = M[i];
}

// ---- make_pairs emission: count, scan, scatter ------------------------------------------------
constexpr int EMIT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS)
k_pairs_count(const DevState* __restrict__ st, const float* __restrict__ mov,
              const int* __restrict__ vlist, int stride, long long n_q, const float* __restrict__ qx,
              const float* __restrict__ qy, const float* __restrict__ qz,
              const unsigned long long* __restrict__ keys, const float4* __restrict__ base4,
              const float4* __restrict__ tris4,
              long long base_off, long long n_local, double thresh, int use_select,
              const uint32_t* __restrict__ pos_slot,
              unsigned char* __restrict__ flags, long long* __restrict__ block_counts)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (p < n_q) {
        const long long t = pos_slot ? (long long)pos_slot[p] : p;
        keep = eval_pair(st, mov, vlist, stride, qx, qy, qz, keys, base4, tris4, base_off, n_local, thresh,
                         use_select, t, p).keep;
        flags[p] = keep ? 1 : 0;
    }
    const int c = __syncthreads_count(keep ? 1 : 0);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void k_scan_counts(long long* __restrict__ block_counts, int n_blocks, long long* total)
{
    // single thread: n_blocks <= a few 10k
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long run = 0;
    for (int i = 0; i < n_blocks; ++i) { const long long c = block_counts[i]; block_counts[i] = run; run += c; }
    *total = run;
}

__global__ void __launch_bounds__(EMIT_THREADS)
k_pairs_emit(const DevState* __restrict__ st, const float* __restrict__ mov,
             const int* __restrict__ vlist, int stride, long long n_q, const float* __restrict__ qx,
             const float* __restrict__ qy, const float* __restrict__ qz,
             const unsigned long long* __restrict__ keys, const float4* __restrict__ base4,
             const float4* __restrict__ tris4,
             long long base_off, long long n_local, double thresh, int use_select,
             const uint32_t* __restrict__ pos_slot,
             const long long* __restrict__ block_offsets, long long K, double* __restrict__ A,
             double* __restrict__ B, double* __restrict__ dist_out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PairEval e;
    e.keep = false;
    if (p < n_q)
        e = eval_pair(st, mov, vlist, stride, qx, qy, qz, keys, base4, tris4, base_off, n_local, thresh,
                      use_select, pos_slot ? (long long)pos_slot[p] : p, p);
    __shared__ int warp_counts[EMIT_THREADS / 32];
    const unsigned ballot = __ballot_sync(0xffffffffu, e.keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_counts[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (e.keep) {
        const long long k = block_offsets[blockIdx.x] + before;
        A[k] = e.ax; A[K + k] = e.ay; A[2 * K + k] = e.az;
        B[k] = e.bx; B[K + k] = e.by; B[2 * K + k] = e.bz;
        dist_out[k] = e.dist;
    }
}

// mean / population std of the kept distances, two-pass like np.mean / np.std (general.py:324-325)
__global__ void k_dist_stats(const double* __restrict__ dist, long long K, double* __restrict__ out)
{
    __shared__ double sh[32];
    __shared__ double mean_s;
    double s = 0;
    for (long long i = threadIdx.x; i < K; i += blockDim.x) s += dist[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        mean_s = t / (double)K;
    }
    __syncthreads();
    const double m = mean_s;
    double v = 0;
    for (long long i = threadIdx.x; i < K; i += blockDim.x) { const double d = dist[i] - m; v += d * d; }
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        out[0] = m;
        out[1] = sqrt(t / (double)K);
    }
}

__global__ void k_unpack_keys(const unsigned long long* __restrict__ keys,
                              const uint32_t* __restrict__ pos_slot, long long n,
                              int* __restrict__ idx, float* __restrict__ d2)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned long long k = keys[pos_slot ? (long long)pos_slot[p] : p];
    if (idx) idx[p] = (k == KEY_NONE) ? -1 : (int)(uint32_t)(k & 0xFFFFFFFFull);
    if (d2) d2[p] = __uint_as_float((uint32_t)(k >> 32));
}

// ---- measurement helpers ----------------------------------------------------------------------
__global__ void k_fma_peak(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
          a6 = a0 + 6, a7 = a0 + 7;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}
__global__ void k_fma2_peak(float* out, int iters)
{
    uint64_t a0 = pk2(threadIdx.x * 1e-3f, 1.f), a1 = pk2(2.f, 3.f), a2 = pk2(4.f, 5.f),
             a3 = pk2(6.f, 7.f), a4 = pk2(8.f, 9.f), a5 = pk2(1.5f, 2.5f), a6 = pk2(3.5f, 4.5f),
             a7 = pk2(5.5f, 6.5f);
    const uint64_t m = pk2(0.999f, 0.998f), c = pk2(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
        a0 = fma2(a0, m, c); a1 = fma2(a1, m, c); a2 = fma2(a2, m, c); a3 = fma2(a3, m, c);
        a4 = fma2(a4, m, c); a5 = fma2(a5, m, c); a6 = fma2(a6, m, c); a7 = fma2(a7, m, c);
    }
    float lo, hi, s = 0;
    unpk2(a0, lo, hi); s += lo + hi; unpk2(a1, lo, hi); s += lo + hi;
    unpk2(a2, lo, hi); s += lo + hi; unpk2(a3, lo, hi); s += lo + hi;
    unpk2(a4, lo, hi); s += lo + hi; unpk2(a5, lo, hi); s += lo + hi;
    unpk2(a6, lo, hi); s += lo + hi; unpk2(a7, lo, hi); s += lo + hi;
    if (s == 12345.678f) out[0] = s;
}
__global__ void k_copy(const float4* __restrict__ src, float4* __restrict__ dst, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace icp

// ==============================================================================================
// host side
// ==============================================================================================
using namespace icp;

namespace {

char g_create_error[512] = "";

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(char* err, size_t n)
    {
        if (lib) return true;
        // RTLD_NOLOAD first: reuse an NCCL already in the process (e.g. the one torch bundles)
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { snprintf(err, n, "dlopen(libnccl.so.2) failed: %s", dlerror()); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy || !GetErrorString) {
            snprintf(err, n, "libnccl is missing required symbols");
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;

struct EventPair { cudaEvent_t a, b; };

}  // namespace

struct icp_handle {
    icp_params prm{};
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // base
    float* d_base_tiles = nullptr;
    float4* d_base4 = nullptr;
    long long n_local = 0, n_global = 0, base_off = 0;
    int n_tiles = 0;
    bool have_base = false;
    // surface mode (ICP_NN_SURFACE): triangles of the base mesh + their LBVH
    bool surface = false;
    float4* d_tris4 = nullptr;             // 3 float4 per triangle, original order
    float4* d_tri_sorted = nullptr;        // Morton order, run-padded
    BvhNode* d_tri_nodes = nullptr;
    int tri_runs = 0;
    float tri_extent = 0.0f;
    // expansion-form filter tiles for the brute-force kernel (nn_variant 20..), built lazily
    float* d_filter_tiles = nullptr;
    FilterInfo* d_filter_info = nullptr;
    bool filter_built = false;
    // cylinder-filter tiles (nn_variant 30..): Morton-ordered, centred per tile, built lazily
    float* d_cyl_ftiles = nullptr;         // [tile][a'|b'|w2|c'][1024]
    float* d_cyl_etiles = nullptr;         // [tile][x|y|z|original index][1024]
    CylTileHdr* d_cyl_hdrs = nullptr;
    int cyl_axes = 0;                      // a | b << 2 | c << 4
    int cyl_sorted = -1;                   // how the current tiles were built (-1: not built)
    float* d_slab_ftiles = nullptr;        // [tile][a'|b'|w2|c'|w1][1024], vertices sorted along a inside a tile
    float* d_slab_etiles = nullptr;        // [tile][x|y|z|original index][1024], same order
    CylTileHdr* d_slab_hdrs = nullptr;
    int slab_axes = 0;
    bool slots_stale = false;              // d_mov_slots must be re-gathered (same slot order, new coordinates)
    int slab_order = -1;                   // -1 not built, 0 sorted along a inside a tile, 1 Morton order inside a tile
    // 16-bit quantised tiles of the q16 kernel (nn_variant 50..): [tile][a|b|c u16 planes, header]
    uint32_t* d_q16_tiles = nullptr;
    float* d_q16_etiles = nullptr;         // [tile][x|y|z|original index][1024], same order
    int q16_axes = 0;
    int q16_order = -1;                    // -1 not built, 0 sorted along a inside a tile, 1 Morton order inside a tile
    // base sharded across ranks: squared pruning radius per slot from the GLOBAL previous winner
    float* d_bound = nullptr;
    long long bound_capacity = 0;
    bool bound_active = false;             // the pass being enqueued has a valid d_bound
    // q16 query classes: slots in class order (similar pruning radius together), rebuilt every pass
    unsigned char* d_q16_cls = nullptr;    // cls | cls_out
    uint32_t* d_q16_ids = nullptr;         // ids | perm
    unsigned int* d_q16_hist = nullptr;    // Q16_CLS_WORDS: two exponent histograms (alternate passes) + the class populations
    int q16_pass = 0;                      // parity selects the exponent histogram
    bool q16_hist_by_k1 = false;           // K1 of the pass being enqueued has accumulated the histogram
    void* d_q16_sort_temp = nullptr;
    size_t q16_sort_temp_bytes = 0;
    long long q16_class_capacity = 0;
    std::vector<const void*> smem_configured;   // kernels whose dynamic-smem limit was raised on THIS device
    // spatial index (nn_mode == ICP_NN_BVH), built lazily from d_base4
    BvhNode* d_bvh_nodes = nullptr;
    float4* d_bvh_sorted = nullptr;
    int bvh_P = 0;
    bool bvh_built = false;
    // 32-ary tree of Morton tiles over the same sorted vertices (nn_tiletree.cuh; default tree search)
    float4* d_tt_boxes = nullptr;
    float* d_tt_tiles = nullptr;
    TileTree* d_tt_desc = nullptr;
    TileTree tt{};
    bool tt_built = false;
    // grow-only scratch for the temporaries of the index builds (codes, orders, sort space): cudaMalloc / cudaFree
    // per build cost 4-140 ms of host time on a busy box, the kernels of a build 3-5 ms
    void* d_build_scratch = nullptr;
    size_t build_scratch_bytes = 0;
    long long bvh_sorted_cap = 0, bvh_nodes_cap = 0, tt_boxes_cap = 0, tt_tiles_cap = 0, q16_tiles_cap = 0, q16_etiles_cap = 0,
              filter_tiles_cap = 0;
    // second-generation walk (k_nn_tiletree_p): batch queue, winner tiles per slot, ticket tree of moment partials
    TTQueue* d_tt_queue = nullptr;
    int* d_win_tile = nullptr;
    long long win_tile_capacity = 0;
    bool tt_win_valid = false;             // d_win_tile matches d_keys (the last pass was a tile-tree walk over these slots)
    double* d_tt_part = nullptr;
    unsigned int* d_tt_tickets = nullptr;
    long long tt_tail_batches = 0;         // what d_tt_part / d_tt_tickets were sized for
    TTTail tt_tail{};
    int tt_grid = 0;                       // persistent grid of the walk (CTAs)
    // query slots in Morton order of the moving vertices (BVH mode) + warm start bookkeeping
    uint32_t* d_slot_pos = nullptr;
    uint32_t* d_pos_slot = nullptr;
    float4* d_mov_slots = nullptr;         // moving vertices gathered into slot order (fused kernel input)
    void* d_order_scratch = nullptr;
    size_t order_scratch_bytes = 0;
    long long order_capacity = 0;
    bool order_valid = false;
    int order_stride = 0;
    bool warm_valid = false;       // d_keys hold last iteration's winners for the current slots
    // brute-force warm start: the previous pass's winners (this rank's own, before the shard
    // resolution) seed the filter kernel's running minimum
    bool bf_warm_valid = false;
    bool bf_warm_active = false;   // the pass being enqueued was seeded
    long long bf_warm_nq = 0;
    unsigned long long* d_keys_local = nullptr;
    long long keys_local_capacity = 0;
    // moving
    float* d_mov = nullptr;
    long long mov_capacity = 0, n_vertices = 0;
    int* d_vlist = nullptr;
    long long vlist_capacity = 0, n_vlist = 0;
    bool has_vlist = false, have_moving = false, have_matrices = false;
    bool pose_dirty = true;               // imx1 / imx2 on the device are older than the matrices (k_prepare_pose needed)
    std::vector<int32_t> h_vlist;         // host copy of the uploaded vertex list (to recognise a repeated upload)
    // queries
    float* d_q = nullptr;                 // 3 planes of q_capacity
    unsigned long long* d_keys = nullptr;
    long long q_capacity = 0;
    // reductions
    double* d_partials = nullptr;
    double* d_partials2 = nullptr;         // second reduction stage (64 slices)
    int partials_capacity = 0;
    double* d_moments = nullptr;
    unsigned int* d_hist = nullptr;
    DevState* d_state = nullptr;
    icp_iter_stats* d_stats = nullptr;
    int stats_capacity = 0;
    // make_pairs scratch
    unsigned char* d_flags = nullptr;
    long long* d_block_counts = nullptr;
    long long* d_total = nullptr;
    double* d_pairs = nullptr;            // A | B | dist
    long long pairs_capacity = 0, flags_capacity = 0;
    // solve_pairs scratch
    double* d_v = nullptr;
    long long v_capacity = 0;
    // nccl
    ncclComm_t comm = nullptr;
    // measurement
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    bool profiling = false;
    std::vector<EventPair> ev_pool[5];
    size_t ev_used[5] = {0, 0, 0, 0, 0};
    double ev_total_ms[5] = {0, 0, 0, 0, 0};
    long long ev_launches[5] = {0, 0, 0, 0, 0};
    long long launches = 0;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    char err[512] = "";
};

namespace {

int fail(icp_handle* h, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(h, expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(h, _e == cudaErrorMemoryAllocation ? ICP_ERR_OOM : ICP_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                        __LINE__);                                                               \
    } while (0)

#define NCCL_TRY(h, expr)                                                                        \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess)                                                                   \
            return fail(h, ICP_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r));     \
    } while (0)

template <typename T>
int ensure(icp_handle* h, T*& ptr, long long& cap, long long need)
{
    if (need <= cap && ptr) return ICP_OK;
    if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
    long long n = std::max<long long>(need, 1);
    CUDA_TRY(h, cudaMalloc((void**)&ptr, (size_t)n * sizeof(T)));
    cap = n;
    return ICP_OK;
}

// carve-up of the build scratch: sizes first (need()), then one ensure, then take() in the same order
struct BuildArena {
    size_t total = 0, off = 0;
    char* base = nullptr;
    static size_t up(size_t v) { return (v + 255) / 256 * 256; }
    void need(size_t bytes) { total += up(bytes ? bytes : 1); }
    int commit(icp_handle* h)
    {
        if (total > h->build_scratch_bytes) {
            if (h->d_build_scratch) cudaFree(h->d_build_scratch);
            h->d_build_scratch = nullptr; h->build_scratch_bytes = 0;
            CUDA_TRY(h, cudaMalloc(&h->d_build_scratch, total));
            h->build_scratch_bytes = total;
        }
        base = (char*)h->d_build_scratch;
        off = 0;
        return ICP_OK;
    }
    template <typename T> T* take(size_t bytes) { T* p = (T*)(base + off); off += up(bytes ? bytes : 1); return p; }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember it per handle
// (a handle is bound to one device), not per process.
template <typename K>
int configure_smem(icp_handle* h, K kern, size_t smem)
{
    const void* f = reinterpret_cast<const void*>(kern);
    for (const void* p : h->smem_configured) if (p == f) return ICP_OK;
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->smem_configured.push_back(f);
    return ICP_OK;
}

// brute-force variants.  0 (default): query slots in Morton order; the unseeded first pass runs the
// expansion-filter kernel (or the direct kernel on small shards), every later pass the slab kernel
// (1-fma plane bound, then the 2-fma cylinder bound, then the exact metric) over Morton-ordered,
// per-tile-centred base tiles, seeded with the previous winners.
// 1..6 direct metric, 20..24 expansion filter, 30..36 cylinder filter on every pass:
// 30 = sorted base + sorted query slots, 31 = base in its own order, 32 = queries in their own
// order, 33 = neither, 34..36 = 30 with other blockings (8 queries/thread, 32-vertex chunks, both).
// 40..43 slab filter (1 fma per pair) in front of the cylinder filter: 40 = 4 queries/thread and 32-vertex
// chunks, 41 = 8 queries/thread, 42 = 64-vertex chunks, 43 = both.
bool cyl_variant(int v) { return v >= 30 && v <= 36; }
bool slab_variant(int v) { return v >= 40 && v <= 45; }   // 44, 45 = 40, 42 with Morton order inside a tile
bool cyl_sorts_base(const icp_handle* h)
{
    const int v = h->prm.nn_variant;
    return v == 0 || v >= 34 || !(v & 1);
}
bool uses_slots(const icp_handle* h)
{
    if (h->prm.nn_mode != ICP_NN_BRUTEFORCE) return true;
    const int v = h->prm.nn_variant;
    return v == 0 || v == 30 || v == 31 || (v >= 34 && v <= 36) || slab_variant(v) || (v >= 50 && v <= 79);
}

long long query_count(const icp_handle* h)
{
    const long long L = h->has_vlist ? h->n_vlist : h->n_vertices;
    const long long s = h->prm.stride > 1 ? h->prm.stride : 1;
    return (L + s - 1) / s;                       // len(vlist[0::s])
}

// ICP_TRACE_BUILD=1: wall-clock time of the one-time index builds to stderr (host side, includes the syncs)
struct BuildTrace {
    const char* what; icp_handle* h; bool on; std::chrono::steady_clock::time_point t0;
    BuildTrace(icp_handle* h_, const char* w) : what(w), h(h_) {
        static const bool enabled = getenv("ICP_TRACE_BUILD") != nullptr;
        on = enabled;
        if (on) t0 = std::chrono::steady_clock::now();
    }
    ~BuildTrace() {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 0.05) fprintf(stderr, "[icp build] %-24s %8.3f ms\n", what, ms);
    }
};

struct ProfScope {
    icp_handle* h; int which; EventPair* ep = nullptr;
    ProfScope(icp_handle* h_, int w) : h(h_), which(w)
    {
        if (!h->profiling) return;
        auto& pool = h->ev_pool[which];
        if (h->ev_used[which] == pool.size()) {
            EventPair p;
            cudaEventCreate(&p.a); cudaEventCreate(&p.b);
            pool.push_back(p);
        }
        ep = &pool[h->ev_used[which]++];
        cudaEventRecord(ep->a, h->stream);
    }
    ~ProfScope() { if (ep) cudaEventRecord(ep->b, h->stream); }
};

void collect_profile(icp_handle* h)
{
    for (int w = 0; w < 5; ++w) {
        for (size_t i = 0; i < h->ev_used[w]; ++i) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->ev_pool[w][i].a, h->ev_pool[w][i].b) == cudaSuccess) {
                h->ev_total_ms[w] += ms;
                h->ev_launches[w] += 1;
            }
        }
        h->ev_used[w] = 0;
    }
}

// ---- K2 launch ---------------------------------------------------------------------------
template <int R, int STAGES, bool PACKED>
int launch_nn_t(icp_handle* h, long long n_q)
{
    auto kern = k_nn_bruteforce<R, STAGES, PACKED>;
    const size_t smem = (size_t)STAGES * NN_TILE_BYTES + STAGES * sizeof(uint64_t);
    { int rc_ = configure_smem(h, kern, smem); if (rc_) return rc_; }
    const long long qb = (n_q + (long long)NN_THREADS * R - 1) / ((long long)NN_THREADS * R);
    // base ranges: enough CTAs for ~8 waves over the chip so the tail is small, but every CTA
    // still streams >= 8 tiles so the TMA ring and the final index recovery are amortised
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_THREADS, smem);
    if (occ < 1) occ = 1;
    const long long slots = (long long)h->sm_count * occ;
    long long want_y = (8 * slots + qb - 1) / qb;
    long long max_y = std::max<long long>(1, h->n_tiles / 8);
    long long gy = std::max<long long>(1, std::min(want_y, max_y));
    int tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy);
    gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    if (gy > 65535) { gy = 65535; tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy); gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta; }
    dim3 grid((unsigned)qb, (unsigned)gy);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    kern<<<grid, NN_THREADS, smem, h->stream>>>(h->d_base_tiles, h->n_tiles, tiles_per_cta, qx, qy, qz,
                                                n_q, h->d_keys, (uint32_t)h->base_off,
                                                &h->d_state->converged);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int build_filter(icp_handle* h)
{
    if (h->filter_built) return ICP_OK;
    BuildTrace bt_(h, "build_filter");
    if (!h->d_filter_info) CUDA_TRY(h, cudaMalloc((void**)&h->d_filter_info, sizeof(FilterInfo)));
    const long long n = h->n_local, n_pad = (long long)h->n_tiles * NN_TILE;
    if (n > 0) {
        const int pb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
        BuildArena ar;
        ar.need((size_t)pb * 24); ar.need(24);
        { int rc = ar.commit(h); if (rc) return rc; }
        float* part = ar.take<float>((size_t)pb * 24);
        float* bbox = ar.take<float>(24);
        { int rc = ensure(h, h->d_filter_tiles, h->filter_tiles_cap, n_pad * 4); if (rc) return rc; }
        k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_base4, n, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        k_filter_info<<<1, 32, 0, h->stream>>>(bbox, h->d_filter_info);
        k_build_filter_tiles<<<(unsigned)((n_pad + 255) / 256), 256, 0, h->stream>>>(h->d_base4, n, n_pad, h->d_filter_info, h->d_filter_tiles);
        h->launches += 4;
        CUDA_TRY(h, cudaGetLastError());
    }
    h->filter_built = true;
    return ICP_OK;
}

template <int R, int STAGES>
int launch_nn_filter_t(icp_handle* h, long long n_q)
{
    int rc = build_filter(h);
    if (rc) return rc;
    auto kern = k_nn_bruteforce_filter<R, STAGES>;
    const size_t smem = (size_t)STAGES * NNF_TILE_BYTES + STAGES * sizeof(uint64_t);
    { int rc_ = configure_smem(h, kern, smem); if (rc_) return rc_; }
    const long long qb = (n_q + (long long)NN_THREADS * R - 1) / ((long long)NN_THREADS * R);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_THREADS, smem);
    if (occ < 1) occ = 1;
    const long long slots = (long long)h->sm_count * occ;
    long long want_y = (8 * slots + qb - 1) / qb;
    long long max_y = std::max<long long>(1, h->n_tiles / 8);
    long long gy = std::max<long long>(1, std::min(want_y, max_y));
    int tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy);
    gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    if (gy > 65535) { gy = 65535; tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy); gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta; }
    dim3 grid((unsigned)qb, (unsigned)gy);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    kern<<<grid, NN_THREADS, smem, h->stream>>>(h->d_filter_tiles, h->d_base_tiles, h->d_filter_info, h->n_tiles,
                                                tiles_per_cta, qx, qy, qz, n_q, h->d_keys, (uint32_t)h->base_off,
                                                h->bf_warm_active ? 1 : 0, &h->d_state->converged, nullptr, nullptr, 0);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int build_cyl(icp_handle* h)
{
    const int want_sorted = cyl_sorts_base(h) ? 1 : 0;
    if (h->cyl_sorted == want_sorted) return ICP_OK;
    for (void* p : {(void*)h->d_cyl_ftiles, (void*)h->d_cyl_etiles, (void*)h->d_cyl_hdrs}) if (p) cudaFree(p);
    h->d_cyl_ftiles = h->d_cyl_etiles = nullptr; h->d_cyl_hdrs = nullptr;
    const long long n = h->n_local;
    if (n > 0) {
        unsigned long long *codes = nullptr, *codes_out = nullptr;
        uint32_t *order = nullptr, *order_out = nullptr;
        float *part = nullptr, *bbox = nullptr;
        void* temp = nullptr;
        size_t temp_bytes = 0;
        const int pb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
        CUDA_TRY(h, cudaMalloc((void**)&part, (size_t)pb * 24));
        CUDA_TRY(h, cudaMalloc((void**)&bbox, 24));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_cyl_ftiles, (size_t)h->n_tiles * CYL_TILE_BYTES));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_cyl_etiles, (size_t)h->n_tiles * CYL_TILE_BYTES));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_cyl_hdrs, (size_t)h->n_tiles * sizeof(CylTileHdr)));
        k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_base4, n, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        h->launches += 2;
        if (want_sorted) {
            CUDA_TRY(h, cudaMalloc((void**)&codes, (size_t)n * 8));
            CUDA_TRY(h, cudaMalloc((void**)&codes_out, (size_t)n * 8));
            CUDA_TRY(h, cudaMalloc((void**)&order, (size_t)n * 4));
            CUDA_TRY(h, cudaMalloc((void**)&order_out, (size_t)n * 4));
            k_morton_codes<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_base4, n, bbox, codes, order);
            CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
            CUDA_TRY(h, cudaMalloc(&temp, temp_bytes ? temp_bytes : 1));
            CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
            h->launches += 2;
        }
        // the dropped axis c is the one of smallest extent: the lower bound loses the least there
        float hb[6];
        CUDA_TRY(h, cudaMemcpyAsync(hb, bbox, 24, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        int c = 2;
        float ext[3];
        for (int k = 0; k < 3; ++k) { ext[k] = hb[3 + k] - hb[k]; if (!(ext[k] >= 0.0f)) ext[k] = 0.0f; }
        for (int k = 1; k >= 0; --k) if (ext[k] < ext[c]) c = k;
        const int a = (c + 1) % 3, b = (c + 2) % 3;
        h->cyl_axes = a | (b << 2) | (c << 4);
        k_build_cyl_tiles<<<(unsigned)h->n_tiles, 256, 0, h->stream>>>(h->d_base4, want_sorted ? order_out : nullptr, n, a, b, c,
                                                                      h->d_cyl_ftiles, h->d_cyl_etiles, h->d_cyl_hdrs);
        h->launches += 1;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (void* p : {(void*)codes, (void*)codes_out, (void*)order, (void*)order_out, (void*)part, (void*)bbox, temp}) if (p) cudaFree(p);
    }
    h->cyl_sorted = want_sorted;
    return ICP_OK;
}

template <int R, int STAGES, int CH>
int launch_nn_cyl_t(icp_handle* h, long long n_q)
{
    auto kern = k_nn_bruteforce_cyl<R, STAGES, CH>;
    const size_t smem = (size_t)STAGES * CYL_TILE_BYTES + STAGES * sizeof(uint64_t);
    { int rc_ = configure_smem(h, kern, smem); if (rc_) return rc_; }
    const long long qb = (n_q + (long long)NN_THREADS * R - 1) / ((long long)NN_THREADS * R);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_THREADS, smem);
    if (occ < 1) occ = 1;
    const long long slots = (long long)h->sm_count * occ;
    long long want_y = (8 * slots + qb - 1) / qb;
    long long max_y = std::max<long long>(1, h->n_tiles / 8);
    long long gy = std::max<long long>(1, std::min(want_y, max_y));
    int tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy);
    gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    if (gy > 65535) { gy = 65535; tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy); gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta; }
    dim3 grid((unsigned)qb, (unsigned)gy);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    kern<<<grid, NN_THREADS, smem, h->stream>>>(h->d_cyl_ftiles, h->d_cyl_etiles, h->d_cyl_hdrs, h->cyl_axes, h->n_tiles,
                                                tiles_per_cta, qx, qy, qz, n_q, h->d_keys, (uint32_t)h->base_off,
                                                h->bf_warm_active ? 1 : 0, &h->d_state->converged);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

// Slab tiles: Morton order decides which 1024 vertices share a tile, then each tile is sorted along
// the axis of largest extent (one more radix sort on (tile, coordinate) keys).  One-time per base.
int build_slab(icp_handle* h)
{
    const int want_order = (h->prm.nn_variant == 44 || h->prm.nn_variant == 45) ? 1 : 0;
    if (h->slab_order == want_order) return ICP_OK;
    for (void* p : {(void*)h->d_slab_ftiles, (void*)h->d_slab_etiles, (void*)h->d_slab_hdrs}) if (p) cudaFree(p);
    h->d_slab_ftiles = h->d_slab_etiles = nullptr; h->d_slab_hdrs = nullptr;
    const long long n = h->n_local;
    if (n > 0) {
        unsigned long long *codes = nullptr, *codes_out = nullptr;
        uint32_t *order = nullptr, *order_out = nullptr;
        float *part = nullptr, *bbox = nullptr;
        void* temp = nullptr;
        size_t temp_bytes = 0;
        const int pb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
        const unsigned gn = (unsigned)((n + 255) / 256);
        CUDA_TRY(h, cudaMalloc((void**)&part, (size_t)pb * 24));
        CUDA_TRY(h, cudaMalloc((void**)&bbox, 24));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_slab_ftiles, (size_t)h->n_tiles * SLAB_TILE_BYTES));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_slab_etiles, (size_t)h->n_tiles * CYL_TILE_BYTES));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_slab_hdrs, (size_t)h->n_tiles * sizeof(CylTileHdr)));
        CUDA_TRY(h, cudaMalloc((void**)&codes, (size_t)n * 8));
        CUDA_TRY(h, cudaMalloc((void**)&codes_out, (size_t)n * 8));
        CUDA_TRY(h, cudaMalloc((void**)&order, (size_t)n * 4));
        CUDA_TRY(h, cudaMalloc((void**)&order_out, (size_t)n * 4));
        k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_base4, n, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        k_morton_codes<<<gn, 256, 0, h->stream>>>(h->d_base4, n, bbox, codes, order);
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 64, h->stream));
        CUDA_TRY(h, cudaMalloc(&temp, temp_bytes ? temp_bytes : 1));
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
        // a = largest extent (the slab axis), c = smallest (dropped by the cylinder bound), b = the third
        float hb[6];
        CUDA_TRY(h, cudaMemcpyAsync(hb, bbox, 24, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ext[3];
        for (int k = 0; k < 3; ++k) { ext[k] = hb[3 + k] - hb[k]; if (!(ext[k] >= 0.0f)) ext[k] = 0.0f; }
        int a = 0, c = 2;
        for (int k = 1; k < 3; ++k) if (ext[k] > ext[a]) a = k;
        for (int k = 1; k >= 0; --k) if (ext[k] < ext[c]) c = k;
        if (c == a) c = (a + 2) % 3;
        const int b = 3 - a - c;
        h->slab_axes = a | (b << 2) | (c << 4);
        if (want_order == 0) {
            k_slab_keys<<<gn, 256, 0, h->stream>>>(h->d_base4, order_out, n, a, codes);
            CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order_out, order, (int)n, 0, 64, h->stream));
        }
        k_build_slab_tiles<<<(unsigned)h->n_tiles, 256, 0, h->stream>>>(h->d_base4, want_order == 0 ? order : order_out, n, a, b, c,
                                                                       h->d_slab_ftiles, h->d_slab_etiles, h->d_slab_hdrs);
        h->launches += 7;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (void* p : {(void*)codes, (void*)codes_out, (void*)order, (void*)order_out, (void*)part, (void*)bbox, temp}) if (p) cudaFree(p);
    }
    h->slab_order = want_order;
    return ICP_OK;
}

template <int R, int STAGES, int CH>
int launch_nn_slab_t(icp_handle* h, long long n_q)
{
    auto kern = k_nn_bruteforce_slab<R, STAGES, CH>;
    const size_t smem = (size_t)STAGES * SLAB_TILE_BYTES + 2 * STAGES * sizeof(uint64_t);
    { int rc_ = configure_smem(h, kern, smem); if (rc_) return rc_; }
    const long long qb = (n_q + (long long)NN_THREADS * R - 1) / ((long long)NN_THREADS * R);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_THREADS, smem);
    if (occ < 1) occ = 1;
    const long long slots = (long long)h->sm_count * occ;
    long long want_y = (8 * slots + qb - 1) / qb;
    long long max_y = std::max<long long>(1, h->n_tiles / 8);
    long long gy = std::max<long long>(1, std::min(want_y, max_y));
    int tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy);
    gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    if (gy > 65535) { gy = 65535; tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy); gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta; }
    dim3 grid((unsigned)qb, (unsigned)gy);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    kern<<<grid, NN_THREADS, smem, h->stream>>>(h->d_slab_ftiles, h->d_slab_etiles, h->d_slab_hdrs, h->slab_axes, h->n_tiles,
                                                tiles_per_cta, qx, qy, qz, n_q, h->d_keys, (uint32_t)h->base_off,
                                                h->bf_warm_active ? 1 : 0, &h->d_state->converged);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

// q16 tiles: the slab kernel's vertex order (Morton order decides the tile, then a sort along the axis of
// largest extent inside the tile -- or Morton order inside the tile too), quantised to 16 bits per
// tile.  One-time per base.
int build_q16(icp_handle* h, int want_order)
{
    if (h->q16_order == want_order) return ICP_OK;
    BuildTrace bt_(h, "build_q16");
    const long long n = h->n_local;
    if (n > 0) {
        size_t temp_bytes = 0;
        const int pb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
        const unsigned gn = (unsigned)((n + 255) / 256);
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64, h->stream));
        BuildArena ar;
        ar.need((size_t)pb * 24); ar.need(24); ar.need((size_t)n * 8); ar.need((size_t)n * 8); ar.need((size_t)n * 4);
        ar.need((size_t)n * 4); ar.need(temp_bytes);
        { int rc = ar.commit(h); if (rc) return rc; }
        float* part = ar.take<float>((size_t)pb * 24);
        float* bbox = ar.take<float>(24);
        unsigned long long* codes = ar.take<unsigned long long>((size_t)n * 8);
        unsigned long long* codes_out = ar.take<unsigned long long>((size_t)n * 8);
        uint32_t* order = ar.take<uint32_t>((size_t)n * 4);
        uint32_t* order_out = ar.take<uint32_t>((size_t)n * 4);
        void* temp = ar.take<char>(temp_bytes);
        { int rc = ensure(h, h->d_q16_tiles, h->q16_tiles_cap, (long long)h->n_tiles * Q16_TILE_WORDS); if (rc) return rc; }
        { int rc = ensure(h, h->d_q16_etiles, h->q16_etiles_cap, (long long)h->n_tiles * CYL_TILE_FLOATS); if (rc) return rc; }
        k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_base4, n, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        // the 1024-vertex tiles of the brute-force kernels follow the MORTON curve: its ranges of 2^10 positions are
        // axis-aligned cells, whose boxes the slab and cull tests prune best (ICP_Q16_TILE_CURVE=hilbert for the A/B)
        static const bool q16_hilbert = getenv("ICP_Q16_TILE_CURVE") && !strcmp(getenv("ICP_Q16_TILE_CURVE"), "hilbert");
        k_morton_codes<<<gn, 256, 0, h->stream>>>(h->d_base4, n, bbox, codes, order, q16_hilbert ? 0 : 1);
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
        // a, b = the two longest axes of the shard (the candidate slab axes), c = the shortest
        float hb[6];
        CUDA_TRY(h, cudaMemcpyAsync(hb, bbox, 24, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ext[3];
        for (int k = 0; k < 3; ++k) { ext[k] = hb[3 + k] - hb[k]; if (!(ext[k] >= 0.0f)) ext[k] = 0.0f; }
        int a = 0, c = 2;
        for (int k = 1; k < 3; ++k) if (ext[k] > ext[a]) a = k;
        for (int k = 1; k >= 0; --k) if (ext[k] < ext[c]) c = k;
        if (c == a) c = (a + 2) % 3;
        const int b = 3 - a - c;
        h->q16_axes = a | (b << 2) | (c << 4);
        if (want_order == 0) {
            k_slab_keys<<<gn, 256, 0, h->stream>>>(h->d_base4, order_out, n, a, codes);
            CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order_out, order, (int)n, 0, 64, h->stream));
        }
        k_build_q16_tiles<<<(unsigned)h->n_tiles, 256, 0, h->stream>>>(h->d_base4, want_order == 0 ? order : order_out, n, a, b, c,
                                                                      h->d_q16_tiles, h->d_q16_etiles);
        h->launches += 7;
        CUDA_TRY(h, cudaGetLastError());
    }
    h->q16_order = want_order;
    return ICP_OK;
}

// The class machinery costs ~85 us per pass (one small kernel and a three-kernel radix sort), which a pass over
// fewer than 2^18 slots does not earn back (C2, 100k x 100k: 0.73 -> 0.81 ms per iteration with it).
// ICP_Q16_CLASS_MIN overrides the threshold (the tests run the class path on small clouds with it).
bool q16_classes_wanted(const icp_handle* h, long long n_q)
{
    if (h->prm.nn_variant == 77) return false;
    const char* e = getenv("ICP_Q16_CLASS_MIN");
    const long long min_slots = e ? atoll(e) : (1ll << 18);
    return n_q >= min_slots;
}

// buffers of the q16 class machinery for n_q slots
int ensure_q16_classes(icp_handle* h, long long n_q)
{
    size_t temp_bytes = 0;
    CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned char*)nullptr, (unsigned char*)nullptr,
                                                (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_q, 0, Q16_CLASS_BITS, h->stream));
    if (!h->d_q16_hist) {
        CUDA_TRY(h, cudaMalloc((void**)&h->d_q16_hist, Q16_CLS_WORDS * sizeof(unsigned int)));
        CUDA_TRY(h, cudaMemsetAsync(h->d_q16_hist, 0, Q16_CLS_WORDS * sizeof(unsigned int), h->stream));
    }
    if (n_q > h->q16_class_capacity || temp_bytes > h->q16_sort_temp_bytes) {
        for (void* p : {(void*)h->d_q16_cls, (void*)h->d_q16_ids, (void*)h->d_q16_sort_temp}) if (p) cudaFree(p);
        h->d_q16_cls = nullptr; h->d_q16_ids = nullptr; h->d_q16_sort_temp = nullptr;
        h->q16_class_capacity = 0; h->q16_sort_temp_bytes = 0;
        const long long cap = std::max<long long>(n_q, 1);
        CUDA_TRY(h, cudaMalloc((void**)&h->d_q16_cls, (size_t)cap * 2));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_q16_ids, (size_t)cap * 2 * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMalloc(&h->d_q16_sort_temp, temp_bytes ? temp_bytes : 1));
        h->q16_class_capacity = cap; h->q16_sort_temp_bytes = temp_bytes;
    }
    return ICP_OK;
}

template <int R, int STAGES, int CH, int MIX, int MINB = 1>
int launch_nn_q16_t(icp_handle* h, long long n_q)
{
    auto kern = k_nn_bruteforce_q16<R, STAGES, CH, MIX, MINB, false>;
    auto kern_gated = k_nn_bruteforce_q16<R, STAGES, CH, MIX, MINB, true>;
    const size_t smem = (size_t)STAGES * ((Q16_TILE_BYTES + 127) / 128 * 128) + 2 * STAGES * sizeof(uint64_t);
    { int rc_ = configure_smem(h, kern, smem); if (rc_) return rc_; }
    { int rc_ = configure_smem(h, kern_gated, smem); if (rc_) return rc_; }
    const long long qb = (n_q + (long long)NN_THREADS * R - 1) / ((long long)NN_THREADS * R);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_THREADS, smem);
    if (occ < 1) occ = 1;
    const long long slots = (long long)h->sm_count * occ;
    // ~24 waves (a CTA's duration varies with its slow-path work: ncu showed 13 % of the SM time idle in the
    // tail at 8.8 waves), every CTA still streams >= 4 tiles
    long long want_y = (24 * slots + qb - 1) / qb;
    long long max_y = std::max<long long>(1, h->n_tiles / 4);
    long long gy = std::max<long long>(1, std::min(want_y, max_y));
    int tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy);
    gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    if (gy > 65535) { gy = 65535; tiles_per_cta = (int)((h->n_tiles + gy - 1) / gy); gy = (h->n_tiles + tiles_per_cta - 1) / tiles_per_cta; }
    dim3 grid((unsigned)qb, (unsigned)gy);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const float* bound = h->bound_active ? h->d_bound : nullptr;
    // slots in class order (nn_variant 77: identity, the A/B of this step)
    const uint32_t* perm = nullptr;
    if (q16_classes_wanted(h, n_q)) {
        { int rc_ = ensure_q16_classes(h, n_q); if (rc_) return rc_; }
        const int* stop = &h->d_state->converged;
        if (!h->q16_hist_by_k1) {          // the seeds of this pass were not made by K1 (cold pass): histogram here
            const int hg = (int)std::min<long long>((n_q + 255) / 256, (long long)h->sm_count * 8);
            k_q16_class_hist<<<hg, 256, 0, h->stream>>>(h->d_keys, bound, n_q, h->d_q16_hist, h->q16_pass, stop);
            h->launches++;
        }
        k_q16_class_keys<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(h->d_keys, bound, n_q, h->d_q16_hist, h->q16_pass,
                                                                              h->d_q16_cls, h->d_q16_ids, stop);
        size_t temp_bytes = h->q16_sort_temp_bytes;
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(h->d_q16_sort_temp, temp_bytes, (const unsigned char*)h->d_q16_cls,
                                                    h->d_q16_cls + h->q16_class_capacity, (const uint32_t*)h->d_q16_ids,
                                                    h->d_q16_ids + h->q16_class_capacity, (int)n_q, 0, Q16_CLASS_BITS, h->stream));
        perm = h->d_q16_ids + h->q16_class_capacity;
        h->launches += 4;
        h->q16_pass++;
        h->q16_hist_by_k1 = false;
    }
    // classes >= hard_class go to the expansion-filter kernel (flat cost per pair, no pruning to lose)
    const char* hard_e = getenv("ICP_Q16_HARD_CLASS");             // tuning / tests; 16 = every class in the q16 kernel
    const int hard_env = hard_e ? atoi(hard_e) : Q16_HARD_CLASS;
    const int hard_class = (perm != nullptr && h->prm.nn_variant != 78) ? hard_env : (1 << Q16_CLASS_BITS);
    const bool gated = perm != nullptr && hard_class < (1 << Q16_CLASS_BITS);
    (gated ? kern_gated : kern)<<<grid, NN_THREADS, smem, h->stream>>>(
        h->d_q16_tiles, h->d_q16_etiles, h->q16_axes, h->n_tiles, tiles_per_cta, qx, qy, qz, n_q, h->d_keys, bound,
        (uint32_t)h->base_off, &h->d_state->converged, perm, h->d_q16_hist, hard_class);
    h->launches++;
    if (gated) {
        auto fk = k_nn_bruteforce_filter<4, 3, true>;
        const size_t fsmem = (size_t)3 * NNF_TILE_BYTES + 3 * sizeof(uint64_t);
        { int rc_ = configure_smem(h, fk, fsmem); if (rc_) return rc_; }
        const long long fqb = (n_q + (long long)NN_THREADS * 4 - 1) / ((long long)NN_THREADS * 4);
        int focc = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&focc, fk, NN_THREADS, fsmem);
        const long long fslots = (long long)h->sm_count * std::max(1, focc);
        // how many hard blocks there are is only known on the device (see the kernel): split the base finely
        (void)fslots;
        long long fgy = std::max<long long>(1, std::min<long long>(32, h->n_tiles / 4));
        int ftiles = (int)((h->n_tiles + fgy - 1) / fgy);
        fgy = (h->n_tiles + ftiles - 1) / ftiles;
        fk<<<dim3((unsigned)fqb, (unsigned)fgy), NN_THREADS, fsmem, h->stream>>>(
            h->d_filter_tiles, h->d_base_tiles, h->d_filter_info, h->n_tiles, ftiles, qx, qy, qz, n_q, h->d_keys,
            (uint32_t)h->base_off, 1, &h->d_state->converged, perm, h->d_q16_hist, hard_class);
        h->launches++;
    }
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

bool q16_variant(int v) { return v >= 50 && v <= 79; }   // 58 = 62 without the global pruning bound (A/B)

bool q16_classes_wanted(const icp_handle* h, long long n_q);
int build_tiletree(icp_handle* h);
int ensure_tiletree_walk(icp_handle* h, long long n_q, bool fused, long long entries = 0);
template <bool FUSED> int launch_tiletree_walk(icp_handle* h, long long n_q, int solve);

int launch_nn(icp_handle* h, long long n_q)
{
    if (n_q <= 0 || h->n_tiles <= 0) return ICP_OK;
    int variant = h->prm.nn_variant;
    // The q16 kernel needs a seed: any real vertex of the shard with its exact S4 distance -- the exhaustive
    // pass that follows is exact whatever the seed's quality, a good one only makes it prune sooner.  First pass
    // of the default path: the tile-tree search (the exact nearest vertex, 0.3-0.5 ms at C3).
    // nn_variant 79: greedy descent of the tile tree instead (the nearest vertex of the nearest Morton tile);
    // nn_variant 59: first pass by the unseeded expansion-filter kernel (round-1 behaviour).
    if ((variant == 0 || (q16_variant(variant) && variant != 59)) && !h->bf_warm_active && h->n_tiles >= 16) {
        int rc = build_tiletree(h);
        if (rc) return rc;
        if (variant != 79 && uses_slots(h)) {
            // the exact tree search as the seed (0.3-0.5 ms at C3): the exhaustive pass that follows then prunes
            // with the true nearest distance from its first tile on (first iteration 126 -> ~30 ms at C3)
            rc = ensure_tiletree_walk(h, n_q, false);
            if (rc) return rc;
            ProfScope ps(h, 3);
            rc = launch_tiletree_walk<false>(h, n_q, 0);
            if (rc) return rc;
        } else {       // nn_variant 79: greedy descent only (round-2 first half)
            ProfScope ps(h, 3);
            float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
            k_tt_seed<<<(unsigned)((n_q + 127) / 128), 128, 0, h->stream>>>(h->d_tt_desc, qx, qy, qz, n_q, h->d_keys,
                                                                            (uint32_t)h->base_off, &h->d_state->converged);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
        h->bf_warm_active = true;
    }
    if (q16_variant(variant) && !h->bf_warm_active) variant = 0;       // no seed (tiny shard / variant 59): cold pass below
    if (variant == 59 || variant == 79) variant = 71;
    if (variant == 0) {
        if (h->bf_warm_active && h->n_tiles >= 16) variant = 71;       // seeded: 16-bit slab kernel
        else variant = h->n_tiles >= 256 ? 20 : 2;                     // cold: filter / direct kernel
    }
    if (q16_variant(variant)) {
        int rc = build_q16(h, (variant == 54 || variant == 55 || variant == 66) ? 1 : 0);
        if (rc) return rc;
        if (variant != 78 && q16_classes_wanted(h, n_q)) {       // the hard classes of the slots are served by the expansion-filter kernel
            rc = build_filter(h);
            if (rc) return rc;
        }
    }
    if (variant >= 20 && variant <= 29) {
        int rc = build_filter(h);      // outside the profiled region (one-time)
        if (rc) return rc;
    }
    if (cyl_variant(variant)) {
        int rc = build_cyl(h);
        if (rc) return rc;
    }
    if (slab_variant(variant)) {
        int rc = build_slab(h);
        if (rc) return rc;
    }
    ProfScope ps(h, 0);
    switch (variant) {
        // q16, alu-only loop (VIADDMNMX): 50 = 4 queries/thread, 64-vertex chunks; 51 = 8 queries; 52, 53 = 128-vertex
        // chunks; 54, 55 = 50, 51 with Morton order inside a tile; 56 = 2 queries; 57 = 32-vertex chunks
        case 50: case 54: return launch_nn_q16_t<4, 4, 64, 0, 3>(h, n_q);
        case 51: case 55: return launch_nn_q16_t<8, 4, 64, 0>(h, n_q);
        case 52: return launch_nn_q16_t<4, 4, 128, 0, 3>(h, n_q);
        case 53: return launch_nn_q16_t<8, 4, 128, 0>(h, n_q);
        case 56: return launch_nn_q16_t<2, 4, 64, 0, 3>(h, n_q);
        case 57: return launch_nn_q16_t<4, 4, 32, 0, 3>(h, n_q);
        // q16, split loops (packed adds as IMAD on the fma pipe + VIMNMX3 on the alu pipe), 4 queries/thread,
        // <= 80 registers (3 CTAs per SM): 60 = ptxas' own IMAD/IADD3 balance, 61 = 3 IMAD + 1 VIADD.16x2 + 2
        // VIMNMX3, 62 = 2 VIADDMNMX + 2 IMAD + 1 VIMNMX3, 63 = 4 IMAD + 2 VIMNMX3 (per 8 vertices), 128-vertex
        // chunks; 64..67 = the same with 64-vertex chunks; 66 Morton order inside a tile (with 62's loop);
        // 68, 69 = 62 with 8 / 2 queries per thread; 58 = 62 without the global pruning bound
        case 60: return launch_nn_q16_t<4, 4, 128, 1, 3>(h, n_q);
        case 61: return launch_nn_q16_t<4, 4, 128, 2, 3>(h, n_q);
        case 62: case 66: case 58: return launch_nn_q16_t<4, 4, 128, 3, 3>(h, n_q);
        case 63: return launch_nn_q16_t<4, 4, 128, 4, 3>(h, n_q);
        case 64: return launch_nn_q16_t<4, 4, 64, 1, 3>(h, n_q);
        case 65: return launch_nn_q16_t<4, 4, 64, 2, 3>(h, n_q);
        case 67: return launch_nn_q16_t<4, 4, 64, 3, 3>(h, n_q);
        case 70: return launch_nn_q16_t<4, 4, 128, 5, 3>(h, n_q);      // one third VIADDMNMX, two thirds IMAD + VIMNMX3
        case 71: return launch_nn_q16_t<4, 8, 128, 3, 3>(h, n_q);      // 62 with an 8-tile ring
        case 72: return launch_nn_q16_t<4, 8, 128, 5, 3>(h, n_q);      // 70 with an 8-tile ring
        case 73: return launch_nn_q16_t<4, 8, 256, 5, 3>(h, n_q);      // 72 with 256-vertex chunks
        case 77: return launch_nn_q16_t<4, 8, 128, 3, 3>(h, n_q);      // 71 without the class order of the slots (A/B)
        case 78: return launch_nn_q16_t<4, 8, 128, 3, 3>(h, n_q);      // 71 with the class order, every class in the q16 kernel (A/B)
        case 68: return launch_nn_q16_t<8, 4, 128, 3>(h, n_q);
        case 69: return launch_nn_q16_t<2, 4, 128, 3, 3>(h, n_q);
        case 40: return launch_nn_slab_t<4, 3, 32>(h, n_q);
        case 41: return launch_nn_slab_t<8, 3, 32>(h, n_q);
        case 42: return launch_nn_slab_t<4, 3, 64>(h, n_q);
        case 43: return launch_nn_slab_t<8, 3, 64>(h, n_q);
        case 44: return launch_nn_slab_t<4, 3, 32>(h, n_q);
        case 45: return launch_nn_slab_t<4, 3, 64>(h, n_q);
        case 30: case 31: case 32: case 33: return launch_nn_cyl_t<4, 3, 16>(h, n_q);
        case 34: return launch_nn_cyl_t<8, 3, 16>(h, n_q);
        case 35: return launch_nn_cyl_t<4, 3, 32>(h, n_q);
        case 36: return launch_nn_cyl_t<8, 3, 32>(h, n_q);
        case 20: return launch_nn_filter_t<4, 3>(h, n_q);   // expansion filter + exact re-check
        case 21: return launch_nn_filter_t<8, 3>(h, n_q);
        case 22: return launch_nn_filter_t<4, 4>(h, n_q);
        case 23: return launch_nn_filter_t<6, 3>(h, n_q);
        case 24: return launch_nn_filter_t<2, 3>(h, n_q);
        case 1: return launch_nn_t<4, 4, false>(h, n_q);   // scalar pipe, 4 queries / thread
        case 2: return launch_nn_t<4, 4, true>(h, n_q);    // packed f32x2
        case 3: return launch_nn_t<8, 4, true>(h, n_q);
        case 4: return launch_nn_t<2, 4, true>(h, n_q);
        case 5: return launch_nn_t<8, 4, false>(h, n_q);
        case 6: return launch_nn_t<6, 4, true>(h, n_q);
        default: return launch_nn_t<4, 4, true>(h, n_q);
    }
}

// ---- spatial index build (one-time per base, the BVHTree.FromObject equivalent) -----------------
int build_bvh(icp_handle* h)
{
    if (h->bvh_built) return ICP_OK;
    BuildTrace bt_(h, "build_bvh");
    const long long n = h->n_local;
    h->bvh_P = 0;
    if (n <= 0) { h->bvh_built = true; return ICP_OK; }
    const int n_runs = (int)((n + BVH_LEAF - 1) / BVH_LEAF);
    const long long n_pad = (long long)n_runs * BVH_LEAF;
    size_t temp_bytes = 0;
    const int pb = std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
    CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 63, h->stream));
    BuildArena ar;
    ar.need((size_t)n * 8); ar.need((size_t)n * 8); ar.need((size_t)n * 4); ar.need((size_t)n * 4);
    ar.need((size_t)n_runs * 8); ar.need((size_t)n_runs * 4); ar.need((size_t)pb * 24); ar.need(24); ar.need(temp_bytes);
    { int rc = ar.commit(h); if (rc) return rc; }
    unsigned long long* codes = ar.take<unsigned long long>((size_t)n * 8);
    unsigned long long* codes_out = ar.take<unsigned long long>((size_t)n * 8);
    uint32_t* order = ar.take<uint32_t>((size_t)n * 4);
    uint32_t* order_out = ar.take<uint32_t>((size_t)n * 4);
    unsigned long long* run_keys = ar.take<unsigned long long>((size_t)n_runs * 8);
    int* leaf_parent = ar.take<int>((size_t)n_runs * 4);
    float* part = ar.take<float>((size_t)pb * 24);
    float* bbox = ar.take<float>(24);
    void* temp = ar.take<char>(temp_bytes);
    { int rc = ensure(h, h->d_bvh_sorted, h->bvh_sorted_cap, n_pad); if (rc) return rc; }
    { int rc = ensure(h, h->d_bvh_nodes, h->bvh_nodes_cap, (long long)std::max(1, n_runs - 1)); if (rc) return rc; }
    k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_base4, n, part);
    k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
    k_morton_codes<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_base4, n, bbox, codes, order);
    CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
    k_bvh_gather<<<(unsigned)((n_pad + 255) / 256), 256, 0, h->stream>>>(h->d_base4, order_out, n, n_pad, h->d_bvh_sorted);
    k_run_keys<<<(unsigned)((n_runs + 255) / 256), 256, 0, h->stream>>>(codes_out, n, n_runs, run_keys);
    if (n_runs > 1) {
        k_lbvh_topology<<<(unsigned)((n_runs + 127) / 128), 128, 0, h->stream>>>(run_keys, n_runs, h->d_bvh_nodes, leaf_parent);
        k_lbvh_fit<<<(unsigned)((n_runs + 127) / 128), 128, 0, h->stream>>>(h->d_bvh_sorted, n, n_runs, h->d_bvh_nodes, leaf_parent);
    }
    h->launches += 8;
    CUDA_TRY(h, cudaGetLastError());
    const int P = n_runs;
    h->bvh_P = P;
    h->bvh_built = true;
    return ICP_OK;
}

// tile tree over the Morton-sorted vertices build_bvh leaves in d_bvh_sorted (one-time per base)
int build_tiletree(icp_handle* h)
{
    if (h->tt_built) return ICP_OK;
    BuildTrace bt_(h, "build_tiletree(+bvh)");
    int rc = build_bvh(h);
    if (rc) return rc;
    memset(&h->tt, 0, sizeof(h->tt));
    const long long n = h->n_local;
    if (n > 0) {
        int total = 0, c = (int)((n + TT_LEAF - 1) / TT_LEAF), levels = 0;
        while (true) {
            h->tt.count[levels] = c; h->tt.offset[levels] = total; total += c; ++levels;
            if (c <= TT_FAN || levels == TT_MAX_LEVELS) break;
            c = (c + TT_FAN - 1) / TT_FAN;
        }
        if (h->tt.count[levels - 1] > TT_FAN) return fail(h, ICP_ERR_BAD_ARG, "tile tree: too many vertices");
        h->tt.n_levels = levels;
        { int rc2 = ensure(h, h->d_tt_boxes, h->tt_boxes_cap, (long long)total * 2); if (rc2) return rc2; }
        { int rc2 = ensure(h, h->d_tt_tiles, h->tt_tiles_cap, (long long)h->tt.count[0] * TT_TILE_FLOATS); if (rc2) return rc2; }
        k_tt_build_tiles<<<(unsigned)(((long long)h->tt.count[0] * 32 + 255) / 256), 256, 0, h->stream>>>(
            h->d_bvh_sorted, n, h->tt.count[0], h->d_tt_tiles, h->d_tt_boxes);
        for (int l = 1; l < levels; ++l)
            k_tt_build_level<<<(unsigned)(((long long)h->tt.count[l] * 32 + 255) / 256), 256, 0, h->stream>>>(
                h->d_tt_boxes + 2 * (size_t)h->tt.offset[l - 1], h->tt.count[l - 1],
                h->d_tt_boxes + 2 * (size_t)h->tt.offset[l], h->tt.count[l]);
        h->launches += levels;
        CUDA_TRY(h, cudaGetLastError());
    }
    h->tt.boxes = h->d_tt_boxes;
    h->tt.tiles = h->d_tt_tiles;
    if (!h->d_tt_desc) CUDA_TRY(h, cudaMalloc((void**)&h->d_tt_desc, sizeof(TileTree)));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_tt_desc, &h->tt, sizeof(TileTree), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->tt_built = true;
    return ICP_OK;
}

// vertex tree searches: 0 (default) = warp walk over the 32-ary tile tree; 12 = binary LBVH packet walk;
// 10 = LBVH per-thread walk; 98, 99 = LBVH walks with visit counters
bool tiletree_variant(const icp_handle* h) { return h->prm.nn_variant == 0 || h->prm.nn_variant == 13; }
bool tiletree_persistent(const icp_handle* h) { return h->prm.nn_variant == 0; }   // 13 = first generation (one CTA per 128 slots)

// queue, winner tiles and (fused) the ticket tree for n_q query slots; the persistent grid
// `entries` (0 = one per batch of 32 slots): level-0 entries of the ticket tree (the surface walk: one per warp of its grid)
int ensure_tiletree_walk(icp_handle* h, long long n_q, bool fused, long long entries)
{
    if (!h->d_tt_queue) {
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tt_queue, sizeof(TTQueue)));
        CUDA_TRY(h, cudaMemsetAsync(h->d_tt_queue, 0, sizeof(TTQueue), h->stream));
    }
    if (n_q > h->win_tile_capacity || !h->d_win_tile) {
        if (h->d_win_tile) cudaFree(h->d_win_tile);
        h->d_win_tile = nullptr; h->win_tile_capacity = 0; h->tt_win_valid = false;
        const long long cap = std::max<long long>(n_q, 1);
        CUDA_TRY(h, cudaMalloc((void**)&h->d_win_tile, (size_t)cap * sizeof(int)));
        h->win_tile_capacity = cap;
    }
    const long long nb = entries > 0 ? entries : (n_q + 31) / 32;
    if (fused && (nb != h->tt_tail_batches || !h->d_tt_part)) {
        TTTail& t = h->tt_tail;
        memset(&t, 0, sizeof(t));
        long long part = 0, tickets = 0, c = nb;
        int lv = 0;
        for (; lv < TT_TAIL_LEVELS; ++lv) {
            const long long groups = (c + 31) / 32;
            t.count[lv] = (int)c; t.part_off[lv] = part; t.ticket_off[lv] = (int)tickets;
            part += groups * 32 * MOM; tickets += groups;
            if (c <= 32) { ++lv; break; }
            c = groups;
        }
        if (nb > 0 && t.count[lv - 1] > 32) return fail(h, ICP_ERR_BAD_ARG, "tile tree: too many query slots");
        if (h->d_tt_part) cudaFree(h->d_tt_part);
        if (h->d_tt_tickets) cudaFree(h->d_tt_tickets);
        h->d_tt_part = nullptr; h->d_tt_tickets = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tt_part, (size_t)std::max<long long>(part, 1) * sizeof(double)));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tt_tickets, (size_t)std::max<long long>(tickets, 1) * sizeof(unsigned int)));
        CUDA_TRY(h, cudaMemsetAsync(h->d_tt_tickets, 0, (size_t)std::max<long long>(tickets, 1) * sizeof(unsigned int), h->stream));
        h->tt_tail_batches = nb;
    }
    if (h->tt_grid == 0) {
        int occ_f = 1, occ_p = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_nn_tiletree_p<true>, TT_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, k_nn_tiletree_p<false>, TT_THREADS, 0);
        h->tt_grid = h->sm_count * std::max(1, std::min(occ_f, occ_p));
    }
    return ICP_OK;
}

TTTail make_tail(const icp_handle* h, int solve)
{
    TTTail tail = h->tt_tail;
    tail.part = h->d_tt_part; tail.tickets = h->d_tt_tickets; tail.moments = h->d_moments;
    tail.stats = h->d_stats; tail.stats_capacity = h->stats_capacity;
    tail.solve = solve; tail.scale_mode = h->prm.align_meth == 1; tail.use_target = h->prm.use_target; tail.target_d = h->prm.target_d;
    return tail;
}

template <bool FUSED>
int launch_tiletree_walk(icp_handle* h, long long n_q, int solve)
{
    const long long nb = (n_q + 31) / 32;
    const int g = (int)std::max<long long>(1, std::min<long long>((nb + TT_THREADS / 32 - 1) / (TT_THREADS / 32), h->tt_grid));
    const TTTail tail = make_tail(h, solve);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const int warm = (h->warm_valid ? 1 : 0) | ((h->warm_valid && h->tt_win_valid) ? 2 : 0);
    k_nn_tiletree_p<FUSED><<<g, TT_THREADS, 0, h->stream>>>(
        h->tt, FUSED ? nullptr : qx, FUSED ? nullptr : qy, FUSED ? nullptr : qz, h->d_state,
        FUSED ? h->d_mov_slots : nullptr, FUSED ? h->d_base4 : nullptr, h->base_off, h->n_local, h->prm.thresh, warm, n_q,
        h->d_keys, h->d_win_tile, (uint32_t)h->base_off, h->d_tt_queue, tail);
    h->launches++;
    h->tt_win_valid = true;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int launch_nn_bvh(icp_handle* h, long long n_q)
{
    if (n_q <= 0 || h->n_local <= 0) return ICP_OK;
    int rc = tiletree_variant(h) ? build_tiletree(h) : build_bvh(h);
    if (rc) return rc;
    if (tiletree_persistent(h)) {
        rc = ensure_tiletree_walk(h, n_q, false);
        if (rc) return rc;
    }
    ProfScope ps(h, 0);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const unsigned g128 = (unsigned)((n_q + 127) / 128);
    const unsigned gpk = (unsigned)((n_q + PKT_THREADS - 1) / PKT_THREADS);
    const int v = h->prm.nn_variant;
    if (tiletree_persistent(h)) return launch_tiletree_walk<false>(h, n_q, 0);
    h->tt_win_valid = false;
    if (tiletree_variant(h)) {
        k_nn_tiletree<false><<<(unsigned)((n_q + TT_THREADS - 1) / TT_THREADS), TT_THREADS, 0, h->stream>>>(
            h->d_tt_desc, qx, qy, qz, h->d_state, nullptr, nullptr, 0, 0, 0.0, 0, n_q, h->d_keys,
            (uint32_t)h->base_off, nullptr);
    } else if (v == 99) {          // debug: per-thread walk with visit counters (icp_debug_counters)
        k_nn_bvh<true><<<g128, 128, 0, h->stream>>>(h->d_bvh_nodes, h->d_bvh_sorted, h->bvh_P, qx, qy, qz, n_q,
                                                    h->d_keys, (uint32_t)h->base_off, &h->d_state->converged);
    } else if (v == 10) {   // per-thread walk
        k_nn_bvh<false><<<g128, 128, 0, h->stream>>>(h->d_bvh_nodes, h->d_bvh_sorted, h->bvh_P, qx, qy, qz, n_q,
                                                     h->d_keys, (uint32_t)h->base_off, &h->d_state->converged);
    } else if (v == 98) {   // debug: warp-packet walk with visit counters
        k_nn_bvh_packet<false, true><<<gpk, PKT_THREADS, 0, h->stream>>>(
            h->d_bvh_nodes, h->d_bvh_sorted, h->bvh_P, qx, qy, qz, h->d_state, nullptr, nullptr, 0, 0, 0.0, 0,
            n_q, h->d_keys, (uint32_t)h->base_off, nullptr);
    } else {                // default: warp-packet walk
        k_nn_bvh_packet<false, false><<<gpk, PKT_THREADS, 0, h->stream>>>(
            h->d_bvh_nodes, h->d_bvh_sorted, h->bvh_P, qx, qy, qz, h->d_state, nullptr, nullptr, 0, 0, 0.0, 0,
            n_q, h->d_keys, (uint32_t)h->base_off, nullptr);
    }
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

// slot order of the queries (one-time per icp_set_moving / stride change, BVH mode only)
int ensure_query_order(icp_handle* h, long long n_q)
{
    const int stride = h->prm.stride > 1 ? h->prm.stride : 1;
    if (h->order_valid && h->order_stride == stride) {
        if (h->slots_stale && n_q > 0) {      // new coordinates, same slot order (icp_set_moving)
            k_gather_mov_slots<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(
                h->d_mov, h->has_vlist ? h->d_vlist : nullptr, stride, h->d_slot_pos, n_q, h->d_mov_slots);
            h->launches += 1;
            CUDA_TRY(h, cudaGetLastError());
        }
        h->slots_stale = false;
        return ICP_OK;
    }
    BuildTrace bt_(h, "query_order");
    h->slots_stale = false;
    h->warm_valid = false;
    if (n_q > h->order_capacity || !h->d_slot_pos) {
        if (h->d_slot_pos) cudaFree(h->d_slot_pos);
        if (h->d_pos_slot) cudaFree(h->d_pos_slot);
        if (h->d_mov_slots) cudaFree(h->d_mov_slots);
        h->d_slot_pos = h->d_pos_slot = nullptr;
        h->d_mov_slots = nullptr;
        const long long cap = std::max<long long>(n_q, 1);
        CUDA_TRY(h, cudaMalloc((void**)&h->d_slot_pos, (size_t)cap * 4));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_pos_slot, (size_t)cap * 4));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_mov_slots, (size_t)cap * sizeof(float4)));
        h->order_capacity = cap;
    }
    if (n_q > 0) {
        // one grow-only scratch buffer, carved up; everything is stream-ordered (no host sync)
        const int* vl = h->has_vlist ? h->d_vlist : nullptr;
        const int pb = (int)std::min<long long>((n_q + 255) / 256, (long long)h->sm_count * 4);
        size_t temp_bytes = 0;
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr,
                                                    (unsigned long long*)nullptr, (uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, (int)n_q, 0, 63, h->stream));
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t o_codes = 0, o_codes_out = o_codes + up((size_t)n_q * 8), o_pos = o_codes_out + up((size_t)n_q * 8),
                     o_part = o_pos + up((size_t)n_q * 4), o_bbox = o_part + up((size_t)pb * 24), o_temp = o_bbox + 256,
                     total = o_temp + up(temp_bytes);
        if (total > h->order_scratch_bytes) {
            if (h->d_order_scratch) cudaFree(h->d_order_scratch);
            h->d_order_scratch = nullptr; h->order_scratch_bytes = 0;
            CUDA_TRY(h, cudaMalloc(&h->d_order_scratch, total));
            h->order_scratch_bytes = total;
        }
        char* sb = (char*)h->d_order_scratch;
        unsigned long long* codes = (unsigned long long*)(sb + o_codes);
        unsigned long long* codes_out = (unsigned long long*)(sb + o_codes_out);
        uint32_t* pos = (uint32_t*)(sb + o_pos);
        float* part = (float*)(sb + o_part);
        float* bbox = (float*)(sb + o_bbox);
        k_query_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_mov, vl, stride, n_q, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        k_query_codes<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(h->d_mov, vl, stride, n_q, bbox, codes, pos);
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(sb + o_temp, temp_bytes, codes, codes_out, pos, h->d_slot_pos, (int)n_q, 0, 63, h->stream));
        k_invert_permutation<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(h->d_slot_pos, n_q, h->d_pos_slot);
        k_gather_mov_slots<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(h->d_mov, vl, stride, h->d_slot_pos, n_q, h->d_mov_slots);
        h->launches += 6;
        CUDA_TRY(h, cudaGetLastError());
    }
    h->order_valid = true;
    h->order_stride = stride;
    return ICP_OK;
}

int moments_grid(const icp_handle* h, long long n)
{
    long long b = (n + MOM_THREADS - 1) / MOM_THREADS;
    return (int)std::max<long long>(1, std::min<long long>(b, (long long)h->sm_count * 4));
}

int ensure_query_buffers(icp_handle* h, long long n_q)
{
    if (n_q > h->q_capacity || !h->d_q) {
        if (h->d_q) cudaFree(h->d_q);
        if (h->d_keys) cudaFree(h->d_keys);
        h->d_q = nullptr; h->d_keys = nullptr; h->q_capacity = 0;
        h->warm_valid = false; h->bf_warm_valid = false;
        const long long cap = std::max<long long>(n_q, 1);
        CUDA_TRY(h, cudaMalloc((void**)&h->d_q, (size_t)cap * 3 * sizeof(float)));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_keys, (size_t)cap * sizeof(unsigned long long)));
        h->q_capacity = cap;
    }
    const int g = moments_grid(h, n_q);
    if (g > h->partials_capacity) {
        if (h->d_partials) cudaFree(h->d_partials);
        h->d_partials = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_partials, (size_t)g * MOM * sizeof(double)));
        h->partials_capacity = g;
    }
    return ICP_OK;
}

int check_ready(icp_handle* h)
{
    if (!h) return ICP_ERR_BAD_ARG;
    if (!h->have_base) return fail(h, ICP_ERR_STATE, "icp_set_base has not been called");
    if (!h->have_moving) return fail(h, ICP_ERR_STATE, "icp_set_moving has not been called");
    if (!h->have_matrices) return fail(h, ICP_ERR_STATE, "icp_set_matrices has not been called");
    if (!(h->prm.thresh > 0))
        return fail(h, ICP_ERR_BAD_ARG, "thresh must be > 0 (the reference's make_pairs returns None otherwise, general.py:277)");
    return ICP_OK;
}

int launch_nn_surface(icp_handle* h, long long n_q);

// transform + nearest neighbour (+ shard resolution + percentile select)
int enqueue_nn(icp_handle* h, long long n_q)
{
    int rc = ensure_query_buffers(h, n_q);
    if (rc) return rc;
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const int stride = h->prm.stride > 1 ? h->prm.stride : 1;
    const bool slots = uses_slots(h);    // Morton-ordered query slots (tree searches, sorted brute force)
    const bool tree = h->prm.nn_mode != ICP_NN_BRUTEFORCE;
    const bool sharded_base = h->prm.world_size > 1 && h->prm.shard_mode == ICP_SHARD_BASE;
    const unsigned long long* warm_src = nullptr;
    if (slots) {
        rc = ensure_query_order(h, n_q);
        if (rc) return rc;
    }
    if (tree) {
        if (h->warm_valid) warm_src = h->d_keys;
    } else {
        h->warm_valid = false;
        // brute force: any real local vertex is a valid seed, so the previous winners stay usable
        // for as long as the base and the query count are unchanged (a new pose or new moving
        // coordinates only change how tight the seed is)
        if (h->bf_warm_valid && h->bf_warm_nq == n_q && h->prm.nn_variant != 1 && n_q > 0)
            warm_src = sharded_base ? h->d_keys_local : h->d_keys;
    }
    h->bf_warm_active = !tree && warm_src != nullptr;
    // sharded base: d_keys still hold the previous pass's GLOBAL winners and d_q its queries
    h->bound_active = h->bf_warm_active && sharded_base && h->prm.nn_variant != 58;
    if (h->bound_active) {
        rc = ensure(h, h->d_bound, h->bound_capacity, n_q);
        if (rc) return rc;
    }
    {
        ProfScope ps(h, 3);
        if (h->pose_dirty) {          // k_solve_update keeps imx1 current between iterations
            k_prepare_pose<<<1, 1, 0, h->stream>>>(h->d_state);
            h->launches += 1;
            h->pose_dirty = false;
        }
        // a seeded pass of the q16 kernel follows: K1 accumulates the exponent histogram of its class machinery
        const int v_ = h->prm.nn_variant;
        h->q16_hist_by_k1 = h->bf_warm_active && n_q > 0 && h->n_tiles >= 16 && (v_ == 0 || q16_variant(v_)) && q16_classes_wanted(h, n_q);
        if (h->q16_hist_by_k1) {
            rc = ensure_q16_classes(h, n_q);
            if (rc) return rc;
        }
        if (n_q > 0)
            k_transform_queries<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(
                h->d_state, h->d_mov, h->has_vlist ? h->d_vlist : nullptr, stride, n_q,
                slots ? h->d_slot_pos : nullptr, qx, qy, qz, h->d_keys,
                warm_src, h->d_base4, h->surface ? h->d_tris4 : nullptr, h->base_off,
                h->n_local, h->bound_active ? h->d_keys : nullptr, h->bound_active ? h->d_bound : nullptr,
                h->q16_hist_by_k1 ? h->d_q16_hist : nullptr, h->q16_pass);
        h->launches += 1;
    }
    if (h->surface != (h->prm.nn_mode == ICP_NN_SURFACE))
        return fail(h, ICP_ERR_STATE, "nn_mode ICP_NN_SURFACE needs icp_set_base_triangles (and the vertex modes need icp_set_base)");
    rc = h->surface ? launch_nn_surface(h, n_q)
                    : (h->prm.nn_mode == ICP_NN_BVH ? launch_nn_bvh(h, n_q) : launch_nn(h, n_q));
    if (rc) return rc;
    if (h->prm.world_size > 1 && h->prm.shard_mode == ICP_SHARD_QUERIES && h->prm.reject_fraction > 0)
        return fail(h, ICP_ERR_BAD_ARG, "reject_fraction needs every query on every rank: use ICP_SHARD_BASE");
    if (sharded_base && n_q > 0) {
        if (!h->comm) return fail(h, ICP_ERR_STATE, "world_size > 1 but icp_comm_init was not called");
        ProfScope ps(h, 4);
        if (!tree) {       // keep this rank's own winners: they seed the next pass
            rc = ensure(h, h->d_keys_local, h->keys_local_capacity, n_q);
            if (rc) return rc;
            CUDA_TRY(h, cudaMemcpyAsync(h->d_keys_local, h->d_keys, (size_t)n_q * sizeof(unsigned long long),
                                        cudaMemcpyDeviceToDevice, h->stream));
        }
        NCCL_TRY(h, g_nccl.AllReduce(h->d_keys, h->d_keys, (size_t)n_q, ncclUint64, ncclMin, h->comm,
                                     h->stream));
    }
    h->warm_valid = tree;         // d_keys now hold this pass's winners, slot-aligned
    h->bf_warm_valid = !tree && n_q > 0;
    h->bf_warm_nq = n_q;
    if (h->prm.reject_fraction > 0 && n_q > 0) {
        ProfScope ps(h, 1);
        const long long keep = n_q - (long long)std::floor(h->prm.reject_fraction * (double)n_q);
        k_select_init<<<1, 256, 0, h->stream>>>(h->d_state, keep - 1, h->d_hist);
        const int g = (int)std::min<long long>((n_q + 255) / 256, (long long)h->sm_count * 8);
        for (int byte = 7; byte >= 0; --byte) {
            k_select_hist<<<g, 256, 0, h->stream>>>(h->d_state, h->d_keys, slots ? h->d_slot_pos : nullptr, n_q, byte, h->d_hist);
            k_select_pick<<<1, 256, 0, h->stream>>>(h->d_state, byte, h->d_hist);
        }
        h->launches += 17;
    }
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int launch_nn_surface(icp_handle* h, long long n_q)
{
    if (n_q <= 0 || h->n_local <= 0) return ICP_OK;
    ProfScope ps(h, 0);
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const unsigned g = (unsigned)((n_q + SURF_THREADS - 1) / SURF_THREADS);
    if (h->prm.nn_variant == 10) {       // per-thread walk (kept for A/B runs and as a cross-check)
        k_nn_surface<<<(unsigned)((n_q + 127) / 128), 128, 0, h->stream>>>(h->d_tri_nodes, h->d_tri_sorted, h->tri_runs, qx, qy, qz,
                                               n_q, h->d_keys, (uint32_t)h->base_off, h->tri_extent,
                                               &h->d_state->converged);
    } else if (h->prm.nn_variant == 11 || h->n_local >= (1ll << 27)) {   // warp-packet walk, eager triangle tests
        k_nn_surface_packet<false><<<g, SURF_THREADS, 0, h->stream>>>(
            h->d_tri_nodes, h->d_tri_sorted, h->tri_runs, qx, qy, qz, h->d_state, nullptr, nullptr, 0, 0, 0.0, 0,
            n_q, h->d_keys, (uint32_t)h->base_off, h->tri_extent, nullptr);
    } else {                             // default: warp-packet walk, deferred (queued) triangle tests
        k_nn_surface_defer<false><<<g, SURF_THREADS, 0, h->stream>>>(
            h->d_tri_nodes, h->d_tri_sorted, h->tri_runs, qx, qy, qz, h->d_state, nullptr, nullptr, 0, 0, 0.0, 0,
            n_q, h->d_keys, (uint32_t)h->base_off, h->tri_extent, nullptr, TTTail{});
    }
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

// sum `g` per-CTA partial vectors into h->d_moments: one CTA, or 64 slice CTAs + one when there are many
int launch_reduce(icp_handle* h, const DevState* st, int g)
{
    constexpr int SLICES = 64;
    if (g <= 2048) {
        k_reduce_moments<<<1, MOM * 32, 0, h->stream>>>(st, h->d_partials, g, h->d_moments);
        h->launches += 1;
        return ICP_OK;
    }
    if (!h->d_partials2) CUDA_TRY(h, cudaMalloc((void**)&h->d_partials2, (size_t)SLICES * MOM * sizeof(double)));
    const int per = (g + SLICES - 1) / SLICES;
    const int nb = (g + per - 1) / per;
    k_reduce_moments<<<nb, MOM * 32, 0, h->stream>>>(st, h->d_partials, g, h->d_partials2, per);
    k_reduce_moments<<<1, MOM * 32, 0, h->stream>>>(st, h->d_partials2, nb, h->d_moments);
    h->launches += 2;
    return ICP_OK;
}

int ensure_partials(icp_handle* h, int g)
{
    if (g > h->partials_capacity) {
        if (h->d_partials) cudaFree(h->d_partials);
        h->d_partials = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_partials, (size_t)g * MOM * sizeof(double)));
        h->partials_capacity = g;
    }
    return ICP_OK;
}

// BVH mode on one GPU without percentile cut: transform + search + filter + moments in ONE kernel
int enqueue_iteration_fused(icp_handle* h, long long n_q)
{
    int rc = ensure_query_buffers(h, n_q);
    if (rc) return rc;
    rc = ensure_query_order(h, n_q);
    if (rc) return rc;
    if (h->surface != (h->prm.nn_mode == ICP_NN_SURFACE))
        return fail(h, ICP_ERR_STATE, "nn_mode ICP_NN_SURFACE needs icp_set_base_triangles (and the vertex modes need icp_set_base)");
    const bool tiletree = !h->surface && tiletree_variant(h);
    if (!h->surface) {
        rc = tiletree ? build_tiletree(h) : build_bvh(h);
        if (rc) return rc;
    }
    const int cta = h->surface ? SURF_THREADS : (tiletree ? TT_THREADS : PKT_THREADS);
    const int g = (int)std::max<long long>(1, (n_q + cta - 1) / cta);
    rc = ensure_partials(h, g);
    if (rc) return rc;
    const bool persistent = tiletree && tiletree_persistent(h) && n_q > 0;
    if (persistent) {
        rc = ensure_tiletree_walk(h, n_q, true);
        if (rc) return rc;
    }
    // surface walk (default variant): its warps feed the same ticket tree, the iteration is one launch as well
    const bool surf_ticket = h->surface && h->prm.nn_variant != 11 && h->prm.nn_variant != 14 && h->n_local < (1ll << 27) && n_q > 0;
    if (surf_ticket) {
        rc = ensure_tiletree_walk(h, n_q, true, (long long)g * SURF_WARPS);
        if (rc) return rc;
    }
    if (h->pose_dirty) {
        ProfScope ps(h, 3);
        k_prepare_pose<<<1, 1, 0, h->stream>>>(h->d_state);
        h->launches += 1;
        h->pose_dirty = false;
    }
    if (persistent) {
        // the whole iteration is ONE launch on one GPU: walk + filter + moments + ticketed reduction + solve
        {
            ProfScope ps(h, 0);
            rc = launch_tiletree_walk<true>(h, n_q, h->prm.world_size == 1 ? 1 : 0);
            if (rc) return rc;
        }
        h->warm_valid = true;
        h->bf_warm_valid = false;
        if (h->prm.world_size > 1) {      // queries sharded: the only exchange is the 20-double moment sum
            if (!h->comm) return fail(h, ICP_ERR_STATE, "world_size > 1 but icp_comm_init was not called");
            {
                ProfScope ps(h, 4);
                NCCL_TRY(h, g_nccl.AllReduce(h->d_moments, h->d_moments, MOM, ncclDouble, ncclSum, h->comm, h->stream));
            }
            ProfScope ps(h, 2);
            k_solve_update<<<1, 32, 0, h->stream>>>(h->d_state, h->d_moments, h->prm.align_meth == 1,
                                                    h->prm.use_target, h->prm.target_d, h->d_stats,
                                                    h->stats_capacity);
            h->launches += 1;
        }
        CUDA_TRY(h, cudaGetLastError());
        return ICP_OK;
    }
    h->tt_win_valid = false;
    {
        ProfScope ps(h, 0);
        if (h->surface && (h->prm.nn_variant == 11 || h->n_local >= (1ll << 27)))
            k_nn_surface_packet<true><<<g, SURF_THREADS, 0, h->stream>>>(
                h->d_tri_nodes, h->d_tri_sorted, h->n_local > 0 ? h->tri_runs : 0, nullptr, nullptr, nullptr, h->d_state,
                h->d_mov_slots, h->d_tris4, h->base_off, h->n_local, h->prm.thresh, h->warm_valid ? 1 : 0, n_q,
                h->d_keys, (uint32_t)h->base_off, h->tri_extent, h->d_partials);
        else if (surf_ticket)
            k_nn_surface_defer<true, true><<<g, SURF_THREADS, 0, h->stream>>>(
                h->d_tri_nodes, h->d_tri_sorted, h->n_local > 0 ? h->tri_runs : 0, nullptr, nullptr, nullptr, h->d_state,
                h->d_mov_slots, h->d_tris4, h->base_off, h->n_local, h->prm.thresh, h->warm_valid ? 1 : 0, n_q,
                h->d_keys, (uint32_t)h->base_off, h->tri_extent, nullptr, make_tail(h, h->prm.world_size == 1 ? 1 : 0));
        else if (h->surface)        // nn_variant 14: walk + separate reduce and solve launches (A/B of the ticketed tail)
            k_nn_surface_defer<true><<<g, SURF_THREADS, 0, h->stream>>>(
                h->d_tri_nodes, h->d_tri_sorted, h->n_local > 0 ? h->tri_runs : 0, nullptr, nullptr, nullptr, h->d_state,
                h->d_mov_slots, h->d_tris4, h->base_off, h->n_local, h->prm.thresh, h->warm_valid ? 1 : 0, n_q,
                h->d_keys, (uint32_t)h->base_off, h->tri_extent, h->d_partials, TTTail{});
        else if (tiletree)
            k_nn_tiletree<true><<<g, TT_THREADS, 0, h->stream>>>(
                h->d_tt_desc, nullptr, nullptr, nullptr, h->d_state,
                h->d_mov_slots, h->d_base4, h->base_off, h->n_local, h->prm.thresh, h->warm_valid ? 1 : 0, n_q,
                h->d_keys, (uint32_t)h->base_off, h->d_partials);
        else
            k_nn_bvh_packet<true, false><<<g, PKT_THREADS, 0, h->stream>>>(
                h->d_bvh_nodes, h->d_bvh_sorted, h->n_local > 0 ? h->bvh_P : 0, nullptr, nullptr, nullptr, h->d_state,
                h->d_mov_slots, h->d_base4, h->base_off, h->n_local, h->prm.thresh, h->warm_valid ? 1 : 0, n_q,
                h->d_keys, (uint32_t)h->base_off, h->d_partials);
        h->launches += 1;
    }
    h->warm_valid = true;
    h->bf_warm_valid = false;
    if (surf_ticket) {
        if (h->prm.world_size > 1) {      // queries sharded: the only exchange is the 20-double moment sum
            if (!h->comm) return fail(h, ICP_ERR_STATE, "world_size > 1 but icp_comm_init was not called");
            {
                ProfScope ps(h, 4);
                NCCL_TRY(h, g_nccl.AllReduce(h->d_moments, h->d_moments, MOM, ncclDouble, ncclSum, h->comm, h->stream));
            }
            ProfScope ps(h, 2);
            k_solve_update<<<1, 32, 0, h->stream>>>(h->d_state, h->d_moments, h->prm.align_meth == 1,
                                                    h->prm.use_target, h->prm.target_d, h->d_stats,
                                                    h->stats_capacity);
            h->launches += 1;
        }
        CUDA_TRY(h, cudaGetLastError());
        return ICP_OK;
    }
    if (h->prm.world_size == 1 && g <= 16384) {      // one launch: reduce + solve + pose update
        ProfScope ps(h, 2);
        k_reduce_solve<<<1, MOM * 32, 0, h->stream>>>(h->d_state, h->d_partials, g, h->d_moments, h->prm.align_meth == 1,
                                                      h->prm.use_target, h->prm.target_d, h->d_stats, h->stats_capacity);
        h->launches += 1;
        CUDA_TRY(h, cudaGetLastError());
        return ICP_OK;
    }
    {
        ProfScope ps(h, 2);
        rc = launch_reduce(h, h->d_state, g);
        if (rc) return rc;
    }
    if (h->prm.world_size > 1) {      // queries sharded: the only exchange is the 20-double moment sum
        if (!h->comm) return fail(h, ICP_ERR_STATE, "world_size > 1 but icp_comm_init was not called");
        ProfScope ps(h, 4);
        NCCL_TRY(h, g_nccl.AllReduce(h->d_moments, h->d_moments, MOM, ncclDouble, ncclSum, h->comm, h->stream));
    }
    {
        ProfScope ps(h, 2);
        k_solve_update<<<1, 32, 0, h->stream>>>(h->d_state, h->d_moments, h->prm.align_meth == 1,
                                                h->prm.use_target, h->prm.target_d, h->d_stats,
                                                h->stats_capacity);
        h->launches += 1;
    }
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int enqueue_iteration(icp_handle* h, long long n_q)
{
    if ((h->prm.nn_mode == ICP_NN_BVH || h->prm.nn_mode == ICP_NN_SURFACE) &&
        (h->prm.world_size == 1 || h->prm.shard_mode == ICP_SHARD_QUERIES) &&
        !(h->prm.reject_fraction > 0) &&
        h->prm.nn_variant != 10 && (h->prm.nn_variant != 11 || h->surface) && h->prm.nn_variant != 99 && h->prm.nn_variant != 98)
        return enqueue_iteration_fused(h, n_q);
    int rc = enqueue_nn(h, n_q);
    if (rc) return rc;
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const int stride = h->prm.stride > 1 ? h->prm.stride : 1;
    const int g = moments_grid(h, n_q);
    {
        ProfScope ps(h, 1);
        k_filter_moments<<<g, MOM_THREADS, 0, h->stream>>>(
            h->d_state, h->d_mov, h->has_vlist ? h->d_vlist : nullptr, stride, n_q, qx, qy, qz,
            h->d_keys, h->d_base4, h->surface ? h->d_tris4 : nullptr, h->base_off, h->n_local, h->prm.thresh,
            h->prm.reject_fraction > 0 ? 1 : 0,
            uses_slots(h) ? h->d_slot_pos : nullptr, uses_slots(h) ? h->d_mov_slots : nullptr, h->d_partials);
        h->launches += 1;
        if (h->prm.world_size == 1) {
            k_reduce_solve<<<1, MOM * 32, 0, h->stream>>>(h->d_state, h->d_partials, g, h->d_moments, h->prm.align_meth == 1,
                                                          h->prm.use_target, h->prm.target_d, h->d_stats, h->stats_capacity);
            h->launches += 1;
            CUDA_TRY(h, cudaGetLastError());
            return ICP_OK;
        }
        k_reduce_moments<<<1, MOM * 32, 0, h->stream>>>(h->d_state, h->d_partials, g, h->d_moments);
        h->launches += 1;
    }
    if (h->prm.world_size > 1) {
        ProfScope ps(h, 4);
        NCCL_TRY(h, g_nccl.AllReduce(h->d_moments, h->d_moments, MOM, ncclDouble, ncclSum, h->comm,
                                     h->stream));
    }
    {
        ProfScope ps(h, 2);
        k_solve_update<<<1, 32, 0, h->stream>>>(h->d_state, h->d_moments, h->prm.align_meth == 1,
                                                h->prm.use_target, h->prm.target_d, h->d_stats,
                                                h->stats_capacity);
        h->launches += 1;
    }
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int ensure_stats(icp_handle* h, int n)
{
    if (n <= h->stats_capacity) return ICP_OK;
    n = std::max(n, std::max(64, 2 * h->stats_capacity));      // geometric: a step()-per-tick caller reallocates O(log) times
    // keep what is there (n_iter indexes into it)
    icp_iter_stats* nw = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&nw, (size_t)n * sizeof(icp_iter_stats)));
    CUDA_TRY(h, cudaMemsetAsync(nw, 0, (size_t)n * sizeof(icp_iter_stats), h->stream));
    if (h->d_stats) {
        CUDA_TRY(h, cudaMemcpyAsync(nw, h->d_stats, (size_t)h->stats_capacity * sizeof(icp_iter_stats),
                                    cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_stats);
    }
    h->d_stats = nw;
    h->stats_capacity = n;
    return ICP_OK;
}

int read_state(icp_handle* h, DevState* out)
{
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ICP_OK;
}

int device_error(icp_handle* h, const DevState& st)
{
    if (st.error == ICP_ERR_TOO_FEW_PAIRS)
        return fail(h, ICP_ERR_TOO_FEW_PAIRS, "input arrays are of wrong shape or type");
    if (st.error == ICP_ERR_SINGULAR)
        return fail(h, ICP_ERR_SINGULAR, "Matrix.invert(ed): matrix does not have an inverse");
    if (st.error) return fail(h, st.error, "device-side error %d", st.error);
    return ICP_OK;
}

// The device-side `converged` flag is the stop flag of icp_run's enqueued iterations: every kernel
// returns early on it.  icp_nn / icp_make_pairs are stand-alone queries and must not be silenced by a
// run that converged earlier on this handle: they lift the latch for their own launches and put it
// back afterwards (the convergence state of the run is untouched).
int suspend_stop(icp_handle* h, int* saved)
{
    DevState st;
    int rc = read_state(h, &st);
    if (rc) return rc;
    *saved = (st.converged && !st.error) ? 1 : 0;
    if (*saved) {
        const int zero = 0;
        CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, converged), &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return ICP_OK;
}

int restore_stop(icp_handle* h, int saved)
{
    if (!saved) return ICP_OK;
    const int one = 1;
    CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, converged), &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ICP_OK;
}

int clear_device_error(icp_handle* h)
{
    // error + its stop flag are cleared so the handle stays usable after a failed call
    DevState st;
    int rc = read_state(h, &st);
    if (rc) return rc;
    if (st.error) {
        h->pose_dirty = true;
        st.error = 0;
        st.converged = 0;
        CUDA_TRY(h, cudaMemcpyAsync(h->d_state, &st, sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    return ICP_OK;
}

}  // namespace

// ==============================================================================================
// C ABI
// ==============================================================================================
extern "C" {

int icp_version(void) { return ICP_B200_VERSION; }

const char* icp_last_error(const icp_handle* h) { return h ? h->err : g_create_error; }

int icp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -ICP_ERR_CUDA; }
    return n;
}

int icp_create(const icp_params* p, icp_handle** out)
{
    if (!p || !out) return fail(nullptr, ICP_ERR_BAD_ARG, "icp_create: null argument");
    if (p->struct_size != (int32_t)sizeof(icp_params))
        return fail(nullptr, ICP_ERR_BAD_ARG, "icp_create: struct_size %d != %zu", p->struct_size, sizeof(icp_params));
    if (p->world_size < 1 || p->rank < 0 || p->rank >= p->world_size)
        return fail(nullptr, ICP_ERR_BAD_ARG, "icp_create: bad rank/world_size");
    *out = nullptr;
    icp_handle* h = new (std::nothrow) icp_handle();
    if (!h) return fail(nullptr, ICP_ERR_OOM, "icp_create: out of host memory");
    h->prm = *p;
#define CREATE_TRY(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fail(nullptr, ICP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));          \
            delete h;                                                                             \
            return ICP_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)
    CREATE_TRY(cudaSetDevice(p->device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major != 10) {
        fail(nullptr, ICP_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             p->device, prop.major, prop.minor);
        delete h;
        return ICP_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&h->ev0));
    CREATE_TRY(cudaEventCreate(&h->ev1));
    CREATE_TRY(cudaMalloc((void**)&h->d_state, sizeof(DevState)));
    CREATE_TRY(cudaMemset(h->d_state, 0, sizeof(DevState)));
    CREATE_TRY(cudaMalloc((void**)&h->d_moments, MOM * sizeof(double)));
    CREATE_TRY(cudaMalloc((void**)&h->d_hist, 256 * sizeof(unsigned int)));
    CREATE_TRY(cudaMalloc((void**)&h->d_total, sizeof(long long)));
#undef CREATE_TRY
    *out = h;
    int rc = icp_reset_convergence(h);
    if (rc) { snprintf(g_create_error, 512, "%s", h->err); icp_destroy(h); *out = nullptr; return rc; }
    return ICP_OK;
}

int icp_destroy(icp_handle* h)
{
    if (!h) return ICP_OK;
    cudaSetDevice(h->prm.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    void* ptrs[] = {h->d_base_tiles, h->d_base4, h->d_mov, h->d_vlist, h->d_q, h->d_keys, h->d_partials,
                    h->d_moments, h->d_hist, h->d_state, h->d_stats, h->d_flags, h->d_block_counts,
                    h->d_total, h->d_pairs, h->d_v, h->d_scratch, h->d_bvh_nodes, h->d_bvh_sorted, h->d_slot_pos, h->d_pos_slot, h->d_mov_slots, h->d_order_scratch, h->d_filter_tiles, h->d_filter_info, h->d_tris4, h->d_tri_sorted, h->d_tri_nodes, h->d_keys_local, h->d_cyl_ftiles, h->d_cyl_etiles, h->d_cyl_hdrs, h->d_slab_ftiles, h->d_slab_etiles, h->d_slab_hdrs, h->d_partials2, h->d_q16_tiles, h->d_q16_etiles, h->d_bound, h->d_tt_boxes, h->d_tt_tiles, h->d_tt_desc, h->d_tt_queue, h->d_win_tile, h->d_tt_part, h->d_tt_tickets, h->d_q16_cls, h->d_q16_ids, h->d_q16_hist, h->d_q16_sort_temp, h->d_build_scratch};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int w = 0; w < 5; ++w)
        for (auto& e : h->ev_pool[w]) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return ICP_OK;
}

int icp_set_params(icp_handle* h, const icp_params* p)
{
    if (!h || !p) return ICP_ERR_BAD_ARG;
    if (p->struct_size != (int32_t)sizeof(icp_params)) return fail(h, ICP_ERR_BAD_ARG, "bad struct_size");
    if (p->device != h->prm.device || p->rank != h->prm.rank || p->world_size != h->prm.world_size)
        return fail(h, ICP_ERR_BAD_ARG, "device / rank / world_size cannot change after icp_create");
    h->prm = *p;
    return ICP_OK;
}

int icp_set_base(icp_handle* h, const float* xyz, int64_t n_local, int64_t global_offset, int64_t n_global)
{
    if (!h) return ICP_ERR_BAD_ARG;
    if (n_local < 0 || global_offset < 0 || n_global < n_local || (n_local > 0 && !xyz))
        return fail(h, ICP_ERR_BAD_ARG, "icp_set_base: bad sizes");
    if (n_global > 0xFFFFFFFELL) return fail(h, ICP_ERR_BAD_ARG, "icp_set_base: more than 2^32-2 base vertices");
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    if (h->d_base_tiles) { cudaFree(h->d_base_tiles); h->d_base_tiles = nullptr; }
    if (h->d_base4) { cudaFree(h->d_base4); h->d_base4 = nullptr; }
    h->n_tiles = (int)((n_local + NN_TILE - 1) / NN_TILE);
    const long long n_pad = (long long)h->n_tiles * NN_TILE;
    h->n_local = n_local; h->n_global = n_global; h->base_off = global_offset;
    h->surface = false;
    h->bvh_built = false;
    h->tt_built = false;
    h->filter_built = false;
    h->cyl_sorted = -1;
    h->slab_order = -1;
    h->q16_order = -1;
    h->warm_valid = false;
    h->bf_warm_valid = false;
    if (n_local > 0) {
        float* d_raw = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_base_tiles, (size_t)n_pad * 3 * sizeof(float)));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_base4, (size_t)n_local * sizeof(float4)));
        CUDA_TRY(h, cudaMalloc((void**)&d_raw, (size_t)n_local * 3 * sizeof(float)));
        CUDA_TRY(h, cudaMemcpyAsync(d_raw, xyz, (size_t)n_local * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        k_build_base<<<(unsigned)((n_pad + 255) / 256), 256, 0, h->stream>>>(d_raw, n_local, n_pad, h->d_base_tiles, h->d_base4);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(d_raw);
    }
    h->have_base = true;
    return ICP_OK;
}

int icp_set_base_triangles(icp_handle* h, const float* tri_xyz, int64_t n_local, int64_t global_offset, int64_t n_global)
{
    if (!h) return ICP_ERR_BAD_ARG;
    if (n_local < 0 || global_offset < 0 || n_global < n_local || (n_local > 0 && !tri_xyz))
        return fail(h, ICP_ERR_BAD_ARG, "icp_set_base_triangles: bad sizes");
    if (n_global > 0xFFFFFFFELL) return fail(h, ICP_ERR_BAD_ARG, "icp_set_base_triangles: too many triangles");
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    for (void* p : {(void*)h->d_tris4, (void*)h->d_tri_sorted, (void*)h->d_tri_nodes}) if (p) cudaFree(p);
    h->d_tris4 = nullptr; h->d_tri_sorted = nullptr; h->d_tri_nodes = nullptr;
    h->n_local = n_local; h->n_global = n_global; h->base_off = global_offset;
    h->n_tiles = 0;
    h->surface = true;
    h->bvh_built = false; h->tt_built = false; h->filter_built = false; h->warm_valid = false; h->bf_warm_valid = false;
    h->q16_order = -1; h->slab_order = -1; h->cyl_sorted = -1;
    h->tri_runs = 0; h->tri_extent = 0.0f;
    if (n_local > 0) {
        const long long n = n_local;
        const int n_runs = (int)((n + TRI_LEAF - 1) / TRI_LEAF);
        const long long n_pad = (long long)n_runs * TRI_LEAF;
        float* d_raw = nullptr; float4* cent = nullptr;
        unsigned long long *codes = nullptr, *codes_out = nullptr, *run_keys = nullptr;
        uint32_t *order = nullptr, *order_out = nullptr;
        int* leaf_parent = nullptr;
        float *part = nullptr, *bbox = nullptr;
        void* temp = nullptr;
        size_t temp_bytes = 0;
        const int pb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 4);
        CUDA_TRY(h, cudaMalloc((void**)&d_raw, (size_t)n * 9 * 4));
        CUDA_TRY(h, cudaMalloc((void**)&cent, (size_t)n * 16));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tris4, (size_t)n * 3 * 16));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tri_sorted, (size_t)n_pad * 3 * 16));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_tri_nodes, (size_t)std::max(1, n_runs - 1) * sizeof(BvhNode)));
        CUDA_TRY(h, cudaMalloc((void**)&codes, (size_t)n * 8));
        CUDA_TRY(h, cudaMalloc((void**)&codes_out, (size_t)n * 8));
        CUDA_TRY(h, cudaMalloc((void**)&order, (size_t)n * 4));
        CUDA_TRY(h, cudaMalloc((void**)&order_out, (size_t)n * 4));
        CUDA_TRY(h, cudaMalloc((void**)&run_keys, (size_t)n_runs * 8));
        CUDA_TRY(h, cudaMalloc((void**)&leaf_parent, (size_t)n_runs * 4));
        CUDA_TRY(h, cudaMalloc((void**)&part, (size_t)pb * 24));
        CUDA_TRY(h, cudaMalloc((void**)&bbox, 24));
        CUDA_TRY(h, cudaMemcpyAsync(d_raw, tri_xyz, (size_t)n * 9 * 4, cudaMemcpyHostToDevice, h->stream));
        k_tri_pack<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_raw, n, h->d_tris4, cent);
        // bounding box over all triangle CORNERS (3n float4, stride 1) -> Morton grid + coordinate extent
        k_bbox_partial<<<pb, 256, 0, h->stream>>>(h->d_tris4, 3 * n, part);
        k_bbox_final<<<1, 192, 0, h->stream>>>(part, pb, bbox);
        k_morton_codes<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(cent, n, bbox, codes, order);
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
        CUDA_TRY(h, cudaMalloc(&temp, temp_bytes ? temp_bytes : 1));
        CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, codes, codes_out, order, order_out, (int)n, 0, 63, h->stream));
        k_tri_gather<<<(unsigned)((n_pad + 255) / 256), 256, 0, h->stream>>>(h->d_tris4, order_out, n, n_pad, h->d_tri_sorted);
        k_run_keys_stride<<<(unsigned)((n_runs + 255) / 256), 256, 0, h->stream>>>(codes_out, n_runs, TRI_LEAF, run_keys);
        if (n_runs > 1) {
            k_lbvh_topology<<<(unsigned)((n_runs + 127) / 128), 128, 0, h->stream>>>(run_keys, n_runs, h->d_tri_nodes, leaf_parent);
            k_lbvh_fit_tris<<<(unsigned)((n_runs + 127) / 128), 128, 0, h->stream>>>(h->d_tri_sorted, n, n_runs, h->d_tri_nodes, leaf_parent);
        }
        float hb[6];
        CUDA_TRY(h, cudaMemcpyAsync(hb, bbox, 24, cudaMemcpyDeviceToHost, h->stream));
        h->launches += 8;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ext = 0.0f;
        for (int k = 0; k < 6; ++k) ext = std::max(ext, std::fabs(hb[k]));
        h->tri_extent = ext;
        h->tri_runs = n_runs;
        for (void* p : {(void*)d_raw, (void*)cent, (void*)codes, (void*)codes_out, (void*)order, (void*)order_out,
                        (void*)run_keys, (void*)leaf_parent, (void*)part, (void*)bbox, temp}) cudaFree(p);
    }
    h->have_base = true;
    return ICP_OK;
}

static int set_moving_impl(icp_handle* h, const float* xyz, int64_t n_vertices, const int32_t* vlist, int64_t n_vlist,
                           bool collective);

int icp_set_moving(icp_handle* h, const float* xyz, int64_t n_vertices, const int32_t* vlist, int64_t n_vlist)
{
    return set_moving_impl(h, xyz, n_vertices, vlist, n_vlist, false);
}

int icp_set_moving_collective(icp_handle* h, const float* xyz, int64_t n_vertices, const int32_t* vlist, int64_t n_vlist)
{
    if (h && h->prm.world_size > 1 && !h->comm)
        return fail(h, ICP_ERR_STATE, "icp_set_moving_collective: icp_comm_init was not called");
    return set_moving_impl(h, xyz, n_vertices, vlist, n_vlist, h && h->prm.world_size > 1);
}

static int set_moving_impl(icp_handle* h, const float* xyz, int64_t n_vertices, const int32_t* vlist, int64_t n_vlist,
                           bool collective)
{
    if (!h) return ICP_ERR_BAD_ARG;
    if (n_vertices < 0 || (n_vertices > 0 && !xyz) || n_vlist < 0 || (vlist == nullptr && n_vlist != 0))
        return fail(h, ICP_ERR_BAD_ARG, "icp_set_moving: bad arguments");
    // validate everything before any handle state changes: a rejected call leaves the handle as it was
    if (vlist)
        for (int64_t i = 0; i < n_vlist; ++i)
            if (vlist[i] < 0 || vlist[i] >= n_vertices)
                return fail(h, ICP_ERR_BAD_ARG, "icp_set_moving: vlist[%lld] = %d out of range", (long long)i, vlist[i]);
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    // collective upload: the coordinate array is cut into world_size slices of `per` floats (16-byte
    // multiples); this rank sends its own slice over the host link and ncclAllGather fills in the rest
    const long long W = h->prm.world_size;
    const long long per = collective ? ((n_vertices * 3 + W - 1) / W + 3) / 4 * 4 : 0;
    int rc = ensure(h, h->d_mov, h->mov_capacity, collective ? per * W : n_vertices * 3);
    if (rc) { h->have_moving = false; return rc; }
    if (vlist) {
        rc = ensure(h, h->d_vlist, h->vlist_capacity, n_vlist);
        if (rc) { h->have_moving = false; return rc; }
    }
    // Same vertex count and the same vertex list: this is normally the same mesh again (the operator
    // and the function surface re-upload per call), so the Morton slot order and the previous winners
    // are kept -- any order and any real base vertex / triangle as seed give the exact result, a stale
    // one only costs speed -- and only the slot-ordered copy of the coordinates is refreshed.
    const bool same_vlist = (vlist != nullptr) == h->has_vlist &&
                            (!vlist || ((int64_t)h->h_vlist.size() == n_vlist &&
                                        (n_vlist == 0 || memcmp(h->h_vlist.data(), vlist, (size_t)n_vlist * sizeof(int32_t)) == 0)));
    const bool same_shape = h->have_moving && h->order_valid && n_vertices == h->n_vertices && same_vlist;
    h->have_moving = false;                          // until every copy below has been queued
    if (n_vertices > 0 && !collective) {
        CUDA_TRY(h, cudaMemcpyAsync(h->d_mov, xyz, (size_t)n_vertices * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else if (n_vertices > 0) {
        const long long lo = per * h->prm.rank, hi = std::min<long long>(lo + per, n_vertices * 3);
        if (hi > lo)
            CUDA_TRY(h, cudaMemcpyAsync(h->d_mov + lo, xyz + lo, (size_t)(hi - lo) * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        NCCL_TRY(h, g_nccl.AllGather(h->d_mov + lo, h->d_mov, (size_t)per, ncclFloat, h->comm, h->stream));
    }
    if (vlist && n_vlist > 0 && !same_vlist)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_vlist, vlist, (size_t)n_vlist * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    // moments are accumulated about the first vertex of the (replicated) moving array, so that
    // every rank of a query-sharded run uses the same reference point
    double shift[3] = {0, 0, 0};
    if (n_vertices > 0)
        for (int k = 0; k < 3; ++k) shift[k] = xyz[k];
    CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, shift), shift, sizeof(shift), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (!same_vlist) {
        if (vlist) h->h_vlist.assign(vlist, vlist + n_vlist);
        else h->h_vlist.clear();
    }
    h->n_vertices = n_vertices;
    h->has_vlist = vlist != nullptr;
    h->n_vlist = n_vlist;
    if (same_shape) {
        h->slots_stale = true;
    } else {
        h->order_valid = false;
        h->warm_valid = false;
    }
    h->have_moving = true;
    return ICP_OK;
}

int icp_set_matrices(icp_handle* h, const float mw_align[16], const float mw_base[16])
{
    if (!h || !mw_align || !mw_base) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    float both[32];
    memcpy(both, mw_align, 64);
    memcpy(both + 16, mw_base, 64);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_state, both, sizeof(both), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_matrices = true;
    h->pose_dirty = true;
    return ICP_OK;
}

int icp_comm_unique_id(void* id128)
{
    if (!id128) return ICP_ERR_BAD_ARG;
    if (!g_nccl.load(g_create_error, sizeof(g_create_error))) return ICP_ERR_NCCL;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, ICP_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, 128);
    return ICP_OK;
}

int icp_comm_init(icp_handle* h, const void* id128)
{
    if (!h || !id128) return ICP_ERR_BAD_ARG;
    if (h->prm.world_size == 1) return ICP_OK;
    if (!g_nccl.load(h->err, sizeof(h->err))) return ICP_ERR_NCCL;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCCL_TRY(h, g_nccl.CommInitRank(&h->comm, h->prm.world_size, id, h->prm.rank));
    return ICP_OK;
}

int icp_query_count(const icp_handle* h, int64_t* n)
{
    if (!h || !n) return ICP_ERR_BAD_ARG;
    *n = query_count(h);
    return ICP_OK;
}

int icp_nn(icp_handle* h, int32_t* idx, float* d2)
{
    int rc = check_ready(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    const long long n_q = query_count(h);
    int latched = 0;
    rc = suspend_stop(h, &latched);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    rc = enqueue_nn(h, n_q);
    if (rc) { restore_stop(h, latched); return rc; }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    if (n_q > 0 && (idx || d2)) {
        int* d_idx = nullptr; float* d_d2 = nullptr;
        if (idx) CUDA_TRY(h, cudaMalloc((void**)&d_idx, (size_t)n_q * sizeof(int)));
        if (d2) CUDA_TRY(h, cudaMalloc((void**)&d_d2, (size_t)n_q * sizeof(float)));
        k_unpack_keys<<<(unsigned)((n_q + 255) / 256), 256, 0, h->stream>>>(h->d_keys, uses_slots(h) ? h->d_pos_slot : nullptr, n_q, d_idx, d_d2);
        h->launches++;
        if (idx) CUDA_TRY(h, cudaMemcpyAsync(idx, d_idx, (size_t)n_q * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        if (d2) CUDA_TRY(h, cudaMemcpyAsync(d2, d_d2, (size_t)n_q * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (d_idx) cudaFree(d_idx);
        if (d_d2) cudaFree(d_d2);
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_ms = ms;
    collect_profile(h);
    DevState st;
    rc = read_state(h, &st);
    if (rc) return rc;
    rc = device_error(h, st);
    if (rc) { clear_device_error(h); return rc; }
    return restore_stop(h, latched);
}

static int make_pairs_body(icp_handle* h, int calc_stats, double* A, double* B, int64_t capacity,
                           int64_t* pairs_out, double d_stats[2]);

int icp_make_pairs(icp_handle* h, int calc_stats, double* A, double* B, int64_t capacity,
                   int64_t* pairs_out, double d_stats[2])
{
    int rc = check_ready(h);
    if (rc) return rc;
    if (h->prm.world_size > 1 && h->prm.shard_mode == ICP_SHARD_BASE)
        return fail(h, ICP_ERR_STATE, "icp_make_pairs materialises pairs on one GPU; use icp_step/icp_run with sharded bases");
    if (!pairs_out) return fail(h, ICP_ERR_BAD_ARG, "icp_make_pairs: pairs_out is null");
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    int latched = 0;
    rc = suspend_stop(h, &latched);
    if (rc) return rc;
    rc = make_pairs_body(h, calc_stats, A, B, capacity, pairs_out, d_stats);
    const int rc2 = restore_stop(h, latched);
    return rc ? rc : rc2;
}

static int make_pairs_body(icp_handle* h, int calc_stats, double* A, double* B, int64_t capacity,
                           int64_t* pairs_out, double d_stats[2])
{
    int rc;
    const long long n_q = query_count(h);
    rc = enqueue_nn(h, n_q);
    if (rc) return rc;
    const int nb = (int)std::max<long long>(1, (n_q + EMIT_THREADS - 1) / EMIT_THREADS);
    if (n_q > h->flags_capacity || !h->d_flags) {
        if (h->d_flags) cudaFree(h->d_flags);
        if (h->d_block_counts) cudaFree(h->d_block_counts);
        h->d_flags = nullptr; h->d_block_counts = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_flags, (size_t)std::max<long long>(n_q, 1)));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_block_counts, (size_t)nb * sizeof(long long)));
        h->flags_capacity = std::max<long long>(n_q, 1);
    }
    float* qx = h->d_q; float* qy = h->d_q + h->q_capacity; float* qz = h->d_q + 2 * h->q_capacity;
    const int stride = h->prm.stride > 1 ? h->prm.stride : 1;
    const int* vl = h->has_vlist ? h->d_vlist : nullptr;
    const int use_sel = h->prm.reject_fraction > 0 ? 1 : 0;
    const uint32_t* pos_slot = uses_slots(h) ? h->d_pos_slot : nullptr;
    k_pairs_count<<<nb, EMIT_THREADS, 0, h->stream>>>(h->d_state, h->d_mov, vl, stride, n_q, qx, qy, qz,
                                                      h->d_keys, h->d_base4, h->surface ? h->d_tris4 : nullptr, h->base_off, h->n_local,
                                                      h->prm.thresh, use_sel, pos_slot, h->d_flags, h->d_block_counts);
    k_scan_counts<<<1, 32, 0, h->stream>>>(h->d_block_counts, nb, h->d_total);
    h->launches += 2;
    long long K = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&K, h->d_total, sizeof(K), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    DevState st;
    rc = read_state(h, &st);
    if (rc) return rc;
    rc = device_error(h, st);
    if (rc) { clear_device_error(h); return rc; }
    *pairs_out = K;
    if (!A && !B) return ICP_OK;
    if (!A || !B) return fail(h, ICP_ERR_BAD_ARG, "icp_make_pairs: A and B must both be given");
    if (capacity < K) return fail(h, ICP_ERR_BAD_ARG, "icp_make_pairs: capacity %lld < %lld pairs", (long long)capacity, K);
    if (K > 0) {
        rc = ensure(h, h->d_pairs, h->pairs_capacity, 7 * K + 2);
        if (rc) return rc;
        double* dA = h->d_pairs; double* dB = dA + 3 * K; double* dD = dB + 3 * K; double* dS = dD + K;
        k_pairs_emit<<<nb, EMIT_THREADS, 0, h->stream>>>(h->d_state, h->d_mov, vl, stride, n_q, qx, qy, qz,
                                                         h->d_keys, h->d_base4, h->surface ? h->d_tris4 : nullptr, h->base_off, h->n_local,
                                                         h->prm.thresh, use_sel, pos_slot, h->d_block_counts, K, dA, dB, dD);
        h->launches++;
        // A, B have row stride K on the device; the caller's have row stride `capacity`
        CUDA_TRY(h, cudaMemcpy2DAsync(A, (size_t)capacity * sizeof(double), dA, (size_t)K * sizeof(double),
                                      (size_t)K * sizeof(double), 3, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpy2DAsync(B, (size_t)capacity * sizeof(double), dB, (size_t)K * sizeof(double),
                                      (size_t)K * sizeof(double), 3, cudaMemcpyDeviceToHost, h->stream));
        if (calc_stats && d_stats) {
            k_dist_stats<<<1, 1024, 0, h->stream>>>(dD, K, dS);
            h->launches++;
            CUDA_TRY(h, cudaMemcpyAsync(d_stats, dS, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    } else if (calc_stats && d_stats) {
        d_stats[0] = d_stats[1] = NAN;                 // np.mean([]) -> nan
    }
    CUDA_TRY(h, cudaGetLastError());
    return ICP_OK;
}

int icp_solve_pairs(icp_handle* h, const double* v0, const double* v1, int64_t K, int scale, double M_out[16])
{
    if (!h || !v0 || !v1 || !M_out || K < 0) return h ? fail(h, ICP_ERR_BAD_ARG, "icp_solve_pairs: bad argument") : ICP_ERR_BAD_ARG;
    if (K < 3) return fail(h, ICP_ERR_TOO_FEW_PAIRS, "input arrays are of wrong shape or type");
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    int rc = ensure(h, h->d_v, h->v_capacity, 6 * K);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_v, v0, (size_t)3 * K * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_v + 3 * K, v1, (size_t)3 * K * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // moments about the first source point; the handle's own shift is restored afterwards
    DevState st;
    rc = read_state(h, &st);
    if (rc) return rc;
    const double keep_shift[3] = {st.shift[0], st.shift[1], st.shift[2]};
    const double shift[3] = {v0[0], v0[K], v0[2 * K]};
    CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, shift), shift, sizeof(shift), cudaMemcpyHostToDevice, h->stream));
    const int g = moments_grid(h, K);
    if (g > h->partials_capacity) {
        if (h->d_partials) cudaFree(h->d_partials);
        h->d_partials = nullptr;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_partials, (size_t)g * MOM * sizeof(double)));
        h->partials_capacity = g;
    }
    k_pairs_moments<<<g, MOM_THREADS, 0, h->stream>>>(h->d_v, h->d_v + 3 * K, K, h->d_state, h->d_partials);
    k_reduce_moments<<<1, MOM * 32, 0, h->stream>>>(nullptr, h->d_partials, g, h->d_moments);
    k_solve_only<<<1, 32, 0, h->stream>>>(h->d_state, h->d_moments, scale ? 1 : 0);
    h->launches += 3;
    CUDA_TRY(h, cudaGetLastError());
    rc = read_state(h, &st);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, shift), keep_shift, sizeof(keep_shift), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    rc = device_error(h, st);
    if (rc) { clear_device_error(h); return rc; }
    memcpy(M_out, st.M, sizeof(st.M));
    return ICP_OK;
}

int icp_run(icp_handle* h, int32_t max_iters, int32_t* iters_done, int32_t* converged, icp_iter_stats* per_iter)
{
    int rc = check_ready(h);
    if (rc) return rc;
    if (max_iters < 0) return fail(h, ICP_ERR_BAD_ARG, "icp_run: max_iters < 0");
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    DevState st0;
    rc = read_state(h, &st0);
    if (rc) return rc;
    rc = ensure_stats(h, st0.n_iter + max_iters);
    if (rc) return rc;
    const long long n_q = query_count(h);
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    for (int it = 0; it < max_iters; ++it) {
        rc = enqueue_iteration(h, n_q);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    DevState st;
    rc = read_state(h, &st);
    if (rc) return rc;
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->last_ms = ms;
    collect_profile(h);
    const int done = st.n_iter - st0.n_iter;
    if (iters_done) *iters_done = done;
    if (converged) *converged = (st.converged && !st.error) ? 1 : 0;
    if (per_iter && done > 0) {
        CUDA_TRY(h, cudaMemcpy(per_iter, h->d_stats + st0.n_iter, (size_t)done * sizeof(icp_iter_stats), cudaMemcpyDeviceToHost));
    }
    rc = device_error(h, st);
    if (rc) { clear_device_error(h); return rc; }
    return ICP_OK;
}

int icp_step(icp_handle* h, double M_out[16], icp_iter_stats* stats)
{
    if (!h) return ICP_ERR_BAD_ARG;
    icp_iter_stats s;
    memset(&s, 0, sizeof(s));
    int done = 0, conv = 0;
    // a single step ignores an earlier "converged" latch: the caller decides when to stop
    DevState st;
    int rc = check_ready(h);
    if (rc) return rc;
    rc = read_state(h, &st);
    if (rc) return rc;
    if (st.converged) {
        int zero = 0;
        CUDA_TRY(h, cudaMemcpyAsync((char*)h->d_state + offsetof(DevState, converged), &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    rc = icp_run(h, 1, &done, &conv, &s);
    if (rc) return rc;
    if (done != 1) return fail(h, ICP_ERR_STATE, "icp_step: iteration did not execute");
    if (M_out) memcpy(M_out, s.M, sizeof(s.M));
    if (stats) *stats = s;
    return ICP_OK;
}

int icp_reset_convergence(icp_handle* h)
{
    if (!h) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    DevState st;
    int rc = read_state(h, &st);
    if (rc) return rc;
    for (int i = 0; i < 5; ++i) st.ring[i] = h->prm.target_d * 2;   // icp_align.py:93
    st.n_iter = 0; st.converged = 0; st.error = 0;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_state, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ICP_OK;
}

int icp_get_matrix_world(icp_handle* h, float out[16])
{
    if (!h || !out) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    DevState st;
    int rc = read_state(h, &st);
    if (rc) return rc;
    memcpy(out, st.mw_align, 64);
    return ICP_OK;
}

int icp_last_elapsed_ms(const icp_handle* h, double* ms)
{
    if (!h || !ms) return ICP_ERR_BAD_ARG;
    *ms = h->last_ms;
    return ICP_OK;
}

int icp_set_profiling(icp_handle* h, int enabled)
{
    if (!h) return ICP_ERR_BAD_ARG;
    h->profiling = enabled != 0;
    for (int w = 0; w < 5; ++w) { h->ev_total_ms[w] = 0; h->ev_launches[w] = 0; h->ev_used[w] = 0; }
    return ICP_OK;
}

int icp_kernel_time_ms(const icp_handle* h, int which, double* total_ms, int64_t* launches)
{
    if (!h || which < 0 || which > 4) return ICP_ERR_BAD_ARG;
    if (total_ms) *total_ms = h->ev_total_ms[which];
    if (launches) *launches = h->ev_launches[which];
    return ICP_OK;
}

int icp_launch_count(const icp_handle* h, int64_t* launches)
{
    if (!h || !launches) return ICP_ERR_BAD_ARG;
    *launches = h->launches;
    return ICP_OK;
}

int icp_synchronize(icp_handle* h)
{
    if (!h) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return ICP_OK;
}

static int ensure_scratch(icp_handle* h, size_t bytes)
{
    if (h->scratch_bytes >= bytes) return ICP_OK;
    if (h->d_scratch) cudaFree(h->d_scratch);
    h->d_scratch = nullptr; h->scratch_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&h->d_scratch, bytes));
    h->scratch_bytes = bytes;
    return ICP_OK;
}

int icp_measure_fp32_peak(icp_handle* h, double* tflops)
{
    if (!h || !tflops) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    int rc = ensure_scratch(h, 1 << 20);
    if (rc) return rc;
    const int iters = 1 << 15, blocks = h->sm_count * 8, threads = 256;
    double best = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 4; ++rep) {
            CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
            if (variant == 0) k_fma_peak<<<blocks, threads, 0, h->stream>>>((float*)h->d_scratch, iters);
            else k_fma2_peak<<<blocks, threads, 0, h->stream>>>((float*)h->d_scratch, iters);
            CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            const double fl = 2.0 * (variant ? 2.0 : 1.0) * 8.0 * iters * (double)blocks * threads;
            if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
    }
    h->launches += 8;
    *tflops = best;
    return ICP_OK;
}

int icp_measure_copy_bandwidth(icp_handle* h, double* gbs)
{
    if (!h || !gbs) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    const size_t half = (size_t)1 << 30;      // 1 GiB read + 1 GiB write
    int rc = ensure_scratch(h, 2 * half);
    if (rc) return rc;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        k_copy<<<h->sm_count * 16, 512, 0, h->stream>>>((const float4*)h->d_scratch, (float4*)((char*)h->d_scratch + half), (long long)(half / 16));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        if (rep > 0) best = std::max(best, 2.0 * half / (ms * 1e-3) / 1e9);
    }
    h->launches += 5;
    *gbs = best;
    return ICP_OK;
}

int icp_flush_l2(icp_handle* h)
{
    if (!h) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    const size_t bytes = (size_t)256 << 20;   // 2x the 126 MB L2
    int rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemsetAsync(h->d_scratch, 0, bytes, h->stream));
    return ICP_OK;
}

int icp_debug_counters(icp_handle* h, uint64_t out[4], int reset)
{
    if (!h || !out) return ICP_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->prm.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    unsigned long long v[4] = {0, 0, 0, 0};
    CUDA_TRY(h, cudaMemcpyFromSymbol(v, g_bvh_counters, sizeof(v)));
    for (int i = 0; i < 4; ++i) out[i] = v[i];
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        CUDA_TRY(h, cudaMemcpyToSymbol(g_bvh_counters, z, sizeof(z)));
    }
    return ICP_OK;
}

int icp_alloc_pinned(size_t bytes, void** out)
{
    if (!out) return ICP_ERR_BAD_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, ICP_ERR_OOM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return ICP_OK;
}

int icp_free_pinned(void* ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return ICP_OK;
}

}  // extern "C"
