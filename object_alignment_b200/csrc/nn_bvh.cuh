// nn_bvh.cuh -- row N2 of SURVEY.md section 8f: a spatially-indexed EXACT nearest-vertex search
// (the "KD-tree equivalent" of BASELINE.json's north_star) that returns bit-identical keys to
// the brute-force kernel K2, in O(M log N) instead of O(M N).
//
// Index (built once per icp_set_base, like the reference's BVHTree.FromObject at
// operators/icp_align.py:53):
//   * base vertices sorted by 63-bit Morton code of their position in the shard's bounding box
//   * leaves = runs of BVH_LEAF consecutive sorted vertices, stored as float4(x, y, z, index bits)
//   * a binary radix tree (LBVH, Karras 2012) of axis-aligned boxes over those runs; a node keeps
//     both children's boxes and references, so one node visit is one 64-byte read
// Query (one thread per query, depth-first, near child first):
//   * a subtree is skipped only when its box lower bound, shrunk by 2e-6 relative, is still
//     STRICTLY greater than the best S4 distance found so far.  fp32 rounding of the bound
//     (<= 6 ulp) and of S4 (<= 5 ulp) is far inside that margin, so no vertex that could win OR
//     TIE is ever skipped, and the winner is min over the same uint64 keys K2 uses
//     (float_bits(d2) << 32 | global index): identical result, ties to the lowest index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"
#include "icp_state.cuh"

namespace icp {

#ifndef ICP_BVH_LEAF
#define ICP_BVH_LEAF 16      // vertices per leaf run; measured 8 -> 16: C3 +2 %, C4 (10M base) +11 %
#endif
constexpr int BVH_LEAF = ICP_BVH_LEAF;
constexpr int BVH_STACK = 96;   // >= 63 code bits + 32 tie-break bits of radix-tree depth

__device__ __forceinline__ uint64_t morton_expand21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// bounding box of the shard: per-block partials, then one block
__global__ void k_bbox_partial(const float4* __restrict__ pts, long long n, float* __restrict__ part)
{
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
    }
    __shared__ float sh[6][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
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
__global__ void k_bbox_final(const float* __restrict__ part, int n_part, float* __restrict__ bbox)
{
    // one warp per component (launch with 6 * 32 threads)
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= 6) return;
    const float inf = __int_as_float(0x7f800000);
    float v = (k < 3) ? inf : -inf;
    for (int i = lane; i < n_part; i += 32) v = (k < 3) ? fminf(v, part[i * 6 + k]) : fmaxf(v, part[i * 6 + k]);
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (k < 3) ? fminf(v, w) : fmaxf(v, w);
    }
    if (lane == 0) bbox[k] = v;
}

// Position of a grid cell (21 bits per axis) along the space-filling curve that orders base vertices, triangles
// and query slots.  Default: the HILBERT curve (Skilling's transpose algorithm) -- consecutive positions are
// always adjacent cells, so 32 consecutive slots (a warp's patch) or 32 / 1024 consecutive vertices (a tile) never
// straddle one of the Morton curve's jumps across the object (the brute-force kernels' 1024-vertex tiles keep the Morton order: `morton` argument).  Any order is valid (the searches are exact
// whatever the order); -DICP_CURVE_MORTON keeps the plain bit interleave for A/B runs.  Both are radix-tree
// friendly: a common prefix of the code is a sub-cube.
__device__ __forceinline__ uint64_t curve_code21(uint32_t gx, uint32_t gy, uint32_t gz, bool morton = false)
{
#ifdef ICP_CURVE_MORTON
    morton = true;
#endif
    if (morton) return morton_expand21(gx) | (morton_expand21(gy) << 1) | (morton_expand21(gz) << 2);
    {
    uint32_t X[3] = {gx, gy, gz};
    for (uint32_t Q = 1u << 20; Q > 1u; Q >>= 1) {          // undo the excess work of the Gray-code walk
        const uint32_t P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];                             // Gray encode
    uint32_t t = 0u;
    for (uint32_t Q = 1u << 20; Q > 1u; Q >>= 1) if (X[2] & Q) t ^= Q - 1u;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (morton_expand21(X[0]) << 2) | (morton_expand21(X[1]) << 1) | morton_expand21(X[2]);
    }
}

__global__ void k_morton_codes(const float4* __restrict__ pts, long long n, const float* __restrict__ bbox,
                               unsigned long long* __restrict__ codes, uint32_t* __restrict__ order, int morton = 0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const float c[3] = {p.x, p.y, p.z};
    uint32_t g[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ext = bbox[3 + k] - bbox[k];
        float t = ext > 0.0f ? (c[k] - bbox[k]) / ext : 0.0f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        g[k] = (uint32_t)fminf(t * 2097152.0f, 2097151.0f);
    }
    codes[i] = curve_code21(g[0], g[1], g[2], morton != 0);
    order[i] = (uint32_t)i;
}

// sorted, leaf-padded vertex array: float4(x, y, z, local index bits); padding is +inf / 0xFFFFFFFF
__global__ void k_bvh_gather(const float4* __restrict__ pts, const uint32_t* __restrict__ order, long long n,
                             long long n_padded, float4* __restrict__ sorted)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    const float inf = __int_as_float(0x7f800000);
    if (i < n) {
        const uint32_t j = order[i];
        const float4 p = pts[j];
        sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float(j));
    } else {
        sorted[i] = make_float4(inf, inf, inf, __uint_as_float(0xFFFFFFFFu));
    }
}

// ---------------------------------------------------------------------------------------------
// Topology: a binary radix tree (Karras 2012) over the RUNS of BVH_LEAF Morton-sorted vertices,
// keyed by the Morton code of each run's first vertex (ties broken by run id).  Any binary tree
// over the sorted runs is valid for an exact search because every box is fitted to the actual
// vertices; the radix splits just make the boxes compact (octree-cell shaped).
// A node stores its two CHILDREN's boxes and references, so one visit is one 64-byte read.
// Child reference: >= 0 internal node id, < 0 leaf run (~ref).
// ---------------------------------------------------------------------------------------------
struct __align__(16) BvhNode {          // 64 bytes
    float box[12];                      // left lo.xyz, left hi.xyz, right lo.xyz, right hi.xyz
    int left, right;
    int parent;                         // build only
    unsigned int arrived;               // build only
};

__global__ void k_run_keys(const unsigned long long* __restrict__ sorted_codes, long long n, int n_runs,
                           unsigned long long* __restrict__ run_keys)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    run_keys[r] = sorted_codes[(long long)r * BVH_LEAF];
}

__global__ void k_run_keys_stride(const unsigned long long* __restrict__ sorted_codes, int n_runs, int run,
                                  unsigned long long* __restrict__ run_keys)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_runs) run_keys[r] = sorted_codes[(long long)r * run];
}

__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ k, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = k[i], b = k[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_lbvh_topology(const unsigned long long* __restrict__ keys, int n_runs,
                                BvhNode* __restrict__ nodes, int* __restrict__ leaf_parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_runs - 1) return;
    const int d = (lbvh_delta(keys, n_runs, i, i + 1) - lbvh_delta(keys, n_runs, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n_runs, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n_runs, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n_runs, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n_runs, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n_runs, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? ~gamma : gamma;
    const int right = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
    nodes[i].left = left;
    nodes[i].right = right;
    nodes[i].arrived = 0;
    if (i == 0) nodes[0].parent = -1;
    if (left < 0) leaf_parent[~left] = i; else nodes[left].parent = i;
    if (right < 0) leaf_parent[~right] = i; else nodes[right].parent = i;
}

// one thread per leaf run: fit the run's box, then climb; the second child to arrive at a node
// (atomic counter) has both boxes visible and carries their union further up
__global__ void k_lbvh_fit(const float4* __restrict__ sorted, long long n, int n_runs,
                           BvhNode* __restrict__ nodes, const int* __restrict__ leaf_parent)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs || n_runs < 2) return;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = 0; i < BVH_LEAF; ++i) {
        const long long j = (long long)r * BVH_LEAF + i;
        if (j >= n) break;
        const float4 p = sorted[j];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
    }
    int child = ~r;
    int node = leaf_parent[r];
    while (node >= 0) {
        BvhNode* nd = nodes + node;
        const int off = (nd->left == child) ? 0 : 6;
        volatile float* vb = nd->box;
        for (int k = 0; k < 3; ++k) { vb[off + k] = lo[k]; vb[off + 3 + k] = hi[k]; }
        __threadfence();
        if (atomicAdd(&nd->arrived, 1u) == 0u) return;       // the sibling will continue
        __threadfence();
        const int o2 = 6 - off;
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], vb[o2 + k]); hi[k] = fmaxf(hi[k], vb[o2 + 3 + k]); }
        child = node;
        node = nd->parent;
    }
}

__device__ __forceinline__ float box_lower_bound(float qx, float qy, float qz, float lox, float loy, float loz,
                                                 float hix, float hiy, float hiz)
{
    const float tx = fmaxf(fmaxf(lox - qx, qx - hix), 0.0f);
    const float ty = fmaxf(fmaxf(loy - qy, qy - hiy), 0.0f);
    const float tz = fmaxf(fmaxf(loz - qz, qz - hiz), 0.0f);
    return fmaf(tz, tz, fmaf(ty, ty, tx * tx));
}

constexpr float BVH_SHRINK = 0.999998f;   // 2e-6 relative: > 11 ulp of fp32, see the header note

__device__ unsigned long long g_bvh_counters[4];   // debug: inner visits, leaf visits, pushes, queries

template <bool COUNT>
__global__ void __launch_bounds__(128)
k_nn_bvh(const BvhNode* __restrict__ nodes, const float4* __restrict__ sorted, int n_runs,
         const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
         long long n_queries, unsigned long long* __restrict__ keys, uint32_t index_offset,
         const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_queries) return;
    const float x = qx[p], y = qy[p], z = qz[p];
    const float inf = __int_as_float(0x7f800000);
    // the slot's key is "none" or a warm-start candidate (a real base vertex under the new pose)
    unsigned long long best = keys[p];
    float best_d = (best == KEY_NONE) ? inf : __uint_as_float((uint32_t)(best >> 32));

    int stack_ref[BVH_STACK];
    float stack_bound[BVH_STACK];
    int sp = 0;
    int cur = (n_runs > 1) ? 0 : ~0;
    unsigned int c_inner = 0, c_leaf = 0, c_push = 0;
    const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
    while (true) {
        if (cur < 0) {
            // leaf run: BVH_LEAF consecutive float4 (128 contiguous bytes)
            const float4* lp = sorted + (long long)(~cur) * BVH_LEAF;
            if (COUNT) ++c_leaf;
#pragma unroll
            for (int i = 0; i < BVH_LEAF; ++i) {
                const float4 b = lp[i];
                const float d = s4_d2(x, y, z, b.x, b.y, b.z);
                if (d < inf) {
                    const unsigned long long k =
                        ((unsigned long long)__float_as_uint(d) << 32) |
                        (unsigned long long)(index_offset + __float_as_uint(b.w));
                    if (k < best) { best = k; best_d = d; }
                }
            }
        } else {
            if (COUNT) ++c_inner;
            const float4 n0 = nodes4[4 * (long long)cur], n1 = nodes4[4 * (long long)cur + 1];
            const float4 n2 = nodes4[4 * (long long)cur + 2], n3 = nodes4[4 * (long long)cur + 3];
            const float bl = box_lower_bound(x, y, z, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
            const float br = box_lower_bound(x, y, z, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
            const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
            const bool left_first = bl <= br;
            const int near = left_first ? left : right, far = left_first ? right : left;
            const float bn = left_first ? bl : br, bf = left_first ? br : bl;
            const bool take_near = !(bn * BVH_SHRINK > best_d) && bn < inf;
            const bool take_far = !(bf * BVH_SHRINK > best_d) && bf < inf;
            if (take_far && sp < BVH_STACK) {
                stack_ref[sp] = far; stack_bound[sp] = bf; ++sp;
                if (COUNT) ++c_push;
            }
            if (take_near) { cur = near; continue; }
        }
        // pop the next subtree that can still contain a winner or a tie
        bool found = false;
        while (sp > 0) {
            --sp;
            if (!(stack_bound[sp] * BVH_SHRINK > best_d)) { cur = stack_ref[sp]; found = true; break; }
        }
        if (!found) break;
    }
    if (best != KEY_NONE) atomicMin(&keys[p], best);
    if (COUNT) {
        atomicAdd(&g_bvh_counters[0], (unsigned long long)c_inner);
        atomicAdd(&g_bvh_counters[1], (unsigned long long)c_leaf);
        atomicAdd(&g_bvh_counters[2], (unsigned long long)c_push);
        atomicAdd(&g_bvh_counters[3], 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-packet variant.  Query slots are in Morton order of the moving vertices, so the 32
// queries of a warp are spatial neighbours and want nearly the same nodes.  The warp walks the
// tree TOGETHER: one shared stack, node and leaf loads are warp-uniform (one transaction,
// broadcast), every lane evaluates its own bounds and distances, and a subtree is entered when
// ANY lane still needs it -- so there is no divergence.  Each lane's own prune test is exactly
// the single-thread kernel's, hence the same exact result.
//
// FUSED (one GPU, no percentile): the kernel also does K1 (q = imx2 @ (mx1 @ a), warm start from
// the previous winner) in its prologue and K3+K4 (world distance, thresh, target, fp64 moment
// partials) in its epilogue, so an ICP iteration is this kernel + one single-CTA solve.
// ---------------------------------------------------------------------------------------------
#ifndef ICP_PKT_THREADS
#define ICP_PKT_THREADS 128
#endif
// 4 warps per CTA.  A CTA's registers are only released when its slowest warp has finished its walk
// (ncu: 20 % of the surface kernel's samples wait in the epilogue barrier); a one-warp build
// (-DICP_PKT_THREADS=32) was 10 % faster on the surface walk but faulted in the vertex walk
// (compute-sanitizer: out-of-range node reference) and is NOT supported until that is understood.
constexpr int PKT_THREADS = ICP_PKT_THREADS;
#ifndef ICP_BVH_RESIDENT
#define ICP_BVH_RESIDENT 1280      // resident threads per SM the walk is compiled for (<= 48 registers; the fused
                                   // kernel needs 46).  Measured at C3: 1280 -> 0.44 ms, 1536 -> 0.46, 2048 -> 0.48 (spills)
#endif

template <bool FUSED, bool COUNT>
__global__ void __launch_bounds__(PKT_THREADS, ICP_BVH_RESIDENT / PKT_THREADS)
k_nn_bvh_packet(const BvhNode* __restrict__ nodes, const float4* __restrict__ sorted, int n_runs,
                const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                const DevState* __restrict__ st, const float4* __restrict__ mov_slots,
                const float4* __restrict__ base4, long long base_off, long long n_local, double thresh,
                int warm, long long n_queries, unsigned long long* __restrict__ keys,
                uint32_t index_offset, double* __restrict__ partials)
{
    if (st->converged) return;
    __shared__ int s_ref[PKT_THREADS / 32][BVH_STACK];
    __shared__ float s_bound[PKT_THREADS / 32][BVH_STACK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < n_queries;
    const float inf = __int_as_float(0x7f800000);

    float x = 0.f, y = 0.f, z = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    unsigned long long best = KEY_NONE;
    if (active) {
        if (FUSED) {
            const float4 a = mov_slots[t];
            ax = a.x; ay = a.y; az = a.z;
            float tx, ty, tz;
            s1_matvec(st->mw_align, ax, ay, az, tx, ty, tz);
            s1_matvec(st->imx2, tx, ty, tz, x, y, z);
            if (warm) {
                const unsigned long long old = keys[t];
                const long long j = (long long)(uint32_t)(old & 0xFFFFFFFFull) - base_off;
                if (old != KEY_NONE && j >= 0 && j < n_local) {
                    const float4 b = base4[j];
                    const float d = s4_d2(x, y, z, b.x, b.y, b.z);
                    if (d < inf) best = ((unsigned long long)__float_as_uint(d) << 32) | (old & 0xFFFFFFFFull);
                }
            }
        } else {
            x = qx[t]; y = qy[t]; z = qz[t];
            best = keys[t];
        }
    }
    float best_d = (best == KEY_NONE) ? inf : __uint_as_float((uint32_t)(best >> 32));

    int sp = 0;
    int cur = (n_runs > 1) ? 0 : ~0;
    unsigned int c_inner = 0, c_leaf = 0, c_push = 0;
    const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
    if (n_runs > 0) {
        while (true) {
            if (cur < 0) {
                const float4* lp = sorted + (long long)(~cur) * BVH_LEAF;
                if (COUNT) ++c_leaf;
                // branch-free: the smallest 64-bit key of the run, then one compare with the lane's best.
                // d >= 0, so its bits order like its value; padding (+inf) and NaN are rejected on the
                // minimum.  The winner's coordinates are fetched once, in the epilogue.
                unsigned long long kmin = KEY_NONE;
#pragma unroll
                for (int i = 0; i < BVH_LEAF; ++i) {
                    const float4 b = __ldg(lp + i);                   // warp-uniform address
                    const float d = s4_d2(x, y, z, b.x, b.y, b.z);
                    const unsigned long long k =
                        ((unsigned long long)__float_as_uint(d) << 32) |
                        (unsigned long long)(index_offset + __float_as_uint(b.w));
                    kmin = k < kmin ? k : kmin;
                }
                if (kmin < best && (uint32_t)(kmin >> 32) < 0x7f800000u) {
                    best = kmin; best_d = __uint_as_float((uint32_t)(kmin >> 32));
                }
            } else {
                if (COUNT) ++c_inner;
                const float4 n0 = __ldg(nodes4 + 4 * (long long)cur), n1 = __ldg(nodes4 + 4 * (long long)cur + 1);
                const float4 n2 = __ldg(nodes4 + 4 * (long long)cur + 2), n3 = __ldg(nodes4 + 4 * (long long)cur + 3);
                const float bl = box_lower_bound(x, y, z, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
                const float br = box_lower_bound(x, y, z, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
                const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
                const bool wl = active && !(bl * BVH_SHRINK > best_d) && bl < inf;
                const bool wr = active && !(br * BVH_SHRINK > best_d) && br < inf;
                const unsigned ml = __ballot_sync(0xffffffffu, wl), mr = __ballot_sync(0xffffffffu, wr);
                if (ml != 0u && mr != 0u) {
                    // enter first the child most lanes would enter first; park the other with the
                    // smallest bound any interested lane has for it
                    const unsigned pl = __ballot_sync(0xffffffffu, wl && (!wr || bl <= br));
                    const bool left_first = 2 * __popc(pl) >= __popc(ml | mr);
                    const float fb = left_first ? (wr ? br : inf) : (wl ? bl : inf);
                    const unsigned fmin = __reduce_min_sync(0xffffffffu, __float_as_uint(fb));
                    if (lane == 0 && sp < BVH_STACK) {
                        s_ref[warp][sp] = left_first ? right : left;
                        s_bound[warp][sp] = __uint_as_float(fmin);
                    }
                    if (sp < BVH_STACK) ++sp;
                    if (COUNT) ++c_push;
                    __syncwarp();
                    cur = left_first ? left : right;
                    continue;
                }
                if (ml != 0u) { cur = left; continue; }
                if (mr != 0u) { cur = right; continue; }
            }
            // pop the next parked subtree that some lane can still use
            bool found = false;
            while (sp > 0) {
                --sp;
                const float pb = s_bound[warp][sp];
                const int pr = s_ref[warp][sp];
                if (__ballot_sync(0xffffffffu, active && !(pb * BVH_SHRINK > best_d)) != 0u) {
                    cur = pr; found = true; break;
                }
            }
            __syncwarp();
            if (!found) break;
        }
    }
    if (active) keys[t] = best;
    if (COUNT && lane == 0) {
        atomicAdd(&g_bvh_counters[0], (unsigned long long)c_inner);
        atomicAdd(&g_bvh_counters[1], (unsigned long long)c_leaf);
        atomicAdd(&g_bvh_counters[2], (unsigned long long)c_push);
        atomicAdd(&g_bvh_counters[3], 1ull);
    }
    if (FUSED) {
        // K3 + K4 for this thread's query (functions/general.py:299-306, :160-181)
        bool keep = false;
        float bx = 0.f, by = 0.f, bz = 0.f;
        double dist = 0.0;
        if (active && best != KEY_NONE) {
            const float4 w = base4[(long long)(uint32_t)(best & 0xFFFFFFFFull) - base_off];   // the winner
            float wqx, wqy, wqz, w1x, w1y, w1z;
            s1_matvec(st->mw_base, x, y, z, wqx, wqy, wqz);
            s1_matvec(st->mw_base, w.x, w.y, w.z, w1x, w1y, w1z);
            dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
            if (dist < thresh) {
                s1_matvec(st->imx1, w1x, w1y, w1z, bx, by, bz);
                keep = true;
            }
        }
        fused_pair_moments<PKT_THREADS>(keep, ax, ay, az, bx, by, bz, dist, st->shift[0], st->shift[1], st->shift[2], partials);
    }
}

}  // namespace icp
