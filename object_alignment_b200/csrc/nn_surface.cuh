// nn_surface.cuh -- row N1 of SURVEY.md section 8f: closest point on the TRIANGULATED SURFACE of the
// base mesh, i.e. what the reference really queries: BVHTree.FromObject(base_obj, depsgraph)
// (operators/icp_align.py:53) + base_bvh.find_nearest(co_find) (functions/general.py:297).
//
// Blender's C BVH is absent from /root/reference; the arithmetic restated here is Blender 3.2's
// closest_on_tri_to_point_v3 (Ericson's region walk, all fp32, no fused multiply-adds) and
// len_squared_v3v3 -- rules S9 / S10 of DESIGN.md, shared bit for bit with oracle/nn_oracle.c.
// Ties between triangles go to the lowest triangle index (Blender's own tie-break depends on its
// tree layout and is not reproducible).
//
// Index: the same LBVH as nn_bvh.cuh, over runs of TRI_LEAF triangles sorted by the Morton code of
// their centroids; boxes are the triangles' bounds.  The walk prunes with the same strict,
// 2e-6-shrunk lower bound, so the winner is exact under S9/S10.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"
#include "nn_bvh.cuh"

namespace icp {

constexpr int TRI_LEAF = 4;

// Small CTAs for the surface walks: a CTA's registers are released only when its slowest warp has
// finished, and the surface walk's tail is long (ncu: 20 % of the samples waited in the epilogue
// barrier with 4 warps per CTA).  Two warps per CTA at <= 48 registers measured best (40 resident
// warps; one-warp CTAs are capped at 32 per SM).  The kernels are never COMPILED for fewer than 64
// threads: with __launch_bounds__(32) ptxas drops the warp syncs around the shared stack (it treats
// the CTA as one converged warp), the votes then see a diverged subset and the walk follows garbage
// references (caught by compute-sanitizer in the vertex walk).
#ifndef ICP_SURF_THREADS
#define ICP_SURF_THREADS 64
#endif
constexpr int SURF_THREADS = ICP_SURF_THREADS;
#ifndef ICP_SURF_RESIDENT
#define ICP_SURF_RESIDENT 1280     // resident threads per SM the deferred walk is compiled for (<= 48 registers, 40 B of
                                   // spills).  Measured at C3s: 64 threads / 1280 -> 1.47 ms, 32 / 1024 -> 1.58, 64 / 1024 -> 1.61
#endif
constexpr int SURF_BOUNDS = SURF_THREADS < 64 ? 64 : SURF_THREADS;
constexpr int SURF_WARPS = (SURF_THREADS + 31) / 32;

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// S9: closest point on triangle (a, b, c) to p
__device__ __forceinline__ void closest_on_tri(float px, float py, float pz, float ax, float ay, float az,
                                               float bx, float by, float bz, float cx, float cy, float cz,
                                               float& rx, float& ry, float& rz)
{
    float abx = __fsub_rn(bx, ax), aby = __fsub_rn(by, ay), abz = __fsub_rn(bz, az);
    float acx = __fsub_rn(cx, ax), acy = __fsub_rn(cy, ay), acz = __fsub_rn(cz, az);
    const float apx = __fsub_rn(px, ax), apy = __fsub_rn(py, ay), apz = __fsub_rn(pz, az);
    const float d1 = dot3_rn(abx, aby, abz, apx, apy, apz);
    const float d2 = dot3_rn(acx, acy, acz, apx, apy, apz);
    if (d1 <= 0.0f && d2 <= 0.0f) { rx = ax; ry = ay; rz = az; return; }             // vertex A
    const float bpx = __fsub_rn(px, bx), bpy = __fsub_rn(py, by), bpz = __fsub_rn(pz, bz);
    const float d3 = dot3_rn(abx, aby, abz, bpx, bpy, bpz);
    const float d4 = dot3_rn(acx, acy, acz, bpx, bpy, bpz);
    if (d3 >= 0.0f && d4 <= d3) { rx = bx; ry = by; rz = bz; return; }               // vertex B
    const float vc = __fsub_rn(__fmul_rn(d1, d4), __fmul_rn(d3, d2));
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {                                    // edge AB
        const float v = __fdiv_rn(d1, __fsub_rn(d1, d3));
        rx = __fadd_rn(ax, __fmul_rn(abx, v)); ry = __fadd_rn(ay, __fmul_rn(aby, v));
        rz = __fadd_rn(az, __fmul_rn(abz, v));
        return;
    }
    const float cpx = __fsub_rn(px, cx), cpy = __fsub_rn(py, cy), cpz = __fsub_rn(pz, cz);
    const float d5 = dot3_rn(abx, aby, abz, cpx, cpy, cpz);
    const float d6 = dot3_rn(acx, acy, acz, cpx, cpy, cpz);
    if (d6 >= 0.0f && d5 <= d6) { rx = cx; ry = cy; rz = cz; return; }               // vertex C
    const float vb = __fsub_rn(__fmul_rn(d5, d2), __fmul_rn(d1, d6));
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {                                    // edge AC
        const float w = __fdiv_rn(d2, __fsub_rn(d2, d6));
        rx = __fadd_rn(ax, __fmul_rn(acx, w)); ry = __fadd_rn(ay, __fmul_rn(acy, w));
        rz = __fadd_rn(az, __fmul_rn(acz, w));
        return;
    }
    const float va = __fsub_rn(__fmul_rn(d3, d6), __fmul_rn(d5, d4));
    const float d43 = __fsub_rn(d4, d3), d56 = __fsub_rn(d5, d6);
    if (va <= 0.0f && d43 >= 0.0f && d56 >= 0.0f) {                                  // edge BC
        const float w = __fdiv_rn(d43, __fadd_rn(d43, d56));
        rx = __fadd_rn(__fmul_rn(__fsub_rn(cx, bx), w), bx);
        ry = __fadd_rn(__fmul_rn(__fsub_rn(cy, by), w), by);
        rz = __fadd_rn(__fmul_rn(__fsub_rn(cz, bz), w), bz);
        return;
    }
    const float denom = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(va, vb), vc));           // face interior
    const float v = __fmul_rn(vb, denom), w = __fmul_rn(vc, denom);
    acx = __fmul_rn(acx, w); acy = __fmul_rn(acy, w); acz = __fmul_rn(acz, w);
    rx = __fadd_rn(__fadd_rn(ax, __fmul_rn(abx, v)), acx);
    ry = __fadd_rn(__fadd_rn(ay, __fmul_rn(aby, v)), acy);
    rz = __fadd_rn(__fadd_rn(az, __fmul_rn(abz, v)), acz);
}

// S10: squared distance as Blender's len_squared_v3v3
__device__ __forceinline__ float s10_d2(float px, float py, float pz, float rx, float ry, float rz)
{
    const float dx = __fsub_rn(rx, px), dy = __fsub_rn(ry, py), dz = __fsub_rn(rz, pz);
    return dot3_rn(dx, dy, dz, dx, dy, dz);
}

// ---- index build ------------------------------------------------------------------------------
// tri_xyz: n x 9 floats (a, b, c).  tris4: 3 float4 per triangle by ORIGINAL index (a.xyz|0, b, c).
__global__ void k_tri_pack(const float* __restrict__ tri_xyz, long long n, float4* __restrict__ tris4,
                           float4* __restrict__ centroids)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = tri_xyz + 9 * i;
    tris4[3 * i] = make_float4(t[0], t[1], t[2], 0.0f);
    tris4[3 * i + 1] = make_float4(t[3], t[4], t[5], 0.0f);
    tris4[3 * i + 2] = make_float4(t[6], t[7], t[8], 0.0f);
    centroids[i] = make_float4((t[0] + t[3] + t[6]) * (1.0f / 3.0f), (t[1] + t[4] + t[7]) * (1.0f / 3.0f),
                               (t[2] + t[5] + t[8]) * (1.0f / 3.0f), 0.0f);
}

// sorted, run-padded triangles: 3 float4 each; .w of the first holds the local triangle index
__global__ void k_tri_gather(const float4* __restrict__ tris4, const uint32_t* __restrict__ order, long long n,
                             long long n_padded, float4* __restrict__ sorted)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    const float inf = __int_as_float(0x7f800000);
    if (i < n) {
        const uint32_t j = order[i];
        float4 a = tris4[3 * (long long)j];
        a.w = __uint_as_float(j);
        sorted[3 * i] = a;
        sorted[3 * i + 1] = tris4[3 * (long long)j + 1];
        sorted[3 * i + 2] = tris4[3 * (long long)j + 2];
    } else {
        sorted[3 * i] = make_float4(inf, inf, inf, __uint_as_float(0xFFFFFFFFu));
        sorted[3 * i + 1] = make_float4(inf, inf, inf, 0.0f);
        sorted[3 * i + 2] = make_float4(inf, inf, inf, 0.0f);
    }
}

__global__ void k_lbvh_fit_tris(const float4* __restrict__ sorted, long long n, int n_runs,
                                BvhNode* __restrict__ nodes, const int* __restrict__ leaf_parent)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs || n_runs < 2) return;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = 0; i < TRI_LEAF; ++i) {
        const long long j = (long long)r * TRI_LEAF + i;
        if (j >= n) break;
        for (int c = 0; c < 3; ++c) {
            const float4 p = sorted[3 * j + c];
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
    }
    int child = ~r;
    int node = leaf_parent[r];
    while (node >= 0) {
        BvhNode* nd = nodes + node;
        const int off = (nd->left == child) ? 0 : 6;
        volatile float* vb = nd->box;
        for (int k = 0; k < 3; ++k) { vb[off + k] = lo[k]; vb[off + 3 + k] = hi[k]; }
        __threadfence();
        if (atomicAdd(&nd->arrived, 1u) == 0u) return;
        __threadfence();
        const int o2 = 6 - off;
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], vb[o2 + k]); hi[k] = fmaxf(hi[k], vb[o2 + 3 + k]); }
        child = node;
        node = nd->parent;
    }
}

// ---- the walk: one thread per query -----------------------------------------------------------
// keys[t] enters as "none" or as a warm-start candidate (previous winner re-evaluated by K1s).
__global__ void __launch_bounds__(128)
k_nn_surface(const BvhNode* __restrict__ nodes, const float4* __restrict__ sorted, int n_runs,
             const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
             long long n_queries, unsigned long long* __restrict__ keys, uint32_t index_offset,
             float coord_extent, const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_queries) return;
    const float x = qx[p], y = qy[p], z = qz[p];
    const float inf = __int_as_float(0x7f800000);
    // S9 returns a point that is within ~eta of the triangle, so its distance can undercut the
    // box bound by eta; prune in DISTANCE terms:  sqrt(bound) > sqrt(best) + eta
    const float eta = 3.9e-6f * fmaxf(fmaxf(fabsf(x), fabsf(y)), fmaxf(fabsf(z), coord_extent));   // 64 u * scale
    unsigned long long best = keys[p];
    float best_d = (best == KEY_NONE) ? inf : __uint_as_float((uint32_t)(best >> 32));
    float prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;
    int stack_ref[BVH_STACK];
    float stack_bound[BVH_STACK];
    int sp = 0;
    int cur = (n_runs > 1) ? 0 : ~0;
    const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
    while (true) {
        if (cur < 0) {
            const float4* lp = sorted + (long long)(~cur) * (TRI_LEAF * 3);
#pragma unroll
            for (int i = 0; i < TRI_LEAF; ++i) {
                const float4 a = lp[3 * i], b = lp[3 * i + 1], c = lp[3 * i + 2];
                if (a.x < inf) {
                    float rx, ry, rz;
                    closest_on_tri(x, y, z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, rx, ry, rz);
                    const float d = s10_d2(x, y, z, rx, ry, rz);
                    if (d <= best_d && d < inf) {
                        const unsigned long long k =
                            ((unsigned long long)__float_as_uint(d) << 32) |
                            (unsigned long long)(index_offset + __float_as_uint(a.w));
                        if (k < best) {
                            best = k; best_d = d;
                            prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;
                        }
                    }
                }
            }
        } else {
            const float4 n0 = nodes4[4 * (long long)cur], n1 = nodes4[4 * (long long)cur + 1];
            const float4 n2 = nodes4[4 * (long long)cur + 2], n3 = nodes4[4 * (long long)cur + 3];
            const float bl = box_lower_bound(x, y, z, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
            const float br = box_lower_bound(x, y, z, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
            const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
            const bool left_first = bl <= br;
            const int near = left_first ? left : right, far = left_first ? right : left;
            const float bn = left_first ? bl : br, bf = left_first ? br : bl;
            const bool take_near = !(bn * BVH_SHRINK > prune_d) && bn < inf;
            const bool take_far = !(bf * BVH_SHRINK > prune_d) && bf < inf;
            if (take_far && sp < BVH_STACK) { stack_ref[sp] = far; stack_bound[sp] = bf; ++sp; }
            if (take_near) { cur = near; continue; }
        }
        bool found = false;
        while (sp > 0) {
            --sp;
            if (!(stack_bound[sp] * BVH_SHRINK > prune_d)) { cur = stack_ref[sp]; found = true; break; }
        }
        if (!found) break;
    }
    keys[p] = best;
}

// ---------------------------------------------------------------------------------------------
// Warp-packet walk (default).  Query slots are in Morton order of the moving vertices, so the 32
// queries of a warp are spatial neighbours: the warp walks the tree TOGETHER with one shared stack,
// node and triangle loads are warp-uniform (one transaction, broadcast), every lane evaluates its
// own bounds and its own S9/S10 distances, and a subtree is entered when ANY lane still needs it.
// Each lane's prune test is exactly the per-thread kernel's, so the winner is the same key.  (The
// per-thread walk diverges badly on a surface: ncu counted 4.3 active threads per warp instruction.)
//
// FUSED (one GPU or sharded queries, no percentile cut): the kernel also does K1 in its prologue
// (q = imx2 @ (mx1 @ a), warm start = previous winning triangle re-evaluated under the new pose) and
// K3+K4 in its epilogue (world distance, thresh, target = closest point on the winning triangle,
// fp64 moment partials), so an ICP iteration is this kernel + one single-CTA solve.
// ---------------------------------------------------------------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(SURF_BOUNDS)
k_nn_surface_packet(const BvhNode* __restrict__ nodes, const float4* __restrict__ sorted, int n_runs,
                    const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                    const DevState* __restrict__ st, const float4* __restrict__ mov_slots,
                    const float4* __restrict__ tris4, long long base_off, long long n_local, double thresh,
                    int warm, long long n_queries, unsigned long long* __restrict__ keys,
                    uint32_t index_offset, float coord_extent, double* __restrict__ partials)
{
    if (st->converged) return;
    __shared__ int s_ref[SURF_WARPS][BVH_STACK];
    __shared__ float s_bound[SURF_WARPS][BVH_STACK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < n_queries;
    const float inf = __int_as_float(0x7f800000);

    float x = 0.f, y = 0.f, z = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    float wx = 0.f, wy = 0.f, wz = 0.f;                 // closest point on the winner (FUSED epilogue)
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
                    const float4 a0 = tris4[3 * j], b0 = tris4[3 * j + 1], c0 = tris4[3 * j + 2];
                    float rx, ry, rz;
                    closest_on_tri(x, y, z, a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, c0.x, c0.y, c0.z, rx, ry, rz);
                    const float d = s10_d2(x, y, z, rx, ry, rz);
                    if (d < inf) {
                        best = ((unsigned long long)__float_as_uint(d) << 32) | (old & 0xFFFFFFFFull);
                        wx = rx; wy = ry; wz = rz;
                    }
                }
            }
        } else {
            x = qx[t]; y = qy[t]; z = qz[t];
            best = keys[t];
        }
    }
    // S9 returns a point within ~eta of the triangle, so its distance can undercut the box bound by
    // eta; prune in DISTANCE terms:  sqrt(bound) > sqrt(best) + eta          (same rule as k_nn_surface)
    const float eta = 3.9e-6f * fmaxf(fmaxf(fabsf(x), fabsf(y)), fmaxf(fabsf(z), coord_extent));
    float best_d = (best == KEY_NONE) ? inf : __uint_as_float((uint32_t)(best >> 32));
    float prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;

    int sp = 0;
    int cur = (n_runs > 1) ? 0 : ~0;
    const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
    if (n_runs > 0) {
        while (true) {
            if (cur < 0) {
                const float4* lp = sorted + (long long)(~cur) * (TRI_LEAF * 3);
#pragma unroll
                for (int i = 0; i < TRI_LEAF; ++i) {
                    const float4 a = __ldg(lp + 3 * i), b = __ldg(lp + 3 * i + 1), c = __ldg(lp + 3 * i + 2);
                    if (a.x < inf) {                                        // warp-uniform (run padding)
                        float rx, ry, rz;
                        closest_on_tri(x, y, z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, rx, ry, rz);
                        const float d = s10_d2(x, y, z, rx, ry, rz);
                        if (d <= best_d && d < inf && active) {
                            const unsigned long long k =
                                ((unsigned long long)__float_as_uint(d) << 32) |
                                (unsigned long long)(index_offset + __float_as_uint(a.w));
                            if (k < best) {
                                best = k; best_d = d;
                                prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;
                                if (FUSED) { wx = rx; wy = ry; wz = rz; }
                            }
                        }
                    }
                }
            } else {
                const float4 n0 = __ldg(nodes4 + 4 * (long long)cur), n1 = __ldg(nodes4 + 4 * (long long)cur + 1);
                const float4 n2 = __ldg(nodes4 + 4 * (long long)cur + 2), n3 = __ldg(nodes4 + 4 * (long long)cur + 3);
                const float bl = box_lower_bound(x, y, z, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
                const float br = box_lower_bound(x, y, z, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
                const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
                const bool wl = active && !(bl * BVH_SHRINK > prune_d) && bl < inf;
                const bool wr = active && !(br * BVH_SHRINK > prune_d) && br < inf;
                const unsigned ml = __ballot_sync(0xffffffffu, wl), mr = __ballot_sync(0xffffffffu, wr);
                if (ml != 0u && mr != 0u) {
                    const unsigned pl = __ballot_sync(0xffffffffu, wl && (!wr || bl <= br));
                    const bool left_first = 2 * __popc(pl) >= __popc(ml | mr);
                    const float fb = left_first ? (wr ? br : inf) : (wl ? bl : inf);
                    const unsigned fmin = __reduce_min_sync(0xffffffffu, __float_as_uint(fb));
                    if (lane == 0 && sp < BVH_STACK) {
                        s_ref[warp][sp] = left_first ? right : left;
                        s_bound[warp][sp] = __uint_as_float(fmin);
                    }
                    if (sp < BVH_STACK) ++sp;
                    __syncwarp();
                    cur = left_first ? left : right;
                    continue;
                }
                if (ml != 0u) { cur = left; continue; }
                if (mr != 0u) { cur = right; continue; }
            }
            bool found = false;
            while (sp > 0) {
                --sp;
                const float pb = s_bound[warp][sp];
                const int pr = s_ref[warp][sp];
                if (__ballot_sync(0xffffffffu, active && !(pb * BVH_SHRINK > prune_d)) != 0u) {
                    cur = pr; found = true; break;
                }
            }
            __syncwarp();
            if (!found) break;
        }
    }
    if (active) keys[t] = best;
    if (FUSED) {
        // K3 + K4 for this thread's query (functions/general.py:299-306, :160-181)
        bool keep = false;
        float bx = 0.f, by = 0.f, bz = 0.f;
        double dist = 0.0;
        if (active && best != KEY_NONE) {
            float wqx, wqy, wqz, w1x, w1y, w1z;
            s1_matvec(st->mw_base, x, y, z, wqx, wqy, wqz);
            s1_matvec(st->mw_base, wx, wy, wz, w1x, w1y, w1z);
            dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
            if (dist < thresh) {
                s1_matvec(st->imx1, w1x, w1y, w1z, bx, by, bz);
                keep = true;
            }
        }
        fused_pair_moments<SURF_THREADS>(keep, ax, ay, az, bx, by, bz, dist, st->shift[0], st->shift[1], st->shift[2], partials);
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-packet walk with DEFERRED triangle tests (default).  At equal base and query density the 32
// queries of a warp share few leaves, so in k_nn_surface_packet every lane pays an S9 test
// (~120 instructions) for every triangle ANY lane needs: ncu put 65 % of the instructions there.
// Here a leaf visit only runs a cheap per-lane box test per triangle and appends the surviving
// (lane, triangle) pairs to a per-warp queue; whenever 32 pairs are queued the warp evaluates them
// in ONE pass, one pair per lane (the owner's query comes by shuffle), and merges the keys into the
// owners' bests with shared-memory atomicMin.  Deferring only delays the tightening of a bound, so
// the walk visits a superset of what the eager walk visits and the winner is the same key.
// ---------------------------------------------------------------------------------------------
constexpr int SURF_QCAP = 32 + 32 * TRI_LEAF;     // queued pairs before a flush + one leaf's worth

// TICKET (with FUSED): the moment sums of every warp enter the ticket tree of nn_tiletree.cuh (`tail`; entry =
// the warp's index in the grid) instead of per-CTA partials, and the warp that completes the top group runs the
// solve and the pose update: the iteration is ONE launch, like the tile-tree walk's.
template <bool FUSED, bool TICKET = false>
__global__ void __launch_bounds__(SURF_BOUNDS, ICP_SURF_RESIDENT / SURF_BOUNDS)
k_nn_surface_defer(const BvhNode* __restrict__ nodes, const float4* __restrict__ sorted, int n_runs,
                   const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                   DevState* st, const float4* __restrict__ mov_slots,
                   const float4* __restrict__ tris4, long long base_off, long long n_local, double thresh,
                   int warm, long long n_queries, unsigned long long* __restrict__ keys,
                   uint32_t index_offset, float coord_extent, double* __restrict__ partials,
                   const __grid_constant__ TTTail tail)
{
    if (*reinterpret_cast<volatile int*>(&st->converged)) return;     // (this launch's own solve cannot have run yet: it waits for every warp's entry)
    __shared__ double s_mom[TICKET ? SURF_WARPS : 1][MOM];
    __shared__ int s_ref[SURF_WARPS][BVH_STACK];
    __shared__ float s_bound[SURF_WARPS][BVH_STACK];
    __shared__ uint32_t s_queue[SURF_WARPS][SURF_QCAP];
    __shared__ unsigned long long s_best[SURF_WARPS][32];
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
                    const float4 a0 = tris4[3 * j], b0 = tris4[3 * j + 1], c0 = tris4[3 * j + 2];
                    float rx, ry, rz;
                    closest_on_tri(x, y, z, a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, c0.x, c0.y, c0.z, rx, ry, rz);
                    const float d = s10_d2(x, y, z, rx, ry, rz);
                    if (d < inf) best = ((unsigned long long)__float_as_uint(d) << 32) | (old & 0xFFFFFFFFull);
                }
            }
        } else {
            x = qx[t]; y = qy[t]; z = qz[t];
            best = keys[t];
        }
    }
    const float eta = 3.9e-6f * fmaxf(fmaxf(fabsf(x), fabsf(y)), fmaxf(fabsf(z), coord_extent));
    float best_d = (best == KEY_NONE) ? inf : __uint_as_float((uint32_t)(best >> 32));
    float prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;
    s_best[warp][lane] = best;
    __syncwarp();

    int count = 0;                                   // queued pairs (warp-uniform)
    // evaluate the last `n` queued pairs, one per lane, and fold the results into the owners' bests
    auto flush = [&](int n) {
        const int start = count - n;
        const bool mine = lane < n;
        const uint32_t e = mine ? s_queue[warp][start + lane] : (uint32_t)lane;
        const int owner = (int)(e & 31u);
        const float ox = __shfl_sync(0xffffffffu, x, owner), oy = __shfl_sync(0xffffffffu, y, owner),
                    oz = __shfl_sync(0xffffffffu, z, owner);
        if (mine) {
            const float4* tp = sorted + 3ll * (long long)(e >> 5);
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float rx, ry, rz;
            closest_on_tri(ox, oy, oz, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, rx, ry, rz);
            const float d = s10_d2(ox, oy, oz, rx, ry, rz);
            if (d < inf)
                atomicMin(&s_best[warp][owner], ((unsigned long long)__float_as_uint(d) << 32) |
                                                   (unsigned long long)(index_offset + __float_as_uint(a.w)));
        }
        __syncwarp();
        const unsigned long long nb = s_best[warp][lane];
        if (nb < best) {
            best = nb; best_d = __uint_as_float((uint32_t)(nb >> 32));
            prune_d = best_d + (2.0f * sqrtf(best_d) + eta) * eta;
        }
        count = start;
        __syncwarp();
    };

    int sp = 0;
    int cur = (n_runs > 1) ? 0 : ~0;
    const float4* nodes4 = reinterpret_cast<const float4*>(nodes);
    if (n_runs > 0) {
        while (true) {
            if (cur < 0) {
                const long long tri0 = (long long)(~cur) * TRI_LEAF;
                const float4* lp = sorted + tri0 * 3;
#pragma unroll
                for (int i = 0; i < TRI_LEAF; ++i) {
                    const float4 a = __ldg(lp + 3 * i), b = __ldg(lp + 3 * i + 1), c = __ldg(lp + 3 * i + 2);
                    if (a.x < inf) {                                        // warp-uniform (run padding)
                        // a sphere about the centroid (radius b.w, rounded up at build time) holds the triangle:
                        // (|q - centroid| - r)^2 bounds the S9 distance from below (up to eta, as for nodes;
                        // the 2e-6 shrink covers the rounding of |q - centroid|)
                        // the triangle's own box bounds the S9 distance from below (up to eta, as for nodes).
                        // (A bounding-sphere test queues 7x fewer pairs but its sqrt made the walk 10 % slower.)
                        const float lb = box_lower_bound(x, y, z, fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y),
                                                         fminf(fminf(a.z, b.z), c.z), fmaxf(fmaxf(a.x, b.x), c.x),
                                                         fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
                        const bool want = active && !(lb * BVH_SHRINK > prune_d);
                        const unsigned m = __ballot_sync(0xffffffffu, want);
                        if (want) s_queue[warp][count + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)(tri0 + i) << 5) | (uint32_t)lane;
                        count += __popc(m);
                    }
                }
                __syncwarp();
                while (count >= 32) flush(32);
            } else {
                const float4 n0 = __ldg(nodes4 + 4 * (long long)cur), n1 = __ldg(nodes4 + 4 * (long long)cur + 1);
                const float4 n2 = __ldg(nodes4 + 4 * (long long)cur + 2), n3 = __ldg(nodes4 + 4 * (long long)cur + 3);
                const float bl = box_lower_bound(x, y, z, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y);
                const float br = box_lower_bound(x, y, z, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w);
                const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
                const bool wl = active && !(bl * BVH_SHRINK > prune_d) && bl < inf;
                const bool wr = active && !(br * BVH_SHRINK > prune_d) && br < inf;
                const unsigned ml = __ballot_sync(0xffffffffu, wl), mr = __ballot_sync(0xffffffffu, wr);
                if (ml != 0u && mr != 0u) {
                    const unsigned pl = __ballot_sync(0xffffffffu, wl && (!wr || bl <= br));
                    const bool left_first = 2 * __popc(pl) >= __popc(ml | mr);
                    const float fb = left_first ? (wr ? br : inf) : (wl ? bl : inf);
                    const unsigned fmin = __reduce_min_sync(0xffffffffu, __float_as_uint(fb));
                    if (lane == 0 && sp < BVH_STACK) {
                        s_ref[warp][sp] = left_first ? right : left;
                        s_bound[warp][sp] = __uint_as_float(fmin);
                    }
                    if (sp < BVH_STACK) ++sp;
                    __syncwarp();
                    cur = left_first ? left : right;
                    continue;
                }
                if (ml != 0u) { cur = left; continue; }
                if (mr != 0u) { cur = right; continue; }
            }
            // before popping, settle the queue if a parked subtree might be pruned by what it holds
            bool found = false;
            while (sp > 0) {
                --sp;
                const float pb = s_bound[warp][sp];
                const int pr = s_ref[warp][sp];
                if (__ballot_sync(0xffffffffu, active && !(pb * BVH_SHRINK > prune_d)) != 0u) {
                    cur = pr; found = true; break;
                }
            }
            __syncwarp();
            if (!found) break;
        }
        while (count > 0) flush(count < 32 ? count : 32);
    }
    if (active) keys[t] = best;
    if (FUSED) {
        // K3 + K4 for this thread's query (functions/general.py:299-306, :160-181); the target is the
        // closest point on the winning triangle, recomputed from the original-order corners (S9)
        bool keep = false;
        float bx = 0.f, by = 0.f, bz = 0.f;
        double dist = 0.0;
        if (active && best != KEY_NONE) {
            const long long j = (long long)(uint32_t)(best & 0xFFFFFFFFull) - base_off;
            const float4 a0 = tris4[3 * j], b0 = tris4[3 * j + 1], c0 = tris4[3 * j + 2];
            float wx, wy, wz;
            closest_on_tri(x, y, z, a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, c0.x, c0.y, c0.z, wx, wy, wz);
            float wqx, wqy, wqz, w1x, w1y, w1z;
            s1_matvec(st->mw_base, x, y, z, wqx, wqy, wqz);
            s1_matvec(st->mw_base, wx, wy, wz, w1x, w1y, w1z);
            dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
            if (dist < thresh) {
                s1_matvec(st->imx1, w1x, w1y, w1z, bx, by, bz);
                keep = true;
            }
        }
        if (!TICKET) {
            fused_pair_moments<SURF_THREADS>(keep, ax, ay, az, bx, by, bz, dist, st->shift[0], st->shift[1], st->shift[2], partials);
        } else {
            const double s0 = st->shift[0], s1 = st->shift[1], s2 = st->shift[2];
            const double k = keep ? 1.0 : 0.0;
            const double a0 = keep ? (double)ax - s0 : 0.0, a1 = keep ? (double)ay - s1 : 0.0, a2 = keep ? (double)az - s2 : 0.0;
            const double b0 = keep ? (double)bx - s0 : 0.0, b1 = keep ? (double)by - s1 : 0.0, b2 = keep ? (double)bz - s2 : 0.0;
            const double d = keep ? dist : 0.0;
            double r0, r1, r2;
            { const double v[8] = {k, a0, a1, a2, b0, b1, b2, b0 * a0}; r0 = warp_sum8(v, lane); }
            { const double v[8] = {b0 * a1, b0 * a2, b1 * a0, b1 * a1, b1 * a2, b2 * a0, b2 * a1, b2 * a2}; r1 = warp_sum8(v, lane); }
            { const double v[8] = {a0 * a0 + a1 * a1 + a2 * a2, b0 * b0 + b1 * b1 + b2 * b2, d, d * d, 0.0, 0.0, 0.0, 0.0}; r2 = warp_sum8(v, lane); }
            tt_tail_push(tail, st, (unsigned)(blockIdx.x * SURF_WARPS + warp), r0, r1, r2, lane, s_mom[TICKET ? warp : 0]);
        }
    }
}

}  // namespace icp
