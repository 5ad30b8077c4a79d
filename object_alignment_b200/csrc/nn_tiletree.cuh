// nn_tiletree.cuh -- row N2 of SURVEY.md section 8f, second form: the exact nearest-vertex search as a
// warp-cooperative walk over a 32-ary tree of Morton tiles (default of nn_mode ICP_NN_BVH).
//
// Replaces the per-vertex `base_bvh.find_nearest(co_find)` of the reference (functions/general.py:297,
// nearest-VERTEX variant) with an O(M log N) search that returns the same 64-bit keys
// as the brute-force kernels (float_bits(d2) << 32 | index: ties to the lowest index).
//
// Why a second tree: the binary LBVH packet walk (nn_bvh.cuh) spends 12.5 k warp instructions per 32
// queries -- ~100 dependent node visits, each one two box tests per LANE plus four warp votes.  Here the
// lanes of a warp change roles by level:
//   * index  = the base vertices in Morton order, cut into LEAF TILES of 32 vertices stored SoA
//              (x[32] | y[32] | z[32] | index[32], 512 B), and per level an array of axis-aligned boxes:
//              level 0 = one box per tile, level l = one box per 32 boxes of level l - 1; the top level has
//              at most 32 boxes.  Implicit 32-ary tree, no pointers (1M vertices: 31250 / 977 / 31 boxes).
//   * upper levels: the 32 queries of a warp (neighbours: slots are in Morton order of the moving
//              vertices) are summarised by ONE patch box = the union of the balls |p - q_i| <= T_i,
//              T_i = the lane's current best distance.  A group of 32 boxes is tested in one step,
//              ONE BOX PER LANE against the patch box, one vote.
//   * leaf tiles: lane = query again.  Every lane folds the 32 vertices of a tile into a running minimum
//              with packed fp32 math (3 FADD2/FMUL2/FFMA2 + half an FMNMX3 per vertex); only a lane whose
//              tile minimum reaches its best re-scans the tile for the exact lowest index.
// A vertex that could beat or tie lane i's best lies inside ball i, hence inside the patch box, hence
// its tile's box intersects the patch box and the tile is scanned: exact.  T_i carries a 2e-6 relative
// margin for the fp32 rounding of S4 and of the box arithmetic (like nn_bvh.cuh's BVH_SHRINK).
// Without a seed (first pass) the warp first descends greedily to the tile nearest its first query and
// scans it, which gives every lane a finite bound.
#pragma once
#include "nn_bruteforce.cuh"
#include "nn_bvh.cuh"

namespace icp {

constexpr int TT_LEAF = 32;            // vertices per leaf tile
constexpr int TT_FAN = 32;             // boxes per group = lanes per warp
constexpr int TT_MAX_LEVELS = 6;       // 32^6 tiles: far beyond the 2^32 - 2 vertex limit of the ABI
constexpr int TT_TILE_FLOATS = 4 * TT_LEAF;
constexpr int TT_STACK = TT_MAX_LEVELS * TT_FAN;
#ifndef ICP_TT_THREADS
#define ICP_TT_THREADS 128
#endif
constexpr int TT_THREADS = ICP_TT_THREADS;

struct TileTree {                      // lives in device memory (the walk indexes count / offset by level)
    const float4* boxes;               // all levels back to back; box b = boxes[2b] (lo.xyz), boxes[2b+1] (hi.xyz)
    const float* tiles;                // [tile][x | y | z | index bits][TT_LEAF]
    int n_levels;
    int count[TT_MAX_LEVELS];          // boxes per level (level 0 = tiles)
    int offset[TT_MAX_LEVELS];         // first box of each level inside `boxes`
};

// sorted (Morton order) float4(x, y, z, index bits), padded with +inf / 0xFFFFFFFF -> SoA tiles + level-0 boxes.
// One warp per tile (TT_LEAF == 32: one vertex per lane).
__global__ void k_tt_build_tiles(const float4* __restrict__ sorted, long long n, int n_tiles,
                                 float* __restrict__ tiles, float4* __restrict__ boxes0)
{
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n_tiles) return;
    const float inf = __int_as_float(0x7f800000);
    const long long j = (long long)t * TT_LEAF + lane;
    float4 p = make_float4(inf, inf, inf, __uint_as_float(0xFFFFFFFFu));
    if (j < n) p = sorted[j];
    float* tp = tiles + (size_t)t * TT_TILE_FLOATS;
    tp[lane] = p.x; tp[TT_LEAF + lane] = p.y; tp[2 * TT_LEAF + lane] = p.z; tp[3 * TT_LEAF + lane] = p.w;
    const bool fin = fabsf(p.x) < inf && fabsf(p.y) < inf && fabsf(p.z) < inf;
    float lo[3] = {fin ? p.x : inf, fin ? p.y : inf, fin ? p.z : inf};
    float hi[3] = {fin ? p.x : -inf, fin ? p.y : -inf, fin ? p.z : -inf};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {                   // an empty box (no finite vertex) keeps lo = +inf, hi = -inf: never intersected
        boxes0[2 * (size_t)t] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        boxes0[2 * (size_t)t + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
}

// one warp per output box: the union of up to 32 boxes of the level below
__global__ void k_tt_build_level(const float4* __restrict__ below, int n_below, float4* __restrict__ above, int n_above)
{
    const int g = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n_above) return;
    const float inf = __int_as_float(0x7f800000);
    const int c = g * TT_FAN + lane;
    float4 l = make_float4(inf, inf, inf, 0.f), h = make_float4(-inf, -inf, -inf, 0.f);
    if (c < n_below) { l = below[2 * (size_t)c]; h = below[2 * (size_t)c + 1]; }
    float lo[3] = {l.x, l.y, l.z}, hi[3] = {h.x, h.y, h.z};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {
        above[2 * (size_t)g] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        above[2 * (size_t)g + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
}

// order-preserving float <-> uint (for redux.sync min / max on floats, NaN-free inputs)
__device__ __forceinline__ uint32_t tt_f2o(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float tt_o2f(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ float tt_warp_min(float v) { return tt_o2f(__reduce_min_sync(0xffffffffu, tt_f2o(v))); }
__device__ __forceinline__ float tt_warp_max(float v) { return tt_o2f(__reduce_max_sync(0xffffffffu, tt_f2o(v))); }

constexpr float TT_GROW = 1.000002f;   // 2e-6 relative on the radius: > the fp32 rounding of S4 (5 ulp) and of the box tests

// lowest index among the vertices of a tile at S4 distance exactly m, looking only at the runs of 8 vertices
// named by `runs` (out of line: keeps the scan's 24 vector loads from staying live; scalar S4 rounds
// like the packed form)
template <bool SMEM> __device__ __forceinline__ float4 tt_ld(const float4* p) { return SMEM ? *p : __ldg(p); }
template <bool SMEM> __device__ __forceinline__ uint4 tt_ld(const uint4* p) { return SMEM ? *p : __ldg(p); }

template <bool SMEM>
__device__ __noinline__ uint32_t tt_lowest_index_at(const float* tp, float x, float y, float z, float m, int runs)
{
    const float4* px = reinterpret_cast<const float4*>(tp);
    const float4* py = reinterpret_cast<const float4*>(tp + TT_LEAF);
    const float4* pz = reinterpret_cast<const float4*>(tp + 2 * TT_LEAF);
    const uint4* pi = reinterpret_cast<const uint4*>(tp + 3 * TT_LEAF);
    uint32_t mi = 0xFFFFFFFFu;
    for (int h = 0; h < TT_LEAF / 8; ++h) {
        if (!((runs >> h) & 1)) continue;
#pragma unroll
        for (int g = 2 * h; g < 2 * h + 2; ++g) {
            const float4 X = tt_ld<SMEM>(px + g), Y = tt_ld<SMEM>(py + g), Z = tt_ld<SMEM>(pz + g);
            const uint4 I = tt_ld<SMEM>(pi + g);
            mi = (s4_d2(x, y, z, X.x, Y.x, Z.x) == m && I.x < mi) ? I.x : mi;
            mi = (s4_d2(x, y, z, X.y, Y.y, Z.y) == m && I.y < mi) ? I.y : mi;
            mi = (s4_d2(x, y, z, X.z, Y.z, Z.z) == m && I.z < mi) ? I.z : mi;
            mi = (s4_d2(x, y, z, X.w, Y.w, Z.w) == m && I.w < mi) ? I.w : mi;
        }
    }
    return mi;
}

// Scan one leaf tile for this lane's query: running minimum with packed math, exact key on demand.
template <bool SMEM>
__device__ __forceinline__ void tt_scan_tile(const float* tp, float x, float y, float z,
                                             uint32_t index_offset, unsigned long long& best, float& best_d)
{
    const float inf = __int_as_float(0x7f800000);
    const float4* px = reinterpret_cast<const float4*>(tp);
    const float4* py = reinterpret_cast<const float4*>(tp + TT_LEAF);
    const float4* pz = reinterpret_cast<const float4*>(tp + 2 * TT_LEAF);
    const uint64_t qx2 = pk2(x, x), qy2 = pk2(y, y), qz2 = pk2(z, z);
    float mg[TT_LEAF / 8];                                   // minimum per run of 8 vertices
#pragma unroll
    for (int g = 0; g < TT_LEAF / 4; ++g) {                  // warp-uniform addresses: one transaction each
        const float4 X = tt_ld<SMEM>(px + g), Y = tt_ld<SMEM>(py + g), Z = tt_ld<SMEM>(pz + g);
        float d0, d1, d2, d3;
        unpk2(s4_d2_x2(qx2, qy2, qz2, pk2(X.x, X.y), pk2(Y.x, Y.y), pk2(Z.x, Z.y)), d0, d1);
        unpk2(s4_d2_x2(qx2, qy2, qz2, pk2(X.z, X.w), pk2(Y.z, Y.w), pk2(Z.z, Z.w)), d2, d3);
        if (g % 2 == 0) mg[g / 2] = fminf(d0, d1); else mg[g / 2] = min3f(mg[g / 2], d0, d1);
        mg[g / 2] = min3f(mg[g / 2], d2, d3);
    }
    float m = fminf(min3f(mg[0], mg[1], mg[2]), mg[3]);
    static_assert(TT_LEAF == 32, "four runs of 8 vertices per tile");
    if (m <= best_d && m < inf) {
        // this tile holds a vertex at least as near as the best so far: the lowest index among the
        // vertices at distance exactly m, looked for only in the runs of 8 whose minimum is m
        const int runs = (mg[0] == m ? 1 : 0) | (mg[1] == m ? 2 : 0) | (mg[2] == m ? 4 : 0) | (mg[3] == m ? 8 : 0);
        const uint32_t mi = tt_lowest_index_at<SMEM>(tp, x, y, z, m, runs);
        const unsigned long long k = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned long long)(index_offset + mi);
        if (k < best) { best = k; best_d = m; }
    }
}

// Seeds for the brute-force kernels' FIRST pass: every query descends the tile tree greedily (the child
// box nearest to it at every level) and takes the nearest vertex of the tile it lands in.  Not the
// nearest neighbour in general -- a real vertex of this shard close to the query, with its exact S4
// distance, which is all the seeded brute-force kernels need to prune (any real vertex is a valid seed).
__global__ void __launch_bounds__(128)
k_tt_seed(const TileTree* __restrict__ tree_ptr, const float* __restrict__ qx, const float* __restrict__ qy,
          const float* __restrict__ qz, long long n_queries, unsigned long long* __restrict__ keys,
          uint32_t index_offset, const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_queries) return;
    const TileTree& tree = *tree_ptr;
    if (tree.n_levels <= 0) return;
    const float inf = __int_as_float(0x7f800000);
    const float x = qx[t], y = qy[t], z = qz[t];
    int group = 0;
    for (int lvl = tree.n_levels - 1; lvl >= 0; --lvl) {
        const int c0 = group * TT_FAN, c1 = min(c0 + TT_FAN, tree.count[lvl]);
        float bd = inf;
        int bc = c0;
        for (int c = c0; c < c1; ++c) {
            const float4 lo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
            const float4 hi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
            if (lo.x > hi.x) continue;                       // empty box
            // nearest box first; among the (fat, overlapping) boxes that contain the query the lower bound is 0
            // for all of them, so the distance to the box centre breaks the tie
            float d = box_lower_bound(x, y, z, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
            const float cx = 0.5f * (lo.x + hi.x) - x, cy = 0.5f * (lo.y + hi.y) - y, cz = 0.5f * (lo.z + hi.z) - z;
            d += 0.25f * (cx * cx + cy * cy + cz * cz);
            d = (d >= 0.0f) ? fminf(d, 3.0e38f) : 3.0e38f;   // NaN query: any non-empty box
            if (d < bd) { bd = d; bc = c; }
        }
        group = bc;
    }
    if (group >= tree.count[0]) return;
    unsigned long long best = KEY_NONE;
    float best_d = inf;
    tt_scan_tile<false>(tree.tiles + (size_t)group * TT_TILE_FLOATS, x, y, z, index_offset, best, best_d);
    if (best != KEY_NONE) keys[t] = best;
}

// one 512-byte leaf tile global -> shared, 16 bytes per lane (LDGSTS), and the group bookkeeping
__device__ __forceinline__ void tt_prefetch_tile(float* dst, const float* __restrict__ src, int lane)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;"
                 ::"r"(smem_u32(dst + 4 * lane)), "l"(src + 4 * lane) : "memory");
}
template <int N> __device__ __forceinline__ void tt_wait_prefetch()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// FUSED (one GPU or query-sharded, no percentile): K1 in the prologue (q = imx2 @ (mx1 @ a), warm start
// from the previous winner) and K3 + K4 (world distance, thresh, target, fp64 moment partials) in the
// epilogue, like k_nn_bvh_packet<true>.
template <bool FUSED>
__global__ void __launch_bounds__(TT_THREADS, 1024 / TT_THREADS)       // <= 64 registers: 8 CTAs of 128 threads per SM
k_nn_tiletree(const TileTree* __restrict__ tree_ptr, const float* __restrict__ qx, const float* __restrict__ qy,
              const float* __restrict__ qz, const DevState* __restrict__ st,
              const float4* __restrict__ mov_slots, const float4* __restrict__ base4, long long base_off,
              long long n_local, double thresh, int warm, long long n_queries,
              unsigned long long* __restrict__ keys, uint32_t index_offset, double* __restrict__ partials)
{
    if (st->converged) return;
    __shared__ int s_stack[TT_THREADS / 32][TT_STACK];
    __shared__ __align__(16) float s_tile[TT_THREADS / 32][2][TT_TILE_FLOATS];   // per warp: two leaf tiles in flight
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
    const TileTree& tree = *tree_ptr;
    const int top = tree.n_levels - 1;

    if (tree.n_levels > 0) {
        // ---- no bound yet on some lane: greedy descent to the tile nearest the warp's first query --------
        if (__any_sync(0xffffffffu, active && !(best_d < inf))) {
            const int src = __ffs(__ballot_sync(0xffffffffu, active)) - 1;      // >= 0: some lane is active here
            const float sx = __shfl_sync(0xffffffffu, x, src), sy = __shfl_sync(0xffffffffu, y, src),
                        sz = __shfl_sync(0xffffffffu, z, src);
            int group = 0;
            for (int lvl = top; lvl >= 0; --lvl) {
                const int c = group * TT_FAN + lane;
                float d = inf;
                if (c < tree.count[lvl]) {
                    const float4 lo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
                    const float4 hi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
                    d = box_lower_bound(sx, sy, sz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
                    if (!(d >= 0.0f)) d = inf;                                  // NaN query: any non-empty box will do
                    if (lo.x > hi.x) d = inf; else d = fminf(d, 3.0e38f);       // empty box never, a real one always finite
                }
                const uint32_t dm = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
                const unsigned who = __ballot_sync(0xffffffffu, __float_as_uint(d) == dm);
                group = group * TT_FAN + (__ffs(who) - 1);
            }
            if (group < tree.count[0])
                tt_scan_tile<false>(tree.tiles + (size_t)group * TT_TILE_FLOATS, x, y, z, index_offset, best, best_d);
        }

        // ---- walk: groups of 32 boxes against the warp's patch box, one box per lane -------------------
        int sp = 0;
        if (lane == 0) s_stack[warp][0] = top << 26;            // entry = level << 26 | group
        sp = 1;
        __syncwarp();
        float plo[3], phi[3];
        bool patch_stale = true;
        while (sp > 0) {
            if (patch_stale) {
                // the ball of a lane: radius sqrt(best_d) (1 + 2e-6) + the rounding of q -+ r itself
                const float r = active ? sqrtf(best_d) * TT_GROW : 0.0f;
                const float q[3] = {x, y, z};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float pad = fabsf(q[k]) * 2.4e-7f;
                    plo[k] = tt_warp_min(active ? (q[k] - r) - pad : inf);
                    phi[k] = tt_warp_max(active ? (q[k] + r) + pad : -inf);
                }
                patch_stale = false;
            }
            --sp;
            const int e = s_stack[warp][sp];
            __syncwarp();
            const int lvl = e >> 26, group = e & 0x3FFFFFF;
            const int c = group * TT_FAN + lane;
            bool hit = false;
            if (c < tree.count[lvl]) {
                const float4 lo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
                const float4 hi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
                // NaN anywhere (a NaN query poisons the patch box) counts as a hit: nothing is skipped
                hit = !(lo.x > phi[0] || hi.x < plo[0] || lo.y > phi[1] || hi.y < plo[1] || lo.z > phi[2] || hi.z < plo[2]);
                if (lo.x > hi.x) hit = false;                                   // empty box
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            // The patch box is only a summary (32 consecutive Morton slots can straddle a jump of the curve
            // and span half the object): every candidate box is re-tested against each lane's OWN ball,
            // lane = query, before anything is pushed or scanned (~20 instructions per candidate).
            float4 clo = make_float4(0.f, 0.f, 0.f, 0.f), chi = clo;
            if (hit) {
                clo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
                chi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
            }
            unsigned keep = 0u;
            bool improved = false;
            // next candidate box (lowest set bit first) that some lane's own ball still reaches, or -1;
            // same prune rule as nn_bvh.cuh: skip only when the shrunk bound is STRICTLY greater
            auto next_needed = [&](unsigned& m_) -> int {
                while (m_) {
                    const int b = __ffs(m_) - 1;
                    m_ &= m_ - 1;
                    const float lx = __shfl_sync(0xffffffffu, clo.x, b), ly = __shfl_sync(0xffffffffu, clo.y, b),
                                lz = __shfl_sync(0xffffffffu, clo.z, b), hx = __shfl_sync(0xffffffffu, chi.x, b),
                                hy = __shfl_sync(0xffffffffu, chi.y, b), hz = __shfl_sync(0xffffffffu, chi.z, b);
                    const float d = box_lower_bound(x, y, z, lx, ly, lz, hx, hy, hz);
                    if (__any_sync(0xffffffffu, active && !(d * BVH_SHRINK > best_d))) return b;
                }
                return -1;
            };
            if (lvl > 0) {
                for (int b = next_needed(mask); b >= 0; b = next_needed(mask)) keep |= 1u << b;
            } else {
                // leaf tiles: the next needed tile is copied into shared memory (one 16-byte LDGSTS per lane)
                // while the current one is scanned, so the walk does not sit out a global-memory latency
                // per tile; a prefetched tile that the improved bounds no longer need is simply dropped
                int buf = 0;
                int b = next_needed(mask);
                if (b >= 0) tt_prefetch_tile(s_tile[warp][0], tree.tiles + (size_t)(group * TT_FAN + b) * TT_TILE_FLOATS, lane);
                while (b >= 0) {
                    const int nb = next_needed(mask);
                    if (nb >= 0) {
                        tt_prefetch_tile(s_tile[warp][buf ^ 1], tree.tiles + (size_t)(group * TT_FAN + nb) * TT_TILE_FLOATS, lane);
                        tt_wait_prefetch<1>();
                    } else {
                        tt_wait_prefetch<0>();
                    }
                    __syncwarp();
                    // the bounds may have improved since `b` was chosen: one more look before the scan
                    const float lx = __shfl_sync(0xffffffffu, clo.x, b), ly = __shfl_sync(0xffffffffu, clo.y, b),
                                lz = __shfl_sync(0xffffffffu, clo.z, b), hx = __shfl_sync(0xffffffffu, chi.x, b),
                                hy = __shfl_sync(0xffffffffu, chi.y, b), hz = __shfl_sync(0xffffffffu, chi.z, b);
                    const float d = box_lower_bound(x, y, z, lx, ly, lz, hx, hy, hz);
                    if (__any_sync(0xffffffffu, active && !(d * BVH_SHRINK > best_d))) {
                        const float before = best_d;
                        tt_scan_tile<true>(s_tile[warp][buf], x, y, z, index_offset, best, best_d);
                        improved = improved || (best_d < before);
                    }
                    __syncwarp();                        // everyone is done with s_tile[buf] before it is refilled
                    b = nb;
                    buf ^= 1;
                }
            }
            if (lvl > 0) {
                // push the surviving groups, lowest child last so that it is popped first
                const int n_hit = __popc(keep);
                if ((keep >> lane) & 1u) {
                    const int pos = sp + n_hit - 1 - __popc(keep & ((1u << lane) - 1u));
                    if (pos < TT_STACK) s_stack[warp][pos] = ((lvl - 1) << 26) | c;
                }
                sp = min(sp + n_hit, TT_STACK);
                __syncwarp();
            } else {
                patch_stale = __any_sync(0xffffffffu, improved);
            }
        }
    }
    if (active) keys[t] = best;
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
        fused_pair_moments<TT_THREADS>(keep, ax, ay, az, bx, by, bz, dist, st->shift[0], st->shift[1], st->shift[2], partials);
    }
}

// =====================================================================================================
// Second generation of the walk (default of nn_mode ICP_NN_BVH; the kernel above stays as nn_variant 13).
// Same tree, same prune rule, same keys.  What changed, each against a counter of the ncu capture of the
// first generation (profiles/r2_ncu_tiletree_summary.txt):
//   * PERSISTENT WARPS.  A batch = 32 consecutive query slots = one warp's work.  Warps pull batches from a
//     global queue until it is empty, so no warp waits at a CTA barrier for the slowest of its CTA
//     (barrier stalls were 1.5 of 10.6 cycles per issue, 19 % of the samples sat in the epilogue).
//   * DEFERRED INDEX.  The scan of a leaf tile only tracks the minimum distance and the tile it came from;
//     the lowest index at that distance is looked up ONCE, after the walk, by all lanes together (every
//     lane in its own tile).  The first generation re-scanned inside the walk, lane by lane (14.6 divergent
//     72-instruction blocks per warp = 12 % of the executed instructions).  Exact ties BETWEEN tiles are
//     resolved on the spot by a slow path.  The winner's tile is kept per slot (win_tile) so that the next
//     iteration's seed enters the walk as "best so far, index unresolved, in that tile".
//   * TICKETED TAIL.  Every batch writes its 20 fp64 moment sums as entry `batch` of level 0; the warp that
//     completes a group of 32 entries (atomic ticket) sums the group -- always in lane order, whoever does
//     it -- into one entry of the next level, and so on; the warp that completes the top group holds the
//     iteration's moments and runs the solve + pose update (general.py:181-217, icp_align.py:116-149)
//     right there.  One launch per iteration instead of walk + a 7813-partial reduce + solve (58 of 368 us).
//   * candidate boxes are broadcast from shared memory (2 LDS.128) instead of 6 SHFL per candidate, and the
//     tree descriptor is a kernel parameter (constant bank) instead of global memory.
// =====================================================================================================
constexpr int TT_TAIL_LEVELS = 6;                 // 32^6 batches
constexpr uint32_t TT_UNRES = 0xFFFFFFFFu;        // "lowest index at best_d inside best tile, not looked up yet"
constexpr int TT_REF_TILE = 0x07FFFFFF;           // best_ref = tile | runs << 27 (runs: which runs of 8 hold the minimum)

struct TTQueue {                                  // device memory, zero between launches
    unsigned int next_batch;
    unsigned int exited;
};

struct TTTail {
    double* part;                                 // level l, entry e: part[part_off[l] + (e / 32) * MOM * 32 + c * 32 + e % 32]
    unsigned int* tickets;                        // one per group of 32 entries; zero between launches
    double* moments;                              // [MOM] the iteration's total
    icp_iter_stats* stats;
    double target_d;
    long long part_off[TT_TAIL_LEVELS];
    int ticket_off[TT_TAIL_LEVELS];
    int count[TT_TAIL_LEVELS];                    // entries per level; count[0] = batches
    int solve;                                    // 1: the top ticket also solves and updates the pose (one GPU)
    int scale_mode, use_target, stats_capacity;
};

__device__ void solve_update_body(DevState* st, const double* __restrict__ moments, int scale_mode,
                                  int use_target, double target_d, icp_iter_stats* __restrict__ stats,
                                  int stats_capacity);

// minimum S4 distance from (x, y, z) to the 32 vertices of a tile, and the minimum per run of 8
template <bool SMEM>
__device__ __forceinline__ float tt_scan_min(const float* tp, float x, float y, float z, float (&mg)[TT_LEAF / 8])
{
    const float4* px = reinterpret_cast<const float4*>(tp);
    const float4* py = reinterpret_cast<const float4*>(tp + TT_LEAF);
    const float4* pz = reinterpret_cast<const float4*>(tp + 2 * TT_LEAF);
    const uint64_t qx2 = pk2(x, x), qy2 = pk2(y, y), qz2 = pk2(z, z);
#pragma unroll
    for (int g = 0; g < TT_LEAF / 4; ++g) {                  // warp-uniform addresses: one transaction each
        const float4 X = tt_ld<SMEM>(px + g), Y = tt_ld<SMEM>(py + g), Z = tt_ld<SMEM>(pz + g);
        float d0, d1, d2, d3;
        unpk2(s4_d2_x2(qx2, qy2, qz2, pk2(X.x, X.y), pk2(Y.x, Y.y), pk2(Z.x, Z.y)), d0, d1);
        unpk2(s4_d2_x2(qx2, qy2, qz2, pk2(X.z, X.w), pk2(Y.z, Y.w), pk2(Z.z, Z.w)), d2, d3);
        if (g % 2 == 0) mg[g / 2] = fminf(d0, d1); else mg[g / 2] = min3f(mg[g / 2], d0, d1);
        mg[g / 2] = min3f(mg[g / 2], d2, d3);
    }
    return fminf(min3f(mg[0], mg[1], mg[2]), mg[3]);
}

// lowest index among the vertices of ONE run of 8 at S4 distance exactly m (every lane its own tile / run)
__device__ __forceinline__ uint32_t tt_lowest_in_run(const float* __restrict__ tp, int h, float x, float y, float z, float m)
{
    const float4* px = reinterpret_cast<const float4*>(tp) + 2 * h;
    const float4* py = reinterpret_cast<const float4*>(tp + TT_LEAF) + 2 * h;
    const float4* pz = reinterpret_cast<const float4*>(tp + 2 * TT_LEAF) + 2 * h;
    const uint4* pi = reinterpret_cast<const uint4*>(tp + 3 * TT_LEAF) + 2 * h;
    uint32_t mi = TT_UNRES;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const float4 X = __ldg(px + g), Y = __ldg(py + g), Z = __ldg(pz + g);
        const uint4 I = __ldg(pi + g);
        mi = (s4_d2(x, y, z, X.x, Y.x, Z.x) == m && I.x < mi) ? I.x : mi;
        mi = (s4_d2(x, y, z, X.y, Y.y, Z.y) == m && I.y < mi) ? I.y : mi;
        mi = (s4_d2(x, y, z, X.z, Y.z, Z.z) == m && I.z < mi) ? I.z : mi;
        mi = (s4_d2(x, y, z, X.w, Y.w, Z.w) == m && I.w < mi) ? I.w : mi;
    }
    return mi;
}

// Two tiles hold a vertex at the very same distance m (exact fp32 tie across tiles: lattices, duplicates):
// look both indices up now and keep the lower one.
__device__ __noinline__ void tt_resolve_tie(const float* __restrict__ tiles, const float* tile_s, int T, int runs_t,
                                            float x, float y, float z, float m, int& best_ref, uint32_t& best_idx)
{
    uint32_t ia = best_idx;
    if (ia == TT_UNRES && best_ref >= 0)
        ia = tt_lowest_index_at<false>(tiles + (size_t)(best_ref & TT_REF_TILE) * TT_TILE_FLOATS, x, y, z, m, (best_ref >> 27) & 0xF);
    const uint32_t ib = tt_lowest_index_at<true>(tile_s, x, y, z, m, runs_t);
    if (ib < ia) { best_idx = ib; best_ref = T | (runs_t << 27); }
    else best_idx = ia;
}

// fold one scanned tile into a lane's running best
template <bool SMEM>
__device__ __forceinline__ bool tt_fold_tile(const float* __restrict__ tiles, const float* tp, int T, bool active,
                                             float x, float y, float z, float& best_d, int& best_ref, uint32_t& best_idx)
{
    const float inf = __int_as_float(0x7f800000);
    float mg[TT_LEAF / 8];
    const float m = tt_scan_min<SMEM>(tp, x, y, z, mg);
    static_assert(TT_LEAF == 32, "four runs of 8 vertices per tile");
    if (!active || !(m <= best_d) || !(m < inf)) return false;
    const int runs = (mg[0] == m ? 1 : 0) | (mg[1] == m ? 2 : 0) | (mg[2] == m ? 4 : 0) | (mg[3] == m ? 8 : 0);
    if (m < best_d) {
        best_d = m; best_ref = T | (runs << 27); best_idx = TT_UNRES;
        return true;
    }
    if (best_ref >= 0 && (best_ref & TT_REF_TILE) == T) {        // the seed's own tile: same vertex set, sharper run mask
        if (best_idx == TT_UNRES) best_ref = T | (runs << 27);
        return false;
    }
    if (SMEM) tt_resolve_tie(tiles, tp, T, runs, x, y, z, m, best_ref, best_idx);
    else {          // tile read straight from global memory (greedy first tile): same look-up through the global path
        uint32_t ia = best_idx;
        if (ia == TT_UNRES && best_ref >= 0)
            ia = tt_lowest_index_at<false>(tiles + (size_t)(best_ref & TT_REF_TILE) * TT_TILE_FLOATS, x, y, z, m, (best_ref >> 27) & 0xF);
        const uint32_t ib = tt_lowest_index_at<false>(tp, x, y, z, m, runs);
        if (ib < ia) { best_idx = ib; best_ref = T | (runs << 27); } else best_idx = ia;
    }
    return false;
}

// The ticket tree above level 0 (out of line: runs once per 32 batches).  Called by the warp that completed
// group `grp` of level `lvl`: sums the group in lane order, passes the sum up as one entry of the next
// level, and repeats while it is again the one that completes a group.
__device__ __noinline__ void tt_tail_climb(const TTTail& tail, DevState* st, int lvl, unsigned grp, int lane, double* s_mom)
{
    for (;;) {
        const unsigned n_here = (unsigned)tail.count[lvl];
        const unsigned target = min(32u, n_here - grp * 32u);
        const double* g = tail.part + tail.part_off[lvl] + (size_t)grp * (MOM * 32);
        double v[MOM];
#pragma unroll
        for (int c = 0; c < MOM; ++c) v[c] = (unsigned)lane < target ? __ldcg(g + c * 32 + lane) : 0.0;
        double mine = 0.0;                                // lane c keeps component c
#pragma unroll
        for (int c = 0; c < MOM; ++c) {
            const double s = warp_sum(v[c]);              // fixed order: the same bits whichever warp does it
            if (lane == c) mine = s;
        }
        if (n_here <= 32u) {                              // the top group: `mine` is the iteration's moment vector
            if (lane < MOM) { tail.moments[lane] = mine; s_mom[lane] = mine; }
            __syncwarp();
            if (tail.solve && lane == 0)
                solve_update_body(st, s_mom, tail.scale_mode, tail.use_target, tail.target_d, tail.stats, tail.stats_capacity);
            return;
        }
        ++lvl;
        const unsigned e = grp;
        grp = e >> 5;
        double* o = tail.part + tail.part_off[lvl] + (size_t)grp * (MOM * 32) + (e & 31);
        if (lane < MOM) o[lane * 32] = mine;
        const unsigned tgt = min(32u, (unsigned)tail.count[lvl] - grp * 32u);
        unsigned int* ctr = tail.tickets + tail.ticket_off[lvl] + grp;
        __threadfence();                                  // this entry is visible before its ticket
        __syncwarp();
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(ctr, 1u);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk != tgt - 1u) return;
        if (lane == 0) *ctr = 0u;                         // every arrival has happened: zero again for the next launch
        __threadfence();
    }
}

// One batch's moment sums (r0: components lane, r1: 8 + lane, r2: 16 + lane; valid in lanes < 8 / 8 / 4)
// enter the ticket tree as entry `e` of level 0.
__device__ __forceinline__ void tt_tail_push(const TTTail& tail, DevState* st, unsigned e, double r0, double r1, double r2,
                                             int lane, double* s_mom)
{
    const unsigned grp = e >> 5;
    double* g = tail.part + tail.part_off[0] + (size_t)grp * (MOM * 32) + (e & 31);
    if (lane < 8) { g[lane * 32] = r0; g[(8 + lane) * 32] = r1; }
    if (lane < 4) g[(16 + lane) * 32] = r2;
    const unsigned target = min(32u, (unsigned)tail.count[0] - grp * 32u);
    unsigned int* ctr = tail.tickets + tail.ticket_off[0] + grp;
    __threadfence();
    __syncwarp();
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(ctr, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk != target - 1u) return;
    if (lane == 0) *ctr = 0u;
    __threadfence();
    tt_tail_climb(tail, st, 0, grp, lane, s_mom);
}

template <bool FUSED>
__global__ void __launch_bounds__(TT_THREADS, 1024 / TT_THREADS)
k_nn_tiletree_p(const __grid_constant__ TileTree tree, const float* __restrict__ qx, const float* __restrict__ qy,
                const float* __restrict__ qz, DevState* st, const float4* __restrict__ mov_slots,
                const float4* __restrict__ base4, long long base_off, long long n_local, double thresh, int warm,
                long long n_queries, unsigned long long* __restrict__ keys, int* __restrict__ win_tile,
                uint32_t index_offset, TTQueue* queue, const __grid_constant__ TTTail tail)
{
    constexpr int W = TT_THREADS / 32;
    __shared__ int s_stack[W][TT_STACK];
    __shared__ __align__(16) float s_tile[W][2][TT_TILE_FLOATS];     // per warp: two leaf tiles in flight
    __shared__ __align__(16) float4 s_box[W][2 * TT_FAN];            // per warp: the candidate boxes of the current group
    __shared__ double s_mom[W][MOM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inf = __int_as_float(0x7f800000);
    const unsigned n_batches = (unsigned)((n_queries + 31) >> 5);
    const bool stop = *reinterpret_cast<volatile int*>(&st->converged) != 0;
    const int top = tree.n_levels - 1;

    unsigned batch = 0xFFFFFFFFu;
    if (!stop) {
        if (lane == 0) batch = atomicAdd(&queue->next_batch, 1u);
        batch = __shfl_sync(0xffffffffu, batch, 0);
    }
    while (batch < n_batches) {
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(&queue->next_batch, 1u);      // in flight during the walk
        const long long t = (long long)batch * 32 + lane;
        const bool active = t < n_queries;
        float x = 0.f, y = 0.f, z = 0.f;
        float best_d = inf;
        uint32_t best_idx = TT_UNRES;
        int best_ref = -1;
        if (active) {
            unsigned long long old = KEY_NONE;
            if (FUSED) {
                const float4 a = mov_slots[t];
                float tx, ty, tz;
                s1_matvec(st->mw_align, a.x, a.y, a.z, tx, ty, tz);
                s1_matvec(st->imx2, tx, ty, tz, x, y, z);
                if (warm & 1) old = keys[t];
                const long long j = (long long)(uint32_t)(old & 0xFFFFFFFFull) - base_off;
                if (old != KEY_NONE && j >= 0 && j < n_local) {
                    const float4 b = base4[j];
                    const float d = s4_d2(x, y, z, b.x, b.y, b.z);
                    if (d < inf) { best_d = d; best_idx = (uint32_t)j; }
                }
            } else {
                x = qx[t]; y = qy[t]; z = qz[t];
                old = keys[t];                     // K1 re-evaluated the previous winner (or wrote "none")
                if (old != KEY_NONE) {
                    best_d = __uint_as_float((uint32_t)(old >> 32));
                    best_idx = (uint32_t)(old & 0xFFFFFFFFull) - index_offset;
                    if (!(best_d < inf)) { best_d = inf; best_idx = TT_UNRES; }
                }
            }
            if ((warm & 2) && best_d < inf) {      // the seed's tile is known: the index waits for the final look-up
                const int wt = win_tile[t];
                if (wt >= 0 && wt < tree.count[0]) { best_ref = wt | (0xF << 27); best_idx = TT_UNRES; }
            }
        }
        if (tree.n_levels > 0) {
            // ---- no bound yet on some lane: greedy descent to the tile nearest the warp's first query --------
            if (__any_sync(0xffffffffu, active && !(best_d < inf))) {
                const int src = __ffs(__ballot_sync(0xffffffffu, active)) - 1;
                const float sx = __shfl_sync(0xffffffffu, x, src), sy = __shfl_sync(0xffffffffu, y, src),
                            sz = __shfl_sync(0xffffffffu, z, src);
                int group = 0;
                for (int lvl = top; lvl >= 0; --lvl) {
                    const int c = group * TT_FAN + lane;
                    float d = inf;
                    if (c < tree.count[lvl]) {
                        const float4 lo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
                        const float4 hi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
                        d = box_lower_bound(sx, sy, sz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
                        if (!(d >= 0.0f)) d = inf;
                        if (lo.x > hi.x) d = inf; else d = fminf(d, 3.0e38f);
                    }
                    const uint32_t dm = __reduce_min_sync(0xffffffffu, __float_as_uint(d));
                    const unsigned who = __ballot_sync(0xffffffffu, __float_as_uint(d) == dm);
                    group = group * TT_FAN + (__ffs(who) - 1);
                }
                if (group < tree.count[0])
                    tt_fold_tile<false>(tree.tiles, tree.tiles + (size_t)group * TT_TILE_FLOATS, group, active, x, y, z,
                                        best_d, best_ref, best_idx);
            }

            // ---- walk: groups of 32 boxes against the warp's patch box, one box per lane -------------------
            int sp = 1;
            if (lane == 0) s_stack[warp][0] = top << 26;            // entry = level << 26 | group
            __syncwarp();
            float plo[3], phi[3];
            bool patch_stale = true;
            while (sp > 0) {
                if (patch_stale) {
                    const float r = active ? sqrtf(best_d) * TT_GROW : 0.0f;
                    const float q[3] = {x, y, z};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float pad = fabsf(q[k]) * 2.4e-7f;
                        plo[k] = tt_warp_min(active ? (q[k] - r) - pad : inf);
                        phi[k] = tt_warp_max(active ? (q[k] + r) + pad : -inf);
                    }
                    patch_stale = false;
                }
                --sp;
                const int e = s_stack[warp][sp];
                __syncwarp();
                const int lvl = e >> 26, group = e & 0x3FFFFFF;
                const int c = group * TT_FAN + lane;
                bool hit = false;
                if (c < tree.count[lvl]) {
                    const float4 lo = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c));
                    const float4 hi = __ldg(tree.boxes + 2 * (size_t)(tree.offset[lvl] + c) + 1);
                    // NaN anywhere (a NaN query poisons the patch box) counts as a hit: nothing is skipped
                    hit = !(lo.x > phi[0] || hi.x < plo[0] || lo.y > phi[1] || hi.y < plo[1] || lo.z > phi[2] || hi.z < plo[2]);
                    if (lo.x > hi.x) hit = false;                                   // empty box
                    if (hit) { s_box[warp][2 * lane] = lo; s_box[warp][2 * lane + 1] = hi; }
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                __syncwarp();
                // The patch box is only a summary: every candidate box is re-tested against each lane's OWN
                // ball, lane = query, before anything is pushed or scanned; same prune rule as nn_bvh.cuh
                // (skip only when the shrunk bound is STRICTLY greater)
                auto needs = [&](int b) -> unsigned {          // the lanes whose own ball reaches candidate box b
                    const float4 l = s_box[warp][2 * b], h = s_box[warp][2 * b + 1];
                    const float d = box_lower_bound(x, y, z, l.x, l.y, l.z, h.x, h.y, h.z);
                    return __ballot_sync(0xffffffffu, active && !(d * BVH_SHRINK > best_d));
                };
                auto next_needed = [&](unsigned& m_) -> int {
                    while (m_) {
                        const int b = __ffs(m_) - 1;
                        m_ &= m_ - 1;
                        if (needs(b)) return b;
                    }
                    return -1;
                };
                if (lvl > 0) {
                    unsigned keep = 0u;
                    for (int b = next_needed(mask); b >= 0; b = next_needed(mask)) keep |= 1u << b;
                    // push the surviving groups, lowest child last so that it is popped first
                    const int n_hit = __popc(keep);
                    if ((keep >> lane) & 1u) {
                        const int pos = sp + n_hit - 1 - __popc(keep & ((1u << lane) - 1u));
                        if (pos < TT_STACK) s_stack[warp][pos] = ((lvl - 1) << 26) | c;
                    }
                    sp = min(sp + n_hit, TT_STACK);
                    __syncwarp();
                } else {
                    // leaf tiles: the next needed tile is copied into shared memory (one 16-byte LDGSTS per lane)
                    // while the current one is scanned; a prefetched tile the improved bounds no longer need is dropped
                    bool improved = false;
                    int buf = 0;
                    int b = next_needed(mask);
                    if (b >= 0) tt_prefetch_tile(s_tile[warp][0], tree.tiles + (size_t)(group * TT_FAN + b) * TT_TILE_FLOATS, lane);
                    while (b >= 0) {
                        const int nb = next_needed(mask);
                        if (nb >= 0) {
                            tt_prefetch_tile(s_tile[warp][buf ^ 1], tree.tiles + (size_t)(group * TT_FAN + nb) * TT_TILE_FLOATS, lane);
                            tt_wait_prefetch<1>();
                        } else {
                            tt_wait_prefetch<0>();
                        }
                        __syncwarp();
                        const unsigned who = needs(b);       // the bounds may have improved since `b` was chosen
#ifdef ICP_TT_STATS
                        if (lane == 0 && who) { const int k = __popc(who); atomicAdd(&g_bvh_counters[k <= 2 ? 0 : (k <= 4 ? 1 : (k <= 8 ? 2 : 3))], 1ull); }
#endif
                        if (who)
                            improved = tt_fold_tile<true>(tree.tiles, s_tile[warp][buf], group * TT_FAN + b, active, x, y, z,
                                                          best_d, best_ref, best_idx) || improved;
                        __syncwarp();                        // everyone is done with s_tile[buf] before it is refilled
                        b = nb;
                        buf ^= 1;
                    }
                    patch_stale = __any_sync(0xffffffffu, improved);
                }
            }
        }
        // ---- the index: lowest one at distance best_d inside the winning tile, all lanes together ----------
        uint32_t idx = best_idx;
        {
            unsigned pend = (active && best_d < inf && idx == TT_UNRES && best_ref >= 0) ? ((unsigned)best_ref >> 27) & 0xFu : 0u;
            const float* tp = tree.tiles + (size_t)(best_ref >= 0 ? (best_ref & TT_REF_TILE) : 0) * TT_TILE_FLOATS;
            while (__any_sync(0xffffffffu, pend != 0u)) {
                if (pend) {
                    const int h = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const uint32_t mi = tt_lowest_in_run(tp, h, x, y, z, best_d);
                    idx = mi < idx ? mi : idx;
                }
            }
        }
        unsigned long long best = KEY_NONE;
        if (active && best_d < inf && idx != TT_UNRES)
            best = ((unsigned long long)__float_as_uint(best_d) << 32) | (unsigned long long)(index_offset + idx);
        if (active) {
            keys[t] = best;
            win_tile[t] = (best != KEY_NONE && best_ref >= 0) ? (best_ref & TT_REF_TILE) : -1;
        }
        if (FUSED) {
            // K3 + K4 for this lane's query (functions/general.py:299-306, :160-181)
            bool keep = false;
            float bx = 0.f, by = 0.f, bz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
            double dist = 0.0;
            if (best != KEY_NONE) {
                const float4 a = mov_slots[t];                // again (an L2 hit) rather than three registers across the walk
                ax = a.x; ay = a.y; az = a.z;
                const float4 w = base4[idx];                  // the winner (local index)
                float wqx, wqy, wqz, w1x, w1y, w1z;
                s1_matvec(st->mw_base, x, y, z, wqx, wqy, wqz);
                s1_matvec(st->mw_base, w.x, w.y, w.z, w1x, w1y, w1z);
                dist = s5_length(__fsub_rn(wqx, w1x), __fsub_rn(wqy, w1y), __fsub_rn(wqz, w1z));
                if (dist < thresh) {
                    s1_matvec(st->imx1, w1x, w1y, w1z, bx, by, bz);
                    keep = true;
                }
            }
            const double s0 = st->shift[0], s1 = st->shift[1], s2 = st->shift[2];
            const double k = keep ? 1.0 : 0.0;
            const double a0 = keep ? (double)ax - s0 : 0.0, a1 = keep ? (double)ay - s1 : 0.0, a2 = keep ? (double)az - s2 : 0.0;
            const double b0 = keep ? (double)bx - s0 : 0.0, b1 = keep ? (double)by - s1 : 0.0, b2 = keep ? (double)bz - s2 : 0.0;
            const double d = keep ? dist : 0.0;
            double r0, r1, r2;
            { const double v[8] = {k, a0, a1, a2, b0, b1, b2, b0 * a0}; r0 = warp_sum8(v, lane); }
            { const double v[8] = {b0 * a1, b0 * a2, b1 * a0, b1 * a1, b1 * a2, b2 * a0, b2 * a1, b2 * a2}; r1 = warp_sum8(v, lane); }
            { const double v[8] = {a0 * a0 + a1 * a1 + a2 * a2, b0 * b0 + b1 * b1 + b2 * b2, d, d * d, 0.0, 0.0, 0.0, 0.0}; r2 = warp_sum8(v, lane); }
            tt_tail_push(tail, st, batch, r0, r1, r2, lane, s_mom[warp]);
        }
        batch = __shfl_sync(0xffffffffu, next, 0);
    }
    // the last warp to leave zeroes the queue for the next launch (nobody touches it after leaving)
    if (lane == 0) {
        const unsigned total = gridDim.x * (unsigned)W;
        const unsigned ex = atomicAdd(&queue->exited, 1u);
        if (ex == total - 1u) { queue->next_batch = 0u; queue->exited = 0u; }
    }
}

}  // namespace icp
