// nn_bruteforce.cuh -- K2: tiled brute-force nearest-vertex search for sm_100a.
//
// Replaces the per-vertex `base_bvh.find_nearest(co_find)` of the reference
// (functions/general.py:297) in its nearest-VERTEX variant with an O(M N)
// min-search:
//   * base vertices live in HBM as tile-blocked SoA  [tile][x|y|z][NN_TILE]  (12 KB per tile);
//     one elected thread streams tiles into a shared-memory ring with TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) while all warps compute on earlier tiles
//   * every thread owns R queries in registers; a tile plane is read with broadcast LDS.128
//   * the metric is spec S4:  d2 = fma(dz,dz, fma(dy,dy, dx*dx)),  evaluated two base points at
//     a time with Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2) and folded into the
//     running minimum with the 3-input FMNMX3 -- 3.5 issue slots per (query, point) pair
//   * the inner loop tracks only the MIN per 16-point chunk and the id of the first chunk that
//     improved it (branch-free); the exact lowest index is recovered once per query afterwards
//   * results are merged into 64-bit keys  float_bits(d2) << 32 | global_index  with atomicMin,
//     so base ranges (grid.y), base shards (ranks) and ties (lowest index) all resolve by MIN.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"

namespace icp {

constexpr int NN_TILE = 1024;     // base points per tile
constexpr int NN_CHUNK = 16;      // points per argmin-tracking chunk
constexpr int NN_THREADS = 256;
constexpr int NN_TILE_BYTES = NN_TILE * 3 * 4;
constexpr unsigned long long KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint64_t pk2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// S4 on two base points at once; the query is broadcast to both halves.
__device__ __forceinline__ uint64_t s4_d2_x2(uint64_t qx2, uint64_t qy2, uint64_t qz2, uint64_t bx2,
                                             uint64_t by2, uint64_t bz2)
{
    uint64_t dx = sub2(qx2, bx2), dy = sub2(qy2, by2), dz = sub2(qz2, bz2);
    return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
}

// ----------------------------------------------------------------------------- the kernel
// grid.x = query blocks of NN_THREADS * R queries, grid.y = base ranges of tiles_per_cta tiles.
// Dynamic shared memory: STAGES * NN_TILE_BYTES + STAGES * 8 (mbarriers).
template <int R, int STAGES, bool PACKED>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_bruteforce(const float* __restrict__ base_tiles, int n_tiles, int tiles_per_cta,
                const float* __restrict__ qx, const float* __restrict__ qy,
                const float* __restrict__ qz, long long n_queries,
                unsigned long long* __restrict__ keys, uint32_t index_offset,
                const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;   // converged earlier: nothing to do

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * NN_TILE_BYTES);

    const int tid = threadIdx.x;
    const int t0 = blockIdx.y * tiles_per_cta;
    const int t1 = min(n_tiles, t0 + tiles_per_cta);
    const int my_tiles = t1 - t0;
    if (my_tiles <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < my_tiles; ++s) {
            mbar_expect_tx(&full[s], NN_TILE_BYTES);
            tma_load_1d(tiles + (size_t)s * (NN_TILE * 3),
                        base_tiles + (size_t)(t0 + s) * (NN_TILE * 3), NN_TILE_BYTES, &full[s]);
        }
    }

    // this thread's queries: strided by the block so the SoA loads coalesce
    const long long q0 = (long long)blockIdx.x * (NN_THREADS * R) + tid;
    float rqx[R], rqy[R], rqz[R], best[R];
    int bchunk[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = q0 + (long long)r * NN_THREADS;
        const bool ok = p < n_queries;
        rqx[r] = ok ? qx[p] : 0.0f;
        rqy[r] = ok ? qy[p] : 0.0f;
        rqz[r] = ok ? qz[p] : 0.0f;
        best[r] = __int_as_float(0x7f800000);
        bchunk[r] = -1;
    }

    for (int it = 0; it < my_tiles; ++it) {
        const int slot = it % STAGES;
        mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        const float4* sx = reinterpret_cast<const float4*>(tiles + (size_t)slot * (NN_TILE * 3));
        const float4* sy = sx + NN_TILE / 4;
        const float4* sz = sy + NN_TILE / 4;
#pragma unroll 1
        for (int c = 0; c < NN_TILE / NN_CHUNK; ++c) {
            float cm[R];
#pragma unroll
            for (int g = 0; g < NN_CHUNK / 4; ++g) {
                const float4 X = sx[c * (NN_CHUNK / 4) + g];
                const float4 Y = sy[c * (NN_CHUNK / 4) + g];
                const float4 Z = sz[c * (NN_CHUNK / 4) + g];
                if (PACKED) {
                    const uint64_t x01 = pk2(X.x, X.y), x23 = pk2(X.z, X.w);
                    const uint64_t y01 = pk2(Y.x, Y.y), y23 = pk2(Y.z, Y.w);
                    const uint64_t z01 = pk2(Z.x, Z.y), z23 = pk2(Z.z, Z.w);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint64_t qx2 = pk2(rqx[r], rqx[r]), qy2 = pk2(rqy[r], rqy[r]),
                                       qz2 = pk2(rqz[r], rqz[r]);
                        float d0, d1, d2, d3;
                        unpk2(s4_d2_x2(qx2, qy2, qz2, x01, y01, z01), d0, d1);
                        unpk2(s4_d2_x2(qx2, qy2, qz2, x23, y23, z23), d2, d3);
                        cm[r] = (g == 0) ? fminf(d0, d1) : min3f(cm[r], d0, d1);
                        cm[r] = min3f(cm[r], d2, d3);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float d0 = s4_d2(rqx[r], rqy[r], rqz[r], X.x, Y.x, Z.x);
                        const float d1 = s4_d2(rqx[r], rqy[r], rqz[r], X.y, Y.y, Z.y);
                        const float d2 = s4_d2(rqx[r], rqy[r], rqz[r], X.z, Y.z, Z.z);
                        const float d3 = s4_d2(rqx[r], rqy[r], rqz[r], X.w, Y.w, Z.w);
                        cm[r] = (g == 0) ? fminf(d0, d1) : min3f(cm[r], d0, d1);
                        cm[r] = min3f(cm[r], d2, d3);
                    }
                }
            }
            const int chunk_id = it * (NN_TILE / NN_CHUNK) + c;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool better = cm[r] < best[r];      // strict: the earliest chunk keeps a tie
                best[r] = better ? cm[r] : best[r];
                bchunk[r] = better ? chunk_id : bchunk[r];
            }
        }
        __syncthreads();                                   // everyone is done with this slot
        if (tid == 0 && it + STAGES < my_tiles) {
            mbar_expect_tx(&full[slot], NN_TILE_BYTES);
            tma_load_1d(tiles + (size_t)slot * (NN_TILE * 3),
                        base_tiles + (size_t)(t0 + it + STAGES) * (NN_TILE * 3), NN_TILE_BYTES,
                        &full[slot]);
        }
    }

    // exact lowest index inside the winning chunk (scalar S4; FFMA2 and FFMA round identically)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = q0 + (long long)r * NN_THREADS;
        if (p >= n_queries || bchunk[r] < 0) continue;
        const long long j0 = (long long)t0 * NN_TILE + (long long)bchunk[r] * NN_CHUNK;
        const float* tp = base_tiles + (j0 / NN_TILE) * (NN_TILE * 3) + (j0 % NN_TILE);
        float bd = __int_as_float(0x7f800000);
        int bi = 0;
#pragma unroll 4
        for (int i = 0; i < NN_CHUNK; ++i) {
            const float d = s4_d2(rqx[r], rqy[r], rqz[r], tp[i], tp[NN_TILE + i], tp[2 * NN_TILE + i]);
            if (d < bd) { bd = d; bi = i; }
        }
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(bd) << 32) |
            (unsigned long long)(uint32_t)(index_offset + (uint32_t)(j0 + bi));
        atomicMin(&keys[p], key);
    }
}

// =============================================================================================
// K2 with an expansion-form FILTER (nn_variant 20..23): half the fp32-pipe work per pair.
//
// The exact metric S4 costs 6 pipe operations per pair.  Here the inner loop evaluates only
//     E(q, b) = fma(-2q'z, b'z, fma(-2q'y, b'y, fma(-2q'x, b'x, |b'|^2)))  ~  |q - b|^2 - |q'|^2
// (3 FFMA2 per two pairs; primes = coordinates centred on the shard's bounding-box centre c) and
// its per-chunk minimum.  E is NOT used to decide anything by itself: it can be off by rounding.
// With u = 2^-24, s = |q'| + B (B >= every |b'|) one can show for the exact S4 winner j* and
// EVERY j in the scanned range
//     E(j*) <= E(j) + 26.7 u s^2                                  (derivation in DESIGN.md)
// so with delta = 2.0e-6 * s^2 (> 32 u s^2, covering the fp32 evaluation of delta itself) every
// chunk that can contain the winner -- or a tie of the winner -- satisfies
//     chunk_min(E) <= running_min(E) + delta
// when it is reached.  Such chunks (a few dozen per query) take a slow path that evaluates the
// EXACT S4 on the points with E <= min(running_min, chunk_min) + delta and keeps the best 64-bit
// key.  The result is bit-identical to the direct kernel (same keys, ties to the lowest index).
// =============================================================================================
constexpr int NNF_TILE_BYTES = NN_TILE * 4 * 4;      // x' | y' | z' | |b'|^2 planes
__device__ __forceinline__ long long q16_first_hard(const unsigned int* __restrict__ scratch, long long n, int hard_class);

struct FilterInfo {
    float cx, cy, cz;      // centre used for the primed coordinates
    float bmax;            // upper bound of |b'| over the shard
};

__global__ void k_filter_info(const float* __restrict__ bbox, FilterInfo* __restrict__ info)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float c[3], h2 = 0.0f;
    for (int k = 0; k < 3; ++k) {
        const float lo = bbox[k], hi = bbox[3 + k];
        c[k] = 0.5f * lo + 0.5f * hi;
        const float h = fmaxf(hi - c[k], c[k] - lo) * 1.000001f;
        h2 += h * h;
    }
    if (!(h2 >= 0.0f) || !(h2 < __int_as_float(0x7f800000))) { c[0] = c[1] = c[2] = 0.0f; h2 = 0.0f; }
    info->cx = c[0]; info->cy = c[1]; info->cz = c[2];
    info->bmax = sqrtf(h2) * 1.00001f;
}

__global__ void k_build_filter_tiles(const float4* __restrict__ base4, long long n, long long n_padded,
                                     const FilterInfo* __restrict__ info, float* __restrict__ ftiles)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_padded) return;
    float x = 0.0f, y = 0.0f, z = 0.0f, bn = __int_as_float(0x7f800000);
    if (j < n) {
        const float4 b = base4[j];
        x = __fsub_rn(b.x, info->cx); y = __fsub_rn(b.y, info->cy); z = __fsub_rn(b.z, info->cz);
        bn = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
    }
    float* t = ftiles + (j / NN_TILE) * (NN_TILE * 4) + (j % NN_TILE);
    t[0] = x; t[NN_TILE] = y; t[2 * NN_TILE] = z; t[3 * NN_TILE] = bn;
}

template <int R, int STAGES, bool GATED = false>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_bruteforce_filter(const float* __restrict__ ftiles, const float* __restrict__ base_tiles,
                       const FilterInfo* __restrict__ info, int n_tiles, int tiles_per_cta,
                       const float* __restrict__ qx, const float* __restrict__ qy,
                       const float* __restrict__ qz, long long n_queries,
                       unsigned long long* __restrict__ keys, uint32_t index_offset, int seeded,
                       const int* __restrict__ stop_flag, const uint32_t* __restrict__ perm,
                       const unsigned int* __restrict__ cls_scratch, int hard_class)
{
    // GATED: perm = the slots in class order (see "query classes" in nn_bruteforce_q16.cuh), q16_first_hard() = the
    // first position of a HARD class.  The launch then serves only the positions from there on -- the queries
    // whose pruning radius is so large that the q16 kernel's slow path would cost more than this kernel's flat
    // 3 FFMA2 + 1 FMNMX3 per two pairs.  How many there are is only known on the device: blockIdx.x counts
    // blocks of hard positions, the CTAs beyond the last one leave at once, and grid.y splits the base finely
    // so that even a handful of hard blocks fill the machine.
    if (stop_flag != nullptr && *stop_flag != 0) return;
    long long first = 0;                      // GATED: the first position this launch serves
    if (GATED) {
        first = q16_first_hard(cls_scratch, n_queries, hard_class);
        if (first + (long long)blockIdx.x * (NN_THREADS * R) >= n_queries) return;     // blockIdx.x counts HARD blocks
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * NNF_TILE_BYTES);

    const int tid = threadIdx.x;
    const int t0 = blockIdx.y * tiles_per_cta;
    const int t1 = min(n_tiles, t0 + tiles_per_cta);
    const int my_tiles = t1 - t0;
    if (my_tiles <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < my_tiles; ++s) {
            mbar_expect_tx(&full[s], NNF_TILE_BYTES);
            tma_load_1d(tiles + (size_t)s * (NN_TILE * 4), ftiles + (size_t)(t0 + s) * (NN_TILE * 4),
                        NNF_TILE_BYTES, &full[s]);
        }
    }

    const float inf = __int_as_float(0x7f800000);
    const float cx = info->cx, cy = info->cy, cz = info->cz, bmax = info->bmax;
    const long long q0 = first + (long long)blockIdx.x * (NN_THREADS * R) + tid;
    uint64_t nqx[R], nqy[R], nqz[R];          // (-2 q') broadcast into both halves
    float be[R], bed[R], delta[R];
    unsigned long long bkey[R];
    uint32_t slot_of[R];                      // GATED: the slot behind this position, 0xFFFFFFFF = not served by this launch
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long i = q0 + (long long)r * NN_THREADS;
        const bool ok = i < n_queries;
        const long long p = ok ? (GATED ? (long long)perm[i] : i) : 0;
        slot_of[r] = (GATED && ok) ? (uint32_t)p : 0xFFFFFFFFu;
        const float px = __fsub_rn(ok ? qx[p] : cx, cx), py = __fsub_rn(ok ? qy[p] : cy, cy),
                    pz = __fsub_rn(ok ? qz[p] : cz, cz);
        const float ax = -2.0f * px, ay = -2.0f * py, az = -2.0f * pz;
        nqx[r] = pk2(ax, ax); nqy[r] = pk2(ay, ay); nqz[r] = pk2(az, az);
        const float nq = sqrtf(fmaf(pz, pz, fmaf(py, py, px * px))) * 1.000001f;
        const float s = nq + bmax;
        delta[r] = 2.0e-6f * s * s;
        be[r] = inf; bed[r] = inf;
        if (GATED && !ok) { be[r] = -inf; bed[r] = -inf; }      // a position of another class never triggers
        bkey[r] = KEY_NONE;
        if (seeded && ok) {
            // keys[p] already names a real vertex j of THIS shard (K1 re-evaluated the previous
            // winner; other CTAs may have lowered it since -- still a real vertex).  The shard
            // winner j* satisfies E(j*) <= E(j) + delta for every j of the shard, wherever j lies,
            // so E(j) is a valid starting value of the running minimum.
            const unsigned long long k0 = *reinterpret_cast<const volatile unsigned long long*>(&keys[p]);
            const long long j = (long long)(uint32_t)(k0 & 0xFFFFFFFFull) - (long long)index_offset;
            if (k0 != KEY_NONE && j >= 0 && j < (long long)n_tiles * NN_TILE) {
                const float* t = ftiles + (j / NN_TILE) * (NN_TILE * 4) + (j % NN_TILE);
                const float e = __fmaf_rn(az, t[2 * NN_TILE], __fmaf_rn(ay, t[NN_TILE], __fmaf_rn(ax, t[0], t[3 * NN_TILE])));
                if (e < inf) { be[r] = e; bed[r] = e + delta[r]; }
            }
        }
    }

    for (int it = 0; it < my_tiles; ++it) {
        const int slot = it % STAGES;
        mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        const float* tile = tiles + (size_t)slot * (NN_TILE * 4);
        const float4* sx = reinterpret_cast<const float4*>(tile);
        const float4* sy = sx + NN_TILE / 4;
        const float4* sz = sy + NN_TILE / 4;
        const float4* sn = sz + NN_TILE / 4;
#pragma unroll 1
        for (int c = 0; c < NN_TILE / NN_CHUNK; ++c) {
            float cm[R];
#pragma unroll
            for (int g = 0; g < NN_CHUNK / 4; ++g) {
                const float4 X = sx[c * (NN_CHUNK / 4) + g];
                const float4 Y = sy[c * (NN_CHUNK / 4) + g];
                const float4 Z = sz[c * (NN_CHUNK / 4) + g];
                const float4 W = sn[c * (NN_CHUNK / 4) + g];
                const uint64_t x01 = pk2(X.x, X.y), x23 = pk2(X.z, X.w);
                const uint64_t y01 = pk2(Y.x, Y.y), y23 = pk2(Y.z, Y.w);
                const uint64_t z01 = pk2(Z.x, Z.y), z23 = pk2(Z.z, Z.w);
                const uint64_t w01 = pk2(W.x, W.y), w23 = pk2(W.z, W.w);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float e0, e1, e2, e3;
                    unpk2(fma2(nqz[r], z01, fma2(nqy[r], y01, fma2(nqx[r], x01, w01))), e0, e1);
                    unpk2(fma2(nqz[r], z23, fma2(nqy[r], y23, fma2(nqx[r], x23, w23))), e2, e3);
                    cm[r] = (g == 0) ? fminf(e0, e1) : min3f(cm[r], e0, e1);
                    cm[r] = min3f(cm[r], e2, e3);
                }
            }
            bool trig = false;
#pragma unroll
            for (int r = 0; r < R; ++r) trig = trig || (cm[r] <= bed[r]);
            if (trig) {
                // slow path: exact S4 on the candidates of this chunk.  Kept INLINE on purpose: with it
                // ptxas routes the tile operands of the hot loop through uniform registers (R2UR), and
                // FFMA2 with a uniform operand avoids the register-bank dispatch stalls of the
                // all-register form (measured: 116 ms inline vs 136 ms out of line at 1M x 1M)
                const long long jbase = ((long long)(t0 + it)) * NN_TILE + (long long)c * NN_CHUNK;
                const float* gp = base_tiles + (long long)(t0 + it) * (NN_TILE * 3) + c * NN_CHUNK;
                const float* fx = tile + c * NN_CHUNK;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!(cm[r] <= bed[r])) continue;
                    const long long p = GATED ? (long long)slot_of[r] : q0 + (long long)r * NN_THREADS;
                    if (GATED ? slot_of[r] != 0xFFFFFFFFu : p < n_queries) {
                        const float thr = fminf(be[r], cm[r]) + delta[r];
                        float ax, ay, az, dummy;
                        unpk2(nqx[r], ax, dummy); unpk2(nqy[r], ay, dummy); unpk2(nqz[r], az, dummy);
                        const float ox = qx[p], oy = qy[p], oz = qz[p];
#pragma unroll
                        for (int i = 0; i < NN_CHUNK; ++i) {
                            const float e = __fmaf_rn(az, fx[2 * NN_TILE + i],
                                            __fmaf_rn(ay, fx[NN_TILE + i],
                                            __fmaf_rn(ax, fx[i], fx[3 * NN_TILE + i])));
                            if (e <= thr) {
                                const float d = s4_d2(ox, oy, oz, gp[i], gp[NN_TILE + i], gp[2 * NN_TILE + i]);
                                if (d < inf) {
                                    const unsigned long long k =
                                        ((unsigned long long)__float_as_uint(d) << 32) |
                                        (unsigned long long)(uint32_t)(index_offset + (uint32_t)(jbase + i));
                                    bkey[r] = k < bkey[r] ? k : bkey[r];
                                }
                            }
                        }
                    }
                    be[r] = fminf(be[r], cm[r]);
                    bed[r] = be[r] + delta[r];
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + STAGES < my_tiles) {
            mbar_expect_tx(&full[slot], NNF_TILE_BYTES);
            tma_load_1d(tiles + (size_t)slot * (NN_TILE * 4),
                        ftiles + (size_t)(t0 + it + STAGES) * (NN_TILE * 4), NNF_TILE_BYTES, &full[slot]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = GATED ? (long long)slot_of[r] : q0 + (long long)r * NN_THREADS;
        if ((GATED ? slot_of[r] != 0xFFFFFFFFu : p < n_queries) && bkey[r] != KEY_NONE) atomicMin(&keys[p], bkey[r]);
    }
}

}  // namespace icp
