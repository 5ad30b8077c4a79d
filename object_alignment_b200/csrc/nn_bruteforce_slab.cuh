// nn_bruteforce_slab.cuh -- K2, fourth form: brute-force nearest vertex with a 1-FMA "slab" filter in
// front of the cylinder filter of nn_bruteforce_cyl.cuh (nn_variant 40..45; the round-1 default for seeded passes, superseded by nn_bruteforce_q16.cuh).
//
// Replaces the per-vertex `base_bvh.find_nearest(co_find)` of the reference
// (functions/general.py:297, nearest-VERTEX variant).  Still O(M N): every
// (query, base vertex) pair is evaluated, but the hot loop evaluates the weakest useful lower bound
// of the squared distance, ONE fma per pair:
//
//     F1(q, b) = fma(-2 q's, b's, b's^2)  ~  (q's - b's)^2 - q's^2
//
// where s is the slab axis and primes are coordinates relative to the centre of the 1024-vertex tile
// the base vertex lives in.  F1 + q's^2 is the squared distance to the PLANE through b normal to s,
// which never exceeds the squared distance to b.  A chunk of CH vertices can hold the winner only if
//
//     chunk_min(F1) <= best_d2 + delta - q's^2          (best_d2: S4 value of a real vertex)
//
// Tiles carry both candidate axes a and b (the two longest axes of the shard; planes
// a' | b' | w2 = a'^2 + b'^2 | c' | a'^2 | b'^2) and every (warp, tile) picks, by one ballot, the axis
// along which its queries are farther from the tile centre: any axis gives a valid bound, that one
// prunes best.  Tiles are Morton-ordered (compact boxes, small delta); inside a tile the vertices are
// sorted along a (nn_variant 40..43) or left in Morton order (44, 45) -- no measurable difference.
// A chunk that triggers is re-tested per query with the packed 2-fma cylinder bound
// (cyl_chunk_min), and only what survives goes through cyl_recheck (cylinder per vertex, 3-axis
// expansion, exact S4) -- the same slow path as the cylinder kernel, hence the same bit-identical
// 64-bit keys, ties to the lowest ORIGINAL index.
//
// Error budget: F1 has one rounding fewer than the cylinder bound F and the same centring error, so
// the cylinder kernel's delta = 4e-6 (|q'|^2 + B^2) (>= 33.5 u s^2 against a need of 24.5 u s^2,
// derivation in DESIGN.md) covers the slab test with room to spare; tests/test_filter_bounds.py
// checks all four inequalities in numpy with this file's operation order.
//
// Issue cost (tools/ubench/pipes2.cu): 1 FFMA2 + 1 FMNMX3 per two pairs = 3.8 cycles per SMSP against
// 6.7 for the cylinder kernel's 2 FFMA2 + 1 FMNMX3.  C3 (1M x 1M): 65.6 ms per pass, 1.52e13 pairs/s.
#pragma once
#include "nn_bruteforce_cyl.cuh"

namespace icp {

constexpr int SLAB_PLANES = 6;                               // a' | b' | w2 | c' | a'^2 | b'^2
constexpr int SLAB_TILE_FLOATS = NN_TILE * SLAB_PLANES;
constexpr int SLAB_TILE_BYTES = SLAB_TILE_FLOATS * 4;        // 24 KB per TMA bulk copy
constexpr int SLAB_WARPS = NN_THREADS / 32;

__device__ __forceinline__ uint32_t orderable_bits(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// sort key of the within-tile order: (tile of the Morton position, coordinate along a)
__global__ void k_slab_keys(const float4* __restrict__ base4, const uint32_t* __restrict__ order, long long n,
                            int axis_a, unsigned long long* __restrict__ keys)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 b = base4[order[j]];
    keys[j] = ((unsigned long long)(j / NN_TILE) << 32) | (unsigned long long)orderable_bits(axis_of(b.x, b.y, b.z, axis_a));
}

// One CTA (256 threads) per tile: like k_build_cyl_tiles plus the a'^2 and b'^2 planes.
__global__ void __launch_bounds__(256)
k_build_slab_tiles(const float4* __restrict__ base4, const uint32_t* __restrict__ order, long long n,
                   int axis_a, int axis_b, int axis_c, float* __restrict__ ftiles,
                   float* __restrict__ etiles, CylTileHdr* __restrict__ hdrs)
{
    __shared__ float red[7][8];
    __shared__ float cen[3];
    const float inf = __int_as_float(0x7f800000);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t = blockIdx.x;
    float x[4], y[4], z[4];
    uint32_t oi[4];
    bool fin[4];
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long j = t * NN_TILE + k * 256 + tid;
        x[k] = y[k] = z[k] = inf;
        oi[k] = 0xFFFFFFFFu;
        fin[k] = false;
        if (j < n) {
            const uint32_t o = order[j];
            const float4 b = base4[o];
            x[k] = b.x; y[k] = b.y; z[k] = b.z; oi[k] = o;
            fin[k] = (fabsf(b.x) < inf) && (fabsf(b.y) < inf) && (fabsf(b.z) < inf);
            if (fin[k]) {
                lo[0] = fminf(lo[0], b.x); hi[0] = fmaxf(hi[0], b.x);
                lo[1] = fminf(lo[1], b.y); hi[1] = fmaxf(hi[1], b.y);
                lo[2] = fminf(lo[2], b.z); hi[2] = fmaxf(hi[2], b.z);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    if (tid < 3) {
        float l = inf, h = -inf;
        for (int w = 0; w < 8; ++w) { l = fminf(l, red[tid][w]); h = fmaxf(h, red[3 + tid][w]); }
        cen[tid] = (l <= h) ? (0.5f * l + 0.5f * h) : 0.0f;      // no finite vertex: centre 0
    }
    __syncthreads();
    const float c0 = cen[0], c1 = cen[1], c2 = cen[2];
    const float ca = axis_of(c0, c1, c2, axis_a), cb = axis_of(c0, c1, c2, axis_b), cc = axis_of(c0, c1, c2, axis_c);
    float* ft = ftiles + t * SLAB_TILE_FLOATS;
    float* et = etiles + t * CYL_TILE_FLOATS;
    float b2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = k * 256 + tid;
        float pa = 0.0f, pb = 0.0f, pc = 0.0f, w2 = inf, wa = inf, wb = inf;
        if (fin[k]) {
            pa = __fsub_rn(axis_of(x[k], y[k], z[k], axis_a), ca);
            pb = __fsub_rn(axis_of(x[k], y[k], z[k], axis_b), cb);
            pc = __fsub_rn(axis_of(x[k], y[k], z[k], axis_c), cc);
            wa = __fmul_rn(pa, pa);
            wb = __fmul_rn(pb, pb);
            w2 = __fmaf_rn(pb, pb, wa);
            b2 = fmaxf(b2, __fmaf_rn(pc, pc, w2));
        }
        ft[i] = pa; ft[NN_TILE + i] = pb; ft[2 * NN_TILE + i] = w2; ft[3 * NN_TILE + i] = pc;
        ft[4 * NN_TILE + i] = wa; ft[5 * NN_TILE + i] = wb;
        et[i] = x[k]; et[NN_TILE + i] = y[k]; et[2 * NN_TILE + i] = z[k];
        et[3 * NN_TILE + i] = __uint_as_float(oi[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b2 = fmaxf(b2, __shfl_xor_sync(0xffffffffu, b2, o));
    if (lane == 0) red[6][warp] = b2;
    __syncthreads();
    if (tid == 0) {
        float m = 0.0f;
        for (int w = 0; w < 8; ++w) m = fmaxf(m, red[6][w]);
        const float b2s = m * 1.00001f;
        CylTileHdr h;
        h.ca = ca; h.cb = cb; h.cc = cc; h.b2 = (b2s < inf) ? b2s : 0.0f;
        hdrs[t] = h;
    }
}

// per (query, tile): the multiplier and the chunk threshold in F1 space; `use_b` (warp-uniform) picks
// the slab axis for this tile -- any axis gives a valid bound, the one along which the warp's queries
// are farther from the tile centre gives the stronger one
__device__ __forceinline__ void slab_setup(float qa, float qb, float qc, const CylTileHdr& h, float best,
                                           bool use_b, float& ax, float& thr1)
{
    const float pa = __fsub_rn(qa, h.ca), pb = __fsub_rn(qb, h.cb), pc = __fsub_rn(qc, h.cc);
    const float na = __fmul_rn(pa, pa), nb = __fmul_rn(pb, pb);
    const float nq3 = __fmaf_rn(pc, pc, __fadd_rn(na, nb));
    const float delta = 4.0e-6f * (nq3 + h.b2);
    ax = -2.0f * (use_b ? pb : pa);
    thr1 = (best + delta) - (use_b ? nb : na);
}

// Second filter level (out of line): minimum of the cylinder bound F over one chunk, packed.
__device__ __noinline__ float cyl_chunk_min(const float* fx, float ax, float ay, int n_groups)
{
    const float4* sa = reinterpret_cast<const float4*>(fx);
    const float4* sb = reinterpret_cast<const float4*>(fx + NN_TILE);
    const float4* sw = reinterpret_cast<const float4*>(fx + 2 * NN_TILE);
    const uint64_t ax2 = pk2(ax, ax), ay2 = pk2(ay, ay);
    float m = __int_as_float(0x7f800000);
#pragma unroll 2
    for (int g = 0; g < n_groups; ++g) {
        const float4 A = sa[g], B = sb[g], W = sw[g];
        float e0, e1, e2, e3;
        unpk2(fma2(ay2, pk2(B.x, B.y), fma2(ax2, pk2(A.x, A.y), pk2(W.x, W.y))), e0, e1);
        unpk2(fma2(ay2, pk2(B.z, B.w), fma2(ax2, pk2(A.z, A.w), pk2(W.z, W.w))), e2, e3);
        m = min3f(m, e0, e1);
        m = min3f(m, e2, e3);
    }
    return m;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid.x = query blocks of NN_THREADS * R slots, grid.y = base ranges of tiles_per_cta tiles.
// Dynamic shared memory: STAGES * SLAB_TILE_BYTES + 2 * STAGES * 8 (full / empty mbarriers).
//
// Tile ring without a CTA-wide barrier: a warp that has finished a tile arrives on the slot's `empty`
// mbarrier; thread 0 refills a slot one tile LATE (at the end of tile it it refills the slot of tile
// it - 1), by which time every warp has normally left that slot, so a warp delayed by slow-path work
// no longer stalls the other seven (ncu: 10 % of the samples sat on the per-tile __syncthreads).
template <int R, int STAGES, int CH>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_bruteforce_slab(const float* __restrict__ ftiles, const float* __restrict__ etiles,
                     const CylTileHdr* __restrict__ hdrs, int axes, int n_tiles, int tiles_per_cta,
                     const float* __restrict__ qx, const float* __restrict__ qy,
                     const float* __restrict__ qz, long long n_queries,
                     unsigned long long* __restrict__ keys, uint32_t index_offset, int seeded,
                     const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SLAB_TILE_BYTES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31;
    const int t0 = blockIdx.y * tiles_per_cta;
    const int t1 = min(n_tiles, t0 + tiles_per_cta);
    const int my_tiles = t1 - t0;
    if (my_tiles <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SLAB_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < my_tiles; ++s) {
            mbar_expect_tx(&full[s], SLAB_TILE_BYTES);
            tma_load_1d(tiles + (size_t)s * SLAB_TILE_FLOATS, ftiles + (size_t)(t0 + s) * SLAB_TILE_FLOATS,
                        SLAB_TILE_BYTES, &full[s]);
        }
    }

    const float inf = __int_as_float(0x7f800000);
    const int axis_a = axes & 3, axis_b = (axes >> 2) & 3, axis_c = (axes >> 4) & 3;
    const long long q0 = (long long)blockIdx.x * (NN_THREADS * R) + tid;
    float qa[R], qb[R], qc[R], best[R];
    unsigned long long bkey[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = q0 + (long long)r * NN_THREADS;
        const bool ok = p < n_queries;
        const float x = ok ? qx[p] : 0.0f, y = ok ? qy[p] : 0.0f, z = ok ? qz[p] : 0.0f;
        qa[r] = axis_of(x, y, z, axis_a); qb[r] = axis_of(x, y, z, axis_b); qc[r] = axis_of(x, y, z, axis_c);
        bkey[r] = KEY_NONE;
        // keys[p] of a seeded pass names a real vertex of this shard with its exact S4 distance
        // under the current pose (K1 re-evaluated it; other CTAs can only have lowered it)
        if (seeded && ok) bkey[r] = *reinterpret_cast<const volatile unsigned long long*>(&keys[p]);
        best[r] = ok ? key_d2(bkey[r]) : -inf;      // padding slots never trigger
    }

    for (int it = 0; it < my_tiles; ++it) {
        const int slot = it % STAGES;
        const CylTileHdr hdr = hdrs[t0 + it];
        // slab axis of this (warp, tile): the one along which most lanes are farther from the tile centre
        const bool use_b = __popc(__ballot_sync(0xffffffffu, fabsf(qb[0] - hdr.cb) > fabsf(qa[0] - hdr.ca))) > 16;
        float ax[R], thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) slab_setup(qa[r], qb[r], qc[r], hdr, best[r], use_b, ax[r], thr[r]);
        mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        const float* tile = tiles + (size_t)slot * SLAB_TILE_FLOATS;
        const float4* sa = reinterpret_cast<const float4*>(tile + (use_b ? NN_TILE : 0));
        const float4* s1 = reinterpret_cast<const float4*>(tile + (use_b ? 5 * NN_TILE : 4 * NN_TILE));
#pragma unroll 1
        for (int c = 0; c < NN_TILE / CH; ++c) {
            float cm[R];
#pragma unroll
            for (int g = 0; g < CH / 4; ++g) {
                const float4 A = sa[c * (CH / 4) + g];
                const float4 W = s1[c * (CH / 4) + g];
                const uint64_t a01 = pk2(A.x, A.y), a23 = pk2(A.z, A.w);
                const uint64_t w01 = pk2(W.x, W.y), w23 = pk2(W.z, W.w);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint64_t ax2 = pk2(ax[r], ax[r]);
                    float e0, e1, e2, e3;
                    unpk2(fma2(ax2, a01, w01), e0, e1);
                    unpk2(fma2(ax2, a23, w23), e2, e3);
                    cm[r] = (g == 0) ? fminf(e0, e1) : min3f(cm[r], e0, e1);
                    cm[r] = min3f(cm[r], e2, e3);
                }
            }
            bool trig = false;
#pragma unroll
            for (int r = 0; r < R; ++r) trig = trig || (cm[r] <= thr[r]);
            if (trig) {
                const float* fx = tile + c * CH;
                const float* ex = etiles + (size_t)(t0 + it) * CYL_TILE_FLOATS + c * CH;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!(cm[r] <= thr[r])) continue;
                    float cax, cay, cthr;
                    cyl_setup(qa[r], qb[r], qc[r], hdr, best[r], cax, cay, cthr);
                    if (!(cyl_chunk_min(fx, cax, cay, CH / 4) <= cthr)) continue;
                    bkey[r] = cyl_recheck(fx, ex, hdr.ca, hdr.cb, hdr.cc, hdr.b2, qa[r], qb[r], qc[r], axes, CH,
                                          index_offset, bkey[r]);
                    best[r] = key_d2(bkey[r]);
                    slab_setup(qa[r], qb[r], qc[r], hdr, best[r], use_b, ax[r], thr[r]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);                   // this warp is done with the slot
        if (tid == 0 && it >= 1 && it - 1 + STAGES < my_tiles) {    // late refill of the previous slot
            const int j = it - 1, js = j % STAGES;
            mbar_wait(&empty[js], (uint32_t)((j / STAGES) & 1));
            mbar_expect_tx(&full[js], SLAB_TILE_BYTES);
            tma_load_1d(tiles + (size_t)js * SLAB_TILE_FLOATS, ftiles + (size_t)(t0 + j + STAGES) * SLAB_TILE_FLOATS,
                        SLAB_TILE_BYTES, &full[js]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = q0 + (long long)r * NN_THREADS;
        if (p < n_queries && bkey[r] != KEY_NONE) atomicMin(&keys[p], bkey[r]);
    }
}

}  // namespace icp
