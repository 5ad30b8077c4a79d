// nn_bruteforce_cyl.cuh -- K2, third form: brute-force nearest vertex with a 2-FMA "cylinder"
// filter over Morton-ordered, per-tile-centred base tiles (nn_variant 30..).
//
// Replaces the per-vertex `base_bvh.find_nearest(co_find)` of the reference
// (functions/general.py:297, nearest-VERTEX variant).  Still O(M N): every
// (query, base vertex) pair is evaluated, but the hot loop evaluates only a LOWER BOUND of the
// squared distance that costs 2 FMA per pair instead of the 6 fp32 operations of the metric S4:
//
//     F(q, b) = fma(-2q'b, b'b, fma(-2q'a, b'a, w2))      w2 = b'a^2 + b'b^2
//             ~ (q'a - b'a)^2 + (q'b - b'b)^2 - (q'a^2 + q'b^2)
//
// where a, b are the two axes of largest extent of the shard, c the dropped axis, and primes are
// coordinates relative to the centre of the 1024-vertex tile the base vertex lives in.  F + |q'ab|^2
// is the squared distance to the LINE through b along axis c, which never exceeds the squared
// distance to b.  A 16-vertex chunk can hold the winner only if
//
//     chunk_min(F) <= best_d2 + delta - |q'ab|^2          (best_d2: S4 value of a real vertex)
//
// and only such chunks (16 or 32 vertices, template CH) take the slow path, which re-tests each vertex with F, then with the 3-axis
// expansion E3 = F + b'c (b'c - 2 q'c), and evaluates the exact S4 on what is left.  The result is
// bit-identical to the direct kernel: same 64-bit keys, ties to the lowest ORIGINAL index.
//
// Error budget (u = 2^-24, s = |q'| + B with B >= every |b'| of the tile; derivation in DESIGN.md):
//   centring 2.01, F and w2 rounding 4.1, |q'ab|^2 rounding 2, S4-vs-exact 5.2, threshold
//   arithmetic 3.1, E3 stage 8  =>  needs delta >= 24.5 u s^2;  delta = 4e-6 (|q'|^2 + B^2)
//   >= 2e-6 s^2 = 33.5 u s^2.
//
// Why tiles are Morton-ordered and centred per tile: delta scales with s^2, i.e. with the distance
// from the query to the TILE centre plus the tile radius -- tiny exactly for the tiles that can hold
// near neighbours -- and the chunks that do trigger are the same few for all queries of a warp
// (query slots are Morton-ordered too).
#pragma once
#include "nn_bruteforce.cuh"

namespace icp {

constexpr int CYL_TILE_FLOATS = NN_TILE * 4;          // filter tile: a' | b' | w2 | c'
constexpr int CYL_TILE_BYTES = CYL_TILE_FLOATS * 4;   // exact tile:  x  | y  | z  | original index

struct CylTileHdr {
    float ca, cb, cc;      // tile centre, permuted axes
    float b2;              // >= max |b'|^2 over the tile
};

__device__ __forceinline__ float axis_of(float x, float y, float z, int a)
{
    return a == 0 ? x : (a == 1 ? y : z);
}

// One CTA (256 threads) per tile: bounding box of the tile's finite vertices -> centre ->
// centred planes, w2, the exact planes and the header.  `order` (nullable) maps a sorted
// position to the original vertex index.
__global__ void __launch_bounds__(256)
k_build_cyl_tiles(const float4* __restrict__ base4, const uint32_t* __restrict__ order, long long n,
                  int axis_a, int axis_b, int axis_c, float* __restrict__ ftiles,
                  float* __restrict__ etiles, CylTileHdr* __restrict__ hdrs)
{
    __shared__ float red[7][8];
    __shared__ float cen[3];
    __shared__ float b2s;
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
            const uint32_t o = order ? order[j] : (uint32_t)j;
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
    float* ft = ftiles + t * CYL_TILE_FLOATS;
    float* et = etiles + t * CYL_TILE_FLOATS;
    float b2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = k * 256 + tid;
        float pa = 0.0f, pb = 0.0f, pc = 0.0f, w2 = inf;
        if (fin[k]) {
            pa = __fsub_rn(axis_of(x[k], y[k], z[k], axis_a), ca);
            pb = __fsub_rn(axis_of(x[k], y[k], z[k], axis_b), cb);
            pc = __fsub_rn(axis_of(x[k], y[k], z[k], axis_c), cc);
            w2 = __fmaf_rn(pb, pb, __fmul_rn(pa, pa));
            b2 = fmaxf(b2, __fmaf_rn(pc, pc, w2));
        }
        ft[i] = pa; ft[NN_TILE + i] = pb; ft[2 * NN_TILE + i] = w2; ft[3 * NN_TILE + i] = pc;
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
        b2s = m * 1.00001f;
        CylTileHdr h;
        h.ca = ca; h.cb = cb; h.cc = cc; h.b2 = (b2s < inf) ? b2s : 0.0f;
        hdrs[t] = h;
    }
}

// per (query, tile): centred query, the two multipliers and the chunk threshold in F space
__device__ __forceinline__ void cyl_setup(float qa, float qb, float qc, const CylTileHdr& h, float best,
                                          float& ax, float& ay, float& thr_f)
{
    const float pa = __fsub_rn(qa, h.ca), pb = __fsub_rn(qb, h.cb), pc = __fsub_rn(qc, h.cc);
    const float nq2 = __fmaf_rn(pb, pb, __fmul_rn(pa, pa));
    const float nq3 = __fmaf_rn(pc, pc, nq2);
    const float delta = 4.0e-6f * (nq3 + h.b2);
    ax = -2.0f * pa;
    ay = -2.0f * pb;
    thr_f = (best + delta) - nq2;
}

__device__ __forceinline__ float key_d2(unsigned long long key)
{
    return key == KEY_NONE ? __int_as_float(0x7f800000) : __uint_as_float((uint32_t)(key >> 32));
}

// Slow path (out of line: keeps the hot loop at 80 registers = 3 CTAs per SM): one query against one 16-vertex chunk.
// fx: the chunk inside the shared-memory filter tile; ex: the same chunk of the exact tile (global).
__device__ __noinline__ unsigned long long cyl_recheck(const float* fx, const float* __restrict__ ex,
                                                       float hca, float hcb, float hcc, float hb2,
                                                       float qa, float qb, float qc, int axes, int n_pts,
                                                       uint32_t index_offset, unsigned long long bkey)
{
    const float inf = __int_as_float(0x7f800000);
    const int axis_a = axes & 3, axis_b = (axes >> 2) & 3;
    // original-order coordinates for the exact metric
    const float ox = axis_a == 0 ? qa : (axis_b == 0 ? qb : qc);
    const float oy = axis_a == 1 ? qa : (axis_b == 1 ? qb : qc);
    const float oz = axis_a == 2 ? qa : (axis_b == 2 ? qb : qc);
    const float best = key_d2(bkey);
    const float pa = __fsub_rn(qa, hca), pb = __fsub_rn(qb, hcb), pc = __fsub_rn(qc, hcc);
    const float nq2 = __fmaf_rn(pb, pb, __fmul_rn(pa, pa));
    const float nq3 = __fmaf_rn(pc, pc, nq2);
    const float delta = 4.0e-6f * (nq3 + hb2);
    const float ax = -2.0f * pa, ay = -2.0f * pb, az = -2.0f * pc;
    const float thr_f = (best + delta) - nq2;
    const float thr_3 = thr_f - __fmul_rn(pc, pc);
#pragma unroll 4
    for (int i = 0; i < n_pts; ++i) {
        const float f = __fmaf_rn(ay, fx[NN_TILE + i], __fmaf_rn(ax, fx[i], fx[2 * NN_TILE + i]));
        if (f <= thr_f) {
            const float zc = fx[3 * NN_TILE + i];
            const float e3 = __fmaf_rn(zc, __fadd_rn(zc, az), f);
            if (e3 <= thr_3) {
                const float d = s4_d2(ox, oy, oz, ex[i], ex[NN_TILE + i], ex[2 * NN_TILE + i]);
                if (d < inf) {
                    const unsigned long long k =
                        ((unsigned long long)__float_as_uint(d) << 32) |
                        (unsigned long long)(index_offset + __float_as_uint(ex[3 * NN_TILE + i]));
                    bkey = k < bkey ? k : bkey;
                }
            }
        }
    }
    return bkey;
}

// grid.x = query blocks of NN_THREADS * R slots, grid.y = base ranges of tiles_per_cta tiles.
// Dynamic shared memory: STAGES * CYL_TILE_BYTES + STAGES * 8 (mbarriers).
template <int R, int STAGES, int CH>
__global__ void __launch_bounds__(NN_THREADS)
k_nn_bruteforce_cyl(const float* __restrict__ ftiles, const float* __restrict__ etiles,
                    const CylTileHdr* __restrict__ hdrs, int axes, int n_tiles, int tiles_per_cta,
                    const float* __restrict__ qx, const float* __restrict__ qy,
                    const float* __restrict__ qz, long long n_queries,
                    unsigned long long* __restrict__ keys, uint32_t index_offset, int seeded,
                    const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * CYL_TILE_BYTES);

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
            mbar_expect_tx(&full[s], CYL_TILE_BYTES);
            tma_load_1d(tiles + (size_t)s * CYL_TILE_FLOATS, ftiles + (size_t)(t0 + s) * CYL_TILE_FLOATS,
                        CYL_TILE_BYTES, &full[s]);
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
        float ax[R], ay[R], thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cyl_setup(qa[r], qb[r], qc[r], hdr, best[r], ax[r], ay[r], thr[r]);
        mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        const float* tile = tiles + (size_t)slot * CYL_TILE_FLOATS;
        const float4* sa = reinterpret_cast<const float4*>(tile);
        const float4* sb = sa + NN_TILE / 4;
        const float4* sw = sb + NN_TILE / 4;
#pragma unroll 1
        for (int c = 0; c < NN_TILE / CH; ++c) {
            float cm[R];
#pragma unroll
            for (int g = 0; g < CH / 4; ++g) {
                const float4 A = sa[c * (CH / 4) + g];
                const float4 B = sb[c * (CH / 4) + g];
                const float4 W = sw[c * (CH / 4) + g];
                const uint64_t a01 = pk2(A.x, A.y), a23 = pk2(A.z, A.w);
                const uint64_t b01 = pk2(B.x, B.y), b23 = pk2(B.z, B.w);
                const uint64_t w01 = pk2(W.x, W.y), w23 = pk2(W.z, W.w);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint64_t ax2 = pk2(ax[r], ax[r]), ay2 = pk2(ay[r], ay[r]);
                    float e0, e1, e2, e3;
                    unpk2(fma2(ay2, b01, fma2(ax2, a01, w01)), e0, e1);
                    unpk2(fma2(ay2, b23, fma2(ax2, a23, w23)), e2, e3);
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
                    bkey[r] = cyl_recheck(fx, ex, hdr.ca, hdr.cb, hdr.cc, hdr.b2, qa[r], qb[r], qc[r], axes, CH,
                                          index_offset, bkey[r]);
                    best[r] = key_d2(bkey[r]);
                    cyl_setup(qa[r], qb[r], qc[r], hdr, best[r], ax[r], ay[r], thr[r]);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + STAGES < my_tiles) {
            mbar_expect_tx(&full[slot], CYL_TILE_BYTES);
            tma_load_1d(tiles + (size_t)slot * CYL_TILE_FLOATS,
                        ftiles + (size_t)(t0 + it + STAGES) * CYL_TILE_FLOATS, CYL_TILE_BYTES, &full[slot]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long p = q0 + (long long)r * NN_THREADS;
        if (p < n_queries && bkey[r] != KEY_NONE) atomicMin(&keys[p], bkey[r]);
    }
}

}  // namespace icp
