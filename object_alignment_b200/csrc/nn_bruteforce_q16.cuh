// nn_bruteforce_q16.cuh -- K2, fifth form: brute-force nearest vertex whose hot loop is ONE integer
// instruction per two (query, base vertex) pairs (nn_variant 50..79; variant 71 is the default for seeded passes).
//
// Replaces the per-vertex `base_bvh.find_nearest(co_find)` of the reference
// (functions/general.py:297, nearest-VERTEX variant).  Still O(M N): every pair is
// evaluated, against the weakest useful lower bound of the distance -- the distance along ONE axis:
//
//     |q_s - b_s| <= T        T = sqrt(best_d2) (1 + eps),  best_d2 the S4 value of a real vertex
//
// A vertex can beat or tie the running best only if it passes this test on every axis: S4 computes
// d2 = fma(dz,dz, fma(dy,dy, dx*dx)) with dx = fl(q.x - b.x), every partial result is >= fl(dx*dx) by
// monotonicity of rounding, so d2 <= best implies |dx| <= sqrt(best)(1 + 2^-23) (same for y, z).
//
// The slab kernel (nn_bruteforce_slab.cuh) spends an FFMA2 and an FMNMX3 per two pairs on this test
// and the two pipes overlap badly (ncu: fma 44 %, alu 48 %).  Here the coordinates of a tile are
// QUANTISED to 16 bits inside the tile's own bounding box and packed two vertices per 32-bit word,
// and the interval test turns into the classic unsigned range check
//
//     u in [lo, hi]   <=>   (u - lo) mod 2^16  <  hi - lo + 1 = W
//
// folded into a running per-chunk minimum by the DPX instruction VIADDMNMX.U16x2:
//
//     acc = min16x2( u_word + (-lo, -lo), acc )           acc starts at (W, W)
//
// After a chunk, acc != (W, W) means some vertex of the chunk lies inside the slab.  One ALU-pipe
// instruction per two pairs, no fp32 work at all, a tile is 6 KB instead of 24 KB.  Quantisation is
// CONSERVATIVE: u = round((b - lo_tile) * inv_tile) is off by at most half a step, the query interval
// is widened by 2 steps plus the fp32 rounding of its own evaluation, so a vertex inside the real
// interval is always inside the integer one (tests/test_filter_bounds.py restates it in numpy).
//
// A chunk that triggers for a query is re-tested per vertex with the 3-axis integer box (three
// VIADDMNMX + one VIMNMX3.U16x2 per two vertices) by the whole warp together, and only vertices inside
// the box get the exact S4 metric from the exact tile in global memory -- so the 64-bit keys are
// bit-identical to the direct kernel's, ties to the lowest ORIGINAL index.
#pragma once
#include "nn_bruteforce_slab.cuh"

namespace icp {

constexpr int Q16_PLANE_WORDS = NN_TILE / 2;                  // two vertices per word
constexpr int Q16_HDR_WORDS = 12;
constexpr int Q16_TILE_WORDS = 3 * Q16_PLANE_WORDS + Q16_HDR_WORDS;   // a | b | c planes + header
constexpr int Q16_TILE_BYTES = Q16_TILE_WORDS * 4;            // 6192 B per TMA bulk copy
constexpr int Q16_WARPS = NN_THREADS / 32;
constexpr float Q16_LEVELS = 65534.0f;                        // 65535 is the padding value (never inside)

// header at the end of every tile (permuted axes a, b, c): lower corner, steps per unit, centre, upper corner
struct Q16Hdr {
    float lo_a, lo_b, lo_c;
    float inv_a, inv_b, inv_c;
    float cen_a, cen_b;
    float hi_a, hi_b, hi_c;       // with lo: the exact bounding box of the tile's finite vertices (the tile-level cull)
    float pad;
};
static_assert(sizeof(Q16Hdr) == Q16_HDR_WORDS * 4, "Q16Hdr layout");

// One CTA (256 threads) per tile.  `order` maps a sorted position to the original vertex index.
__global__ void __launch_bounds__(256)
k_build_q16_tiles(const float4* __restrict__ base4, const uint32_t* __restrict__ order, long long n,
                  int axis_a, int axis_b, int axis_c, uint32_t* __restrict__ qtiles,
                  float* __restrict__ etiles)
{
    __shared__ float red[6][8];
    __shared__ float box[6];
    __shared__ unsigned short us[3][NN_TILE];
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
        if (!(l <= h)) { l = 0.0f; h = 0.0f; }                  // no finite vertex in this tile
        box[tid] = l; box[3 + tid] = h;
    }
    __syncthreads();
    Q16Hdr hd;
    float lo_p[3], inv_p[3];
    const int ax[3] = {axis_a, axis_b, axis_c};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float l = box[ax[k]], h = box[3 + ax[k]];
        // a degenerate extent (all vertices on a plane normal to this axis) still gets a fine grid, so
        // that the axis keeps pruning: steps of about 2^-20 of the coordinate magnitude
        float ext = __fsub_rn(h, l);
        ext = fmaxf(ext, fmaxf(fmaxf(fabsf(l), fabsf(h)) * 1.0e-6f, 1.0e-30f));
        lo_p[k] = l;
        inv_p[k] = __fdiv_rn(Q16_LEVELS, ext);
    }
    hd.lo_a = lo_p[0]; hd.lo_b = lo_p[1]; hd.lo_c = lo_p[2];
    hd.inv_a = inv_p[0]; hd.inv_b = inv_p[1]; hd.inv_c = inv_p[2];
    hd.cen_a = 0.5f * box[axis_a] + 0.5f * box[3 + axis_a];
    hd.cen_b = 0.5f * box[axis_b] + 0.5f * box[3 + axis_b];
    hd.hi_a = box[3 + axis_a]; hd.hi_b = box[3 + axis_b]; hd.hi_c = box[3 + axis_c];
    hd.pad = 0.0f;
    float* et = etiles + t * CYL_TILE_FLOATS;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = k * 256 + tid;
        const float p[3] = {axis_of(x[k], y[k], z[k], axis_a), axis_of(x[k], y[k], z[k], axis_b),
                            axis_of(x[k], y[k], z[k], axis_c)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            unsigned short u = 65535;
            if (fin[k]) {
                // exact product of two floats in double; |u - (p - lo) inv| <= 0.5 (+ 1e-11)
                double tq = ((double)p[a] - (double)lo_p[a]) * (double)inv_p[a];
                tq = floor(tq + 0.5);
                tq = tq < 0.0 ? 0.0 : (tq > (double)Q16_LEVELS ? (double)Q16_LEVELS : tq);
                u = (unsigned short)tq;
            }
            us[a][i] = u;
        }
        et[i] = x[k]; et[NN_TILE + i] = y[k]; et[2 * NN_TILE + i] = z[k];
        et[3 * NN_TILE + i] = __uint_as_float(oi[k]);
    }
    __syncthreads();
    uint32_t* qt = qtiles + t * Q16_TILE_WORDS;
    for (int w = tid; w < 3 * Q16_PLANE_WORDS; w += 256) {
        const int a = w / Q16_PLANE_WORDS, i = w % Q16_PLANE_WORDS;
        qt[w] = (uint32_t)us[a][2 * i] | ((uint32_t)us[a][2 * i + 1] << 16);
    }
    if (tid == 0) *reinterpret_cast<Q16Hdr*>(qt + 3 * Q16_PLANE_WORDS) = hd;
}

// per (query, tile, axis): the packed (-lo, -lo) addend and the packed width (W, W); W = 0 means the
// query's interval misses the tile's range altogether ("never": min(x, 0) stays 0).  T < 0 marks a
// padding slot (never).  NaN / inf inputs give the full range (the exact path then rejects them).
// CARRY: the hot loop adds the packed addend with a plain 32-bit add (IMAD on the fma pipe), so a
// carry out of the low half can raise the high half by one: the width grows by one step to keep the
// test conservative (no carry can occur when lo == 0, where the addend is 0).
template <bool CARRY = false>
__device__ __forceinline__ void q16_setup(float q, float T, float lo, float inv, uint32_t& nlo, uint32_t& wp)
{
    const float c = __fmul_rn(__fsub_rn(q, lo), inv);
    const float r = __fmul_rn(T, inv);
    // 2 steps: half a step of the stored u, the rest for floor/ceil; 2^-21 (|c| + r): the fp32
    // rounding of c, r and of the two sums below
    const float e = __fmaf_rn(4.8e-7f, __fadd_rn(fabsf(c), fabsf(r)), 2.0f);
    const float lf = __fsub_rn(__fsub_rn(c, r), e), hf = __fadd_rn(__fadd_rn(c, r), e);
    const bool never = (hf < 0.0f) || (lf > Q16_LEVELS) || (T < 0.0f);
    const int li = __float2int_rd(fminf(fmaxf(lf, 0.0f), Q16_LEVELS));
    const int hi = __float2int_ru(fminf(fmaxf(hf, 0.0f), Q16_LEVELS));
    nlo = ((uint32_t)(65536 - li) & 0xFFFFu) * 0x10001u;
    wp = never ? 0u : (uint32_t)(hi - li + 1 + ((CARRY && li > 0) ? 1 : 0)) * 0x10001u;
}

// 32-bit add on the fma pipe (IMAD.IADD), leaving the alu pipe to the min instructions
__device__ __forceinline__ uint32_t iadd_fma(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// a + b as IMAD a * one + b with a run-time multiplier (== 1): stays on the fma pipe for certain
__device__ __forceinline__ uint32_t imad_add(uint32_t a, uint32_t one, uint32_t b)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}

__device__ __forceinline__ float q16_radius(float best)
{
    // sqrtf is correctly rounded; 1.000001 covers S4's own roundings (see the header comment)
    return __fmul_rn(sqrtf(best), 1.000001f);
}

// any vertex of the chunk inside the query's interval along one axis?  (chunk = n_words words of `plane`)
__device__ __forceinline__ bool q16_chunk_any(const uint32_t* plane, int n_words, uint32_t nlo, uint32_t wp)
{
    const uint4* p4 = reinterpret_cast<const uint4*>(plane);
    uint32_t m0 = wp, m1 = wp;
#pragma unroll 4
    for (int g = 0; g < n_words / 4; ++g) {
        const uint4 v = p4[g];
        m0 = __viaddmin_u16x2(v.x, nlo, m0);
        m1 = __viaddmin_u16x2(v.y, nlo, m1);
        m0 = __viaddmin_u16x2(v.z, nlo, m0);
        m1 = __viaddmin_u16x2(v.w, nlo, m1);
    }
    return m0 != wp || m1 != wp;
}

// Slow path (out of line, every lane for its own query -- the lanes of a warp are neighbours and
// mostly trigger on the same chunks together): one query against one chunk of n_words * 2 vertices.
//   tile : the shared-memory tile (planes a | b | c, header)      w0 : first word of the chunk
//   ex   : the same chunk of the exact tile (global: x | y | z | original index planes)
//   T    : the query's pruning radius;  flags = axes | (slab axis was b) << 8
// Stage 1: the OTHER in-plane axis, chunk level (one VIADDMNMX per word).  Stage 2: the third axis.
// Stage 3: the 3-axis box per vertex, and the exact S4 metric on the vertices inside it.
__device__ __noinline__ unsigned long long q16_recheck(const uint32_t* tile, int w0, int n_words,
                                                       const float* __restrict__ ex, float qa, float qb,
                                                       float qc, float T, int flags, uint32_t index_offset,
                                                       unsigned long long bkey)
{
    const float inf = __int_as_float(0x7f800000);
    const Q16Hdr hd = *reinterpret_cast<const Q16Hdr*>(tile + 3 * Q16_PLANE_WORDS);
    const bool slab_b = (flags >> 8) & 1;
    const uint32_t* pa = tile + w0;
    const uint32_t* pb = pa + Q16_PLANE_WORDS;
    const uint32_t* pc = pb + Q16_PLANE_WORDS;
    uint32_t na, wa, nb, wb, nc, wc;
    if (slab_b) {
        q16_setup(qa, T, hd.lo_a, hd.inv_a, na, wa);
        if (wa == 0u || !q16_chunk_any(pa, n_words, na, wa)) return bkey;
    } else {
        q16_setup(qb, T, hd.lo_b, hd.inv_b, nb, wb);
        if (wb == 0u || !q16_chunk_any(pb, n_words, nb, wb)) return bkey;
    }
    q16_setup(qc, T, hd.lo_c, hd.inv_c, nc, wc);
    if (wc == 0u || !q16_chunk_any(pc, n_words, nc, wc)) return bkey;
    if (slab_b) q16_setup(qb, T, hd.lo_b, hd.inv_b, nb, wb);
    else q16_setup(qa, T, hd.lo_a, hd.inv_a, na, wa);
    if (wa == 0u || wb == 0u) return bkey;
    const int axis_a = flags & 3, axis_b = (flags >> 2) & 3;
    const float ox = axis_a == 0 ? qa : (axis_b == 0 ? qb : qc);
    const float oy = axis_a == 1 ? qa : (axis_b == 1 ? qb : qc);
    const float oz = axis_a == 2 ? qa : (axis_b == 2 ? qb : qc);
    uint32_t fa = ~wa, fb = ~wb, fc = ~wc;          // per half: 0xFFFF - W
    for (int w = 0; w < n_words; ++w) {
        // per half: e in [0, W], e < W iff inside; e + (0xFFFF - W) < 0xFFFF iff inside (no carry)
        const uint32_t ea = __viaddmin_u16x2(pa[w], na, wa) + fa;
        const uint32_t eb = __viaddmin_u16x2(pb[w], nb, wb) + fb;
        const uint32_t ec = __viaddmin_u16x2(pc[w], nc, wc) + fc;
        const uint32_t f = __vimax3_u16x2(ea, eb, ec);
        if (f == 0xFFFFFFFFu) continue;             // neither vertex of the word is inside the box
        bool improved = false;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (((f >> (16 * half)) & 0xFFFFu) == 0xFFFFu) continue;
            const int i = 2 * w + half;                          // index inside the chunk
            const float d = s4_d2(ox, oy, oz, ex[i], ex[NN_TILE + i], ex[2 * NN_TILE + i]);
            if (d < inf) {
                const unsigned long long k =
                    ((unsigned long long)__float_as_uint(d) << 32) |
                    (unsigned long long)(index_offset + __float_as_uint(ex[3 * NN_TILE + i]));
                if (k < bkey) { bkey = k; improved = true; }
            }
        }
        if (improved) {                             // a nearer real vertex: tighter box for the rest of the chunk
            const float T2 = q16_radius(key_d2(bkey));
            if (T2 < T) {
                T = T2;
                q16_setup(qa, T, hd.lo_a, hd.inv_a, na, wa);
                q16_setup(qb, T, hd.lo_b, hd.inv_b, nb, wb);
                q16_setup(qc, T, hd.lo_c, hd.inv_c, nc, wc);
                if (wa == 0u || wb == 0u || wc == 0u) return bkey;
                fa = ~wa; fb = ~wb; fc = ~wc;
            }
        }
    }
    return bkey;
}

// ---- query classes ---------------------------------------------------------------------------------------
// The hot loop costs the same for every query, the slow path does not: a query whose pruning radius T is
// large (an outlier of the moving cloud, a part of the scan the base does not cover) has a slab that crosses
// a large share of ALL chunks, and since a warp runs the slow path whenever ANY of its lanes triggers, a few
// such queries per warp make the whole warp slow (C5, 25 % uniform outliers spread over every warp: 2369 ms
// per pass against 432 ms at the no-outlier rate).  So the pass walks the slots through a permutation that
// groups queries of similar radius: class 0 = squared radius up to ~8x the median's power of two, then one
// class per factor 4; the order inside a class stays the Morton order of the slots (stable sort), so a warp's
// queries are still neighbours.  Easy warps run at full speed, hard warps run their slow path with all lanes.
constexpr int Q16_CLASS_BITS = 4;
constexpr int Q16_HARD_CLASS = 3;                // classes >= this one (radius > ~11x the median's) go to the expansion-filter kernel

// device scratch of the class machinery (unsigned int[Q16_CLS_WORDS]): two exponent histograms used by alternate
// passes (a pass zeroes the OTHER one for its successor), then the populations of the classes of the current pass
constexpr int Q16_CLS_EXP = 0;                    // + 256 * (pass & 1)
constexpr int Q16_CLS_COUNT = 512;                // [1 << Q16_CLASS_BITS]
constexpr int Q16_CLS_WORDS = 512 + (1 << Q16_CLASS_BITS);

__device__ __forceinline__ uint32_t q16_exponent_of(unsigned long long key, float bound)
{
    float d = key_d2(key);                           // KEY_NONE decodes as NaN bits: exponent 255, the last class
    d = fminf(d, bound);
    return (__float_as_uint(d) >> 23) & 0xFFu;
}
// Exponent histogram of the pruning radii for a pass whose seeds do not come from K1 (the cold pass: K1 adds to
// the histogram itself otherwise).  Block 0 also zeroes the class populations of the pass.
__global__ void k_q16_class_hist(const unsigned long long* __restrict__ keys, const float* __restrict__ bound,
                                 long long n, unsigned int* __restrict__ scratch, int pass, const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    __shared__ unsigned int sh[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x < (1 << Q16_CLASS_BITS)) scratch[Q16_CLS_COUNT + threadIdx.x] = 0;
    __syncthreads();
    const float inf = __int_as_float(0x7f800000);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        atomicAdd(&sh[q16_exponent_of(keys[t], bound != nullptr ? bound[t] : inf)], 1u);
    __syncthreads();
    unsigned int* hist = scratch + Q16_CLS_EXP + 256 * (pass & 1);
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// class of every slot (relative to the exponent bin that holds the median), slot ids for the sort, and the
// populations of the classes; block 0 zeroes the other exponent histogram for the next pass
__global__ void __launch_bounds__(256)
k_q16_class_keys(const unsigned long long* __restrict__ keys, const float* __restrict__ bound, long long n,
                 unsigned int* __restrict__ scratch, int pass, unsigned char* __restrict__ cls,
                 uint32_t* __restrict__ ids, const int* __restrict__ stop_flag)
{
    if (stop_flag != nullptr && *stop_flag != 0) return;
    __shared__ unsigned int s_scan[8];
    __shared__ int s_m;
    __shared__ unsigned int s_cnt[1 << Q16_CLASS_BITS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int* hist = scratch + Q16_CLS_EXP + 256 * (pass & 1);
    if (blockIdx.x == 0) scratch[Q16_CLS_EXP + 256 * ((pass + 1) & 1) + tid] = 0;
    if (tid < (1 << Q16_CLASS_BITS)) s_cnt[tid] = 0;
    if (tid == 0) s_m = 255;
    // inclusive prefix sum of the 256 bins (one per thread), median bin = the first one whose prefix reaches n / 2
    unsigned int v = hist[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) s_scan[warp] = v;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += s_scan[w];
    v += before;
    const unsigned int below = v - hist[tid];
    if (2ull * v >= (unsigned long long)n && 2ull * below < (unsigned long long)n) s_m = tid;     // exactly one thread (or none: n == 0)
    __syncthreads();
    const int m = s_m;
    const long long t = (long long)blockIdx.x * blockDim.x + tid;
    if (t < n) {
        const int e = (int)q16_exponent_of(keys[t], bound != nullptr ? bound[t] : __int_as_float(0x7f800000));
        const int c = e <= m + 2 ? 0 : min((1 << Q16_CLASS_BITS) - 1, (e - m - 1) / 2);
        cls[t] = (unsigned char)c;
        ids[t] = (uint32_t)t;
        atomicAdd(&s_cnt[c], 1u);
    }
    __syncthreads();
    if (tid < (1 << Q16_CLASS_BITS) && s_cnt[tid]) atomicAdd(&scratch[Q16_CLS_COUNT + tid], s_cnt[tid]);
}
// The first position (in class order) that goes to the expansion-filter launch.  Which classes do is decided from
// the populations: when more than a tenth of the queries lie above class 0 the cloud has outliers (or overlaps the
// base only in part) and everything above class 0 is cheaper in the flat-cost kernel (C5: 824 ms against 1093 ms
// per pass); otherwise only classes >= hard_class go there -- the thin upper tail of an ordinary cloud is cheaper
// in the q16 kernel (C3: 29.2 ms against 32.4 ms).
__device__ __forceinline__ long long q16_first_hard(const unsigned int* __restrict__ scratch, long long n, int hard_class)
{
    const unsigned int* cnt = scratch + Q16_CLS_COUNT;
    const long long c0 = (long long)cnt[0];
    if ((n - c0) * 10 > n) return c0;
    long long f = c0;
    for (int c = 1; c < hard_class && c < (1 << Q16_CLASS_BITS); ++c) f += (long long)cnt[c];
    return f;
}

// grid.x = query blocks of NN_THREADS * R slots, grid.y = base ranges of tiles_per_cta tiles.
// Dynamic shared memory: STAGES * Q16_TILE_BYTES + 2 * STAGES * 8 (full / empty mbarriers).
// `bound` (nullable): per-slot squared pruning radius from another source than the seed (the global
// winner of the previous pass moved by the query's displacement, when the base is sharded).
// MIX = 0: one VIADDMNMX.U16x2 per word (2 vertices), alu pipe only: 2.0 cycles per warp instruction.
// MIX >= 1: the packed add moves (partly) to the fma pipe as IMAD and VIMNMX3.U16x2 folds two sums per
//          instruction; both are 2-cycle instructions on 16-lane pipes (IMAD runs on fmaheavy only), so the
//          mixes differ in how they balance alu / fmaheavy / issue slots -- see the loop body.
template <int R, int STAGES, int CH, int MIX, int MINB, bool GATED = false>
__global__ void __launch_bounds__(NN_THREADS, MINB)
k_nn_bruteforce_q16(const uint32_t* __restrict__ qtiles, const float* __restrict__ etiles, int axes,
                    int n_tiles, int tiles_per_cta, const float* __restrict__ qx,
                    const float* __restrict__ qy, const float* __restrict__ qz, long long n_queries,
                    unsigned long long* __restrict__ keys, const float* __restrict__ bound,
                    uint32_t index_offset, const int* __restrict__ stop_flag, const uint32_t* __restrict__ perm,
                    const unsigned int* __restrict__ cls_scratch, int hard_class)
{
    // GATED: the positions from q16_first_hard() on belong to the expansion-filter launch that follows
    // (k_nn_bruteforce_filter<.., true>: the classes with a large pruning radius); a CTA that holds only those leaves at once
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const long long n_mine = GATED ? q16_first_hard(cls_scratch, n_queries, hard_class) : n_queries;
    // Positions are in class order, easy first; the query blocks are taken from the END: a block of hard
    // queries takes many times longer than an easy one, and the blocks that start last set the tail of the launch.
    const long long qblock = (perm != nullptr) ? (long long)(gridDim.x - 1 - blockIdx.x) : (long long)blockIdx.x;
    if (GATED && qblock * (NN_THREADS * R) >= n_mine) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TILE_STRIDE = (Q16_TILE_BYTES + 127) / 128 * 128;      // 6272: keeps every slot 128-B aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * TILE_STRIDE);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31;
    const int t0 = blockIdx.y * tiles_per_cta;
    const int t1 = min(n_tiles, t0 + tiles_per_cta);
    const int my_tiles = t1 - t0;
    if (my_tiles <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], Q16_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < my_tiles; ++s) {
            mbar_expect_tx(&full[s], Q16_TILE_BYTES);
            tma_load_1d(smem_raw + (size_t)s * TILE_STRIDE, qtiles + (size_t)(t0 + s) * Q16_TILE_WORDS,
                        Q16_TILE_BYTES, &full[s]);
        }
    }

    const int axis_a = axes & 3, axis_b = (axes >> 2) & 3, axis_c = (axes >> 4) & 3;
    // R CONSECUTIVE slots per thread: slots are in Morton order of the moving vertices, so a warp's
    // 32 R queries form one compact patch and one slab axis per (warp, tile) suits all of them
    const long long q0 = qblock * (NN_THREADS * R) + (long long)tid * R;
    float qa[R], qb[R], qc[R], bnd[R], rad[R];      // bnd: squared pruning radius, rad: its root (< 0: padding slot)
    unsigned long long bkey[R];
    // `perm` (nullable = identity): the slots in class order (see "query classes" above)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool ok = q0 + r < n_mine;
        const long long p = ok ? (perm != nullptr ? (long long)perm[q0 + r] : q0 + r) : 0;
        const float x = ok ? qx[p] : 0.0f, y = ok ? qy[p] : 0.0f, z = ok ? qz[p] : 0.0f;
        qa[r] = axis_of(x, y, z, axis_a); qb[r] = axis_of(x, y, z, axis_b); qc[r] = axis_of(x, y, z, axis_c);
        bkey[r] = KEY_NONE;
        // keys[p] of a seeded pass names a real vertex of this shard with its exact S4 distance
        // under the current pose (K1 re-evaluated it; other CTAs can only have lowered it)
        if (ok) bkey[r] = *reinterpret_cast<const volatile unsigned long long*>(&keys[p]);
        bnd[r] = key_d2(bkey[r]);
        if (ok && bound != nullptr) bnd[r] = fminf(bnd[r], bound[p]);
        rad[r] = ok ? q16_radius(bnd[r]) : -1.0f;    // padding slot: q16_setup answers "never"
    }
    // a multiplier ptxas cannot prove to be 1: keeps the packed adds on the fma pipe (IMAD) instead of
    // letting it turn them into IADD3 on the alu pipe, which the min instructions need
    const uint32_t one = ((uint32_t)n_tiles >> 30) | 1u;

    for (int it = 0; it < my_tiles; ++it) {
        const int slot = it % STAGES;
        mbar_wait(&full[slot], (uint32_t)((it / STAGES) & 1));
        const uint32_t* tile = reinterpret_cast<const uint32_t*>(smem_raw + (size_t)slot * TILE_STRIDE);
        const Q16Hdr hd = *reinterpret_cast<const Q16Hdr*>(tile + 3 * Q16_PLANE_WORDS);
        // slab axis of this (warp, tile): the one along which most lanes are farther from the tile centre
        const bool use_b = __popc(__ballot_sync(0xffffffffu, fabsf(qb[R / 2] - hd.cen_b) > fabsf(qa[R / 2] - hd.cen_a))) > 16;
        const float lo_s = use_b ? hd.lo_b : hd.lo_a, inv_s = use_b ? hd.inv_b : hd.inv_a;
        uint32_t nlo[R], wp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            q16_setup<MIX != 0>(use_b ? qb[r] : qa[r], rad[r], lo_s, inv_s, nlo[r], wp[r]);
            // Tile-level cull on all three axes: a vertex can beat or tie only if it is within T of the query on
            // EVERY axis, so a tile whose bounding box misses the query's box on any axis cannot trigger at all.
            // Without it a query triggers in every tile that overlaps its slab along the ONE slab axis -- a whole
            // cross-section of the object -- and the slow path sorts that out chunk by chunk (stage 1 / 2).
            // 1.000001 T covers the rounding of the two differences; a NaN compares false (no cull).
            const float tm = rad[r] * 1.000001f;
            const bool miss = (qa[r] - hd.hi_a > tm) || (hd.lo_a - qa[r] > tm) || (qb[r] - hd.hi_b > tm) ||
                              (hd.lo_b - qb[r] > tm) || (qc[r] - hd.hi_c > tm) || (hd.lo_c - qc[r] > tm);
            if (miss) wp[r] = 0u;
        }
        const uint4* plane = reinterpret_cast<const uint4*>(tile + (use_b ? Q16_PLANE_WORDS : 0));
#pragma unroll 1
        for (int c = 0; c < NN_TILE / CH; ++c) {
            uint32_t acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = wp[r];
            if (MIX == 5) {
                // every third word by VIADDMNMX, the other two by IMAD + IMAD + VIMNMX3: alu and fma pipes both
                // at 4 cycles per 6 vertices (0.667 per pair), 4 issue slots per 6 vertices
                uint32_t pend[R];
                int n_pend = 0;
#pragma unroll
                for (int g = 0; g < CH / 8; ++g) {
                    const uint4 v = plane[c * (CH / 8) + g];
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int wi = 4 * g + k;                 // compile-time after unrolling
                        if (wi % 3 == 0) {
#pragma unroll
                            for (int r = 0; r < R; ++r) acc[r] = __viaddmin_u16x2(w4[k], nlo[r], acc[r]);
                        } else if (n_pend == 0) {
#pragma unroll
                            for (int r = 0; r < R; ++r) pend[r] = imad_add(w4[k], one, nlo[r]);
                            n_pend = 1;
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r) acc[r] = __vimin3_u16x2(acc[r], pend[r], imad_add(w4[k], one, nlo[r]));
                            n_pend = 0;
                        }
                    }
                }
                if (n_pend) {
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = __vminu2(acc[r], pend[r]);
                }
            } else
#pragma unroll
            for (int g = 0; g < CH / 8; ++g) {
                const uint4 v = plane[c * (CH / 8) + g];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (MIX == 0) {             // alu pipe only: 4 VIADDMNMX
                        acc[r] = __viaddmin_u16x2(v.x, nlo[r], acc[r]);
                        acc[r] = __viaddmin_u16x2(v.y, nlo[r], acc[r]);
                        acc[r] = __viaddmin_u16x2(v.z, nlo[r], acc[r]);
                        acc[r] = __viaddmin_u16x2(v.w, nlo[r], acc[r]);
                    } else if (MIX == 1) {      // 4 adds left to ptxas' own IMAD / IADD3 balance + 2 VIMNMX3
                        acc[r] = __vimin3_u16x2(acc[r], iadd_fma(v.x, nlo[r]), iadd_fma(v.y, nlo[r]));
                        acc[r] = __vimin3_u16x2(acc[r], iadd_fma(v.z, nlo[r]), iadd_fma(v.w, nlo[r]));
                    } else if (MIX == 2) {      // 3 IMAD (fma pipe) + 1 VIADD.16x2 + 2 VIMNMX3 (alu pipe): 6 + 6 cycles
                        acc[r] = __vimin3_u16x2(acc[r], imad_add(v.x, one, nlo[r]), imad_add(v.y, one, nlo[r]));
                        acc[r] = __vimin3_u16x2(acc[r], imad_add(v.z, one, nlo[r]), __vadd2(v.w, nlo[r]));
                    } else if (MIX == 3) {      // 2 VIADDMNMX + 2 IMAD + 1 VIMNMX3: 5 instructions, alu 6 / fma 4 cycles
                        acc[r] = __viaddmin_u16x2(v.x, nlo[r], acc[r]);
                        acc[r] = __vimin3_u16x2(acc[r], imad_add(v.y, one, nlo[r]), imad_add(v.z, one, nlo[r]));
                        acc[r] = __viaddmin_u16x2(v.w, nlo[r], acc[r]);
                    } else {                    // MIX 4: all four adds as IMAD with the opaque multiplier + 2 VIMNMX3
                        acc[r] = __vimin3_u16x2(acc[r], imad_add(v.x, one, nlo[r]), imad_add(v.y, one, nlo[r]));
                        acc[r] = __vimin3_u16x2(acc[r], imad_add(v.z, one, nlo[r]), imad_add(v.w, one, nlo[r]));
                    }
                }
            }
            bool trig = false;
#pragma unroll
            for (int r = 0; r < R; ++r) trig = trig || (acc[r] != wp[r]);
            if (trig) {
                const float* ex = etiles + (size_t)(t0 + it) * CYL_TILE_FLOATS + c * CH;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (acc[r] == wp[r]) continue;
                    const unsigned long long k = q16_recheck(tile, c * (CH / 2), CH / 2, ex, qa[r], qb[r], qc[r], rad[r],
                                                             axes | (use_b ? 256 : 0), index_offset, bkey[r]);
                    if (k != bkey[r]) {
                        bkey[r] = k;
                        const float d = key_d2(k);
                        if (d < bnd[r]) {
                            bnd[r] = d;
                            rad[r] = q16_radius(d);
                            q16_setup<MIX != 0>(use_b ? qb[r] : qa[r], rad[r], lo_s, inv_s, nlo[r], wp[r]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);                   // this warp is done with the slot
        if (tid == 0 && it >= 1 && it - 1 + STAGES < my_tiles) {    // late refill of the previous slot
            const int j = it - 1, js = j % STAGES;
            mbar_wait(&empty[js], (uint32_t)((j / STAGES) & 1));
            mbar_expect_tx(&full[js], Q16_TILE_BYTES);
            tma_load_1d(smem_raw + (size_t)js * TILE_STRIDE, qtiles + (size_t)(t0 + j + STAGES) * Q16_TILE_WORDS,
                        Q16_TILE_BYTES, &full[js]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (q0 + r < n_mine && bkey[r] != KEY_NONE) {
            const long long p = perm != nullptr ? (long long)perm[q0 + r] : q0 + r;
            atomicMin(&keys[p], bkey[r]);
        }
    }
}

}  // namespace icp
