// icp_state.cuh -- device-resident state shared by all kernels of one handle.
#pragma once
#include <stdint.h>

namespace icp {

// fp64 moment vector accumulated over the kept pairs (about the point `shift`):
// K, sum a'(3), sum b'(3), sum b' a'^T (9), sum|a'|^2, sum|b'|^2, sum d, sum d^2
constexpr int MOM = 20;
constexpr int MOM_THREADS = 256;
constexpr int MOM_BATCH = 4;       // queries each k_filter_moments thread keeps in flight

struct DevState {
    float mw_align[16];          // align_obj.matrix_world (updated every iteration, rule S3)
    float mw_base[16];           // base_obj.matrix_world
    float imx1[16];              // inverse of mw_align (rule S2)
    float imx2[16];              // inverse of mw_base
    double M[16];                // last solved 4x4
    double ring[5];              // translation ring of the convergence test (icp_align.py:93,136-142)
    double shift[3];             // a0: moments are accumulated about this point (cancellation guard)
    int n_iter;
    int converged;               // also the "stop" flag read by every kernel
    int error;                   // icp_status raised on the device (too few pairs, singular)
    int pad;
    unsigned long long sel_prefix, sel_mask, sel_threshold;   // percentile radix-select state
    long long sel_rank;
    int sel_valid;
    int pad2;
};

// one kept pair's contribution to the moment vector (a, b in align-local space, about `shift`)
__device__ __forceinline__ void moments_add(double* acc, float ax, float ay, float az, float bx, float by,
                                            float bz, double dist, double s0, double s1, double s2)
{
    const double a0 = (double)ax - s0, a1 = (double)ay - s1, a2 = (double)az - s2;
    const double b0 = (double)bx - s0, b1 = (double)by - s1, b2 = (double)bz - s2;
    acc[0] += 1.0;
    acc[1] += a0; acc[2] += a1; acc[3] += a2;
    acc[4] += b0; acc[5] += b1; acc[6] += b2;
    acc[7] += b0 * a0; acc[8] += b0 * a1; acc[9] += b0 * a2;
    acc[10] += b1 * a0; acc[11] += b1 * a1; acc[12] += b1 * a2;
    acc[13] += b2 * a0; acc[14] += b2 * a1; acc[15] += b2 * a2;
    acc[16] += a0 * a0 + a1 * a1 + a2 * a2;
    acc[17] += b0 * b0 + b1 * b1 + b2 * b2;
    acc[18] += dist;
    acc[19] += dist * dist;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of a per-thread moment vector -> partials[blockIdx.x][MOM]
template <int THREADS>
__device__ __forceinline__ void block_reduce_moments(const double* acc, double* __restrict__ partials)
{
    __shared__ double sh[THREADS / 32][MOM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MOM; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) sh[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < MOM) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * MOM + threadIdx.x] = v;
    }
}

// Sum over the warp of 8 per-lane values at once: recursive halving (4 + 2 + 1 adds) leaves lane l with the
// total of component (l & 7) over its group of 8 lanes, two butterfly steps finish it across the 4 groups.
// 9 fp64 adds and 9 64-bit shuffles for 8 components, against 40 + 40 for eight separate butterflies --
// fp64 adds are the expensive instruction here (ncu: the 100-DADD epilogue was 19 % of the fused
// tile-tree kernel's samples).  The order of the additions is fixed, so the result is deterministic.
__device__ __forceinline__ double warp_sum8(const double (&v)[8], int lane)
{
    double a[4], b[2];
    const bool u4 = lane & 4, u2 = lane & 2, u1 = lane & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = u4 ? v[i] : v[i + 4], keep = u4 ? v[i + 4] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = u2 ? a[i] : a[i + 2], keep = u2 ? a[i + 2] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    double r = (u1 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, u1 ? b[0] : b[1], 1);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;                                   // component ((lane & 4) | (lane & 2) | (lane & 1)) == lane & 7
}

// Epilogue of the fused search kernels: one kept pair (or none) per thread -> partials[blockIdx.x][MOM].
// The 20 moments are formed and warp-summed eight at a time, so the fp64 vector never sets the register
// count (and the occupancy) of the walk.
template <int THREADS>
__device__ __forceinline__ void fused_pair_moments(bool keep, float ax, float ay, float az, float bx, float by,
                                                   float bz, double dist, double s0, double s1, double s2,
                                                   double* __restrict__ partials)
{
    __shared__ double sh[THREADS / 32][24];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double k = keep ? 1.0 : 0.0;
    const double a0 = keep ? (double)ax - s0 : 0.0, a1 = keep ? (double)ay - s1 : 0.0, a2 = keep ? (double)az - s2 : 0.0;
    const double b0 = keep ? (double)bx - s0 : 0.0, b1 = keep ? (double)by - s1 : 0.0, b2 = keep ? (double)bz - s2 : 0.0;
    const double d = keep ? dist : 0.0;
    {
        const double v[8] = {k, a0, a1, a2, b0, b1, b2, b0 * a0};                       // moments 0..7
        const double r = warp_sum8(v, lane);
        if (lane < 8) sh[warp][lane] = r;
    }
    {
        const double v[8] = {b0 * a1, b0 * a2, b1 * a0, b1 * a1, b1 * a2, b2 * a0, b2 * a1, b2 * a2};   // 8..15
        const double r = warp_sum8(v, lane);
        if (lane < 8) sh[warp][8 + lane] = r;
    }
    {
        const double v[8] = {a0 * a0 + a1 * a1 + a2 * a2, b0 * b0 + b1 * b1 + b2 * b2, d, d * d, 0.0, 0.0, 0.0, 0.0};   // 16..19
        const double r = warp_sum8(v, lane);
        if (lane < 8) sh[warp][16 + lane] = r;
    }
    __syncthreads();
    if (threadIdx.x < MOM) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * MOM + threadIdx.x] = v;
    }
}

}  // namespace icp
