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

// Epilogue of the fused search kernels: one kept pair (or none) per thread -> partials[blockIdx.x][MOM].
// Each moment is formed and warp-summed on the spot instead of living in a 20-double per-thread
// accumulator, so the fp64 vector does not set the register count (and the occupancy) of the walk.
template <int THREADS>
__device__ __forceinline__ void fused_pair_moments(bool keep, float ax, float ay, float az, float bx, float by,
                                                   float bz, double dist, double s0, double s1, double s2,
                                                   double* __restrict__ partials)
{
    __shared__ double sh[THREADS / 32][MOM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double k = keep ? 1.0 : 0.0;
    const double a0 = keep ? (double)ax - s0 : 0.0, a1 = keep ? (double)ay - s1 : 0.0, a2 = keep ? (double)az - s2 : 0.0;
    const double b0 = keep ? (double)bx - s0 : 0.0, b1 = keep ? (double)by - s1 : 0.0, b2 = keep ? (double)bz - s2 : 0.0;
    const double d = keep ? dist : 0.0;
#define ICP_MOM_TERM(i, expr) { const double v_ = warp_sum(expr); if (lane == 0) sh[warp][i] = v_; }
    ICP_MOM_TERM(0, k)
    ICP_MOM_TERM(1, a0) ICP_MOM_TERM(2, a1) ICP_MOM_TERM(3, a2)
    ICP_MOM_TERM(4, b0) ICP_MOM_TERM(5, b1) ICP_MOM_TERM(6, b2)
    ICP_MOM_TERM(7, b0 * a0) ICP_MOM_TERM(8, b0 * a1) ICP_MOM_TERM(9, b0 * a2)
    ICP_MOM_TERM(10, b1 * a0) ICP_MOM_TERM(11, b1 * a1) ICP_MOM_TERM(12, b1 * a2)
    ICP_MOM_TERM(13, b2 * a0) ICP_MOM_TERM(14, b2 * a1) ICP_MOM_TERM(15, b2 * a2)
    ICP_MOM_TERM(16, a0 * a0 + a1 * a1 + a2 * a2)
    ICP_MOM_TERM(17, b0 * b0 + b1 * b1 + b2 * b2)
    ICP_MOM_TERM(18, d)
    ICP_MOM_TERM(19, d * d)
#undef ICP_MOM_TERM
    __syncthreads();
    if (threadIdx.x < MOM) {
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += sh[w][threadIdx.x];
        partials[(size_t)blockIdx.x * MOM + threadIdx.x] = v;
    }
}

}  // namespace icp
