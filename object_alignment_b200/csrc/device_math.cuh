// device_math.cuh -- the exact arithmetic rules S1..S5 and the 3x3 solver, device side.
//
// These mirror, operation for operation, what the reference gets from Blender's mathutils
// (restated in oracle/mathutils_shim.py) so that q, the world distance and the targets come
// out bit-identical:
//   S1  Matrix @ Vector : float products, double accumulation, one rounding to float
//       (reference call sites: functions/general.py:287,299,304)
//   S2  Matrix.inverted(): fp32 adjugate / determinant            (general.py:265-266)
//   S3  Matrix @ Matrix : same rule as S1                          (operators/icp_align.py:121)
//   S5  Vector.length   : sqrt of a double sum of float squares, taken z, y, x (general.py:299)
// Explicit _rn intrinsics keep nvcc from contracting anything into an FMA.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace icp {

struct Mat4f { float m[16]; };   // row-major

__device__ __forceinline__ float s1_row(const float* __restrict__ row, float x, float y, float z)
{
    double acc = (double)__fmul_rn(row[0], x);
    acc = __dadd_rn(acc, (double)__fmul_rn(row[1], y));
    acc = __dadd_rn(acc, (double)__fmul_rn(row[2], z));
    acc = __dadd_rn(acc, (double)row[3]);            // row[3] * 1.0f is exact
    return __double2float_rn(acc);
}

// S1: out = (M @ (x, y, z, 1)).xyz
__device__ __forceinline__ void s1_matvec(const float* __restrict__ m, float x, float y, float z,
                                          float& ox, float& oy, float& oz)
{
    ox = s1_row(m + 0, x, y, z);
    oy = s1_row(m + 4, x, y, z);
    oz = s1_row(m + 8, x, y, z);
}

// S5: length of (dx, dy, dz) as a double
__device__ __forceinline__ double s5_length(float dx, float dy, float dz)
{
    double acc = (double)__fmul_rn(dz, dz);
    acc = __dadd_rn(acc, (double)__fmul_rn(dy, dy));
    acc = __dadd_rn(acc, (double)__fmul_rn(dx, dx));
    return sqrt(acc);                                // IEEE-correct in double
}

// S3: c = a @ b (4x4, row-major)
__device__ inline void s3_matmul(const float* a, const float* b, float* c)
{
    float t[16];
    for (int r = 0; r < 4; ++r)
        for (int col = 0; col < 4; ++col) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k)
                acc = __dadd_rn(acc, (double)__fmul_rn(a[r * 4 + k], b[k * 4 + col]));
            t[r * 4 + col] = __double2float_rn(acc);
        }
    for (int i = 0; i < 16; ++i) c[i] = t[i];
}

__device__ __forceinline__ float det2f(float a, float b, float c, float d)
{
    return __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, c));
}
__device__ __forceinline__ float det3f(float a1, float a2, float a3, float b1, float b2, float b3,
                                       float c1, float c2, float c3)
{
    float t0 = __fmul_rn(a1, det2f(b2, b3, c2, c3));
    float t1 = __fmul_rn(b1, det2f(a2, a3, c2, c3));
    float t2 = __fmul_rn(c1, det2f(a2, a3, b2, b3));
    return __fadd_rn(__fsub_rn(t0, t1), t2);
}

// S2: fp32 adjugate/determinant inverse over the column-major view; returns false if det == 0.
__device__ inline bool s2_invert(const float* mrow, float* out_row)
{
    // column-major view: m[i][j] = element (row j, col i)
    float m[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m[i][j] = mrow[j * 4 + i];
    const float a1 = m[0][0], b1 = m[0][1], c1 = m[0][2], d1 = m[0][3];
    const float a2 = m[1][0], b2 = m[1][1], c2 = m[1][2], d2 = m[1][3];
    const float a3 = m[2][0], b3 = m[2][1], c3 = m[2][2], d3 = m[2][3];
    const float a4 = m[3][0], b4 = m[3][1], c4 = m[3][2], d4 = m[3][3];
    float det = __fmul_rn(a1, det3f(b2, b3, b4, c2, c3, c4, d2, d3, d4));
    det = __fsub_rn(det, __fmul_rn(b1, det3f(a2, a3, a4, c2, c3, c4, d2, d3, d4)));
    det = __fadd_rn(det, __fmul_rn(c1, det3f(a2, a3, a4, b2, b3, b4, d2, d3, d4)));
    det = __fsub_rn(det, __fmul_rn(d1, det3f(a2, a3, a4, b2, b3, b4, c2, c3, c4)));
    if (det == 0.0f) return false;
    float r[4][4];
    r[0][0] = det3f(b2, b3, b4, c2, c3, c4, d2, d3, d4);
    r[1][0] = -det3f(a2, a3, a4, c2, c3, c4, d2, d3, d4);
    r[2][0] = det3f(a2, a3, a4, b2, b3, b4, d2, d3, d4);
    r[3][0] = -det3f(a2, a3, a4, b2, b3, b4, c2, c3, c4);
    r[0][1] = -det3f(b1, b3, b4, c1, c3, c4, d1, d3, d4);
    r[1][1] = det3f(a1, a3, a4, c1, c3, c4, d1, d3, d4);
    r[2][1] = -det3f(a1, a3, a4, b1, b3, b4, d1, d3, d4);
    r[3][1] = det3f(a1, a3, a4, b1, b3, b4, c1, c3, c4);
    r[0][2] = det3f(b1, b2, b4, c1, c2, c4, d1, d2, d4);
    r[1][2] = -det3f(a1, a2, a4, c1, c2, c4, d1, d2, d4);
    r[2][2] = det3f(a1, a2, a4, b1, b2, b4, d1, d2, d4);
    r[3][2] = -det3f(a1, a2, a4, b1, b2, b4, c1, c2, c4);
    r[0][3] = -det3f(b1, b2, b3, c1, c2, c3, d1, d2, d3);
    r[1][3] = det3f(a1, a2, a3, c1, c2, c3, d1, d2, d3);
    r[2][3] = -det3f(a1, a2, a3, b1, b2, b3, d1, d2, d3);
    r[3][3] = det3f(a1, a2, a3, b1, b2, b3, c1, c2, c3);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out_row[j * 4 + i] = __fdiv_rn(r[i][j], det);
    return true;
}

// S4: the nearest-vertex metric (shared bit for bit with oracle/nn_oracle.c)
__device__ __forceinline__ float s4_d2(float qx, float qy, float qz, float bx, float by, float bz)
{
    float dx = __fsub_rn(qx, bx), dy = __fsub_rn(qy, by), dz = __fsub_rn(qz, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD by one-sided (Hestenes) Jacobi in fp64, then the proper rotation
//     R = U diag(1, 1, sign(det U det V)) V^T
// which equals the reference's `R = u @ vh; if det(R) < 0: R -= outer(u[:,2], 2 vh[2,:])`
// (functions/general.py:181-187) and does not depend on the sign LAPACK picks for u[:,2].
// h is row-major 3x3 (the cross-covariance sum_k (b_k - cb)(a_k - ca)^T, general.py:181).
// ---------------------------------------------------------------------------------------------
__device__ inline void rotation_from_covariance(const double* h, double* rot /*3x3 row-major*/)
{
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { a[i][j] = h[i * 3 + j]; v[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-300) continue;
                double lim = sqrt(alpha) * sqrt(beta);
                if (fabs(gamma) <= 1e-17 * lim) continue;
                off = fmax(off, fabs(gamma) / (lim > 0 ? lim : 1.0));
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    double ap = a[i][p], aq = a[i][q];
                    a[i][p] = c * ap - s * aq;
                    a[i][q] = s * ap + c * aq;
                    double vp = v[i][p], vq = v[i][q];
                    v[i][p] = c * vp - s * vq;
                    v[i][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-16) break;
    }
    double sig[3];
    for (int j = 0; j < 3; ++j)
        sig[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
    int o[3] = {0, 1, 2};                       // descending by singular value
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (sig[o[j]] < sig[o[j + 1]]) { int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }
    double u[3][3], w[3][3];
    for (int k = 0; k < 3; ++k) {
        int j = o[k];
        for (int i = 0; i < 3; ++i) {
            w[i][k] = v[i][j];
            u[i][k] = (sig[j] > 0.0) ? a[i][j] / sig[j] : 0.0;
        }
    }
    const double s0 = sig[o[0]], s1v = sig[o[1]], s2v = sig[o[2]];
    if (!(s0 > 0.0)) {                          // zero covariance: identity rotation
        for (int i = 0; i < 9; ++i) rot[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    if (s1v <= 1e-14 * s0) {
        // rank 1: complete u1 with any unit vector orthogonal to it (degenerate input; LAPACK's
        // completion is just as arbitrary).  Do the same for the right vectors.
        for (int pass = 0; pass < 2; ++pass) {
            double (*x)[3] = pass ? w : u;
            int mi = 0;
            for (int i = 1; i < 3; ++i) if (fabs(x[i][0]) < fabs(x[mi][0])) mi = i;
            double e[3] = {0, 0, 0}; e[mi] = 1.0;
            double d = e[0] * x[0][0] + e[1] * x[1][0] + e[2] * x[2][0];
            double y[3], n = 0;
            for (int i = 0; i < 3; ++i) { y[i] = e[i] - d * x[i][0]; n += y[i] * y[i]; }
            n = sqrt(n);
            for (int i = 0; i < 3; ++i) x[i][1] = y[i] / n;
        }
    }
    if (s2v <= 1e-14 * s0) {
        // rank <= 2 (planar data): u3 = u1 x u2, v3 = v1 x v2; the sign is irrelevant below
        u[0][2] = u[1][0] * u[2][1] - u[2][0] * u[1][1];
        u[1][2] = u[2][0] * u[0][1] - u[0][0] * u[2][1];
        u[2][2] = u[0][0] * u[1][1] - u[1][0] * u[0][1];
        w[0][2] = w[1][0] * w[2][1] - w[2][0] * w[1][1];
        w[1][2] = w[2][0] * w[0][1] - w[0][0] * w[2][1];
        w[2][2] = w[0][0] * w[1][1] - w[1][0] * w[0][1];
    }
    auto det3 = [](double (*x)[3]) {
        return x[0][0] * (x[1][1] * x[2][2] - x[1][2] * x[2][1])
             - x[0][1] * (x[1][0] * x[2][2] - x[1][2] * x[2][0])
             + x[0][2] * (x[1][0] * x[2][1] - x[1][1] * x[2][0]);
    };
    const double sgn = (det3(u) * det3(w) < 0.0) ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            rot[i * 3 + j] = u[i][0] * w[j][0] + u[i][1] * w[j][1] + sgn * u[i][2] * w[j][2];
}

}  // namespace icp
