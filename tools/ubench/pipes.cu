// Micro-benchmark: issue cost of the instructions of the brute-force NN inner loop on sm_100a.
// Each kernel runs ITER iterations of an unrolled body on 8 independent chains per thread;
// reports cycles per warp-instruction per SMSP (1024 threads/SM = 8 warps per SMSP resident).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi){uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi){asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ float min3f(float a, float b, float c){float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}
__device__ __forceinline__ int imin3(int a, int b, int c){int r; asm("min.s32 %0, %1, %2;\n\tmin.s32 %0, %0, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;}

#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float* in, long long* cyc)
{
    __shared__ __align__(16) float sm[1024];
    sm[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    float a[8], b[8], c[8];
    uint64_t A[8], B[8], C[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i];
        A[i] = pk2(a[i], b[i]); B[i] = pk2(b[i], c[i]); C[i] = pk2(c[i], a[i]);
    }
    const float u0 = in[blockIdx.x], u1 = in[blockIdx.x + 1];      // warp-uniform values
    const uint64_t U = pk2(u0, u1);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {            // FFMA scalar, 8 chains
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b[i], c[i]);
        } else if (MODE == 1) {     // FFMA2 all-register, 8 chains
#pragma unroll
            for (int i = 0; i < 8; ++i) A[i] = fma2(A[i], B[i], C[i]);
        } else if (MODE == 2) {     // FFMA2 with a broadcast scalar multiplicand (R.F32 form)
#pragma unroll
            for (int i = 0; i < 8; ++i) A[i] = fma2(pk2(b[i], b[i]), B[i], A[i]);
        } else if (MODE == 3) {     // FFMA2 broadcast scalar x uniform pair (UR form?) + acc
#pragma unroll
            for (int i = 0; i < 8; ++i) A[i] = fma2(pk2(b[i], b[i]), U, A[i]);
        } else if (MODE == 4) {     // FMNMX3, 8 chains
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = min3f(a[i], b[i], c[i]);
        } else if (MODE == 5) {     // FMNMX (2-input)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fminf(a[i], b[i]);
        } else if (MODE == 6) {     // integer min3 on the bits
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __int_as_float(imin3(__float_as_int(a[i]), __float_as_int(b[i]), __float_as_int(c[i])));
        } else if (MODE == 7) {     // 2 FFMA2 (bcast x reg) : 1 FMNMX3  -- the cylinder loop's mix
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint64_t e = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(e, e0, e1);
                a[i] = min3f(a[i], e0, e1);
            }
        } else if (MODE == 8) {     // same with integer min3
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint64_t e = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(e, e0, e1);
                a[i] = __int_as_float(imin3(__float_as_int(a[i]), __float_as_int(e0), __float_as_int(e1)));
            }
        } else if (MODE == 9) {     // 3 FFMA2 : 1 FMNMX3 -- the expansion filter's mix
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint64_t e = fma2(pk2(a[i], a[i]), A[i], fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i])));
                float e0, e1; unpk2(e, e0, e1);
                c[i] = min3f(c[i], e0, e1);
            }
        } else if (MODE == 10) {    // LDS.128 broadcast + 2 FFMA2 per pair of points x 4 queries + FMNMX3 (tile operands from smem)
            const float4 X = reinterpret_cast<const float4*>(sm)[it & 63];
            const float4 Y = reinterpret_cast<const float4*>(sm)[64 + (it & 63)];
            const float4 W = reinterpret_cast<const float4*>(sm)[128 + (it & 63)];
            const uint64_t x01 = pk2(X.x, X.y), x23 = pk2(X.z, X.w), y01 = pk2(Y.x, Y.y), y23 = pk2(Y.z, Y.w),
                           w01 = pk2(W.x, W.y), w23 = pk2(W.z, W.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float e0, e1, e2, e3;
                unpk2(fma2(pk2(b[i], b[i]), y01, fma2(pk2(a[i], a[i]), x01, w01)), e0, e1);
                unpk2(fma2(pk2(b[i], b[i]), y23, fma2(pk2(a[i], a[i]), x23, w23)), e2, e3);
                c[i] = min3f(c[i], e0, e1);
                c[i] = min3f(c[i], e2, e3);
            }
        } else if (MODE == 11) {    // FADD2-like: fma2 with 1.0 (packed add) all-register
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(A[i]), "l"(B[i])); A[i] = r; }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float l, h; unpk2(A[i], l, h); s += a[i] + c[i] + l + h; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, float* out, float* in, long long* cyc)
{
    k<MODE><<<148, 1024>>>(out, in, cyc);
    k<MODE><<<148, 1024>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 8 warps per SMSP resident, each issues instr_per_iter * ITER instructions
    double per = (double)h / ((double)ITER * instr_per_iter * 8);
    printf("%-52s %6.2f cycles per warp-instruction per SMSP   (%lld cycles)\n", name, per, h);
}

int main()
{
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 4);
    run<0>("FFMA  R,R,R", 8, out, in, cyc);
    run<1>("FFMA2 R2,R2,R2", 8, out, in, cyc);
    run<2>("FFMA2 R.F32(bcast),R2,R2", 8, out, in, cyc);
    run<3>("FFMA2 R.F32(bcast),uniform pair,R2", 8, out, in, cyc);
    run<4>("FMNMX3", 8, out, in, cyc);
    run<5>("FMNMX", 8, out, in, cyc);
    run<6>("IMNMX x2 (int min3)", 16, out, in, cyc);
    run<7>("mix 2 FFMA2 + 1 FMNMX3 (per 3 instr)", 24, out, in, cyc);
    run<8>("mix 2 FFMA2 + 2 IMNMX (per 4 instr)", 32, out, in, cyc);
    run<9>("mix 3 FFMA2 + 1 FMNMX3 (per 4 instr)", 32, out, in, cyc);
    run<10>("cyl body: 3 LDS.128 + 16 FFMA2 + 8 FMNMX3 (per 27)", 27, out, in, cyc);
    run<11>("FADD2 R2,R2", 8, out, in, cyc);
    return 0;
}
