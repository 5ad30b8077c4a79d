// Micro-benchmark 2: do the fma pipe (FFMA2 / HFMA2) and the alu pipe (FMNMX*, IMNMX, VIMNMX3) overlap
// on sm_100a, and what do the min instructions cost?  Every value computed in the loop is loop-carried
// (nothing can be hoisted); the SASS of each mode is checked with cuobjdump before trusting a number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu && ./pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi){uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;}
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi){asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ float min3f(float a, float b, float c){float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}

#define ITER 2048
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, const float* in, long long* cyc)
{
    float a[NCH], b[NCH], c[NCH];
    uint64_t A[NCH], B[NCH], C[NCH];
    uint32_t ua[NCH];
    __half2 ha[NCH], hb[NCH], hc[NCH], hm[NCH];
    for (int i = 0; i < NCH; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i];
        A[i] = pk2(a[i], b[i]); B[i] = pk2(b[i], c[i]); C[i] = pk2(c[i], a[i]);
        ua[i] = __float_as_uint(a[i]);
        ha[i] = __floats2half2_rn(a[i], b[i]); hb[i] = __floats2half2_rn(b[i], c[i]); hc[i] = __floats2half2_rn(c[i], a[i]);
        hm[i] = ha[i];
    }
    long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) {                 // FFMA2 bcast x2 (the cylinder filter's two FMAs), loop-carried
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
            } else if (MODE == 1) {          // FMNMX3 x1
                a[i] = min3f(a[i], b[i], c[i]); b[i] = min3f(b[i], a[i], c[i]);
            } else if (MODE == 2) {          // 2 FFMA2 + 1 FMNMX3  (current hot loop mix)
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(A[i], e0, e1);
                a[i] = min3f(a[i], e0, e1);
            } else if (MODE == 3) {          // 2 FFMA2 + 2 FMNMX
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(A[i], e0, e1);
                a[i] = fminf(fminf(a[i], e0), e1);
            } else if (MODE == 4) {          // 2 FFMA2 + 2 IMNMX (unsigned max on the bits)
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(A[i], e0, e1);
                ua[i] = max(max(ua[i], __float_as_uint(e0)), __float_as_uint(e1));
            } else if (MODE == 5) {          // 2 FFMA2 + 1 VIMNMX3 (unsigned max3 on the bits)
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(A[i], e0, e1);
                ua[i] = __vimax3_u32(ua[i], __float_as_uint(e0), __float_as_uint(e1));
            } else if (MODE == 6) {          // VIMNMX3 alone (2 per chain step)
                ua[i] = __vimax3_u32(ua[i], __float_as_uint(b[i]), __float_as_uint(c[i]));
                ua[i] = __vimax3_u32(ua[i], __float_as_uint(c[i]), __float_as_uint(a[i]));
            } else if (MODE == 7) {          // 4 scalar FFMA + 1 FMNMX3 (same work as mode 2 on the scalar pipes)
                float e0 = fmaf(b[i], b[(i + 1) % NCH], fmaf(c[i], c[(i + 1) % NCH], a[i]));
                float e1 = fmaf(b[i], c[(i + 2) % NCH], fmaf(c[i], b[(i + 2) % NCH], a[i]));
                a[i] = min3f(a[i], e0, e1);
            } else if (MODE == 8) {          // 2 HFMA2 + 1 HMNMX2: the same two FMAs + min on 2 points per 32-bit register
                ha[i] = __hfma2(hb[i], hb[(i + 1) % NCH], __hfma2(hc[i], hc[(i + 1) % NCH], ha[i]));
                hm[i] = __hmin2(hm[i], ha[i]);
            } else if (MODE == 9) {          // HFMA2 alone x2
                ha[i] = __hfma2(hb[i], hb[(i + 1) % NCH], __hfma2(hc[i], hc[(i + 1) % NCH], ha[i]));
            } else if (MODE == 10) {         // HMNMX2 alone x2
                hm[i] = __hmin2(hm[i], hb[i]); ha[i] = __hmin2(ha[i], hm[i]);
            } else if (MODE == 11) {         // 1 FFMA2 + 2 scalar FFMA: does fmalite add throughput next to FFMA2?
                A[i] = fma2(pk2(b[i], b[i]), B[i], A[i]);
                a[i] = fmaf(a[i], b[i], c[i]);
                c[i] = fmaf(c[i], b[i], a[i]);
            } else if (MODE == 12) {         // scalar FFMA alone x2 (distinct banks permitting)
                a[i] = fmaf(a[i], b[i], c[i]);
                c[i] = fmaf(c[i], b[i], a[i]);
            } else if (MODE == 13) {         // 2 FFMA2 + 1 FMNMX (2-input: min of the two lanes only, then folded later)
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                float e0, e1; unpk2(A[i], e0, e1);
                a[i] = fminf(e0, e1);
            } else if (MODE == 15) {         // slab filter mix: 1 FFMA2 + 1 FMNMX3 per two pairs
                A[i] = fma2(pk2(b[i], b[i]), B[i], A[i]);
                float e0, e1; unpk2(A[i], e0, e1);
                a[i] = min3f(a[i], e0, e1);
            } else if (MODE == 16) {         // slab mix, non-accumulating FMA (fresh addend each time, like w1 from the tile)
                uint64_t e = fma2(pk2(b[i], b[i]), B[i], C[i]);
                float e0, e1; unpk2(e, e0, e1);
                a[i] = min3f(a[i], e0, e1);
                C[i] = e;
            } else if (MODE == 17) {         // slab mix with 2-input mins: 1 FFMA2 + 2 FMNMX
                A[i] = fma2(pk2(b[i], b[i]), B[i], A[i]);
                float e0, e1; unpk2(A[i], e0, e1);
                a[i] = fminf(a[i], e0); c[i] = fminf(c[i], e1);
            } else if (MODE == 14) {         // 4 FFMA2 + 1 FFMA2-packed-min substitute: FMNMX3 on every second pair only
                A[i] = fma2(pk2(b[i], b[i]), B[i], fma2(pk2(c[i], c[i]), C[i], A[i]));
                B[i] = fma2(pk2(b[i], b[i]), A[i], fma2(pk2(c[i], c[i]), C[i], B[i]));
                float e0, e1; unpk2(B[i], e0, e1);
                a[i] = min3f(a[i], e0, e1);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < NCH; ++i) {
        float l, h; unpk2(A[i], l, h); float l2, h2; unpk2(B[i], l2, h2);
        s += a[i] + b[i] + c[i] + l + h + l2 + h2 + __uint_as_float(ua[i]) + __low2float(ha[i]) + __high2float(ha[i]) +
             __low2float(hm[i]) + __high2float(hm[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_step, float* out, float* in, long long* cyc, int threads)
{
    k<MODE><<<148, threads>>>(out, in, cyc);
    k<MODE><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    double per_step = (double)h / ((double)ITER * NCH * warps_per_smsp);   // cycles per chain step per SMSP
    printf("%-64s %2d w/SMSP  %6.2f cycles per step  (%5.2f per instr, %d instr)\n", name, warps_per_smsp, per_step,
           per_step / instr_per_step, instr_per_step);
}

int main()
{
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 8);
    float hin[4096];
    for (int i = 0; i < 4096; ++i) hin[i] = 0.25f + 0.5f * (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.0f;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int threads = 512; threads >= 256; threads -= 256) {
        run<0>("2 FFMA2(bcast)", 2, out, in, cyc, threads);
        run<1>("2 FMNMX3", 2, out, in, cyc, threads);
        run<2>("2 FFMA2 + 1 FMNMX3 (current mix)", 3, out, in, cyc, threads);
        run<3>("2 FFMA2 + 2 FMNMX", 4, out, in, cyc, threads);
        run<4>("2 FFMA2 + 2 IMNMX.U32", 4, out, in, cyc, threads);
        run<5>("2 FFMA2 + 1 VIMNMX3.U32", 3, out, in, cyc, threads);
        run<6>("2 VIMNMX3.U32", 2, out, in, cyc, threads);
        run<7>("4 FFMA + 1 FMNMX3 (scalar)", 5, out, in, cyc, threads);
        run<8>("2 HFMA2 + 1 HMNMX2 (two points per register)", 3, out, in, cyc, threads);
        run<9>("2 HFMA2", 2, out, in, cyc, threads);
        run<10>("2 HMNMX2", 2, out, in, cyc, threads);
        run<11>("1 FFMA2 + 2 FFMA", 3, out, in, cyc, threads);
        run<12>("2 FFMA", 2, out, in, cyc, threads);
        run<13>("2 FFMA2 + 1 FMNMX", 3, out, in, cyc, threads);
        run<14>("4 FFMA2 + 1 FMNMX3", 5, out, in, cyc, threads);
        run<15>("1 FFMA2 + 1 FMNMX3 (slab mix, in-place)", 2, out, in, cyc, threads);
        run<16>("1 FFMA2 + 1 FMNMX3 (slab mix, fresh addend)", 2, out, in, cyc, threads);
        run<17>("1 FFMA2 + 2 FMNMX", 3, out, in, cyc, threads);
    }
    return 0;
}
