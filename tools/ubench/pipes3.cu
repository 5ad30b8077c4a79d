// Micro-benchmark 3: what does the q16 kernel's hot instruction cost on sm_100a?
//   VIADDMNMX.U16x2 (DPX: packed 16-bit add + min) alone, next to VIMNMX3.U16x2, and the mixes that
//   could move part of the fold to the fma pipe (IMAD-add + VIMNMX3).  Every value is loop-carried;
//   check the SASS with cuobjdump before trusting a number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes3 pipes3.cu && ./pipes3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float min3f(float a, float b, float c){float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}

#define ITER 2048
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, const uint32_t* in, long long* cyc)
{
    uint32_t a[NCH], b[NCH], c[NCH];
    float fa[NCH], fb[NCH], fc[NCH];
    for (int i = 0; i < NCH; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i];
        fa[i] = __uint_as_float(a[i] & 0x3fffffffu); fb[i] = __uint_as_float(b[i] & 0x3fffffffu); fc[i] = __uint_as_float(c[i] & 0x3fffffffu);
    }
    long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) {                 // 2 VIADDMNMX.U16x2 (the q16 hot loop: acc = min(v + nlo, acc))
                a[i] = __viaddmin_u16x2(b[i], c[i], a[i]);
                b[i] = __viaddmin_u16x2(c[i], a[i], b[i]);
            } else if (MODE == 1) {          // 2 VIMNMX3.U16x2
                a[i] = __vimin3_u16x2(a[i], b[i], c[i]);
                b[i] = __vimin3_u16x2(b[i], a[i], c[i]);
            } else if (MODE == 2) {          // 2 IMAD-adds (fma pipe) + 1 VIMNMX3.U16x2: four vertices per 3 instructions
                uint32_t x, y;
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(x) : "r"(b[i]), "r"(c[i]));
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(y) : "r"(c[i]), "r"(a[i]));
                a[i] = __vimin3_u16x2(a[i], x, y);
            } else if (MODE == 3) {          // 2 VIADDMNMX + (2 IMAD + 1 VIMNMX3): the balanced mix
                a[i] = __viaddmin_u16x2(b[i], c[i], a[i]);
                c[i] = __viaddmin_u16x2(c[i], a[i], b[i]);
                uint32_t x, y;
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(x) : "r"(b[i]), "r"(a[i]));
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(y) : "r"(c[i]), "r"(a[i]));
                b[i] = __vimin3_u16x2(b[i], x, y);
            } else if (MODE == 4) {          // 2 VIADDMNMX.U32 (is the packed form slower than the plain one?)
                a[i] = __viaddmin_u32(b[i], c[i], a[i]);
                b[i] = __viaddmin_u32(c[i], a[i], b[i]);
            } else if (MODE == 5) {          // 2 FADD + 1 FMNMX3 with |abs| (fp32 slab without the FMA)
                const float t0 = fb[i] - fa[i], t1 = fc[i] - fa[i];
                fa[i] = min3f(fa[i], fabsf(t0), fabsf(t1));
            } else if (MODE == 6) {          // 2 VIADDMNMX + 1 FFMA (does fp32 work ride along for free?)
                a[i] = __viaddmin_u16x2(b[i], c[i], a[i]);
                b[i] = __viaddmin_u16x2(c[i], a[i], b[i]);
                fa[i] = fmaf(fa[i], fb[i], fc[i]);
            } else if (MODE == 7) {          // 4 VIADDMNMX + 1 LDS.128 (the real ratio at R = 1; smem operand)
                extern __shared__ uint4 sm[];
                const uint4 v = sm[(it * NCH + i) & 255];
                a[i] = __viaddmin_u16x2(v.x, c[i], a[i]);
                a[i] = __viaddmin_u16x2(v.y, c[i], a[i]);
                a[i] = __viaddmin_u16x2(v.z, c[i], a[i]);
                a[i] = __viaddmin_u16x2(v.w, c[i], a[i]);
            } else if (MODE == 8) {          // 2 VIADD.16x2 + 1 VIMNMX3 (what ptxas makes of two viaddmin with a shared sum)
                const uint32_t x = __vadd2(b[i], c[i]), y = __vadd2(c[i], a[i]);
                a[i] = __vimin3_u16x2(a[i], x, y);
            } else if (MODE == 9) {          // 2 VIADD.16x2 alone
                a[i] = __vadd2(a[i], b[i]);
                b[i] = __vadd2(b[i], c[i]);
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < NCH; ++i) s += a[i] + b[i] + c[i] + __float_as_uint(fa[i]) + __float_as_uint(fb[i]) + __float_as_uint(fc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_step, uint32_t* out, uint32_t* in, long long* cyc, int threads)
{
    k<MODE><<<148, threads, 4096>>>(out, in, cyc);
    k<MODE><<<148, threads, 4096>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    double per_step = (double)h / ((double)ITER * NCH * warps_per_smsp);   // cycles per chain step per SMSP
    printf("%-72s %2d w/SMSP  %6.2f cycles per step  (%5.2f per instr, %d instr)\n", name, warps_per_smsp, per_step,
           per_step / instr_per_step, instr_per_step);
}

int main()
{
    uint32_t *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 8);
    uint32_t hin[4096];
    for (int i = 0; i < 4096; ++i) hin[i] = (uint32_t)i * 2654435761u;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int threads = 512; threads >= 128; threads /= 2) {
        run<0>("2 VIADDMNMX.U16x2", 2, out, in, cyc, threads);
        run<1>("2 VIMNMX3.U16x2", 2, out, in, cyc, threads);
        run<2>("2 IMAD(add) + 1 VIMNMX3.U16x2 (4 vertices)", 3, out, in, cyc, threads);
        run<3>("2 VIADDMNMX + 2 IMAD + 1 VIMNMX3 (8 vertices)", 5, out, in, cyc, threads);
        run<4>("2 VIADDMNMX.U32", 2, out, in, cyc, threads);
        run<5>("2 FADD + 1 FMNMX3|abs| (2 vertices)", 3, out, in, cyc, threads);
        run<6>("2 VIADDMNMX + 1 FFMA", 3, out, in, cyc, threads);
        run<7>("4 VIADDMNMX + 1 LDS.128", 5, out, in, cyc, threads);
        run<8>("2 VIADD.16x2 + 1 VIMNMX3.U16x2 (4 vertices)", 3, out, in, cyc, threads);
        run<9>("2 VIADD.16x2", 2, out, in, cyc, threads);
    }
    return 0;
}
