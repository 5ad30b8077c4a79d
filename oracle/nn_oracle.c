/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the nearest-vertex query.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never does.
 *
 * What it restates: the per-vertex query `base_bvh.find_nearest(co_find)` at
 * /root/reference/functions/general.py:297, in the nearest-VERTEX variant that
 * SURVEY.md section 0 (D2) fixes for the graded path (the reference's own provider is
 * Blender's C BVHTree over triangles -- mathutils.bvhtree, Blender 3.2.2 per
 * /root/reference/__init__.py:22 -- which is absent from /root/reference).
 *
 * Metric (spec S4, shared bit for bit with the CUDA kernels):
 *     dx = q.x - b.x; dy = q.y - b.y; dz = q.z - b.z           (fp32, round-to-nearest)
 *     d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx))                 (fp32, two fused steps)
 * The winner is the smallest d2; ties go to the LOWEST base index.
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -mfma -ffp-contract=off -fopenmp).
 * -ffp-contract=off keeps the compiler from fusing anything we did not write as fmaf.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define LANES 8

static inline float d2_spec(float qx, float qy, float qz, float bx, float by, float bz)
{
    float dx = qx - bx, dy = qy - by, dz = qz - bz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* Exported scalar form so Python can evaluate the metric on a handful of candidates. */
float oracle_d2(const float *q, const float *b)
{
    return d2_spec(q[0], q[1], q[2], b[0], b[1], b[2]);
}

/* d2 of one query against a list of candidate base indices. */
void oracle_d2_candidates(const float *q, const float *base_xyz, const int64_t *cand, int64_t ncand,
                          float *out)
{
    for (int64_t i = 0; i < ncand; ++i) {
        const float *b = base_xyz + 3 * cand[i];
        out[i] = d2_spec(q[0], q[1], q[2], b[0], b[1], b[2]);
    }
}

/* Brute-force nearest vertex for M queries against N base points (AoS xyz, fp32).
 * index_offset is added to the reported indices (used to emulate base shards).
 * threads <= 0 -> all OpenMP threads.  Returns 0, or -1 on allocation failure. */
int oracle_nn_bruteforce(const float *q_xyz, int64_t M, const float *base_xyz, int64_t N,
                         int32_t index_offset, int32_t *idx_out, float *d2_out, int threads)
{
    if (N <= 0) {
        for (int64_t i = 0; i < M; ++i) {
            idx_out[i] = -1;
            d2_out[i] = INFINITY;
        }
        return 0;
    }
    /* SoA copy padded to a multiple of LANES with +inf so padded lanes never win. */
    int64_t Np = (N + LANES - 1) / LANES * LANES;
    float *bx = (float *)aligned_alloc(64, (size_t)Np * sizeof(float) * 3 + 64);
    if (!bx) return -1;
    float *by = bx + Np, *bz = by + Np;
    for (int64_t j = 0; j < N; ++j) {
        bx[j] = base_xyz[3 * j];
        by[j] = base_xyz[3 * j + 1];
        bz[j] = base_xyz[3 * j + 2];
    }
    for (int64_t j = N; j < Np; ++j) bx[j] = by[j] = bz[j] = INFINITY;

#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < M; ++i) {
        const float qx = q_xyz[3 * i], qy = q_xyz[3 * i + 1], qz = q_xyz[3 * i + 2];
        float best[LANES];
        int32_t bidx[LANES];
        for (int l = 0; l < LANES; ++l) {
            best[l] = INFINITY;
            bidx[l] = 0x7fffffff;
        }
        for (int64_t j = 0; j < Np; j += LANES) {
#pragma omp simd
            for (int l = 0; l < LANES; ++l) {
                float d = d2_spec(qx, qy, qz, bx[j + l], by[j + l], bz[j + l]);
                /* strict < keeps the lowest index inside a lane; NaN (inf-inf) never wins */
                int take = d < best[l];
                best[l] = take ? d : best[l];
                bidx[l] = take ? (int32_t)(j + l) : bidx[l];
            }
        }
        float b = best[0];
        int32_t bi = bidx[0];
        for (int l = 1; l < LANES; ++l) {
            if (best[l] < b || (best[l] == b && bidx[l] < bi)) {
                b = best[l];
                bi = bidx[l];
            }
        }
        if (bi == 0x7fffffff) { /* every distance was NaN/inf */
            idx_out[i] = -1;
            d2_out[i] = INFINITY;
        } else {
            idx_out[i] = bi + index_offset;
            d2_out[i] = b;
        }
    }
    free(bx);
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
