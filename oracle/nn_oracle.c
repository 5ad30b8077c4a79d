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

/* ================================================================================================
 * Surface variant: closest point on the TRIANGULATED SURFACE, i.e.
 * what the reference really asks of Blender: BVHTree.FromObject(base_obj, depsgraph)
 * (/root/reference/operators/icp_align.py:53) and base_bvh.find_nearest(co_find)
 * (/root/reference/functions/general.py:297).  Blender's C BVH (mathutils.bvhtree, Blender 3.2.2) is
 * not in /root/reference; restated here is its published per-triangle arithmetic:
 *   S9   closest_on_tri_to_point_v3 -- Ericson's region walk, every operation a separate fp32
 *        rounding (this file is compiled with -ffp-contract=off)
 *   S10  len_squared_v3v3           -- (dx*dx + dy*dy) + dz*dz
 * The winner is the smallest S10 distance; ties go to the LOWEST triangle index (Blender's own
 * tie-break depends on its tree layout).
 * ================================================================================================ */
static inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

void oracle_closest_on_tri(const float *p, const float *a, const float *b, const float *c, float *r)
{
    float ab[3], ac[3], ap[3], bp[3], cp[3];
    for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; return; }
    for (int i = 0; i < 3; ++i) bp[i] = p[i] - b[i];
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { r[0] = b[0]; r[1] = b[1]; r[2] = b[2]; return; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        const float v = d1 / (d1 - d3);
        for (int i = 0; i < 3; ++i) r[i] = a[i] + ab[i] * v;
        return;
    }
    for (int i = 0; i < 3; ++i) cp[i] = p[i] - c[i];
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { r[0] = c[0]; r[1] = c[1]; r[2] = c[2]; return; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        const float w = d2 / (d2 - d6);
        for (int i = 0; i < 3; ++i) r[i] = a[i] + ac[i] * w;
        return;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; ++i) { float t = c[i] - b[i]; t = t * w; r[i] = t + b[i]; }
        return;
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    for (int i = 0; i < 3; ++i) {
        const float acw = ac[i] * w;
        float t = a[i] + ab[i] * v;
        r[i] = t + acw;
    }
}

float oracle_len_squared(const float *p, const float *r)
{
    const float d[3] = {r[0] - p[0], r[1] - p[1], r[2] - p[2]};
    return dot3(d, d);
}

/* Brute force over all triangles (tri_xyz: F x 9 floats).  co_out (nullable): M x 3 closest points. */
int oracle_nn_surface_bruteforce(const float *q_xyz, int64_t M, const float *tri_xyz, int64_t F,
                                 int32_t index_offset, int32_t *idx_out, float *d2_out, float *co_out,
                                 int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < M; ++i) {
        const float *p = q_xyz + 3 * i;
        float best = INFINITY, br[3] = {0, 0, 0};
        int64_t bi = -1;
        for (int64_t f = 0; f < F; ++f) {
            const float *t = tri_xyz + 9 * f;
            float r[3];
            oracle_closest_on_tri(p, t, t + 3, t + 6, r);
            const float d = oracle_len_squared(p, r);
            if (d < best) { best = d; bi = f; br[0] = r[0]; br[1] = r[1]; br[2] = r[2]; }
        }
        idx_out[i] = bi < 0 ? -1 : (int32_t)(bi + index_offset);
        d2_out[i] = best;
        if (co_out) { co_out[3 * i] = br[0]; co_out[3 * i + 1] = br[1]; co_out[3 * i + 2] = br[2]; }
    }
    return 0;
}
