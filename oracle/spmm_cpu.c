/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under pla-gnn_b200/ may link or call this.
 *
 * CPU restatement of the sparse aggregation that PLA-GNN reaches through DGL:
 *   model.py:13-15,20,22,24  SAGEConv(.., 'pool')  ->  update_all(copy_u, max)
 *   (DGL 0.8.2 SpMMCsr on the in-edge CSR; arithmetic lives in the un-vendored
 *    third-party wheel dgl_cu113 0.8.2.post1, README.md:27 — restated from its published
 *    semantics: row-parallel loop over destination nodes, strict '>' compare so the FIRST
 *    maximum in in-edge order wins, rows with no in-edge yield 0.)
 * Parity status: UNPINNED for this file (DGL cannot be imported here).
 *
 * Also holds the copy_u/u_mul_e + sum reducers used by the synthetic throughput configs
 * (BASELINE.json configs[3]) and the reverse (scatter) passes that DGL's autograd runs.
 *
 * Built by oracle/Makefile into oracle/_build/liboracle_spmm.so (OpenMP, host cores).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* out[v,:] = max_{e in [indptr[v],indptr[v+1])} x[indices[e],:]; arg = winning source id (-1 if none) */
void oracle_spmm_max_f32(int64_t n_dst, int64_t feat,
                         const int32_t *indptr, const int32_t *indices,
                         const float *x, int64_t ldx,
                         float *out, int32_t *arg, int64_t ldo)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t v = 0; v < n_dst; ++v) {
        float *o = out + v * ldo;
        int32_t *a = arg + v * ldo;
        const int32_t beg = indptr[v], end = indptr[v + 1];
        if (beg == end) {
            for (int64_t f = 0; f < feat; ++f) { o[f] = 0.0f; a[f] = -1; }
            continue;
        }
        for (int64_t f = 0; f < feat; ++f) { o[f] = -INFINITY; a[f] = -1; }
        for (int32_t e = beg; e < end; ++e) {
            const int32_t u = indices[e];
            const float *xr = x + (int64_t)u * ldx;
            for (int64_t f = 0; f < feat; ++f) {
                if (xr[f] > o[f]) { o[f] = xr[f]; a[f] = u; }
            }
        }
        for (int64_t f = 0; f < feat; ++f) if (a[f] < 0) o[f] = 0.0f; /* all -inf/NaN column */
    }
}

/* reverse of the max reducer: dx[arg[v,f], f] += dz[v,f].  Column-block parallel so the
 * accumulation order over v is the sequential one for every (u,f): deterministic. */
void oracle_spmm_max_bwd_f32(int64_t n_dst, int64_t n_src, int64_t feat,
                             const float *dz, const int32_t *arg, int64_t ldz,
                             float *dx, int64_t ldx)
{
#pragma omp parallel
    {
#ifdef _OPENMP
        const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
        const int nt = 1, tid = 0;
#endif
        const int64_t f0 = feat * tid / nt, f1 = feat * (tid + 1) / nt;
        for (int64_t u = 0; u < n_src; ++u)
            for (int64_t f = f0; f < f1; ++f) dx[u * ldx + f] = 0.0f;
        for (int64_t v = 0; v < n_dst; ++v)
            for (int64_t f = f0; f < f1; ++f) {
                const int32_t u = arg[v * ldz + f];
                if (u >= 0) dx[(int64_t)u * ldx + f] += dz[v * ldz + f];
            }
    }
}

/* out[v,:] = scale[v] * sum_e w[eid(e)] * x[indices[e],:]   (w, eids, scale optional)
 * Used forward (in-edge CSR) and, on the out-edge CSR, as the exact transpose for backward. */
void oracle_spmm_sum_f32(int64_t n_dst, int64_t feat,
                         const int32_t *indptr, const int32_t *indices,
                         const int32_t *eids, const float *w, const float *scale,
                         const float *x, int64_t ldx, float *out, int64_t ldo)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t v = 0; v < n_dst; ++v) {
        float *o = out + v * ldo;
        for (int64_t f = 0; f < feat; ++f) o[f] = 0.0f;
        for (int32_t e = indptr[v]; e < indptr[v + 1]; ++e) {
            const float *xr = x + (int64_t)indices[e] * ldx;
            if (w) {
                const float we = w[eids ? eids[e] : e];
                for (int64_t f = 0; f < feat; ++f) o[f] += we * xr[f];
            } else {
                for (int64_t f = 0; f < feat; ++f) o[f] += xr[f];
            }
        }
        if (scale) { const float s = scale[v]; for (int64_t f = 0; f < feat; ++f) o[f] *= s; }
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
