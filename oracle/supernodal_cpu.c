/*
 * supernodal_cpu.c -- CPU BASELINE (test/bench infrastructure, NOT product code).
 *
 * A multithreaded, BLAS-3 supernodal multifrontal Cholesky on the host cores: the stand-in for the reference's
 * CHOLMOD path (`cholesky!(F, S; check=false)`, /root/reference/src/workspace/backend.jl:184), which cannot be
 * run in this image (no Julia, no libcholmod). It does what CHOLMOD's supernodal numeric phase does -- dense
 * POTRF / TRSM / SYRK on supernodal fronts with assembly through relative indices -- with LAPACK/BLAS kernels
 * taken from the OpenBLAS that ships inside SciPy (function pointers are passed in from Python, see
 * oracle/cpu_baseline.py), OpenMP across independent fronts of a level and threaded BLAS inside the big fronts.
 * It is only ever TIMED (bench.py cpu_baseline / --impl reference) and cross-checked against the simplicial oracle
 * in tests/; it is never linked into libgmrf_b200.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef void (*dpotrf_t)(char *, int *, double *, int *, int *);
typedef void (*dtrsm_t)(char *, char *, char *, char *, int *, int *, double *, double *, int *, double *, int *);
typedef void (*dsyrk_t)(char *, char *, int *, int *, double *, double *, int *, double *, double *, int *);

typedef struct {
    const i64 *super_ptr, *row_ptr, *row_idx, *rel_idx, *panel_off, *panel_ld, *upd_off, *upd_ld, *child_ptr, *child_idx;
    double *Lx, *upd;
    dpotrf_t potrf;
    dtrsm_t trsm;
    dsyrk_t syrk;
} cpu_ctx;

static int factor_one(const cpu_ctx *c, i64 s)
{
    const i64 ns = c->super_ptr[s + 1] - c->super_ptr[s];
    const i64 nrow = c->row_ptr[s + 1] - c->row_ptr[s];
    const i64 nr = nrow - ns, ld = c->panel_ld[s], uld = c->upd_ld[s];
    double *P = c->Lx + c->panel_off[s];
    double *U = c->upd + c->upd_off[s];
    for (i64 j = 0; j < nr; j++) memset(U + j * uld + j, 0, sizeof(double) * (size_t)(nr - j));
    for (i64 ci = c->child_ptr[s]; ci < c->child_ptr[s + 1]; ci++) {
        const i64 ch = c->child_idx[ci];
        const i64 cns = c->super_ptr[ch + 1] - c->super_ptr[ch];
        const i64 cnr = (c->row_ptr[ch + 1] - c->row_ptr[ch]) - cns, culd = c->upd_ld[ch];
        const i64 *rel = c->rel_idx + c->row_ptr[ch] + cns;
        const double *Uc = c->upd + c->upd_off[ch];
        for (i64 jc = 0; jc < cnr; jc++) {
            const i64 pc = rel[jc];
            const double *src = Uc + jc * culd;
            if (pc < ns) {
                double *dst = P + pc * ld;
                for (i64 ic = jc; ic < cnr; ic++) dst[rel[ic]] += src[ic];
            } else {
                double *dst = U + (pc - ns) * uld - ns;
                for (i64 ic = jc; ic < cnr; ic++) dst[rel[ic]] += src[ic];
            }
        }
    }
    int n_ = (int)ns, ld_ = (int)ld, info = 0;
    c->potrf("L", &n_, P, &ld_, &info);
    if (info != 0) return (int)(c->super_ptr[s] + info);
    if (nr > 0) {
        int m_ = (int)nr, uld_ = (int)uld;
        double one = 1.0, mone = -1.0;
        c->trsm("R", "L", "T", "N", &m_, &n_, &one, P, &ld_, P + ns, &ld_);
        c->syrk("L", "N", &m_, &n_, &mone, P + ns, &ld_, &one, U, &uld_);
    }
    return 0;
}

/* Factor the supernodes supers[0..count) of one level. parallel != 0: OpenMP over fronts (BLAS must be limited
 * to one thread by the caller); parallel == 0: fronts one after another, BLAS threaded. Returns 0 or k+1. */
int cpu_supernodal_factor_level(i64 count, const i64 *supers, const i64 *super_ptr, const i64 *row_ptr,
                                const i64 *row_idx, const i64 *rel_idx, const i64 *panel_off, const i64 *panel_ld,
                                const i64 *upd_off, const i64 *upd_ld, const i64 *child_ptr, const i64 *child_idx,
                                double *Lx, double *upd, void *potrf, void *trsm, void *syrk, int parallel)
{
    cpu_ctx c = {super_ptr, row_ptr, row_idx, rel_idx, panel_off, panel_ld, upd_off, upd_ld, child_ptr, child_idx,
                 Lx, upd, (dpotrf_t)potrf, (dtrsm_t)trsm, (dsyrk_t)syrk};
    int status = 0;
    if (parallel) {
#pragma omp parallel for schedule(dynamic, 1)
        for (i64 t = 0; t < count; t++) {
            int r = factor_one(&c, supers[t]);
            if (r) {
#pragma omp critical
                if (status == 0 || r < status) status = r;
            }
        }
    } else {
        for (i64 t = 0; t < count; t++) {
            int r = factor_one(&c, supers[t]);
            if (r && (status == 0 || r < status)) status = r;
        }
    }
    return status;
}

/* Lx[dst[k]] = nz[src[k]] after zeroing the panels */
void cpu_scatter(i64 total, double *Lx, i64 cnt, const i64 *src, const i64 *dst, const double *nz)
{
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < total; k++) Lx[k] = 0.0;
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < cnt; k++) Lx[dst[k]] = nz[src[k]];
}

double cpu_logdet(i64 n, const i64 *diag_pos, const double *Lx)
{
    double acc = 0.0;
    for (i64 j = 0; j < n; j++) acc += log(Lx[diag_pos[j]]);
    return 2.0 * acc;
}

int cpu_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
