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
 * in tests/; it is never linked into libgmrf_b200.so. The second half of the file is the matching supernodal selected
 * inversion (the CPU stand-in for SelectedInversion.selinv), so that `selinv ms` has a host-core number beside it too.
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

/* ------------------------------------------------------------------------------------------------------------------
 * Supernodal Takahashi selected inversion on the host cores (the CPU stand-in for SelectedInversion.selinv on a
 * supernodal CHOLMOD factor, /root/reference/src/workspace/backend.jl:226-236), top-down over the assembly tree:
 *   W  = Z[R,R] gathered from the parent's Z panel and the parent's own W,
 *   T' = -W L21,  G = I - L21' T',  [H; Z_RS] = [G; T'] L11^-1,  Z_SS = H' L11^-1
 * (the formulation the device plan uses, validated in tests/replay.py). Same panel layout as the factor; the diagonal
 * block of a Z panel is stored as a full square. W matrices are heap blocks owned by their supernode and freed when the
 * last child has gathered from them. Only ever timed / cross-checked; never linked into libgmrf_b200.so.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef void (*dgemm_t)(char *, char *, int *, int *, int *, double *, double *, int *, double *, int *, double *, double *, int *);

typedef struct {
    const i64 *super_ptr, *super_parent, *row_ptr, *rel_idx, *panel_off, *panel_ld, *child_ptr;
    const double *Lx;
    double *Zx;
    double **W;        /* [nsuper] heap block nr x nr (ld = nr) or NULL */
    int *pending;      /* [nsuper] children that still have to gather from W[s] */
    dgemm_t gemm;
    dtrsm_t trsm;
    int inner_parallel;   /* fronts are processed one after another: thread the gather of a big W as well */
} selinv_ctx;

static int selinv_one(const selinv_ctx *c, i64 s)
{
    const i64 ns = c->super_ptr[s + 1] - c->super_ptr[s];
    const i64 nrow = c->row_ptr[s + 1] - c->row_ptr[s];
    const i64 nr = nrow - ns, ld = c->panel_ld[s];
    const double *P = c->Lx + c->panel_off[s];
    double *Z = c->Zx + c->panel_off[s];
    double one = 1.0, zero = 0.0, mone = -1.0;
    int ns_ = (int)ns, nr_ = (int)nr, ld_ = (int)ld, nrow_ = (int)nrow;
    double *Ws = NULL;
    /* X = [G; T'] lives in the Z panel itself (nrow x ns, ld) */
    if (nr > 0) {
        const i64 p = c->super_parent[s];
        const i64 pns = c->super_ptr[p + 1] - c->super_ptr[p], pld = c->panel_ld[p];
        const i64 pnr = (c->row_ptr[p + 1] - c->row_ptr[p]) - pns;
        const double *Zp = c->Zx + c->panel_off[p];
        const double *Wp = c->W[p];
        const i64 *rel = c->rel_idx + c->row_ptr[s] + ns;
        Ws = (double *)malloc(sizeof(double) * (size_t)nr * (size_t)nr);
        if (!Ws) return -1;
#pragma omp parallel for schedule(dynamic, 16) if (c->inner_parallel && nr >= 512)
        for (i64 b = 0; b < nr; b++) {
            const i64 pb = rel[b];
            for (i64 a = b; a < nr; a++) {
                const i64 pa = rel[a];     /* pa >= pb: row lists are sorted */
                const double v = pb < pns ? Zp[pa + pb * pld] : Wp[(pa - pns) + (pb - pns) * pnr];
                Ws[a + b * nr] = v;
                Ws[b + a * nr] = v;
            }
        }
        {   /* this child is done reading the parent's W: the last one to finish frees it */
            int left;
#pragma omp atomic capture
            left = --c->pending[p];
            if (left == 0 && c->W[p]) { free(c->W[p]); c->W[p] = NULL; }
        }
        /* T' = -W L21 -> rows ns.. of the Z panel */
        c->gemm("N", "N", &nr_, &ns_, &nr_, &mone, Ws, &nr_, (double *)P + ns, &ld_, &zero, Z + ns, &ld_);
        /* G = I - L21' T' -> rows 0..ns of the Z panel */
        c->gemm("T", "N", &ns_, &ns_, &nr_, &mone, (double *)P + ns, &ld_, Z + ns, &ld_, &zero, Z, &ld_);
        for (i64 j = 0; j < ns; j++) Z[j + j * ld] += 1.0;
    } else {
        for (i64 j = 0; j < ns; j++) {
            memset(Z + j * ld, 0, sizeof(double) * (size_t)ns);
            Z[j + j * ld] = 1.0;
        }
    }
    /* [H; Z_RS] = [G; T'] L11^-1 */
    c->trsm("R", "L", "N", "N", &nrow_, &ns_, &one, (double *)P, &ld_, Z, &ld_);
    /* Z_SS = H' L11^-1: transpose H in place (ns x ns block), then the same solve on it */
    for (i64 j = 0; j < ns; j++)
        for (i64 i = j + 1; i < ns; i++) {
            const double t = Z[i + j * ld];
            Z[i + j * ld] = Z[j + i * ld];
            Z[j + i * ld] = t;
        }
    c->trsm("R", "L", "N", "N", &ns_, &ns_, &one, (double *)P, &ld_, Z, &ld_);
    c->W[s] = Ws;
    if (c->pending[s] == 0 && Ws) { free(Ws); c->W[s] = NULL; }     /* a leaf of the assembly tree: nobody gathers from it */
    return 0;
}

/* Selected inversion of the supernodes supers[0..count) of one level (call the levels top-down). parallel as above.
 * W / pending are caller-allocated arrays of nsuper entries (W zero-initialised, pending[s] = number of children of s).
 * Returns 0, or -1 if a W block could not be allocated. */
int cpu_supernodal_selinv_level(i64 count, const i64 *supers, const i64 *super_ptr, const i64 *super_parent, const i64 *row_ptr,
                                const i64 *rel_idx, const i64 *panel_off, const i64 *panel_ld, const i64 *child_ptr,
                                const double *Lx, double *Zx, void **W, int *pending, void *gemm, void *trsm, int parallel)
{
    selinv_ctx c = {super_ptr, super_parent, row_ptr, rel_idx, panel_off, panel_ld, child_ptr, Lx, Zx, (double **)W, pending,
                    (dgemm_t)gemm, (dtrsm_t)trsm, !parallel};
    int status = 0;
#pragma omp parallel for schedule(dynamic, 1) if (parallel)
    for (i64 t = 0; t < count; t++) {
        const i64 s = supers[t];
        if (selinv_one(&c, s)) {
#pragma omp atomic write
            status = -1;
        }
    }
    return status;
}

/* free whatever W blocks are still alive (error paths) */
void cpu_supernodal_selinv_cleanup(i64 nsuper, void **W)
{
    for (i64 s = 0; s < nsuper; s++)
        if (W[s]) { free(W[s]); W[s] = NULL; }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Supernodal triangular solves on the host cores (the CPU stand-in for `F \ rhs` / `F.UP \ x`,
 * /root/reference/src/workspace/backend.jl:191-193, :281-284), one right-hand side, level by level:
 *   forward  (levels ascending):  the parent pulls its children's update vectors, x_S = L11^-1 y_S, u = L21 x_S (+ pulled);
 *   backward (levels descending): x_S = L11^-T (y_S - L21' x_R), x_R read from the ancestors' finished entries.
 * `y` is the right-hand side in elimination order (overwritten by the solution), `u` a work array indexed like row_idx
 * (u_s lives at u[row_ptr[s] + ns .. row_ptr[s+1])). No two fronts of a level write the same entry.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef void (*dtrsv_t)(char *, char *, char *, int *, double *, int *, double *, int *);
typedef void (*dgemv_t)(char *, int *, int *, double *, double *, int *, double *, int *, double *, double *, int *);

int cpu_supernodal_solve_level(i64 count, const i64 *supers, const i64 *super_ptr, const i64 *row_ptr, const i64 *row_idx,
                               const i64 *rel_idx, const i64 *panel_off, const i64 *panel_ld, const i64 *child_ptr,
                               const i64 *child_idx, const double *Lx, double *y, double *u, void *trsv_, void *gemv_,
                               int backward, int parallel)
{
    dtrsv_t trsv = (dtrsv_t)trsv_;
    dgemv_t gemv = (dgemv_t)gemv_;
#pragma omp parallel for schedule(dynamic, 4) if (parallel)
    for (i64 t = 0; t < count; t++) {
        const i64 s = supers[t];
        const i64 f = super_ptr[s], ns = super_ptr[s + 1] - f;
        const i64 nrow = row_ptr[s + 1] - row_ptr[s], nr = nrow - ns, ld = panel_ld[s];
        double *P = (double *)Lx + panel_off[s];
        double *us = u + row_ptr[s] + ns;
        int ns_ = (int)ns, nr_ = (int)nr, ld_ = (int)ld, inc = 1;
        double one = 1.0, zero = 0.0, mone = -1.0;
        if (!backward) {
            for (i64 i = 0; i < nr; i++) us[i] = 0.0;
            for (i64 ci = child_ptr[s]; ci < child_ptr[s + 1]; ci++) {
                const i64 c = child_idx[ci];
                const i64 cns = super_ptr[c + 1] - super_ptr[c], cnr = (row_ptr[c + 1] - row_ptr[c]) - cns;
                const i64 *rel = rel_idx + row_ptr[c] + cns;
                const double *uc = u + row_ptr[c] + cns;
                for (i64 k = 0; k < cnr; k++) {
                    const i64 pos = rel[k];
                    if (pos < ns) y[f + pos] -= uc[k];
                    else us[pos - ns] += uc[k];
                }
            }
            trsv("L", "N", "N", &ns_, P, &ld_, y + f, &inc);
            if (nr > 0) {
                if (parallel || nr < 2048) gemv("N", &nr_, &ns_, &one, P + ns, &ld_, y + f, &inc, &one, us, &inc);
                else {
                    /* one big front at a time (top of the tree): its rows are split over the OpenMP threads, so that the whole
                     * sweep runs on ONE thread pool (BLAS single-threaded throughout; alternating pools costs more than the solve) */
                    const i64 CH = 1024, nch = (nr + CH - 1) / CH;
#pragma omp parallel for schedule(static)
                    for (i64 b = 0; b < nch; b++) {
                        int m_ = (int)((b + 1) * CH <= nr ? CH : nr - b * CH), one_i = 1;
                        double o = 1.0;
                        gemv("N", &m_, &ns_, &o, P + ns + b * CH, &ld_, y + f, &one_i, &o, us + b * CH, &one_i);
                    }
                }
            }
        } else {
            if (nr > 0) {
                const i64 *rows = row_idx + row_ptr[s] + ns;
                for (i64 i = 0; i < nr; i++) us[i] = y[rows[i]];
                if (parallel || ns < 256 || nr < 2048) gemv("T", &nr_, &ns_, &mone, P + ns, &ld_, us, &inc, &one, y + f, &inc);
                else {
                    const i64 CH = 64, nch = (ns + CH - 1) / CH;                /* columns of L21 split over the threads */
#pragma omp parallel for schedule(static)
                    for (i64 b = 0; b < nch; b++) {
                        int n_ = (int)((b + 1) * CH <= ns ? CH : ns - b * CH), one_i = 1;
                        double o = 1.0, mo = -1.0;
                        gemv("T", &nr_, &n_, &mo, P + ns + b * CH * ld, &ld_, us, &one_i, &o, y + f + b * CH, &one_i);
                    }
                }
            }
            trsv("L", "T", "N", &ns_, P, &ld_, y + f, &inc);
        }
        (void)zero;
    }
    return 0;
}
