/* julia_ccall_replay.c -- plain-C replay of the ccall sequence of julia/B200Backend.jl (and the entry points
 * julia/B200LinearSolve.jl reuses), argument for argument, compiled against include/gmrf_b200.h only:
 *   create (1-based SparseMatrixCSC arrays, 1-based perm, index_base = 1)   B200Backend.jl:64-66
 *   refactorize / logdet / solve (vector, matrix) / solve_Lt                  :77, :115, :86-94, :100-109
 *   selinv_diag, selinv_nnz + selinv_pattern(…, 1) + selinv_values            :127, :138-146
 *   selinv_extract(…, 1), selinv_dot(…, 1)                                     :155, :167
 *   factor_nnz + factor_pattern(…, 1) + factor_values                         :177-184
 *   analysis_export (size query, then the blob) + create_from_analysis(…, 1)  :192-194, :59-61
 *   get_perm(…, 1), destroy                                                    :70
 *   selinv_quadform_rows(…, 1), set_base_values / refactorize_base_minus_diag, set_hessian_pattern(…, 1) /
 *   refactorize_base_minus_sparse                                            (the device-side consumers of the glue)
 * The matrix is the reference's first deterministic fixture (test/workspace/test_backend_ordering.jl:9-17: 12 x 12 grid
 * Laplacian + 0.1 I with a dense border row/column), answers are checked against a dense Cholesky written here.
 * Exit code 0 = every check passed.  cc julia_ccall_replay.c -I include -L lib -lgmrf_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gmrf_b200.h"

#define CHECK(rc, what) do { int r_ = (rc); if (r_ != 0) { fprintf(stderr, "%s failed: %d (%s)\n", what, r_, gmrf_b200_last_error(h)); return 1; } } while (0)
#define REQUIRE(cond, what) do { if (!(cond)) { fprintf(stderr, "CHECK FAILED: %s\n", what); return 1; } } while (0)

static double *dense_fixture(int *n_out) {
    const int g = 12, N = g * g + 1;
    double *A = calloc((size_t)N * N, sizeof(double));
    for (int y = 0; y < g; y++)
        for (int x = 0; x < g; x++) {
            int i = y * g + x;
            A[i + (size_t)i * N] = 4.0 + 0.1;              /* kron(I, A1) + kron(A1, I) + 0.1 I with A1 = tridiag(-1, 2, -1) */
            if (x + 1 < g) { A[i + (size_t)(i + 1) * N] = -1.0; A[(i + 1) + (size_t)i * N] = -1.0; }
            if (y + 1 < g) { A[i + (size_t)(i + g) * N] = -1.0; A[(i + g) + (size_t)i * N] = -1.0; }
        }
    for (int i = 0; i < N - 1; i++) { A[i + (size_t)(N - 1) * N] = 0.01; A[(N - 1) + (size_t)i * N] = 0.01; }
    A[(N - 1) + (size_t)(N - 1) * N] = 2.0;
    *n_out = N;
    return A;
}

int main(void) {
    gmrf_b200_handle *h = NULL;
    int n;
    double *A = dense_fixture(&n);
    /* full symmetric CSC, 1-based, sorted rows: what SparseMatrixCSC{Float64,Int} holds */
    int64_t *colptr = malloc(sizeof(int64_t) * (n + 1)), nnz = 0;
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) if (A[i + (size_t)j * n] != 0.0) nnz++;
    int64_t *rowval = malloc(sizeof(int64_t) * nnz);
    double *nzval = malloc(sizeof(double) * nnz);
    nnz = 0;
    for (int j = 0; j < n; j++) {
        colptr[j] = nnz + 1;
        for (int i = 0; i < n; i++) if (A[i + (size_t)j * n] != 0.0) { rowval[nnz] = i + 1; nzval[nnz++] = A[i + (size_t)j * n]; }
    }
    colptr[n] = nnz + 1;
    /* dense reference: Cholesky, logdet, inverse */
    double *Ld = malloc(sizeof(double) * n * n), *inv = calloc((size_t)n * n, sizeof(double));
    memcpy(Ld, A, sizeof(double) * n * n);
    double logdet_ref = 0.0;
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) for (int i = j; i < n; i++) Ld[i + (size_t)j * n] -= Ld[i + (size_t)k * n] * Ld[j + (size_t)k * n];
        double d = sqrt(Ld[j + (size_t)j * n]);
        for (int i = j; i < n; i++) Ld[i + (size_t)j * n] /= d;
        logdet_ref += 2.0 * log(d);
    }
    for (int c = 0; c < n; c++) {                       /* inv(:, c) = A \ e_c */
        double *x = inv + (size_t)c * n;
        x[c] = 1.0;
        for (int i = 0; i < n; i++) { for (int k = 0; k < i; k++) x[i] -= Ld[i + (size_t)k * n] * x[k]; x[i] /= Ld[i + (size_t)i * n]; }
        for (int i = n - 1; i >= 0; i--) { for (int k = i + 1; k < n; k++) x[i] -= Ld[k + (size_t)i * n] * x[k]; x[i] /= Ld[i + (size_t)i * n]; }
    }
    /* a caller-supplied 1-based permutation (reverse order), as `ordering = perm` passes it */
    int64_t *perm = malloc(sizeof(int64_t) * n);
    for (int k = 0; k < n; k++) perm[k] = n - k;

    for (int pass = 0; pass < 2; pass++) {              /* pass 0: library ordering (perm = NULL), pass 1: user perm */
        CHECK(gmrf_b200_create(&h, n, colptr, rowval, 1, pass ? perm : NULL, 1, 0), "create");
        CHECK(gmrf_b200_refactorize(h, nzval, nnz), "refactorize");
        double ld = 0.0;
        CHECK(gmrf_b200_logdet(h, &ld), "logdet");
        REQUIRE(fabs(ld - logdet_ref) <= 1e-10 * fabs(logdet_ref), "logdet vs dense Cholesky");
        int64_t *p = malloc(sizeof(int64_t) * n);
        CHECK(gmrf_b200_get_perm(h, p, 1), "get_perm");
        char *seen = calloc(n + 1, 1);
        for (int k = 0; k < n; k++) { REQUIRE(p[k] >= 1 && p[k] <= n && !seen[p[k]], "get_perm returns a 1-based permutation"); seen[p[k]] = 1; }
        /* solve: vector and 3-column matrix */
        double *B = malloc(sizeof(double) * n * 3), *X = malloc(sizeof(double) * n * 3);
        for (int i = 0; i < 3 * n; i++) B[i] = sin(0.37 * i) + 0.1;
        CHECK(gmrf_b200_solve(h, B, X, n, 1), "solve vector");
        CHECK(gmrf_b200_solve(h, B, X, n, 3), "solve matrix");
        double err = 0.0, nrm = 0.0;
        for (int c = 0; c < 3; c++) for (int i = 0; i < n; i++) {
            double s = 0.0; for (int k = 0; k < n; k++) s += inv[i + (size_t)k * n] * B[k + (size_t)c * n];
            err += (s - X[i + (size_t)c * n]) * (s - X[i + (size_t)c * n]); nrm += s * s;
        }
        REQUIRE(sqrt(err) <= 1e-10 * sqrt(nrm), "solve vs dense inverse");
        /* half solve: x = P' L^-T z  =>  x' A x = z' z */
        CHECK(gmrf_b200_solve_Lt(h, B, X, n, 1), "solve_Lt");
        double q = 0.0, zz = 0.0;
        for (int i = 0; i < n; i++) { double s = 0.0; for (int k = 0; k < n; k++) s += A[i + (size_t)k * n] * X[k]; q += X[i] * s; zz += B[i] * B[i]; }
        REQUIRE(fabs(q - zz) <= 1e-10 * zz, "half solve quadratic form");
        /* selected inverse: diagonal, CSC (1-based pattern), extract and dot at Q's own pattern */
        double *d = malloc(sizeof(double) * n);
        CHECK(gmrf_b200_selinv_diag(h, d), "selinv_diag");
        for (int i = 0; i < n; i++) REQUIRE(fabs(d[i] - inv[i + (size_t)i * n]) <= 1e-8 * inv[i + (size_t)i * n], "selinv_diag vs dense inverse");
        int64_t znz = 0;
        CHECK(gmrf_b200_selinv_nnz(h, &znz), "selinv_nnz");
        int64_t *zcp = malloc(sizeof(int64_t) * (n + 1)), *zrv = malloc(sizeof(int64_t) * znz);
        double *zv = malloc(sizeof(double) * znz);
        CHECK(gmrf_b200_selinv_pattern(h, zcp, zrv, 1), "selinv_pattern");
        CHECK(gmrf_b200_selinv_values(h, zv), "selinv_values");
        REQUIRE(zcp[0] == 1 && zcp[n] == znz + 1, "selinv colptr is 1-based");
        for (int j = 0; j < n; j++) for (int64_t k = zcp[j] - 1; k < zcp[j + 1] - 1; k++) {
            REQUIRE(zrv[k] >= 1 && zrv[k] <= n, "selinv rowval is 1-based");
            double ref = inv[(zrv[k] - 1) + (size_t)j * n];
            REQUIRE(fabs(zv[k] - ref) <= 1e-6 * fabs(ref) + 1e-14, "selinv entries vs dense inverse");
            if (zrv[k] - 1 == j) REQUIRE(zv[k] == d[j], "diag(selinv) is bit-identical to selinv_diag");
        }
        double *ex = malloc(sizeof(double) * nnz), tr = 0.0, tr_ref = 0.0;
        CHECK(gmrf_b200_selinv_extract(h, n, colptr, rowval, 1, ex), "selinv_extract");
        CHECK(gmrf_b200_selinv_dot(h, n, colptr, rowval, 1, nzval, &tr), "selinv_dot");
        for (int j = 0; j < n; j++) for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; k++) {
            double ref = inv[(rowval[k] - 1) + (size_t)j * n];
            REQUIRE(fabs(ex[k] - ref) <= 1e-6 * fabs(ref) + 1e-14, "selinv_extract vs dense inverse");
            tr_ref += ref * nzval[k];
        }
        REQUIRE(fabs(tr - n) <= 1e-8 * n && fabs(tr - tr_ref) <= 1e-8 * n, "tr(Q^-1 Q) = n");
        /* factor export P'L (1-based pattern): R R' = Q */
        int64_t lnz = 0;
        CHECK(gmrf_b200_factor_nnz(h, &lnz), "factor_nnz");
        int64_t *lcp = malloc(sizeof(int64_t) * (n + 1)), *lrv = malloc(sizeof(int64_t) * lnz);
        double *lv = malloc(sizeof(double) * lnz), *RRt = calloc((size_t)n * n, sizeof(double));
        CHECK(gmrf_b200_factor_pattern(h, lcp, lrv, 1), "factor_pattern");
        CHECK(gmrf_b200_factor_values(h, lv), "factor_values");
        for (int k = 0; k < n; k++) for (int64_t a = lcp[k] - 1; a < lcp[k + 1] - 1; a++) for (int64_t b2 = lcp[k] - 1; b2 < lcp[k + 1] - 1; b2++)
            RRt[(lrv[a] - 1) + (size_t)(lrv[b2] - 1) * n] += lv[a] * lv[b2];
        for (int i = 0; i < n * n; i++) REQUIRE(fabs(RRt[i] - A[i]) <= 1e-10, "R R' = Q for the exported square root");
        /* analysis blob -> second handle through create_from_analysis with the same 1-based arrays: identical bits */
        int64_t nb = 0;
        CHECK(gmrf_b200_analysis_export(h, NULL, 0, &nb), "analysis_export (size)");
        unsigned char *blob = malloc((size_t)nb);
        CHECK(gmrf_b200_analysis_export(h, blob, nb, &nb), "analysis_export");
        gmrf_b200_handle *h2 = NULL;
        { gmrf_b200_handle *keep = h; h = NULL; int rc = gmrf_b200_create_from_analysis(&h2, n, colptr, rowval, 1, blob, nb, 0); h = keep; CHECK(rc, "create_from_analysis"); }
        if (gmrf_b200_refactorize(h2, nzval, nnz) != 0) { fprintf(stderr, "refactorize (restored handle) failed\n"); return 1; }
        double ld2 = 0.0; gmrf_b200_logdet(h2, &ld2);
        REQUIRE(ld2 == ld, "restored analysis reproduces the log-determinant bit for bit");
        /* diag(A Sigma A') of a sparse design matrix in CSR, 1-based (row_diag_AΣAt in the glue): point observations and
         * one two-point row on grid neighbours (pairs inside Q's pattern, hence inside the factor's) */
        {
            int64_t rp[5] = {1, 2, 3, 5, 5}, ci[4] = {1, 77, 10, 11};          /* 4 rows; the last one is empty */
            double av[4] = {1.0, 2.0, 0.5, -1.5}, qf[4];
            CHECK(gmrf_b200_selinv_quadform_rows(h, 4, rp, ci, av, 1, qf), "selinv_quadform_rows");
            double want2 = 0.25 * inv[9 + 9 * (size_t)n] + 2.25 * inv[10 + 10 * (size_t)n] + 2.0 * 0.5 * -1.5 * inv[9 + 10 * (size_t)n];
            REQUIRE(fabs(qf[0] - inv[0]) <= 1e-8 * inv[0], "quadform row 1 = Sigma_11");
            REQUIRE(fabs(qf[1] - 4.0 * inv[76 + 76 * (size_t)n]) <= 1e-8 * 4.0 * inv[76 + 76 * (size_t)n], "quadform row 2 = 4 Sigma_77,77");
            REQUIRE(fabs(qf[2] - want2) <= 1e-8 * fabs(want2) + 1e-14, "quadform row 3 (two neighbouring sites)");
            REQUIRE(qf[3] == 0.0, "quadform of an empty row");
        }
        /* Newton iterates formed in HBM: prior values resident, a diagonal and a sparse Hessian update (1-based positions) */
        {
            CHECK(gmrf_b200_set_base_values(h, nzval, nnz), "set_base_values");
            double *hd = malloc(sizeof(double) * n), *nz2 = malloc(sizeof(double) * nnz);
            memcpy(nz2, nzval, sizeof(double) * nnz);
            int64_t *hp = malloc(sizeof(int64_t) * n);
            for (int j = 0; j < n; j++) {
                hd[j] = -0.25 - 0.01 * j;
                for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; k++) if (rowval[k] - 1 == j) { nz2[k] -= hd[j]; hp[j] = k + 1; }
            }
            double ld_host = 0.0, ld_dev = 0.0, ld_sp = 0.0;
            CHECK(gmrf_b200_refactorize(h, nz2, nnz), "refactorize (host-assembled iterate)");
            gmrf_b200_logdet(h, &ld_host);
            CHECK(gmrf_b200_refactorize_base_minus_diag(h, hd, n), "refactorize_base_minus_diag");
            gmrf_b200_logdet(h, &ld_dev);
            CHECK(gmrf_b200_set_hessian_pattern(h, hp, n, 1), "set_hessian_pattern");
            CHECK(gmrf_b200_refactorize_base_minus_sparse(h, hd, n), "refactorize_base_minus_sparse");
            gmrf_b200_logdet(h, &ld_sp);
            REQUIRE(ld_dev == ld_host && ld_sp == ld_host, "device-formed iterates are bit-identical to the host-assembled one");
            CHECK(gmrf_b200_refactorize(h, nzval, nnz), "refactorize (back to Q)");
        }
        /* error behaviour the glue maps to ArgumentError: nnz mismatch */
        REQUIRE(gmrf_b200_refactorize(h, nzval, nnz - 1) == GMRF_B200_ERR_ARG, "nnz mismatch is GMRF_B200_ERR_ARG");
        gmrf_b200_destroy(h2);
        gmrf_b200_destroy(h);
        h = NULL;
        printf("pass %d ok (logdet %.12f, nnz(Z) %lld, nnz(L) %lld)\n", pass, ld, (long long)znz, (long long)lnz);
    }
    printf("julia_ccall_replay: all checks passed\n");
    return 0;
}
