/*
 * gmrf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the arithmetic the reference delegates to third-party packages at
 * its backend boundary (/root/reference/src/workspace/backend.jl):
 *   cholesky(Q; perm)              backend.jl:148-149  -> oracle_symbolic + oracle_factor
 *   cholesky!(F, S; check=false)   backend.jl:184      -> oracle_factor (numeric only)
 *   F \ rhs                        backend.jl:192,208  -> oracle_solve
 *   logdet(F)                      backend.jl:212      -> oracle_logdet
 *   SelectedInversion.selinv(F).Z  backend.jl:232,253  -> oracle_selinv (Takahashi recursion)
 *   F.UP \ x                       backend.jl:283      -> oracle_ltsolve (P' L^-T x)
 *
 * The libraries that own this arithmetic are absent from /root/reference and from this image:
 * SuiteSparse CHOLMOD (via the SparseArrays stdlib, compat "<0.0.1, 1"), SelectedInversion.jl
 * (compat "0.2.1"), CliqueTrees.jl ("1.19.1") -- Project.toml:60-100, no Manifest. What is
 * restated here is their published algorithm in its simplest (simplicial, scalar) form:
 *   - elimination tree (Liu 1990) and row-subtree reach (Davis, "Direct Methods for Sparse
 *     Linear Systems", ch. 4) for the exact column counts of L = chol(P Q P'),
 *   - up-looking sparse Cholesky (ibid. ch. 4.7),
 *   - Takahashi/Erisman-Tinney selected inversion on the pattern of L.
 * PARITY PINNING: CHOLMOD itself cannot be run here (no Julia, no libcholmod). The oracle is
 * pinned the way the reference's own tests pin their backends -- against dense LinearAlgebra
 * identities (inv, logdet, \) on the deterministic fixtures of SURVEY.md section 8c
 * (tests/test_oracle.py). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call into this file.
 *
 * Conventions: 0-based int64 indices; A is the FULL symmetric CSC, only entries with
 * row <= col are read (Symmetric(Q) = upper triangle, src/workspace/gmrf_workspace.jl:176);
 * perm[k] = original index of the k-th pivot; L is lower-triangular CSC in the permuted order
 * with sorted rows and the diagonal first in every column.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* Upper triangle of C = P A P' in CSC (column k holds rows i <= k). Caller frees. */
static int build_permuted_upper(i64 n, const i64 *Ap, const i64 *Ai, const double *Ax,
                                const i64 *perm, i64 **Cp_out, i64 **Ci_out, double **Cx_out)
{
    i64 *iperm = (i64 *)malloc(sizeof(i64) * (size_t)n);
    i64 *Cp = (i64 *)calloc((size_t)n + 1, sizeof(i64));
    if (!iperm || !Cp) return -1;
    for (i64 k = 0; k < n; k++) iperm[perm ? perm[k] : k] = k;
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            i64 i = Ai[p];
            if (i > j) continue;
            i64 a = iperm[i], b = iperm[j];
            Cp[(a > b ? a : b) + 1]++;
        }
    for (i64 j = 0; j < n; j++) Cp[j + 1] += Cp[j];
    i64 nz = Cp[n];
    i64 *Ci = (i64 *)malloc(sizeof(i64) * (size_t)(nz > 0 ? nz : 1));
    double *Cx = Ax ? (double *)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1)) : NULL;
    i64 *w = (i64 *)malloc(sizeof(i64) * (size_t)n);
    memcpy(w, Cp, sizeof(i64) * (size_t)n);
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            i64 i = Ai[p];
            if (i > j) continue;
            i64 a = iperm[i], b = iperm[j];
            i64 lo = a < b ? a : b, hi = a < b ? b : a;
            i64 q = w[hi]++;
            Ci[q] = lo;
            if (Cx) Cx[q] = Ax[p];
        }
    free(w);
    free(iperm);
    *Cp_out = Cp; *Ci_out = Ci; *Cx_out = Cx;
    return 0;
}

/* Liu's elimination tree of a matrix given by its upper triangle. */
static void etree_upper(i64 n, const i64 *Cp, const i64 *Ci, i64 *parent)
{
    i64 *anc = (i64 *)malloc(sizeof(i64) * (size_t)n);
    for (i64 k = 0; k < n; k++) {
        parent[k] = -1;
        anc[k] = -1;
        for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
            i64 i = Ci[p];
            while (i != -1 && i < k) {
                i64 next = anc[i];
                anc[i] = k;
                if (next == -1) parent[i] = k;
                i = next;
            }
        }
    }
    free(anc);
}

/* Pattern of row k of L (excluding the diagonal), topologically ordered in s[top..n-1]. */
static i64 row_reach(i64 n, const i64 *Cp, const i64 *Ci, i64 k, const i64 *parent,
                     i64 *s, i64 *mark)
{
    i64 top = n;
    mark[k] = k;
    for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
        i64 i = Ci[p];
        if (i >= k) continue;
        i64 len = 0;
        for (; mark[i] != k; i = parent[i]) {
            s[len++] = i;
            mark[i] = k;
        }
        while (len > 0) s[--top] = s[--len];
    }
    return top;
}

/* Exact column counts (incl. diagonal) and etree of L = chol(P A P'). Returns nnz(L). */
i64 oracle_symbolic(i64 n, const i64 *Ap, const i64 *Ai, const i64 *perm,
                    i64 *parent, i64 *colcount)
{
    i64 *Cp, *Ci; double *Cx;
    if (build_permuted_upper(n, Ap, Ai, NULL, perm, &Cp, &Ci, &Cx)) return -1;
    etree_upper(n, Cp, Ci, parent);
    i64 *s = (i64 *)malloc(sizeof(i64) * (size_t)n);
    i64 *mark = (i64 *)malloc(sizeof(i64) * (size_t)n);
    for (i64 k = 0; k < n; k++) { mark[k] = -1; colcount[k] = 1; }
    i64 nnz = n;
    for (i64 k = 0; k < n; k++) {
        i64 top = row_reach(n, Cp, Ci, k, parent, s, mark);
        for (i64 t = top; t < n; t++) colcount[s[t]]++;
        nnz += n - top;
    }
    free(s); free(mark); free(Cp); free(Ci);
    return nnz;
}

/* Up-looking numeric Cholesky. Lp must hold the cumulative column counts from
 * oracle_symbolic. Returns 0, or k+1 for the first non-positive pivot k (permuted order). */
int oracle_factor(i64 n, const i64 *Ap, const i64 *Ai, const double *Ax, const i64 *perm,
                  const i64 *Lp, i64 *Li, double *Lx)
{
    i64 *Cp, *Ci; double *Cx;
    if (build_permuted_upper(n, Ap, Ai, Ax, perm, &Cp, &Ci, &Cx)) return -1;
    i64 *parent = (i64 *)malloc(sizeof(i64) * (size_t)n);
    etree_upper(n, Cp, Ci, parent);
    i64 *s = (i64 *)malloc(sizeof(i64) * (size_t)n);
    i64 *mark = (i64 *)malloc(sizeof(i64) * (size_t)n);
    i64 *c = (i64 *)malloc(sizeof(i64) * (size_t)n);
    double *x = (double *)calloc((size_t)n, sizeof(double));
    int status = 0;
    for (i64 k = 0; k < n; k++) { mark[k] = -1; c[k] = Lp[k]; }
    for (i64 k = 0; k < n; k++) {
        i64 top = row_reach(n, Cp, Ci, k, parent, s, mark);
        double d = 0.0;
        for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
            if (Ci[p] < k) x[Ci[p]] += Cx[p];
            else if (Ci[p] == k) d += Cx[p];
        }
        for (i64 t = top; t < n; t++) {
            i64 i = s[t];
            double lki = x[i] / Lx[Lp[i]];
            x[i] = 0.0;
            for (i64 p = Lp[i] + 1; p < c[i]; p++) x[Li[p]] -= Lx[p] * lki;
            d -= lki * lki;
            i64 q = c[i]++;
            Li[q] = k;
            Lx[q] = lki;
        }
        if (!(d > 0.0)) { status = (int)(k + 1); break; }
        i64 q = c[k]++;
        Li[q] = k;
        Lx[q] = sqrt(d);
    }
    free(x); free(c); free(mark); free(s); free(parent); free(Cp); free(Ci); free(Cx);
    return status;
}

double oracle_logdet(i64 n, const i64 *Lp, const double *Lx)
{
    double acc = 0.0;
    for (i64 j = 0; j < n; j++) acc += log(Lx[Lp[j]]);
    return 2.0 * acc;
}

/* y := L^-1 y (in place, permuted order) */
void oracle_lsolve(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, double *y)
{
    for (i64 j = 0; j < n; j++) {
        y[j] /= Lx[Lp[j]];
        double yj = y[j];
        for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) y[Li[p]] -= Lx[p] * yj;
    }
}

/* y := L^-T y (in place, permuted order) */
void oracle_ltsolve_inplace(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, double *y)
{
    for (i64 j = n - 1; j >= 0; j--) {
        double acc = y[j];
        for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) acc -= Lx[p] * y[Li[p]];
        y[j] = acc / Lx[Lp[j]];
    }
}

/* x = Q^-1 b = P' L^-T L^-1 P b, nrhs columns, column-major with leading dimension n. */
void oracle_solve(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, const i64 *perm,
                  const double *b, double *x, i64 nrhs)
{
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    for (i64 r = 0; r < nrhs; r++) {
        for (i64 k = 0; k < n; k++) y[k] = b[r * n + (perm ? perm[k] : k)];
        oracle_lsolve(n, Lp, Li, Lx, y);
        oracle_ltsolve_inplace(n, Lp, Li, Lx, y);
        for (i64 k = 0; k < n; k++) x[r * n + (perm ? perm[k] : k)] = y[k];
    }
    free(y);
}

/* x = P' L^-T z  (the `factor.UP \ z` half solve used for sampling; z is NOT permuted,
 * exactly like CHOLMOD's UP component: UP = L' P, so UP \ z = P' (L' \ z)). */
void oracle_ltsolve(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, const i64 *perm,
                    const double *z, double *x, i64 nrhs)
{
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    for (i64 r = 0; r < nrhs; r++) {
        memcpy(y, z + r * n, sizeof(double) * (size_t)n);
        oracle_ltsolve_inplace(n, Lp, Li, Lx, y);
        for (i64 k = 0; k < n; k++) x[r * n + (perm ? perm[k] : k)] = y[k];
    }
    free(y);
}

/* Takahashi recursion: Zx[p] = (Q_perm^-1)[Li[p], j] for every stored entry p of column j of L.
 *   Z_ij = -(1/L_jj) sum_{k in struct(j), k>j} L_kj Z_{max(i,k),min(i,k)}      (i > j)
 *   Z_jj =  1/L_jj^2 - (1/L_jj) sum_{k>j} L_kj Z_kj
 * Columns are processed from n-1 down to 0; every Z entry a column needs lives in a column
 * to its right because struct(j)\{j} is a clique of the filled graph. */
void oracle_selinv(i64 n, const i64 *Lp, const i64 *Li, const double *Lx, double *Zx)
{
    i64 *where = (i64 *)malloc(sizeof(i64) * (size_t)n); /* row -> position in column k of Z */
    for (i64 i = 0; i < n; i++) where[i] = -1;
    for (i64 j = n - 1; j >= 0; j--) {
        i64 p0 = Lp[j], p1 = Lp[j + 1];
        double ljj = Lx[p0];
        for (i64 p = p0 + 1; p < p1; p++) Zx[p] = 0.0;
        /* accumulate sum_k L_kj Z_{ik} for all i in struct(j): loop k, scatter column k of Z */
        for (i64 pk = p0 + 1; pk < p1; pk++) {
            i64 k = Li[pk];
            double lkj = Lx[pk];
            for (i64 q = Lp[k]; q < Lp[k + 1]; q++) where[Li[q]] = q;
            /* i >= k: Z_ik sits in column k; contributes to row i (and symmetric to row k) */
            for (i64 pi = pk; pi < p1; pi++) {
                i64 q = where[Li[pi]];
                double zik = Zx[q];
                Zx[pi] += lkj * zik;
                if (pi != pk) Zx[pk] += Lx[pi] * zik;
            }
            for (i64 q = Lp[k]; q < Lp[k + 1]; q++) where[Li[q]] = -1;
        }
        double dsum = 0.0;
        for (i64 p = p0 + 1; p < p1; p++) {
            Zx[p] = -Zx[p] / ljj;
            dsum += Lx[p] * Zx[p];
        }
        Zx[p0] = 1.0 / (ljj * ljj) - dsum / ljj;
    }
    free(where);
}

/* y = A x for the full symmetric CSC A (used for residual checks in tests). */
void oracle_spmv(i64 n, const i64 *Ap, const i64 *Ai, const double *Ax, const double *x, double *y)
{
    for (i64 i = 0; i < n; i++) y[i] = 0.0;
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) y[Ai[p]] += Ax[p] * x[j];
}
