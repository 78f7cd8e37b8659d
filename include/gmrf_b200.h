/*
 * gmrf_b200.h -- C-ABI of libgmrf_b200.so: a B200 (sm_100a) sparse-Cholesky backend that sits
 * behind GaussianMarkovRandomFields.jl's `WorkspaceBackend` protocol
 * (reference: src/workspace/backend.jl:8-30; second implementation to imitate:
 * src/workspace/cliquetrees_backend.jl:21-150).
 *
 * Conventions
 *   - plain pointers and sizes only; every output buffer is CALLER-allocated; the library never
 *     keeps a host pointer after a call returns and never frees caller memory;
 *   - matrices are column-major Float64; sparse patterns are CSC with Int64 indices, FULL symmetric
 *     pattern as stored in `GMRFWorkspace.Q` (src/workspace/gmrf_workspace.jl:32); only entries with
 *     row <= col are read (`Symmetric(ws.Q)`, gmrf_workspace.jl:176). `index_base` is 1 for Julia,
 *     0 for C/Python callers;
 *   - return value: 0 = ok, <0 = usage / CUDA error (text via gmrf_b200_last_error), >0 = matrix
 *     not positive definite, value = 1-based column (in the factor's elimination order) of the
 *     first non-positive pivot. The reference factorizes with `check=false`
 *     (backend.jl:184), so the Julia glue ignores >0 unless asked;
 *   - a handle is bound to one device and one CUDA stream and is NOT thread-safe; distinct handles
 *     may be used concurrently from different host threads (no global lock; cf. the CHOLMOD lock
 *     note in src/workspace/workspace_pool.jl:15-21);
 *   - results are run-to-run bit-reproducible (no floating-point atomics anywhere)
 *     (test/gaussian_approximation/test_predictive_convergence.jl:20-28).
 *   - there is NO CPU fallback: every numeric entry point fails with GMRF_B200_ERR_NO_DEVICE on a
 *     handle created with device < 0 (analysis-only handle).
 */
#ifndef GMRF_B200_H
#define GMRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmrf_b200_handle gmrf_b200_handle;

enum {
    GMRF_B200_OK = 0,
    GMRF_B200_ERR_ARG = -1,        /* bad argument (size / pattern / nnz mismatch -> ArgumentError) */
    GMRF_B200_ERR_CUDA = -2,       /* CUDA runtime error */
    GMRF_B200_ERR_NO_DEVICE = -3,  /* numeric call on an analysis-only handle */
    GMRF_B200_ERR_STATE = -4,      /* e.g. solve before the first refactorize */
    GMRF_B200_ERR_ALLOC = -5
};

/* Fill-reducing ordering used when `perm` is NULL. The reference resolves `ordering=` on the host
 * (ordering_permutation, backend.jl:73-133) and may hand over the permutation instead. */
enum {
    GMRF_B200_ORDER_NATURAL = 0,
    GMRF_B200_ORDER_ND = 1,        /* nested dissection (METIS_NodeND), default */
    GMRF_B200_ORDER_AMD = 2        /* approximate minimum degree (own implementation) */
};

/* ---- lifecycle ------------------------------------------------------------------------------
 * replaces  CHOLMODBackend(Q::Symmetric; ordering)   backend.jl:147-153  (symbolic part)
 *           CliqueTreesBackend(Q; alg)               cliquetrees_backend.jl:28-41
 * Symbolic analysis on the host (ordering, etree, exact column counts, supernodes, level
 * schedule, scatter maps), uploaded once to `device`. No numeric work happens here.
 * perm (optional): perm[k] = index (in `index_base`) of the k-th pivot. device < 0 -> analysis only. */
int gmrf_b200_create(gmrf_b200_handle **out, int64_t n, const int64_t *colptr, const int64_t *rowval,
                     int index_base, const int64_t *perm, int ordering, int device);
/* Persisting / sharing the symbolic analysis (the reference has no serialization: the symbolic half of its CHOLMOD
 * factor lives and dies with the backend, backend.jl:32-61; SURVEY.md section 5 names the analysis as the one state worth
 * keeping). `analysis_export` writes a self-describing byte stream (call with buf = NULL for the size);
 * `create_from_analysis` is `create` with the ordering / etree / supernodes / schedules read from such a stream instead
 * of recomputed -- for a second session on the same mesh, or for the other handles of a pool (one analysis, one handle
 * per GPU: workspace_pool.jl:55-58 "resolve the permutation ONCE"). The stream is tied to the pattern by a hash; a
 * different pattern, a truncated or foreign stream give -1 with a message. */
int gmrf_b200_create_from_analysis(gmrf_b200_handle **out, int64_t n, const int64_t *colptr, const int64_t *rowval,
                                   int index_base, const void *analysis, int64_t analysis_bytes, int device);
int gmrf_b200_analysis_export(const gmrf_b200_handle *h, void *buf, int64_t capacity, int64_t *bytes);
/* 1 if the two handles hold identical analyses (every table), 0 if not: test hook for the round trip above */
int gmrf_b200_analysis_equal(const gmrf_b200_handle *a, const gmrf_b200_handle *b);
void gmrf_b200_destroy(gmrf_b200_handle *h);
const char *gmrf_b200_last_error(const gmrf_b200_handle *h);   /* h may be NULL: last create error */

/* ---- numeric factorization -------------------------------------------------------------------
 * replaces  refactorize!(b, Q::Symmetric)   backend.jl:178-189 (+ _copy_sparse_values! :165-176)
 * nzval: the nnz values of the full symmetric CSC, positionally matching the pattern given to
 * create (nnz mismatch -> GMRF_B200_ERR_ARG, like the ArgumentError at backend.jl:168-173).
 * Invalidates every selected-inverse cache (backend.jl:185-187). The log-determinant is reduced
 * inside the same pass. `_device` takes nzval already resident in this device's HBM. */
int gmrf_b200_refactorize(gmrf_b200_handle *h, const double *nzval, int64_t nnz);
int gmrf_b200_refactorize_device(gmrf_b200_handle *h, const double *d_nzval, int64_t nnz);

/* Device-side value assembly for hyperparameter loops. replaces the per-theta host re-assembly + upload of nzval:
 * `(model)(ws; theta...)` -> _pad_to_workspace_pattern (src/workspace/latent_model_integration.jl:151-250) followed by
 * _copy_sparse_values! (backend.jl:165-176). set_value_basis uploads nbasis (<= 8) value arrays laid out on the pattern
 * given to create (row-major nbasis x nnz) once; refactorize_combination forms nzval = sum_j coeff[j] * basis_j in HBM
 * and factorizes (same return convention as refactorize). */
int gmrf_b200_set_value_basis(gmrf_b200_handle *h, const double *basis, int nbasis);
int gmrf_b200_refactorize_combination(gmrf_b200_handle *h, const double *coeff, int nbasis);

/* Newton loops with a diagonal observation Hessian: set_base_values keeps the prior's nzval in HBM; every iterate then
 * refactorizes Q_prior - Diagonal(diag) from n doubles. replaces the host rebuild + upload in _update_hessian!
 * (src/workspace/gaussian_approximation.jl:96-129, _subtract_diagonal_hessian! :63-72) + refactorize! (backend.jl:178-189). */
int gmrf_b200_set_base_values(gmrf_b200_handle *h, const double *nzval, int64_t nnz);
int gmrf_b200_refactorize_base_minus_diag(gmrf_b200_handle *h, const double *diag, int64_t n);
/* The same for a SPARSE observation Hessian whose pattern lies inside the workspace's. replaces _sparse_hessian_map (built
 * once per Newton loop, gaussian_approximation.jl:31-61) + _subtract_sparse_hessian! (:74-83) + refactorize!:
 * set_hessian_pattern uploads the nzval positions (in `index_base`) of the Hessian's stored entries once -- they must be
 * distinct (a CSC matrix has no duplicates; one owner per entry, no floating-point atomics) --, every iterate then moves
 * `count` values and refactorizes Q_prior - H formed in HBM. */
int gmrf_b200_set_hessian_pattern(gmrf_b200_handle *h, const int64_t *nzpos, int64_t count, int index_base);
int gmrf_b200_refactorize_base_minus_sparse(gmrf_b200_handle *h, const double *values, int64_t count);

/* Lanes: B independent value sets of the SAME pattern factorized side by side by the same launches -- the workload of a
 * hyperparameter sweep over a `WorkspacePool` (src/workspace/workspace_pool.jl:42-119, `(model)(ws; theta...)`
 * src/workspace/latent_model_integration.jl:151-185) whose outputs are log-determinants. Capacity is the process-wide
 * option "lanes" at create time (memory: B copies of the numeric arrays). nzval holds `lanes` arrays of nnz values back
 * to back (coeff: lanes x nbasis); logdet[lanes] and status[lanes] (0 ok / k>0 first non-positive pivot; may be NULL) are
 * filled. Lane 0 remains the handle's ordinary factor for solves and selected inversion. */
int gmrf_b200_lane_capacity(const gmrf_b200_handle *h);
int gmrf_b200_refactorize_lanes(gmrf_b200_handle *h, const double *nzval, int64_t nnz, int lanes, double *logdet, int *status);
int gmrf_b200_refactorize_combination_lanes(gmrf_b200_handle *h, const double *coeff, int nbasis, int lanes, double *logdet,
                                            int *status);

/* replaces  compute_logdet(b) = logdet(factor)   backend.jl:211-213 */
int gmrf_b200_logdet(gmrf_b200_handle *h, double *out);

/* ---- solves ----------------------------------------------------------------------------------
 * replaces  backend_solve(b, rhs::Vector) / (b, RHS::Matrix)   backend.jl:191-209
 * X = Q^-1 B, original ordering, B and X are n x nrhs column-major with leading dimension ld
 * (B == X allowed). */
int gmrf_b200_solve(gmrf_b200_handle *h, const double *B, double *X, int64_t ld, int64_t nrhs);
/* replaces  backend_backward_solve(b, x) = factor.UP \ x   backend.jl:281-284
 * X = P' L^-T Z so that Cov(X) = Q^-1 for Z ~ N(0, I) (sampling). */
int gmrf_b200_solve_Lt(gmrf_b200_handle *h, const double *Z, double *X, int64_t ld, int64_t nrhs);
/* Same two operations on buffers resident in this device's HBM. Stream contract of every *_device entry point: the
 * library works on the handle's own (non-blocking) stream and returns after that stream has drained; it does NOT wait
 * for work the caller queued on other streams, so buffers produced elsewhere (e.g. on torch's current stream) must be
 * complete -- synchronize that stream or the device -- before the call. */
int gmrf_b200_solve_device(gmrf_b200_handle *h, const double *dB, double *dX, int64_t ld, int64_t nrhs);
int gmrf_b200_solve_Lt_device(gmrf_b200_handle *h, const double *dZ, double *dX, int64_t ld, int64_t nrhs);

/* ---- selected inversion (Takahashi) ----------------------------------------------------------
 * replaces  compute_selinv! / get_selinv_Z   backend.jl:215-236  (SelectedInversion.selinv(F).Z)
 * Runs the recursion once per refactorization and caches Z on the device; the getters below call
 * it on demand (the reference's compute_selinv! is lazy too). */
int gmrf_b200_selinv_compute(gmrf_b200_handle *h);
/* replaces  get_selinv_diag(b)   backend.jl:248-257 ; bit-identical to diag of selinv_values */
int gmrf_b200_selinv_diag(gmrf_b200_handle *h, double *out);
/* replaces  get_selinv(b) = sparse(Z)   backend.jl:238-246 : full symmetric CSC on the FACTOR's
 * pattern (superset of Q's), original ordering, sorted rows. Pattern depends on the symbolic
 * analysis only: fetch it once, then only values after each refactorize. */
int gmrf_b200_selinv_nnz(gmrf_b200_handle *h, int64_t *nnz);
int gmrf_b200_selinv_pattern(gmrf_b200_handle *h, int64_t *colptr, int64_t *rowval, int index_base);
int gmrf_b200_selinv_values(gmrf_b200_handle *h, double *nzval);
/* replaces  selinv_extract_at(b, B)   backend.jl:275-279 : Sigma read at a caller pattern (CSC,
 * nnz entries), 0.0 where the position is outside the factor's pattern. */
int gmrf_b200_selinv_extract(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr,
                             const int64_t *rowval, int index_base, double *out);
/* replaces  selinv_dot(b, B) = dot(Z, B) = tr(Q^-1 B)   backend.jl:265-267 (generic fallback :30), for
 * Float64-valued B (n x n CSC, both triangles as stored): Sigma is gathered at B's pattern and contracted
 * on the device (fixed-shape reduction, bit-reproducible); positions outside the factor's pattern count 0.
 * Dual-valued B (ext/forwarddiff/logdetcov.jl:23) keeps using selinv_extract + a host dot. */
int gmrf_b200_selinv_dot(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                         int index_base, const double *values, double *out);
/* out[j] = tr(Q^-1 B_j) for the resident value basis (set_value_basis): with Q(theta) = sum_j c_j(theta) B_j
 * this is d logdet Q / d c_j, the contraction the logdetcov / logpdf pullbacks (src/workspace/autodiff.jl:8-91,
 * compute_precision_gradient src/autodiff/precision_gradient.jl:137-146) apply to Q-bar = c * selinv(ws)
 * when Q is a fixed-pattern combination. The B_j are read as symmetric matrices through their stored upper
 * triangle, like the factorization reads Q. Nothing but nbasis doubles crosses PCIe. */
int gmrf_b200_selinv_dot_basis(gmrf_b200_handle *h, double *out, int nbasis);

/* replaces  _row_diag_AΣAt(ws, A)   src/linear_predictor_marginals.jl:137-165 : out[i] = sum_{j,k} A_ij A_ik Sigma_jk for
 * the m rows of a sparse design matrix A (CSR, m x n, indices in `index_base`) -- the marginal variances of a linear
 * predictor eta = A x -- contracted against Z on the device (one warp per row, fixed reduction order). Sigma entries
 * outside the factor's pattern count 0, like selinv_extract. */
int gmrf_b200_selinv_quadform_rows(gmrf_b200_handle *h, int64_t m, const int64_t *rowptr, const int64_t *colidx,
                                   const double *values, int index_base, double *out);

/* ---- factor export ------------------------------------------------------------------------------
 * replaces  sparse_cho_sqrt(cho) = sparse(cho.L)[invperm(cho.p), :]   src/linear_maps/cholesky_sqrt.jl:6-21
 * (CholeskySqrt, used by cholesky_factorized_map.jl:40): the square root R = P'L of Q as CSC, n x n,
 * column k = k-th pivot of the elimination order, row indices in the ORIGINAL numbering, sorted, so
 * that R R' = Q. The pattern is the stored (relaxed-supernode) one: a superset of the exact pattern
 * of L whose extra entries are explicit zeros; it depends on the symbolic analysis only (fetch once,
 * valid on analysis-only handles), values after each refactorize. */
int gmrf_b200_factor_nnz(gmrf_b200_handle *h, int64_t *nnz);
int gmrf_b200_factor_pattern(gmrf_b200_handle *h, int64_t *colptr, int64_t *rowval, int index_base);
int gmrf_b200_factor_values(gmrf_b200_handle *h, double *nzval);

/* ---- introspection (symbolic facts; all host-side, valid on analysis-only handles) -----------*/
enum {
    GMRF_B200_INFO_N = 0,
    GMRF_B200_INFO_NNZ_Q = 1,          /* nnz of the full symmetric input pattern */
    GMRF_B200_INFO_NNZ_L = 2,          /* exact nnz(L) = sum of column counts */
    GMRF_B200_INFO_NNZ_L_STORED = 3,   /* doubles in the supernodal panels (incl. relaxation zeros + padding) */
    GMRF_B200_INFO_NSUPER = 4,
    GMRF_B200_INFO_NLEVELS = 5,
    GMRF_B200_INFO_MAX_FRONT = 6,      /* largest front order (ns + nr) */
    GMRF_B200_INFO_MAX_NS = 7,
    GMRF_B200_INFO_UPDATE_POOL = 8,    /* doubles in the update-matrix pool */
    GMRF_B200_INFO_FLOPS_CHOL = 9,     /* sum_j cc_j^2 (exact), the algorithmic factorization flop count */
    GMRF_B200_INFO_FLOPS_CHOL_STORED = 10, /* flops actually executed on the relaxed panels */
    GMRF_B200_INFO_DEVICE_BYTES = 11,
    GMRF_B200_INFO_GRAPH_NODES = 12,   /* kernel launches in one refactorization */
    GMRF_B200_INFO_SELINV_NODES = 13,  /* kernel launches in one selected inversion */
    GMRF_B200_INFO_PATTERN_CACHE_HITS = 14, /* selinv_extract / selinv_dot calls that reused the previous pattern's lookup */
    GMRF_B200_INFO_LARGE_TILE_LAUNCHES = 15, /* factorization launches of the 128 x 64-tile GEMM (plan introspection for tests) */
    GMRF_B200_INFO_SPLITK_TASKS = 16,  /* products of the factorization plan that were cut along k */
    GMRF_B200_INFO_FAST_ROOTS = 17,    /* root supernodes on the triangular route of the selected inversion (after its plan exists) */
    GMRF_B200_INFO_CHAIN_LAUNCHES = 18, /* fused chain-step launches (one per 128 columns of a level's chains) */
    GMRF_B200_INFO_FRONT_LAUNCHES = 19, /* one-CTA-per-front launches (one per tree level of small fronts) */
    GMRF_B200_INFO_COUNT = 20
};
int gmrf_b200_info(const gmrf_b200_handle *h, int64_t *info, int n_info);
int gmrf_b200_get_perm(const gmrf_b200_handle *h, int64_t *perm, int index_base);       /* final elimination order */
int gmrf_b200_get_colcounts(const gmrf_b200_handle *h, int64_t *colcount);              /* exact, elimination order */
int gmrf_b200_get_etree(const gmrf_b200_handle *h, int64_t *parent);                    /* 0-based, -1 = root */
/* Supernodal schedule tables (0-based), used by the host-side replay test:
 *   super_ptr[nsuper+1], super_parent[nsuper], level[nsuper], row_ptr[nsuper+1],
 *   panel_off[nsuper+1], panel_ld[nsuper], upd_off[nsuper], upd_ld[nsuper]. Any pointer may be NULL. */
int gmrf_b200_get_supernodes(const gmrf_b200_handle *h, int64_t *super_ptr, int64_t *super_parent,
                             int64_t *level, int64_t *row_ptr, int64_t *panel_off, int64_t *panel_ld,
                             int64_t *upd_off, int64_t *upd_ld);
int gmrf_b200_get_rows(const gmrf_b200_handle *h, int64_t *row_idx, int64_t *rel_idx);  /* row_ptr[nsuper] entries each */
int gmrf_b200_get_scatter(const gmrf_b200_handle *h, int64_t *n_entries, int64_t *src, int64_t *dst);
/* panel offset of every entry of a caller pattern (n x n CSC; -1 outside the factor's stored pattern): the host-side
 * lookup behind selinv_extract / selinv_dot, exposed for tests of the index arithmetic without a device */
int gmrf_b200_pattern_positions(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                                int index_base, int64_t *pos);
/* Wall-clock of the last call's phases in milliseconds (CUDA events on the handle's stream):
 * [0] h2d of nzval, [1] numeric factorization + logdet, [2] solve, [3] selinv, [4] host analysis. */
int gmrf_b200_last_timings(const gmrf_b200_handle *h, double *ms, int n);
/* Live per-kernel-family profile of one refactorization with the values of the previous refactorize (graphs off,
 * CUDA event pairs around every launch on the handle's stream): ms[4]/count[4] for {0: DMMA GEMM (incl. TRSM by
 * inverted block), 1: panel potrf + inverse, 2: extend-add assembly, 3: memset+scatter+logdet}; *gemm_flops = algorithmic flops of the GEMM tasks. */
int gmrf_b200_profile_refactorize(gmrf_b200_handle *h, double *ms, int64_t *count, double *gemm_flops);
/* Per-launch device times of one phase (diagnostics; graphs off, event pair around every launch): phase 0 = numeric
 * factorization, 1 = selected inversion, 2 = forward sweep, 3 = backward sweep with `nrhs` (<= 8) columns. Up to `cap`
 * entries of kind[] (internal launch kind), grid[] (CTAs) and ms[] are filled; *count = launches in the phase. */
int gmrf_b200_profile_plan(gmrf_b200_handle *h, int phase, int nrhs, int64_t cap, int *kind, int *grid, double *ms,
                           int64_t *count);
/* Static facts of the launches of a phase, in profile_plan's order: algorithmic flops and longest contraction of the GEMM
 * launches, tasks per launch (diagnostics: per-launch TFLOP/s = flops / ms). */
int gmrf_b200_plan_launch_info(gmrf_b200_handle *h, int phase, int64_t cap, double *flops, int *kmax, int *ntasks, int64_t *count);
/* Diagnostics: SM clock stamps of the phases of the last fused chain-step launch of one refactorization (7 values). */
int gmrf_b200_debug_chain_phases(gmrf_b200_handle *h, int64_t *stamps, int n);
/* Numeric factor / selected inverse panels copied back to the host (tests, CholeskySqrt-style export). */
int gmrf_b200_get_factor_panels(gmrf_b200_handle *h, double *Lx, int64_t n_doubles);
int gmrf_b200_get_selinv_panels(gmrf_b200_handle *h, double *Zx, int64_t n_doubles);

/* ---- factor sharing inside a multi-GPU pool (WorkspacePool, src/workspace/workspace_pool.jl:42-67, one workspace per
 * GPU with the SAME ordering): device pointer + length in doubles of which = 0 the supernodal panels of L, 1 the
 * inverted 64-column diagonal blocks, 2 the selected-inverse panels. One rank factorizes, the arrays travel to the
 * peers (ncclBroadcast over NVLink) and each peer calls adopt_factor; its solves / samples then need no factorization
 * of their own (right-hand-side blocks sharded over GPUs). */
int gmrf_b200_device_array(gmrf_b200_handle *h, int which, void **ptr, int64_t *n_doubles);
int gmrf_b200_adopt_factor(gmrf_b200_handle *h, double logdet, int with_selinv);
/* The same with a guard: `fingerprint` (gmrf_b200_analysis_fingerprint of the SENDER: pattern, elimination order,
 * supernode partition, panel and inverse-block layout) must equal the receiver's, else GMRF_B200_ERR_ARG; the sender's
 * factorization status (0 / first non-positive pivot) becomes the receiver's and is returned. */
int gmrf_b200_analysis_fingerprint(const gmrf_b200_handle *h, uint64_t *fingerprint);
int gmrf_b200_adopt_factor_checked(gmrf_b200_handle *h, uint64_t sender_fingerprint, double logdet, int sender_status,
                                   int with_selinv);

/* Page-lock / unlock a caller-owned host buffer (typically the workspace's nzval array, `ws.Q.nzval`) so that
 * gmrf_b200_refactorize moves it with an asynchronous DMA. Optional; ownership stays with the caller. */
int gmrf_b200_host_register(void *ptr, int64_t bytes);
int gmrf_b200_host_unregister(void *ptr);
/* dst[0..n) = src[0..n) (doubles) with all host threads: the host mirror `ws.Q.nzval .= nzval` of update_precision_values
 * (src/workspace/gmrf_workspace.jl:154-165) for callers without a threaded copy of their own. Regions must not overlap. */
int gmrf_b200_host_copy(double *dst, const double *src, int64_t n);

/* Tunables, set BEFORE create (process-wide defaults): key in {"relax_n0","relax_n1","relax_n2",
 * "relax_z0","relax_z1","relax_z2","use_graph","outer_block","naive_kernels","selinv_fast_root","splitk_min_k","wide_rhs_min","bwd_row_chunk","large_tile_mask","lanes","fused_front","fused_chain","chain_max_tiles","front_smem_kb","asm_gather","syrk_gather","level_alap","wide_steps","syrk_split",
 * "pdl" (few-RHS sweeps launched with programmatic stream serialization, default 1), "pdl_multi" (same for the wide right-hand-side
 * GEMM sweeps, default 1), "pdl_factor" (same for the factorization / selected-inversion plans, default 0)}. */
int gmrf_b200_set_option(const char *key, double value);

/* Dense-kernel unit-test hooks (tests/ only). HOST pointers, column-major; operands are staged to `device`
 * and back. gemm: C[m x n] = (beta ? C : 0) + alpha * Aop * Bop^T, where transa/transb = 0 means the operand is
 * stored m x k (resp. n x k), 1 means k x m (resp. k x n); `flags` bits: 1 lower-only, 2 alpha=+1 (default -1),
 * 4 add identity, 8 naive debug kernel, 16 large (128x128) tile. */
int gmrf_b200_test_gemm(int device, int transa, int transb, int flags, int m, int n, int k,
                        const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
int gmrf_b200_test_potrf(int device, int n, double *A, int lda, int *info);
/* y[i] = the kernels' straight-line reciprocal square root of x[i] (pivot tiles of the panel factorizations). */
int gmrf_b200_test_rsqrt(int device, int n, const double *x, double *y);
/* same kernel, also returning inv(L) (n x n, leading dimension n, zeros above the diagonal) */
int gmrf_b200_test_potrf_inv(int device, int n, double *A, int lda, double *inv, int *info);
/* Device-timed micro-benchmark of the library's own FP64 GEMM kernel (zero-filled device operands, best of
 * `reps`, CUDA events): used to compare against the measured cuBLAS DGEMM peak in profiles/. */
int gmrf_b200_bench_gemm(int device, int transa, int transb, int flags, int m, int n, int k, int reps, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* GMRF_B200_H */
