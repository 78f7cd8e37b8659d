// front_kernels.cuh -- fused kernels for the latency-bound part of the numeric factorization: the many small fronts at
// the bottom of the assembly tree and the 64-column chains of the separators above them (north star: "register-tiled
// kernels for the many small ones"). Measured on a B200 before these existed (profiles/r02_plan_2d224_before.log): a
// 50 k-dof 2D refactorization was 172 dependent launches -- per 64 columns of a chain one single-CTA potrf+inverse
// (31 us), one TRSM-by-inverse GEMM launch and one k = 64 update launch (12-15 us each at grids of 10-200 CTAs) -- and
// ran at 4 % of the FP64 roofline.
//
//   front_small_kernel   ONE CTA per front, whole panel resident in shared memory: children pulled in (extend-add),
//                        diagonal blocks factored, rows below solved, update matrix formed and the children's
//                        remaining contributions added -- one launch per tree level instead of 4-5.
//   chain_step_kernel    ONE launch per 128 columns of the chains of a level, one CTA per 64-row tile and NO
//                        communication between CTAs: every CTA factors the 128 x 128 diagonal square REDUNDANTLY in its
//                        own shared memory (bit-identical everywhere: same code, same inputs) and then solves / updates
//                        its own rows. The diagonal CTA parks the factored square in a scratch array, because the other
//                        CTAs of the launch are still reading the unfactored one from the panel.
//   chain_finalize_kernel  one launch per factorization: copies the parked squares into the panels and inverts every
//                        64-column diagonal block (the inverses serve the solve and selected-inversion phases; taking
//                        them off the chain's critical path is half of the gain).
//
// Determinism: every entry has one owner thread and a fixed summation order; no floating-point atomics.
#pragma once
#include "kernels.cuh"

namespace gmrf {

constexpr int FB = 64;                 // block order
constexpr int FLD = 72;                // leading dimension of a 64 x 64 tile in shared memory (column-major; even ->
                                       // 16-byte aligned row pairs; 72 = 8 mod 16 makes the DMMA fragment loads -- 8 rows
                                       // of 4 consecutive columns per warp -- hit every bank exactly twice, the minimum)
constexpr int FTILE = FB * FLD;        // doubles per tile
constexpr int FPLD = 68;               // leading dimension of the published 4-column panel (k-major)
constexpr int FSCRATCH = 2 * 3 * 4 * FPLD + FB;   // panel-step scratch: double-buffered published panel (diagonal tile + 2 row tiles) + reciprocal pivots

// ------------------------------------------------------------------------------------------------
// Panel step on 64 columns in shared memory, 256 threads: Cholesky of the diagonal tile sD (column-major
// sD[j * FLD + i], identity-padded beyond the true order nb) TOGETHER with the solve X L^T = B of up to RT further
// 64-row tiles of the same columns -- i.e. the factorization of a (1 + RT) * 64 x 64 panel. Everything is held as
// 4 x 4 register tiles on a 16 x 16 thread grid (thread (ty, tx) owns rows 4ty.., columns 4tx.. of every tile);
// nb/4 column-panel steps, rolled (the first version solved the rows with a fully unrolled one-thread-per-row
// substitution: 4000 straight-line instructions per call that two warps execute once -- ncu showed 80 % of its
// samples stalled on instruction fetch, 20 us per call):
//   (a) the diagonal thread factors its 4 x 4 tile and publishes it                         [FACTOR only]
//   (b) the threads of column-panel P solve their tiles against it and publish the (1 + RT) * 64 x 4 panel,
//   (c) everybody to the right applies the rank-4 update from the published panel (k-major: conflict-free reads).
// FACTOR = false: sD already holds L (and srinv the reciprocal pivots); only the row tiles are solved.
// `scratch`: panel_scratch_doubles(RT) doubles (FSCRATCH covers RT <= 2). Non-positive pivots: integer atomicMin of the 1-based column, NaNs propagate.
// ------------------------------------------------------------------------------------------------
// Cholesky of the 4 x 4 diagonal tile held by one thread (chol4x4_lower: two rsqrt latencies); publishes the tile (k-major)
// and the reciprocal pivots.
__device__ __forceinline__ void factor_diag_tile(double (&a)[4][4], double *__restrict__ pb, double *__restrict__ srinv, int P, int nb,
                                                 int col0, int *fail_col, bool report) {
    double ri[4], piv[4];
    chol4x4_lower(a, ri, piv);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (report && !(piv[c] > 0.0) && 4 * P + c < nb) atomicMin(fail_col, col0 + 4 * P + c + 1);
        srinv[4 * P + c] = ri[c];
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) pb[c * FPLD + 4 * P + r] = (c <= r) ? a[r][c] : 0.0;
}

struct RowTile {
    double *p;      // element (i, j) of the tile at p[j * ld + i]
    int ld, nr;     // rows >= nr read as zero and are not written back
};
__host__ __device__ constexpr int panel_scratch_doubles(int RT) { return 2 * (1 + RT) * 4 * FPLD + FB; }

template <int RT, bool FACTOR>
__device__ __forceinline__ void panel_solve64(double *__restrict__ sD, double *__restrict__ scratch, const RowTile (&rt)[RT > 0 ? RT : 1],
                                              int nb, int col0, int *fail_col, bool report) {
    constexpr int PB = (1 + RT) * 4 * FPLD;          // doubles per published panel (diagonal tile rows first)
    double *srinv = scratch + 2 * PB;
    // column tiles are spread over the WARPS (warp w owns tx = 2w, 2w + 1), row tiles over the lanes: at step P only the
    // warps with a column tile right of P issue anything (with tx on the lanes every warp kept issuing for a shrinking
    // set of active lanes: ncu counted 2.4x the ideal FP64 issue slots), and all shared-memory reads are 512-byte rows
    const int tid = threadIdx.x, ty = tid & 15, tx = tid >> 4;
    const bool active = ty >= tx;
    const int nsteps = (nb + 3) >> 2;                // the identity padding needs no work
    double a[4][4];
    double x[RT > 0 ? RT : 1][4][4];
    if (FACTOR) {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++) a[r][c] = active ? sD[(4 * tx + c) * FLD + 4 * ty + r] : 0.0;
        for (int i = 4 * nsteps + tid; i < FB; i += 256) srinv[i] = 1.0;   // (entries below are written by the diagonal threads)
    }
#pragma unroll
    for (int q = 0; q < RT; q++)
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                x[q][r][c] = (4 * ty + r < rt[q].nr && 4 * tx + c < nb) ? rt[q].p[(4 * tx + c) * rt[q].ld + 4 * ty + r] : 0.0;
#pragma unroll 1
    for (int P = 0; P < nsteps; P++) {
        double *pb = scratch + (P & 1) * PB;
        if (FACTOR && P == 0) {
            if (ty == 0 && tx == 0) factor_diag_tile(a, pb, srinv, 0, nb, col0, fail_col, report);
            __syncthreads();
        }
        if (tx == P) {
            // X * L_PP^T = B on 4 x 4 tiles, column by column
            double lpp[4][4], ri[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                ri[c] = srinv[4 * P + c];
#pragma unroll
                for (int k = 0; k < 4; k++) lpp[c][k] = FACTOR ? pb[k * FPLD + 4 * P + c] : sD[(4 * P + k) * FLD + 4 * P + c];
            }
            if (FACTOR && ty > P) {
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        double s = a[r][c];
#pragma unroll
                        for (int k = 0; k < c; k++) s -= a[r][k] * lpp[c][k];
                        a[r][c] = s * ri[c];
                    }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pb[c * FPLD + 4 * ty + r] = a[r][c];
            }
#pragma unroll
            for (int q = 0; q < RT; q++) {
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        double s = x[q][r][c];
#pragma unroll
                        for (int k = 0; k < c; k++) s -= x[q][r][k] * lpp[c][k];
                        x[q][r][c] = s * ri[c];
                    }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pb[(4 * (q + 1) + c) * FPLD + 4 * ty + r] = x[q][r][c];
            }
        }
        __syncthreads();
        if (tx > P) {
            double pc[4][4];          // L[4tx + c][4P + k]
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int c = 0; c < 4; c++) pc[c][k] = FACTOR ? pb[k * FPLD + 4 * tx + c] : sD[(4 * P + k) * FLD + 4 * tx + c];
            if (FACTOR && active) {
                double pr[4][4];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pr[r][k] = pb[k * FPLD + 4 * ty + r];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int k = 0; k < 4; k++) a[r][c] -= pr[r][k] * pc[c][k];
                // look-ahead: the next diagonal tile is final now -- factor it while the other warps are still updating
                if (ty == P + 1 && tx == P + 1 && P + 1 < nsteps)
                    factor_diag_tile(a, scratch + ((P + 1) & 1) * PB, srinv, P + 1, nb, col0, fail_col, report);
            }
#pragma unroll
            for (int q = 0; q < RT; q++) {
                double pr[4][4];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pr[r][k] = pb[(4 * (q + 1) + k) * FPLD + 4 * ty + r];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int k = 0; k < 4; k++) x[q][r][c] -= pr[r][k] * pc[c][k];
            }
        }
        if (FACTOR) __syncthreads();      // the next diagonal tile is published
        // (!FACTOR: the next step publishes into the other panel buffer, one barrier per step is enough)
    }
    if (FACTOR && active) {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++) sD[(4 * tx + c) * FLD + 4 * ty + r] = (4 * tx + c <= 4 * ty + r) ? a[r][c] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < RT; q++)
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (4 * ty + r < rt[q].nr && 4 * tx + c < nb) rt[q].p[(4 * tx + c) * rt[q].ld + 4 * ty + r] = x[q][r][c];
    __syncthreads();
}

// Warp-level FP64 tensor-core product on shared-memory operands (mma.sync.m8n8k4.f64 -> DMMA.8x8x4): the warp accumulates
// a 32 x 16 tile  acc += A[r0.., :K] * B[c0.., :K]^T  with A(r, k) at A[k * lda + r] and B(c, k) at B[k * ldb + c]
// (both column-major in k); rows >= arows / brows and columns k >= K read as zero. Fragment layout as in gemm_dmma_kernel:
// lane = 4 grp + tig holds acc[i][j][e] = C[8 i + grp][8 j + 2 tig + e]. Per 4 k's: 6 fragment loads, 8 DMMA -- the
// 4 x 4 register-tile FMA loops this replaces were bound by shared-memory bandwidth (8 loads per 16 FMA: ncu, 10 us for
// the two rank-64 updates of a chain step against 2 us of FP64 issue time).
__device__ __forceinline__ void warp_dmma_32x16(const double *__restrict__ A, int lda, int r0, int arows, const double *__restrict__ B, int ldb,
                                                int c0, int brows, int K, double (&acc)[4][2][2]) {
    const int lane = threadIdx.x & 31, grp = lane >> 2, tig = lane & 3;
    bool aok[4], bok[2];
#pragma unroll
    for (int i = 0; i < 4; i++) aok[i] = r0 + 8 * i + grp < arows;
#pragma unroll
    for (int j = 0; j < 2; j++) bok[j] = c0 + 8 * j + grp < brows;
    const double *ap = A + r0 + grp + tig * lda, *bp = B + c0 + grp + tig * ldb;
    for (int k0 = 0; k0 < K; k0 += 4) {
        const bool kok = k0 + tig < K;
        double a[4], b[2];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = (aok[i] && kok) ? ap[k0 * lda + 8 * i] : 0.0;
#pragma unroll
        for (int j = 0; j < 2; j++) b[j] = (bok[j] && kok) ? bp[k0 * ldb + 8 * j] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// The same panel factorization, blocked: [D; rows of rt[0..RT)] (64 columns) is factored 16 columns at a time.
//   * inside a 16-column sub-panel the rank-4 steps run on ONE 4 x 4 register tile per thread -- thread = (group g: D or a
//     row tile, row tile ty, column tile ctl of the sub-panel), 64 (1 + RT) threads -- so the tile solves of a step are spread
//     over 1 + RT half-warps instead of one, and the rank-4 updates touch 16 columns instead of up to 64;
//   * the columns right of the sub-panel get ONE rank-16 update from shared memory on the FP64 tensor cores
//     (warp_dmma_32x16, 8 warps).
// panel_solve64<RT, true> updates every remaining column after every 4 columns from 4 x 4 register tiles (64 shared-memory
// loads per thread for 192 FMAs) and its tile solves are done by a single warp holding 1 + RT tiles per thread.
// Measured (clock64 stamps inside the kernel, profiles/r02_chain_phases.log): a rank-4 step of the sub-panel is ~1,400
// cycles instead of ~2,400 -- the tile solves ~100, the 4 x 4 update of ONE tile per thread still ~750 (shared-memory
// instruction throughput: 32 LDS.64 per thread, not bank conflicts -- a [k][r][row-tile] layout of the published panels
// without any conflict measured slower), the look-ahead pivot tile ~450 -- but the rank-16 update costs ~9,200 cycles for the
// first sub-panel: ~1,500 per 32 x 16 warp tile for 32 DMMAs (the SM's DMMA rate, equal to its DFMA rate) and ~2,000 for
// reading and writing the 16 accumulators of C through shared memory. Net: the 192-row panel of a chain step is 10 %
// slower (22 us against 20), the one-CTA fronts are faster; 2D factorizations gain 2-3 %, a 16-lane sweep 5 %.
template <int RT>
__device__ __forceinline__ void panel_factor64b(double *__restrict__ sD, double *__restrict__ scratch, const RowTile (&rt)[RT], int nb, int col0,
                                                int *fail_col, bool report, long long *prof = nullptr) {
    static_assert(RT >= 1 && RT <= 2, "row-tile groups");
    constexpr int PB = (1 + RT) * 4 * FPLD;
    double *srinv = scratch + 2 * PB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = tid >> 6;                            // 0: rows of the diagonal block, 1..RT: row tile g - 1
    const int ty = tid & 15, ctl = (tid >> 4) & 3;     // row tile inside the group, column tile inside the sub-panel
    const bool have = g <= RT;
    const RowTile R = (g >= 1 && g <= RT) ? rt[g - 1] : RowTile{sD, FLD, FB};
    const int nsteps = (nb + 3) >> 2;
    for (int i = 4 * nsteps + tid; i < FB; i += 256) srinv[i] = 1.0;
#pragma unroll 1
    for (int s = 0; 4 * s < nsteps; s++) {
        const int tx = 4 * s + ctl;
        const bool act = have && tx < nsteps && (g > 0 || ty >= tx);
        double a[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                a[r][c] = (act && 4 * ty + r < R.nr && (g == 0 || 4 * tx + c < nb)) ? R.p[(4 * tx + c) * R.ld + 4 * ty + r] : 0.0;
        const int jmax = min(4, nsteps - 4 * s);
        if (prof && s == 0 && tid == 0) prof[5] = clock64();
#pragma unroll 1
        for (int j = 0; j < jmax; j++) {
            const int P = 4 * s + j;
            double *pb = scratch + (P & 1) * PB;
            const bool stamp = prof && P == 1 && tid == 34;          // (the look-ahead thread of step 1: g = 0, ty = 2, ctl = 2)
            if (stamp) prof[0] = clock64();
            if (j == 0) {          // (no look-ahead across sub-panels: this tile was final only after the rank-16 update)
                if (g == 0 && ty == P && ctl == 0) factor_diag_tile(a, pb, srinv, P, nb, col0, fail_col, report);
                __syncthreads();
            }
            if (act && ctl == j && (g > 0 || ty > P)) {
                // X * L_PP^T = B on the 4 x 4 tile, column by column
                double lpp[4][4], ri[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    ri[c] = srinv[4 * P + c];
#pragma unroll
                    for (int k = 0; k < 4; k++) lpp[c][k] = pb[k * FPLD + 4 * P + c];
                }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        double sum = a[r][c];
#pragma unroll
                        for (int k = 0; k < c; k++) sum -= a[r][k] * lpp[c][k];
                        a[r][c] = sum * ri[c];
                    }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pb[(4 * g + c) * FPLD + 4 * ty + r] = a[r][c];
            }
            __syncthreads();
            if (stamp) prof[1] = clock64();
            if (act && ctl > j) {
                double pc[4][4], pr[4][4];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int c = 0; c < 4; c++) pc[c][k] = pb[k * FPLD + 4 * tx + c];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int r = 0; r < 4; r++) pr[r][k] = pb[(4 * g + k) * FPLD + 4 * ty + r];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int k = 0; k < 4; k++) a[r][c] -= pr[r][k] * pc[c][k];
                // look-ahead: the next diagonal tile of the sub-panel is final now
                if (stamp) prof[2] = clock64() + (long long)(a[0][0] == 12345.678);      // (keeps the stamp behind the updates)
                if (g == 0 && ty == P + 1 && ctl == j + 1) factor_diag_tile(a, scratch + ((P + 1) & 1) * PB, srinv, P + 1, nb, col0, fail_col, report);
                if (stamp) prof[3] = clock64();
            }
            __syncthreads();
            if (stamp) prof[4] = clock64();
        }
        if (prof && s == 0 && tid == 0) prof[6] = clock64();
        // the sub-panel goes back to shared memory ...
        if (act) {
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (4 * ty + r < R.nr && (g == 0 || 4 * tx + c < nb))
                        R.p[(4 * tx + c) * R.ld + 4 * ty + r] = (g == 0 && 4 * tx + c > 4 * ty + r) ? 0.0 : a[r][c];
        }
        __syncthreads();
        if (prof && s == 0 && tid == 0) prof[7] = clock64();
        // ... and updates the columns to its right: C[r, c] -= sum_{k in sub-panel} L[r, k] L[c, k], 32 x 16 warp tiles
        const int cbeg = 16 * (s + 1), cend = min(FB, (4 * nsteps + 15) & ~15);
        if (cbeg < cend) {
            const int ncb = (cend - cbeg) >> 4;
            const int rb0 = (FB - cbeg + 31) >> 5;                 // 32-row blocks of the diagonal block (rows cbeg..63)
            const int nrb = rb0 + 2 * RT;
            for (int w = warp; w < nrb * ncb; w += 8) {
                const int rb = w / ncb, cb = cbeg + 16 * (w - rb * ncb);
                const int G = rb < rb0 ? 0 : 1 + ((rb - rb0) >> 1);
                const RowTile C = G == 0 ? RowTile{sD, FLD, FB} : rt[G - 1];
                const int r0 = G == 0 ? cbeg + 32 * rb : 32 * ((rb - rb0) & 1);
                if (r0 >= C.nr) continue;
                double acc[4][2][2];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) acc[i][jj][0] = acc[i][jj][1] = 0.0;
                const bool stamp2 = prof && s == 0 && w == 0 && lane == 0;
                if (stamp2) prof[9] = clock64();
                warp_dmma_32x16(C.p + 16 * s * C.ld, C.ld, r0, C.nr, sD + 16 * s * FLD, FLD, cb, FB, 16, acc);
                if (stamp2) prof[10] = clock64() + (long long)(acc[0][0][0] == 12345.678);
                const int grp = lane >> 2, tig = lane & 3;
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int jj = 0; jj < 2; jj++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int r = r0 + 8 * i + grp, c = cb + 8 * jj + 2 * tig + e;
                            if (r < C.nr && (G == 0 ? r >= c : c < nb)) C.p[c * C.ld + r] -= acc[i][jj][e];
                        }
                if (stamp2) prof[11] = clock64();
            }
            __syncthreads();
        }
        if (prof && s == 0 && tid == 0) prof[8] = clock64();
    }
}

// C[64 x 64] -= A[64 x 64] B[64 x 64]^T on column-major shared-memory tiles (leading dimension FLD), 8 warps x (32 x 16);
// lower: warp tiles strictly above the diagonal are skipped (entries above the diagonal inside the others are computed
// and ignored by the caller).
__device__ __forceinline__ void rank64_update(double *__restrict__ sC, const double *__restrict__ sA, const double *__restrict__ sB, bool lower) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 2, tig = lane & 3;
    const int r0 = (warp & 1) * 32, c0 = (warp >> 1) * 16;
    if (lower && c0 >= r0 + 32) return;
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    warp_dmma_32x16(sA, FLD, r0, FB, sB, FLD, c0, FB, FB, acc);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) sC[(c0 + 8 * j + 2 * tig + e) * FLD + r0 + 8 * i + grp] -= acc[i][j][e];
}

// global (column-major, leading dimension ld) -> shared tile; entries outside nr x nc read as 0 (tri: only i >= j is
// read, the diagonal beyond the true order is padded with ones so that the padded tile stays positive definite)
// Asynchronous (cp.async, 8 bytes per element: the row offsets of a panel tile may be odd) -- a plain load / store
// loop serialises on in-order issue: the store of one element stalls the load of the next. Call tile_pad_identity
// after cp_async_wait + barrier for `tri` tiles.
__device__ __forceinline__ void load_tile(double *__restrict__ s, const double *__restrict__ g, long long ld, int nr, int nc, bool tri) {
    const int i = threadIdx.x & 63, j0 = threadIdx.x >> 6;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s + j0 * FLD + i);
    const double *src = g + i + (long long)j0 * ld;
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int j = j0 + 4 * q;
        const bool ok = i < nr && j < nc && (!tri || i >= j);
        cp_async8s(sa + q * 4 * FLD * 8, ok ? (const void *)(src + (long long)q * 4 * ld) : (const void *)g, ok ? 8 : 0);
    }
}
__device__ __forceinline__ void tile_pad_identity(double *__restrict__ s, int nb) {
    if (threadIdx.x < FB && threadIdx.x >= nb) s[threadIdx.x * FLD + threadIdx.x] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// Chain step: columns [k0, k0 + nb0 + nb1) of one supernode panel (nb0 <= 64; nb1 > 0 only if nb0 == 64).
// CTA 0 of a task is the diagonal CTA, CTA t >= 1 owns rows [k2 + 64 (t - 1), ...) with k2 = k0 + nb0 + nb1.
// ------------------------------------------------------------------------------------------------
struct ChainTask {
    double *P;          // panel base (column-major, leading dimension ld)
    double *sq;         // parked factored square: 128 x 128 doubles, column-major, leading dimension 128
    int ld, nrow;
    int k0, nb0, nb1;
    int col0;           // global (permuted) column of k0, for pivot reporting
};

// Phase timestamps (clock64) of tile 1 of the LAST chain launch, for tests/gpu_chain_phases.py: nullptr = off.
__device__ long long *g_chain_prof = nullptr;
#define CHAIN_STAMP(i) do { if (prof && tile == 1 && tid == 0) prof[i] = clock64(); } while (0)

constexpr int CHAIN_SMEM_BYTES = (5 * FTILE + FSCRATCH) * (int)sizeof(double);

template <bool BLK>
__global__ void __launch_bounds__(256, 1)
chain_step_kernel(const ChainTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int *__restrict__ fail_col,
                  long long bstride) {
    extern __shared__ __align__(16) double fsm[];
    double *sD0 = fsm, *sH = fsm + FTILE, *sD1 = fsm + 2 * FTILE, *sB0 = fsm + 3 * FTILE, *sB1 = fsm + 4 * FTILE;
    double *scr = fsm + 5 * FTILE;
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    ChainTask T = tasks[t];
    T.P = lane_ptr(T.P, bstride); T.sq = lane_ptr(T.sq, bstride); fail_col = lane_ptr(fail_col, bstride);
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x;
    const long long ld = T.ld;
    const bool two = T.nb1 > 0, diag = tile == 0;
    const int k1 = T.k0 + FB, k2 = T.k0 + T.nb0 + T.nb1;
    const int row0 = k2 + (tile - 1) * FB;
    const int nrv = diag ? 0 : min(FB, T.nrow - row0);
    long long *prof = g_chain_prof;
    pdl_wait();
    CHAIN_STAMP(0);
    // ---- loads (asynchronous, all in flight together) ----
    load_tile(sD0, T.P + (long long)T.k0 * ld + T.k0, ld, T.nb0, T.nb0, true);
    load_tile(sH, T.P + (long long)T.k0 * ld + k1, ld, T.nb1, FB, false);
    load_tile(sD1, T.P + (long long)k1 * ld + k1, ld, T.nb1, T.nb1, true);
    load_tile(sB0, T.P + (long long)T.k0 * ld + row0, ld, nrv, T.nb0, false);
    load_tile(sB1, T.P + (long long)k1 * ld + row0, ld, nrv, T.nb1, false);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    tile_pad_identity(sD0, T.nb0);
    tile_pad_identity(sD1, T.nb1);
    __syncthreads();
    CHAIN_STAMP(1);
    // ---- block 0: [D0; H; B0] is one 192 x 64 panel ----
    {
        const RowTile rt[2] = {{sH, FLD, T.nb1}, {sB0, FLD, nrv}};
        if constexpr (BLK) panel_factor64b<2>(sD0, scr, rt, T.nb0, T.col0, fail_col, diag, (prof && tile == 1) ? prof + 8 : nullptr);
        else panel_solve64<2, true>(sD0, scr, rt, T.nb0, T.col0, fail_col, diag);
    }
    CHAIN_STAMP(2);
    for (int e = tid; e < FB * FB; e += 256) {          // X0 -> HBM (coalesced along the rows)
        const int i = e & 63, j = e >> 6;
        if (i < nrv && j < T.nb0) T.P[(long long)(T.k0 + j) * ld + row0 + i] = sB0[j * FLD + i];
    }
    CHAIN_STAMP(3);
    if (two) {
        rank64_update(sD1, sH, sH, true);
        if (!diag) rank64_update(sB1, sB0, sH, false);
        __syncthreads();
        CHAIN_STAMP(4);
        const RowTile rt[1] = {{sB1, FLD, nrv}};
        if constexpr (BLK) panel_factor64b<1>(sD1, scr, rt, T.nb1, T.col0 + FB, fail_col, diag);
        else panel_solve64<1, true>(sD1, scr, rt, T.nb1, T.col0 + FB, fail_col, diag);
        CHAIN_STAMP(5);
        for (int e = tid; e < FB * FB; e += 256) {
            const int i = e & 63, j = e >> 6;
            if (i < nrv && j < T.nb1) T.P[(long long)(k1 + j) * ld + row0 + i] = sB1[j * FLD + i];
        }
    }
    if (diag) {
        // park the factored square (the other CTAs of this launch may still be reading the panel's copy)
        for (int e = tid; e < FB * FB; e += 256) {
            const int i = e & 63, j = e >> 6;
            T.sq[j * 128 + i] = sD0[j * FLD + i];
            if (two) {
                T.sq[j * 128 + 64 + i] = sH[j * FLD + i];
                T.sq[(64 + j) * 128 + 64 + i] = sD1[j * FLD + i];
            }
        }
    }
    CHAIN_STAMP(6);
}

// ------------------------------------------------------------------------------------------------
// Finalize: park -> panel copy of the factored diagonal squares and inversion of every 64-column diagonal block that
// a fused kernel produced (inv = L_kk^-1, nb x nb, leading dimension nb, zeros above the diagonal).
// ------------------------------------------------------------------------------------------------
struct FinalizeTask {
    const double *sq;   // parked square (leading dimension 128) or nullptr: the factor already sits in the panel
    double *P;          // panel + k0 * ld + k0 (the diagonal square's corner)
    double *inv0, *inv1;
    int ld, nb0, nb1, pad_;
};

constexpr int FINALIZE_SMEM_BYTES = (2 * FTILE + panel_scratch_doubles(1)) * (int)sizeof(double);

__global__ void __launch_bounds__(256, 2)
chain_finalize_kernel(const FinalizeTask *__restrict__ tasks, long long bstride) {
    extern __shared__ __align__(16) double fsm[];
    double *sL = fsm, *sY = fsm + FTILE, *scr = fsm + 2 * FTILE;
    double *srinv = scr + 2 * 2 * 4 * FPLD;
    const int which = blockIdx.x & 1;                    // CTA pair: block 0 / block 1 of the task
    FinalizeTask T = tasks[blockIdx.x >> 1];
    const bool parked = T.sq != nullptr;
    T.sq = lane_ptr(T.sq, bstride); T.P = lane_ptr(T.P, bstride); T.inv0 = lane_ptr(T.inv0, bstride); T.inv1 = lane_ptr(T.inv1, bstride);
    const int tid = threadIdx.x;
    const long long ld = T.ld;
    if (which == 1 && T.nb1 == 0) return;
    const int nb = which ? T.nb1 : T.nb0, o = which ? FB : 0;
    double *inv = which ? T.inv1 : T.inv0;
    for (int e = tid; e < FB * FB; e += 256) {
        const int i = e & 63, j = e >> 6;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j <= i) v = parked ? T.sq[(o + j) * 128 + o + i] : T.P[(o + i) + (o + j) * ld];
        sL[j * FLD + i] = v;
        sY[j * FLD + i] = (i == j) ? 1.0 : 0.0;
        if (parked) {
            if (i < nb && j <= i) T.P[(o + i) + (o + j) * ld] = v;
            if (which == 1 && i < nb) T.P[(FB + i) + j * ld] = T.sq[j * 128 + FB + i];     // the rows under block 0
        }
    }
    __syncthreads();
    if (tid < FB) srinv[tid] = 1.0 / sL[tid * FLD + tid];
    __syncthreads();
    // Y L^T = I  =>  Y = L^-T; the inverse is its transpose
    const RowTile rt[1] = {{sY, FLD, FB}};
    panel_solve64<1, false>(sL, scr, rt, nb, 0, nullptr, false);
    for (int e = tid; e < nb * nb; e += 256) {
        const int i = e % nb, c = e / nb;
        inv[e] = (i >= c) ? sY[i * FLD + c] : 0.0;        // inv(i, c) = Y(c, i)
    }
}

// ------------------------------------------------------------------------------------------------
// potrf + inverse of one 64 x 64 diagonal block per CTA (the bulk path's panel step), on the blocked panel code: the
// identity rides along as a row tile, so ONE pass gives L and Y = I L^-T whose transpose is L^-1 -- the rank-4 steps are
// latency bound and the extra group costs them nothing. potrf_inv64_kernel (kernels.cuh) factors without look-ahead
// (three barriers per 4 columns with one active thread / warp between them) and then inverts with 64 threads by a fully
// unrolled substitution while 192 threads wait: 30-37 us per launch, 7 barrier stalls per issue under ncu
// (profiles/r02_ncu_full_potrf_3d48.txt).
// ------------------------------------------------------------------------------------------------
constexpr int POTRF_LA_SMEM_BYTES = (2 * FTILE + panel_scratch_doubles(1)) * (int)sizeof(double);

__global__ void __launch_bounds__(256)
potrf_inv64_la_kernel(const PanelTask *__restrict__ tasks, int *__restrict__ fail_col, long long bstride) {
    extern __shared__ __align__(16) double fsm[];
    double *sL = fsm, *sY = fsm + FTILE, *scr = fsm + 2 * FTILE;
    pdl_launch_dependents();
    PanelTask T = tasks[blockIdx.x];
    T.D = lane_ptr(T.D, bstride); T.inv = lane_ptr(T.inv, bstride); fail_col = lane_ptr(fail_col, bstride);
    const int nb = T.nb, tid = threadIdx.x;
    pdl_wait();
    for (int e = tid; e < FB * FB; e += 256) {
        const int i = e & 63, j = e >> 6;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j <= i) v = T.D[i + (long long)j * T.ld];
        sL[j * FLD + i] = v;
        sY[j * FLD + i] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    const RowTile rt[1] = {{sY, FLD, nb}};
    panel_factor64b<1>(sL, scr, rt, nb, T.col0, fail_col, true);
    for (int e = tid; e < nb * nb; e += 256) {
        const int i = e % nb, c = e / nb;
        if (i >= c) T.D[i + (long long)c * T.ld] = sL[c * FLD + i];
        T.inv[e] = (i >= c) ? sY[i * FLD + c] : 0.0;        // inv(i, c) = Y(c, i)
    }
}

// ------------------------------------------------------------------------------------------------
// Small fronts: one CTA per supernode; the nrow x ns panel lives in shared memory (column-major, leading dimension
// ldp = nrow rounded up to odd). Extend-add is a GATHER: for every child the inverse of its relative-index list
// (parent front row -> child update row, -1 if absent) is built in shared memory, and every entry of the parent front
// sums its children's contributions itself (fixed child order) -- no read-modify-write chains, no atomics, all loads
// independent. Phases: panel <- Q values already scattered into HBM (cp.async) + children; per 64-column block: potrf,
// rows below solved (thread per row, values in registers), later columns updated; panel -> HBM; the update matrix
// U = -L21 L21^T + children goes straight from registers to HBM.
// ------------------------------------------------------------------------------------------------
struct FrontTask {
    int super;
    int smem_doubles;   // doubles of dynamic shared memory beyond the potrf tile + scratch (panel + inverse maps)
};

constexpr int FRONT_MAXC = 4;       // children whose inverse row maps are resident at once (more: several passes)

// panel stride: the smallest value >= nrow that is 8 mod 16 (the DMMA fragment loads of the update matrix then touch every bank twice)
__host__ __device__ __forceinline__ int front_ldp(int nrow) { return ((nrow + 7) & ~15) + 8; }
__host__ __device__ __forceinline__ int front_smem_doubles(int nrow, int ns) {
    return ns * front_ldp(nrow) + (FRONT_MAXC * nrow + 1) / 2;
}

// (two CTAs per SM: the second one hides the barriers of the small panel steps)
template <bool BLK>
__global__ void __launch_bounds__(256, 2)
front_small_kernel(const FrontTask *__restrict__ tasks, const SuperMeta *__restrict__ meta, const int *__restrict__ child_idx,
                   const int *__restrict__ relidx, double *__restrict__ Lx0, double *__restrict__ upd0, int *__restrict__ fail_col,
                   long long bstride) {
    extern __shared__ __align__(16) double fsm[];
    __shared__ const double *cU[FRONT_MAXC];
    __shared__ int cUld[FRONT_MAXC];
    double *sD = fsm;                       // 64 x 64 tile for the diagonal block
    double *scr = fsm + FTILE;
    double *sP = fsm + FTILE + FSCRATCH;    // panel
    double *__restrict__ Lx = lane_ptr_pinned(Lx0, bstride);
    double *__restrict__ upd = lane_ptr_pinned(upd0, bstride);
    fail_col = lane_ptr(fail_col, bstride);
    pdl_launch_dependents();
    const SuperMeta S = meta[tasks[blockIdx.x].super];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ns = S.ns, nrow = S.nrow, nr = nrow - ns;
    const int ldp = front_ldp(nrow);
    int *sInv = reinterpret_cast<int *>(sP + ns * ldp);     // [FRONT_MAXC][nrow]
    const int nch = S.child_end - S.child_begin;
    double *Lp = Lx + S.panel_off;
    pdl_wait();
    // ---- panel <- HBM (asynchronous; the strictly upper part of the diagonal block reads as zero) ----
    for (int c = warp; c < ns; c += 8) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(sP + c * ldp);
        const double *src = Lp + (long long)c * S.ld;
        for (int r = lane; r < nrow; r += 32) cp_async8s(sa + r * 8, src + r, r >= c ? 8 : 0);
    }
    cp_async_commit();
    // inverse maps of children [cb, cb + cn)
    auto build_maps = [&](int cb, int cn) {
        for (int e = tid; e < cn * nrow; e += 256) sInv[e] = -1;
        __syncthreads();
        for (int k = 0; k < cn; k++) {
            const SuperMeta C = meta[child_idx[S.child_begin + cb + k]];
            const int cnr = C.nrow - C.ns;
            const int *rel = relidx + C.rowptr + C.ns;
            for (int i = tid; i < cnr; i += 256) sInv[k * nrow + rel[i]] = i;
            if (tid == 0) { cU[k] = upd + C.upd_off; cUld[k] = C.uld; }
        }
        __syncthreads();
    };
    // ---- children: contributions to the panel ----
    for (int cb = 0; cb < nch; cb += FRONT_MAXC) {
        const int cn = min(FRONT_MAXC, nch - cb);
        build_maps(cb, cn);
        if (cb == 0) { cp_async_wait<0>(); __syncthreads(); }
        for (int c = warp; c < ns; c += 8) {
            double *col = sP + c * ldp;
            const double *Uck[FRONT_MAXC];
            bool has[FRONT_MAXC];
#pragma unroll
            for (int k = 0; k < FRONT_MAXC; k++) {
                const int ic = k < cn ? sInv[k * nrow + c] : -1;        // (uniform across the warp)
                has[k] = ic >= 0;
                Uck[k] = has[k] ? cU[k] + (long long)ic * cUld[k] : nullptr;
            }
            // 8 rows per lane and all children at once: up to 32 independent loads in flight per thread
            for (int r0 = c; r0 < nrow; r0 += 256) {
                double v[8][FRONT_MAXC];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + lane + 32 * u;
#pragma unroll
                    for (int k = 0; k < FRONT_MAXC; k++) {
                        const int ir = (has[k] && r < nrow) ? sInv[k * nrow + r] : -1;
                        v[u][k] = ir >= 0 ? Uck[k][ir] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + lane + 32 * u;
                    double add = 0.0;
#pragma unroll
                    for (int k = 0; k < FRONT_MAXC; k++) add += v[u][k];
                    if (r < nrow) col[r] += add;
                }
            }
        }
        __syncthreads();
    }
    if (nch == 0) { cp_async_wait<0>(); __syncthreads(); }
    // ---- blocked right-looking factorization of the panel ----
    for (int k0 = 0; k0 < ns; k0 += FB) {
        const int nb = min(FB, ns - k0), k1 = k0 + nb;
        for (int e = tid; e < FB * FB; e += 256) {
            const int i = e & 63, j = e >> 6;
            double v = (i == j) ? 1.0 : 0.0;
            if (i < nb && j <= i) v = sP[(k0 + j) * ldp + k0 + i];
            sD[j * FLD + i] = v;
        }
        __syncthreads();
        // the diagonal block together with the first 128 rows below it, then the remaining rows 128 at a time
        {
            const RowTile rt[2] = {{sP + k0 * ldp + k1, ldp, min(FB, nrow - k1)}, {sP + k0 * ldp + k1 + FB, ldp, max(0, min(FB, nrow - k1 - FB))}};
            if constexpr (BLK) panel_factor64b<2>(sD, scr, rt, nb, S.first + k0, fail_col, true);
            else panel_solve64<2, true>(sD, scr, rt, nb, S.first + k0, fail_col, true);
        }
        for (int e = tid; e < nb * nb; e += 256) {
            const int i = e % nb, j = e / nb;
            if (i >= j) sP[(k0 + j) * ldp + k0 + i] = sD[j * FLD + i];
        }
        for (int r0 = k1 + 2 * FB; r0 < nrow; r0 += 2 * FB) {
            const RowTile rt[2] = {{sP + k0 * ldp + r0, ldp, min(FB, nrow - r0)}, {sP + k0 * ldp + r0 + FB, ldp, max(0, min(FB, nrow - r0 - FB))}};
            panel_solve64<2, false>(sD, scr, rt, nb, 0, nullptr, false);
        }
        __syncthreads();
        // later columns of the panel: A[r, j] -= sum_c X[r, c] X[j, c]  (j in [k1, ns), r >= j), 4 x 4 register tiles
        if (k1 < ns) {
            const int nj = ns - k1, ni = nrow - k1;
            const int tj = (nj + 3) >> 2, ti = (ni + 3) >> 2;
            for (int tt = tid; tt < tj * ti; tt += 256) {
                const int bj = tt / ti, bi = tt - bj * ti;
                if (4 * bi + 3 < 4 * bj) continue;          // tile strictly above the diagonal
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
                for (int c = 0; c < nb; c++) {
                    const double *col = sP + (k0 + c) * ldp + k1;
                    double a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        a[i] = (4 * bi + i < ni) ? col[4 * bi + i] : 0.0;
                        b[i] = (4 * bj + i < nj) ? col[4 * bj + i] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int rr = 4 * bi + i, jj = 4 * bj + j;
                        if (rr < ni && jj < nj && rr >= jj) sP[(k1 + jj) * ldp + k1 + rr] -= acc[i][j];
                    }
            }
            __syncthreads();
        }
    }
    // ---- panel -> HBM ----
    for (int c = warp; c < ns; c += 8)
        for (int r = c + lane; r < nrow; r += 32) Lp[r + (long long)c * S.ld] = sP[c * ldp + r];
    if (nr == 0) return;
    // ---- update matrix U = -L21 L21^T + children (lower), registers -> HBM ----
    double *Up = upd + S.upd_off;
    const int tn = (nr + 3) >> 2;
    for (int cb = 0; cb == 0 || cb < nch; cb += FRONT_MAXC) {
        const int cn = max(0, min(FRONT_MAXC, nch - cb));
        if (nch > FRONT_MAXC) build_maps(cb, cn);            // (otherwise the maps of the panel phase are still in place)
        // FP64 tensor cores: the warps take 32 x 16 tiles of the lower triangle round robin (column-major over the tile
        // grid, so the warps of a round share the column operand); the children's contributions are gathered straight
        // into the accumulator fragments and every entry is stored once
        const int tr = (nr + 31) >> 5, tc = (nr + 15) >> 4;
        const int grp = lane >> 2, tig = lane & 3;
        for (int t = warp; t < tr * tc; t += 8) {
            const int bj = t / tr, bi = t - bj * tr;
            const int r0 = 32 * bi, c0 = 16 * bj;
            if (c0 >= r0 + 32) continue;                                   // tile strictly above the diagonal
            double acc[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            if (cb == 0) {
                warp_dmma_32x16(sP + ns, ldp, r0, nr, sP + ns, ldp, c0, nr, ns, acc);
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int rr = r0 + 8 * i + grp, jj = c0 + 8 * j + 2 * tig + e;
                            acc[i][j][e] = (rr < nr && jj < nr && rr >= jj) ? Up[rr + (long long)jj * S.uld] : 0.0;
                        }
            }
            for (int k = 0; k < cn; k++) {
                const int *inv = sInv + k * nrow + ns;
                const double *Uc = cU[k];
                const long long uld = cUld[k];
                int ii[4], ij[2][2];
#pragma unroll
                for (int i = 0; i < 4; i++) ii[i] = (r0 + 8 * i + grp < nr) ? inv[r0 + 8 * i + grp] : -1;
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) ij[j][e] = (c0 + 8 * j + 2 * tig + e < nr) ? inv[c0 + 8 * j + 2 * tig + e] : -1;
                double g[4][2][2];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++)
                            g[i][j][e] = (ii[i] >= 0 && ij[j][e] >= 0 && ii[i] >= ij[j][e]) ? Uc[ii[i] + ij[j][e] * uld] : 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) { acc[i][j][0] += g[i][j][0]; acc[i][j][1] += g[i][j][1]; }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int rr = r0 + 8 * i + grp, jj = c0 + 8 * j + 2 * tig + e;
                        if (rr < nr && jj < nr && rr >= jj) Up[rr + (long long)jj * S.uld] = acc[i][j][e];
                    }
        }
        if (nch > FRONT_MAXC) __syncthreads();
    }
}

}  // namespace gmrf
