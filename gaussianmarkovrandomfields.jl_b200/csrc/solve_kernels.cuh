// solve_kernels.cuh -- level-scheduled supernodal triangular solves for a few right-hand sides (nrhs <= 8).
//
// The solves are HBM-bound: the factor panels are streamed exactly once per direction, everything else (the
// right-hand sides, the 64x64 inverted diagonal blocks) lives in L2. Design points, all driven by the first B200
// measurements (profiles/r01_perf_1m_first.log: 242 GB/s with the one-thread-per-row / one-warp-per-column kernels):
//   * ONE launch per 64-column block step of a supernode chain, no cross-CTA reductions:
//       forward  (L y = b)  is right-looking over row tiles   : b[rows below] -= L[rows, K_j] x_j
//       backward (L' x = y) is right-looking over column tiles: t[cols left]  -= L[K_j, cols]' x_j
//     so every output element has one owner CTA and a fixed summation order (bit-reproducible).
//   * the small triangular solves with the diagonal blocks are replaced by products with the explicitly inverted
//     64x64 blocks that the factorization already produced (potrf_inv_kernel, kernels.cuh); the CTA that owns
//     the tile holding the NEXT block of the chain applies that inverse right after its update ("look-ahead"), so
//     the block solve never needs a launch of its own.
//   * a CTA owns a 64 x 64 tile (32 KB of L): 8 warps x 8 columns, every thread issues its 16 independent 8-byte
//     loads before the first use -> ~32 KB in flight per CTA, several CTAs per SM.
#pragma once
#include "kernels.cuh"

namespace gmrf {

constexpr int SOLVE_NB = 64;
constexpr int STEP_SMEM_BYTES = SOLVE_NB * SOLVE_NB * 8;     // dynamic shared memory of the block-step kernels (staged inverse)


struct FwdStepTask {      // block column K_j = [k0, k1) of a supernode, x_j final in `x`
    const double *L;      // first row below the diagonal block: panel + k0*ld + k1
    const double *inv_next;   // inverse of the next diagonal block (nb_next x nb_next) or nullptr
    const double *x;      // x_j (nb entries per right-hand side, stride ldy)
    double *y;            // rows k1.. of the supernode's own columns (ms of them), stride ldy
    double *u;            // update vector of the supernode (rows beyond ns), stride ldu
    int ld, nb, ms, m;    // m = rows below the block (ms own-column rows + nr update rows)
    int nb_next, pad_;
};

struct BwdGatherTask {    // t_S = y_S - L21' x_R for one supernode, then x of the LAST block of its chain
    const double *L21;    // panel + ns (nr x ns, leading dimension ld)
    const int *idx;       // global (permuted) rows of R (nr entries)
    const double *inv_last;   // inverse of the last diagonal block (nb_last x nb_last)
    double *y;            // own columns of the supernode (ns entries), stride ldy
    int ld, ns, nr, nb_last;
    int tile0, pad_;      // first 64-column tile this task launches (nr == 0: only the last one)
    double *part;         // nullptr: the task covers all nr rows and finishes t_S itself. Otherwise the task is ONE ROW
                          // CHUNK of a tall L21 (L21 / idx / nr describe the chunk) and only stores its partial sums
                          // part[c + q*ns]; bwd_reduce_kernel folds the chunks in fixed order afterwards.
};

constexpr int BWD_PART_Q = 8;         // right-hand-side planes per chunk in the partial buffer

struct BwdReduceTask {    // t_S = y_S - sum_chunks part, then the last block solve (one CTA per 64 own columns)
    const double *part;   // chunk k, plane q at part + (k * BWD_PART_Q + q) * ns
    const double *inv_last;
    double *y;
    int ns, nchunks, nb_last, pad_;
};

struct BwdStepTask {      // block row K_j = [k0, k1) of L11, x_j final in `x`
    const double *L;      // panel + k0 (row k0, column 0), leading dimension ld
    const double *inv_prev;   // inverse of diagonal block j-1 (64 x 64)
    const double *x;      // x_j (nb entries), stride ldy
    double *y;            // own columns 0.. of the supernode, stride ldy
    int ld, nb, ncols, pad_;  // ncols = k0 (a multiple of 64)
};

// Fixed-tree reduction of 8 per-lane values over the 32 lanes of a warp with 9 shuffles (instead of 40):
// three transposing rounds halve the number of live values while folding lane bits 4, 3, 2, two plain rounds fold
// bits 1, 0. On return every lane holds the full sum of value index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1).
__device__ __forceinline__ double warp_reduce8(double (&v)[8], int lane) {
    const unsigned full = 0xffffffffu;
    double a[4];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double send = hi ? v[i] : v[i + 4];
            const double keep = hi ? v[i + 4] : v[i];
            a[i] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    double b[2];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double send = hi ? a[i] : a[i + 2];
            const double keep = hi ? a[i + 2] : a[i];
            b[i] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    double c;
    {
        const bool hi = lane & 4;
        const double send = hi ? b[0] : b[1];
        const double keep = hi ? b[1] : b[0];
        c = keep + __shfl_xor_sync(full, send, 4);
    }
    c += __shfl_xor_sync(full, c, 2);
    c += __shfl_xor_sync(full, c, 1);
    return c;
}
__device__ __forceinline__ int warp_reduce8_index(int lane) {
    return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

// The inverted diagonal block is fetched into registers at the START of the kernel by the one CTA that will need it
// (16 independent loads per thread, overlapped with the tile loads), and applied from registers afterwards.
//   lower   : x[r, q] = sum_c inv[r + c*nb] * b[c][q]    thread (r = tid & 63, part = tid >> 6) owns 16 columns
//   lower^T : x[c, q] = sum_r inv[r + c*nb] * t[r][q]    warp owns 8 columns, lanes own rows lane, lane + 32
__device__ __forceinline__ void load_inv_lower(double (&g)[16], const double *__restrict__ inv, int nb, int tid) {
    const int r = tid & 63, part = tid >> 6;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int c = part * 16 + i;
        g[i] = (r < nb && c <= r) ? inv[r + (long long)c * nb] : 0.0;
    }
}
template <int RB>
__device__ __forceinline__ void apply_inv_lower(const double (&g)[16], int nb, const double (*sb)[RB], double (*sp)[SOLVE_NB][RB],
                                                double *__restrict__ out, long long ldo, int nrhs, int tid) {
    const int r = tid & 63, part = tid >> 6;
#pragma unroll
    for (int q = 0; q < RB; q++) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            s0 += g[i] * sb[part * 16 + i][q];
            s1 += g[i + 1] * sb[part * 16 + i + 1][q];
        }
        sp[part][r][q] = s0 + s1;
    }
    __syncthreads();
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int rr = e % SOLVE_NB, q = e / SOLVE_NB;
        if (rr < nb && q < nrhs) out[rr + q * ldo] = (sp[0][rr][q] + sp[1][rr][q]) + (sp[2][rr][q] + sp[3][rr][q]);
    }
}

__device__ __forceinline__ void load_inv_lower_t(double (&g)[16], const double *__restrict__ inv, int nb, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = warp * 8 + i;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lane + 32 * h;
            g[2 * i + h] = (c < nb && r < nb && r >= c) ? inv[r + (long long)c * nb] : 0.0;
        }
    }
}
template <int RB>
__device__ __forceinline__ void apply_inv_lower_t(const double (&g)[16], int nb, const double (*st)[RB],
                                                  double *__restrict__ out, long long ldo, int nrhs, int warp, int lane) {
#pragma unroll
    for (int q = 0; q < RB; q++) {
        if (q >= nrhs) break;
        double p[8];
        const double t0 = lane < nb ? st[lane][q] : 0.0, t1 = lane + 32 < nb ? st[lane + 32][q] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = g[2 * i] * t0 + g[2 * i + 1] * t1;
        const double s = warp_reduce8(p, lane);
        const int c = warp * 8 + warp_reduce8_index(lane);
        if ((lane & 3) == 0 && c < nb) out[c + q * ldo] = s;
    }
}

// The block-step kernels keep the inverted diagonal block of their ONE look-ahead CTA in shared memory instead (cp.async,
// no registers): with the 16 doubles per thread above, every CTA of the launch paid 32 registers for a block only one of
// them applies, which capped the step kernels at 3 CTAs per SM (ncu, profiles/r02_ncu_full_fwd_step_3d48.txt: 33 % of the
// warp slots, long-scoreboard stalls of 8-10 per issue, 1.2-2.4 TB/s per launch -- not enough loads in flight).
__device__ __forceinline__ void stage_inv_block(double *__restrict__ sinv, const double *__restrict__ inv, int nb, int tid) {
    for (int e = tid; e < nb * nb; e += 256) cp_async8(sinv + e, inv + e, 8);      // (8-byte granules: block bases may be odd)
    cp_async_commit();
}
template <int RB>
__device__ __forceinline__ void apply_inv_lower_smem(const double *__restrict__ sinv, int nb, const double (*sb)[RB],
                                                     double (*sp)[SOLVE_NB][RB], double *__restrict__ out, long long ldo, int nrhs, int tid) {
    const int r = tid & 63, part = tid >> 6;
    double g[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int c = part * 16 + i;
        g[i] = (r < nb && c <= r) ? sinv[r + c * nb] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < RB; q++) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            s0 += g[i] * sb[part * 16 + i][q];
            s1 += g[i + 1] * sb[part * 16 + i + 1][q];
        }
        sp[part][r][q] = s0 + s1;
    }
    __syncthreads();
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int rr = e % SOLVE_NB, q = e / SOLVE_NB;
        if (rr < nb && q < nrhs) out[rr + q * ldo] = (sp[0][rr][q] + sp[1][rr][q]) + (sp[2][rr][q] + sp[3][rr][q]);
    }
}
template <int RB>
__device__ __forceinline__ void apply_inv_lower_t_smem(const double *__restrict__ sinv, int nb, const double (*st)[RB],
                                                       double *__restrict__ out, long long ldo, int nrhs, int warp, int lane) {
    double g[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = warp * 8 + i;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lane + 32 * h;
            g[2 * i + h] = (c < nb && r < nb && r >= c) ? sinv[r + c * nb] : 0.0;
        }
    }
#pragma unroll
    for (int q = 0; q < RB; q++) {
        if (q >= nrhs) break;
        double p[8];
        const double t0 = lane < nb ? st[lane][q] : 0.0, t1 = lane + 32 < nb ? st[lane + 32][q] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = g[2 * i] * t0 + g[2 * i + 1] * t1;
        const double s = warp_reduce8(p, lane);
        const int c = warp * 8 + warp_reduce8_index(lane);
        if ((lane & 3) == 0 && c < nb) out[c + q * ldo] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward assembly of one level: supernode s pulls its children's update vectors (fixed order) into its own
// rows of y and into u_s, then the first block of its chain is solved: x_0 = inv(L_00) b_0.
// One CTA per (supernode, right-hand side).
// ------------------------------------------------------------------------------------------------
constexpr int FWD_ASM_ROWS = 4096;      // front rows per CTA: the top supernodes (30,000 rows) are split over several CTAs

// First position of the sorted list rel[0..n) whose value is >= x, found by the whole CTA: 256 evenly spaced probes narrow
// the window by a factor of 256 per round (two rounds for 65,000 entries). Uniform result; contains barriers.
__device__ __forceinline__ int cta_lower_bound(const int *__restrict__ rel, int n, int x) {
    int a = 0, b = n;                       // invariant: rel[i] < x for i < a, rel[i] >= x for i >= b
    while (b - a > 0) {
        const int stride = (b - a + 255) / 256;
        const int i = a + (int)threadIdx.x * stride;
        const int below = __syncthreads_count(i < b && rel[i] < x);       // probes a, a + stride, ...: a prefix of them is < x
        if (below == 0) { b = a; break; }
        const int na = a + (below - 1) * stride + 1;                      // the last probe < x
        b = min(b, a + below * stride);                                   // the first probe >= x (or the old end)
        a = na;
    }
    return b;
}

__global__ void __launch_bounds__(256)
fwd_assemble_x0_kernel(const int *__restrict__ supers, const SuperMeta *__restrict__ meta,
                       const int *__restrict__ child_idx, const int *__restrict__ relidx,
                       const double *__restrict__ Linv, const long long *__restrict__ inv_base,
                       double *__restrict__ y, long long ldy, double *__restrict__ uvec, long long ldu) {
    __shared__ double sb[SOLVE_NB][1];
    __shared__ double sp[4][SOLVE_NB][1];
    pdl_launch_dependents();
    const int s = supers[blockIdx.x];
    // a supernode with more than FWD_ASM_ROWS front rows appears once per row chunk in the list: this CTA's chunk is its
    // position among the repeats. Every row has ONE owner CTA, which adds the children in their fixed order (the first
    // version ran the root's 2 x 30,000 entries through one CTA: 233 us of serialized read-modify-writes per solve).
    int first = blockIdx.x;
    while (first > 0 && supers[first - 1] == s) first--;
    const int chunk = blockIdx.x - first;
    const SuperMeta P = meta[s];
    const int nr = P.nrow - P.ns;
    const int r = blockIdx.y;
    const int nb0 = min(P.ns, SOLVE_NB);
    const int row_lo = chunk * FWD_ASM_ROWS, row_hi = min(P.nrow, row_lo + FWD_ASM_ROWS);
    const bool solve0 = chunk == 0 && P.wide == 0;       // long chains: the first 256-column block is solved by fwd_wide_diag_kernel
    double g[16];
    if (solve0) load_inv_lower(g, Linv + inv_base[s], nb0, threadIdx.x);
    pdl_wait();
    double *us = uvec + P.uvec_off + (long long)r * ldu;
    for (int i = max(row_lo, P.ns) - P.ns + threadIdx.x; i < row_hi - P.ns && i < nr; i += 256) us[i] = 0.0;
    __syncthreads();
    double *ys = y + P.first + (long long)r * ldy;
    for (int ci = P.child_begin; ci < P.child_end; ci++) {
        const SuperMeta C = meta[child_idx[ci]];
        const int cnr = C.nrow - C.ns;
        const int *rel = relidx + C.rowptr + C.ns;
        const double *uc = uvec + C.uvec_off + (long long)r * ldu;
        // the child's rows that land in this chunk are a contiguous run of its (sorted) list
        int run_lo = 0, run_hi = cnr;
        if (P.nrow > FWD_ASM_ROWS) {
            run_lo = cta_lower_bound(rel, cnr, row_lo);
            run_hi = row_hi >= P.nrow ? cnr : cta_lower_bound(rel, cnr, row_hi);
        }
        // four entries per thread and pass, loads before stores: the rows of one child are distinct, so the four
        // read-modify-writes are independent (written as a plain loop they serialize on possible aliasing)
        for (int i0 = run_lo + threadIdx.x; i0 < run_hi; i0 += 4 * 256) {
            double *dst[4];
            double v[4], old[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + 256 * u;
                const int p = i < run_hi ? rel[i] : -1;
                ok[u] = p >= row_lo && p < row_hi;
                dst[u] = p < P.ns ? ys + p : us + (p - P.ns);
                v[u] = ok[u] ? uc[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) old[u] = ok[u] ? *dst[u] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (ok[u]) *dst[u] = old[u] + v[u];
        }
        __syncthreads();
    }
    if (!solve0) return;
    if (threadIdx.x < SOLVE_NB) sb[threadIdx.x][0] = threadIdx.x < nb0 ? ys[threadIdx.x] : 0.0;
    __syncthreads();
    apply_inv_lower<1>(g, nb0, sb, sp, ys, 0, 1, threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Forward block step: rows below block column K_j:  b[r] -= sum_k L[r, k] x_j[k]; the tile that holds the next
// diagonal block then solves it (x_{j+1} = inv(L_{j+1,j+1}) b_{j+1}) and stores x in place of b.
// ------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(256, RB == 1 ? 4 : 2)      // (one right-hand side: 4 CTAs per SM = at most 64 registers)
fwd_step_kernel(const FwdStepTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs,
                long long ldy, long long ldu) {
    __shared__ double xs[SOLVE_NB][RB];
    __shared__ double part[8][SOLVE_NB][RB];
    __shared__ double sb[SOLVE_NB][RB];
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const FwdStepTask T = tasks[t];
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = tile * SOLVE_NB;
    const bool head = (tile == 0 && T.nb_next > 0);
    extern __shared__ __align__(16) double step_sinv[];      // [64 x 64] inverse of the next diagonal block (look-ahead CTA only)
    if (head) stage_inv_block(step_sinv, T.inv_next, T.nb_next, tid);
    // issue the tile loads first: 8 columns x 2 rows per thread
    double l[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int kk = warp * 8 + i;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = row0 + lane + 32 * h;
            l[i][h] = (kk < T.nb && r < T.m) ? T.L[r + (long long)kk * T.ld] : 0.0;
        }
    }
    pdl_wait();        // (everything above reads only the factor and the task tables)
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int kk = e % SOLVE_NB, q = e / SOLVE_NB;
        xs[kk][q] = (kk < T.nb && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
    }
    // the entries this thread will update at the end: fetched now, so that their L2 round trip overlaps the products
    constexpr int NOUT = (SOLVE_NB * RB + 255) / 256;
    double yold[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        const int e = tid + 256 * k, rr = e % SOLVE_NB, q = e / SOLVE_NB, r = row0 + rr;
        yold[k] = (e < SOLVE_NB * RB && r < T.m && q < nrhs) ? ((r < T.ms) ? T.y[r + q * ldy] : T.u[(r - T.ms) + q * ldu]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RB; q++) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double xv = xs[warp * 8 + i][q];
            a0 += l[i][0] * xv;
            a1 += l[i][1] * xv;
        }
        part[warp][lane][q] = a0;
        part[warp][lane + 32][q] = a1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        const int e = tid + 256 * k;
        if (e >= SOLVE_NB * RB) break;
        const int rr = e % SOLVE_NB, q = e / SOLVE_NB;
        const int r = row0 + rr;
        if (head && rr >= T.nb_next) sb[rr][q] = 0.0;
        if (r >= T.m || q >= nrhs) continue;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += part[w][rr][q];
        double *dst = (r < T.ms) ? (T.y + r + q * ldy) : (T.u + (r - T.ms) + q * ldu);
        const double v = yold[k] - s;
        if (head && rr < T.nb_next) sb[rr][q] = v; else *dst = v;
    }
    if (!head) return;
    cp_async_wait<0>();
    __syncthreads();
    apply_inv_lower_smem<RB>(step_sinv, T.nb_next, sb, (double (*)[SOLVE_NB][RB])part, T.y, ldy, nrhs, tid);
}

// ------------------------------------------------------------------------------------------------
// Backward, first phase of a level: t_S = y_S - L21' x_R with x_R gathered through the row structure; the tile
// that holds the last block of the chain then solves it: x_last = inv(L_last)' t_last.
// One CTA per 64 own columns, one warp per 8 columns, lanes stride the nr rows (coalesced column reads).
// ------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(256)
bwd_gather_kernel(const BwdGatherTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs,
                  const double *__restrict__ xg, long long ldy) {
    __shared__ double st[SOLVE_NB][RB];
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const BwdGatherTask T = tasks[t];
    const int tile = T.tile0 + (blockIdx.x - tile_prefix[t]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = tile * SOLVE_NB + warp * 8;
    const int nblk = (T.ns + SOLVE_NB - 1) / SOLVE_NB;
    const bool tail = (tile == nblk - 1) && !T.pad_;      // pad_ = 1: long chain, bwd_wide_diag_kernel solves the last block
    double g[16];
    if (tail) load_inv_lower_t(g, T.inv_last, T.nb_last, warp, lane);
    pdl_wait();
    double acc[8][RB];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < RB; q++) acc[i][q] = 0.0;
    if (c0 < T.ns) {
        // Rows are walked in blocks of 32*HB; inside a block the warp visits its 8 columns four at a time, so every
        // column visit reads 256*HB contiguous bytes with 4*HB independent loads per lane in flight (ncu on the first
        // version: 256-byte visits of 8 interleaved column streams per warp held DRAM at 45 % of the copy rate).
        constexpr int HB = (RB >= 4) ? 2 : 4;
        for (int r0 = 0; r0 < T.nr; r0 += 32 * HB) {
            double xv[HB][RB];
            bool ok[HB];
#pragma unroll
            for (int h = 0; h < HB; h++) {
                const int r = r0 + lane + 32 * h;
                ok[h] = r < T.nr;
                const long long gr = ok[h] ? T.idx[r] : 0;
#pragma unroll
                for (int q = 0; q < RB; q++) xv[h][q] = (ok[h] && q < nrhs) ? xg[gr + q * ldy] : 0.0;
            }
#pragma unroll
            for (int half = 0; half < 2; half++) {
                double lv[4][HB];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int c = c0 + 4 * half + i;
                    const double *col = T.L21 + (long long)c * T.ld + r0 + lane;
#pragma unroll
                    for (int h = 0; h < HB; h++) lv[i][h] = (ok[h] && c < T.ns) ? col[32 * h] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int h = 0; h < HB; h++)
#pragma unroll
                        for (int q = 0; q < RB; q++) acc[4 * half + i][q] += lv[i][h] * xv[h][q];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RB; q++) {
        double p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = acc[i][q];
        const double s = warp_reduce8(p, lane);
        const int cl = warp * 8 + warp_reduce8_index(lane);      // column within the tile
        const int c = tile * SOLVE_NB + cl;
        if ((lane & 3) == 0 && c < T.ns && q < nrhs) {
            if (T.part) { T.part[c + (long long)q * T.ns] = s; continue; }
            const double v = T.y[c + q * ldy] - s;
            if (tail) st[cl][q] = v; else T.y[c + q * ldy] = v;
        }
    }
    if (!tail || T.part) return;
    __syncthreads();
    apply_inv_lower_t<RB>(g, T.nb_last, st, T.y + (long long)tile * SOLVE_NB, ldy, nrhs, warp, lane);
}

// Second pass for the supernodes whose L21 was processed in row chunks: fold the partial sums in chunk order.
template <int RB>
__global__ void __launch_bounds__(256)
bwd_reduce_kernel(const BwdReduceTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs,
                  long long ldy) {
    __shared__ double st[SOLVE_NB][RB];
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const BwdReduceTask T = tasks[t];
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (T.ns + SOLVE_NB - 1) / SOLVE_NB;
    const bool tail = (tile == nblk - 1) && !T.pad_;
    double g[16];
    if (tail) load_inv_lower_t(g, T.inv_last, T.nb_last, warp, lane);
    pdl_wait();
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int cl = e % SOLVE_NB, q = e / SOLVE_NB;
        const int c = tile * SOLVE_NB + cl;
        if (c >= T.ns || q >= nrhs) continue;
        double s = 0.0;
        for (int k = 0; k < T.nchunks; k++) s += T.part[((long long)k * BWD_PART_Q + q) * T.ns + c];
        const double v = T.y[c + q * ldy] - s;
        if (tail) st[cl][q] = v; else T.y[c + q * ldy] = v;
    }
    if (!tail) return;
    __syncthreads();
    apply_inv_lower_t<RB>(g, T.nb_last, st, T.y + (long long)tile * SOLVE_NB, ldy, nrhs, warp, lane);
}

// ------------------------------------------------------------------------------------------------
// Backward block step: columns left of block row K_j:  t[c] -= sum_{r in K_j} L[r, c] x_j[r]; the tile that
// holds block j-1 then solves it: x_{j-1} = inv(L_{j-1,j-1})' t_{j-1}.
// ------------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(256, RB == 1 ? 4 : 2)
bwd_step_kernel(const BwdStepTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs,
                long long ldy) {
    __shared__ double xs[SOLVE_NB][RB];
    __shared__ double st[SOLVE_NB][RB];
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const BwdStepTask T = tasks[t];
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = tile * SOLVE_NB + warp * 8;
    const int ntiles = T.ncols / SOLVE_NB;
    const bool tail = (tile == ntiles - 1);
    extern __shared__ __align__(16) double step_sinv[];      // [64 x 64] inverse of the previous diagonal block (look-ahead CTA only)
    if (tail) stage_inv_block(step_sinv, T.inv_prev, SOLVE_NB, tid);
    double l[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lane + 32 * h;
            l[i][h] = (r < T.nb) ? T.L[r + (long long)(c0 + i) * T.ld] : 0.0;
        }
    pdl_wait();        // (everything above reads only the factor and the task tables)
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int kk = e % SOLVE_NB, q = e / SOLVE_NB;
        xs[kk][q] = (kk < T.nb && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
    }
    // the entries this thread will update at the end: fetched now, so that their L2 round trip overlaps the products
    const int cl = warp * 8 + warp_reduce8_index(lane);
    double yold[RB];
#pragma unroll
    for (int q = 0; q < RB; q++) yold[q] = ((lane & 3) == 0 && q < nrhs) ? T.y[(long long)tile * SOLVE_NB + cl + q * ldy] : 0.0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RB; q++) {
        if (q >= nrhs) break;
        const double x0 = xs[lane][q], x1 = xs[lane + 32][q];
        double p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = l[i][0] * x0 + l[i][1] * x1;
        const double s = warp_reduce8(p, lane);
        if ((lane & 3) == 0) {
            double *dst = T.y + (long long)tile * SOLVE_NB + cl + q * ldy;
            const double v = yold[q] - s;
            if (tail) st[cl][q] = v; else *dst = v;
        }
    }
    if (!tail) return;
    cp_async_wait<0>();
    __syncthreads();
    apply_inv_lower_t_smem<RB>(step_sinv, SOLVE_NB, st, T.y + (long long)tile * SOLVE_NB, ldy, nrhs, warp, lane);
}

// ------------------------------------------------------------------------------------------------
// Wide steps for long supernode chains (ns > 256): the sweeps above advance 64 columns per launch, so the top separators
// of a 3D problem (30,000 columns) cost ~480 dependent launches per direction whose ~10 us of latency each -- not the
// 15 MB they stream -- bound the 1 M-dof solve (18.7 ms against 10 ms of pure streaming). Here a step is 256 columns and two
// launches: (A) every 64-row (forward) / 64-column (backward) tile applies the whole 256-column block of already final
// unknowns, 64 independent loads per thread in flight; (B) ONE CTA per supernode solves the next 256 x 256 diagonal block
// (four inverse-block products with the updates in between; 260 KB, L2-resident right after the factorization's writes).
// A quarter of the dependent launches, each moving four times the bytes.
// ------------------------------------------------------------------------------------------------
constexpr int SOLVE_WB = 256;          // columns per wide step
constexpr int SOLVE_WIDE_MIN = 256;    // supernodes with more own columns than this take the wide steps

struct WideStepTask {     // forward: rows below block column K = [k0, k0 + nbw): dst[r] -= L[r, K] x_K
    const double *L;      // forward: panel + k0*ld + k1 (k1 = k0 + nbw);  backward: panel + k0 (row k0, column 0)
    const double *x;      // x_K (nbw entries per right-hand side, stride ldy)
    double *y;            // forward: own rows k1.. of the supernode;  backward: own columns 0.. of the supernode
    double *u;            // forward: update vector of the supernode (rows beyond ns)
    const double *inv_head;   // look-ahead (hrows > 0): inverted 64 x 64 blocks of the NEXT diagonal block of the chain
    int ld, nbw;          // nbw <= 256
    int ms, m;            // forward: own rows below / all rows below;  backward: m = columns left of K (a multiple of 64 tiles)
    int hrows, pad_;      // look-ahead: tile 0 of the task is a HEAD CTA that updates the hrows unknowns of the next diagonal
                          // block (forward: the first hrows rows below K, backward: the 256 columns left of K) and solves that
                          // block right away, so the chain needs one launch per 256 columns; the other tiles cover the rest
};

struct WideDiagTask {     // the 256 x 256 diagonal block D = L[K, K] with its (up to four) inverted 64 x 64 blocks
    const double *D;      // panel + k0*ld + k0
    const double *inv;    // inverse of the first 64-column block of K (the others follow at stride 64*64)
    double *y;            // the nbw unknowns of K (stride ldy)
    int ld, nbw;
};

// x_K := L[K, K]^-1 y_K on the unknowns held in shared memory, block by block: x_b = inv_b y_b, then y_c -= L[c, b] x_b for
// the later blocks c of K
template <int RB>
__device__ __forceinline__ void fwd_diag_solve_smem(double (*ys)[RB], double (*sp)[SOLVE_NB][RB], const double *__restrict__ D, int ld,
                                                    const double *__restrict__ inv, int nbw, int tid) {
    const int nblk = (nbw + SOLVE_NB - 1) / SOLVE_NB;
    for (int b = 0; b < nblk; b++) {
        const int c0 = b * SOLVE_NB, nb = min(SOLVE_NB, nbw - c0);
        // the rows of the later blocks against this block's columns: issued now, used after the block solve
        const int r = c0 + SOLVE_NB + tid;                     // one thread per later row (<= 192 of them)
        double lrow[SOLVE_NB];
        const bool has = r < nbw;
#pragma unroll
        for (int c = 0; c < SOLVE_NB; c++) lrow[c] = (has && c < nb) ? D[r + (long long)(c0 + c) * ld] : 0.0;
        double g[16];
        load_inv_lower(g, inv + (long long)b * SOLVE_NB * SOLVE_NB, nb, tid);
        {
            const int rr = tid & 63, prt = tid >> 6;
#pragma unroll
            for (int q = 0; q < RB; q++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    s0 += g[i] * ys[c0 + prt * 16 + i][q];
                    s1 += g[i + 1] * ys[c0 + prt * 16 + i + 1][q];
                }
                sp[prt][rr][q] = s0 + s1;
            }
        }
        __syncthreads();
        for (int e = tid; e < SOLVE_NB * RB; e += 256) {
            const int rr = e % SOLVE_NB, q = e / SOLVE_NB;
            if (rr < nb) ys[c0 + rr][q] = (sp[0][rr][q] + sp[1][rr][q]) + (sp[2][rr][q] + sp[3][rr][q]);
        }
        __syncthreads();
        if (has) {
#pragma unroll
            for (int q = 0; q < RB; q++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < SOLVE_NB; c += 2) { s0 += lrow[c] * ys[c0 + c][q]; s1 += lrow[c + 1] * ys[c0 + c + 1][q]; }
                ys[r][q] -= s0 + s1;
            }
        }
        __syncthreads();
    }
}

// x_K := L[K, K]^-T t_K on the unknowns held in shared memory, blocks in reverse: x_b = inv_b^T t_b, then
// t_c -= L[b, c]^T x_b for the earlier blocks c of K
template <int RB>
__device__ __forceinline__ void bwd_diag_solve_smem(double (*ys)[RB], double (*st)[RB], const double *__restrict__ D, int ld,
                                                    const double *__restrict__ inv, int nbw, int nrhs, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int nblk = (nbw + SOLVE_NB - 1) / SOLVE_NB;
    for (int b = nblk - 1; b >= 0; b--) {
        const int r0 = b * SOLVE_NB, nb = min(SOLVE_NB, nbw - r0);
        // x_b = inv_b^T t_b: warp owns 8 columns of the inverse, lanes own rows lane, lane + 32
        double g[16];
        load_inv_lower_t(g, inv + (long long)b * SOLVE_NB * SOLVE_NB, nb, warp, lane);
        for (int e = tid; e < SOLVE_NB * RB; e += 256) st[e % SOLVE_NB][e / SOLVE_NB] = ys[r0 + e % SOLVE_NB][e / SOLVE_NB];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < RB; q++) {
            if (q >= nrhs) break;
            double p[8];
            const double t0 = lane < nb ? st[lane][q] : 0.0, t1 = lane + 32 < nb ? st[lane + 32][q] : 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = g[2 * i] * t0 + g[2 * i + 1] * t1;
            const double sum = warp_reduce8(p, lane);
            const int c = warp * 8 + warp_reduce8_index(lane);
            if ((lane & 3) == 0 && c < nb) ys[r0 + c][q] = sum;
        }
        __syncthreads();
        if (b > 0) {
            // t_c -= L[b-rows, c]^T x_b for the r0 earlier columns (r0 is a multiple of 64): 64 columns per pass, a warp
            // takes 8 of them with two rows per lane -- 16 loads in flight, then the fixed 8-way shuffle tree
            for (int cb = 0; cb < r0; cb += SOLVE_NB) {
                double l[8][2];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const double *col = D + (long long)(cb + warp * 8 + i) * ld + r0;
                    l[i][0] = lane < nb ? col[lane] : 0.0;
                    l[i][1] = lane + 32 < nb ? col[lane + 32] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < RB; q++) {
                    if (q >= nrhs) break;
                    const double x0 = lane < nb ? ys[r0 + lane][q] : 0.0, x1 = lane + 32 < nb ? ys[r0 + lane + 32][q] : 0.0;
                    double p[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) p[i] = l[i][0] * x0 + l[i][1] * x1;
                    const double sum = warp_reduce8(p, lane);
                    const int c = cb + warp * 8 + warp_reduce8_index(lane);
                    if ((lane & 3) == 0) ys[c][q] -= sum;
                }
            }
            __syncthreads();
        }
    }
}

template <int RB>
__global__ void __launch_bounds__(256)
fwd_wide_step_kernel(const WideStepTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs,
                     long long ldy, long long ldu) {
    __shared__ double xs[SOLVE_WB][RB];
    __shared__ double part[8][SOLVE_NB][RB];
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const WideStepTask T = tasks[t];
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int head = T.hrows > 0 ? 1 : 0;
    if (head && tile == 0) {
        // HEAD: the hrows unknowns of the next diagonal block. One thread per row streams the row's nbw entries (64 loads
        // in flight per thread), then the block is solved in shared memory and published: the next launch finds x_{K+1}.
        for (int e = tid; e < SOLVE_WB * RB; e += 256) {
            const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
            xs[kk][q] = (kk < T.nbw && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
        }
        __syncthreads();
        double acc0[RB], acc1[RB];
#pragma unroll
        for (int q = 0; q < RB; q++) acc0[q] = acc1[q] = 0.0;
        const bool mine = tid < T.hrows;
#pragma unroll 1
        for (int b = 0; b < SOLVE_WB; b += SOLVE_NB) {
            if (b >= T.nbw) break;
            double lr[SOLVE_NB];
#pragma unroll
            for (int i = 0; i < SOLVE_NB; i++) lr[i] = (mine && b + i < T.nbw) ? T.L[tid + (long long)(b + i) * T.ld] : 0.0;
#pragma unroll
            for (int q = 0; q < RB; q++)
#pragma unroll
                for (int i = 0; i < SOLVE_NB; i += 2) {
                    acc0[q] += lr[i] * xs[b + i][q];
                    acc1[q] += lr[i + 1] * xs[b + i + 1][q];
                }
        }
        double (*ys)[RB] = reinterpret_cast<double (*)[RB]>(&part[0][0][0]);                       // [256][RB]
        double (*sp)[SOLVE_NB][RB] = reinterpret_cast<double (*)[SOLVE_NB][RB]>(&part[4][0][0]);   // [4][64][RB]
#pragma unroll
        for (int q = 0; q < RB; q++) ys[tid][q] = (mine && q < nrhs) ? T.y[tid + q * ldy] - (acc0[q] + acc1[q]) : 0.0;
        __syncthreads();
        fwd_diag_solve_smem<RB>(ys, sp, T.L + (long long)T.nbw * T.ld, T.ld, T.inv_head, T.hrows, tid);
#pragma unroll
        for (int q = 0; q < RB; q++)
            if (mine && q < nrhs) T.y[tid + q * ldy] = ys[tid][q];
        return;
    }
    const int row0 = T.hrows + (tile - head) * SOLVE_NB;
    // all 64 loads of the thread's 2 rows x 32 columns are issued before anything waits on them
    double l[4][8][2];
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int kk = b * SOLVE_NB + warp * 8 + i;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = row0 + lane + 32 * h;
                l[b][i][h] = (kk < T.nbw && r < T.m) ? T.L[r + (long long)kk * T.ld] : 0.0;
            }
        }
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        xs[kk][q] = (kk < T.nbw && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RB; q++) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double xv = xs[b * SOLVE_NB + warp * 8 + i][q];
                a0 += l[b][i][0] * xv;
                a1 += l[b][i][1] * xv;
            }
        part[warp][lane][q] = a0;
        part[warp][lane + 32][q] = a1;
    }
    __syncthreads();
    for (int e = tid; e < SOLVE_NB * RB; e += 256) {
        const int rr = e % SOLVE_NB, q = e / SOLVE_NB;
        const int r = row0 + rr;
        if (r >= T.m || q >= nrhs) continue;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += part[w][rr][q];
        double *dst = (r < T.ms) ? (T.y + r + q * ldy) : (T.u + (r - T.ms) + q * ldu);
        *dst -= s;
    }
}

template <int RB>
__global__ void __launch_bounds__(256)
fwd_wide_diag_kernel(const WideDiagTask *__restrict__ tasks, int nrhs, long long ldy) {
    __shared__ double ys[SOLVE_WB][RB];
    __shared__ double sp[4][SOLVE_NB][RB];
    const WideDiagTask T = tasks[blockIdx.x];
    const int tid = threadIdx.x;
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        ys[kk][q] = (kk < T.nbw && q < nrhs) ? T.y[kk + q * ldy] : 0.0;
    }
    __syncthreads();
    fwd_diag_solve_smem<RB>(ys, sp, T.D, T.ld, T.inv, T.nbw, tid);
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        if (kk < T.nbw && q < nrhs) T.y[kk + q * ldy] = ys[kk][q];
    }
}

// backward: columns left of block row K = [k0, k0 + nbw): t[c] -= sum_{r in K} L[r, c] x_K[r]; a CTA owns 64 columns
template <int RB>
__global__ void __launch_bounds__(256)
bwd_wide_step_kernel(const WideStepTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, int nrhs, long long ldy) {
    __shared__ double xs[SOLVE_WB][RB];
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const WideStepTask T = tasks[t];
    const int tile = blockIdx.x - tile_prefix[t];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int head = T.hrows > 0 ? 1 : 0;
    if (head && tile == 0) {
        // HEAD: the 256 unknowns of the previous diagonal block = columns [m - 256, m). Four passes of 64 columns (the
        // ordinary tile code), then the transposed block solve in shared memory: the next launch finds x_{K-1}.
        __shared__ double ys[SOLVE_WB][RB];
        __shared__ double st[SOLVE_NB][RB];
        for (int e = tid; e < SOLVE_WB * RB; e += 256) {
            const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
            xs[kk][q] = (kk < T.nbw && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
            ys[kk][q] = 0.0;
        }
        __syncthreads();
        const int cbase = T.m - SOLVE_WB;
#pragma unroll 1
        for (int cb = 0; cb < SOLVE_WB; cb += SOLVE_NB) {
            const int cc = cbase + cb + warp * 8;
            double lh[8][8];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int h = 0; h < 8; h++) {
                    const int r = lane + 32 * h;
                    lh[i][h] = r < T.nbw ? T.L[r + (long long)(cc + i) * T.ld] : 0.0;
                }
#pragma unroll
            for (int q = 0; q < RB; q++) {
                if (q >= nrhs) break;
                double xv[8];
#pragma unroll
                for (int h = 0; h < 8; h++) xv[h] = xs[lane + 32 * h][q];
                double p[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int h = 0; h < 8; h += 2) { s0 += lh[i][h] * xv[h]; s1 += lh[i][h + 1] * xv[h + 1]; }
                    p[i] = s0 + s1;
                }
                const double sum = warp_reduce8(p, lane);
                const int cl = cb + warp * 8 + warp_reduce8_index(lane);
                if ((lane & 3) == 0) ys[cl][q] = T.y[cbase + cl + q * ldy] - sum;
            }
        }
        __syncthreads();
        bwd_diag_solve_smem<RB>(ys, st, T.L - SOLVE_WB + (long long)cbase * T.ld, T.ld, T.inv_head, SOLVE_WB, nrhs, tid);
        for (int e = tid; e < SOLVE_WB * RB; e += 256) {
            const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
            if (q < nrhs) T.y[cbase + kk + q * ldy] = ys[kk][q];
        }
        return;
    }
    const int mcols = T.m - (head ? SOLVE_WB : 0);             // columns the ordinary tiles cover
    const int c0 = (tile - head) * SOLVE_NB + warp * 8;
    double l[8][8];                                            // 8 columns x 8 rows (lane, lane + 32, ...)
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int h = 0; h < 8; h++) {
            const int r = lane + 32 * h;
            l[i][h] = (r < T.nbw && c0 + i < mcols) ? T.L[r + (long long)(c0 + i) * T.ld] : 0.0;
        }
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        xs[kk][q] = (kk < T.nbw && q < nrhs) ? T.x[kk + q * ldy] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RB; q++) {
        if (q >= nrhs) break;
        double xv[8];
#pragma unroll
        for (int h = 0; h < 8; h++) xv[h] = xs[lane + 32 * h][q];
        double p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int h = 0; h < 8; h += 2) { s0 += l[i][h] * xv[h]; s1 += l[i][h + 1] * xv[h + 1]; }
            p[i] = s0 + s1;
        }
        const double sum = warp_reduce8(p, lane);
        const int cl = warp * 8 + warp_reduce8_index(lane);
        const int c = (tile - head) * SOLVE_NB + cl;
        if ((lane & 3) == 0 && c < mcols) T.y[c + q * ldy] -= sum;
    }
}

template <int RB>
__global__ void __launch_bounds__(256)
bwd_wide_diag_kernel(const WideDiagTask *__restrict__ tasks, int nrhs, long long ldy) {
    __shared__ double ys[SOLVE_WB][RB];
    __shared__ double st[SOLVE_NB][RB];
    const WideDiagTask T = tasks[blockIdx.x];
    const int tid = threadIdx.x;
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        ys[kk][q] = (kk < T.nbw && q < nrhs) ? T.y[kk + q * ldy] : 0.0;
    }
    __syncthreads();
    bwd_diag_solve_smem<RB>(ys, st, T.D, T.ld, T.inv, T.nbw, nrhs, tid);
    for (int e = tid; e < SOLVE_WB * RB; e += 256) {
        const int kk = e % SOLVE_WB, q = e / SOLVE_WB;
        if (kk < T.nbw && q < nrhs) T.y[kk + q * ldy] = ys[kk][q];
    }
}

// ------------------------------------------------------------------------------------------------
// Wide right-hand-side blocks (sampling, constraints, many columns): the sweeps are DMMA GEMMs on 64-column blocks of
// right-hand sides (plans built in gmrf_b200.cu); only the two irregular steps need kernels of their own.
// ------------------------------------------------------------------------------------------------
constexpr int MULTI_NW = 3;    // block widths of the wide path: 64, 128, 256 right-hand sides per pass
constexpr int MULTI_WMAX = 256;
constexpr int MULTI_QB = 8;    // columns per CTA in the two kernels below

// Forward assembly for a block of right-hand sides: u_s := 0, then the children's update blocks are added into the
// supernode's own rows of y and into u_s (fixed child order). One CTA per (supernode, 8 columns).
__global__ void __launch_bounds__(256)
fwd_assemble_multi_kernel(const int *__restrict__ supers, const SuperMeta *__restrict__ meta,
                          const int *__restrict__ child_idx, const int *__restrict__ relidx,
                          double *__restrict__ y, long long ldy, double *__restrict__ uvec, long long ldu) {
    const SuperMeta P = meta[supers[blockIdx.x]];
    const int nr = P.nrow - P.ns;
    const int q0 = blockIdx.y * MULTI_QB;
    double *us = uvec + P.uvec_off + (long long)q0 * ldu;
    for (int e = threadIdx.x; e < nr * MULTI_QB; e += 256) {
        const int i = e % nr, q = e / nr;
        us[i + q * ldu] = 0.0;
    }
    __syncthreads();
    double *ys = y + P.first + (long long)q0 * ldy;
    for (int ci = P.child_begin; ci < P.child_end; ci++) {
        const SuperMeta C = meta[child_idx[ci]];
        const int cnr = C.nrow - C.ns;
        const int *rel = relidx + C.rowptr + C.ns;
        const double *uc = uvec + C.uvec_off + (long long)q0 * ldu;
        // the rows of one child are distinct and so are the 8 columns: all loads of a pass are issued before its first
        // store (as a plain `+=` loop the 8 read-modify-writes of an entry serialize on possible aliasing)
        for (int i = threadIdx.x; i < cnr; i += 256) {
            const int p = rel[i];
            double *dst = (p < P.ns) ? (ys + p) : (us + (p - P.ns));
            const long long ldd = (p < P.ns) ? ldy : ldu;
            double v[MULTI_QB], old[MULTI_QB];
#pragma unroll
            for (int q = 0; q < MULTI_QB; q++) v[q] = uc[i + q * ldu];
#pragma unroll
            for (int q = 0; q < MULTI_QB; q++) old[q] = dst[q * ldd];
#pragma unroll
            for (int q = 0; q < MULTI_QB; q++) dst[q * ldd] = old[q] + v[q];
        }
        __syncthreads();
    }
}

// Backward gather for a block of right-hand sides: u_s[r, q] = y[rowidx_s[ns + r], q] (the solved rows of the
// ancestors), so that t_S = y_S - L21' u_s is a plain GEMM. One CTA per (supernode row tile of 256, 8 columns).
struct RowGatherTask { const int *idx; double *u; int nr, pad_; };
__global__ void __launch_bounds__(256)
rows_gather_kernel(const RowGatherTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks,
                   const double *__restrict__ y, long long ldy, long long ldu) {
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const RowGatherTask T = tasks[t];
    const int r = (blockIdx.x - tile_prefix[t]) * 256 + threadIdx.x;
    if (r >= T.nr) return;
    const long long g = T.idx[r];
    const int q0 = blockIdx.y * MULTI_QB;
#pragma unroll
    for (int q = 0; q < MULTI_QB; q++) T.u[r + (q0 + q) * ldu] = y[g + (q0 + q) * ldy];
}

}  // namespace gmrf
