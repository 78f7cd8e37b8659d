// kernels.cuh -- hand-written sm_100a device code for the supernodal multifrontal Cholesky, the
// level-scheduled triangular solves and Takahashi selected inversion. All arithmetic is FP64.
//
// B200 facts that shape these kernels (profiles/r01_fp64_probe.json): the FP64 tensor pipe (DMMA.8x8x4,
// reached through mma.sync.m8n8k4.f64 -- there is no FP64 tcgen05/UMMA) peaks at 37.2 TFLOP/s, the same as
// the DFMA pipe, but needs 1/8 of the issue slots and 1/4 of the operand traffic, so every GEMM-shaped
// contraction goes through DMMA with cp.async-staged shared-memory tiles; everything else (assembly,
// scatter, skinny panel products of the solves) is HBM/L2-bound and written for coalesced streaming.
//
// Determinism: no floating-point atomics; every output element has exactly one owner thread and all
// reductions use a fixed tree, so reruns are bit-identical.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace gmrf {

// ------------------------------------------------------------------------------------------------
// Task descriptors (built on the host at analysis time, resident in HBM)
// ------------------------------------------------------------------------------------------------
enum : int {
    GEMM_LOWER = 1,      // write only entries with local row >= local col
    GEMM_BETA0 = 2,      // C = alpha*op(A)op(B) (+I) instead of C += ...
    GEMM_ALPHA_POS = 4,  // alpha = +1 (default alpha = -1)
    GEMM_ADD_I = 8,      // add the identity on the local diagonal
    GEMM_GATHER = 16     // (GATHER instantiation, C = update matrix of supernode pad_ - 1) C = children's contributions - A B^T,
                         // written once: the extend-add of the update matrix happens in this epilogue
    ,GEMM_A_CONST = 32   // operand A is not written by any kernel of the same graph (the factor during a triangular sweep): a
                         // programmatically launched instance loads its first A tiles before waiting for the predecessor
};

struct GemmTask {        // C[m x n] (+)= alpha * Aop[m x k] * Bop[n x k]^T
    const double *A;     // TA=false: element (i,kk) at A[i + kk*lda]   (m-contiguous)
    const double *B;     // TB=false: element (j,kk) at B[j + kk*ldb]   (n-contiguous)
    double *C;           // TA/TB=true: element at X[kk + i*ldx]        (k-contiguous)
    int m, n, k;
    int lda, ldb, ldc;
    int flags;
    int pad_;            // GEMM_GATHER: parent supernode + 1
};

struct PanelTask {       // diagonal block step: D := chol(D) (nb x nb, in place, lower) and inv := D^-1
    double *D;           // diagonal block, leading dimension ld
    double *inv;         // nb x nb, leading dimension nb, zeros above the diagonal (reused by the solve phase)
    int ld, nb;
    int col0;            // global (permuted) column of the block's first column, for pivot reporting
    int pad_;
};

struct AsmItem {         // one column tile of one parent front
    int super;
    int col0;            // first front column of the tile (0 .. nrow)
};

struct SuperMeta {
    long long panel_off;   // doubles
    long long upd_off;     // doubles (update pool / selinv W pool share the layout fields below)
    long long zw_off;
    long long rowptr;      // offset into rowidx / relidx
    long long uvec_off;    // rows
    int first, ns, nrow, ld, uld, parent;
    int child_begin, child_end;   // range in child_idx
    int wide, pad_;               // wide = 1: a long chain whose triangular solves advance 256 columns per step
};

// Tables the gathering epilogue of the update-matrix products reads (one copy per handle, in HBM).
struct GatherCtx {
    const SuperMeta *meta;
    const int *child_idx;
    const int *relidx;
    const int *relpos;            // per child: position in its relative-index list of every 256-row boundary of the parent
    const long long *relpos_off;
    double *upd;                  // update pool (lane 0)
};
constexpr int GATHER_MAXC = 4;    // children whose inverse maps are resident at once

// ------------------------------------------------------------------------------------------------
// Cholesky of a 4 x 4 tile held in the registers of one thread (lower triangle of a[r][c], r >= c), the innermost serial
// piece of every panel factorization: its latency is paid once per 4 columns on the critical path of the 2D configs.
// Column by column it is four dependent rsqrt chains (~480 cycles each on sm_100a: MUFU seed + Newton steps + scale +
// update). Here the tile is factored as two 2 x 2 blocks whose second pivot comes from the block's determinant,
//     l00 = sqrt(a00),  l10 = a10 / l00,  l11 = sqrt(a00 a11 - a10^2) / sqrt(a00),
// so both square roots of a block start together and the chain is two rsqrt latencies instead of four. The conditioning is
// that of the ordinary recurrence (a11 - l10^2 = (a00 a11 - a10^2) / a00: the same cancellation). piv[c] gets a number with
// the sign of the c-th pivot (for the not-positive-definite report), ri[c] = 1 / l_cc.
// 1 / sqrt(x) as straight-line code: the hardware seed (rsqrt.approx.ftz.f64 -> MUFU.RSQ64H, ~20 good bits) and one
// third-order correction y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2 (error after the step ~ e^3: below the rounding of the
// five operations, max 1 ulp-ish measured against numpy in tests/test_gpu_kernels.py). The library rsqrt() carries a
// slow-path branch for subnormal / special arguments, which keeps the compiler from interleaving two of them; pivots of a
// factorization that is going to be used are normal positive numbers, and anything else (<= 0, NaN, inf) still comes out
// as NaN / inf / 0 and is caught by the sign test on the pivot itself.
__device__ __forceinline__ double rsqrt_inline(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}

__device__ __forceinline__ void chol4x4_lower(double (&a)[4][4], double (&ri)[4], double (&piv)[4]) {
    const double a00 = a[0][0], a10 = a[1][0];
    const double det1 = fma(a00, a[1][1], -(a10 * a10));
    const double r0 = rsqrt_inline(a00), q1 = rsqrt_inline(det1);
    const double l00 = a00 * r0, l10 = a10 * r0;
    const double ri1 = l00 * q1;
    const double l20 = a[2][0] * r0, l30 = a[3][0] * r0;
    const double l21 = fma(-l20, l10, a[2][1]) * ri1, l31 = fma(-l30, l10, a[3][1]) * ri1;
    const double s22 = fma(-l21, l21, fma(-l20, l20, a[2][2]));
    const double s32 = fma(-l31, l21, fma(-l30, l20, a[3][2]));
    const double s33 = fma(-l31, l31, fma(-l30, l30, a[3][3]));
    const double det2 = fma(s22, s33, -(s32 * s32));
    const double r2 = rsqrt_inline(s22), q3 = rsqrt_inline(det2);
    const double l22 = s22 * r2;
    piv[0] = a00; piv[1] = det1; piv[2] = s22; piv[3] = det2;
    ri[0] = r0; ri[1] = ri1; ri[2] = r2; ri[3] = l22 * q3;
    a[0][0] = l00; a[1][0] = l10; a[1][1] = det1 * q1 * r0;
    a[2][0] = l20; a[2][1] = l21; a[3][0] = l30; a[3][1] = l31;
    a[2][2] = l22; a[3][2] = s32 * r2; a[3][3] = det2 * q3 * r2;
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol). A kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream still runs: it reads its task tables (and whatever else no kernel
// of the same graph writes -- for the triangular sweeps that is the whole factor) and then waits for the predecessor's
// completion before touching anything the predecessor produces. The few-RHS sweeps are ~900 dependent launches per
// direction whose factor tiles do NOT depend on the previous step, only the 64 unknowns x_j do: their HBM latency
// overlaps the previous step. Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Batched factorization: a handle may hold several numeric "lanes" (independent value sets on the same pattern, e.g.
// the points of a hyperparameter sweep). All numeric arrays of a lane live in one arena; lane b sits `bstride` bytes
// after lane 0, the task tables point into lane 0, and blockIdx.y selects the lane. One launch then advances every
// lane by the same step, which is what fills the GPU on small (2D) problems where a single factorization cannot.
template <class T>
__device__ __forceinline__ T *lane_ptr(T *p, long long bstride) {
    return reinterpret_cast<T *>(reinterpret_cast<char *>(const_cast<typename std::remove_const<T>::type *>(p)) + (long long)blockIdx.y * bstride);
}
// Same, but opaque to the optimiser: for kernels whose inner loops select between lane-offset base pointers. Without
// it the offset was rematerialised (special-register read, 64-bit multiply, a branch) inside the extend-add loops,
// which then ran 91 ms instead of 55 ms per 1 M-dof refactorization. (The GEMM kernel is faster with the plain form.)
template <class T>
__device__ __forceinline__ T *lane_ptr_pinned(T *p, long long bstride) {
    char *q = reinterpret_cast<char *>(const_cast<typename std::remove_const<T>::type *>(p)) + (long long)blockIdx.y * bstride;
    asm volatile("" : "+l"(q));
    return reinterpret_cast<T *>(q);
}

__device__ __forceinline__ int find_task(const int *__restrict__ prefix, int ntasks, int bid) {
    int lo = 0, hi = ntasks;  // prefix has ntasks+1 entries; find t with prefix[t] <= bid < prefix[t+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= bid) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16s(unsigned saddr, const void *gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8s(unsigned saddr, const void *gmem, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(saddr), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core GEMM:  C (+)= alpha * Aop * Bop^T  over a list of tasks (one launch per schedule step).
// CTA tile BM x BN, warp grid WGM x WGN, K tile 16, 3-stage cp.async pipeline.
// smem tiles are stored k-major: As[stage][kk][m] (+4 padding -> conflict-free 8-byte fragment loads).
// ------------------------------------------------------------------------------------------------
// Shared-memory tile of one operand: k-major [KT][BX + 4] for an operand that is x-contiguous in global memory,
// x-major [BX][KT + 4] for a k-contiguous one (so that its cp.async stores run along the contiguous direction);
// both paddings make the 8-byte DMMA fragment loads bank-conflict free.
template <int BX, int KT, bool TR>
__host__ __device__ constexpr int gemm_tile_doubles() { return TR ? BX * (KT + 4) : KT * (BX + 4); }
template <int BM, int BN, int KT = 16, int ST = 3, bool TA = false, bool TB = false>
constexpr int gemm_smem_bytes() { return ST * (gemm_tile_doubles<BM, KT, TA>() + gemm_tile_doubles<BN, KT, TB>()) * (int)sizeof(double); }

// Loader of one operand: [KT x BX] tiles, k-major in shared memory (LDS = BX + 4). TR=false: global is x-contiguous
// (16-byte cp.async when pointer and leading dimension allow it, 8-byte otherwise); TR=true: global is k-contiguous.
// Everything that does not change along k (this thread's x position, its validity, the base pointer, the shared-memory
// offset) is computed once per tile; a k-step costs one pointer bump and one bound compare per cp.async. (The first
// version recomputed the index arithmetic per element: ncu showed 4 integer/branch instructions per DMMA.)
template <int BX, int NT, bool TR, int KT>
struct TileLoader {
    static constexpr int LDS = BX + 4;                // !TR: k-major rows
    static constexpr int LDK = KT + 4;                // TR: x-major rows
    static constexpr int IT_V = (BX / 2) * KT / NT;   // 16-byte chunks per thread per tile (same count for both layouts)
    static constexpr int IT_S = BX * KT / NT;         // 8-byte elements per thread per tile
    const double *p;       // source of this thread's first element at k = 0
    int ld;
    int kk;                // k offset (within a tile) of the first element
    int soff;              // shared-memory offset (doubles) of the first element
    int xbytes;            // !TR: valid bytes at this thread's x (0, 8, 16)
    unsigned xmask;        // TR: bit i set <=> x of iteration i is in range
    bool vec;

    __device__ __forceinline__ void init(const double *g, int ld_, int x0, int xmax, bool aligned16, int tid) {
        ld = ld_;
        vec = aligned16;
        xmask = 0;
        xbytes = 0;
        if (!TR) {
            if (vec) {
                const int x = (tid % (BX / 2)) * 2;
                kk = tid / (BX / 2);
                const int gx = x0 + x;
                xbytes = gx < xmax ? ((xmax - gx >= 2) ? 16 : 8) : 0;
                p = g + (xbytes ? gx : 0) + (long long)kk * ld;
                soff = kk * LDS + x;
            } else {
                const int x = tid % BX;
                kk = tid / BX;
                const int gx = x0 + x;
                xbytes = gx < xmax ? 8 : 0;
                p = g + (xbytes ? gx : 0) + (long long)kk * ld;
                soff = kk * LDS + x;
            }
        } else {
            // chunks run along k: a row of the tile is KT contiguous doubles in global AND in shared memory
            const int per_row = vec ? KT / 2 : KT;         // chunks per x row
            const int xb = tid / per_row;
            kk = (tid % per_row) * (vec ? 2 : 1);
            const int its = vec ? IT_V : IT_S, xs = NT / per_row;
            for (int i = 0; i < its; i++)
                if (x0 + xb + i * xs < xmax) xmask |= 1u << i;
            p = g + kk + (long long)(x0 + xb) * ld;
            soff = xb * LDK + kk;
        }
    }
    // Issue the copies of the tile that starts at k0 into the stage at shared address `sbase` (rows >= kmax are
    // zero-filled: src-size 0 reads nothing, so the source pointer may run past the operand). Pointers advance
    // additively; interior tiles (the steady state) skip the k-bound tests altogether.
    __device__ __forceinline__ void load(unsigned sbase, int k0, int kmax) const {
        const unsigned sa = sbase + (unsigned)soff * 8u;
        const bool interior = k0 + KT <= kmax;
        if (!TR) {
            const char *src = reinterpret_cast<const char *>(p + (long long)k0 * ld);
            if (vec) {
                constexpr int KS = NT / (BX / 2);
                const long long stride = (long long)KS * ld * 8;
                if (interior) {
#pragma unroll
                    for (int i = 0; i < IT_V; i++) cp_async16s(sa + i * KS * LDS * 8, src + i * stride, xbytes);
                } else {
                    const int kleft = kmax - k0 - kk;
#pragma unroll
                    for (int i = 0; i < IT_V; i++) cp_async16s(sa + i * KS * LDS * 8, src + i * stride, (i * KS < kleft) ? xbytes : 0);
                }
            } else {
                constexpr int KS = NT / BX;
                const long long stride = (long long)KS * ld * 8;
                if (interior) {
#pragma unroll
                    for (int i = 0; i < IT_S; i++) cp_async8s(sa + i * KS * LDS * 8, src + i * stride, xbytes);
                } else {
                    const int kleft = kmax - k0 - kk;
#pragma unroll
                    for (int i = 0; i < IT_S; i++) cp_async8s(sa + i * KS * LDS * 8, src + i * stride, (i * KS < kleft) ? xbytes : 0);
                }
            }
        } else {
            const char *src = reinterpret_cast<const char *>(p + k0);
            const int kleft = kmax - k0 - kk;               // valid elements from this thread's k position on
            if (vec) {
                constexpr int XS = NT / (KT / 2);
                const long long stride = (long long)XS * ld * 8;
                const int kb = kleft >= 2 ? 16 : (kleft == 1 ? 8 : 0);
#pragma unroll
                for (int i = 0; i < IT_V; i++)
                    cp_async16s(sa + i * XS * LDK * 8, src + i * stride, ((xmask >> i) & 1u) ? kb : 0);
            } else {
                constexpr int XS = NT / KT;
                const long long stride = (long long)XS * ld * 8;
                const int kb = kleft >= 1 ? 8 : 0;
#pragma unroll
                for (int i = 0; i < IT_S; i++)
                    cp_async8s(sa + i * XS * LDK * 8, src + i * stride, ((xmask >> i) & 1u) ? kb : 0);
            }
        }
    }
};

// Tail of an update-matrix product of a supernode with children (GEMM_GATHER): the extend-add of the update part happens
// HERE. The tile holds -L21 L21^T (just written by this CTA, still in L2); every entry adds what the children contribute
// to it and is stored once more -- HBM sees the children's entries read once and the parent's written once (the scatter
// form wrote the parent's update matrix -- zeros included --, and this product read and rewrote it: 24 B per entry where
// 8 B suffice). Per child the rows / columns that land in the tile are found by scanning the (<= 512-entry) stretch of
// its relative-index list that the 256-row position table brackets; the inverse maps live in the pipeline's shared
// memory, free by now. A separate function on purpose: nothing of it may cost the main loop a register.
struct GatherChild {          // per child of the tile's supernode (shared memory)
    const double *U;
    const int *rel;
    long long uld;
    int rlo, rhi, clo, chi;   // stretches of the child's relative-index list bracketing the tile's rows / columns
};

template <int BM, int BN, int NT>
__device__ __noinline__ void gather_children_into_tile(const GemmTask &T, int m0, int n0, const GatherCtx *__restrict__ gctx, long long bstride,
                                                       int *__restrict__ smem_i) {
    int *invR = smem_i;                         // [GATHER_MAXC][BM]
    int *invC = smem_i + GATHER_MAXC * BM;      // [GATHER_MAXC][BN]
    GatherChild *ch = reinterpret_cast<GatherChild *>(smem_i + GATHER_MAXC * (BM + BN));
    const int tid = threadIdx.x;
    const GatherCtx G = *gctx;
    const SuperMeta P = G.meta[T.pad_ - 1];
    const double *upd = lane_ptr(G.upd, bstride);
    const int fr0 = P.ns + m0, fc0 = P.ns + n0;                  // first front row / column of the tile
    const int nch = P.child_end - P.child_begin;
    const int qmax = (P.nrow - 1) >> 8;
    for (int cb = 0; cb < nch; cb += GATHER_MAXC) {
        const int cn = min(GATHER_MAXC, nch - cb);
        if (tid < cn) {                          // one thread per child walks child -> meta -> position table
            const int cs = G.child_idx[P.child_begin + cb + tid];
            const SuperMeta C = G.meta[cs];
            const int *rp = G.relpos + G.relpos_off[cs];
            GatherChild c;
            c.U = upd + C.upd_off; c.rel = G.relidx + C.rowptr + C.ns; c.uld = C.uld;
            c.rlo = rp[fr0 >> 8]; c.rhi = rp[min((fr0 + BM - 1) >> 8, qmax) + 1];
            c.clo = rp[fc0 >> 8]; c.chi = rp[min((fc0 + BN - 1) >> 8, qmax) + 1];
            ch[tid] = c;
        }
        for (int e = tid; e < GATHER_MAXC * (BM + BN); e += NT) invR[e] = -1;
        __syncthreads();
        for (int k = 0; k < cn; k++) {
            const GatherChild c = ch[k];
            for (int i = c.rlo + tid; i < c.rhi; i += NT) {
                const int p = c.rel[i];
                if (p >= fr0 && p < fr0 + BM) invR[k * BM + p - fr0] = i;
            }
            for (int i = c.clo + tid; i < c.chi; i += NT) {
                const int p = c.rel[i];
                if (p >= fc0 && p < fc0 + BN) invC[k * BN + p - fc0] = i;
            }
        }
        __syncthreads();
        // rows across the threads (coalesced), eight columns per pass and all children at once: up to 32 loads in flight
        constexpr int UN = 8;
        for (int e0 = tid; e0 < BM * BN; e0 += UN * NT) {
            double v[UN][GATHER_MAXC];
            bool ok[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int e = e0 + u * NT, rr = e % BM, cc = e / BM;
                const int r = m0 + rr, c = n0 + cc;
                ok[u] = e < BM * BN && r < T.m && c < T.n && r >= c;
#pragma unroll
                for (int k = 0; k < GATHER_MAXC; k++) {
                    v[u][k] = 0.0;
                    if (k < cn && ok[u]) {
                        const int ir = invR[k * BM + rr], ic = invC[k * BN + cc];
                        if (ic >= 0 && ir >= ic) v[u][k] = ch[k].U[ir + ic * ch[k].uld];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int e = e0 + u * NT, rr = e % BM, cc = e / BM;
                const double add = (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]);
                if (ok[u] && add != 0.0) T.C[(m0 + rr) + (long long)(n0 + cc) * T.ldc] += add;
            }
        }
        __syncthreads();
    }
}

template <int BM, int BN, int WGM, int WGN, bool TA, bool TB, int GEMM_KT = 16, int GEMM_STAGES = 3, bool GATHER = false>
__global__ void __launch_bounds__(WGM *WGN * 32, (WGM * WGN == 4) ? ((BM * BN <= 64 * 64) ? 4 : 3) : 1)
gemm_dmma_kernel(const GemmTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, long long bstride,
                 const GatherCtx *__restrict__ gctx = nullptr) {
    constexpr int NT = WGM * WGN * 32;
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4, LDK = GEMM_KT + 4;
    constexpr int A_TILE = gemm_tile_doubles<BM, GEMM_KT, TA>(), B_TILE = gemm_tile_doubles<BN, GEMM_KT, TB>();
    constexpr int WM = BM / WGM, WN = BN / WGN;
    constexpr int MI = WM / 8, NI = WN / 8;
    extern __shared__ __align__(16) double gemm_smem[];
    double *As = gemm_smem;
    double *Bs = gemm_smem + GEMM_STAGES * A_TILE;

    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    GemmTask T = tasks[t];
    T.A = lane_ptr(T.A, bstride); T.B = lane_ptr(T.B, bstride); T.C = lane_ptr(T.C, bstride);
    const int local = blockIdx.x - tile_prefix[t];
    const int mt = (T.m + BM - 1) / BM, nt = (T.n + BN - 1) / BN;
    // L2-friendly rasterisation: consecutive CTAs sweep bands of GEMM_BAND row tiles column by column, so one wave of
    // resident CTAs covers a roughly square patch of C and every operand strip it streams is shared by many CTAs
    // (ncu, 8192^2 x 2048 with the plain column-major order: 15 GB of DRAM reads for 0.27 GB of operands).
    constexpr int GEMM_BAND = 16;
    const int band = local / (GEMM_BAND * nt);
    const int rem = local - band * GEMM_BAND * nt;
    const int bh = min(GEMM_BAND, mt - band * GEMM_BAND);
    const int tm = band * GEMM_BAND + rem % bh, tn = rem / bh;
    const int m0 = tm * BM, n0 = tn * BN;
    if ((T.flags & GEMM_LOWER) && n0 >= m0 + BM) return;  // tile strictly above the diagonal

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp % WGM) * WM, wn0 = (warp / WGM) * WN;
    const int grp = lane >> 2, tig = lane & 3;
    const bool a16 = ((((uintptr_t)T.A) & 15) == 0) && ((T.lda & 1) == 0);
    const bool b16 = ((((uintptr_t)T.B) & 15) == 0) && ((T.ldb & 1) == 0);

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = (T.k + GEMM_KT - 1) / GEMM_KT;
    const unsigned as_base = (unsigned)__cvta_generic_to_shared(As), bs_base = (unsigned)__cvta_generic_to_shared(Bs);
    TileLoader<BM, NT, TA, GEMM_KT> la;
    TileLoader<BN, NT, TB, GEMM_KT> lb;
    la.init(T.A, T.lda, m0, T.m, a16, tid);
    lb.init(T.B, T.ldb, n0, T.n, b16, tid);
    // A tiles of the first stages, then B tiles; the wait for a programmatic predecessor sits before the first operand it
    // may have written (all A tiles join the first commit group: they were issued back to back anyway)
    const bool a_ahead = (T.flags & GEMM_A_CONST) != 0;
    if (!a_ahead) pdl_wait();
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; s++)
        if (s < nk) la.load(as_base + s * A_TILE * 8, s * GEMM_KT, T.k);
    if (a_ahead) pdl_wait();
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; s++) {
        if (s < nk) lb.load(bs_base + s * B_TILE * 8, s * GEMM_KT, T.k);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + GEMM_STAGES - 1;
            if (nx < nk) {
                int st = nx % GEMM_STAGES;
                la.load(as_base + st * A_TILE * 8, nx * GEMM_KT, T.k);
                lb.load(bs_base + st * B_TILE * 8, nx * GEMM_KT, T.k);
            }
            cp_async_commit();
        }
        const double *as = As + (kt % GEMM_STAGES) * A_TILE;
        const double *bs = Bs + (kt % GEMM_STAGES) * B_TILE;
#pragma unroll
        for (int ks = 0; ks < GEMM_KT / 4; ks++) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; i++)
                a[i] = TA ? as[(wm0 + 8 * i + grp) * LDK + ks * 4 + tig] : as[(ks * 4 + tig) * LDA_S + wm0 + 8 * i + grp];
#pragma unroll
            for (int j = 0; j < NI; j++)
                b[j] = TB ? bs[(wn0 + 8 * j + grp) * LDK + ks * 4 + tig] : bs[(ks * 4 + tig) * LDB_S + wn0 + 8 * j + grp];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    const double alpha = (T.flags & GEMM_ALPHA_POS) ? 1.0 : -1.0;
    const bool gather = GATHER && (T.flags & GEMM_GATHER);
    const bool beta0 = (T.flags & GEMM_BETA0) || gather, lower = T.flags & GEMM_LOWER, addi = T.flags & GEMM_ADD_I;
    // Epilogue in chunks of 16 outputs: all reads of C are issued before the first store of the chunk, so a tile pays
    // one memory round trip per chunk instead of one per element (loads could not be hoisted over the stores otherwise;
    // with k = 64 the element-wise read-modify-write chain cost more than the tile's arithmetic).
    constexpr int TOT = MI * NI * 2, CH = (BM * BN > 64 * 64) ? 4 : 8;
    static_assert(TOT % CH == 0, "epilogue chunking");
    double *const Cb = T.C + (m0 + wm0 + grp) + (long long)(n0 + wn0 + 2 * tig) * T.ldc;
#pragma unroll
    for (int base = 0; base < TOT; base += CH) {
        double cv[CH];
#pragma unroll
        for (int u = 0; u < CH; u++) {
            const int idx = base + u, e = idx & 1, j = (idx >> 1) % NI, i = (idx >> 1) / NI;
            const int r = m0 + wm0 + 8 * i + grp, c = n0 + wn0 + 8 * j + 2 * tig + e;
            const bool ok = r < T.m && c < T.n && !(lower && r < c);
            cv[u] = (ok && !beta0) ? Cb[8 * i + (long long)(8 * j + e) * T.ldc] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < CH; u++) {
            const int idx = base + u, e = idx & 1, j = (idx >> 1) % NI, i = (idx >> 1) / NI;
            const int r = m0 + wm0 + 8 * i + grp, c = n0 + wn0 + 8 * j + 2 * tig + e;
            const bool ok = r < T.m && c < T.n && !(lower && r < c);
            if (ok) {
                double v = alpha * acc[i][j][e] + cv[u];
                if (addi && r == c) v += 1.0;
                Cb[8 * i + (long long)(8 * j + e) * T.ldc] = v;
            }
        }
    }
    if (gather) {
        __syncthreads();           // the tile is written (-A B^T) and the operand stages are free
        gather_children_into_tile<BM, BN, NT>(T, m0, n0, gctx, bstride, reinterpret_cast<int *>(gemm_smem));
    }
}

// Debug kernel with the same contract (one thread per output element); selected with the "naive_kernels" option.
template <bool TA, bool TB>
__global__ void gemm_naive_kernel(const GemmTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, long long bstride) {
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    GemmTask T = tasks[t];
    T.A = lane_ptr(T.A, bstride); T.B = lane_ptr(T.B, bstride); T.C = lane_ptr(T.C, bstride);
    const int local = blockIdx.x - tile_prefix[t];
    const int mt = (T.m + 15) / 16;
    const int r = (local % mt) * 16 + (threadIdx.x & 15), c = (local / mt) * 16 + (threadIdx.x >> 4);
    if (r >= T.m || c >= T.n) return;
    if ((T.flags & GEMM_LOWER) && r < c) return;
    double s = 0.0;
    for (int kk = 0; kk < T.k; kk++) {
        double a = TA ? T.A[kk + (long long)r * T.lda] : T.A[r + (long long)kk * T.lda];
        double b = TB ? T.B[kk + (long long)c * T.ldb] : T.B[c + (long long)kk * T.ldb];
        s += a * b;
    }
    double *p = T.C + r + (long long)c * T.ldc;
    double v = ((T.flags & GEMM_ALPHA_POS) ? 1.0 : -1.0) * s;
    if (!(T.flags & GEMM_BETA0)) v += *p;
    if ((T.flags & GEMM_ADD_I) && r == c) v += 1.0;
    *p = v;
}

constexpr int POTRF_NB = 64;

// ------------------------------------------------------------------------------------------------
// Diagonal-block step of the panel factorization (latency-critical: one launch per 64 columns of a chain):
// ONE CTA factors the nb x nb block in shared memory (right-looking, 256 threads on the trailing update) and then
// inverts the triangular factor (thread c builds column c of L^-1 by forward substitution, registers only).
// The rows below the block are then solved by a DMMA GEMM with the inverted block (X = B * L^-T, in place), and
// the same inverted blocks serve the triangular solves later (solve_kernels.cuh), so they cost nothing extra there.
// Non-positive pivots are recorded with an integer atomicMin (deterministic) and produce NaNs downstream,
// matching the reference's `check=false` factorization (backend.jl:184).
// ------------------------------------------------------------------------------------------------
template <int NBT>
__global__ void __launch_bounds__(256)
potrf_inv_kernel(const PanelTask *__restrict__ tasks, int *__restrict__ fail_col, long long bstride) {
    __shared__ double sA[NBT][NBT + 1];   // identity-padded beyond nb
    pdl_launch_dependents();
    PanelTask T = tasks[blockIdx.x];
    T.D = lane_ptr(T.D, bstride); T.inv = lane_ptr(T.inv, bstride); fail_col = lane_ptr(fail_col, bstride);
    const int nb = T.nb, tid = threadIdx.x;
    pdl_wait();
    for (int e = tid; e < NBT * NBT; e += 256) {
        const int i = e % NBT, j = e / NBT;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j <= i) v = T.D[i + (long long)j * T.ld];
        sA[i][j] = v;
    }
    __syncthreads();
    const int ti = tid % NBT, tk = tid / NBT;        // trailing update: thread owns row ti, columns tk, tk + KS, ...
    constexpr int KS = 256 / NBT;
    for (int j = 0; j < NBT; j++) {
        const double d = sA[j][j];
        if (tid == 0 && !(d > 0.0) && j < nb) atomicMin(fail_col, T.col0 + j + 1);
        const double r = rsqrt(d);
        __syncthreads();                              // everyone has read the pivot
        if (tid == j) sA[j][j] = d * r;
        else if (tid > j && tid < NBT) sA[tid][j] *= r;
        __syncthreads();
        const double lij = sA[ti][j];
        for (int k = j + 1 + tk; k <= ti; k += KS) sA[ti][k] -= lij * sA[k][j];
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += 256) {
        const int i = e % nb, j = e / nb;
        if (i >= j) T.D[i + (long long)j * T.ld] = sA[i][j];
    }
    if (tid >= NBT) return;
    // column c = tid of L^-1: v_i = (delta_ic - sum_{k<i} L_ik v_k) / L_ii, four partial sums to shorten the chain
    double v[NBT];
#pragma unroll
    for (int i = 0; i < NBT; i++) {
        double s0 = (i == tid) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int k = 0; k + 3 < i; k += 4) {
            s0 -= sA[i][k] * v[k]; s1 -= sA[i][k + 1] * v[k + 1];
            s2 -= sA[i][k + 2] * v[k + 2]; s3 -= sA[i][k + 3] * v[k + 3];
        }
#pragma unroll
        for (int k = (i / 4) * 4; k < i; k++) s0 -= sA[i][k] * v[k];
        v[i] = (i >= tid) ? ((s0 + s1) + (s2 + s3)) / sA[i][i] : 0.0;
    }
    if (tid < nb) {
#pragma unroll
        for (int i = 0; i < NBT; i++)
            if (i < nb) T.inv[i + (long long)tid * nb] = v[i];
    }
}

// 64 x 64 variant with register tiling: the block is held as 4 x 4 tiles in the registers of a 16 x 16 thread grid
// (thread (ty, tx), ty >= tx, owns rows 4ty.., columns 4tx..). 16 panel steps, two barriers each:
//   (a) the diagonal thread factors its 4 x 4 tile (scalar code, 4 dependent rsqrt) and publishes it,
//   (b) the threads below it solve their tiles against it and publish the 64 x 4 column panel,
//   (c) everybody to the right applies the rank-4 update to its own tile from the published panel.
// Then the factor sits in shared memory and the first 64 threads invert it column by column as above.
__global__ void __launch_bounds__(256)
potrf_inv64_kernel(const PanelTask *__restrict__ tasks, int *__restrict__ fail_col, long long bstride) {
    constexpr int N = 64, LDS_ = 66;                  // even leading dimension: 16-byte aligned row pairs
    __shared__ __align__(16) double sA[N * LDS_];     // sA[i * LDS_ + j]
    __shared__ double sdiag[16];                      // published 4 x 4 diagonal tile (lower) ...
    __shared__ double srinv[4];                       // ... and the reciprocals of its diagonal
    __shared__ double srinv_all[N];                   // 1 / L_ii of every row, reused by the inversion
    __shared__ __align__(16) double spanel2[2][N][4];  // published column panel (rows 0..63 of the current 4 columns),
                                                       // double-buffered: step P+1 publishes while step P is still read
    pdl_launch_dependents();
    PanelTask T = tasks[blockIdx.x];
    T.D = lane_ptr(T.D, bstride); T.inv = lane_ptr(T.inv, bstride); fail_col = lane_ptr(fail_col, bstride);
    const int nb = T.nb, tid = threadIdx.x;
    pdl_wait();
    for (int e = tid; e < N * N; e += 256) {
        const int i = e % N, j = e / N;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j <= i) v = T.D[i + (long long)j * T.ld];
        sA[i * LDS_ + j] = v;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    const bool active = ty >= tx;
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) a[r][c] = active ? sA[(4 * ty + r) * LDS_ + 4 * tx + c] : 0.0;
#pragma unroll 1
    for (int P = 0; P < 16; P++) {
        double (*spanel)[4] = spanel2[P & 1];
        if (ty == P && tx == P) {
            double ri[4], piv[4];
            chol4x4_lower(a, ri, piv);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (!(piv[c] > 0.0) && 4 * P + c < nb) atomicMin(fail_col, T.col0 + 4 * P + c + 1);
                srinv[c] = ri[c];
                srinv_all[4 * P + c] = ri[c];
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    sdiag[r * 4 + c] = (c <= r) ? a[r][c] : 0.0;
                    spanel[4 * P + r][c] = (c <= r) ? a[r][c] : 0.0;
                }
        }
        __syncthreads();
        if (tx == P && ty > P) {
            // X * L_PP^T = B on the 4 x 4 tile, column by column
#pragma unroll
            for (int c = 0; c < 4; c++) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    double s = a[r][c];
#pragma unroll
                    for (int k = 0; k < c; k++) s -= a[r][k] * sdiag[c * 4 + k];
                    a[r][c] = s * srinv[c];
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) spanel[4 * ty + r][c] = a[r][c];
        }
        __syncthreads();
        if (tx > P && active) {
            double pr[4][4], pc[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int k = 0; k < 4; k++) { pr[r][k] = spanel[4 * ty + r][k]; pc[r][k] = spanel[4 * tx + r][k]; }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int k = 0; k < 4; k++) a[r][c] -= pr[r][k] * pc[c][k];
        }
        // the next step's (a) works on registers and publishes into the other panel buffer
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) sA[(4 * ty + r) * LDS_ + 4 * tx + c] = (4 * tx + c <= 4 * ty + r) ? a[r][c] : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += 256) {
        const int i = e % nb, j = e / nb;
        if (i >= j) T.D[i + (long long)j * T.ld] = sA[i * LDS_ + j];
    }
    double v[N];
    if (tid < N) {
        // column c = tid of L^-1 by forward substitution; rows of L are read as 16-byte pairs (broadcast), the
        // divisions are multiplications with the reciprocal pivots kept from the factorization
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s0 = (i == tid) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double2 *row = reinterpret_cast<const double2 *>(sA + i * LDS_);
#pragma unroll
            for (int k = 0; k + 3 < i; k += 4) {
                const double2 l0 = row[k / 2], l1 = row[k / 2 + 1];
                s0 -= l0.x * v[k]; s1 -= l0.y * v[k + 1];
                s2 -= l1.x * v[k + 2]; s3 -= l1.y * v[k + 3];
            }
#pragma unroll
            for (int k = (i / 4) * 4; k < i; k++) s0 -= sA[i * LDS_ + k] * v[k];
            v[i] = (i >= tid) ? ((s0 + s1) + (s2 + s3)) * srinv_all[i] : 0.0;
        }
    }
    __syncthreads();                                   // everybody is done with the factor in shared memory
    if (tid < N) {
#pragma unroll
        for (int i = 0; i < N; i++) sA[i * LDS_ + tid] = v[i];   // inverse, element (i, c)
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += 256) {         // coalesced store: inv[i + c*nb]
        const int i = e % nb, c = e / nb;
        T.inv[e] = sA[i * LDS_ + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Split-K epilogue: C -= part_0 + part_1 + ... (fixed order -> bit-reproducible). The k-range of a product with too
// few output tiles to fill the GPU (the left-looking block-column updates of the top supernodes: m x 256 outputs,
// k up to 30,000) is cut into `nsplit` slices whose partial products go to a scratch buffer; this kernel folds them.
// ------------------------------------------------------------------------------------------------
struct SplitTask {
    double *C;
    const double *part;      // nsplit slabs of ldp x n doubles, slab stride `stride`
    long long stride;
    int m, n, ldc, ldp, nsplit, lower;
};

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const SplitTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks, long long bstride) {
    pdl_launch_dependents();
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    SplitTask T = tasks[t];
    T.C = lane_ptr(T.C, bstride); T.part = lane_ptr(T.part, bstride);
    pdl_wait();
    const int local = blockIdx.x - tile_prefix[t];
    const int mt = (T.m + 255) / 256;                 // a CTA owns 256 rows x 4 columns
    const int r = (local % mt) * 256 + threadIdx.x;
    const int c0 = (local / mt) * 4;
    if (r >= T.m) return;
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
        const int c = c0 + cc;
        if (c >= T.n || (T.lower && r < c)) continue;
        const double *p = T.part + r + (long long)c * T.ldp;
        double s = p[0];
        for (int q = 1; q < T.nsplit; q++) s += p[q * T.stride];
        T.C[r + (long long)c * T.ldc] -= s;
    }
}

// ------------------------------------------------------------------------------------------------
// Q.nzval -> panels (after the panel array has been zeroed): Lx[dst[k]] = nzval[src[k]]
// ------------------------------------------------------------------------------------------------
__global__ void scatter_q_kernel(double *__restrict__ Lx, const double *__restrict__ nz,
                                 const long long *__restrict__ src, const long long *__restrict__ dst, long long cnt,
                                 long long bstride) {
    Lx = lane_ptr(Lx, bstride); nz = lane_ptr(nz, bstride);
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < cnt; k += stride) Lx[dst[k]] = nz[src[k]];
}

// nzval[k] = sum_j coeff[j] * basis[j * nnz + k]: fixed-pattern value assembly for hyperparameter loops (the Matern
// precision is a polynomial in kappa^2 on a fixed pattern: matern_spde.jl:332-356, fem_utils.jl:313-335). HBM-bound:
// (nbasis + 1) * 8 * nnz bytes; the coefficients travel as kernel arguments.
constexpr int MAX_VALUE_BASIS = 8;
struct BasisCoeff { double c[MAX_VALUE_BASIS]; };
__global__ void __launch_bounds__(256)
combine_basis_kernel(double *__restrict__ nz, const double *__restrict__ basis, BasisCoeff coeff, int nbasis, long long nnz) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        double v = 0.0;
        for (int j = 0; j < nbasis; j++) v += coeff.c[j] * basis[(long long)j * nnz + k];
        nz[k] = v;
    }
}

// nz[diagpos[i]] = base[diagpos[i]] - diag[i]: the Newton iterate Q_prior - H(x_k) for a diagonal observation Hessian
// (_subtract_diagonal_hessian!, src/workspace/gaussian_approximation.jl:63-72) formed on values resident in HBM.
__global__ void __launch_bounds__(256)
minus_diag_kernel(double *__restrict__ nz, const long long *__restrict__ diagpos, const double *__restrict__ diag, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long p = diagpos[i];
    if (p >= 0) nz[p] -= diag[i];
}

__global__ void fill_zero_kernel(double *__restrict__ p, long long cnt) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < cnt; k += stride) p[k] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// Extend-add ("parent pulls"): each CTA owns a tile of ASM_CW consecutive front columns of one parent and
// walks the parent's children in fixed order, so every front entry has one owner and one summation order.
// Columns < ns land in the parent's panel, columns >= ns in its update matrix (zero-filled here first).
// ------------------------------------------------------------------------------------------------
constexpr int ASM_CW = 32;

__global__ void __launch_bounds__(256)
assemble_kernel(const AsmItem *__restrict__ items, const SuperMeta *__restrict__ meta,
                const int *__restrict__ child_idx, const int *__restrict__ relidx,
                double *__restrict__ Lx0, double *__restrict__ upd0, long long bstride) {
    double *__restrict__ Lx = lane_ptr_pinned(Lx0, bstride);
    double *__restrict__ upd = lane_ptr_pinned(upd0, bstride);
    pdl_launch_dependents();
    const AsmItem it = items[blockIdx.x];
    const SuperMeta P = meta[it.super];
    pdl_wait();
    const int c_lo = it.col0, c_hi = min(it.col0 + ASM_CW, P.nrow);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nr = P.nrow - P.ns;
    double *Lp = Lx + P.panel_off;
    double *Up = upd + P.upd_off;
    // zero-fill the owned columns of the update matrix (lower triangle only)
    for (int c = max(c_lo, P.ns) + warp; c < c_hi; c += 8) {
        const int uc = c - P.ns;
        for (int r = uc + lane; r < nr; r += 32) Up[r + (long long)uc * P.uld] = 0.0;
    }
    __syncthreads();
    for (int ci = P.child_begin; ci < P.child_end; ci++) {
        const SuperMeta C = meta[child_idx[ci]];
        const int cnr = C.nrow - C.ns;
        const int *rel = relidx + C.rowptr + C.ns;   // cnr entries, strictly increasing
        // child update columns whose parent column falls in [c_lo, c_hi)
        int a, b;
        {
            int lo = 0, hi = cnr;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rel[mid] < c_lo) lo = mid + 1; else hi = mid; }
            a = lo;
            hi = cnr;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rel[mid] < c_hi) lo = mid + 1; else hi = mid; }
            b = lo;
        }
        const double *Uc = upd + C.upd_off;
        for (int jc = a + warp; jc < b; jc += 8) {
            const int pc = rel[jc];
            const double *src = Uc + (long long)jc * C.uld;
            // rel is strictly increasing, so the 4 read-modify-writes of a batch touch distinct entries: all loads of a
            // batch are issued before its first store (the plain `dst[rel[i]] += src[i]` loop serialised on possible
            // aliasing: ncu showed one memory round trip per element, long-scoreboard stalls 67 per issue)
            double *dst = (pc < P.ns) ? (Lp + (long long)pc * P.ld) : (Up + (long long)(pc - P.ns) * P.uld - P.ns);
            for (int ic = jc + lane; ic < cnr; ic += 128) {
                int ri[4];
                double sv[4], dv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = ic + 32 * u;
                    ri[u] = i < cnr ? rel[i] : -1;
                    sv[u] = i < cnr ? src[i] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) dv[u] = ri[u] >= 0 ? dst[ri[u]] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (ri[u] >= 0) dst[ri[u]] = dv[u] + sv[u];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// log det Q = 2 * sum_j log L_jj, fused into the factorization graph. Two fixed-shape passes.
// ------------------------------------------------------------------------------------------------
constexpr int LOGDET_BLOCKS = 256;

__device__ __forceinline__ double block_sum_256(double v, double *sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    return sh[0];
}

__global__ void __launch_bounds__(256)
logdet_partial_kernel(const double *__restrict__ Lx, const long long *__restrict__ diag_pos, long long n,
                      double *__restrict__ partial, long long bstride) {
    Lx = lane_ptr(Lx, bstride); partial = lane_ptr(partial, bstride);
    __shared__ double sh[256];
    double acc = 0.0;
    // fixed assignment of indices to (block, thread) and fixed sequential order per thread
    for (long long j = blockIdx.x * 256LL + threadIdx.x; j < n; j += 256LL * LOGDET_BLOCKS) acc += log(Lx[diag_pos[j]]);
    double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) logdet_final_kernel(const double *__restrict__ partial, double *__restrict__ out, long long bstride) {
    partial = lane_ptr(partial, bstride); out = lane_ptr(out, bstride);
    __shared__ double sh[256];
    double s = block_sum_256(threadIdx.x < LOGDET_BLOCKS ? partial[threadIdx.x] : 0.0, sh);
    if (threadIdx.x == 0) out[0] = 2.0 * s;
}

// ------------------------------------------------------------------------------------------------
// Solve phase: right-hand sides live in a permuted n x nrhs column-major work array `y` (sweeps: solve_kernels.cuh).
// ------------------------------------------------------------------------------------------------
// y[k, r] = b[perm[k], r]  (gather)  /  x[perm[k], r] = y[k, r]  (scatter)
__global__ void permute_rows_kernel(double *__restrict__ dst, const double *__restrict__ src,
                                    const long long *__restrict__ perm, long long n, long long ld_dst,
                                    long long ld_src, int nrhs, int scatter) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long p = perm[k];
    for (int r = 0; r < nrhs; r++) {
        if (scatter) dst[p + r * ld_dst] = src[k + r * ld_src];
        else dst[k + r * ld_dst] = src[p + r * ld_src];
    }
}

// ------------------------------------------------------------------------------------------------
// Selected inversion kernels
// ------------------------------------------------------------------------------------------------
// W_s[a,b] = Z(R_a, R_b), gathered from the parent's front [Z panel of p ; W_p] through the relative
// indices; written as a full symmetric nr x nr matrix so the following products are plain GEMMs.
__global__ void __launch_bounds__(256)
selinv_gather_kernel(const AsmItem *__restrict__ items, const SuperMeta *__restrict__ meta,
                     const int *__restrict__ relidx, const double *__restrict__ Zx, double *__restrict__ zw) {
    // A CTA owns 32 columns of W_s and walks down the 32-row tiles from the diagonal: the lower-triangle entries are
    // gathered with lanes along the rows (coalesced reads of the parent column, coalesced writes of W), the mirrored
    // upper-triangle entries go through a shared-memory tile so that their writes are coalesced as well
    // (the direct transposed store cost 4x write amplification: 430 GB/s in the first 1 M-dof profile).
    __shared__ double tile[32][33];
    const AsmItem it = items[blockIdx.x];
    const SuperMeta S = meta[it.super];
    const SuperMeta P = meta[S.parent];
    const int nr = S.nrow - S.ns;
    const int *rel = relidx + S.rowptr + S.ns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *Zp = Zx + P.panel_off;
    const double *Wp = zw + P.zw_off;
    double *Ws = zw + S.zw_off;
    const int c0 = it.col0;
    const double *src[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = c0 + warp * 4 + q;
        const int pb = b < nr ? rel[b] : 0;
        src[q] = (pb < P.ns) ? (Zp + (long long)pb * P.ld) : (Wp + (long long)(pb - P.ns) * P.uld - P.ns);
    }
    for (int a0 = c0; a0 < nr; a0 += 32) {
        const int a = a0 + lane;
        const int ra = a < nr ? rel[a] : 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int bb = warp * 4 + q, b = c0 + bb;
            double v = 0.0;
            if (a < nr && b < nr && a >= b) {
                v = src[q][ra];
                Ws[a + (long long)b * S.uld] = v;
            }
            tile[bb][lane] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int aa = warp * 4 + q, ar = a0 + aa, b = c0 + lane;
            if (ar < nr && b < nr && ar > b) Ws[b + (long long)ar * S.uld] = tile[lane][aa];
        }
        __syncthreads();
    }
}

// In-place transpose of the leading ns x ns block of a panel (tile pairs swapped through smem).
struct TransTask { double *A; int n, ld; };
__global__ void __launch_bounds__(256)
transpose_inplace_kernel(const TransTask *__restrict__ tasks, const int *__restrict__ tile_prefix, int ntasks) {
    __shared__ double s1[32][33], s2[32][33];
    const int t = find_task(tile_prefix, ntasks, blockIdx.x);
    const TransTask T = tasks[t];
    // enumerate tile pairs (ti >= tj) from the linear index
    int local = blockIdx.x - tile_prefix[t];
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= local) ti++;
    const int tj = local - ti * (ti + 1) / 2;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int yy = ty; yy < 32; yy += 8) {
        int r = ti * 32 + tx, c = tj * 32 + yy;
        s1[yy][tx] = (r < T.n && c < T.n) ? T.A[r + (long long)c * T.ld] : 0.0;
        r = tj * 32 + tx; c = ti * 32 + yy;
        s2[yy][tx] = (r < T.n && c < T.n) ? T.A[r + (long long)c * T.ld] : 0.0;
    }
    __syncthreads();
    for (int yy = ty; yy < 32; yy += 8) {
        // block (tj, ti) := transpose of old block (ti, tj): new A[tj*32+tx][ti*32+yy] = old A[ti*32+yy][tj*32+tx] = s1[tx][yy]
        int r = tj * 32 + tx, c = ti * 32 + yy;
        if (r < T.n && c < T.n) T.A[r + (long long)c * T.ld] = s1[tx][yy];
        if (ti != tj) {
            r = ti * 32 + tx; c = tj * 32 + yy;
            if (r < T.n && c < T.n) T.A[r + (long long)c * T.ld] = s2[tx][yy];
        }
    }
}

// out[k] = Zx[pos[k]] (pos < 0 -> 0.0): diagonal extraction, CSC materialisation, extract-at-pattern
__global__ void gather_values_kernel(double *__restrict__ out, const double *__restrict__ Zx,
                                     const long long *__restrict__ pos, long long cnt) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < cnt; k += stride) {
        const long long p = pos[k];
        out[k] = p >= 0 ? Zx[p] : 0.0;
    }
}


// ------------------------------------------------------------------------------------------------
// tr(Q^-1 B) = sum_k w_k * Z[pos_k] * B[idx_k]  with the contraction on the device (selinv_dot, backend.jl:265-267):
// the gathered values never leave HBM, 8 bytes per value set travel back. Fixed assignment of entries to
// (block, thread), fixed order per thread, fixed-shape tree per block and over the blocks => bit-reproducible.
// blockIdx.y selects the value set (stride `vstride` doubles); idx == nullptr: B[k]; w == nullptr: weight 1.
// HBM-bound: 8 (pos) + 8 (Z, scattered sector reads) + 8 (B) [+ 8 idx + 8 w] bytes per entry.
// ------------------------------------------------------------------------------------------------
constexpr int DOT_BLOCKS = 256;

__global__ void __launch_bounds__(256)
gather_dot_partial_kernel(const double *__restrict__ Zx, const long long *__restrict__ pos,
                          const long long *__restrict__ idx, const double *__restrict__ w,
                          const double *__restrict__ val, long long vstride, long long cnt,
                          double *__restrict__ partial) {
    __shared__ double sh[256];
    val += (long long)blockIdx.y * vstride;
    double acc = 0.0;
    for (long long k = blockIdx.x * 256LL + threadIdx.x; k < cnt; k += 256LL * DOT_BLOCKS) {
        const long long p = pos[k];
        if (p < 0) continue;                       // outside the factor's pattern: Sigma is not available there, counts 0
        const double b = val[idx ? idx[k] : k];
        const double z = Zx[p];
        acc += (w ? w[k] : 1.0) * (z * b);
    }
    const double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * DOT_BLOCKS + blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) gather_dot_final_kernel(const double *__restrict__ partial, double *__restrict__ out) {
    __shared__ double sh[256];
    const double s = block_sum_256(threadIdx.x < DOT_BLOCKS ? partial[(long long)blockIdx.y * DOT_BLOCKS + threadIdx.x] : 0.0, sh);
    if (threadIdx.x == 0) out[blockIdx.y] = s;
}

// ------------------------------------------------------------------------------------------------
// out[i] = sum over the entries p of segment i of w[p] * Z[pos[p]] (pos < 0: outside the factor's pattern, counts 0):
// diag(A Sigma A') of a sparse design matrix, one segment per row of A holding its nnz^2 index pairs
// (_row_diag_AΣAt, src/linear_predictor_marginals.jl:137-165). One warp per row, lanes stride the segment, fixed
// shuffle tree: bit-reproducible.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segment_dot_kernel(const double *__restrict__ Zx, const long long *__restrict__ pos, const double *__restrict__ w,
                   const long long *__restrict__ segptr, long long nseg, double *__restrict__ out) {
    const long long i = (blockIdx.x * 256LL + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= nseg) return;
    double acc = 0.0;
    for (long long p = segptr[i] + lane; p < segptr[i + 1]; p += 32) {
        const long long q = pos[p];
        if (q >= 0) acc += w[p] * Zx[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = acc;
}

// nz[pos[k]] -= v[k] (positions unique): the Newton iterate Q_prior - H(x_k) for a SPARSE observation Hessian
// (_subtract_sparse_hessian!, src/workspace/gaussian_approximation.jl:74-83) formed on values resident in HBM.
__global__ void __launch_bounds__(256)
minus_sparse_kernel(double *__restrict__ nz, const long long *__restrict__ pos, const double *__restrict__ v, long long cnt) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k < cnt) nz[pos[k]] -= v[k];
}

}  // namespace gmrf
