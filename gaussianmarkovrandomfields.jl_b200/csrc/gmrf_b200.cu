// gmrf_b200.cu -- handle, schedule construction and the C-ABI of libgmrf_b200.so (see include/gmrf_b200.h).
//
// Execution model: every numeric phase (factorization+logdet, forward/backward solves, selected inversion) is a
// static, level-ordered list of kernel launches over task tables that were built on the host at analysis time
// and live in HBM next to the factor. The factorization and selected-inversion lists are captured into CUDA
// graphs once per handle, so Newton / hyperparameter loops replay a graph per refactorization.
#include "../../include/gmrf_b200.h"
#include "kernels.cuh"
#include "solve_kernels.cuh"
#include "front_kernels.cuh"
#include "assemble_kernels.cuh"
#include "symbolic.hpp"

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler is attached

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

using namespace gmrf;

namespace {

thread_local std::string g_create_error;
int g_alloc_fail_countdown = 0;     // test hook ("debug_alloc_fail_after"): the k-th dev_alloc from now fails

enum LaunchKind : int {
    K_ASSEMBLE, K_CHAIN, K_FINALIZE, K_FRONT, K_GEMM_NN_S, K_GEMM_NN_L, K_GEMM_NT_S, K_GEMM_NT_L, K_GEMM_TT_S, K_GEMM_TT_L,
    K_GATHER, K_TRANSPOSE, K_FWD_ASM, K_FWD_STEP, K_BWD_GATHER, K_BWD_STEP, K_PANEL, K_SPLIT_REDUCE, K_FWD_ASM_M, K_ROWS_GATHER, K_BWD_REDUCE, K_ASSEMBLE_G,
    K_FWD_WSTEP, K_FWD_WDIAG, K_BWD_WSTEP, K_BWD_WDIAG
};

struct Launch {
    int kind;
    int aux;            // panel step: NBT bucket; front launch: dynamic shared memory (bytes)
    int aux2 = 0;       // front launch: widest diagonal block (selects the kernel instantiation)
    double flops = 0;   // GEMM launches: algorithmic flops of the launch's tasks (diagnostics)
    int kmax = 0;       // GEMM launches: longest contraction
    i64 task_off;       // first task in the kind's task array
    int ntasks;
    i64 prefix_off;     // offset into the tile-prefix array (ntasks+1 entries) or -1
    int grid;
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void free_() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Plan {
    std::vector<Launch> launches;
};

}  // namespace

struct gmrf_b200_handle {
    Symbolic S;
    Options opt;
    int device = -1;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;    // side stream: work off the factorization's critical path (block inverses)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool factored = false, selinv_valid = false;
    int fail_col = 0;
    double logdet = 0.0;
    double t_ms[5] = {0, 0, 0, 0, 0};
    size_t device_bytes = 0;
    double gemm_flops_factor = 0, gemm_flops_selinv = 0;

    // device arrays
    double *d_Lx = nullptr, *d_upd = nullptr, *d_nz = nullptr, *d_Zx = nullptr, *d_zw = nullptr;
    double *d_y = nullptr, *d_uvec = nullptr, *d_io = nullptr, *d_partial = nullptr, *d_scalars = nullptr;  // scalars: [0]=logdet
    int *d_fail = nullptr;
    i64 io_cap = 0;
    long long *d_qsrc = nullptr, *d_qdst = nullptr, *d_diagpos = nullptr, *d_perm = nullptr;
    int *d_rowidx = nullptr, *d_relidx = nullptr, *d_child = nullptr, *d_prefix = nullptr, *d_superlist = nullptr;
    SuperMeta *d_meta = nullptr;
    GemmTask *d_gemm = nullptr;
    PanelTask *d_panel = nullptr;
    AsmItem *d_items = nullptr;
    FwdStepTask *d_fwd = nullptr;
    BwdGatherTask *d_bwdg = nullptr;
    BwdStepTask *d_bwds = nullptr;
    BwdReduceTask *d_bwdr = nullptr;
    WideStepTask *d_wstep = nullptr;   // 256-column steps of the long chains (solve_kernels.cuh)
    WideDiagTask *d_wdiag = nullptr;
    double *d_bwdpart = nullptr;       // partial sums of the row-chunked L21' x_R products (backward sweep)
    double *d_Linv = nullptr;          // inverted 64-column diagonal blocks: written by the factorization (TRSM by
                                       // GEMM), reused by the solve phase
    long long *d_invbase = nullptr;    // per supernode: offset of its first inverted block in d_Linv
    std::vector<long long> inv_base;   // host copy
    std::map<int, cudaGraphExec_t> solve_graphs;   // key = nrhs * 2 + mode
    // wide right-hand-side path (blocks of 64 / 128 / 256 columns, DMMA GEMM sweeps), one plan per width, built lazily;
    // the work arrays are shared (sized for the widest block built so far)
    struct Multi {
        bool built = false;
        GemmTask *d_gemm = nullptr;
        RowGatherTask *d_rg = nullptr;
        int *d_prefix = nullptr, *d_superlist = nullptr;
        Plan fwd_plan, bwd_plan;
        cudaGraphExec_t graph[2] = {nullptr, nullptr};
    } multi[MULTI_NW];
    double *d_ym = nullptr, *d_um = nullptr;
    int multi_wcap = 0;                // columns the shared work arrays hold
    TransTask *d_trans = nullptr;
    SplitTask *d_split = nullptr, *d_split_z = nullptr;
    ChainTask *d_chain = nullptr;      // fused chain steps (front_kernels.cuh)
    FinalizeTask *d_final = nullptr;
    FrontTask *d_front = nullptr;
    GatherCtx *d_gctx = nullptr;       // tables of the gathering epilogue of the update-matrix products
    AsmTile *d_asmtiles = nullptr;     // gather extend-add (assemble_kernels.cuh)
    int *d_relpos = nullptr;           // per child: position in its relative-index list of every 256-row boundary of the parent
    long long *d_relpos_off = nullptr;
    double *d_sq = nullptr;            // parked factored diagonal squares of the fused chain steps (128 x 128 each)
    int front_smem_max = 0;            // dynamic shared memory the one-CTA-per-front kernel was configured for
    i64 n_large_tile_launches = 0, n_splitk_tasks = 0, n_fast_roots = 0, n_chain_launches = 0, n_front_launches = 0;
    double *d_base = nullptr;          // optional resident copy of a prior's nzval (Newton loops: Q_prior - H on the device)
    double *d_hdiag = nullptr;
    long long *d_hpos = nullptr;       // nzval positions of a sparse observation Hessian (set_hessian_pattern)
    i64 hpos_count = 0;
    long long *d_diagnz = nullptr;     // nzval position of every diagonal entry of the input pattern (-1: not stored)
    std::vector<long long> diag_nzpos;
    double *d_basis = nullptr;         // optional value basis (nbasis x nnz) for device-side assembly of nzval
    int nbasis = 0;
    double *d_dot = nullptr;           // selinv_dot: DOT_BLOCKS partial sums per value set, then the results
    double *d_qzw = nullptr;           // basis traces: weight per scatter-map entry (1 diagonal, 2 off-diagonal)
    double *d_splitk = nullptr;        // scratch for split-K partial products
    i64 splitk_cap = 0;                // doubles
    // selinv task tables are built lazily (they need d_Zx / d_zw)
    GemmTask *d_gemm_z = nullptr;
    AsmItem *d_items_z = nullptr;
    int *d_prefix_z = nullptr;
    // selinv CSC materialisation (lazy)
    std::vector<i64> z_colptr, z_rowval;
    long long *d_zpos = nullptr, *d_zdiagpos = nullptr;
    double *d_zout = nullptr;
    bool z_pattern_built = false;
    bool selinv_tables_built = false;  // set only after EVERY table of the selected-inversion plan is resident
    // last caller pattern looked up by selinv_extract / selinv_dot (host-side, content-compared): a repeated pattern
    // (gradient loops, linear-predictor marginals) skips the lookup
    std::vector<i64> pp_colptr, pp_rowval;
    std::vector<long long> pp_pos;
    int pp_base = 0;
    bool pp_valid = false;
    i64 pp_hits = 0;
    long long *d_ppos = nullptr;       // the same positions resident in HBM: a repeated pattern costs no upload either
    i64 d_ppos_cap = 0;
    bool d_ppos_valid = false;
    unsigned long long pattern_hash = 0;   // of the (0-based) input pattern: ties an exported analysis to its matrix
    // factor export P'L as CSC (lazy; CholeskySqrt-style consumers)
    std::vector<i64> l_colptr, l_rowval;
    long long *d_lpos = nullptr;
    bool l_pattern_built = false;

    Plan factor_plan, selinv_plan, fwd_plan, bwd_plan;
    std::map<int, cudaGraphExec_t> factor_graphs;   // key = number of lanes advanced by the graph
    cudaGraphExec_t selinv_graph = nullptr;
    // lanes: independent value sets on the same pattern factorized by the same launches (blockIdx.y); every numeric
    // array of the factorization lives in one arena per lane, lane b starts arena_bytes * b after lane 0
    int lanes = 1;
    int last_lanes = 1;     // lanes advanced by the last factorization
    long long arena_bytes = 0;
    double *d_arena = nullptr;
    std::vector<double> lane_logdet;
    std::vector<int> lane_fail;
    int rhs_block = 8;

    std::vector<void *> owned;   // every cudaMalloc for destroy
    std::map<void *, size_t> owned_bytes;
};

namespace {

// NVTX range per phase (`ncu --nvtx --nvtx-include "gmrf_b200:factor/"` etc. narrows a capture to one phase)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (h)->err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #expr;       \
            return GMRF_B200_ERR_CUDA;                                                            \
        }                                                                                         \
    } while (0)

template <class T>
int dev_alloc(gmrf_b200_handle *h, T **p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    if (g_alloc_fail_countdown > 0 && --g_alloc_fail_countdown == 0) {
        h->err = "cudaMalloc failed (injected by debug_alloc_fail_after)";
        return GMRF_B200_ERR_ALLOC;
    }
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc failed (") + std::to_string(count * sizeof(T)) + " bytes): " + cudaGetErrorString(e);
        return GMRF_B200_ERR_ALLOC;
    }
    h->owned.push_back(*p);
    h->owned_bytes[(void *)*p] = count * sizeof(T);
    h->device_bytes += count * sizeof(T);
    return 0;
}

template <class T>
int dev_upload(gmrf_b200_handle *h, T **p, const std::vector<T> &v) {
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(h, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Release one allocation made through dev_alloc (grow-only buffers that are being replaced).
template <class T>
void dev_free(gmrf_b200_handle *h, T *&p) {
    if (!p) return;
    auto it = std::find(h->owned.begin(), h->owned.end(), (void *)p);
    if (it != h->owned.end()) h->owned.erase(it);
    auto ib = h->owned_bytes.find((void *)p);
    if (ib != h->owned_bytes.end()) { h->device_bytes -= ib->second; h->owned_bytes.erase(ib); }
    cudaFree(p);
    p = nullptr;
}

inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Host-side plan builder
// ------------------------------------------------------------------------------------------------
struct Builder {
    std::vector<GemmTask> gemm;
    std::vector<PanelTask> panel;
    std::vector<AsmItem> items;
    std::vector<FwdStepTask> fwd;
    std::vector<BwdGatherTask> bwdg;
    std::vector<BwdStepTask> bwds;
    std::vector<BwdReduceTask> bwdr;
    std::vector<WideStepTask> wstep;
    std::vector<WideDiagTask> wdiag;
    std::vector<TransTask> trans;
    std::vector<SplitTask> split;
    std::vector<RowGatherTask> rowgather;
    std::vector<ChainTask> chain;
    std::vector<FinalizeTask> finalize;
    std::vector<FrontTask> front;
    std::vector<AsmTile> asmtiles;
    i64 n_large_tile_launches = 0, n_splitk_tasks = 0;
    double *splitk_base = nullptr;     // device scratch for split-K partial products
    i64 splitk_cap = 0;
    int splitk_min_k = 1024;
    int large_tile_mask = 1;           // bit v set: operand-layout variant v (0 NN, 1 NT, 2 TT) may use the 128 x 64 tile
    std::vector<int> superlist;
    std::vector<int> prefix;
    bool naive = false;
    double gemm_flops = 0;

    // GEMM launches: tasks are split into a small-tile and a large-tile launch
    // GEMM launches: tasks are split into a small-tile and a large-tile launch. With allow_split, products that leave
    // the GPU under-filled (few output tiles, very long k) are cut along k into slices that write partial products to
    // the split-K scratch, folded afterwards in fixed order by splitk_reduce_kernel.
    void add_gemm(Plan &plan, std::vector<GemmTask> &tasks, int variant /*0=NN,1=NT,2=TT*/, bool force_small = false,
                  double flop_weight = 1.0, bool allow_split = false) {
        if (tasks.empty()) return;
        std::vector<SplitTask> reduce;
        if (allow_split && !naive && splitk_base) {
            const i64 slots = 148 * 4;                       // resident 64 x 64-tile CTAs on a B200
            i64 tiles = 0;
            for (auto &t : tasks) tiles += (i64)cdiv(t.m, 64) * cdiv(t.n, 64);
            if (tiles < 4 * slots) {
                std::vector<GemmTask> out;
                i64 used = 0;
                for (auto &t : tasks) {
                    int S = (int)std::min<i64>({8, (4 * slots + tiles - 1) / std::max<i64>(tiles, 1), (i64)t.k / splitk_min_k});
                    const i64 ldp = (t.m + 1) & ~1LL;
                    const bool plain = (t.flags & ~GEMM_LOWER) == 0;          // C -= A B^T, optionally lower-only
                    if (S < 2 || !plain || t.m <= 0 || t.n <= 0 || used + (i64)S * ldp * t.n > splitk_cap) { out.push_back(t); continue; }
                    const int kslice = ((t.k + S - 1) / S + 15) & ~15;
                    S = (t.k + kslice - 1) / kslice;
                    double *part = splitk_base + used;
                    const long long stride = ldp * t.n;
                    used += (i64)S * stride;
                    const bool ta = variant == 2, tb = variant >= 1;
                    for (int q = 0; q < S; q++) {
                        GemmTask g = t;
                        const int kk0 = q * kslice;
                        g.k = std::min(kslice, t.k - kk0);
                        g.A = t.A + (ta ? (long long)kk0 : (long long)kk0 * t.lda);
                        g.B = t.B + (tb ? (long long)kk0 : (long long)kk0 * t.ldb);
                        g.C = part + q * stride;
                        g.ldc = (int)ldp;
                        g.flags = (t.flags & GEMM_LOWER) | GEMM_BETA0 | GEMM_ALPHA_POS;
                        out.push_back(g);
                    }
                    reduce.push_back(SplitTask{t.C, part, stride, t.m, t.n, t.ldc, (int)ldp, S, (t.flags & GEMM_LOWER) ? 1 : 0});
                }
                tasks.swap(out);
            }
        }
        std::vector<GemmTask> small, large;
        for (auto &t : tasks) {
            if (t.m <= 0 || t.n <= 0) continue;
            gemm_flops += flop_weight * 2.0 * t.m * (double)t.n * t.k * ((t.flags & GEMM_LOWER) ? 0.5 * (1.0 + 1.0 / std::max(1, t.m)) * ((double)t.n <= t.m ? (2.0 - (double)t.n / t.m) : 1.0) : 1.0);
            // measured on B200 (profiles/r01_gemm_variants.log): the 64x64 tile (4 CTAs/SM) wins everywhere except very
            // large square-ish products, where the 128x64 tile (warp tile 64x32) is ~4% faster
            bool big = !naive && !force_small && (i64)t.m * t.n >= 4096LL * 4096LL && t.n >= 1024 && ((large_tile_mask >> variant) & 1);
            (big ? large : small).push_back(t);
        }
        for (int pass = 0; pass < 2; pass++) {
            auto &v = pass ? large : small;
            if (v.empty()) continue;
            int BM = naive ? 16 : (pass ? 128 : 64), BN = naive ? 16 : 64;
            Launch L;
            L.kind = (variant == 0 ? K_GEMM_NN_S : variant == 1 ? K_GEMM_NT_S : K_GEMM_TT_S) + pass;
            L.aux = 0;                                     // 1: the launch holds gathering update-matrix products
            for (auto &t : v) if (t.flags & GEMM_GATHER) L.aux = 1;
            L.task_off = (i64)gemm.size();
            L.ntasks = (int)v.size();
            L.prefix_off = (i64)prefix.size();
            i64 tot = 0;
            for (auto &t : v) {
                prefix.push_back((int)tot);
                tot += (i64)cdiv(t.m, BM) * cdiv(t.n, BN);
                gemm.push_back(t);
                L.flops += 2.0 * t.m * (double)t.n * t.k * ((t.flags & GEMM_LOWER) ? 0.5 * ((double)t.n <= t.m ? (2.0 - (double)t.n / t.m) : 1.0) : 1.0);
                L.kmax = std::max(L.kmax, t.k);
            }
            prefix.push_back((int)tot);
            if (tot > INT_MAX) throw std::runtime_error("too many GEMM tiles in one launch");
            L.grid = (int)tot;
            plan.launches.push_back(L);
            if (pass) n_large_tile_launches++;
        }
        tasks.clear();
        n_splitk_tasks += (i64)reduce.size();
        if (!reduce.empty()) {
            Launch L;
            L.kind = K_SPLIT_REDUCE;
            L.aux = 0;
            L.task_off = (i64)split.size();
            L.ntasks = (int)reduce.size();
            L.prefix_off = (i64)prefix.size();
            i64 tot = 0;
            for (auto &t : reduce) {
                prefix.push_back((int)tot);
                tot += (i64)cdiv(t.m, 256) * cdiv(t.n, 4);
                split.push_back(t);
            }
            prefix.push_back((int)tot);
            L.grid = (int)tot;
            plan.launches.push_back(L);
        }
    }
    void add_panel(Plan &plan, std::vector<PanelTask> &tasks) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = K_PANEL;
        int mx = 0;
        for (auto &t : tasks) mx = std::max(mx, t.nb);
        L.aux = mx <= 8 ? 8 : mx <= 16 ? 16 : mx <= 32 ? 32 : 64;
        L.task_off = (i64)panel.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = -1;
        L.grid = (int)tasks.size();
        panel.insert(panel.end(), tasks.begin(), tasks.end());
        plan.launches.push_back(L);
        tasks.clear();
    }
    void add_items(Plan &plan, std::vector<AsmItem> &its, int kind) {
        if (its.empty()) return;
        Launch L;
        L.kind = kind;
        L.aux = 0;
        L.task_off = (i64)items.size();
        L.ntasks = (int)its.size();
        L.prefix_off = -1;
        L.grid = (int)its.size();
        items.insert(items.end(), its.begin(), its.end());
        plan.launches.push_back(L);
        its.clear();
    }
    // solve-phase launches: `tiles(t)` CTAs per task, tasks appended to the kind's table
    template <class Task, class TilesFn>
    void add_tiled(Plan &plan, std::vector<Task> &tasks, std::vector<Task> &table, int kind, TilesFn tiles) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = kind;
        L.aux = 0;
        L.task_off = (i64)table.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = (i64)prefix.size();
        i64 tot = 0;
        for (auto &t : tasks) {
            prefix.push_back((int)tot);
            tot += tiles(t);
            table.push_back(t);
        }
        prefix.push_back((int)tot);
        if (tot > INT_MAX) throw std::runtime_error("too many tiles in one solve launch");
        L.grid = (int)tot;
        plan.launches.push_back(L);
        tasks.clear();
    }
    void add_trans(Plan &plan, std::vector<TransTask> &tasks) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = K_TRANSPOSE;
        L.aux = 0;
        L.task_off = (i64)trans.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = (i64)prefix.size();
        i64 tot = 0;
        for (auto &t : tasks) {
            prefix.push_back((int)tot);
            i64 nt = cdiv(t.n, 32);
            tot += nt * (nt + 1) / 2;
            trans.push_back(t);
        }
        prefix.push_back((int)tot);
        L.grid = (int)tot;
        plan.launches.push_back(L);
        tasks.clear();
    }
    void add_asm_tiles(Plan &plan, std::vector<AsmTile> &tiles) {
        if (tiles.empty()) return;
        Launch L;
        L.kind = K_ASSEMBLE_G;
        L.aux = 0;
        L.task_off = (i64)asmtiles.size();
        L.ntasks = (int)tiles.size();
        L.prefix_off = -1;
        L.grid = (int)tiles.size();
        asmtiles.insert(asmtiles.end(), tiles.begin(), tiles.end());
        plan.launches.push_back(L);
        tiles.clear();
    }
    // fused chain step: 1 diagonal CTA + one CTA per 64-row tile below the step's columns
    void add_chain(Plan &plan, std::vector<ChainTask> &tasks) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = K_CHAIN;
        L.aux = 0;
        L.task_off = (i64)chain.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = (i64)prefix.size();
        i64 tot = 0;
        for (auto &t : tasks) {
            prefix.push_back((int)tot);
            tot += 1 + cdiv(std::max(0, t.nrow - (t.k0 + t.nb0 + t.nb1)), FB);
            chain.push_back(t);
        }
        prefix.push_back((int)tot);
        if (tot > INT_MAX) throw std::runtime_error("too many tiles in one chain step");
        L.grid = (int)tot;
        plan.launches.push_back(L);
        tasks.clear();
    }
    void add_front(Plan &plan, std::vector<FrontTask> &tasks, int max_nb) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = K_FRONT;
        L.aux = 0;                                     // dynamic shared memory of the launch (bytes)
        for (auto &t : tasks) L.aux = std::max(L.aux, (int)((FTILE + FSCRATCH + t.smem_doubles) * sizeof(double)));
        L.aux2 = max_nb;
        L.task_off = (i64)front.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = -1;
        L.grid = (int)tasks.size();
        front.insert(front.end(), tasks.begin(), tasks.end());
        plan.launches.push_back(L);
        tasks.clear();
    }
    i64 finalize_done = 0;
    void add_finalize(Plan &plan) {          // one launch for the tasks pushed since the previous call
        if ((i64)finalize.size() == finalize_done) return;
        Launch L;
        L.kind = K_FINALIZE;
        L.aux = 0;
        L.task_off = finalize_done;
        L.ntasks = (int)((i64)finalize.size() - finalize_done);
        L.prefix_off = -1;
        L.grid = L.ntasks;
        plan.launches.push_back(L);
        finalize_done = (i64)finalize.size();
    }
    void add_wdiag(Plan &plan, std::vector<WideDiagTask> &tasks, int kind) {
        if (tasks.empty()) return;
        Launch L;
        L.kind = kind;
        L.aux = 0;
        L.task_off = (i64)wdiag.size();
        L.ntasks = (int)tasks.size();
        L.prefix_off = -1;
        L.grid = (int)tasks.size();
        wdiag.insert(wdiag.end(), tasks.begin(), tasks.end());
        plan.launches.push_back(L);
        tasks.clear();
    }
    void add_superlist(Plan &plan, std::vector<int> &lst, int kind) {
        if (lst.empty()) return;
        Launch L;
        L.kind = kind;
        L.aux = 0;
        L.task_off = (i64)superlist.size();
        L.ntasks = (int)lst.size();
        L.prefix_off = -1;
        L.grid = (int)lst.size();
        superlist.insert(superlist.end(), lst.begin(), lst.end());
        plan.launches.push_back(L);
        lst.clear();
    }
};

constexpr int NB = POTRF_NB;

// Which path every level of the assembly tree takes in the numeric factorization (decided once per analysis, before the
// numeric arena is sized):
//   2  front_small_kernel: every panel of the level fits in shared memory -> ONE launch for the level;
//   1  fused chain steps (chain_step_kernel): few enough 64-row tiles that the chain's latency dominates -> one launch
//      per 128 columns (+ the left-looking block-column GEMM) instead of 3 launches per 64 columns;
//   0  the bulk path: potrf+inverse per 64 columns, TRSM and updates as DMMA GEMM launches (throughput-bound levels).
struct FusedInfo {
    std::vector<int> mode;        // per level
    std::vector<i64> sq_base;     // per supernode: first parked square (units of 128 x 128 doubles), mode 1 only
    i64 sq_total = 0;
    int front_smem_max = 0;       // bytes
};


FusedInfo plan_fused(const Symbolic &S, const Options &opt) {
    FusedInfo F;
    F.mode.assign((size_t)S.nlevels, 0);
    F.sq_base.assign((size_t)S.nsuper, -1);
    if (opt.naive_kernels) return F;
    const i64 cap = std::min<i64>(220, std::max(40, opt.front_smem_kb)) * 1024;
    for (i64 l = 0; l < S.nlevels; l++) {
        bool fit = opt.fused_front != 0;
        i64 tiles = 0, need_max = 0;
        for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
            const i64 s = S.level_idx[t], ns = S.ns(s), nrow = S.nrow(s);
            const i64 need = ns * nrow > (1 << 20) ? (i64)1 << 40 : (i64)(FTILE + FSCRATCH + front_smem_doubles((int)nrow, (int)ns)) * (i64)sizeof(double);
            need_max = std::max(need_max, need);
            if (need > cap) fit = false;
            tiles += 1 + cdiv(std::max<i64>(0, nrow - std::min<i64>(ns, 2 * FB)), FB);
        }
        if (fit) {
            F.mode[l] = 2;
            F.front_smem_max = std::max(F.front_smem_max, (int)need_max);
        } else if (opt.fused_chain && tiles <= opt.chain_max_tiles) {
            F.mode[l] = 1;
            for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
                const i64 s = S.level_idx[t];
                F.sq_base[s] = F.sq_total;
                F.sq_total += cdiv(S.ns(s), 2 * FB);
            }
        }
    }
    return F;
}

// Panel factorization of a supernode (nrow x ns, column-major, in place), two-level blocking:
//   outer blocks of OB columns: left-looking update with ALL previous columns in one large-k DMMA GEMM,
//   inner blocks of NB=64 columns: single-CTA potrf + inverse, TRSM as a GEMM with the inverted block, then a k=64
//   trailing update confined to the outer block (bulk path), or one fused chain step per 128 columns (latency path).
// All supernodes of a level advance in lockstep, so one launch serves every front of the level.
void build_factor_plan(gmrf_b200_handle *h, Builder &B, const FusedInfo &F) {
    const Symbolic &S = h->S;
    Plan &plan = h->factor_plan;
    const i64 OB_bulk = std::max<i64>(NB, (i64)h->opt.outer_block / NB * NB);
    const i64 BBLK = (i64)NB * NB;
    std::vector<AsmItem> its;
    std::vector<PanelTask> pt;
    std::vector<GemmTask> gt, st;
    std::vector<ChainTask> ct;
    std::vector<FrontTask> ft;
    std::vector<AsmTile> ats;
    // the update part of the extend-add moves into the epilogue of the update-matrix product (needs the gather tables)
    const bool syrk_gather = h->opt.syrk_gather && h->opt.asm_gather && !B.naive;
    auto push_finalize = [&](i64 s, i64 k0, const double *sq) {
        const i64 ns = S.ns(s), ld = S.panel_ld[s];
        FinalizeTask f;
        f.sq = sq;
        f.P = h->d_Lx + S.panel_off[s] + k0 * ld + k0;
        f.inv0 = h->d_Linv + h->inv_base[s] + (k0 / NB) * BBLK;
        f.inv1 = f.inv0 + BBLK;
        f.ld = (int)ld;
        f.nb0 = (int)std::min<i64>(NB, ns - k0);
        f.nb1 = (int)std::max<i64>(0, std::min<i64>(NB, ns - k0 - NB));
        f.pad_ = 0;
        B.finalize.push_back(f);
    };
    for (i64 l = 0; l < S.nlevels; l++) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        const int mode = F.mode[(size_t)l];
        if (mode == 2) {
            for (const i64 *sp = sb; sp < se; sp++) {
                const i64 s = *sp;
                ft.push_back(FrontTask{(int)s, front_smem_doubles((int)S.nrow(s), (int)S.ns(s))});
                for (i64 k0 = 0; k0 < S.ns(s); k0 += 2 * NB) push_finalize(s, k0, nullptr);
            }
            int max_nb = 0;
            for (const i64 *sp = sb; sp < se; sp++) max_nb = std::max<int>(max_nb, (int)std::min<i64>(NB, S.ns(*sp)));
            B.add_front(plan, ft, max_nb);
            B.add_finalize(plan);
            h->n_front_launches++;
            continue;
        }
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp;
            if (S.child_ptr[s + 1] == S.child_ptr[s]) continue;
            if (h->opt.asm_gather && !B.naive) {
                const i64 cend = syrk_gather ? S.ns(s) : S.nrow(s);        // panel columns only / the whole front
                for (i64 c0 = 0; c0 < cend; c0 += AG_CW)
                    for (i64 r0 = c0 / AG_RH * AG_RH; r0 < S.nrow(s); r0 += AG_RH) ats.push_back(AsmTile{(int)s, (int)c0, (int)r0, syrk_gather ? 1 : 0});
            } else {
                for (i64 c0 = 0; c0 < S.nrow(s); c0 += ASM_CW) its.push_back(AsmItem{(int)s, (int)c0});
            }
        }
        B.add_items(plan, its, K_ASSEMBLE);
        B.add_asm_tiles(plan, ats);
        const i64 OB = mode == 1 ? 2 * NB : OB_bulk;
        i64 maxouter = 0;
        for (const i64 *sp = sb; sp < se; sp++) maxouter = std::max<i64>(maxouter, cdiv(S.ns(*sp), OB));
        for (i64 J = 0; J < maxouter; J++) {
            const i64 J0 = J * OB;
            if (J0 > 0) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                    if (J0 >= ns) continue;
                    i64 J1 = std::min(J0 + OB, ns);
                    double *P = h->d_Lx + S.panel_off[s];
                    GemmTask g;
                    g.A = P + J0; g.B = P + J0; g.C = P + J0 * ld + J0;
                    g.m = (int)(nrow - J0); g.n = (int)(J1 - J0); g.k = (int)J0;
                    g.lda = g.ldb = g.ldc = (int)ld;
                    g.flags = GEMM_LOWER; g.pad_ = 0;
                    gt.push_back(g);
                }
                B.add_gemm(plan, gt, 0, false, 1.0, /*allow_split=*/true);
            }
            if (mode == 1) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                    if (J0 >= ns) continue;
                    ChainTask c;
                    c.P = h->d_Lx + S.panel_off[s];
                    c.sq = h->d_sq + (F.sq_base[(size_t)s] + J) * (i64)(4 * BBLK);
                    c.ld = (int)ld; c.nrow = (int)nrow; c.k0 = (int)J0;
                    c.nb0 = (int)std::min<i64>(NB, ns - J0);
                    c.nb1 = (int)std::max<i64>(0, std::min<i64>(NB, ns - J0 - NB));
                    c.col0 = (int)(S.sfirst[s] + J0);
                    ct.push_back(c);
                    push_finalize(s, J0, c.sq);
                }
                B.add_chain(plan, ct);
                h->n_chain_launches++;
                continue;
            }
            for (i64 jj = 0; jj < OB / NB; jj++) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                    i64 k0 = J0 + jj * NB;
                    if (k0 >= ns) continue;
                    i64 nb = std::min<i64>(NB, ns - k0), k1 = k0 + nb, J1 = std::min(J0 + OB, ns);
                    double *P = h->d_Lx + S.panel_off[s];
                    double *inv = h->d_Linv + h->inv_base[s] + (k0 / NB) * BBLK;
                    pt.push_back(PanelTask{P + k0 * ld + k0, inv, (int)ld, (int)nb, (int)(S.sfirst[s] + k0), 0});
                    if (nrow > k1) {   // rows below the block: X = B * inv(L_kk)^T, in place (one 64-wide tile per row strip)
                        GemmTask g;
                        g.A = P + k0 * ld + k1; g.B = inv; g.C = P + k0 * ld + k1;
                        g.m = (int)(nrow - k1); g.n = (int)nb; g.k = (int)nb;
                        g.lda = g.ldc = (int)ld; g.ldb = (int)nb;
                        g.flags = GEMM_BETA0 | GEMM_ALPHA_POS; g.pad_ = 0;
                        st.push_back(g);
                    }
                    if (J1 > k1) {
                        GemmTask g;
                        g.A = P + k0 * ld + k1; g.B = P + k0 * ld + k1; g.C = P + k1 * ld + k1;
                        g.m = (int)(nrow - k1); g.n = (int)(J1 - k1); g.k = (int)nb;
                        g.lda = g.ldb = g.ldc = (int)ld;
                        g.flags = GEMM_LOWER; g.pad_ = 0;
                        gt.push_back(g);
                    }
                }
                B.add_panel(plan, pt);
                B.add_gemm(plan, st, 0, /*force_small=*/true, /*triangular operand: half the flops are algorithmic*/ 0.5);
                B.add_gemm(plan, gt, 0);
            }
        }
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
            if (nr == 0) continue;
            double *P = h->d_Lx + S.panel_off[s];
            GemmTask g;
            g.A = P + ns; g.B = P + ns;
            g.C = h->d_upd + S.upd_off[s];
            g.m = g.n = (int)nr; g.k = (int)ns;
            g.lda = g.ldb = (int)ld; g.ldc = S.upd_ld[s];
            const bool kids = S.child_ptr[s + 1] > S.child_ptr[s];
            g.flags = GEMM_LOWER | (kids ? (syrk_gather ? GEMM_GATHER : 0) : GEMM_BETA0);
            g.pad_ = (kids && syrk_gather) ? (int)s + 1 : 0;
            gt.push_back(g);
        }
        B.add_gemm(plan, gt, 0, false, 1.0, /*allow_split=*/h->opt.syrk_split != 0);
        B.add_finalize(plan);
    }
}

// Solve plans (few right-hand sides; kernels in solve_kernels.cuh). Per level: forward = children assembly + first
// block solve, then one launch per 64-column step of the level's supernode chains (lockstep); backward = gather
// phase (t_S = y_S - L21' x_R, last block solve), then one launch per block step, descending.
void build_solve_plans(gmrf_b200_handle *h, Builder &B) {
    const Symbolic &S = h->S;
    std::vector<int> lst;
    std::vector<FwdStepTask> ft;
    std::vector<BwdGatherTask> gt;
    std::vector<BwdReduceTask> rt;
    std::vector<BwdStepTask> bt;
    std::vector<WideStepTask> wt;
    std::vector<WideDiagTask> wd;
    auto is_wide = [&](i64 s) { return h->opt.wide_steps && S.ns(s) > SOLVE_WIDE_MIN; };
    const bool lookahead = h->opt.wide_steps >= 2;
    i64 part_off = 0;
    const i64 RCH = std::max(32, h->opt.bwd_row_chunk);
    const i64 BB = (i64)SOLVE_NB * SOLVE_NB;
    auto inv_ptr = [&](i64 s, i64 j) { return (const double *)(h->d_Linv + h->inv_base[s] + j * BB); };
    // forward: L y = b
    for (i64 l = 0; l < S.nlevels; l++) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        for (const i64 *sp = sb; sp < se; sp++)      // (one entry per FWD_ASM_ROWS front rows: fwd_assemble_x0_kernel)
            for (i64 c = 0; c < std::max<i64>(1, cdiv(S.nrow(*sp), FWD_ASM_ROWS)); c++) lst.push_back((int)*sp);
        B.add_superlist(h->fwd_plan, lst, K_FWD_ASM);
        i64 maxsteps = 0, maxwide = 0;
        for (const i64 *sp = sb; sp < se; sp++) {
            if (is_wide(*sp)) maxwide = std::max<i64>(maxwide, cdiv(S.ns(*sp), SOLVE_WB));
            else maxsteps = std::max<i64>(maxsteps, cdiv(S.ns(*sp), SOLVE_NB));
        }
        // long chains: per 256 columns one diagonal-block solve (one CTA per supernode), then one update of everything below
        for (i64 J = 0; J < maxwide; J++) {
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                i64 k0 = J * SOLVE_WB;
                if (!is_wide(s) || k0 >= ns) continue;
                i64 nbw = std::min<i64>(SOLVE_WB, ns - k0), k1 = k0 + nbw;
                const double *P = h->d_Lx + S.panel_off[s];
                // wide_steps = 2: the step's head CTA solves the NEXT diagonal block, only the first one needs a launch
                if (!lookahead || J == 0)
                    wd.push_back(WideDiagTask{P + k0 * ld + k0, inv_ptr(s, k0 / SOLVE_NB), h->d_y + S.sfirst[s] + k0, (int)ld, (int)nbw});
                if (nrow > k1) {
                    const i64 hrows = lookahead ? std::min<i64>(SOLVE_WB, ns - k1) : 0;
                    wt.push_back(WideStepTask{P + k0 * ld + k1, h->d_y + S.sfirst[s] + k0, h->d_y + S.sfirst[s] + k1, h->d_uvec + S.uvec_off[s],
                                              hrows > 0 ? inv_ptr(s, k1 / SOLVE_NB) : nullptr, (int)ld, (int)nbw, (int)(ns - k1),
                                              (int)(nrow - k1), (int)hrows, 0});
                }
            }
            B.add_wdiag(h->fwd_plan, wd, K_FWD_WDIAG);
            B.add_tiled(h->fwd_plan, wt, B.wstep, K_FWD_WSTEP, [](const WideStepTask &t) { return (i64)((t.hrows > 0) + cdiv(t.m - t.hrows, SOLVE_NB)); });
        }
        for (i64 j = 0; j < maxsteps; j++) {
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                i64 k0 = j * SOLVE_NB;
                if (k0 >= ns || is_wide(s)) continue;
                i64 nb = std::min<i64>(SOLVE_NB, ns - k0), k1 = k0 + nb;
                if (nrow == k1) continue;
                const double *P = h->d_Lx + S.panel_off[s];
                FwdStepTask t;
                t.L = P + k0 * ld + k1;
                t.x = h->d_y + S.sfirst[s] + k0;
                t.y = h->d_y + S.sfirst[s] + k1;
                t.u = h->d_uvec + S.uvec_off[s];
                t.ld = (int)ld; t.nb = (int)nb; t.ms = (int)(ns - k1); t.m = (int)(nrow - k1);
                t.nb_next = (int)std::min<i64>(SOLVE_NB, ns - k1);
                t.inv_next = t.nb_next > 0 ? inv_ptr(s, j + 1) : nullptr;
                t.pad_ = 0;
                ft.push_back(t);
            }
            B.add_tiled(h->fwd_plan, ft, B.fwd, K_FWD_STEP, [](const FwdStepTask &t) { return (i64)cdiv(t.m, SOLVE_NB); });
        }
    }
    // backward: L^T x = y
    for (i64 l = S.nlevels - 1; l >= 0; l--) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        i64 maxsteps = 0, maxwide = 0;
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
            i64 nblk = cdiv(ns, SOLVE_NB);
            if (is_wide(s)) maxwide = std::max<i64>(maxwide, cdiv(ns, SOLVE_WB));
            else maxsteps = std::max(maxsteps, nblk);
            BwdGatherTask t;
            t.L21 = h->d_Lx + S.panel_off[s] + ns;
            t.idx = h->d_rowidx + S.rowptr[s] + ns;
            t.inv_last = inv_ptr(s, nblk - 1);
            t.y = h->d_y + S.sfirst[s];
            t.ld = (int)ld; t.ns = (int)ns; t.nr = (int)nr; t.nb_last = (int)(ns - (nblk - 1) * SOLVE_NB);
            t.tile0 = nr > 0 ? 0 : (int)(nblk - 1);
            t.pad_ = is_wide(s) ? 1 : 0;          // long chain: the last block is solved by the wide diagonal kernel
            t.part = nullptr;
            if (nr > RCH) {
                // tall L21: row chunks write partial sums, a second pass folds them (fixed order) and finishes t_S
                const i64 nch = cdiv(nr, RCH);
                double *part = h->d_bwdpart + part_off;
                part_off += nch * BWD_PART_Q * ns;
                for (i64 k = 0; k < nch; k++) {
                    BwdGatherTask c = t;
                    const i64 r0 = k * RCH;
                    c.L21 = t.L21 + r0; c.idx = t.idx + r0;
                    c.nr = (int)std::min<i64>(RCH, nr - r0);
                    c.part = part + k * BWD_PART_Q * ns;
                    gt.push_back(c);
                }
                rt.push_back(BwdReduceTask{part, t.inv_last, t.y, (int)ns, (int)nch, t.nb_last, t.pad_});
            } else {
                gt.push_back(t);
            }
        }
        B.add_tiled(h->bwd_plan, gt, B.bwdg, K_BWD_GATHER, [](const BwdGatherTask &t) {
            return (i64)(cdiv(t.ns, SOLVE_NB) - t.tile0);
        });
        B.add_tiled(h->bwd_plan, rt, B.bwdr, K_BWD_REDUCE, [](const BwdReduceTask &t) { return (i64)cdiv(t.ns, SOLVE_NB); });
        for (i64 tJ = 0; tJ < maxwide; tJ++) {
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), ld = S.panel_ld[s];
                if (!is_wide(s)) continue;
                i64 J = cdiv(ns, SOLVE_WB) - 1 - tJ;
                if (J < 0) continue;
                i64 k0 = J * SOLVE_WB, nbw = std::min<i64>(SOLVE_WB, ns - k0);
                const double *P = h->d_Lx + S.panel_off[s];
                if (!lookahead || tJ == 0)
                    wd.push_back(WideDiagTask{P + k0 * ld + k0, inv_ptr(s, k0 / SOLVE_NB), h->d_y + S.sfirst[s] + k0, (int)ld, (int)nbw});
                if (k0 > 0)     // (k0 is a multiple of 256: the head's previous block is always a full one)
                    wt.push_back(WideStepTask{P + k0, h->d_y + S.sfirst[s] + k0, h->d_y + S.sfirst[s], nullptr,
                                              lookahead ? inv_ptr(s, (k0 - SOLVE_WB) / SOLVE_NB) : nullptr, (int)ld, (int)nbw, 0, (int)k0,
                                              lookahead ? SOLVE_WB : 0, 0});
            }
            B.add_wdiag(h->bwd_plan, wd, K_BWD_WDIAG);
            B.add_tiled(h->bwd_plan, wt, B.wstep, K_BWD_WSTEP, [](const WideStepTask &t) { return (i64)((t.hrows > 0) + cdiv(t.m - t.hrows, SOLVE_NB)); });
        }
        for (i64 tt = 0; tt + 1 < maxsteps; tt++) {
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), ld = S.panel_ld[s];
                i64 nblk = cdiv(ns, SOLVE_NB);
                i64 j = nblk - 1 - tt;
                if (j < 1 || is_wide(s)) continue;
                i64 k0 = j * SOLVE_NB, nb = std::min<i64>(SOLVE_NB, ns - k0);
                BwdStepTask t;
                t.L = h->d_Lx + S.panel_off[s] + k0;
                t.inv_prev = inv_ptr(s, j - 1);
                t.x = h->d_y + S.sfirst[s] + k0;
                t.y = h->d_y + S.sfirst[s];
                t.ld = (int)ld; t.nb = (int)nb; t.ncols = (int)k0; t.pad_ = 0;
                bt.push_back(t);
            }
            B.add_tiled(h->bwd_plan, bt, B.bwds, K_BWD_STEP, [](const BwdStepTask &t) { return (i64)(t.ncols / SOLVE_NB); });
        }
    }
}

// Wide right-hand-side sweeps (blocks of W = 64, 128 or 256 columns): every step is a DMMA GEMM on the block.
// Two-level blocking like the factorization: inside an outer block of OB columns the 64-column diagonal blocks are
// applied through their inverses (one 64 x 64-tile GEMM each, in place) with small k = 64 updates confined to the outer
// block; the rows below / columns left of the outer block get ONE k = OB update (forward: right-looking over rows,
// backward: right-looking over columns, exactly the dependency structure of the few-RHS kernels).
void build_multi_plans(gmrf_b200_handle *h, Builder &B, int W, Plan &fwd_plan, Plan &bwd_plan) {
    const Symbolic &S = h->S;
    const i64 OB = std::max<i64>(NB, (i64)h->opt.outer_block / NB * NB);
    const i64 BBLK = (i64)NB * NB;
    const int ldy = (int)S.n, ldu = (int)S.uvec_total;
    std::vector<int> lst;
    std::vector<GemmTask> st, gt;
    std::vector<RowGatherTask> rg;
    auto inv_ptr = [&](i64 s, i64 k0) { return (const double *)(h->d_Linv + h->inv_base[s] + (k0 / NB) * BBLK); };
    auto task = [](const double *A, int lda, const double *Bm, int ldb, double *C, int ldc, i64 m, i64 n, i64 k, int flags) {
        GemmTask g;
        g.A = A; g.B = Bm; g.C = C; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
        g.m = (int)m; g.n = (int)n; g.k = (int)k; g.pad_ = 0;
        g.flags = flags | GEMM_A_CONST;      // A is always a piece of the factor (panel or inverted diagonal block)
        return g;
    };
    // ---- forward: L y = b ------------------------------------------------------------------------------------
    for (i64 l = 0; l < S.nlevels; l++) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        i64 maxouter = 0;
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp;
            if (S.nr(s) > 0 || S.child_ptr[s + 1] > S.child_ptr[s]) lst.push_back((int)s);
            maxouter = std::max<i64>(maxouter, cdiv(S.ns(s), OB));
        }
        B.add_superlist(fwd_plan, lst, K_FWD_ASM_M);
        for (i64 J = 0; J < maxouter; J++) {
            for (i64 jj = 0; jj < OB / NB; jj++) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), ld = S.panel_ld[s];
                    i64 J0 = J * OB, k0 = J0 + jj * NB;
                    if (k0 >= ns) continue;
                    i64 J1 = std::min(J0 + OB, ns), nb = std::min<i64>(NB, ns - k0), k1 = k0 + nb;
                    const double *P = h->d_Lx + S.panel_off[s];
                    double *ys = h->d_ym + S.sfirst[s];
                    // x_K = inv(L_KK) y_K : C[nb x W] = inv * y_K, in place (one tile)
                    st.push_back(task(inv_ptr(s, k0), (int)nb, ys + k0, ldy, ys + k0, ldy, nb, W, nb, GEMM_BETA0 | GEMM_ALPHA_POS));
                    // y[k1:J1] -= L[k1:J1, K] x_K
                    if (J1 > k1) gt.push_back(task(P + k0 * ld + k1, (int)ld, ys + k0, ldy, ys + k1, ldy, J1 - k1, W, nb, 0));
                }
                B.add_gemm(fwd_plan, st, 1, true, 0.5);
                B.add_gemm(fwd_plan, gt, 1);
            }
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
                i64 J0 = J * OB;
                if (J0 >= ns) continue;
                i64 J1 = std::min(J0 + OB, ns);
                const double *P = h->d_Lx + S.panel_off[s];
                double *ys = h->d_ym + S.sfirst[s];
                if (ns > J1) gt.push_back(task(P + J0 * ld + J1, (int)ld, ys + J0, ldy, ys + J1, ldy, ns - J1, W, J1 - J0, 0));
                if (nr > 0) gt.push_back(task(P + J0 * ld + ns, (int)ld, ys + J0, ldy, h->d_um + S.uvec_off[s], ldu, nr, W, J1 - J0, 0));
            }
            B.add_gemm(fwd_plan, gt, 1);
        }
    }
    // ---- backward: L^T x = y ---------------------------------------------------------------------------------
    for (i64 l = S.nlevels - 1; l >= 0; l--) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        i64 maxouter = 0;
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
            maxouter = std::max<i64>(maxouter, cdiv(ns, OB));
            if (nr == 0) continue;
            rg.push_back(RowGatherTask{h->d_rowidx + S.rowptr[s] + ns, h->d_um + S.uvec_off[s], (int)nr, 0});
            // y_S -= L21^T u_s
            gt.push_back(task(h->d_Lx + S.panel_off[s] + ns, (int)ld, h->d_um + S.uvec_off[s], ldu, h->d_ym + S.sfirst[s], ldy, ns, W, nr, 0));
        }
        B.add_tiled(bwd_plan, rg, B.rowgather, K_ROWS_GATHER, [](const RowGatherTask &t) { return (i64)cdiv(t.nr, 256); });
        B.add_gemm(bwd_plan, gt, 2);
        for (i64 tJ = 0; tJ < maxouter; tJ++) {
            for (i64 tj = 0; tj < OB / NB; tj++) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), ld = S.panel_ld[s];
                    i64 J = cdiv(ns, OB) - 1 - tJ;
                    if (J < 0) continue;
                    i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                    i64 jj = cdiv(J1 - J0, NB) - 1 - tj;
                    if (jj < 0) continue;
                    i64 k0 = J0 + jj * NB, nb = std::min<i64>(NB, J1 - k0);
                    const double *P = h->d_Lx + S.panel_off[s];
                    double *ys = h->d_ym + S.sfirst[s];
                    // x_K = inv(L_KK)^T t_K : Aop[c][r] = inv[r + c*nb]
                    st.push_back(task(inv_ptr(s, k0), (int)nb, ys + k0, ldy, ys + k0, ldy, nb, W, nb, GEMM_BETA0 | GEMM_ALPHA_POS));
                    // t[J0:k0] -= L[K, J0:k0]^T x_K
                    if (k0 > J0) gt.push_back(task(P + J0 * ld + k0, (int)ld, ys + k0, ldy, ys + J0, ldy, k0 - J0, W, nb, 0));
                }
                B.add_gemm(bwd_plan, st, 2, true, 0.5);
                B.add_gemm(bwd_plan, gt, 2);
            }
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp, ns = S.ns(s), ld = S.panel_ld[s];
                i64 J = cdiv(ns, OB) - 1 - tJ;
                if (J < 1) continue;
                i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                const double *P = h->d_Lx + S.panel_off[s];
                double *ys = h->d_ym + S.sfirst[s];
                // t[0:J0] -= L[J, 0:J0]^T x_J
                gt.push_back(task(P + J0, (int)ld, ys + J0, ldy, ys, ldy, J0, W, J1 - J0, 0));
            }
            B.add_gemm(bwd_plan, gt, 2);
        }
    }
}

// Selected inversion (Takahashi), per supernode s with panel L = [L11; L21], W = Z[R,R] gathered from the parent:
//   T' = -W L21 ;  G = I - L21^T T' = I + L21^T W L21 ;  [H; Z_RS] = [G; T'] L11^-1 ;  Z_SS = H^T L11^-1
// (Z_SS = L11^-T (I + L21^T W L21) L11^-1, Z_RS = -W L21 L11^-1.) Right solves with the 64-column diagonal blocks are
// GEMMs with the inverted blocks the factorization left in d_Linv (in place: one 64-wide tile owns a whole row strip).
//
// Root supernodes of the top level (nr == 0, nothing to gather) take a cheaper route that exploits triangularity:
//   H = L11^-1 is built as a LOWER TRIANGULAR matrix in the (then empty) update pool, restricted to its nonzero rows
//   and, per row block, to its nonzero k-range (trtri, ~ns^3/3 flops instead of ns^3), and
//   Z_SS = H^T H is one triangular product written straight into the Z panel (lauum, ~ns^3/3 instead of ns^3).
// At 1 M dofs the root (ns = 30,502) was half of the selected-inversion time with the generic path.
void build_selinv_plan(gmrf_b200_handle *h, Builder &B) {
    const Symbolic &S = h->S;
    Plan &plan = h->selinv_plan;
    std::vector<AsmItem> its;
    std::vector<GemmTask> gt, st;
    std::vector<TransTask> tr;
    const i64 OB = std::max<i64>(NB, (i64)h->opt.outer_block / NB * NB);
    const i64 BBLK = (i64)NB * NB;
    auto inv_ptr = [&](i64 s, i64 k0) { return (const double *)(h->d_Linv + h->inv_base[s] + (k0 / NB) * BBLK); };
    // X[row0:, K] := X[row0:, K] * inv(L_KK) (in place), K = [k0, k0 + nb)
    auto push_block_solve = [&](double *X, i64 ldx, i64 s, i64 k0, i64 nb, i64 m) {
        if (m <= 0) return;
        GemmTask g;
        g.A = X; g.lda = (int)ldx;
        g.B = inv_ptr(s, k0); g.ldb = (int)nb;             // Bop[j][kk] = inv[kk + j*nb]  (k-contiguous)
        g.C = X; g.ldc = (int)ldx;
        g.m = (int)m; g.n = (int)nb; g.k = (int)nb;
        g.flags = GEMM_BETA0 | GEMM_ALPHA_POS; g.pad_ = 0;
        st.push_back(g);
    };
    std::vector<char> fast(S.nsuper, 0);
    std::vector<i64> scratch_off(S.nsuper, 0);
    if (S.nlevels > 0 && h->opt.selinv_fast_root) {
        i64 l = S.nlevels - 1, used = 0;
        const i64 pool = std::max(S.upd_total, S.zw_total);
        for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
            i64 s = S.level_idx[t], ns = S.ns(s);
            i64 need = (i64)S.panel_ld[s] * ns;
            if (S.nr(s) == 0 && ns >= 4 * NB && used + need <= pool) { fast[s] = 1; scratch_off[s] = used; used += need; h->n_fast_roots++; }
        }
    }
    for (i64 l = S.nlevels - 1; l >= 0; l--) {
        const i64 *sb = S.level_idx.data() + S.level_ptr[l], *se = S.level_idx.data() + S.level_ptr[l + 1];
        // ---- fast roots -------------------------------------------------------------------------------------
        {
            const i64 RBK = 2048;   // row block of the triangular products (k-range is cut per row block)
            i64 maxouter = 0;
            for (const i64 *sp = sb; sp < se; sp++) {
                i64 s = *sp;
                if (!fast[s]) continue;
                i64 ns = S.ns(s), ld = S.panel_ld[s];
                maxouter = std::max<i64>(maxouter, cdiv(ns, OB));
                GemmTask g;   // scratch := I
                g.A = h->d_Lx; g.B = h->d_Lx; g.lda = g.ldb = 1;
                g.C = h->d_zw + scratch_off[s]; g.ldc = (int)ld;
                g.m = g.n = (int)ns; g.k = 0;
                g.flags = GEMM_BETA0 | GEMM_ADD_I; g.pad_ = 0;
                gt.push_back(g);
            }
            B.add_gemm(plan, gt, 1, true);
            for (i64 tJ = 0; tJ < maxouter; tJ++) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp;
                    if (!fast[s]) continue;
                    i64 ns = S.ns(s), ld = S.panel_ld[s];
                    i64 J = cdiv(ns, OB) - 1 - tJ;
                    if (J < 0) continue;
                    i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                    double *X = h->d_zw + scratch_off[s];
                    const double *P = h->d_Lx + S.panel_off[s];
                    for (i64 a = J1; a < ns; a += RBK) {     // X[a:b, J] -= X[a:b, J1:b] L[J1:b, J]
                        i64 b = std::min(a + RBK, ns);
                        GemmTask g;
                        g.A = X + J1 * ld + a; g.lda = (int)ld;
                        g.B = P + J0 * ld + J1; g.ldb = (int)ld;
                        g.C = X + J0 * ld + a; g.ldc = (int)ld;
                        g.m = (int)(b - a); g.n = (int)(J1 - J0); g.k = (int)(b - J1);
                        g.flags = 0; g.pad_ = 0;
                        gt.push_back(g);
                    }
                }
                B.add_gemm(plan, gt, 1, false, 1.0, /*allow_split=*/true);
                for (i64 tj = 0; tj < OB / NB; tj++) {
                    for (const i64 *sp = sb; sp < se; sp++) {
                        i64 s = *sp;
                        if (!fast[s]) continue;
                        i64 ns = S.ns(s), ld = S.panel_ld[s];
                        i64 J = cdiv(ns, OB) - 1 - tJ;
                        if (J < 0) continue;
                        i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                        i64 jj = cdiv(J1 - J0, NB) - 1 - tj;
                        if (jj < 0) continue;
                        i64 k0 = J0 + jj * NB, nb = std::min<i64>(NB, J1 - k0), k1 = k0 + nb;
                        double *X = h->d_zw + scratch_off[s];
                        const double *P = h->d_Lx + S.panel_off[s];
                        if (J1 > k1) {                        // X[k1:, K] -= X[k1:, k1:J1] L[k1:J1, K]
                            GemmTask g;
                            g.A = X + k1 * ld + k1; g.lda = (int)ld;
                            g.B = P + k0 * ld + k1; g.ldb = (int)ld;
                            g.C = X + k0 * ld + k1; g.ldc = (int)ld;
                            g.m = (int)(ns - k1); g.n = (int)nb; g.k = (int)(J1 - k1);
                            g.flags = 0; g.pad_ = 0;
                            gt.push_back(g);
                        }
                        push_block_solve(X + k0 * ld + k0, ld, s, k0, nb, ns - k0);
                    }
                    B.add_gemm(plan, gt, 1);
                    B.add_gemm(plan, st, 1, true, 0.5);
                }
            }
            for (const i64 *sp = sb; sp < se; sp++) {       // Z_SS (lower) = H^T H, one task per row block
                i64 s = *sp;
                if (!fast[s]) continue;
                i64 ns = S.ns(s), ld = S.panel_ld[s];
                const double *X = h->d_zw + scratch_off[s];
                double *Z = h->d_Zx + S.panel_off[s];
                for (i64 a = 0; a < ns; a += RBK) {
                    i64 b = std::min(a + RBK, ns);
                    GemmTask g;
                    g.A = X + a * ld + a; g.lda = (int)ld;     // Aop[i][kk] = H[a + kk, a + i]
                    g.B = X + a; g.ldb = (int)ld;              // Bop[j][kk] = H[a + kk, j]
                    g.C = Z + a; g.ldc = (int)ld;
                    g.m = (int)(b - a); g.n = (int)b; g.k = (int)(ns - a);
                    g.flags = GEMM_BETA0 | GEMM_ALPHA_POS; g.pad_ = 0;
                    gt.push_back(g);
                }
            }
            B.add_gemm(plan, gt, 2);
        }
        // ---- generic path -----------------------------------------------------------------------------------
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp;
            if (fast[s]) continue;
            for (i64 c0 = 0; c0 < S.nr(s); c0 += ASM_CW) its.push_back(AsmItem{(int)s, (int)c0});
        }
        B.add_items(plan, its, K_GATHER);
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
            if (nr == 0 || fast[s]) continue;
            GemmTask g;
            g.A = h->d_zw + S.zw_off[s]; g.lda = S.upd_ld[s];
            g.B = h->d_Lx + S.panel_off[s] + ns; g.ldb = (int)ld;
            g.C = h->d_Zx + S.panel_off[s] + ns; g.ldc = (int)ld;
            g.m = (int)nr; g.n = (int)ns; g.k = (int)nr;
            g.flags = GEMM_BETA0; g.pad_ = 0;
            gt.push_back(g);
        }
        B.add_gemm(plan, gt, 1);
        for (const i64 *sp = sb; sp < se; sp++) {
            i64 s = *sp, ns = S.ns(s), nr = S.nr(s), ld = S.panel_ld[s];
            if (fast[s]) continue;
            GemmTask g;
            g.A = h->d_Lx + S.panel_off[s] + ns; g.lda = (int)ld;
            g.B = h->d_Zx + S.panel_off[s] + ns; g.ldb = (int)ld;
            g.C = h->d_Zx + S.panel_off[s]; g.ldc = (int)ld;
            g.m = (int)ns; g.n = (int)ns; g.k = (int)nr;
            g.flags = GEMM_BETA0 | GEMM_ADD_I; g.pad_ = 0;
            gt.push_back(g);
        }
        B.add_gemm(plan, gt, 2);
        // X := X L11^-1 by block back-substitution, two-level blocking (outer OB, inner NB), lockstep over the level
        i64 maxouter = 0;
        for (const i64 *sp = sb; sp < se; sp++)
            if (!fast[*sp]) maxouter = std::max<i64>(maxouter, cdiv(S.ns(*sp), OB));
        for (int pass = 0; pass < 2; pass++) {
            // pass 0: all nrow rows of [G; T'];  pass 1: the (transposed) ns x ns block only
            if (pass == 1) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp;
                    if (S.ns(s) > 1 && !fast[s]) tr.push_back(TransTask{h->d_Zx + S.panel_off[s], (int)S.ns(s), S.panel_ld[s]});
                }
                B.add_trans(plan, tr);
            }
            for (i64 tJ = 0; tJ < maxouter; tJ++) {
                for (const i64 *sp = sb; sp < se; sp++) {
                    i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                    if (fast[s]) continue;
                    i64 J = cdiv(ns, OB) - 1 - tJ;
                    if (J < 0) continue;
                    i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                    if (J1 >= ns) continue;
                    i64 m = pass == 0 ? nrow : ns;
                    double *Z = h->d_Zx + S.panel_off[s];
                    const double *P = h->d_Lx + S.panel_off[s];
                    GemmTask g;
                    g.A = Z + J1 * ld; g.lda = (int)ld;             // X[:, J1:ns]
                    g.B = P + J0 * ld + J1; g.ldb = (int)ld;        // L11[J1:ns, J0:J1]  (k-contiguous)
                    g.C = Z + J0 * ld; g.ldc = (int)ld;
                    g.m = (int)m; g.n = (int)(J1 - J0); g.k = (int)(ns - J1);
                    g.flags = 0; g.pad_ = 0;
                    gt.push_back(g);
                }
                B.add_gemm(plan, gt, 1, false, 1.0, /*allow_split=*/true);
                for (i64 tj = 0; tj < OB / NB; tj++) {
                    for (const i64 *sp = sb; sp < se; sp++) {
                        i64 s = *sp, ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
                        if (fast[s]) continue;
                        i64 J = cdiv(ns, OB) - 1 - tJ;
                        if (J < 0) continue;
                        i64 J0 = J * OB, J1 = std::min(J0 + OB, ns);
                        i64 jj = cdiv(J1 - J0, NB) - 1 - tj;
                        if (jj < 0) continue;
                        i64 k0 = J0 + jj * NB, nb = std::min<i64>(NB, J1 - k0), k1 = k0 + nb;
                        i64 m = pass == 0 ? nrow : ns;
                        double *Z = h->d_Zx + S.panel_off[s];
                        const double *P = h->d_Lx + S.panel_off[s];
                        if (J1 > k1) {
                            GemmTask g;
                            g.A = Z + k1 * ld; g.lda = (int)ld;          // X[:, k1:J1]
                            g.B = P + k0 * ld + k1; g.ldb = (int)ld;     // L11[k1:J1, K]
                            g.C = Z + k0 * ld; g.ldc = (int)ld;
                            g.m = (int)m; g.n = (int)nb; g.k = (int)(J1 - k1);
                            g.flags = 0; g.pad_ = 0;
                            gt.push_back(g);
                        }
                        push_block_solve(Z + k0 * ld, ld, s, k0, nb, m);
                    }
                    B.add_gemm(plan, gt, 1);
                    B.add_gemm(plan, st, 1, true, 0.5);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Launch dispatch
// ------------------------------------------------------------------------------------------------
// Kernel launch with (optionally) the programmatic-stream-serialization attribute: the kernel may become resident while
// its predecessor in the stream still runs and synchronizes itself with griddepcontrol.wait (kernels.cuh).
template <typename... KA, typename... A>
static inline void launch_k(bool pdl, void (*kern)(KA...), dim3 grid, int block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}

template <int BM, int BN, int WGM, int WGN, bool TA, bool TB>
void launch_gemm_t(const GemmTask *tasks, const int *prefix, int ntasks, int grid, cudaStream_t st, int lanes, long long bstride,
                   const GatherCtx *gctx = nullptr, bool pdl = false) {
    if (!TA && !TB && gctx)      // update-matrix products whose epilogue gathers the children's contributions
        launch_k(pdl, gemm_dmma_kernel<BM, BN, WGM, WGN, false, false, 16, 3, true>, dim3(grid, lanes), WGM * WGN * 32, gemm_smem_bytes<BM, BN>(), st,
                 tasks, prefix, ntasks, bstride, gctx);
    else
        launch_k(pdl, gemm_dmma_kernel<BM, BN, WGM, WGN, TA, TB>, dim3(grid, lanes), WGM * WGN * 32, gemm_smem_bytes<BM, BN, 16, 3, TA, TB>(), st,
                 tasks, prefix, ntasks, bstride, (const GatherCtx *)nullptr);
}

// Opt in to > 48 KB dynamic shared memory for every GEMM instantiation (per device; must run outside stream capture).
template <int BM, int BN, int WGM, int WGN>
cudaError_t configure_gemm_tile() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WGM, WGN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>()))) return e;
    if ((e = cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WGM, WGN, false, false, 16, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>()))) return e;
    if ((e = cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WGM, WGN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN, 16, 3, false, true>()))) return e;
    if ((e = cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WGM, WGN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN, 16, 3, true, true>()))) return e;
    return cudaSuccess;
}
cudaError_t configure_kernels(int front_smem = 0) {
    cudaError_t e;
    if ((e = configure_gemm_tile<128, 64, 2, 2>())) return e;
    if ((e = configure_gemm_tile<64, 64, 2, 2>())) return e;
    if ((e = cudaFuncSetAttribute(chain_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(chain_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM_BYTES))) return e;
    // (static + dynamic shared memory of the wider forward step kernels exceeds 48 KB)
    if ((e = cudaFuncSetAttribute(fwd_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(fwd_step_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEP_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(assemble_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AG_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(chain_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FINALIZE_SMEM_BYTES))) return e;
    if ((e = cudaFuncSetAttribute(potrf_inv64_la_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_LA_SMEM_BYTES))) return e;
    // (per device, sticky: handles of one process may need different sizes -> always opt in to the cap)
    const int fs = std::max(front_smem, 220 * 1024);
    if ((e = cudaFuncSetAttribute(front_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs))) return e;
    if ((e = cudaFuncSetAttribute(front_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs))) return e;
    return cudaSuccess;
}

template <bool TA, bool TB>
void launch_gemm(bool large, bool naive, const GemmTask *tasks, const int *prefix, int ntasks, int grid, cudaStream_t st,
                 int lanes = 1, long long bstride = 0, const GatherCtx *gctx = nullptr, bool pdl = false) {
    if (naive) gemm_naive_kernel<TA, TB><<<dim3(grid, lanes), 256, 0, st>>>(tasks, prefix, ntasks, bstride);
    else if (large) launch_gemm_t<128, 64, 2, 2, TA, TB>(tasks, prefix, ntasks, grid, st, lanes, bstride, gctx, pdl);
    else launch_gemm_t<64, 64, 2, 2, TA, TB>(tasks, prefix, ntasks, grid, st, lanes, bstride, gctx, pdl);
}

struct TableSet {
    const GemmTask *gemm;
    const AsmItem *items;
    const int *prefix;
    const SplitTask *split;
    const int *superlist = nullptr;         // solve phases: supernode lists
    const RowGatherTask *rowgather = nullptr;
    double *y = nullptr, *u = nullptr;      // solve phases: permuted work array and update-vector pool
    int lanes = 1;                          // factorization: lanes advanced per launch (grid.y)
    bool pdl = false;                       // launch the kernels that synchronize themselves (griddepcontrol.wait) programmatically
};

void run_launch(gmrf_b200_handle *h, const Launch &L, const TableSet &T, int nrhs, cudaStream_t stream_override = nullptr) {
    cudaStream_t st = stream_override ? stream_override : h->stream;
    const bool naive = h->opt.naive_kernels != 0;
    const int *pf = L.prefix_off >= 0 ? T.prefix + L.prefix_off : nullptr;
    const long long bstride = T.lanes > 1 ? h->arena_bytes : 0;
    if (L.grid <= 0) return;
    switch (L.kind) {
        case K_ASSEMBLE:
            launch_k(T.pdl, assemble_kernel, dim3(L.grid, T.lanes), 256, 0, st, T.items + L.task_off, h->d_meta, h->d_child, h->d_relidx, h->d_Lx, h->d_upd, bstride);
            break;
        case K_ASSEMBLE_G:
            launch_k(T.pdl, assemble_gather_kernel, dim3(L.grid, T.lanes), 256, AG_SMEM_BYTES, st, h->d_asmtiles + L.task_off, h->d_meta, h->d_child,
                     h->d_relidx, h->d_relpos, h->d_relpos_off, h->d_Lx, h->d_upd, bstride);
            break;
        case K_CHAIN:
            if (h->opt.panel_blocked >= 2)
                launch_k(T.pdl, chain_step_kernel<true>, dim3(L.grid, T.lanes), 256, CHAIN_SMEM_BYTES, st, h->d_chain + L.task_off, pf, (int)L.ntasks, h->d_fail, bstride);
            else
                launch_k(T.pdl, chain_step_kernel<false>, dim3(L.grid, T.lanes), 256, CHAIN_SMEM_BYTES, st, h->d_chain + L.task_off, pf, (int)L.ntasks, h->d_fail, bstride);
            break;
        case K_FRONT:
            if (h->opt.panel_blocked)
                launch_k(T.pdl, front_small_kernel<true>, dim3(L.grid, T.lanes), 256, (size_t)L.aux, st, h->d_front + L.task_off, h->d_meta, h->d_child,
                         h->d_relidx, h->d_Lx, h->d_upd, h->d_fail, bstride);
            else
                launch_k(T.pdl, front_small_kernel<false>, dim3(L.grid, T.lanes), 256, (size_t)L.aux, st, h->d_front + L.task_off, h->d_meta, h->d_child,
                         h->d_relidx, h->d_Lx, h->d_upd, h->d_fail, bstride);
            break;
        case K_FINALIZE:
            chain_finalize_kernel<<<dim3(2 * L.grid, T.lanes), 256, FINALIZE_SMEM_BYTES, st>>>(h->d_final + L.task_off, bstride);
            break;
        case K_PANEL:
            switch (L.aux) {
                case 8: launch_k(T.pdl, potrf_inv_kernel<8>, dim3(L.grid, T.lanes), 256, 0, st, h->d_panel + L.task_off, h->d_fail, bstride); break;
                case 16: launch_k(T.pdl, potrf_inv_kernel<16>, dim3(L.grid, T.lanes), 256, 0, st, h->d_panel + L.task_off, h->d_fail, bstride); break;
                case 32: launch_k(T.pdl, potrf_inv_kernel<32>, dim3(L.grid, T.lanes), 256, 0, st, h->d_panel + L.task_off, h->d_fail, bstride); break;
                default:
                    if (h->opt.potrf_lookahead)
                        launch_k(T.pdl, potrf_inv64_la_kernel, dim3(L.grid, T.lanes), 256, POTRF_LA_SMEM_BYTES, st, h->d_panel + L.task_off, h->d_fail, bstride);
                    else
                        launch_k(T.pdl, potrf_inv64_kernel, dim3(L.grid, T.lanes), 256, 0, st, h->d_panel + L.task_off, h->d_fail, bstride);
                    break;
            }
            break;
        case K_GEMM_NN_S: case K_GEMM_NN_L:
            launch_gemm<false, false>(L.kind == K_GEMM_NN_L, naive, T.gemm + L.task_off, pf, L.ntasks, L.grid, st, T.lanes, bstride,
                                      L.aux ? h->d_gctx : nullptr, T.pdl);
            break;
        case K_GEMM_NT_S: case K_GEMM_NT_L:
            launch_gemm<false, true>(L.kind == K_GEMM_NT_L, naive, T.gemm + L.task_off, pf, L.ntasks, L.grid, st, T.lanes, bstride, nullptr, T.pdl); break;
        case K_GEMM_TT_S: case K_GEMM_TT_L:
            launch_gemm<true, true>(L.kind == K_GEMM_TT_L, naive, T.gemm + L.task_off, pf, L.ntasks, L.grid, st, T.lanes, bstride, nullptr, T.pdl); break;
        case K_SPLIT_REDUCE:
            launch_k(T.pdl, splitk_reduce_kernel, dim3(L.grid, T.lanes), 256, 0, st, T.split + L.task_off, pf, (int)L.ntasks, bstride);
            break;
        case K_GATHER:
            selinv_gather_kernel<<<L.grid, 256, 0, st>>>(T.items + L.task_off, h->d_meta, h->d_relidx, h->d_Zx, h->d_zw);
            break;
        case K_TRANSPOSE:
            transpose_inplace_kernel<<<L.grid, 256, 0, st>>>(h->d_trans + L.task_off, pf, L.ntasks);
            break;
        case K_FWD_ASM: {
            dim3 g(L.grid, nrhs);
            launch_k(T.pdl, fwd_assemble_x0_kernel, g, 256, 0, st, h->d_superlist + L.task_off, h->d_meta, h->d_child, h->d_relidx, h->d_Linv,
                     h->d_invbase, h->d_y, h->S.n, h->d_uvec, h->S.uvec_total);
            break;
        }
        case K_FWD_ASM_M: {
            dim3 g(L.grid, nrhs / MULTI_QB);
            fwd_assemble_multi_kernel<<<g, 256, 0, st>>>(T.superlist + L.task_off, h->d_meta, h->d_child, h->d_relidx,
                                                         T.y, h->S.n, T.u, h->S.uvec_total);
            break;
        }
        case K_ROWS_GATHER: {
            dim3 g(L.grid, nrhs / MULTI_QB);
            rows_gather_kernel<<<g, 256, 0, st>>>(T.rowgather + L.task_off, pf, L.ntasks, T.y, h->S.n, h->S.uvec_total);
            break;
        }
#define SOLVE_RB_DISPATCH(KERNEL, ...)                                                         \
    do {                                                                                       \
        if (nrhs <= 1) KERNEL<1><<<L.grid, 256, 0, st>>>(__VA_ARGS__);                         \
        else if (nrhs <= 2) KERNEL<2><<<L.grid, 256, 0, st>>>(__VA_ARGS__);                    \
        else if (nrhs <= 4) KERNEL<4><<<L.grid, 256, 0, st>>>(__VA_ARGS__);                    \
        else KERNEL<8><<<L.grid, 256, 0, st>>>(__VA_ARGS__);                                   \
    } while (0)
// the kernels of the few-RHS sweeps: programmatic dependent launch (solve_kernels.cuh) -- the next step's CTAs may start
// loading their factor tiles while this one finishes
#define SOLVE_RB_DISPATCH_PDL(SMEM, KERNEL, ...)                                              \
    do {                                                                                       \
        const bool pdl = T.pdl;                                                                \
        if (nrhs <= 1) launch_k(pdl, KERNEL<1>, dim3(L.grid), 256, SMEM, st, __VA_ARGS__);     \
        else if (nrhs <= 2) launch_k(pdl, KERNEL<2>, dim3(L.grid), 256, SMEM, st, __VA_ARGS__); \
        else if (nrhs <= 4) launch_k(pdl, KERNEL<4>, dim3(L.grid), 256, SMEM, st, __VA_ARGS__); \
        else launch_k(pdl, KERNEL<8>, dim3(L.grid), 256, SMEM, st, __VA_ARGS__);               \
    } while (0)
        case K_FWD_STEP:
            SOLVE_RB_DISPATCH_PDL(STEP_SMEM_BYTES, fwd_step_kernel, (const FwdStepTask *)(h->d_fwd + L.task_off), pf, (int)L.ntasks, nrhs, (long long)h->S.n, (long long)h->S.uvec_total);
            break;
        case K_BWD_GATHER:
            SOLVE_RB_DISPATCH_PDL(0, bwd_gather_kernel, h->d_bwdg + L.task_off, pf, (int)L.ntasks, nrhs, h->d_y, (long long)h->S.n);
            break;
        case K_BWD_REDUCE:
            SOLVE_RB_DISPATCH_PDL(0, bwd_reduce_kernel, h->d_bwdr + L.task_off, pf, (int)L.ntasks, nrhs, (long long)h->S.n);
            break;
        case K_BWD_STEP:
            SOLVE_RB_DISPATCH_PDL(STEP_SMEM_BYTES, bwd_step_kernel, (const BwdStepTask *)(h->d_bwds + L.task_off), pf, (int)L.ntasks, nrhs, (long long)h->S.n);
            break;
        case K_FWD_WSTEP:
            SOLVE_RB_DISPATCH(fwd_wide_step_kernel, h->d_wstep + L.task_off, pf, L.ntasks, nrhs, (long long)h->S.n, (long long)h->S.uvec_total);
            break;
        case K_BWD_WSTEP:
            SOLVE_RB_DISPATCH(bwd_wide_step_kernel, h->d_wstep + L.task_off, pf, L.ntasks, nrhs, (long long)h->S.n);
            break;
        case K_FWD_WDIAG:
            SOLVE_RB_DISPATCH(fwd_wide_diag_kernel, h->d_wdiag + L.task_off, nrhs, (long long)h->S.n);
            break;
        case K_BWD_WDIAG:
            SOLVE_RB_DISPATCH(bwd_wide_diag_kernel, h->d_wdiag + L.task_off, nrhs, (long long)h->S.n);
            break;
#undef SOLVE_RB_DISPATCH
#undef SOLVE_RB_DISPATCH_PDL
    }
}

int check_launch(gmrf_b200_handle *h, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        h->err = std::string("kernel launch failed in ") + what + ": " + cudaGetErrorString(e);
        return GMRF_B200_ERR_CUDA;
    }
    return 0;
}

void enqueue_factor(gmrf_b200_handle *h, int lanes = 1) {
    const Symbolic &S = h->S;
    cudaStream_t st = h->stream;
    const long long bstride = lanes > 1 ? h->arena_bytes : 0;
    for (int b = 0; b < lanes; b++) {
        char *base = reinterpret_cast<char *>(h->d_Lx) + (long long)b * h->arena_bytes;
        cudaMemsetAsync(base, 0, sizeof(double) * (size_t)S.panel_total, st);
        cudaMemsetAsync(reinterpret_cast<char *>(h->d_fail) + (long long)b * h->arena_bytes, 0x7f, sizeof(int), st);   // "no failure" sentinel
    }
    i64 cnt = (i64)S.q_src.size();
    if (cnt > 0) {
        int grid = (int)std::min<i64>((cnt + 255) / 256, 148 * 16);
        scatter_q_kernel<<<dim3(grid, lanes), 256, 0, st>>>(h->d_Lx, h->d_nz, h->d_qsrc, h->d_qdst, cnt, bstride);
    }
    TableSet T{h->d_gemm, h->d_items, h->d_prefix, h->d_split};
    T.lanes = lanes;
    T.pdl = h->opt.pdl_factor != 0;
    // The finalize launches (block inverses for the solve / selected-inversion phases, parked diagonal squares -> panels)
    // feed nothing later in the factorization: they run on a side stream beside the upper tree levels, which leave
    // most SMs idle, and join before the log-determinant reads the diagonal.
    bool forked = false;
    for (const Launch &L : h->factor_plan.launches) {
        if (L.kind == K_FINALIZE && h->stream2) {
            cudaEventRecord(h->ev_fork, st);
            cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
            run_launch(h, L, T, 0, h->stream2);
            forked = true;
        } else {
            run_launch(h, L, T, 0);
        }
    }
    if (forked) {
        cudaEventRecord(h->ev_join, h->stream2);
        cudaStreamWaitEvent(st, h->ev_join, 0);
    }
    logdet_partial_kernel<<<dim3(LOGDET_BLOCKS, lanes), 256, 0, st>>>(h->d_Lx, h->d_diagpos, S.n, h->d_partial, bstride);
    logdet_final_kernel<<<dim3(1, lanes), 256, 0, st>>>(h->d_partial, h->d_scalars, bstride);
}

void enqueue_selinv(gmrf_b200_handle *h) {
    TableSet T{h->d_gemm_z, h->d_items_z, h->d_prefix_z, h->d_split_z};
    T.pdl = h->opt.pdl_factor != 0;
    for (const Launch &L : h->selinv_plan.launches) run_launch(h, L, T, 0);
}

int ensure_device(gmrf_b200_handle *h) {
    if (!h) return GMRF_B200_ERR_ARG;
    if (h->device < 0) {
        h->err = "numeric call on an analysis-only handle (device < 0): there is no CPU fallback";
        return GMRF_B200_ERR_NO_DEVICE;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    return 0;
}

int do_factor(gmrf_b200_handle *h, int lanes = 1) {
    NvtxRange nvtx_("gmrf_b200:factor");
    cudaStream_t st = h->stream;
    CUDA_TRY(h, cudaEventRecord(h->ev[0], st));
    if (h->opt.use_graph) {
        auto it = h->factor_graphs.find(lanes);
        if (it == h->factor_graphs.end()) {
            cudaGraph_t g;
            cudaGraphExec_t ge;
            CUDA_TRY(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_factor(h, lanes);
            CUDA_TRY(h, cudaStreamEndCapture(st, &g));
            CUDA_TRY(h, cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            it = h->factor_graphs.emplace(lanes, ge).first;
        }
        CUDA_TRY(h, cudaGraphLaunch(it->second, st));
    } else {
        enqueue_factor(h, lanes);
        int rc = check_launch(h, "factorization");
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[1], st));
    h->last_lanes = lanes;
    h->lane_logdet.assign(lanes, 0.0);
    h->lane_fail.assign(lanes, 0);
    for (int b = 0; b < lanes; b++) {      // (a strided 2-D copy would exceed the pitch limit for multi-GB arenas)
        const char *sc = reinterpret_cast<const char *>(h->d_scalars) + (long long)b * h->arena_bytes;
        const char *fl = reinterpret_cast<const char *>(h->d_fail) + (long long)b * h->arena_bytes;
        CUDA_TRY(h, cudaMemcpyAsync(&h->lane_logdet[b], sc, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaMemcpyAsync(&h->lane_fail[b], fl, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(h, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->t_ms[1] = ms;
    for (auto &f : h->lane_fail) f = (f == 0x7f7f7f7f) ? 0 : f;
    h->logdet = h->lane_logdet[0];
    h->fail_col = h->lane_fail[0];
    h->factored = true;
    h->selinv_valid = false;
    return h->fail_col > 0 ? h->fail_col : 0;
}

int build_selinv_tables(gmrf_b200_handle *h) {
    if (h->selinv_tables_built) return 0;
    const Symbolic &S = h->S;
    // failure-atomic: whatever step fails (Z is a second panel array -- an out-of-memory here is realistic), everything
    // built so far is released and the plan cleared, so that a retry starts from scratch instead of replaying a
    // half-built plan or launching kernels on null tables
    auto rollback = [&](int rc) {
        if (h->selinv_graph) { cudaGraphExecDestroy(h->selinv_graph); h->selinv_graph = nullptr; }
        dev_free(h, h->d_Zx); dev_free(h, h->d_gemm_z); dev_free(h, h->d_items_z); dev_free(h, h->d_prefix_z);
        dev_free(h, h->d_trans); dev_free(h, h->d_split_z); dev_free(h, h->d_zdiagpos);
        h->selinv_plan.launches.clear();
        h->n_fast_roots = 0;
        h->selinv_valid = false;
        return rc;
    };
    int rc;
    if ((rc = dev_alloc(h, &h->d_Zx, (size_t)S.panel_total))) return rollback(rc);
    h->d_zw = h->d_upd;
    if (cudaMemsetAsync(h->d_Zx, 0, sizeof(double) * (size_t)S.panel_total, h->stream) != cudaSuccess) { h->err = "cudaMemset failed"; return rollback(GMRF_B200_ERR_CUDA); }
    Builder B;
    B.naive = h->opt.naive_kernels != 0;
    B.splitk_base = h->d_splitk; B.splitk_cap = h->splitk_cap; B.splitk_min_k = h->opt.splitk_min_k; B.large_tile_mask = h->opt.large_tile_mask;
    try {
        build_selinv_plan(h, B);
    } catch (std::exception &e) {
        h->err = e.what();
        return rollback(GMRF_B200_ERR_ARG);
    }
    if ((rc = dev_upload(h, &h->d_gemm_z, B.gemm))) return rollback(rc);
    if ((rc = dev_upload(h, &h->d_items_z, B.items))) return rollback(rc);
    if ((rc = dev_upload(h, &h->d_prefix_z, B.prefix))) return rollback(rc);
    if ((rc = dev_upload(h, &h->d_trans, B.trans))) return rollback(rc);
    if ((rc = dev_upload(h, &h->d_split_z, B.split))) return rollback(rc);
    // diagonal of Z in ORIGINAL ordering: out[perm[k]] = Z[diag_pos[k]]  ->  pos_orig[i] = diag_pos[iperm[i]]
    std::vector<long long> zp(S.n);
    for (i64 i = 0; i < S.n; i++) zp[i] = S.diag_pos[S.iperm[i]];
    if ((rc = dev_upload(h, &h->d_zdiagpos, zp))) return rollback(rc);
    // the uploads above went through the legacy stream from pageable memory: make them (and the memset) visible to the
    // handle's non-blocking stream before anything is launched on it
    if (cudaDeviceSynchronize() != cudaSuccess) { h->err = "device synchronize failed"; return rollback(GMRF_B200_ERR_CUDA); }
    h->selinv_tables_built = true;
    return 0;
}

int ensure_io(gmrf_b200_handle *h, i64 count) {
    if (h->io_cap >= count) return 0;
    // grow-only staging buffer for host<->device transfers of right-hand sides / results
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    dev_free(h, h->d_io);
    h->io_cap = 0;
    int rc = dev_alloc(h, &h->d_io, (size_t)count);
    if (rc) return rc;
    h->io_cap = count;
    return 0;
}

// Lazily build the wide right-hand-side path for block width 64 << wi (GEMM task tables, plans) and make sure the
// shared work arrays hold that many columns.
int ensure_multi(gmrf_b200_handle *h, int wi) {
    const Symbolic &S = h->S;
    const int W = 64 << wi;
    int rc;
    if (h->multi_wcap < W) {
        // grow-only; plans of narrower widths keep pointing into the old arrays, so every built plan is dropped and
        // rebuilt on demand against the new buffers (happens at most twice per handle)
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (auto &m : h->multi) {
            for (auto &g : m.graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
            m.built = false;
            m.fwd_plan.launches.clear();
            m.bwd_plan.launches.clear();
            dev_free(h, m.d_gemm); dev_free(h, m.d_rg); dev_free(h, m.d_prefix); dev_free(h, m.d_superlist);
        }
        dev_free(h, h->d_ym);
        dev_free(h, h->d_um);
        if ((rc = dev_alloc(h, &h->d_ym, (size_t)(S.n * W)))) return rc;
        if ((rc = dev_alloc(h, &h->d_um, (size_t)(S.uvec_total * W)))) return rc;
        h->multi_wcap = W;
    }
    gmrf_b200_handle::Multi &M = h->multi[wi];
    if (M.built) return 0;
    Builder B;
    B.naive = h->opt.naive_kernels != 0;
    try {
        build_multi_plans(h, B, W, M.fwd_plan, M.bwd_plan);
    } catch (std::exception &e) {
        h->err = e.what();
        return GMRF_B200_ERR_ARG;
    }
    if ((rc = dev_upload(h, &M.d_gemm, B.gemm))) return rc;
    if ((rc = dev_upload(h, &M.d_prefix, B.prefix))) return rc;
    if ((rc = dev_upload(h, &M.d_superlist, B.superlist))) return rc;
    if ((rc = dev_upload(h, &M.d_rg, B.rowgather))) return rc;
    CUDA_TRY(h, cudaDeviceSynchronize());      // pageable uploads on the legacy stream -> visible to the handle's stream
    M.built = true;
    return 0;
}

void enqueue_multi_sweeps(gmrf_b200_handle *h, int wi, int mode) {
    gmrf_b200_handle::Multi &M = h->multi[wi];
    TableSet T{M.d_gemm, nullptr, M.d_prefix, nullptr};
    T.superlist = M.d_superlist;
    T.rowgather = M.d_rg;
    T.y = h->d_ym;
    T.u = h->d_um;
    T.pdl = false;               // (first kernel of a sweep: ordinary launch, see enqueue_sweeps)
    const int W = 64 << wi;
    if (mode == 0)
        for (const Launch &L : M.fwd_plan.launches) { run_launch(h, L, T, W); T.pdl = h->opt.pdl_multi != 0; }
    for (const Launch &L : M.bwd_plan.launches) { run_launch(h, L, T, W); T.pdl = h->opt.pdl_multi != 0; }
}

// Wide path: the right-hand sides go through the GEMM sweeps in blocks of 256, then 128, then 64 columns (the last
// block is zero-padded up to its width); wider blocks stream the factor fewer times per column.
int do_solve_device(gmrf_b200_handle *h, const double *dB, double *dX, i64 ld, i64 nrhs, int mode);

int do_solve_device_wide(gmrf_b200_handle *h, const double *dB, double *dX, i64 ld, i64 nrhs_all, int mode) {
    NvtxRange nvtx_("gmrf_b200:solve_wide");
    const Symbolic &S = h->S;
    int rc;
    cudaStream_t st = h->stream;
    // a short tail (<= 8 columns beyond whole 64-column blocks) is cheaper through the few-RHS kernels than as a padded block
    const i64 tail = (nrhs_all > 64 && nrhs_all % 64 != 0 && nrhs_all % 64 <= h->opt.wide_rhs_min) ? nrhs_all % 64 : 0;
    const i64 nrhs = nrhs_all - tail;
    // widths needed by this call; the widest first so the work arrays are sized once
    bool need[MULTI_NW] = {false, false, false};
    for (i64 left = nrhs; left > 0;) {
        int wi = left > 128 ? 2 : left > 64 ? 1 : 0;
        need[wi] = true;
        left -= 64 << wi;
    }
    for (int wi = MULTI_NW - 1; wi >= 0; wi--) {
        if (!need[wi]) continue;
        if ((rc = ensure_multi(h, wi))) return rc;
    }
    for (int wi = 0; wi < MULTI_NW; wi++) {
        if (!need[wi] || !h->opt.use_graph || h->multi[wi].graph[mode]) continue;
        if ((rc = ensure_multi(h, wi))) return rc;            // (re)build if a wider block dropped it
        cudaGraph_t g;
        CUDA_TRY(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        enqueue_multi_sweeps(h, wi, mode);
        CUDA_TRY(h, cudaStreamEndCapture(st, &g));
        CUDA_TRY(h, cudaGraphInstantiate(&h->multi[wi].graph[mode], g, 0));
        cudaGraphDestroy(g);
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[2], st));
    const int tpb = 256;
    const int gridn = (int)((S.n + tpb - 1) / tpb);
    for (i64 r0 = 0; r0 < nrhs;) {
        const i64 left = nrhs - r0;
        const int wi = left > 128 ? 2 : left > 64 ? 1 : 0;
        const int W = 64 << wi;
        const int nb = (int)std::min<i64>(W, left);
        if (nb < W)
            CUDA_TRY(h, cudaMemsetAsync(h->d_ym + (size_t)nb * S.n, 0, sizeof(double) * (size_t)(W - nb) * S.n, st));
        if (mode == 0) {
            permute_rows_kernel<<<gridn, tpb, 0, st>>>(h->d_ym, dB + r0 * ld, h->d_perm, S.n, S.n, ld, nb, 0);
        } else {
            CUDA_TRY(h, cudaMemcpy2DAsync(h->d_ym, sizeof(double) * S.n, dB + r0 * ld, sizeof(double) * ld,
                                          sizeof(double) * S.n, nb, cudaMemcpyDeviceToDevice, st));
        }
        if (h->opt.use_graph) CUDA_TRY(h, cudaGraphLaunch(h->multi[wi].graph[mode], st));
        else enqueue_multi_sweeps(h, wi, mode);
        permute_rows_kernel<<<gridn, tpb, 0, st>>>(dX + r0 * ld, h->d_ym, h->d_perm, S.n, ld, S.n, nb, 1);
        r0 += nb;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], st));
    if ((rc = check_launch(h, "wide solve"))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    if (tail > 0) {
        if ((rc = do_solve_device(h, dB + nrhs * ld, dX + nrhs * ld, ld, tail, mode))) return rc;
        ms += (float)h->t_ms[2];
    }
    h->t_ms[2] = ms;
    return 0;
}

// Enqueue the level-scheduled sweeps on the permuted work array d_y (mode 0: forward + backward, 1: backward only).
void enqueue_sweeps(gmrf_b200_handle *h, int nb, int mode) {
    TableSet T{h->d_gemm, h->d_items, h->d_prefix, h->d_split};
    // the first kernel of a sweep follows a copy / permutation that is not part of the plan: ordinary launch
    T.pdl = false;
    if (mode == 0)
        for (const Launch &L : h->fwd_plan.launches) { run_launch(h, L, T, nb); T.pdl = h->opt.pdl != 0; }
    for (const Launch &L : h->bwd_plan.launches) { run_launch(h, L, T, nb); T.pdl = h->opt.pdl != 0; }
}

// X = Q^-1 B (mode 0) or X = P' L^-T Z (mode 1) on device buffers; nrhs processed in blocks of rhs_block.
// The sweeps of one block are a static launch list on fixed buffers -> captured once per (block width, mode) into a
// CUDA graph and replayed (thousands of dependent micro-launches on the chains of the top supernodes).
int do_solve_device(gmrf_b200_handle *h, const double *dB, double *dX, i64 ld, i64 nrhs, int mode) {
    NvtxRange nvtx_("gmrf_b200:solve");
    const Symbolic &S = h->S;
    if (!h->factored) { h->err = "solve before the first refactorize"; return GMRF_B200_ERR_STATE; }
    if (nrhs < 0 || ld < S.n) { h->err = "solve: need nrhs >= 0 and ld >= n"; return GMRF_B200_ERR_ARG; }
    if (nrhs > h->opt.wide_rhs_min && S.n > 0) return do_solve_device_wide(h, dB, dX, ld, nrhs, mode);
    cudaStream_t st = h->stream;
    // graphs for the block widths this call needs are built before the timed region starts
    auto sweep_graph = [&](int nb, cudaGraphExec_t *out) -> int {
        const int key = nb * 2 + mode;
        auto it = h->solve_graphs.find(key);
        if (it == h->solve_graphs.end()) {
            cudaGraph_t g;
            cudaGraphExec_t ge;
            CUDA_TRY(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_sweeps(h, nb, mode);
            CUDA_TRY(h, cudaStreamEndCapture(st, &g));
            CUDA_TRY(h, cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            it = h->solve_graphs.emplace(key, ge).first;
        }
        *out = it->second;
        return 0;
    };
    cudaGraphExec_t g_full = nullptr, g_tail = nullptr;
    if (h->opt.use_graph && S.n > 0 && nrhs > 0) {
        int rc;
        if (nrhs >= h->rhs_block && (rc = sweep_graph(h->rhs_block, &g_full))) return rc;
        if (nrhs % h->rhs_block && (rc = sweep_graph((int)(nrhs % h->rhs_block), &g_tail))) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[2], st));
    const int tpb = 256;
    const int gridn = (int)((S.n + tpb - 1) / tpb);
    for (i64 r0 = 0; r0 < nrhs; r0 += h->rhs_block) {
        int nb = (int)std::min<i64>(h->rhs_block, nrhs - r0);
        if (S.n == 0) break;
        if (mode == 0) {
            permute_rows_kernel<<<gridn, tpb, 0, st>>>(h->d_y, dB + r0 * ld, h->d_perm, S.n, S.n, ld, nb, 0);
        } else {
            // the half solve takes z in the factor's own ordering (CHOLMOD's `UP \ z`): no input permutation
            CUDA_TRY(h, cudaMemcpy2DAsync(h->d_y, sizeof(double) * S.n, dB + r0 * ld, sizeof(double) * ld,
                                          sizeof(double) * S.n, nb, cudaMemcpyDeviceToDevice, st));
        }
        if (h->opt.use_graph) CUDA_TRY(h, cudaGraphLaunch(nb == h->rhs_block ? g_full : g_tail, st));
        else enqueue_sweeps(h, nb, mode);
        permute_rows_kernel<<<gridn, tpb, 0, st>>>(dX + r0 * ld, h->d_y, h->d_perm, S.n, ld, S.n, nb, 1);
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], st));
    int rc = check_launch(h, "solve");
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[2] = ms;
    return 0;
}

int do_solve_host(gmrf_b200_handle *h, const double *B, double *X, i64 ld, i64 nrhs, int mode) {
    int rc = ensure_device(h);
    if (rc) return rc;
    const Symbolic &S = h->S;
    if (!B || !X) { h->err = "null buffer"; return GMRF_B200_ERR_ARG; }
    if (nrhs < 0 || ld < S.n) { h->err = "solve: need nrhs >= 0 and ld >= n"; return GMRF_B200_ERR_ARG; }
    if (nrhs == 0 || S.n == 0) return 0;
    // stage in chunks so the device footprint stays bounded for very wide right-hand sides; chunks are whole multiples
    // of the wide path's block widths so that no chunk ends in a mostly padded block
    i64 chunk = std::max<i64>(h->rhs_block, std::min<i64>(nrhs, (i64)(256LL << 20) / std::max<i64>(S.n, 1)));
    if (chunk < nrhs) chunk = chunk >= MULTI_WMAX ? chunk / MULTI_WMAX * MULTI_WMAX : chunk >= 64 ? chunk / 64 * 64 : chunk;
    if ((rc = ensure_io(h, S.n * std::min(chunk, nrhs)))) return rc;
    double solve_ms = 0;
    for (i64 r0 = 0; r0 < nrhs; r0 += chunk) {
        i64 nb = std::min(chunk, nrhs - r0);
        CUDA_TRY(h, cudaMemcpy2DAsync(h->d_io, sizeof(double) * S.n, B + r0 * ld, sizeof(double) * ld, sizeof(double) * S.n,
                                      nb, cudaMemcpyHostToDevice, h->stream));
        rc = do_solve_device(h, h->d_io, h->d_io, S.n, nb, mode);
        if (rc) return rc;
        solve_ms += h->t_ms[2];
        CUDA_TRY(h, cudaMemcpy2DAsync(X + r0 * ld, sizeof(double) * ld, h->d_io, sizeof(double) * S.n, sizeof(double) * S.n,
                                      nb, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    h->t_ms[2] = solve_ms;      // device time of the sweeps of all chunks (copies excluded)
    return 0;
}

void build_z_pattern(gmrf_b200_handle *h) {
    // full symmetric CSC on the stored factor pattern, original ordering, sorted rows; pos = Z panel position
    if (h->z_pattern_built) return;
    const Symbolic &S = h->S;
    const i64 n = S.n;
    std::vector<i64> cnt(n + 1, 0);
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 ns = S.ns(s), nrow = S.nrow(s);
        const i32 *rows = S.rowidx.data() + S.rowptr[s];
        for (i64 lc = 0; lc < ns; lc++) {
            i64 cj = S.perm[S.sfirst[s] + lc];
            for (i64 i = lc; i < nrow; i++) {
                i64 ri = S.perm[rows[i]];
                cnt[cj + 1]++;
                if (i != lc) cnt[ri + 1]++;
            }
        }
    }
    for (i64 j = 0; j < n; j++) cnt[j + 1] += cnt[j];
    h->z_colptr = cnt;
    i64 nnz = cnt[n];
    std::vector<i64> rowv(nnz);
    std::vector<long long> pos(nnz);
    std::vector<i64> w(cnt.begin(), cnt.end() - 1);
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
        const i32 *rows = S.rowidx.data() + S.rowptr[s];
        for (i64 lc = 0; lc < ns; lc++) {
            i64 cj = S.perm[S.sfirst[s] + lc];
            for (i64 i = lc; i < nrow; i++) {
                i64 ri = S.perm[rows[i]];
                long long p = S.panel_off[s] + lc * ld + i;
                rowv[w[cj]] = ri; pos[w[cj]++] = p;
                if (i != lc) { rowv[w[ri]] = cj; pos[w[ri]++] = p; }
            }
        }
    }
    // sort rows within each column
    std::vector<std::pair<i64, long long>> tmp;
    for (i64 j = 0; j < n; j++) {
        i64 a = cnt[j], b = cnt[j + 1];
        tmp.resize(b - a);
        for (i64 k = a; k < b; k++) tmp[k - a] = {rowv[k], pos[k]};
        std::sort(tmp.begin(), tmp.end());
        for (i64 k = a; k < b; k++) { rowv[k] = tmp[k - a].first; pos[k] = tmp[k - a].second; }
    }
    h->z_rowval.swap(rowv);
    if (h->device >= 0) {
        dev_upload(h, &h->d_zpos, pos);
    }
    h->z_pattern_built = true;
}

void build_l_pattern(gmrf_b200_handle *h) {
    // P'L as CSC: column k = k-th pivot (elimination order), row indices in the original numbering, sorted;
    // pos = position of the entry in the factor panels (stored pattern: relaxed supernodes carry explicit zeros)
    if (h->l_pattern_built) return;
    const Symbolic &S = h->S;
    const i64 n = S.n;
    h->l_colptr.assign((size_t)n + 1, 0);
    for (i64 s = 0; s < S.nsuper; s++)
        for (i64 lc = 0; lc < S.ns(s); lc++) h->l_colptr[(size_t)(S.sfirst[s] + lc) + 1] = S.nrow(s) - lc;
    for (i64 j = 0; j < n; j++) h->l_colptr[(size_t)j + 1] += h->l_colptr[(size_t)j];
    const i64 nnz = h->l_colptr[(size_t)n];
    std::vector<i64> rowv((size_t)nnz);
    std::vector<long long> pos((size_t)nnz);
    std::vector<std::pair<i64, long long>> tmp;
    for (i64 s = 0; s < S.nsuper; s++) {
        const i64 ns = S.ns(s), nrow = S.nrow(s), ld = S.panel_ld[s];
        const i32 *rows = S.rowidx.data() + S.rowptr[s];
        for (i64 lc = 0; lc < ns; lc++) {
            tmp.resize((size_t)(nrow - lc));
            for (i64 i = lc; i < nrow; i++) tmp[(size_t)(i - lc)] = {S.perm[rows[i]], (long long)(S.panel_off[s] + lc * ld + i)};
            std::sort(tmp.begin(), tmp.end());
            i64 w = h->l_colptr[(size_t)(S.sfirst[s] + lc)];
            for (auto &t : tmp) { rowv[(size_t)w] = t.first; pos[(size_t)w++] = t.second; }
        }
    }
    h->l_rowval.swap(rowv);
    if (h->device >= 0) dev_upload(h, &h->d_lpos, pos);
    h->l_pattern_built = true;
}

}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int gmrf_b200_set_option(const char *key, double value) {
    Options &o = global_options();
    std::string k = key ? key : "";
    if (k == "relax_n0") o.relax_n[0] = value;
    else if (k == "relax_n1") o.relax_n[1] = value;
    else if (k == "relax_n2") o.relax_n[2] = value;
    else if (k == "relax_z0") o.relax_z[1] = value;
    else if (k == "relax_z1") o.relax_z[2] = value;
    else if (k == "relax_z2") o.relax_z[3] = value;
    else if (k == "use_graph") o.use_graph = (int)value;
    else if (k == "outer_block") o.outer_block = (int)value;
    else if (k == "naive_kernels") o.naive_kernels = (int)value;
    else if (k == "selinv_fast_root") o.selinv_fast_root = (int)value;
    else if (k == "wide_rhs_min") o.wide_rhs_min = std::max(0, (int)value);
    else if (k == "bwd_row_chunk") o.bwd_row_chunk = std::max(32, (int)value);
    else if (k == "large_tile_mask") o.large_tile_mask = (int)value & 7;
    else if (k == "lanes") o.lanes = std::max(1, std::min(64, (int)value));
    else if (k == "splitk_min_k") o.splitk_min_k = std::max(16, (int)value);
    else if (k == "debug_alloc_fail_after") g_alloc_fail_countdown = std::max(0, (int)value);
    else if (k == "asm_gather") o.asm_gather = (int)value;
    else if (k == "level_alap") o.level_alap = (int)value;
    else if (k == "syrk_gather") o.syrk_gather = (int)value;
    else if (k == "wide_steps") o.wide_steps = (int)value;
    else if (k == "pdl") o.pdl = value != 0;
    else if (k == "pdl_factor") o.pdl_factor = value != 0;
    else if (k == "pdl_multi") o.pdl_multi = value != 0;
    else if (k == "panel_blocked") o.panel_blocked = (int)value;
    else if (k == "potrf_lookahead") o.potrf_lookahead = value != 0;
    else if (k == "syrk_split") o.syrk_split = (int)value;
    else if (k == "fused_front") o.fused_front = (int)value;
    else if (k == "fused_chain") o.fused_chain = (int)value;
    else if (k == "chain_max_tiles") o.chain_max_tiles = std::max(0, (int)value);
    else if (k == "front_smem_kb") o.front_smem_kb = (int)value;
    else return GMRF_B200_ERR_ARG;
    return 0;
}

const char *gmrf_b200_last_error(const gmrf_b200_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int create_impl(gmrf_b200_handle **out, int64_t n, const int64_t *colptr, const int64_t *rowval, int index_base,
                       const int64_t *perm, int ordering, int device, const void *analysis, int64_t analysis_bytes) {
    g_create_error.clear();
    if (!out) { g_create_error = "out is null"; return GMRF_B200_ERR_ARG; }
    *out = nullptr;
    if (n < 0 || !colptr || (!rowval && n > 0) || (index_base != 0 && index_base != 1)) {
        g_create_error = "bad arguments to gmrf_b200_create";
        return GMRF_B200_ERR_ARG;
    }
    std::unique_ptr<gmrf_b200_handle> h(new gmrf_b200_handle());
    h->opt = global_options();
    h->device = device;
    try {
        std::vector<i64> cp(colptr, colptr + n + 1), rv, pm;
        for (auto &v : cp) v -= index_base;
        if (cp[0] != 0) throw std::runtime_error("colptr[0] must equal index_base");
        i64 nnz = cp[n];
        if (nnz < 0) throw std::runtime_error("negative nnz");
        for (i64 j = 0; j < n; j++)
            if (cp[j + 1] < cp[j] || cp[j + 1] > nnz) throw std::runtime_error("colptr must be non-decreasing");
        rv.assign(rowval, rowval + nnz);
        for (auto &v : rv) v -= index_base;
        if (perm) {
            pm.assign(perm, perm + n);
            for (auto &v : pm) v -= index_base;
        }
        h->diag_nzpos.assign((size_t)n, -1);
        for (i64 j = 0; j < n; j++) {
            const i64 *b = rv.data() + cp[j], *e = rv.data() + cp[j + 1];
            const i64 *it = std::lower_bound(b, e, j);
            if (it != e && *it == j) h->diag_nzpos[(size_t)j] = (long long)(it - rv.data());
        }
        NvtxRange nvtx_("gmrf_b200:analysis");
        h->pattern_hash = gmrf::pattern_hash(n, cp.data(), rv.data());
        if (analysis) gmrf::deserialize(h->S, static_cast<const char *>(analysis), (size_t)analysis_bytes, n, nnz, h->pattern_hash);
        else analyze(h->S, n, cp.data(), rv.data(), perm ? pm.data() : nullptr, ordering, h->opt);
    } catch (std::exception &e) {
        g_create_error = e.what();
        return GMRF_B200_ERR_ARG;
    }
    h->t_ms[4] = h->S.analysis_ms;
    if (device < 0) {
        *out = h.release();
        return 0;
    }
    // ---- device side -------------------------------------------------------------------------------
    gmrf_b200_handle *H = h.get();
    auto fail = [&](int rc) { g_create_error = H->err; gmrf_b200_destroy(h.release()); return rc; };
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || device >= ndev) {
            H->err = "CUDA device " + std::to_string(device) + " is not available (libgmrf_b200 has no CPU fallback)";
            return fail(GMRF_B200_ERR_NO_DEVICE);
        }
        if (cudaSetDevice(device) != cudaSuccess) { H->err = "cudaSetDevice failed"; return fail(GMRF_B200_ERR_CUDA); }
        if (cudaStreamCreateWithFlags(&H->stream, cudaStreamNonBlocking) != cudaSuccess) { H->err = "stream creation failed"; return fail(GMRF_B200_ERR_CUDA); }
        if (cudaStreamCreateWithFlags(&H->stream2, cudaStreamNonBlocking) != cudaSuccess) { H->err = "stream creation failed"; return fail(GMRF_B200_ERR_CUDA); }
        cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&H->ev_join, cudaEventDisableTiming);
        for (auto &e2 : H->ev) cudaEventCreate(&e2);
        if (configure_kernels() != cudaSuccess) { H->err = "cudaFuncSetAttribute failed (is this an sm_100a device?)"; return fail(GMRF_B200_ERR_CUDA); }
    }
    const Symbolic &S = H->S;
    int rc;
#define TRY_RC(x) do { rc = (x); if (rc) return fail(rc); } while (0)
    // ---- numeric arena: every array a factorization writes, once per lane -------------------------------------
    i64 inv_total = 0;   // doubles in the inverted diagonal blocks (nb x nb each, 64-column blocks)
    H->inv_base.assign(S.nsuper, 0);
    for (i64 s = 0; s < S.nsuper; s++) {
        H->inv_base[s] = inv_total;
        for (i64 k0 = 0; k0 < S.ns(s); k0 += SOLVE_NB) {
            i64 nb = std::min<i64>(SOLVE_NB, S.ns(s) - k0);
            inv_total += nb * nb;
        }
    }
    H->splitk_cap = std::min<i64>(32LL << 20, std::max<i64>(1, 24 * S.max_front * (i64)H->opt.outer_block));
    H->lanes = std::max(1, H->opt.lanes);
    if (H->lanes >= 4) {
        // A handle with lanes exists for THROUGHPUT (a sweep of independent value sets fills the GPU by itself): the
        // latency-oriented paths cost it 5 % each (measured, profiles/r02_sweep_lanes.log: redundant diagonal factorization
        // in every chain CTA; the gather extend-add's per-tile set-up), so such a handle takes the bulk path for those.
        H->opt.fused_chain = 0;
        H->opt.asm_gather = 0;
        H->opt.syrk_gather = 0;
    }
    const FusedInfo fused = plan_fused(S, H->opt);
    H->front_smem_max = fused.front_smem_max;
    {
        auto al = [](i64 x) { return (std::max<i64>(x, 1) + 31) & ~31LL; };   // 256-byte granules
        // one pool serves the update matrices of the factorization and, afterwards, the gathered Z[R,R] blocks of the
        // selected inversion (the two phases never overlap)
        const i64 n_lx = al(S.panel_total), n_upd = al(std::max(S.upd_total, S.zw_total)), n_nz = al(S.nnzA),
                  n_inv = al(inv_total), n_split = al(H->splitk_cap), n_part = al(LOGDET_BLOCKS), n_sc = al(8), n_fail = al(2),
                  n_sq = al(fused.sq_total * 4 * (i64)NB * NB);
        const i64 arena = n_lx + n_upd + n_nz + n_inv + n_split + n_part + n_sc + n_fail + n_sq;
        if (H->lanes == 1) {
            // single lane: separate allocations (measured on B200: with panels and update pool inside ONE 78 GB
            // allocation the extend-add kernel ran 91 ms instead of 55 ms per 1 M-dof refactorization)
            H->arena_bytes = 0;
            TRY_RC(dev_alloc(H, &H->d_Lx, (size_t)n_lx));
            TRY_RC(dev_alloc(H, &H->d_upd, (size_t)n_upd));
            TRY_RC(dev_alloc(H, &H->d_nz, (size_t)n_nz));
            TRY_RC(dev_alloc(H, &H->d_Linv, (size_t)n_inv));
            TRY_RC(dev_alloc(H, &H->d_splitk, (size_t)n_split));
            TRY_RC(dev_alloc(H, &H->d_partial, (size_t)n_part));
            TRY_RC(dev_alloc(H, &H->d_scalars, (size_t)n_sc));
            TRY_RC(dev_alloc(H, &H->d_sq, (size_t)n_sq));
            double *f = nullptr;
            TRY_RC(dev_alloc(H, &f, (size_t)n_fail));
            H->d_fail = reinterpret_cast<int *>(f);
        } else {
            H->arena_bytes = arena * (long long)sizeof(double);
            TRY_RC(dev_alloc(H, &H->d_arena, (size_t)arena * (size_t)H->lanes));
            double *p = H->d_arena;
            H->d_Lx = p; p += n_lx;
            H->d_upd = p; p += n_upd;
            H->d_nz = p; p += n_nz;
            H->d_Linv = p; p += n_inv;
            H->d_splitk = p; p += n_split;
            H->d_partial = p; p += n_part;
            H->d_scalars = p; p += n_sc;
            H->d_sq = p; p += n_sq;
            H->d_fail = reinterpret_cast<int *>(p);
        }
    }
    TRY_RC(dev_upload(H, &H->d_invbase, H->inv_base));
    TRY_RC(dev_alloc(H, &H->d_y, (size_t)(S.n * H->rhs_block)));
    TRY_RC(dev_alloc(H, &H->d_uvec, (size_t)(S.uvec_total * H->rhs_block)));
    {
        std::vector<long long> a(S.q_src.begin(), S.q_src.end()), b(S.q_dst.begin(), S.q_dst.end()),
            c(S.diag_pos.begin(), S.diag_pos.end()), d(S.perm.begin(), S.perm.end());
        TRY_RC(dev_upload(H, &H->d_qsrc, a));
        TRY_RC(dev_upload(H, &H->d_qdst, b));
        TRY_RC(dev_upload(H, &H->d_diagpos, c));
        TRY_RC(dev_upload(H, &H->d_perm, d));
        std::vector<int> ri(S.rowidx.begin(), S.rowidx.end()), rl(S.relidx.begin(), S.relidx.end()),
            ch(S.child_idx.begin(), S.child_idx.end());
        TRY_RC(dev_upload(H, &H->d_rowidx, ri));
        TRY_RC(dev_upload(H, &H->d_relidx, rl));
        TRY_RC(dev_upload(H, &H->d_child, ch));
        std::vector<SuperMeta> meta(S.nsuper);
        for (i64 s = 0; s < S.nsuper; s++) {
            SuperMeta &m = meta[s];
            m.panel_off = S.panel_off[s]; m.upd_off = S.upd_off[s]; m.zw_off = S.zw_off[s];
            m.rowptr = S.rowptr[s]; m.uvec_off = S.uvec_off[s];
            m.first = (int)S.sfirst[s]; m.ns = (int)S.ns(s); m.nrow = (int)S.nrow(s);
            m.ld = S.panel_ld[s]; m.uld = S.upd_ld[s]; m.parent = (int)S.sparent[s];
            m.child_begin = (int)S.child_ptr[s]; m.child_end = (int)S.child_ptr[s + 1];
            m.wide = (H->opt.wide_steps && S.ns(s) > SOLVE_WIDE_MIN) ? 1 : 0; m.pad_ = 0;
        }
        TRY_RC(dev_upload(H, &H->d_meta, meta));
        // relpos[s][q] = first position in s's relative-index list (rows below its own columns) that lands at or beyond
        // row 256 q of its parent's front; one extra entry closes the last block
        std::vector<long long> rpo((size_t)S.nsuper + 1, 0);
        for (i64 s = 0; s < S.nsuper; s++) {
            const i64 p = S.sparent[s];
            rpo[(size_t)s + 1] = rpo[(size_t)s] + (p < 0 ? 0 : cdiv(S.nrow(p), AG_RH) + 1);
        }
        std::vector<int> rp((size_t)rpo[(size_t)S.nsuper]);
        for (i64 s = 0; s < S.nsuper; s++) {
            const i64 p = S.sparent[s];
            if (p < 0) continue;
            const i32 *rel = S.relidx.data() + S.rowptr[s] + S.ns(s);
            const i64 cnr = S.nr(s), Q = cdiv(S.nrow(p), AG_RH);
            i64 i = 0;
            for (i64 q = 0; q <= Q; q++) {
                while (i < cnr && rel[i] < q * AG_RH) i++;
                rp[(size_t)(rpo[(size_t)s] + q)] = (int)i;
            }
        }
        TRY_RC(dev_upload(H, &H->d_relpos, rp));
        TRY_RC(dev_upload(H, &H->d_relpos_off, rpo));
        std::vector<GatherCtx> gc(1);
        gc[0] = GatherCtx{H->d_meta, H->d_child, H->d_relidx, H->d_relpos, H->d_relpos_off, H->d_upd};
        TRY_RC(dev_upload(H, &H->d_gctx, gc));
    }
    {
        i64 part_total = 0;   // partial sums of the row-chunked backward products
        for (i64 s = 0; s < S.nsuper; s++)
            if (S.nr(s) > std::max(32, H->opt.bwd_row_chunk))
                part_total += cdiv(S.nr(s), std::max(32, H->opt.bwd_row_chunk)) * (i64)BWD_PART_Q * S.ns(s);
        TRY_RC(dev_alloc(H, &H->d_bwdpart, (size_t)part_total));
    }
    {
        Builder B;
        B.naive = H->opt.naive_kernels != 0;
        B.splitk_base = H->d_splitk; B.splitk_cap = H->splitk_cap; B.splitk_min_k = H->opt.splitk_min_k; B.large_tile_mask = H->opt.large_tile_mask;
        try {
            build_factor_plan(H, B, fused);
            build_solve_plans(H, B);
        } catch (std::exception &e) {
            H->err = e.what();
            return fail(GMRF_B200_ERR_ARG);
        }
        H->gemm_flops_factor = B.gemm_flops;
        H->n_large_tile_launches = B.n_large_tile_launches;
        H->n_splitk_tasks = B.n_splitk_tasks;
        TRY_RC(dev_upload(H, &H->d_chain, B.chain));
        TRY_RC(dev_upload(H, &H->d_final, B.finalize));
        TRY_RC(dev_upload(H, &H->d_front, B.front));
        TRY_RC(dev_upload(H, &H->d_asmtiles, B.asmtiles));
        TRY_RC(dev_upload(H, &H->d_gemm, B.gemm));
        TRY_RC(dev_upload(H, &H->d_panel, B.panel));
        TRY_RC(dev_upload(H, &H->d_items, B.items));
        TRY_RC(dev_upload(H, &H->d_fwd, B.fwd));
        TRY_RC(dev_upload(H, &H->d_bwdg, B.bwdg));
        TRY_RC(dev_upload(H, &H->d_bwds, B.bwds));
        TRY_RC(dev_upload(H, &H->d_bwdr, B.bwdr));
        TRY_RC(dev_upload(H, &H->d_wstep, B.wstep));
        TRY_RC(dev_upload(H, &H->d_wdiag, B.wdiag));
        TRY_RC(dev_upload(H, &H->d_superlist, B.superlist));
        TRY_RC(dev_upload(H, &H->d_prefix, B.prefix));
        TRY_RC(dev_upload(H, &H->d_split, B.split));
    }
    // every table above was copied from pageable memory through the legacy stream; the handle's stream is non-blocking
    // (not ordered against it), so make the uploads visible before the first launch
    if (cudaDeviceSynchronize() != cudaSuccess) { H->err = "device synchronize failed after the table uploads"; return fail(GMRF_B200_ERR_CUDA); }
#undef TRY_RC
    *out = h.release();
    return 0;
}

int gmrf_b200_create(gmrf_b200_handle **out, int64_t n, const int64_t *colptr, const int64_t *rowval, int index_base,
                     const int64_t *perm, int ordering, int device) {
    return create_impl(out, n, colptr, rowval, index_base, perm, ordering, device, nullptr, 0);
}

// ---- persisting / sharing the symbolic analysis ------------------------------------------------------------------
int gmrf_b200_create_from_analysis(gmrf_b200_handle **out, int64_t n, const int64_t *colptr, const int64_t *rowval, int index_base,
                                   const void *analysis, int64_t analysis_bytes, int device) {
    if (!analysis || analysis_bytes <= 0) {
        g_create_error = "create_from_analysis: empty analysis blob";
        if (out) *out = nullptr;
        return GMRF_B200_ERR_ARG;
    }
    return create_impl(out, n, colptr, rowval, index_base, nullptr, 0, device, analysis, analysis_bytes);
}

int gmrf_b200_analysis_export(const gmrf_b200_handle *h, void *buf, int64_t capacity, int64_t *bytes) {
    if (!h || !bytes) return GMRF_B200_ERR_ARG;
    std::vector<char> blob;
    gmrf::serialize(h->S, h->pattern_hash, blob);
    *bytes = (int64_t)blob.size();
    if (!buf) return 0;                                   // size query
    if (capacity < (int64_t)blob.size()) return GMRF_B200_ERR_ARG;
    std::memcpy(buf, blob.data(), blob.size());
    return 0;
}

int gmrf_b200_analysis_equal(const gmrf_b200_handle *a, const gmrf_b200_handle *b) {
    if (!a || !b) return GMRF_B200_ERR_ARG;
    return gmrf::equal(a->S, b->S) && a->diag_nzpos == b->diag_nzpos ? 1 : 0;
}

void gmrf_b200_destroy(gmrf_b200_handle *h) {
    if (!h) return;
    if (h->device >= 0) {
        cudaSetDevice(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        for (auto &kv : h->factor_graphs) cudaGraphExecDestroy(kv.second);
        if (h->selinv_graph) cudaGraphExecDestroy(h->selinv_graph);
        for (auto &kv : h->solve_graphs) cudaGraphExecDestroy(kv.second);
        for (auto &m : h->multi)
            for (auto &g : m.graph) if (g) cudaGraphExecDestroy(g);
        for (void *p : h->owned) cudaFree(p);
        for (auto &e : h->ev) if (e) cudaEventDestroy(e);
        if (h->ev_fork) cudaEventDestroy(h->ev_fork);
        if (h->ev_join) cudaEventDestroy(h->ev_join);
        if (h->stream2) cudaStreamDestroy(h->stream2);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

int gmrf_b200_refactorize_device(gmrf_b200_handle *h, const double *d_nzval, int64_t nnz) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (nnz != h->S.nnzA) {
        h->err = "nzval holds " + std::to_string(nnz) + " values but the pattern has " + std::to_string(h->S.nnzA) +
                 " nonzeros; the sparsity pattern must be invariant across refactorizations";
        return GMRF_B200_ERR_ARG;
    }
    if (!d_nzval && nnz > 0) { h->err = "null nzval"; return GMRF_B200_ERR_ARG; }
    if (d_nzval != h->d_nz && nnz > 0)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_nz, d_nzval, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
    h->t_ms[0] = 0;
    return do_factor(h);
}

int gmrf_b200_refactorize(gmrf_b200_handle *h, const double *nzval, int64_t nnz) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (nnz != h->S.nnzA) {
        h->err = "nzval holds " + std::to_string(nnz) + " values but the pattern has " + std::to_string(h->S.nnzA) +
                 " nonzeros; the sparsity pattern must be invariant across refactorizations";
        return GMRF_B200_ERR_ARG;
    }
    if (!nzval && nnz > 0) { h->err = "null nzval"; return GMRF_B200_ERR_ARG; }
    CUDA_TRY(h, cudaEventRecord(h->ev[2], h->stream));
    if (nnz > 0) CUDA_TRY(h, cudaMemcpyAsync(h->d_nz, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
    rc = do_factor(h);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[0] = ms;
    return rc;
}

// ---- device-side value assembly (hyperparameter loops) ------------------------------------------------------
// replaces the host-side re-assembly + upload of nzval per theta: `(model)(ws; theta...)` ->
// _pad_to_workspace_pattern (src/workspace/latent_model_integration.jl:151-250) + _copy_sparse_values! (backend.jl:165-176).
// set_value_basis uploads `nbasis` value arrays laid out on the workspace pattern ONCE; refactorize_combination forms
// nzval = sum_j coeff[j] * basis_j in HBM and factorizes, so a theta evaluation moves `nbasis` doubles over PCIe.
int gmrf_b200_set_value_basis(gmrf_b200_handle *h, const double *basis, int nbasis) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!basis || nbasis < 1 || nbasis > MAX_VALUE_BASIS) { h->err = "set_value_basis: need 1 <= nbasis <= 8 value arrays"; return GMRF_B200_ERR_ARG; }
    const size_t cnt = (size_t)nbasis * (size_t)h->S.nnzA;
    if (h->d_basis && h->nbasis != nbasis) dev_free(h, h->d_basis);
    if (!h->d_basis && (rc = dev_alloc(h, &h->d_basis, cnt))) return rc;
    h->nbasis = nbasis;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_basis, basis, sizeof(double) * cnt, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gmrf_b200_refactorize_combination(gmrf_b200_handle *h, const double *coeff, int nbasis) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->d_basis || nbasis != h->nbasis || !coeff) { h->err = "refactorize_combination: call set_value_basis first (same nbasis)"; return GMRF_B200_ERR_STATE; }
    BasisCoeff c;
    for (int j = 0; j < MAX_VALUE_BASIS; j++) c.c[j] = j < nbasis ? coeff[j] : 0.0;
    const i64 nnz = h->S.nnzA;
    if (nnz > 0) {
        const int grid = (int)std::min<i64>((nnz + 255) / 256, 148 * 32);
        combine_basis_kernel<<<grid, 256, 0, h->stream>>>(h->d_nz, h->d_basis, c, nbasis, nnz);
        if ((rc = check_launch(h, "value assembly"))) return rc;
    }
    h->t_ms[0] = 0;
    return do_factor(h);
}

// ---- Newton loops: Q_k = Q_prior - Diagonal(h_k) formed in HBM -------------------------------------------------
// replaces, for diagonal observation Hessians, the per-iterate host rebuild + upload of nzval in
// _update_hessian! (src/workspace/gaussian_approximation.jl:96-129) followed by refactorize! (backend.jl:178-189):
// the prior's values are uploaded once, an iterate moves n doubles.
int gmrf_b200_set_base_values(gmrf_b200_handle *h, const double *nzval, int64_t nnz) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (nnz != h->S.nnzA || (!nzval && nnz > 0)) { h->err = "set_base_values: nzval must hold nnz(Q) values"; return GMRF_B200_ERR_ARG; }
    if (!h->d_base) {
        if ((rc = dev_alloc(h, &h->d_base, (size_t)nnz))) return rc;
        if ((rc = dev_alloc(h, &h->d_hdiag, (size_t)h->S.n))) return rc;
        if ((rc = dev_upload(h, &h->d_diagnz, h->diag_nzpos))) return rc;
    }
    if (nnz > 0) CUDA_TRY(h, cudaMemcpyAsync(h->d_base, nzval, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gmrf_b200_refactorize_base_minus_diag(gmrf_b200_handle *h, const double *diag, int64_t n) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->d_base) { h->err = "refactorize_base_minus_diag: call set_base_values first"; return GMRF_B200_ERR_STATE; }
    if (n != h->S.n || (!diag && n > 0)) { h->err = "refactorize_base_minus_diag: diag must hold n values"; return GMRF_B200_ERR_ARG; }
    cudaStream_t st = h->stream;
    CUDA_TRY(h, cudaEventRecord(h->ev[2], st));
    if (n > 0) {
        CUDA_TRY(h, cudaMemcpyAsync(h->d_hdiag, diag, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_nz, h->d_base, sizeof(double) * (size_t)h->S.nnzA, cudaMemcpyDeviceToDevice, st));
        minus_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d_nz, h->d_diagnz, h->d_hdiag, n);
        if ((rc = check_launch(h, "diagonal update"))) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], st));
    rc = do_factor(h);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[0] = ms;
    return rc;
}

// Sparse observation Hessians (replaces _sparse_hessian_map + _subtract_sparse_hessian! + refactorize!,
// src/workspace/gaussian_approximation.jl:31-83, backend.jl:178-189): the nzval positions of the Hessian's stored entries
// are uploaded once per Newton loop, an iterate then moves nnz(H) doubles and forms Q_prior - H in HBM.
int gmrf_b200_set_hessian_pattern(gmrf_b200_handle *h, const int64_t *nzpos, int64_t count, int index_base) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (count < 0 || (!nzpos && count > 0) || (index_base != 0 && index_base != 1)) { h->err = "set_hessian_pattern: bad arguments"; return GMRF_B200_ERR_ARG; }
    std::vector<long long> p((size_t)count);
    for (i64 k = 0; k < count; k++) {
        p[(size_t)k] = nzpos[k] - index_base;
        if (p[(size_t)k] < 0 || p[(size_t)k] >= h->S.nnzA) { h->err = "set_hessian_pattern: position outside the workspace pattern"; return GMRF_B200_ERR_ARG; }
    }
    {   // one owner per entry (no floating-point atomics): duplicates must be summed by the caller (a CSC matrix has none)
        std::vector<long long> q(p);
        std::sort(q.begin(), q.end());
        if (std::adjacent_find(q.begin(), q.end()) != q.end()) { h->err = "set_hessian_pattern: duplicate positions"; return GMRF_B200_ERR_ARG; }
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    dev_free(h, h->d_hpos);
    h->hpos_count = 0;
    if ((rc = dev_upload(h, &h->d_hpos, p))) return rc;
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->hpos_count = count;
    return 0;
}

int gmrf_b200_refactorize_base_minus_sparse(gmrf_b200_handle *h, const double *values, int64_t count) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->d_base) { h->err = "refactorize_base_minus_sparse: call set_base_values first"; return GMRF_B200_ERR_STATE; }
    if (count != h->hpos_count || (!values && count > 0)) { h->err = "refactorize_base_minus_sparse: values must match set_hessian_pattern"; return GMRF_B200_ERR_ARG; }
    cudaStream_t st = h->stream;
    if ((rc = ensure_io(h, std::max<i64>(count, 1)))) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev[2], st));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_nz, h->d_base, sizeof(double) * (size_t)h->S.nnzA, cudaMemcpyDeviceToDevice, st));
    if (count > 0) {
        CUDA_TRY(h, cudaMemcpyAsync(h->d_io, values, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st));
        minus_sparse_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(h->d_nz, h->d_hpos, h->d_io, count);
        if ((rc = check_launch(h, "sparse Hessian update"))) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], st));
    rc = do_factor(h);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[0] = ms;
    return rc;
}

// ---- lanes: several value sets factorized side by side ------------------------------------------------------
// The workload of a hyperparameter sweep (`WorkspacePool` + `(model)(ws; theta...)`, workspace_pool.jl:42-119): many
// independent numeric factorizations of ONE pattern whose only outputs are log-determinants. A handle created with the
// option "lanes" = B holds B copies of its numeric arrays and advances all of them with the same launches, which is
// what fills the GPU on problems (2D meshes) too small to do so alone. Lane 0 is the handle's ordinary factor: solves
// and selected inversion after a lane call see lane 0.
static int check_lanes(gmrf_b200_handle *h, int lanes, const void *a, const void *b) {
    if (lanes < 1 || lanes > h->lanes) {
        h->err = "lanes = " + std::to_string(lanes) + " but the handle was created with capacity " + std::to_string(h->lanes) +
                 " (set_option(\"lanes\", B) before create)";
        return GMRF_B200_ERR_ARG;
    }
    if (!a || !b) { h->err = "null buffer"; return GMRF_B200_ERR_ARG; }
    return 0;
}
static void lanes_out(gmrf_b200_handle *h, int lanes, double *logdet, int *status) {
    for (int b = 0; b < lanes; b++) {
        logdet[b] = h->lane_logdet[b];
        if (status) status[b] = h->lane_fail[b];
    }
}

int gmrf_b200_lane_capacity(const gmrf_b200_handle *h) { return h ? h->lanes : 0; }

int gmrf_b200_refactorize_lanes(gmrf_b200_handle *h, const double *nzval, int64_t nnz, int lanes, double *logdet, int *status) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if ((rc = check_lanes(h, lanes, nzval, logdet))) return rc;
    if (nnz != h->S.nnzA) { h->err = "nzval length does not match the pattern"; return GMRF_B200_ERR_ARG; }
    CUDA_TRY(h, cudaEventRecord(h->ev[2], h->stream));
    for (int b = 0; b < lanes && nnz > 0; b++)
        CUDA_TRY(h, cudaMemcpyAsync(reinterpret_cast<char *>(h->d_nz) + (long long)b * h->arena_bytes, nzval + (size_t)b * (size_t)nnz,
                                    sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev[3], h->stream));
    rc = do_factor(h, lanes);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[0] = ms;
    if (rc < 0) return rc;
    lanes_out(h, lanes, logdet, status);
    return 0;
}

int gmrf_b200_refactorize_combination_lanes(gmrf_b200_handle *h, const double *coeff, int nbasis, int lanes, double *logdet, int *status) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if ((rc = check_lanes(h, lanes, coeff, logdet))) return rc;
    if (!h->d_basis || nbasis != h->nbasis) { h->err = "refactorize_combination_lanes: call set_value_basis first (same nbasis)"; return GMRF_B200_ERR_STATE; }
    const i64 nnz = h->S.nnzA;
    for (int b = 0; b < lanes && nnz > 0; b++) {
        BasisCoeff c;
        for (int j = 0; j < MAX_VALUE_BASIS; j++) c.c[j] = j < nbasis ? coeff[(size_t)b * nbasis + j] : 0.0;
        double *nz = reinterpret_cast<double *>(reinterpret_cast<char *>(h->d_nz) + (long long)b * h->arena_bytes);
        const int grid = (int)std::min<i64>((nnz + 255) / 256, 148 * 32);
        combine_basis_kernel<<<grid, 256, 0, h->stream>>>(nz, h->d_basis, c, nbasis, nnz);
    }
    if ((rc = check_launch(h, "value assembly"))) return rc;
    h->t_ms[0] = 0;
    rc = do_factor(h, lanes);
    if (rc < 0) return rc;
    lanes_out(h, lanes, logdet, status);
    return 0;
}

int gmrf_b200_logdet(gmrf_b200_handle *h, double *out) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored) { h->err = "logdet before the first refactorize"; return GMRF_B200_ERR_STATE; }
    if (!out) { h->err = "null out"; return GMRF_B200_ERR_ARG; }
    *out = h->logdet;
    return 0;
}

int gmrf_b200_solve(gmrf_b200_handle *h, const double *B, double *X, int64_t ld, int64_t nrhs) {
    return do_solve_host(h, B, X, ld, nrhs, 0);
}
int gmrf_b200_solve_Lt(gmrf_b200_handle *h, const double *Z, double *X, int64_t ld, int64_t nrhs) {
    return do_solve_host(h, Z, X, ld, nrhs, 1);
}
int gmrf_b200_solve_device(gmrf_b200_handle *h, const double *dB, double *dX, int64_t ld, int64_t nrhs) {
    int rc = ensure_device(h);
    if (rc) return rc;
    return do_solve_device(h, dB, dX, ld, nrhs, 0);
}
int gmrf_b200_solve_Lt_device(gmrf_b200_handle *h, const double *dZ, double *dX, int64_t ld, int64_t nrhs) {
    int rc = ensure_device(h);
    if (rc) return rc;
    return do_solve_device(h, dZ, dX, ld, nrhs, 1);
}

int gmrf_b200_selinv_compute(gmrf_b200_handle *h) {
    NvtxRange nvtx_("gmrf_b200:selinv");
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored) { h->err = "selinv before the first refactorize"; return GMRF_B200_ERR_STATE; }
    if (h->selinv_valid) return 0;
    if ((rc = build_selinv_tables(h))) return rc;
    cudaStream_t st = h->stream;
    CUDA_TRY(h, cudaEventRecord(h->ev[2], st));
    if (h->opt.use_graph) {
        if (!h->selinv_graph) {
            cudaGraph_t g;
            CUDA_TRY(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_selinv(h);
            CUDA_TRY(h, cudaStreamEndCapture(st, &g));
            CUDA_TRY(h, cudaGraphInstantiate(&h->selinv_graph, g, 0));
            cudaGraphDestroy(g);
        }
        CUDA_TRY(h, cudaGraphLaunch(h->selinv_graph, st));
    } else {
        enqueue_selinv(h);
        if ((rc = check_launch(h, "selected inversion"))) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev[3], st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
    h->t_ms[3] = ms;
    h->selinv_valid = true;
    return 0;
}

static int gather_to_host(gmrf_b200_handle *h, const long long *d_pos, i64 cnt, double *out, const double *d_src = nullptr) {
    int rc;
    if ((rc = ensure_io(h, std::max<i64>(cnt, 1)))) return rc;
    if (cnt == 0) return 0;
    int grid = (int)std::min<i64>((cnt + 255) / 256, 148 * 16);
    gather_values_kernel<<<grid, 256, 0, h->stream>>>(h->d_io, d_src ? d_src : h->d_Zx, d_pos, cnt);
    if ((rc = check_launch(h, "gather"))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_io, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gmrf_b200_selinv_diag(gmrf_b200_handle *h, double *out) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (!out) { h->err = "null out"; return GMRF_B200_ERR_ARG; }
    return gather_to_host(h, h->d_zdiagpos, h->S.n, out);
}

int gmrf_b200_selinv_nnz(gmrf_b200_handle *h, int64_t *nnz) {
    if (!h || !nnz) return GMRF_B200_ERR_ARG;
    if (h->device >= 0) cudaSetDevice(h->device);
    build_z_pattern(h);
    *nnz = h->z_colptr[h->S.n];
    return 0;
}

int gmrf_b200_selinv_pattern(gmrf_b200_handle *h, int64_t *colptr, int64_t *rowval, int index_base) {
    if (!h || !colptr || !rowval) return GMRF_B200_ERR_ARG;
    if (h->device >= 0) cudaSetDevice(h->device);
    build_z_pattern(h);
    for (i64 j = 0; j <= h->S.n; j++) colptr[j] = h->z_colptr[j] + index_base;
    for (size_t k = 0; k < h->z_rowval.size(); k++) rowval[k] = h->z_rowval[k] + index_base;
    return 0;
}

int gmrf_b200_selinv_values(gmrf_b200_handle *h, double *nzval) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (!nzval) { h->err = "null out"; return GMRF_B200_ERR_ARG; }
    build_z_pattern(h);
    if (!h->d_zpos) { h->err = "selinv pattern upload failed"; return GMRF_B200_ERR_ALLOC; }
    return gather_to_host(h, h->d_zpos, h->z_colptr[h->S.n], nzval);
}

// ---- factor export: the square root P'L of Q as a sparse matrix -----------------------------------------------------
int gmrf_b200_factor_nnz(gmrf_b200_handle *h, int64_t *nnz) {
    if (!h || !nnz) return GMRF_B200_ERR_ARG;
    if (h->device >= 0) cudaSetDevice(h->device);
    build_l_pattern(h);
    *nnz = h->l_colptr[(size_t)h->S.n];
    return 0;
}

int gmrf_b200_factor_pattern(gmrf_b200_handle *h, int64_t *colptr, int64_t *rowval, int index_base) {
    if (!h || !colptr || !rowval) return GMRF_B200_ERR_ARG;
    if (h->device >= 0) cudaSetDevice(h->device);
    build_l_pattern(h);
    for (i64 j = 0; j <= h->S.n; j++) colptr[j] = h->l_colptr[(size_t)j] + index_base;
    for (size_t k = 0; k < h->l_rowval.size(); k++) rowval[k] = h->l_rowval[k] + index_base;
    return 0;
}

int gmrf_b200_factor_values(gmrf_b200_handle *h, double *nzval) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored) { h->err = "factor_values before the first refactorize"; return GMRF_B200_ERR_STATE; }
    if (!nzval) { h->err = "null out"; return GMRF_B200_ERR_ARG; }
    build_l_pattern(h);
    if (!h->d_lpos) { h->err = "factor pattern upload failed"; return GMRF_B200_ERR_ALLOC; }
    return gather_to_host(h, h->d_lpos, h->l_colptr[(size_t)h->S.n], nzval, h->d_Lx);
}

// Position of Sigma_ij in the Z panels for every entry of a caller pattern (n x n CSC); -1 outside the factor's pattern.
// The result lives in the handle (h->pp_pos) together with a copy of the pattern it belongs to.
static int pattern_positions(gmrf_b200_handle *h, const char *who, int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                             int index_base, const std::vector<long long> **out) {
    const Symbolic &S = h->S;
    *out = &h->pp_pos;
    if (ncol != S.n || !colptr || (!rowval && colptr[ncol] - index_base > 0)) {
        h->err = std::string(who) + ": pattern must be n x n";
        return GMRF_B200_ERR_ARG;
    }
    const i64 cnt = colptr[ncol] - index_base;
    if (h->pp_valid && h->pp_base == index_base && (i64)h->pp_rowval.size() == cnt && (i64)h->pp_colptr.size() == ncol + 1 &&
        std::memcmp(h->pp_colptr.data(), colptr, sizeof(i64) * (size_t)(ncol + 1)) == 0 &&
        (cnt <= 0 || std::memcmp(h->pp_rowval.data(), rowval, sizeof(i64) * (size_t)cnt) == 0)) {
        h->pp_hits++;
        return 0;
    }
    h->pp_valid = false;
    h->d_ppos_valid = false;
    for (i64 j = 0; j < ncol; j++)
        if (colptr[j + 1] < colptr[j] || colptr[j] < index_base) { h->err = std::string(who) + ": colptr must be non-decreasing from index_base"; return GMRF_B200_ERR_ARG; }
    h->pp_pos.assign((size_t)std::max<i64>(cnt, 0), -1LL);
    if (cnt > 0 && gmrf::pattern_positions(S, colptr, rowval, index_base, h->pp_pos.data()) >= 0) {
        h->err = std::string(who) + ": row index out of range";
        return GMRF_B200_ERR_ARG;
    }
    if (cnt <= (i64)1 << 27) {       // remember the pattern (16 bytes per entry) unless it is huge
        h->pp_colptr.assign(colptr, colptr + ncol + 1);
        h->pp_rowval.assign(rowval, rowval + std::max<i64>(cnt, 0));
        h->pp_base = index_base;
        h->pp_valid = true;
    }
    return 0;
}

// device copy of the positions of the last looked-up pattern (grow-only buffer, reused while the pattern repeats)
static int resident_positions(gmrf_b200_handle *h, const std::vector<long long> &pos, const long long **d_pos) {
    const i64 cnt = (i64)pos.size();
    if (!h->d_ppos_valid) {
        if (h->d_ppos_cap < cnt) {
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            dev_free(h, h->d_ppos);
            h->d_ppos_cap = 0;
            int rc = dev_alloc(h, &h->d_ppos, (size_t)cnt);
            if (rc) return rc;
            h->d_ppos_cap = cnt;
        }
        CUDA_TRY(h, cudaMemcpyAsync(h->d_ppos, pos.data(), sizeof(long long) * (size_t)cnt, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));       // pos is pageable host memory owned by the handle
        h->d_ppos_valid = h->pp_valid;
    }
    *d_pos = h->d_ppos;
    return 0;
}

// introspection twin of the lookup above (host-side, valid on analysis-only handles): the panel offsets themselves
int gmrf_b200_pattern_positions(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr, const int64_t *rowval, int index_base,
                                int64_t *pos) {
    if (!h || !pos) return GMRF_B200_ERR_ARG;
    const std::vector<long long> *p;
    int rc = pattern_positions(h, "pattern_positions", ncol, colptr, rowval, index_base, &p);
    if (rc) return rc;
    std::copy(p->begin(), p->end(), pos);
    return 0;
}

int gmrf_b200_selinv_extract(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr, const int64_t *rowval,
                             int index_base, double *out) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (!out) { h->err = "selinv_extract: null out"; return GMRF_B200_ERR_ARG; }
    const std::vector<long long> *ppos;
    if ((rc = pattern_positions(h, "selinv_extract", ncol, colptr, rowval, index_base, &ppos))) return rc;
    const std::vector<long long> &pos = *ppos;
    const i64 cnt = (i64)pos.size();
    if (cnt <= 0) return 0;
    const long long *d_pos = nullptr;
    if ((rc = resident_positions(h, pos, &d_pos))) return rc;
    return gather_to_host(h, d_pos, cnt, out);
}

// ---- traces against the selected inverse, contracted on the device ---------------------------------------------
static int ensure_dot(gmrf_b200_handle *h) {
    if (h->d_dot) return 0;
    return dev_alloc(h, &h->d_dot, (size_t)MAX_VALUE_BASIS * (DOT_BLOCKS + 1));
}

// partial sums + fixed-shape final reduction for `nsets` value sets, results to the host
static int run_dot(gmrf_b200_handle *h, const long long *d_pos, const long long *d_idx, const double *d_w, const double *d_val,
                   long long vstride, i64 cnt, int nsets, double *out) {
    int rc;
    double *d_res = h->d_dot + (size_t)MAX_VALUE_BASIS * DOT_BLOCKS;
    gather_dot_partial_kernel<<<dim3(DOT_BLOCKS, nsets), 256, 0, h->stream>>>(h->d_Zx, d_pos, d_idx, d_w, d_val, vstride, cnt, h->d_dot);
    gather_dot_final_kernel<<<dim3(1, nsets), 256, 0, h->stream>>>(h->d_dot, d_res);
    if ((rc = check_launch(h, "selinv_dot"))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, d_res, sizeof(double) * (size_t)nsets, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int gmrf_b200_selinv_dot(gmrf_b200_handle *h, int64_t ncol, const int64_t *colptr, const int64_t *rowval, int index_base,
                         const double *values, double *out) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (!out) { h->err = "selinv_dot: null out"; return GMRF_B200_ERR_ARG; }
    const std::vector<long long> *ppos;
    if ((rc = pattern_positions(h, "selinv_dot", ncol, colptr, rowval, index_base, &ppos))) return rc;
    const std::vector<long long> &pos = *ppos;
    const i64 cnt = (i64)pos.size();
    *out = 0.0;
    if (cnt <= 0) return 0;
    if (!values) { h->err = "selinv_dot: null values"; return GMRF_B200_ERR_ARG; }
    if ((rc = ensure_dot(h))) return rc;
    if ((rc = ensure_io(h, cnt))) return rc;
    const long long *d_pos = nullptr;
    if ((rc = resident_positions(h, pos, &d_pos))) return rc;
    cudaError_t e = cudaMemcpyAsync(h->d_io, values, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) rc = run_dot(h, d_pos, nullptr, nullptr, h->d_io, 0, cnt, 1, out);
    else { h->err = "H2D copy failed"; rc = GMRF_B200_ERR_CUDA; }
    if (rc) cudaStreamSynchronize(h->stream);   // values are pageable host memory: nothing may still be in flight
    return rc;
}

int gmrf_b200_selinv_dot_basis(gmrf_b200_handle *h, double *out, int nbasis) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (!h->d_basis || nbasis != h->nbasis || !out) { h->err = "selinv_dot_basis: call set_value_basis first (same nbasis)"; return GMRF_B200_ERR_STATE; }
    const Symbolic &S = h->S;
    const i64 cnt = (i64)S.q_src.size();
    for (int j = 0; j < nbasis; j++) out[j] = 0.0;
    if (cnt == 0) return 0;
    if ((rc = ensure_dot(h))) return rc;
    if (!h->d_qzw) {
        // the factor and Z share one panel layout, so the Q -> panel scatter map doubles as the gather map; it lists
        // the stored upper triangle once per unordered pair: off-diagonal entries count twice in the trace
        std::vector<double> w((size_t)cnt, 2.0);
        std::vector<long long> dn;
        for (long long p : h->diag_nzpos) if (p >= 0) dn.push_back(p);
        std::sort(dn.begin(), dn.end());
        for (i64 k = 0; k < cnt; k++)
            if (std::binary_search(dn.begin(), dn.end(), (long long)S.q_src[(size_t)k])) w[(size_t)k] = 1.0;
        if ((rc = dev_upload(h, &h->d_qzw, w))) return rc;
    }
    return run_dot(h, h->d_qdst, h->d_qsrc, h->d_qzw, h->d_basis, (long long)S.nnzA, cnt, nbasis, out);
}

// diag(A Sigma A') for a sparse m x n design matrix A in CSR (replaces _row_diag_AΣAt, src/linear_predictor_marginals.jl:
// 137-165: Sigma read at the pattern of A'A and contracted row by row): the panel positions of every index pair of a row
// are looked up on the host (OpenMP over rows), positions and weights A_ia * A_ib go to the device, one warp per row
// contracts them against Z in HBM. Pairs outside the factor's pattern count 0, like selinv_extract.
int gmrf_b200_selinv_quadform_rows(gmrf_b200_handle *h, int64_t m, const int64_t *rowptr, const int64_t *colidx, const double *values,
                                   int index_base, double *out) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (m < 0 || !rowptr || !out || (index_base != 0 && index_base != 1)) { h->err = "selinv_quadform_rows: bad arguments"; return GMRF_B200_ERR_ARG; }
    if (m == 0) return 0;
    const Symbolic &S = h->S;
    const i64 nnz = rowptr[m] - index_base;
    if (nnz < 0 || (nnz > 0 && (!colidx || !values))) { h->err = "selinv_quadform_rows: bad arguments"; return GMRF_B200_ERR_ARG; }
    std::vector<long long> seg((size_t)m + 1, 0);
    for (i64 i = 0; i < m; i++) {
        const i64 c = rowptr[i + 1] - rowptr[i];
        if (c < 0) { h->err = "selinv_quadform_rows: rowptr must be non-decreasing"; return GMRF_B200_ERR_ARG; }
        seg[(size_t)i + 1] = seg[(size_t)i] + c * c;
    }
    for (i64 p = 0; p < nnz; p++)
        if (colidx[p] - index_base < 0 || colidx[p] - index_base >= S.n) { h->err = "selinv_quadform_rows: column index out of range"; return GMRF_B200_ERR_ARG; }
    const i64 tot = seg[(size_t)m];
    std::vector<long long> pos((size_t)tot);
    std::vector<double> w((size_t)tot);
#pragma omp parallel for schedule(dynamic, 256)
    for (i64 i = 0; i < m; i++) {
        const i64 lo = rowptr[i] - index_base, c = rowptr[i + 1] - rowptr[i];
        long long q = seg[(size_t)i];
        for (i64 a = 0; a < c; a++)
            for (i64 b = 0; b < c; b++, q++) {
                pos[(size_t)q] = gmrf::entry_position(S, colidx[lo + a] - index_base, colidx[lo + b] - index_base);
                w[(size_t)q] = values[lo + a] * values[lo + b];
            }
    }
    long long *d_pos = nullptr, *d_seg = nullptr;
    double *d_w = nullptr, *d_out = nullptr;
    auto cleanup = [&](int r) { dev_free(h, d_pos); dev_free(h, d_seg); dev_free(h, d_w); dev_free(h, d_out); return r; };
    if ((rc = dev_upload(h, &d_pos, pos)) || (rc = dev_upload(h, &d_seg, seg)) || (rc = dev_upload(h, &d_w, w)) || (rc = dev_alloc(h, &d_out, (size_t)m)))
        return cleanup(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) { h->err = "device synchronize failed"; return cleanup(GMRF_B200_ERR_CUDA); }
    segment_dot_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, h->stream>>>(h->d_Zx, d_pos, d_w, d_seg, m, d_out);
    if ((rc = check_launch(h, "selinv_quadform_rows"))) return cleanup(rc);
    if (cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "D2H copy failed"; return cleanup(GMRF_B200_ERR_CUDA); }
    return cleanup(0);
}

// ---- introspection -------------------------------------------------------------------------------
int gmrf_b200_info(const gmrf_b200_handle *h, int64_t *info, int n_info) {
    if (!h || !info) return GMRF_B200_ERR_ARG;
    const Symbolic &S = h->S;
    int64_t v[GMRF_B200_INFO_COUNT];
    v[GMRF_B200_INFO_N] = S.n;
    v[GMRF_B200_INFO_NNZ_Q] = S.nnzA;
    v[GMRF_B200_INFO_NNZ_L] = S.nnzL;
    v[GMRF_B200_INFO_NNZ_L_STORED] = S.panel_total;
    v[GMRF_B200_INFO_NSUPER] = S.nsuper;
    v[GMRF_B200_INFO_NLEVELS] = S.nlevels;
    v[GMRF_B200_INFO_MAX_FRONT] = S.max_front;
    v[GMRF_B200_INFO_MAX_NS] = S.max_ns;
    v[GMRF_B200_INFO_UPDATE_POOL] = S.upd_total;
    v[GMRF_B200_INFO_FLOPS_CHOL] = (int64_t)S.flops;
    v[GMRF_B200_INFO_FLOPS_CHOL_STORED] = (int64_t)S.flops_stored;
    v[GMRF_B200_INFO_DEVICE_BYTES] = (int64_t)h->device_bytes;
    v[GMRF_B200_INFO_GRAPH_NODES] = (int64_t)h->factor_plan.launches.size() + 5;
    v[GMRF_B200_INFO_SELINV_NODES] = (int64_t)h->selinv_plan.launches.size();
    v[GMRF_B200_INFO_PATTERN_CACHE_HITS] = (int64_t)h->pp_hits;
    v[GMRF_B200_INFO_LARGE_TILE_LAUNCHES] = (int64_t)h->n_large_tile_launches;
    v[GMRF_B200_INFO_SPLITK_TASKS] = (int64_t)h->n_splitk_tasks;
    v[GMRF_B200_INFO_FAST_ROOTS] = (int64_t)h->n_fast_roots;
    v[GMRF_B200_INFO_CHAIN_LAUNCHES] = (int64_t)h->n_chain_launches;
    v[GMRF_B200_INFO_FRONT_LAUNCHES] = (int64_t)h->n_front_launches;
    for (int i = 0; i < n_info && i < GMRF_B200_INFO_COUNT; i++) info[i] = v[i];
    return 0;
}

int gmrf_b200_get_perm(const gmrf_b200_handle *h, int64_t *perm, int index_base) {
    if (!h || !perm) return GMRF_B200_ERR_ARG;
    for (i64 k = 0; k < h->S.n; k++) perm[k] = h->S.perm[k] + index_base;
    return 0;
}
int gmrf_b200_get_colcounts(const gmrf_b200_handle *h, int64_t *cc) {
    if (!h || !cc) return GMRF_B200_ERR_ARG;
    std::copy(h->S.colcount.begin(), h->S.colcount.end(), cc);
    return 0;
}
int gmrf_b200_get_etree(const gmrf_b200_handle *h, int64_t *parent) {
    if (!h || !parent) return GMRF_B200_ERR_ARG;
    std::copy(h->S.parent.begin(), h->S.parent.end(), parent);
    return 0;
}
int gmrf_b200_get_supernodes(const gmrf_b200_handle *h, int64_t *super_ptr, int64_t *super_parent, int64_t *level,
                             int64_t *row_ptr, int64_t *panel_off, int64_t *panel_ld, int64_t *upd_off, int64_t *upd_ld) {
    if (!h) return GMRF_B200_ERR_ARG;
    const Symbolic &S = h->S;
    if (super_ptr) std::copy(S.sfirst.begin(), S.sfirst.end(), super_ptr);
    if (super_parent) std::copy(S.sparent.begin(), S.sparent.end(), super_parent);
    if (level) for (i64 s = 0; s < S.nsuper; s++) level[s] = S.level[s];
    if (row_ptr) std::copy(S.rowptr.begin(), S.rowptr.end(), row_ptr);
    if (panel_off) std::copy(S.panel_off.begin(), S.panel_off.end(), panel_off);
    if (panel_ld) for (i64 s = 0; s < S.nsuper; s++) panel_ld[s] = S.panel_ld[s];
    if (upd_off) std::copy(S.upd_off.begin(), S.upd_off.end(), upd_off);
    if (upd_ld) for (i64 s = 0; s < S.nsuper; s++) upd_ld[s] = S.upd_ld[s];
    return 0;
}
int gmrf_b200_get_rows(const gmrf_b200_handle *h, int64_t *row_idx, int64_t *rel_idx) {
    if (!h) return GMRF_B200_ERR_ARG;
    const Symbolic &S = h->S;
    if (row_idx) for (size_t k = 0; k < S.rowidx.size(); k++) row_idx[k] = S.rowidx[k];
    if (rel_idx) for (size_t k = 0; k < S.relidx.size(); k++) rel_idx[k] = S.relidx[k];
    return 0;
}
int gmrf_b200_get_scatter(const gmrf_b200_handle *h, int64_t *n_entries, int64_t *src, int64_t *dst) {
    if (!h) return GMRF_B200_ERR_ARG;
    const Symbolic &S = h->S;
    if (n_entries) *n_entries = (int64_t)S.q_src.size();
    if (src) std::copy(S.q_src.begin(), S.q_src.end(), src);
    if (dst) std::copy(S.q_dst.begin(), S.q_dst.end(), dst);
    return 0;
}
int gmrf_b200_last_timings(const gmrf_b200_handle *h, double *ms, int n) {
    if (!h || !ms) return GMRF_B200_ERR_ARG;
    for (int i = 0; i < n && i < 5; i++) ms[i] = h->t_ms[i];
    return 0;
}
int gmrf_b200_get_factor_panels(gmrf_b200_handle *h, double *Lx, int64_t n_doubles) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored || n_doubles != h->S.panel_total) { h->err = "get_factor_panels: bad state or size"; return GMRF_B200_ERR_ARG; }
    CUDA_TRY(h, cudaMemcpy(Lx, h->d_Lx, sizeof(double) * (size_t)n_doubles, cudaMemcpyDeviceToHost));
    return 0;
}
int gmrf_b200_get_selinv_panels(gmrf_b200_handle *h, double *Zx, int64_t n_doubles) {
    int rc = gmrf_b200_selinv_compute(h);
    if (rc) return rc;
    if (n_doubles != h->S.panel_total) { h->err = "get_selinv_panels: bad size"; return GMRF_B200_ERR_ARG; }
    CUDA_TRY(h, cudaMemcpy(Zx, h->d_Zx, sizeof(double) * (size_t)n_doubles, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- dense-kernel unit-test hooks (HOST pointers; operands are staged to `device` and back) ------------
static int test_fail(const char *msg) { g_create_error = msg; return GMRF_B200_ERR_CUDA; }

int gmrf_b200_test_gemm(int device, int transa, int transb, int lower, int m, int n, int k, const double *A, int lda,
                        const double *B, int ldb, double beta, double *C, int ldc) {
    if (cudaSetDevice(device) != cudaSuccess || configure_kernels() != cudaSuccess) return test_fail("no usable device");
    // operand extents in doubles
    size_t sa = (size_t)lda * (transa ? m : k), sb = (size_t)ldb * (transb ? n : k), sc = (size_t)ldc * n;
    double *dA, *dB, *dC;
    GemmTask *dT;
    int *dP;
    if (cudaMalloc(&dA, sa * 8 + 16) || cudaMalloc(&dB, sb * 8 + 16) || cudaMalloc(&dC, sc * 8 + 16) ||
        cudaMalloc(&dT, sizeof(GemmTask)) || cudaMalloc(&dP, 8))
        return test_fail("alloc");
    cudaMemcpy(dA, A, sa * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B, sb * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C, sc * 8, cudaMemcpyHostToDevice);
    GemmTask T;
    T.A = dA; T.B = dB; T.C = dC; T.m = m; T.n = n; T.k = k; T.lda = lda; T.ldb = ldb; T.ldc = ldc;
    T.flags = (lower & 1 ? GEMM_LOWER : 0) | (beta == 0.0 ? GEMM_BETA0 : 0) | (lower & 2 ? GEMM_ALPHA_POS : 0) | (lower & 4 ? GEMM_ADD_I : 0);
    T.pad_ = 0;
    const bool naive = (lower & 8) != 0, large = (lower & 16) != 0;
    int BM = naive ? 16 : large ? 128 : 64, BN = naive ? 16 : 64;
    int tiles = cdiv(m, BM) * cdiv(n, BN);
    int pf[2] = {0, tiles};
    cudaMemcpy(dT, &T, sizeof(T), cudaMemcpyHostToDevice);
    cudaMemcpy(dP, pf, 8, cudaMemcpyHostToDevice);
    if (!transa && !transb) launch_gemm<false, false>(large, naive, dT, dP, 1, tiles, 0);
    else if (!transa && transb) launch_gemm<false, true>(large, naive, dT, dP, 1, tiles, 0);
    else if (transa && transb) launch_gemm<true, true>(large, naive, dT, dP, 1, tiles, 0);
    else return test_fail("unsupported variant");
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(C, dC, sc * 8, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dT); cudaFree(dP);
    if (e != cudaSuccess) return test_fail(cudaGetErrorString(e));
    return 0;
}

int gmrf_b200_test_potrf_inv(int device, int n, double *A, int lda, double *inv, int *info) {
    if (cudaSetDevice(device) != cudaSuccess) return test_fail("no device");
    if (n > POTRF_NB || n < 1) return GMRF_B200_ERR_ARG;
    double *dA, *dI; PanelTask *dT; int *dF;
    if (cudaMalloc(&dA, (size_t)lda * n * 8) || cudaMalloc(&dI, (size_t)n * n * 8) || cudaMalloc(&dT, sizeof(PanelTask)) || cudaMalloc(&dF, 4))
        return test_fail("alloc");
    cudaMemcpy(dA, A, (size_t)lda * n * 8, cudaMemcpyHostToDevice);
    PanelTask T{dA, dI, lda, n, 0, 0};
    cudaMemcpy(dT, &T, sizeof(T), cudaMemcpyHostToDevice);
    cudaMemset(dF, 0x7f, 4);
    if (n <= 8) potrf_inv_kernel<8><<<1, 256>>>(dT, dF, 0LL);
    else if (n <= 16) potrf_inv_kernel<16><<<1, 256>>>(dT, dF, 0LL);
    else if (n <= 32) potrf_inv_kernel<32><<<1, 256>>>(dT, dF, 0LL);
    else if (global_options().potrf_lookahead) {
        if (configure_kernels() != cudaSuccess) return test_fail("kernel attributes");
        potrf_inv64_la_kernel<<<1, 256, POTRF_LA_SMEM_BYTES>>>(dT, dF, 0LL);
    } else potrf_inv64_kernel<<<1, 256>>>(dT, dF, 0LL);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(A, dA, (size_t)lda * n * 8, cudaMemcpyDeviceToHost);
    if (inv) cudaMemcpy(inv, dI, (size_t)n * n * 8, cudaMemcpyDeviceToHost);
    int f = 0;
    cudaMemcpy(&f, dF, 4, cudaMemcpyDeviceToHost);
    if (info) *info = (f == 0x7f7f7f7f) ? 0 : f;
    cudaFree(dA); cudaFree(dI); cudaFree(dT); cudaFree(dF);
    if (e != cudaSuccess) return test_fail(cudaGetErrorString(e));
    return 0;
}
// y[i] = rsqrt_inline(x[i]) (the straight-line reciprocal square root of the 4 x 4 pivot tiles), for the accuracy test
__global__ void test_rsqrt_kernel(const double *x, double *y, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) y[i] = rsqrt_inline(x[i]);
}
int gmrf_b200_test_rsqrt(int device, int n, const double *x, double *y) {
    if (cudaSetDevice(device) != cudaSuccess) return test_fail("no device");
    if (n < 1 || !x || !y) return GMRF_B200_ERR_ARG;
    double *dx, *dy;
    if (cudaMalloc(&dx, (size_t)n * 8) || cudaMalloc(&dy, (size_t)n * 8)) return test_fail("alloc");
    cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice);
    test_rsqrt_kernel<<<(n + 255) / 256, 256>>>(dx, dy, n);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(y, dy, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dy);
    if (e != cudaSuccess) return test_fail(cudaGetErrorString(e));
    return 0;
}
int gmrf_b200_test_potrf(int device, int n, double *A, int lda, int *info) {
    return gmrf_b200_test_potrf_inv(device, n, A, lda, nullptr, info);
}

// Device-timed GEMM micro-benchmark of the library's own kernel (tests/ and profiling only): operands are
// allocated and filled on the device; returns the best-of-reps time in ms (CUDA events on the default stream).
int gmrf_b200_bench_gemm(int device, int transa, int transb, int flags, int m, int n, int k, int reps, double *ms_out) {
    if (cudaSetDevice(device) != cudaSuccess || configure_kernels() != cudaSuccess) return test_fail("no usable device");
    const int lda = (transa ? k : m) + 2, ldb = (transb ? k : n) + 2, ldc = m + 2;
    size_t sa = (size_t)lda * (transa ? m : k), sb = (size_t)ldb * (transb ? n : k), sc = (size_t)ldc * n;
    double *dA, *dB, *dC; GemmTask *dT; int *dP;
    if (cudaMalloc(&dA, sa * 8) || cudaMalloc(&dB, sb * 8) || cudaMalloc(&dC, sc * 8) || cudaMalloc(&dT, sizeof(GemmTask)) || cudaMalloc(&dP, 8))
        return test_fail("alloc");
    cudaMemset(dA, 0, sa * 8); cudaMemset(dB, 0, sb * 8); cudaMemset(dC, 0, sc * 8);
    GemmTask T;
    T.A = dA; T.B = dB; T.C = dC; T.m = m; T.n = n; T.k = k; T.lda = lda; T.ldb = ldb; T.ldc = ldc;
    T.flags = (flags & 1 ? GEMM_LOWER : 0); T.pad_ = 0;
    const bool large = (flags & 16) != 0;
    int BM = large ? 128 : 64, BN = 64;
    int tiles = cdiv(m, BM) * cdiv(n, BN);
    int pf[2] = {0, tiles};
    cudaMemcpy(dT, &T, sizeof(T), cudaMemcpyHostToDevice);
    cudaMemcpy(dP, pf, 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, 0);
        if (!transa && !transb) launch_gemm<false, false>(large, false, dT, dP, 1, tiles, 0);
        else if (!transa && transb) launch_gemm<false, true>(large, false, dT, dP, 1, tiles, 0);
        else launch_gemm<true, true>(large, false, dT, dP, 1, tiles, 0);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) best = std::min(best, ms);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dT); cudaFree(dP);
    if (e != cudaSuccess) return test_fail(cudaGetErrorString(e));
    if (ms_out) *ms_out = best;
    return 0;
}

// Live per-kernel-family profile of ONE refactorization (graphs off, a CUDA event pair around every launch on the
// handle's stream). kinds: 0 gemm (DMMA, incl. TRSM-by-inverse), 1 panel (potrf + inverse), 2 assemble (extend-add), 3 scatter/memset/logdet.
// ms[k] = summed device time, count[k] = launches, flops[0] = algorithmic GEMM flops issued (lower-only tasks count half).
int gmrf_b200_profile_refactorize(gmrf_b200_handle *h, double *ms, int64_t *count, double *flops) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored) { h->err = "profile_refactorize needs a previous refactorize (values resident)"; return GMRF_B200_ERR_STATE; }
    const Symbolic &S = h->S;
    cudaStream_t st = h->stream;
    for (int k = 0; k < 4; k++) { ms[k] = 0; count[k] = 0; }
    std::vector<cudaEvent_t> evs;
    std::vector<int> kinds;
    auto mark = [&](int kind) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        evs.push_back(e);
        kinds.push_back(kind);
    };
    mark(3);
    cudaMemsetAsync(h->d_Lx, 0, sizeof(double) * (size_t)S.panel_total, st);
    cudaMemsetAsync(h->d_fail, 0x7f, sizeof(int), st);
    i64 cnt = (i64)S.q_src.size();
    if (cnt > 0) scatter_q_kernel<<<(int)std::min<i64>((cnt + 255) / 256, 148 * 16), 256, 0, st>>>(h->d_Lx, h->d_nz, h->d_qsrc, h->d_qdst, cnt, 0LL);
    TableSet T{h->d_gemm, h->d_items, h->d_prefix, h->d_split};
    for (const Launch &L : h->factor_plan.launches) {
        int kind = (L.kind >= K_GEMM_NN_S && L.kind <= K_GEMM_TT_L) ? 0
                   : (L.kind == K_PANEL || L.kind == K_CHAIN || L.kind == K_FRONT || L.kind == K_FINALIZE) ? 1
                   : (L.kind == K_ASSEMBLE || L.kind == K_ASSEMBLE_G) ? 2 : 3;
        mark(kind);
        run_launch(h, L, T, 0);
    }
    mark(3);
    logdet_partial_kernel<<<LOGDET_BLOCKS, 256, 0, st>>>(h->d_Lx, h->d_diagpos, S.n, h->d_partial, 0LL);
    logdet_final_kernel<<<1, 256, 0, st>>>(h->d_partial, h->d_scalars, 0LL);
    mark(-1);
    CUDA_TRY(h, cudaStreamSynchronize(st));
    for (size_t i = 0; i + 1 < evs.size(); i++) {
        float t = 0;
        cudaEventElapsedTime(&t, evs[i], evs[i + 1]);
        ms[kinds[i]] += t;
        count[kinds[i]] += 1;
    }
    for (auto e : evs) cudaEventDestroy(e);
    if (flops) *flops = h->gemm_flops_factor;
    if ((rc = check_launch(h, "profile_refactorize"))) return rc;
    return 0;
}

// Per-launch device times of one phase (graphs off, a CUDA event pair around every launch on the handle's stream):
// phase 0 = factorization plan, 1 = selected-inversion plan, 2 = forward sweep, 3 = backward sweep (nrhs columns of
// whatever the work array holds). Fills up to `cap` entries of kind[] (LaunchKind), grid[] (CTAs), ms[]; *count = launches.
// Profiling/diagnostics only: phases 0 and 1 recompute from the resident values, phases 2/3 leave the work array dirty.
int gmrf_b200_profile_plan(gmrf_b200_handle *h, int phase, int nrhs, int64_t cap, int *kind, int *grid, double *ms, int64_t *count) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored) { h->err = "profile_plan needs a previous refactorize"; return GMRF_B200_ERR_STATE; }
    const Plan *plan = nullptr;
    TableSet T{h->d_gemm, h->d_items, h->d_prefix, h->d_split};
    cudaStream_t st = h->stream;
    if (phase == 0) {
        plan = &h->factor_plan;
        // (a lane handle is profiled with the lanes of its last sweep: same grids as the sweep itself)
        T.lanes = std::max(1, h->last_lanes);
        const long long bstride = T.lanes > 1 ? h->arena_bytes : 0;
        for (int b = 0; b < T.lanes; b++) {
            cudaMemsetAsync(reinterpret_cast<char *>(h->d_Lx) + (long long)b * h->arena_bytes, 0, sizeof(double) * (size_t)h->S.panel_total, st);
            cudaMemsetAsync(reinterpret_cast<char *>(h->d_fail) + (long long)b * h->arena_bytes, 0x7f, sizeof(int), st);
        }
        i64 cnt = (i64)h->S.q_src.size();
        if (cnt > 0) scatter_q_kernel<<<dim3((int)std::min<i64>((cnt + 255) / 256, 148 * 16), T.lanes), 256, 0, st>>>(h->d_Lx, h->d_nz, h->d_qsrc, h->d_qdst, cnt, bstride);
        h->selinv_valid = false;
    } else if (phase == 1) {
        if ((rc = build_selinv_tables(h))) return rc;
        plan = &h->selinv_plan;
        T = TableSet{h->d_gemm_z, h->d_items_z, h->d_prefix_z, h->d_split_z};
    } else if (phase == 2 || phase == 3) {
        plan = phase == 2 ? &h->fwd_plan : &h->bwd_plan;
        if (nrhs < 1 || nrhs > h->rhs_block) { h->err = "profile_plan: 1 <= nrhs <= 8"; return GMRF_B200_ERR_ARG; }
    } else {
        h->err = "profile_plan: phase must be 0..3";
        return GMRF_B200_ERR_ARG;
    }
    const size_t nl = plan->launches.size();
    std::vector<cudaEvent_t> evs(nl + 1);
    for (auto &e : evs) cudaEventCreate(&e);
    for (size_t i = 0; i < nl; i++) {
        cudaEventRecord(evs[i], st);
        run_launch(h, plan->launches[i], T, nrhs);
    }
    cudaEventRecord(evs[nl], st);
    CUDA_TRY(h, cudaStreamSynchronize(st));
    for (size_t i = 0; i < nl; i++) {
        float t = 0;
        cudaEventElapsedTime(&t, evs[i], evs[i + 1]);
        if ((int64_t)i < cap) {
            if (kind) kind[i] = plan->launches[i].kind;
            if (grid) grid[i] = plan->launches[i].grid;
            if (ms) ms[i] = t;
        }
    }
    for (auto e : evs) cudaEventDestroy(e);
    if (count) *count = (int64_t)nl;
    if (phase == 0) {
        logdet_partial_kernel<<<LOGDET_BLOCKS, 256, 0, st>>>(h->d_Lx, h->d_diagpos, h->S.n, h->d_partial, 0LL);
        logdet_final_kernel<<<1, 256, 0, st>>>(h->d_partial, h->d_scalars, 0LL);
        CUDA_TRY(h, cudaStreamSynchronize(st));
    }
    return check_launch(h, "profile_plan");
}

// Static facts of the launches of a phase (same order as profile_plan): flops = algorithmic flops of GEMM launches, kmax
// = their longest contraction, ntasks = tasks in the launch.
int gmrf_b200_plan_launch_info(gmrf_b200_handle *h, int phase, int64_t cap, double *flops, int *kmax, int *ntasks, int64_t *count) {
    if (!h || !count) return GMRF_B200_ERR_ARG;
    const Plan *plan = phase == 0 ? &h->factor_plan : phase == 1 ? &h->selinv_plan : phase == 2 ? &h->fwd_plan : phase == 3 ? &h->bwd_plan : nullptr;
    if (!plan) { h->err = "plan_launch_info: phase must be 0..3"; return GMRF_B200_ERR_ARG; }
    *count = (int64_t)plan->launches.size();
    for (size_t i = 0; i < plan->launches.size() && (int64_t)i < cap; i++) {
        if (flops) flops[i] = plan->launches[i].flops;
        if (kmax) kmax[i] = plan->launches[i].kmax;
        if (ntasks) ntasks[i] = plan->launches[i].ntasks;
    }
    return 0;
}

// clock64 stamps of the phases of the last fused chain launch of one refactorization (tile 1): [0] start, [1] tiles
// loaded, [2] first 64-column panel done, [3] its rows stored, [4] rank-64 updates done, [5] second panel done, [6] end.
int gmrf_b200_debug_chain_phases(gmrf_b200_handle *h, int64_t *stamps, int n) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!h->factored || !stamps || n < 7) { h->err = "debug_chain_phases: needs a factored handle and 7 slots"; return GMRF_B200_ERR_ARG; }
    long long *d = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&d, 24 * sizeof(long long)));
    cudaMemset(d, 0, 24 * sizeof(long long));
    CUDA_TRY(h, cudaMemcpyToSymbol(g_chain_prof, &d, sizeof(d)));
    const int ug = h->opt.use_graph;
    h->opt.use_graph = 0;
    rc = do_factor(h);
    h->opt.use_graph = ug;
    long long host[24];
    cudaMemcpy(host, d, sizeof(host), cudaMemcpyDeviceToHost);
    long long *null = nullptr;
    cudaMemcpyToSymbol(g_chain_prof, &null, sizeof(null));
    cudaFree(d);
    for (int i = 0; i < 7; i++) stamps[i] = host[i];
    for (int i = 7; i < n && i < 24; i++) stamps[i] = host[i];        // [8..16]: inside the blocked panel (front_kernels.cuh)
    return rc < 0 ? rc : 0;
}

// ---- factor sharing between the handles of a multi-GPU pool ------------------------------------------------
// Device pointers and lengths (doubles) of the numeric state a solve needs: which = 0 supernodal panels of L,
// 1 inverted diagonal blocks, 2 selected-inverse panels (after selinv_compute). A pool factorizes on ONE device,
// moves these arrays to its peers (ncclBroadcast over NVLink, or any device copy) and calls adopt_factor there.
int gmrf_b200_device_array(gmrf_b200_handle *h, int which, void **ptr, int64_t *n_doubles) {
    int rc = ensure_device(h);
    if (rc) return rc;
    if (!ptr || !n_doubles) { h->err = "device_array: null output"; return GMRF_B200_ERR_ARG; }
    const Symbolic &S = h->S;
    switch (which) {
        case 0: *ptr = h->d_Lx; *n_doubles = S.panel_total; break;
        case 1: {
            i64 tot = 0;
            for (i64 s = 0; s < S.nsuper; s++)
                for (i64 k0 = 0; k0 < S.ns(s); k0 += SOLVE_NB) { i64 nb = std::min<i64>(SOLVE_NB, S.ns(s) - k0); tot += nb * nb; }
            *ptr = h->d_Linv; *n_doubles = tot;
            break;
        }
        case 2:
            if ((rc = build_selinv_tables(h))) return rc;
            *ptr = h->d_Zx; *n_doubles = S.panel_total;
            break;
        default: h->err = "device_array: which must be 0, 1 or 2"; return GMRF_B200_ERR_ARG;
    }
    return 0;
}

// Declare the panels / inverted blocks currently in this handle's HBM (received from a peer that factorized the same
// pattern with the same ordering) to be its numeric factor; `logdet` is the sender's log-determinant. with_selinv != 0
// also adopts the selected-inverse panels.
int gmrf_b200_adopt_factor(gmrf_b200_handle *h, double logdet, int with_selinv) {
    int rc = ensure_device(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->logdet = logdet;
    h->fail_col = 0;
    h->factored = true;
    h->selinv_valid = with_selinv != 0 && h->selinv_tables_built;
    return 0;
}

// Fingerprint of everything a received factor must agree on with the handle that adopts it: the pattern, the elimination
// order, the supernode partition and the panel / inverse-block layout (which also encode the relaxation and blocking options).
int gmrf_b200_analysis_fingerprint(const gmrf_b200_handle *h, uint64_t *out) {
    if (!h || !out) return GMRF_B200_ERR_ARG;
    const Symbolic &S = h->S;
    unsigned long long f = 1469598103934665603ULL ^ h->pattern_hash;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; i++) { f ^= c[i]; f *= 1099511628211ULL; }
    };
    mix(S.perm.data(), S.perm.size() * sizeof(i64));
    mix(S.sfirst.data(), S.sfirst.size() * sizeof(i64));
    mix(S.panel_off.data(), S.panel_off.size() * sizeof(i64));
    mix(S.panel_ld.data(), S.panel_ld.size() * sizeof(i32));
    mix(h->inv_base.data(), h->inv_base.size() * sizeof(long long));
    *out = f;
    return 0;
}

// adopt_factor with the sender's fingerprint and pivot status: a peer built with another ordering / other options is
// refused (GMRF_B200_ERR_ARG) instead of silently solving with foreign panels, and a non-positive pivot travels along.
int gmrf_b200_adopt_factor_checked(gmrf_b200_handle *h, uint64_t sender_fingerprint, double logdet, int sender_status, int with_selinv) {
    int rc = ensure_device(h);
    if (rc) return rc;
    uint64_t mine = 0;
    gmrf_b200_analysis_fingerprint(h, &mine);
    if (mine != sender_fingerprint) {
        h->err = "adopt_factor: the sender's symbolic analysis (pattern / ordering / supernodes / panel layout) differs from this handle's";
        return GMRF_B200_ERR_ARG;
    }
    rc = gmrf_b200_adopt_factor(h, logdet, with_selinv);
    if (rc) return rc;
    h->fail_col = sender_status > 0 ? sender_status : 0;
    return h->fail_col;
}

// Page-lock / unlock a caller-owned host buffer (e.g. the workspace's nzval array) so refactorize() copies it with
// a true asynchronous DMA instead of a staged pageable copy. Purely an optimisation; the buffer stays caller-owned.
int gmrf_b200_host_register(void *ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return GMRF_B200_ERR_ARG;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = cudaGetErrorString(e); return GMRF_B200_ERR_CUDA; }
    return 0;
}
int gmrf_b200_host_unregister(void *ptr) {
    if (!ptr) return GMRF_B200_ERR_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return GMRF_B200_ERR_CUDA; }
    return 0;
}


// dst[0..n) = src[0..n) with all host threads: the workspace mirror `ws.Q.nzval .= nzval` of update_precision_values
// (gmrf_workspace.jl:154-165) is a 520 MB host copy at 1 M dofs -- 35 ms on one thread, inside every end-to-end step.
int gmrf_b200_host_copy(double *dst, const double *src, int64_t n) {
    if (n < 0 || (n > 0 && (!dst || !src))) return GMRF_B200_ERR_ARG;
    const int64_t chunk = 1 << 18;                     // 2 MB pieces
    const int64_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(static) if (nchunks > 4)
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t lo = c * chunk, len = std::min<int64_t>(chunk, n - lo);
        std::memcpy(dst + lo, src + lo, (size_t)len * sizeof(double));
    }
    return 0;
}

}  // extern "C"
