// symbolic.hpp -- host-side symbolic analysis for the B200 supernodal multifrontal Cholesky.
//
// This is the work the north star leaves on the CPU, done once per sparsity pattern and cached in the
// handle (reference: the symbolic half of `cholesky(Q; perm)` in src/workspace/backend.jl:147-153 and
// the "symbolic factorization is computed once and reused" contract of backend.jl:32-50):
// fill-reducing ordering, elimination tree, exact column counts, supernode partition with relaxed
// amalgamation, per-supernode row structures, assembly-tree levels, relative indices for extend-add,
// the Q.nzval -> panel scatter map (cf. _build_full_to_lower_map, cliquetrees_backend.jl:90-123) and the
// lifetime-packed pools for update matrices.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gmrf {

typedef int64_t i64;
typedef int32_t i32;

struct Options {
    // relaxed amalgamation: merge a child into its parent if merged ns <= relax_n[0]; else if the zero
    // fraction of the merged panel is < relax_z[k] for the first k with ns <= relax_n[k]; relax_z[2]
    // applies beyond relax_n[2] ... (the shape of CHOLMOD's nrelax/zrelax rule, GPU-tuned values).
    double relax_n[3] = {8, 32, 96};
    double relax_z[4] = {0.8, 0.3, 0.1, 0.05};
    int use_graph = 1;
    int outer_block = 256;   // outer panel block (columns) of the two-level blocked factorization
    int naive_kernels = 0;
    int splitk_min_k = 128;   // split-K: a k-slice is at least this long (1024 in round 1; 128 measured 4-7 % faster on the 2D configs, neutral at 1 M dofs: profiles/r02_sweep_opts_splitk.log)
    int wide_rhs_min = 8;     // more right-hand sides than this take the GEMM (wide) solve path in blocks of 64 columns
    int bwd_row_chunk = 2048; // backward sweep: rows of L21 per partial task (taller panels are reduced in a second pass)
    int large_tile_mask = 1;  // GEMM operand-layout variants (bit 0 NN, 1 NT, 2 TT) allowed to use the 128 x 64 tile (measured: only NN gains)
    int lanes = 1;            // value sets a handle can factorize side by side (batched hyperparameter evaluations)
    int selinv_fast_root = 1; // triangular (trtri + lauum) route for top-level root supernodes in the selected inversion
    int syrk_gather = 0;      // update-matrix products gather their children's contributions in a tail of the GEMM (U written once). Measured at 1 M dofs (profiles/r02_plan_3d100_syrk_gather.log): extend-add 60 -> 33 ms but the products +50 ms (the tail is not hidden behind the other CTAs), so off by default
    int wide_steps = 0;       // few-RHS triangular solves advance 256 columns per step on supernodes with more than 256 columns (two launches per step: update + one-CTA diagonal-block solve). Measured at 1 M dofs (profiles/r02_plan_solve_3d100_wide_vs_narrow.log): 20.3 ms against 18.7 ms for the 64-column steps with their look-ahead block solve -- the serial diagonal kernels (15 us x 222 per direction) cost more than the halved launch count saves; off by default
    int pdl = 1;              // the kernels of the few-RHS sweeps are launched with programmatic stream serialization (griddepcontrol): the next step's CTAs load their factor tiles while the current step finishes and wait for its completion before touching the unknowns. Measured at 1 M dofs: 18.70 -> 15.71 ms per solve, bit-identical (profiles/r02_solve_pdl.log)
    int pdl_factor = 0;       // same for the kernels of the factorization / selected inversion plans: only task tables and symbolic maps can be read ahead there, so it hides launch latency (2D configs), not HBM latency
    int pdl_multi = 1;        // same for the GEMM sweeps of the wide right-hand-side blocks: operand A of every product is a piece of the factor, its first tiles are loaded before the wait (64 columns: -12 % at 50 k dofs, -8 % at 117 k, -1.6 % at 1 M; bit-identical; profiles/r02_multi_pdl.log)
    int panel_blocked = 1;    // 1: the one-CTA front kernel factors its 64-column panels 16 columns at a time (rank-4 steps on one register tile per thread inside the sub-panel, one rank-16 DMMA update to its right); 2: the fused chain steps as well (measured slower there: the rank-16 update of a 192-row panel through shared memory costs more than it saves); 0: rank-4 updates of the whole panel after every 4 columns everywhere
    int potrf_lookahead = 1;  // bulk path: 64 x 64 diagonal blocks factored AND inverted in one pass of the blocked panel code (identity as a row tile); 0 = the round-1 kernel (no look-ahead, 64-thread substitution for the inverse)
    int syrk_split = 0;       // allow split-K on the update-matrix products as well (few-tile launches at the top of 2D trees)
    int level_alap = 1;       // assembly-tree levels counted from the roots (as late as possible) instead of from the leaves
    int asm_gather = 1;       // extend-add as a gather through TMA-staged shared memory (0: the first, scatter-shaped kernel)
    int fused_front = 1;      // one-CTA-per-front kernel for tree levels whose panels all fit in shared memory
    int fused_chain = 1;      // one launch per 128 columns of a level's supernode chains (redundant diagonal factorization per CTA)
    int chain_max_tiles = 160;   // a level takes the fused chain path if its first step has at most this many 64-row tiles
    int front_smem_kb = 200;  // shared-memory budget of the one-CTA-per-front kernel
};
Options &global_options();

struct Symbolic {
    i64 n = 0, nnzA = 0;
    std::vector<i64> perm, iperm;        // final elimination order (ordering o etree postorder)
    std::vector<i64> parent;             // etree, final order
    std::vector<i64> colcount;           // exact column counts of L incl. diagonal, final order
    i64 nnzL = 0;                        // sum colcount
    double flops = 0;                    // sum colcount^2

    i64 nsuper = 0;
    std::vector<i64> sfirst;             // [nsuper+1] first column of each supernode
    std::vector<i64> sparent;            // [nsuper]  parent supernode or -1
    std::vector<i64> col2super;          // [n]
    std::vector<i64> rowptr;             // [nsuper+1] offsets into rowidx
    std::vector<i32> rowidx;             // row structure, sorted; first ns entries are the own columns
    std::vector<i32> relidx;             // same offsets as rowidx: position of the row in the PARENT's structure
                                         // (valid for the below-diagonal rows of non-root supernodes)
    std::vector<i64> panel_off;          // [nsuper+1] offset (doubles) of the nrow x ns panel, column-major
    std::vector<i32> panel_ld;           // [nsuper]
    i64 panel_total = 0;
    double flops_stored = 0;

    std::vector<i64> child_ptr, child_idx;   // children lists (ascending)
    std::vector<i32> level;                  // height above the leaves
    i64 nlevels = 0;
    std::vector<i64> level_ptr, level_idx;   // supernodes grouped by level

    // update-matrix pool (nr x nr, ld = upd_ld), packed by lifetime [level(s), level(parent)]
    std::vector<i64> upd_off;
    std::vector<i32> upd_ld;
    i64 upd_total = 0;
    // selected-inversion pool: W_s = Z[R_s, R_s] (nr x nr full symmetric), lifetime
    // [min level of children, level(s)], processed top-down
    std::vector<i64> zw_off;
    i64 zw_total = 0;
    // forward-solve pool in units of rows (u_s has nr_s rows per right-hand side)
    std::vector<i64> uvec_off;
    i64 uvec_total = 0;

    // scatter of the upper triangle of Q into the panels: Lx[q_dst[k]] = nzval[q_src[k]]
    std::vector<i64> q_src, q_dst;
    std::vector<i64> diag_pos;           // [n] position of L_jj in the panel array

    i64 max_front = 0, max_ns = 0;
    double analysis_ms = 0;

    inline i64 ns(i64 s) const { return sfirst[s + 1] - sfirst[s]; }
    inline i64 nrow(i64 s) const { return rowptr[s + 1] - rowptr[s]; }
    inline i64 nr(i64 s) const { return nrow(s) - ns(s); }
};

// colptr/rowval: 0-based full symmetric pattern. user_perm may be null. Throws std::runtime_error.
void analyze(Symbolic &S, i64 n, const i64 *colptr, const i64 *rowval, const i64 *user_perm, int ordering,
             const Options &opt);

// Offset in the panel array of entry (i, j) = (j, i) of a symmetric matrix living on the factor's stored pattern, for
// every entry of an n x n CSC pattern (pos[p] = -1 outside the pattern). Columns are independent: OpenMP over columns.
// Returns -1, or the index p of the first entry whose row index is out of range.
i64 pattern_positions(const Symbolic &S, const i64 *colptr, const i64 *rowval, i64 index_base, long long *pos);

// Offset in the panel array of entry (i, j) = (j, i) (ORIGINAL 0-based indices) of a symmetric matrix on the factor's stored
// pattern, -1 outside it.
long long entry_position(const Symbolic &S, i64 i, i64 j);

// Persisting the analysis (the only state worth keeping across sessions / sharing between the handles of a pool):
// a self-describing little-endian byte stream of every member of Symbolic plus a hash of the pattern it belongs to.
// deserialize throws std::runtime_error on a malformed / truncated stream or a pattern mismatch.
unsigned long long pattern_hash(i64 n, const i64 *colptr, const i64 *rowval);   // 0-based CSC pattern
void serialize(const Symbolic &S, unsigned long long pattern_hash, std::vector<char> &out);
void deserialize(Symbolic &S, const char *data, size_t len, i64 n, i64 nnz, unsigned long long pattern_hash);
bool equal(const Symbolic &a, const Symbolic &b);   // every member (analysis_ms excepted)

// orderings (0-based adjacency without self loops: xadj[n+1], adj[])
void order_metis_nd(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, std::vector<i64> &perm);
void order_amd(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, std::vector<i64> &perm);

}  // namespace gmrf
