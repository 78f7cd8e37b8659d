// symbolic.cpp -- see symbolic.hpp. Integer graph work only; no floating-point matrix values are touched.
#include "symbolic.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>

// METIS ships inside the CUDA toolkit as a static archive without a header
// (/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a, 64-bit idx_t); prototypes are hand-declared.
extern "C" {
int METIS_NodeND(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *options,
                 int64_t *perm, int64_t *iperm);
int METIS_SetDefaultOptions(int64_t *options);
}

namespace gmrf {

Options &global_options() {
    static Options o;
    return o;
}

void order_metis_nd(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, std::vector<i64> &perm) {
    perm.resize(n);
    if (n == 0) return;
    if (adj.empty()) {
        std::iota(perm.begin(), perm.end(), 0);
        return;
    }
    std::vector<i64> xa(xadj), ad(adj), ip(n);
    i64 opts[40];
    METIS_SetDefaultOptions(opts);
    i64 nn = n;
    int rc = METIS_NodeND(&nn, xa.data(), ad.data(), nullptr, opts, perm.data(), ip.data());
    if (rc != 1) throw std::runtime_error("METIS_NodeND failed");
    // METIS: A(perm, perm) is the reordered matrix, i.e. perm[k] = original index of the k-th pivot.
}

// ------------------------------------------------------------------------------------------------
// Approximate minimum degree on the quotient graph (Amestoy, Davis & Duff 1996, restated from the paper): elements
// replace eliminated pivots, the external degree of a variable is bounded by
//     d_i <= min( n_left - nv_i,  d_i(old) + |L_p \ i|,  |A_i \ i| + |L_p \ i| + sum_{e in E_i \ p} |L_e \ L_p| ),
// the set differences |L_e \ L_p| come from one counting pass over the new element, elements contained in L_p are
// absorbed (aggressive absorption), variables whose quotient-graph adjacency coincides are merged into supervariables
// (found by hashing) and variables adjacent to nothing but the new element are eliminated with it (mass elimination).
void order_amd(i64 n, const std::vector<i64> &xadj, const std::vector<i64> &adj, std::vector<i64> &perm) {
    perm.clear();
    perm.reserve(n);
    if (n == 0) return;
    typedef int32_t I;
    std::vector<std::vector<I>> A(n), E(n), Le(n);      // variable lists, element lists, element patterns
    for (i64 v = 0; v < n; v++) {
        A[v].reserve(xadj[v + 1] - xadj[v]);
        for (i64 p = xadj[v]; p < xadj[v + 1]; p++)
            if (adj[p] != v) A[v].push_back((I)adj[p]);
    }
    enum : char { VAR = 0, ELEM = 1, DEAD = 2, MERGED = 3 };
    std::vector<char> state(n, VAR);
    std::vector<i64> nv(n, 1), degree(n), elem_deg(n, 0), w(n, 0), inLp(n, -1), hashv(n, 0);
    std::vector<I> head(n + 1, -1), next(n, -1), prev(n, -1), hhead(n, -1), hnext(n, -1);
    std::vector<std::vector<I>> members(n);              // variables merged into / eliminated with a principal variable
    auto list_insert = [&](I v) {
        i64 d = std::min<i64>(degree[v], n);
        prev[v] = -1; next[v] = head[d];
        if (head[d] != -1) prev[head[d]] = v;
        head[d] = v;
    };
    auto list_remove = [&](I v) {
        i64 d = std::min<i64>(degree[v], n);
        if (prev[v] != -1) next[prev[v]] = next[v]; else head[d] = next[v];
        if (next[v] != -1) prev[next[v]] = prev[v];
    };
    for (i64 v = 0; v < n; v++) { degree[v] = (i64)A[v].size(); list_insert((I)v); }
    i64 mindeg = 0, wflg = 1, nleft = n, pstamp = 0;
    std::vector<I> Lp, order_heads;
    order_heads.reserve(n);
    auto emit = [&](I root) {                            // root followed by everything merged into it (iteratively)
        std::vector<I> stack{root};
        while (!stack.empty()) {
            I v = stack.back(); stack.pop_back();
            perm.push_back(v);
            for (I m : members[v]) stack.push_back(m);
        }
    };
    while (nleft > 0) {
        while (mindeg <= n && head[mindeg] == -1) mindeg++;
        const I p = head[mindeg];
        list_remove(p);
        // ---- 1. new element p: L_p = (A_p  U  union of L_e, e in E_p) \ {p}, principal live variables only -------
        pstamp++;
        Lp.clear();
        i64 degp = 0;
        inLp[p] = pstamp;
        for (I u : A[p])
            if (state[u] == VAR && nv[u] > 0 && inLp[u] != pstamp) { inLp[u] = pstamp; Lp.push_back(u); degp += nv[u]; }
        for (I e : E[p]) {
            if (state[e] != ELEM) continue;
            for (I u : Le[e])
                if (state[u] == VAR && nv[u] > 0 && inLp[u] != pstamp) { inLp[u] = pstamp; Lp.push_back(u); degp += nv[u]; }
            state[e] = DEAD;                              // absorbed into p
            std::vector<I>().swap(Le[e]);
        }
        std::vector<I>().swap(A[p]);
        std::vector<I>().swap(E[p]);
        state[p] = ELEM;
        const i64 nvp = nv[p];
        nleft -= nvp;
        // ---- 2. |L_e \ L_p| for every element adjacent to a variable of L_p ----------------------------------------
        if (wflg + n + 2 < wflg) { std::fill(w.begin(), w.end(), 0); wflg = 1; }
        for (I i : Lp)
            for (I e : E[i]) {
                if (state[e] != ELEM) continue;
                if (w[e] >= wflg) w[e] -= nv[i];
                else w[e] = elem_deg[e] + wflg - nv[i];
            }
        // ---- 3. update the variables of L_p -----------------------------------------------------------------------
        std::vector<I> touched_hash;
        for (I i : Lp) {
            list_remove(i);
            i64 deg = 0;
            u_int64_t hsh = 0;
            auto &Ei = E[i];
            size_t k = 0;
            for (size_t t = 0; t < Ei.size(); t++) {
                I e = Ei[t];
                if (state[e] != ELEM) continue;
                const i64 dext = w[e] - wflg;
                if (dext > 0) { deg += dext; Ei[k++] = e; hsh += (u_int64_t)e; }
                else { state[e] = DEAD; std::vector<I>().swap(Le[e]); }   // aggressive absorption: L_e inside L_p
            }
            Ei.resize(k);
            auto &Ai = A[i];
            k = 0;
            for (size_t t = 0; t < Ai.size(); t++) {
                I j = Ai[t];
                if (state[j] != VAR || nv[j] <= 0 || inLp[j] == pstamp) continue;   // covered by element p or gone
                deg += nv[j];
                Ai[k++] = j;
                hsh += (u_int64_t)j;
            }
            Ai.resize(k);
            if (Ei.empty() && Ai.empty()) {
                // mass elimination: i is adjacent to nothing but p -> eliminated together with p
                degp -= nv[i];
                nleft -= nv[i];
                members[p].push_back(i);
                nv[i] = 0;
                state[i] = MERGED;
                continue;
            }
            Ei.push_back(p);
            hsh += (u_int64_t)p;
            degree[i] = std::min<i64>(degree[i], deg);   // first two terms of the bound; |L_p \ i| added below
            hashv[i] = (i64)(hsh % (u_int64_t)n);
            I hb = (I)hashv[i];
            if (hhead[hb] == -1) touched_hash.push_back(hb);
            hnext[i] = hhead[hb];
            hhead[hb] = i;
        }
        // ---- 4. supervariables: identical quotient-graph adjacency --------------------------------------------------
        wflg += n + 1;                                    // step-2 counters live in [wflg, wflg + n]: step past them
        for (I hb : touched_hash) {
            for (I i = hhead[hb]; i != -1; i = hnext[i]) {
                if (nv[i] <= 0) continue;
                // tag the adjacency of i
                wflg++;
                for (I e : E[i]) w[e] = wflg;
                for (I j : A[i]) w[j] = wflg;
                I prevj = i;
                for (I j = hnext[i]; j != -1; j = hnext[j]) {
                    if (nv[j] <= 0 || E[j].size() != E[i].size() || A[j].size() != A[i].size()) { prevj = j; continue; }
                    bool same = true;
                    for (I e : E[j]) if (w[e] != wflg) { same = false; break; }
                    if (same) for (I x : A[j]) if (w[x] != wflg) { same = false; break; }
                    if (same) {
                        nv[i] += nv[j];
                        nv[j] = 0;
                        state[j] = MERGED;
                        members[i].push_back(j);
                        std::vector<I>().swap(A[j]);
                        std::vector<I>().swap(E[j]);
                        hnext[prevj] = hnext[j];           // unlink j from the bucket
                    } else {
                        prevj = j;
                    }
                }
            }
            hhead[hb] = -1;
        }
        wflg += n + 1;                                    // invalidate every w[] stamp of this round
        // ---- 5. final element list, degrees back into the lists ---------------------------------------------------
        size_t k = 0;
        for (I i : Lp) {
            if (nv[i] <= 0) continue;
            Lp[k++] = i;
        }
        Lp.resize(k);
        i64 lp_weight = 0;
        for (I i : Lp) lp_weight += nv[i];
        for (I i : Lp) {
            i64 d = degree[i] + lp_weight - nv[i];
            d = std::min<i64>(d, nleft - nv[i]);
            degree[i] = std::max<i64>(d, 0);
            list_insert(i);
            if (degree[i] < mindeg) mindeg = degree[i];
        }
        elem_deg[p] = lp_weight;
        Le[p] = Lp;
        if (Lp.empty()) state[p] = DEAD;
        order_heads.push_back(p);
        (void)degp;
    }
    for (I p : order_heads) emit(p);
    if ((i64)perm.size() != n) throw std::runtime_error("order_amd: internal error (incomplete permutation)");
}

// ------------------------------------------------------------------------------------------------
namespace {

struct UpperCSC {  // upper triangle (row <= col) of the permuted matrix, column k = row pattern seeds
    std::vector<i64> ptr, idx;
};

// Counting sort of the (destination column, value) pairs that `emit(j, sink)` produces for the source columns
// j = 0 .. n_src-1, in parallel and in EXACTLY the order of the serial double loop: the source columns are cut into one
// contiguous chunk per thread, every thread counts its pairs per destination column, the per-thread counts become write
// offsets (threads in chunk order), and a second enumeration writes the values. Deterministic for any thread count.
template <class Emit>
void bucket_by_column(i64 n_src, i64 n_dst, Emit emit, std::vector<i64> &ptr, std::vector<i64> &idx) {
    int T = 1;
#ifdef _OPENMP
    T = std::max(1, omp_get_max_threads());
#endif
    if (n_src < 100000 || (double)T * (double)n_dst > 4e8) T = 1;          // small inputs / huge histograms: serial
    std::vector<std::vector<i64>> hist((size_t)T);
    ptr.assign((size_t)n_dst + 1, 0);
#pragma omp parallel num_threads(T)
    {
        int t = 0, Tn = 1;                                                   // the team the runtime actually gave us
#ifdef _OPENMP
        t = omp_get_thread_num();
        Tn = omp_get_num_threads();
#endif
        const i64 a = n_src * t / Tn, b = n_src * (t + 1) / Tn;
        std::vector<i64> &h = hist[(size_t)t];
        h.assign((size_t)n_dst, 0);
        for (i64 j = a; j < b; j++) emit(j, [&](i64 dst, i64) { h[(size_t)dst]++; });
#pragma omp barrier
#pragma omp for schedule(static)
        for (i64 d = 0; d < n_dst; d++) {                                    // counts -> offsets within the column
            i64 run = 0;
            for (int u = 0; u < Tn; u++) { const i64 c = hist[(size_t)u][(size_t)d]; hist[(size_t)u][(size_t)d] = run; run += c; }
            ptr[(size_t)d + 1] = run;
        }
#pragma omp single
        {
            for (i64 d = 0; d < n_dst; d++) ptr[(size_t)d + 1] += ptr[(size_t)d];
            idx.resize((size_t)ptr[(size_t)n_dst]);
        }
        for (i64 j = a; j < b; j++) emit(j, [&](i64 dst, i64 val) { idx[(size_t)(ptr[(size_t)dst] + h[(size_t)dst]++)] = val; });
    }
}

void build_permuted_upper(i64 n, const i64 *Ap, const i64 *Ai, const std::vector<i64> &iperm, UpperCSC &C) {
    bucket_by_column(n, n, [&](i64 j, auto sink) {
        const i64 b = iperm[j];
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i > j) continue;
            const i64 a = iperm[i];
            sink(std::max(a, b), std::min(a, b));
        }
    }, C.ptr, C.idx);
}

// lower-triangular column lists of C: for column j the rows i > j with A_ij != 0 (transpose of the strict upper part)
void lower_lists(i64 n, const UpperCSC &C, std::vector<i64> &lptr, std::vector<i64> &lidx) {
    bucket_by_column(n, n, [&](i64 k, auto sink) {
        for (i64 p = C.ptr[k]; p < C.ptr[k + 1]; p++)
            if (C.idx[p] < k) sink(C.idx[p], k);
    }, lptr, lidx);
}

void etree(i64 n, const UpperCSC &C, std::vector<i64> &parent) {
    parent.assign(n, -1);
    std::vector<i64> anc(n, -1);
    for (i64 k = 0; k < n; k++)
        for (i64 p = C.ptr[k]; p < C.ptr[k + 1]; p++) {
            i64 i = C.idx[p];
            while (i != -1 && i < k) {
                i64 nx = anc[i];
                anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
        }
}

// Postorder with children visited in the order given by `key` ascending (ties by index).
void postorder(i64 n, const std::vector<i64> &parent, const std::vector<i64> *key, std::vector<i64> &post) {
    std::vector<i64> order(n);
    std::iota(order.begin(), order.end(), 0);
    if (key) std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return (*key)[a] < (*key)[b]; });
    // build child lists so that popping from head yields ascending key: insert in descending order
    std::vector<i64> head(n, -1), next(n, -1);
    std::vector<i64> roots;
    for (i64 t = n - 1; t >= 0; t--) {
        i64 v = order[t];
        if (parent[v] == -1) roots.push_back(v);
        else { next[v] = head[parent[v]]; head[parent[v]] = v; }
    }
    std::reverse(roots.begin(), roots.end());
    post.clear();
    post.reserve(n);
    std::vector<i64> stack;
    for (i64 r : roots) {
        stack.push_back(r);
        while (!stack.empty()) {
            i64 v = stack.back();
            i64 c = head[v];
            if (c == -1) { post.push_back(v); stack.pop_back(); }
            else { head[v] = next[c]; stack.push_back(c); }
        }
    }
}

// Gilbert-Ng-Peyton column counts; matrix must already be labelled in postorder (parent[j] > j).
void column_counts(i64 n, const UpperCSC &C, const std::vector<i64> &parent, std::vector<i64> &cc) {
    // lower-triangular column lists: for column j the rows i > j with A_ij != 0  == transpose of C's strict upper
    std::vector<i64> lptr, lidx;
    lower_lists(n, C, lptr, lidx);
    std::vector<i64> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n), delta(n);
    for (i64 k = 0; k < n; k++) {
        i64 j = k;
        delta[j] = (first[j] == -1) ? 1 : 0;
        for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
    }
    std::iota(anc.begin(), anc.end(), 0);
    for (i64 j = 0; j < n; j++) {
        if (parent[j] != -1) delta[parent[j]]--;
        for (i64 p = lptr[j]; p < lptr[j + 1]; p++) {
            i64 i = lidx[p];
            if (first[j] <= maxfirst[i]) continue;  // j is not a leaf of the i-th row subtree
            maxfirst[i] = first[j];
            i64 jprev = prevleaf[i];
            prevleaf[i] = j;
            delta[j]++;
            if (jprev != -1) {
                i64 q = jprev;
                while (q != anc[q]) q = anc[q];
                for (i64 s = jprev; s != q;) { i64 sp = anc[s]; anc[s] = q; s = sp; }
                delta[q]--;
            }
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    cc = delta;
    for (i64 j = 0; j < n; j++)
        if (parent[j] != -1) cc[parent[j]] += cc[j];
}

// First-fit interval packing. Items are (birth, death, size); processed in phase order.
struct PoolAlloc {
    std::map<i64, i64> free_;  // offset -> size
    i64 top = 0;
    i64 alloc(i64 size) {
        if (size == 0) return 0;
        for (auto it = free_.begin(); it != free_.end(); ++it)
            if (it->second >= size) {
                i64 off = it->first, rem = it->second - size;
                free_.erase(it);
                if (rem > 0) free_[off + size] = rem;
                return off;
            }
        // extend: if the last free block touches the top, grow it
        if (!free_.empty()) {
            auto last = std::prev(free_.end());
            if (last->first + last->second == top) {
                i64 off = last->first;
                free_.erase(last);
                top = off + size;
                return off;
            }
        }
        i64 off = top;
        top += size;
        return off;
    }
    void release(i64 off, i64 size) {
        if (size == 0) return;
        auto it = free_.emplace(off, size).first;
        auto nx = std::next(it);
        if (nx != free_.end() && it->first + it->second == nx->first) { it->second += nx->second; free_.erase(nx); }
        if (it != free_.begin()) {
            auto pv = std::prev(it);
            if (pv->first + pv->second == it->first) { pv->second += it->second; free_.erase(it); }
        }
    }
};

inline i64 round_up(i64 x, i64 m) { return (x + m - 1) / m * m; }

}  // namespace

// ------------------------------------------------------------------------------------------------
void analyze(Symbolic &S, i64 n, const i64 *Ap, const i64 *Ai, const i64 *user_perm, int ordering,
             const Options &opt) {
    auto t0 = std::chrono::steady_clock::now();
    // GMRF_B200_TRACE_ANALYSIS=1: per-phase wall clock of the analysis on stderr (diagnostics)
    const bool trace = std::getenv("GMRF_B200_TRACE_ANALYSIS") != nullptr;
    auto tphase = t0;
    auto phase = [&](const char *name) {
        if (!trace) return;
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[gmrf_b200 analysis] %-28s %9.1f ms\n", name, std::chrono::duration<double, std::milli>(t - tphase).count());
        tphase = t;
    };
    if (n < 0) throw std::runtime_error("n must be non-negative");
    if (n > 2000000000LL) throw std::runtime_error("n exceeds 32-bit row index range");
    S = Symbolic();
    S.n = n;
    S.nnzA = Ap[n];
    if (Ap[0] != 0) throw std::runtime_error("colptr[0] must equal index_base");
    for (i64 j = 0; j < n; j++)
        if (Ap[j + 1] < Ap[j]) throw std::runtime_error("colptr must be non-decreasing");
    // SparseMatrixCSC invariants (the scatter map and the diagonal lookups rely on them): row indices in range and
    // strictly increasing within a column (sorted, no duplicates)
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            if (Ai[p] < 0 || Ai[p] >= n) throw std::runtime_error("row index out of range");
            if (p > Ap[j] && Ai[p] <= Ai[p - 1]) throw std::runtime_error("row indices must be strictly increasing within each column (sorted, no duplicates)");
        }

    // ---- 1. ordering ---------------------------------------------------------------------------
    std::vector<i64> perm0(n);
    if (user_perm) {
        std::vector<char> seen(n, 0);
        for (i64 k = 0; k < n; k++) {
            i64 v = user_perm[k];
            if (v < 0 || v >= n || seen[v]) throw std::runtime_error("perm is not a permutation");
            seen[v] = 1;
            perm0[k] = v;
        }
    } else if (ordering == 0) {
        std::iota(perm0.begin(), perm0.end(), 0);
    } else {
        // adjacency from the upper triangle, symmetrised, no self loops
        std::vector<i64> xadj(n + 1, 0), adj;
        for (i64 j = 0; j < n; j++)
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++)
                if (Ai[p] < j) { xadj[Ai[p] + 1]++; xadj[j + 1]++; }
        for (i64 j = 0; j < n; j++) xadj[j + 1] += xadj[j];
        adj.resize(xadj[n]);
        std::vector<i64> w(xadj.begin(), xadj.end() - 1);
        for (i64 j = 0; j < n; j++)
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++)
                if (Ai[p] < j) { adj[w[Ai[p]]++] = j; adj[w[j]++] = Ai[p]; }
        // duplicates are possible if the input repeats entries; METIS tolerates none -> dedupe
        for (i64 v = 0; v < n; v++) std::sort(adj.begin() + xadj[v], adj.begin() + xadj[v + 1]);
        {
            std::vector<i64> xa2(n + 1, 0), ad2;
            ad2.reserve(adj.size());
            for (i64 v = 0; v < n; v++) {
                i64 last = -1;
                for (i64 p = xadj[v]; p < xadj[v + 1]; p++)
                    if (adj[p] != last) { ad2.push_back(adj[p]); last = adj[p]; }
                xa2[v + 1] = (i64)ad2.size();
            }
            xadj.swap(xa2);
            adj.swap(ad2);
        }
        if (ordering == 1) order_metis_nd(n, xadj, adj, perm0);
        else if (ordering == 2) order_amd(n, xadj, adj, perm0);
        else throw std::runtime_error("unknown ordering");
    }

    phase("ordering");
    // ---- 2. etree, postorder (children by ascending column count), exact column counts ------------
    std::vector<i64> iperm0(n);
    for (i64 k = 0; k < n; k++) iperm0[perm0[k]] = k;
    UpperCSC C;
    build_permuted_upper(n, Ap, Ai, iperm0, C);
    phase("  permuted upper triangle");
    std::vector<i64> parent0, post, cc0;
    etree(n, C, parent0);
    phase("  etree");
    // first postorder -> relabel -> counts; then second postorder with the heaviest child last. A postorder is an
    // equivalent reordering: an upper-triangular entry (a, b), a < b, has b an ancestor of a, so it stays upper
    // triangular under the relabelling, the tree is the relabelled tree and every node keeps its column count --
    // nothing is recomputed from A, the permuted matrix / parent / counts are relabelled in place.
    std::vector<i64> ipo(n), tmp(n);
    auto relabel = [&](const std::vector<i64> &po, std::vector<i64> *cc) {
        // po[k] = old label of new label k
        for (i64 k = 0; k < n; k++) ipo[po[k]] = k;
        for (i64 k = 0; k < n; k++) tmp[k] = perm0[po[k]];
        perm0.swap(tmp);
        for (i64 k = 0; k < n; k++) iperm0[perm0[k]] = k;
        for (i64 k = 0; k < n; k++) tmp[k] = parent0[po[k]] == -1 ? -1 : ipo[parent0[po[k]]];
        parent0.swap(tmp);
        if (cc) {
            for (i64 k = 0; k < n; k++) tmp[k] = (*cc)[po[k]];
            cc->swap(tmp);
        }
        UpperCSC D;
        D.ptr.assign(n + 1, 0);
        for (i64 k = 0; k < n; k++) D.ptr[k + 1] = D.ptr[k] + (C.ptr[po[k] + 1] - C.ptr[po[k]]);
        D.idx.resize(C.idx.size());
#pragma omp parallel for schedule(static)
        for (i64 k = 0; k < n; k++) {
            i64 w = D.ptr[k];
            for (i64 p = C.ptr[po[k]]; p < C.ptr[po[k] + 1]; p++) D.idx[w++] = ipo[C.idx[p]];
        }
        C.ptr.swap(D.ptr);
        C.idx.swap(D.idx);
    };
    postorder(n, parent0, nullptr, post);
    relabel(post, nullptr);
    phase("  postorder + relabel");
    column_counts(n, C, parent0, cc0);
    phase("  column counts");
    postorder(n, parent0, &cc0, post);
    {
        bool ident = true;
        for (i64 k = 0; k < n; k++) if (post[k] != k) { ident = false; break; }
        if (!ident) relabel(post, &cc0);
    }
    S.perm = perm0;
    S.iperm = iperm0;
    S.parent = parent0;
    S.colcount = cc0;
    S.nnzL = 0;
    S.flops = 0;
    for (i64 j = 0; j < n; j++) {
        if (parent0[j] != -1 && parent0[j] <= j) throw std::runtime_error("internal: etree not postordered");
        S.nnzL += cc0[j];
        S.flops += (double)cc0[j] * (double)cc0[j];
    }

    phase("etree/postorder/colcounts");
    // ---- 3. supernodes: maximal chains, then relaxed amalgamation ------------------------------------
    struct Grp { i64 first, ns, nrow; double exact; i64 last_orig; };
    std::vector<i64> fund_first;  // first column of each fundamental supernode
    for (i64 j = 0; j < n; j++) {
        bool join = j > 0 && parent0[j - 1] == j && cc0[j] == cc0[j - 1] - 1;
        if (!join) fund_first.push_back(j);
    }
    i64 nf = (i64)fund_first.size();
    fund_first.push_back(n);
    std::vector<i64> fcol2s(n);
    for (i64 s = 0; s < nf; s++)
        for (i64 j = fund_first[s]; j < fund_first[s + 1]; j++) fcol2s[j] = s;
    std::vector<Grp> stack;
    auto zfrac_ok = [&](double ns, double nrow, double exact) {
        double total = ns * nrow - ns * (ns - 1) / 2.0;
        double z = (total - exact) / total;
        if (ns <= opt.relax_n[0]) return true;
        if (ns <= opt.relax_n[1]) return z < opt.relax_z[1];
        if (ns <= opt.relax_n[2]) return z < opt.relax_z[2];
        return z < opt.relax_z[3];
    };
    for (i64 s = 0; s < nf; s++) {
        Grp g;
        g.first = fund_first[s];
        g.ns = fund_first[s + 1] - fund_first[s];
        g.nrow = cc0[g.first];
        g.exact = 0;
        for (i64 j = g.first; j < g.first + g.ns; j++) g.exact += (double)cc0[j];
        g.last_orig = s;
        stack.push_back(g);
        while (stack.size() >= 2) {
            Grp &top = stack.back();
            Grp &ch = stack[stack.size() - 2];
            // parent (fundamental) supernode of the child group's root
            i64 lastcol = ch.first + ch.ns - 1;
            i64 pj = parent0[lastcol];
            // mergeable only if the child group's root hangs off a column of `top` (it then is the group
            // immediately preceding `top` in postorder). rows(merged) = cols(ch) U rows(top) exactly.
            if (pj == -1 || pj < top.first || pj >= top.first + top.ns) break;
            double ns = (double)(ch.ns + top.ns);
            double nrow = (double)ch.ns + (double)top.nrow;
            double exact = ch.exact + top.exact;
            if (!zfrac_ok(ns, nrow, exact)) break;
            Grp m;
            m.first = ch.first;
            m.ns = ch.ns + top.ns;
            m.nrow = ch.ns + top.nrow;
            m.exact = exact;
            m.last_orig = top.last_orig;
            stack.pop_back();
            stack.pop_back();
            stack.push_back(m);
        }
    }
    S.nsuper = (i64)stack.size();
    S.sfirst.resize(S.nsuper + 1);
    for (i64 s = 0; s < S.nsuper; s++) S.sfirst[s] = stack[s].first;
    S.sfirst[S.nsuper] = n;
    S.col2super.resize(n);
    for (i64 s = 0; s < S.nsuper; s++)
        for (i64 j = S.sfirst[s]; j < S.sfirst[s + 1]; j++) S.col2super[j] = s;
    S.sparent.assign(S.nsuper, -1);
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 pj = parent0[S.sfirst[s + 1] - 1];
        S.sparent[s] = pj == -1 ? -1 : S.col2super[pj];
    }
    // children lists
    S.child_ptr.assign(S.nsuper + 1, 0);
    for (i64 s = 0; s < S.nsuper; s++)
        if (S.sparent[s] != -1) S.child_ptr[S.sparent[s] + 1]++;
    for (i64 s = 0; s < S.nsuper; s++) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child_idx.resize(S.child_ptr[S.nsuper]);
    {
        std::vector<i64> w(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (i64 s = 0; s < S.nsuper; s++)
            if (S.sparent[s] != -1) S.child_idx[w[S.sparent[s]]++] = s;
    }

    phase("supernodes");
    // ---- 4. row structures ----------------------------------------------------------------------------
    // lower-triangular column lists of the permuted matrix
    std::vector<i64> lptr, lidx;
    lower_lists(n, C, lptr, lidx);
    phase("  lower-triangular column lists");
    // the sizes are known from the supernode partition (a merged group has exactly cols(child) + rows(parent) rows), so the
    // structures are written straight into their final place; a child's list is read from there by its parent
    S.rowptr.assign(S.nsuper + 1, 0);
    for (i64 s = 0; s < S.nsuper; s++) S.rowptr[s + 1] = S.rowptr[s] + stack[s].nrow;
    S.rowidx.resize(S.rowptr[S.nsuper]);
    {
        // a supernode needs only its children's finished lists: the supernodes of one assembly-tree level are independent
        // and are built concurrently, each thread with its own marker array (stamp = supernode index)
        std::vector<i32> lev(S.nsuper, 0);
        i32 nlev = 0;
        for (i64 s = 0; s < S.nsuper; s++) {
            const i64 p = S.sparent[s];
            if (p != -1) lev[p] = std::max(lev[p], lev[s] + 1);
            nlev = std::max(nlev, lev[s] + 1);
        }
        std::vector<i64> lptr_(nlev + 1, 0), lidx_(S.nsuper);
        for (i64 s = 0; s < S.nsuper; s++) lptr_[lev[s] + 1]++;
        for (i32 v = 0; v < nlev; v++) lptr_[v + 1] += lptr_[v];
        {
            std::vector<i64> w(lptr_.begin(), lptr_.end() - 1);
            for (i64 s = 0; s < S.nsuper; s++) lidx_[w[lev[s]]++] = s;
        }
        int nthreads = 1;
#ifdef _OPENMP
        nthreads = omp_get_max_threads();
#endif
        std::vector<std::vector<i64>> marks((size_t)nthreads);
        int bad = 0;
        for (i32 v = 0; v < nlev; v++) {
            const i64 a = lptr_[v], b = lptr_[v + 1];
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad) if (b - a >= 64)
            for (i64 t = a; t < b; t++) {
                int tid = 0;
#ifdef _OPENMP
                tid = omp_get_thread_num();
#endif
                std::vector<i64> &mark = marks[(size_t)tid];
                if (mark.empty()) mark.assign((size_t)n, -1);
                const i64 s = lidx_[t];
                const i64 f = S.sfirst[s], l = S.sfirst[s + 1], cap = S.rowptr[s + 1] - S.rowptr[s];
                i32 *r = S.rowidx.data() + S.rowptr[s];
                i64 cnt = 0;
                auto push = [&](i64 i) {
                    if (mark[i] == s) return;
                    if (cnt >= cap) { bad |= 1; return; }
                    mark[i] = s;
                    r[cnt++] = (i32)i;
                };
                for (i64 j = f; j < l; j++) push(j);
                for (i64 j = f; j < l; j++)
                    for (i64 p = lptr[j]; p < lptr[j + 1]; p++) push(lidx[p]);
                for (i64 cp = S.child_ptr[s]; cp < S.child_ptr[s + 1]; cp++) {
                    const i64 c = S.child_idx[cp];
                    const i64 cns = S.sfirst[c + 1] - S.sfirst[c];
                    const i32 *cr = S.rowidx.data() + S.rowptr[c];
                    const i64 cn = S.rowptr[c + 1] - S.rowptr[c];
                    for (i64 u = cns; u < cn; u++) push(cr[u]);
                }
                if (cnt != cap) bad |= 2;
                std::sort(r + (l - f), r + cnt);
            }
            if (bad) break;
        }
        if (bad & 1) throw std::runtime_error("internal: supernode row structure larger than its column count");
        if (bad & 2) throw std::runtime_error("internal: supernode row structure smaller than its column count");
    }
    phase("row structures");
    // relative indices
    S.relidx.assign(S.rowidx.size(), -1);
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 p = S.sparent[s];
        if (p == -1) {
            if (S.nr(s) != 0) throw std::runtime_error("internal: root supernode with rows below");
            continue;
        }
        i64 a = S.rowptr[s] + S.ns(s), ae = S.rowptr[s + 1];
        i64 b = S.rowptr[p], be = S.rowptr[p + 1];
        for (; a < ae; a++) {
            while (b < be && S.rowidx[b] < S.rowidx[a]) b++;
            if (b == be || S.rowidx[b] != S.rowidx[a]) throw std::runtime_error("internal: child row missing in parent");
            S.relidx[a] = (i32)(b - S.rowptr[p]);
        }
    }

    phase("relative indices");
    // ---- 5. panel layout -----------------------------------------------------------------------------
    S.panel_off.assign(S.nsuper + 1, 0);
    S.panel_ld.resize(S.nsuper);
    S.flops_stored = 0;
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 nrow = S.nrow(s), ns = S.ns(s);
        i64 ld = round_up(nrow, 2);
        S.panel_ld[s] = (i32)ld;
        S.panel_off[s + 1] = S.panel_off[s] + round_up(ld * ns, 2);
        S.max_front = std::max(S.max_front, nrow);
        S.max_ns = std::max(S.max_ns, ns);
        for (i64 j = 0; j < ns; j++) S.flops_stored += (double)(nrow - j) * (double)(nrow - j);
    }
    S.panel_total = S.panel_off[S.nsuper];
    S.diag_pos.resize(n);
    for (i64 s = 0; s < S.nsuper; s++)
        for (i64 j = S.sfirst[s]; j < S.sfirst[s + 1]; j++) {
            i64 lc = j - S.sfirst[s];
            S.diag_pos[j] = S.panel_off[s] + lc * S.panel_ld[s] + lc;
        }

    // ---- 6. levels ------------------------------------------------------------------------------------
    S.level.assign(S.nsuper, 0);
    for (i64 s = 0; s < S.nsuper; s++) {
        i64 p = S.sparent[s];
        if (p != -1) S.level[p] = std::max(S.level[p], S.level[s] + 1);
    }
    S.nlevels = 0;
    for (i64 s = 0; s < S.nsuper; s++) S.nlevels = std::max<i64>(S.nlevels, S.level[s] + 1);
    if (opt.level_alap) {
        // as late as possible: a supernode runs one level below its parent (the roots at the height of the forest), so
        // siblings with subtrees of different height share a launch instead of each paying its own chain of block steps,
        // and every update matrix lives for exactly one level (smaller pool). Children precede parents in the numbering.
        for (i64 s = S.nsuper - 1; s >= 0; s--) {
            i64 p = S.sparent[s];
            S.level[s] = (i32)(p == -1 ? S.nlevels - 1 : S.level[p] - 1);
        }
    }
    S.level_ptr.assign(S.nlevels + 1, 0);
    for (i64 s = 0; s < S.nsuper; s++) S.level_ptr[S.level[s] + 1]++;
    for (i64 l = 0; l < S.nlevels; l++) S.level_ptr[l + 1] += S.level_ptr[l];
    S.level_idx.resize(S.nsuper);
    {
        std::vector<i64> w(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (i64 s = 0; s < S.nsuper; s++) S.level_idx[w[S.level[s]]++] = s;
    }

    phase("layout/levels");
    // ---- 7. pools -------------------------------------------------------------------------------------
    S.upd_off.assign(S.nsuper, 0);
    S.upd_ld.assign(S.nsuper, 0);
    S.uvec_off.assign(S.nsuper, 0);
    S.zw_off.assign(S.nsuper, 0);
    {
        PoolAlloc pa, pv;
        for (i64 l = 0; l < S.nlevels; l++) {
            for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
                i64 s = S.level_idx[t];
                i64 nr = S.nr(s);
                i64 ld = round_up(nr, 2);
                S.upd_ld[s] = (i32)ld;
                S.upd_off[s] = pa.alloc(ld * nr);
                S.uvec_off[s] = pv.alloc(nr);
            }
            for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
                i64 s = S.level_idx[t];
                for (i64 cp = S.child_ptr[s]; cp < S.child_ptr[s + 1]; cp++) {
                    i64 c = S.child_idx[cp];
                    pa.release(S.upd_off[c], (i64)S.upd_ld[c] * S.nr(c));
                    pv.release(S.uvec_off[c], S.nr(c));
                }
            }
        }
        S.upd_total = pa.top;
        S.uvec_total = pv.top;
        // selected inversion, top-down: W_s is born at level(s) and dies after its lowest child level
        PoolAlloc pz;
        std::vector<i64> minchild(S.nsuper);
        for (i64 s = 0; s < S.nsuper; s++) {
            i64 m = S.level[s];
            for (i64 cp = S.child_ptr[s]; cp < S.child_ptr[s + 1]; cp++) m = std::min<i64>(m, S.level[S.child_idx[cp]]);
            minchild[s] = m;
        }
        std::vector<std::vector<i64>> dies_at(S.nlevels);
        for (i64 s = 0; s < S.nsuper; s++) dies_at[minchild[s]].push_back(s);
        for (i64 l = S.nlevels - 1; l >= 0; l--) {
            for (i64 t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
                i64 s = S.level_idx[t];
                S.zw_off[s] = pz.alloc((i64)S.upd_ld[s] * S.nr(s));
            }
            for (i64 s : dies_at[l]) pz.release(S.zw_off[s], (i64)S.upd_ld[s] * S.nr(s));
        }
        S.zw_total = pz.top;
    }

    phase("pools");
    // ---- 8. scatter map Q.nzval -> panels ------------------------------------------------------------
    {
        // upper-triangle entries in storage order; columns are independent (one binary search per entry): OpenMP
        std::vector<i64> start(n + 1, 0);
        for (i64 j = 0; j < n; j++) {
            i64 c = 0;
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++)
                if (Ai[p] <= j) c++;
            start[j + 1] = start[j] + c;
        }
        const i64 cnt = start[n];
        S.q_src.resize(cnt);
        S.q_dst.resize(cnt);
        i64 bad = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : bad)
        for (i64 j = 0; j < n; j++) {
            i64 k = start[j];
            const i64 b = S.iperm[j];
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
                i64 i = Ai[p];
                if (i > j) continue;
                i64 a = S.iperm[i];
                i64 col = std::min(a, b), row = std::max(a, b);
                i64 s = S.col2super[col];
                const i32 *rb = S.rowidx.data() + S.rowptr[s];
                const i32 *re = S.rowidx.data() + S.rowptr[s + 1];
                const i32 *it = std::lower_bound(rb, re, (i32)row);
                if (it == re || *it != (i32)row) { bad++; it = rb; }
                S.q_src[k] = p;
                S.q_dst[k] = S.panel_off[s] + (col - S.sfirst[s]) * (i64)S.panel_ld[s] + (it - rb);
                k++;
            }
        }
        if (bad) throw std::runtime_error("internal: Q entry outside factor pattern");
    }
    phase("scatter map");
    auto t1 = std::chrono::steady_clock::now();
    S.analysis_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// ------------------------------------------------------------------------------------------------
// caller pattern -> panel offsets (selinv_extract / selinv_dot): one binary search in the owning supernode's row list
// per entry, ~40 ns each single-threaded (cache-missing), so the columns are spread over the host cores
// ------------------------------------------------------------------------------------------------
long long entry_position(const Symbolic &S, i64 i, i64 j) {
    const i64 a = S.iperm[i], b = S.iperm[j];
    const i64 col = std::min(a, b), row = std::max(a, b);
    const i64 s = S.col2super[col];
    const i32 *rb = S.rowidx.data() + S.rowptr[s], *re = S.rowidx.data() + S.rowptr[s + 1];
    const i32 *it = std::lower_bound(rb, re, (i32)row);
    return (it != re && *it == (i32)row) ? (long long)(S.panel_off[s] + (col - S.sfirst[s]) * (i64)S.panel_ld[s] + (it - rb)) : -1LL;
}

i64 pattern_positions(const Symbolic &S, const i64 *colptr, const i64 *rowval, i64 index_base, long long *pos) {
    const i64 n = S.n;
    i64 bad = -1;
#pragma omp parallel for schedule(dynamic, 1024)
    for (i64 j = 0; j < n; j++) {
        const i64 b = S.iperm[j];
        for (i64 p = colptr[j] - index_base; p < colptr[j + 1] - index_base; p++) {
            const i64 i = rowval[p] - index_base;
            if (i < 0 || i >= n) {
                pos[p] = -1;
#pragma omp critical(gmrf_pattern_positions_bad)
                if (bad < 0 || p < bad) bad = p;
                continue;
            }
            const i64 a = S.iperm[i];
            const i64 col = std::min(a, b), row = std::max(a, b);
            const i64 s = S.col2super[col];
            const i32 *rb = S.rowidx.data() + S.rowptr[s], *re = S.rowidx.data() + S.rowptr[s + 1];
            const i32 *it = std::lower_bound(rb, re, (i32)row);
            pos[p] = (it != re && *it == (i32)row) ? (long long)(S.panel_off[s] + (col - S.sfirst[s]) * (i64)S.panel_ld[s] + (it - rb)) : -1LL;
        }
    }
    return bad;
}

// ------------------------------------------------------------------------------------------------
// serialization of the analysis
// ------------------------------------------------------------------------------------------------
namespace {

// ONE enumeration of the members of Symbolic, shared by serialize / deserialize / equal (keep in step with symbolic.hpp)
template <class V>
void visit_members(Symbolic &S, V &v) {
    v.scalar(S.n); v.scalar(S.nnzA);
    v.vec(S.perm); v.vec(S.iperm); v.vec(S.parent); v.vec(S.colcount);
    v.scalar(S.nnzL); v.scalar(S.flops);
    v.scalar(S.nsuper);
    v.vec(S.sfirst); v.vec(S.sparent); v.vec(S.col2super); v.vec(S.rowptr); v.vec(S.rowidx); v.vec(S.relidx);
    v.vec(S.panel_off); v.vec(S.panel_ld);
    v.scalar(S.panel_total); v.scalar(S.flops_stored);
    v.vec(S.child_ptr); v.vec(S.child_idx); v.vec(S.level);
    v.scalar(S.nlevels);
    v.vec(S.level_ptr); v.vec(S.level_idx);
    v.vec(S.upd_off); v.vec(S.upd_ld);
    v.scalar(S.upd_total);
    v.vec(S.zw_off);
    v.scalar(S.zw_total);
    v.vec(S.uvec_off);
    v.scalar(S.uvec_total);
    v.vec(S.q_src); v.vec(S.q_dst); v.vec(S.diag_pos);
    v.scalar(S.max_front); v.scalar(S.max_ns);
}
// symbolic.hpp: 13 stored scalars + analysis_ms (not stored) + 24 vectors; a new member changes this size
static_assert(sizeof(Symbolic) == 13 * 8 + 8 /*analysis_ms*/ + 24 * sizeof(std::vector<i64>), "Symbolic changed: update visit_members");

struct Writer {
    std::vector<char> &out;
    template <class T> void raw(const T *p, size_t cnt) { const char *c = reinterpret_cast<const char *>(p); out.insert(out.end(), c, c + cnt * sizeof(T)); }
    template <class T> void scalar(T &x) { raw(&x, 1); }
    template <class T> void vec(std::vector<T> &x) { i64 cnt = (i64)x.size(), w = (i64)sizeof(T); raw(&cnt, 1); raw(&w, 1); raw(x.data(), x.size()); }
};
struct Reader {
    const char *p, *end;
    template <class T> void raw(T *dst, size_t cnt) {
        if ((size_t)(end - p) < cnt * sizeof(T)) throw std::runtime_error("analysis blob is truncated");
        std::memcpy(dst, p, cnt * sizeof(T));
        p += cnt * sizeof(T);
    }
    template <class T> void scalar(T &x) { raw(&x, 1); }
    template <class T> void vec(std::vector<T> &x) {
        i64 cnt = 0, w = 0;
        raw(&cnt, 1); raw(&w, 1);
        if (cnt < 0 || w != (i64)sizeof(T) || (size_t)(end - p) < (size_t)cnt * sizeof(T)) throw std::runtime_error("analysis blob is malformed");
        x.resize((size_t)cnt);
        raw(x.data(), (size_t)cnt);
    }
};
constexpr char BLOB_MAGIC[8] = {'G', 'M', 'R', 'F', 'S', 'Y', 'M', '1'};

}  // namespace

unsigned long long pattern_hash(i64 n, const i64 *colptr, const i64 *rowval) {
    unsigned long long h = 1469598103934665603ULL;          // FNV-1a, one 64-bit word per step
    auto mix = [&](unsigned long long v) { h = (h ^ v) * 1099511628211ULL; h ^= h >> 29; };
    mix((unsigned long long)n);
    for (i64 j = 0; j <= n; j++) mix((unsigned long long)colptr[j]);
    for (i64 p = 0; p < colptr[n]; p++) mix((unsigned long long)rowval[p]);
    return h;
}

void serialize(const Symbolic &S, unsigned long long ph, std::vector<char> &out) {
    out.clear();
    Writer w{out};
    w.raw(BLOB_MAGIC, 8);
    w.scalar(ph);
    visit_members(const_cast<Symbolic &>(S), w);
}

void deserialize(Symbolic &S, const char *data, size_t len, i64 n, i64 nnz, unsigned long long pattern_hash_) {
    auto t0 = std::chrono::steady_clock::now();
    if (!data || len < 16 || std::memcmp(data, BLOB_MAGIC, 8) != 0) throw std::runtime_error("not an analysis blob of this library version");
    Reader r{data + 8, data + len};
    unsigned long long ph = 0;
    r.scalar(ph);
    S = Symbolic();
    visit_members(S, r);
    if (r.p != r.end) throw std::runtime_error("analysis blob has trailing bytes");
    if (S.n != n || S.nnzA != nnz) throw std::runtime_error("analysis blob belongs to a different matrix (size / nnz)");
    if (ph != pattern_hash_) throw std::runtime_error("analysis blob belongs to a different sparsity pattern");
    // cheap structural sanity of what the device code indexes with
    if ((i64)S.perm.size() != n || (i64)S.iperm.size() != n || (i64)S.col2super.size() != n || (i64)S.diag_pos.size() != n ||
        (i64)S.sfirst.size() != S.nsuper + 1 || (i64)S.rowptr.size() != S.nsuper + 1 || (i64)S.panel_off.size() != S.nsuper + 1 ||
        (i64)S.level_ptr.size() != S.nlevels + 1 || S.rowidx.size() != S.relidx.size() || S.q_src.size() != S.q_dst.size())
        throw std::runtime_error("analysis blob is inconsistent");
    S.analysis_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

bool equal(const Symbolic &a, const Symbolic &b) {
    std::vector<char> x, y;
    Writer wx{x}, wy{y};
    visit_members(const_cast<Symbolic &>(a), wx);
    visit_members(const_cast<Symbolic &>(b), wy);
    return x == y;
}

}  // namespace gmrf
