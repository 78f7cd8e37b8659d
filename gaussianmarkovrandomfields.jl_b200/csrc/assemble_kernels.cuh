// assemble_kernels.cuh -- extend-add as a GATHER through TMA-staged shared memory (north star: "extend-add scatter kernels
// that read and write HBM through TMA/shared-memory staging").
//
// The first extend-add (assemble_kernel, kernels.cuh) was a scatter: zero-fill of the parent's update matrix, then per
// child a read-modify-write of the parent entries -- 8 B written + 24 B per child entry moved for 16 B of algorithmic
// traffic, the relative-index list re-read for every child column, and dependent load -> add -> store chains (ncu,
// profiles/r01_ncu_full_assemble_kernel_3d48_v2.txt: 2.0 TB/s, DRAM 45 %). Here every parent entry is written ONCE:
//   * a CTA owns a tile of 32 columns x 256 rows of one parent front and keeps its 8192 sums in registers;
//   * per child, the rows / columns of the child that land in the tile form contiguous runs of its (sorted) relative
//     index list; a host-built table (relpos: position of every 256-row boundary of the parent in the child's list)
//     gives the runs without any search;
//   * the child's column segments of that run are contiguous in HBM: one elected warp moves them into shared memory
//     with 1-D bulk-tensor copies (cp.async.bulk ... mbarrier::complete_tx, the TMA engine; SASS: UBLKCP) -- every
//     HBM read is a full-width contiguous burst and costs no registers;
//   * the irregular part (parent row -> child row) happens in shared memory through small inverse maps;
//   * panel columns are read-modify-written once (the Q values are already there), update-matrix columns are plain
//     stores (this replaces the zero-fill).
// One owner per entry, children summed in fixed order: bit-reproducible, no atomics.
#pragma once
#include "kernels.cuh"

namespace gmrf {

constexpr int AG_CW = 32;          // tile columns
constexpr int AG_RH = 256;         // tile rows
constexpr int AG_LDS = AG_RH + 2;  // staged column stride (even: every column start stays 16-byte aligned)
constexpr int AG_MAXCH = 4;        // children whose maps are resident at once

struct AsmTile {
    int super;      // parent supernode
    int col0;       // first front column of the tile (multiple of 32)
    int row0;       // first front row of the tile (multiple of 256)
    int pad_;       // 1: panel columns only
};

struct AgChild {
    const double *U;    // child's update matrix (lane 0)
    const int *rel;     // its relative-index list (rows below its own columns)
    int uld;
    int a, b;           // run [a, b) of the child's row list that lands in the tile's rows
    int a_base;         // a rounded down to even (16-byte aligned staging)
    int pa, pb;         // run that lands in the 256-row block holding the tile's columns
    int ca, cb;         // run that lands in the tile's columns proper (= the child's update columns to fetch): first as
                        // counts of entries below c0 / below c0 + 32, then turned into positions
};

constexpr int AG_SMEM_BYTES = AG_CW * AG_LDS * 8 + AG_MAXCH * (AG_RH + AG_CW) * 4 + AG_MAXCH * (int)sizeof(AgChild) + 16;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(256, 3)
assemble_gather_kernel(const AsmTile *__restrict__ tiles, const SuperMeta *__restrict__ meta, const int *__restrict__ child_idx,
                       const int *__restrict__ relidx, const int *__restrict__ relpos, const long long *__restrict__ relpos_off,
                       double *__restrict__ Lx0, double *__restrict__ upd0, long long bstride) {
    extern __shared__ __align__(16) unsigned char ag_smem[];
    double *stage = reinterpret_cast<double *>(ag_smem);                                     // [AG_CW][AG_LDS]
    int *invRow = reinterpret_cast<int *>(ag_smem + AG_CW * AG_LDS * 8);                     // [AG_MAXCH][AG_RH]
    int *invCol = invRow + AG_MAXCH * AG_RH;                                                 // [AG_MAXCH][AG_CW]
    AgChild *ch = reinterpret_cast<AgChild *>(invCol + AG_MAXCH * AG_CW);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(ch + AG_MAXCH);
    double *__restrict__ Lx = lane_ptr_pinned(Lx0, bstride);
    double *__restrict__ upd = lane_ptr_pinned(upd0, bstride);
    pdl_launch_dependents();
    const AsmTile it = tiles[blockIdx.x];
    const SuperMeta P = meta[it.super];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = it.col0, r0 = it.row0;
    const int qr = r0 / AG_RH, qc = c0 / AG_RH;
    const int nch = P.child_end - P.child_begin;
    if (tid == 0) mbar_init(bar, 1);
    double acc[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int u = 0; u < 8; u++) acc[j][u] = 0.0;
    unsigned phase = 0;
    for (int cb0 = 0; cb0 < nch; cb0 += AG_MAXCH) {
        const int cn = min(AG_MAXCH, nch - cb0);
        // ---- maps and runs of this group of children ----
        // one thread per child walks the dependent chain child -> meta -> relpos (all children in parallel: the first
        // version did it child after child on every thread, ~5 dependent global round trips per child and tile)
        if (tid < cn) {
            const int cs = child_idx[P.child_begin + cb0 + tid];
            const SuperMeta C = meta[cs];
            const int *rp = relpos + relpos_off[cs];
            AgChild c;
            c.U = upd + C.upd_off; c.rel = relidx + C.rowptr + C.ns; c.uld = C.uld;
            c.a = rp[qr]; c.b = rp[qr + 1]; c.a_base = c.a & ~1;
            c.pa = rp[qc]; c.pb = rp[qc + 1];
            c.ca = 0; c.cb = 0;
            ch[tid] = c;
        }
        for (int e = tid; e < cn * AG_RH; e += 256) invRow[e] = -1;
        for (int e = tid; e < cn * AG_CW; e += 256) invCol[e] = -1;
        __syncthreads();
        for (int e = tid; e < cn * AG_RH; e += 256) {           // (both runs are at most 256 entries long)
            const int k = e >> 8, t = e & 255;
            const AgChild &c = ch[k];
            if (c.a + t < c.b) invRow[k * AG_RH + c.rel[c.a + t] - r0] = c.a + t;
            if (c.pa + t < c.pb) {
                const int p = c.rel[c.pa + t];
                if (p < c0) atomicAdd(&ch[k].ca, 1);
                if (p < c0 + AG_CW) atomicAdd(&ch[k].cb, 1);
                if (p >= c0 && p < c0 + AG_CW) invCol[k * AG_CW + p - c0] = c.pa + t;
            }
        }
        __syncthreads();
        if (tid < cn) { ch[tid].ca += ch[tid].pa; ch[tid].cb += ch[tid].pa; }     // counts -> positions (the list is sorted)
        __syncthreads();
        // ---- children one after the other: bulk copies into the stage, gather from it ----
        pdl_wait();          // (the maps above come from the symbolic tables only; the update matrices are the predecessor's)
        for (int k = 0; k < cn; k++) {
            const AgChild c = ch[k];
            const bool any = c.cb > c.ca && c.b > c.a;
            if (any) {
                if (warp == 0) {
                    // lane j moves the column segment of child column ca + j: rows [max(a, jc) & ~1, roundup2(b))
                    const int jc = c.ca + lane;
                    int start = 0, bytes = 0;
                    if (jc < c.cb) {
                        start = max(c.a, jc) & ~1;
                        const int end = min((c.b + 1) & ~1, c.uld);
                        bytes = max(0, end - start) * 8;
                    }
                    unsigned tot = (unsigned)bytes;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    if (lane == 0) mbar_expect_tx(bar, tot);
                    __syncwarp();
                    if (bytes > 0) bulk_g2s(stage + lane * AG_LDS + (start - c.a_base), c.U + (long long)jc * c.uld + start, (unsigned)bytes, bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int cc = warp + 8 * j;
                    const int jc = invCol[k * AG_CW + cc];
                    if (jc < 0) continue;                              // (uniform across the warp)
                    const double *col = stage + (jc - c.ca) * AG_LDS - c.a_base;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int rr = lane + 32 * u;
                        const int ir = invRow[k * AG_RH + rr];
                        if (ir >= jc) acc[j][u] += col[ir];            // (ir >= jc <=> the entry is in the child's lower triangle)
                    }
                }
            }
            __syncthreads();                                           // the stage is free again
        }
    }
    // ---- one write per parent entry ----
    double *Lp = Lx + P.panel_off;
    double *Up = upd + P.upd_off;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = c0 + warp + 8 * j;
        if (c >= P.nrow) continue;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int r = r0 + lane + 32 * u;
            if (r >= P.nrow || r < c) continue;
            if (c < P.ns) {
                if (acc[j][u] != 0.0) Lp[r + (long long)c * P.ld] += acc[j][u];
            } else if (!it.pad_) {            // (pad_ = 1: the update part is gathered by the update-matrix product itself)
                Up[(r - P.ns) + (long long)(c - P.ns) * P.uld] = acc[j][u];
            }
        }
    }
}

}  // namespace gmrf
