// Stage-2 kernels: gapped 3Di+AA Smith-Waterman with the exact semantics of the reference's striped
// AVX2 kernel (src/commons/StructureSmithWaterman.cpp:1094-1322 sw_sse2_word, :1326-1554 sw_sse2_int).
//
// Recurrence reproduced (see DESIGN.md "SW semantics"):  with s = profAA[tAA_c][q] + prof3Di[t3Di_c][q]
//   Hp(c,q) = max(H(c-1,q-1)+s, E(c,q), Fs(c,q), 0)      Fs = F restricted to the SIMD stripe of q
//   H (c,q) = max(Hp(c,q), F(c,q))                        F  = the usual vertical gap state
//   E(c+1,q) = max(E(c,q)-ge, Hp(c,q)-go)                 (E opens from Hp, NOT from H: the lazy-F loop of the
//                                                          reference fixes H but not E, SSW.cpp:1230-1259)
//   Fs/F(c,q+1) = max(Fs/F(c,q)-ge, Hp(c,q)-go), Fs := 0 at q = l*segLen, segLen = ceil(L/16)
// score = max H; dbEnd = first column reaching the maximum; qEnd = smallest row holding it in that column.
//
// Mapping to the GPU (sw_packed_kernel):
//   * 16 threads per alignment, thread g owns stripe g (rows g*segLen .. (g+1)*segLen-1): the AVX2 lane
//     striping of the reference *is* the thread decomposition, so Fs is thread-local and F crosses threads by shuffle.
//   * anti-diagonal wavefront inside the 16-thread group: thread g works on column s-g at step s; H of the last
//     row, F and the target letters are handed down with __shfl_up_sync.
//   * every 32-bit register holds the same cell of TWO different targets (s16x2), so the DP runs on the DPX
//     instructions VIADDMNMX.S16x2(.RELU) / VIMNMX(3).S16x2 / VIADD.16x2.
//   * the per-query score profile over all 21x21 (AA,3Di) letter pairs (+1 padding letter) is built once per work
//     item in shared memory as int8, in a layout that makes every LDS.64 of the wavefront bank-conflict free;
//     one PRMT turns two int8 scores into the sign-extended s16x2 operand.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace fsb {

constexpr int SW_ALPH = 21;
constexpr int SW_COMBOS = SW_ALPH * SW_ALPH + 1;  // + padding letter
constexpr int SW_PADLETTER = SW_ALPH * SW_ALPH;
constexpr int SW_GROUP = 16;                      // threads per alignment == AVX2 int16 lanes of the reference
constexpr int SW_MAXR = 24;                       // rows per thread supported by the packed kernel (L <= 384)
constexpr int SW_BLOCK = 512;

struct SwProfileDesc {   // one per (query, direction)
    uint64_t off;        // offset into qAA/q3Di/biasAA/biasSS
    uint32_t len;
    uint32_t pad;
};

struct SwPair {          // one alignment job
    uint32_t profile;    // index into SwProfileDesc
    uint32_t target;     // target id
    uint32_t tlen;       // effective target length (<= real length; end-row mode truncates at dbEnd+1)
    uint32_t want;       // end-row mode: the score whose first row is searched in the last column
};

struct SwTask {          // 4 alignments of the same profile processed by one warp
    uint32_t pair[4];    // indices into SwPair, 0xFFFFFFFF = empty
};

struct SwItem {          // unit of CTA scheduling: tasks [taskBegin, taskEnd) share one profile
    uint32_t profile;
    uint32_t taskBegin;
    uint32_t taskEnd;
    uint32_t pad;        // 1 = end-row mode
};

struct SwOut {           // result of one alignment
    uint32_t score;      // maximum H (if >= SW_OVERFLOW_GUARD the pair must be redone by the int32 kernel)
    int32_t dbEnd;       // first column where the maximum is reached (0 if score == 0, like the reference)
    int32_t qEnd;        // end-row mode: smallest row == want in the last column, else -1
    uint32_t flags;      // bit0: possible int16 overflow (redo), bit1: int32 path taken, bit2: produced by the generic kernel (qEnd valid)
};

constexpr uint32_t SW_OVERFLOW_GUARD = 32000;

struct SwArgs {
    const uint8_t *tAA;          // target residues (codes 0..20), concatenated
    const uint8_t *t3Di;
    const uint64_t *tOff;        // per target offset
    const uint8_t *qAA;          // profile sequences (forward and reversed queries), concatenated
    const uint8_t *q3Di;
    const int8_t *biasAA;        // per position composition bias (SSW.cpp:1563-1575)
    const int8_t *biasSS;
    const SwProfileDesc *profiles;
    const SwPair *pairs;
    const SwTask *tasks;
    const SwItem *items;
    uint32_t nItems;
    uint32_t *itemCounter;       // persistent-CTA work counter (zeroed by the host before launch)
    SwOut *out;
    int gapOpen, gapExtend;
};

__constant__ int8_t c_matAA[SW_ALPH * SW_ALPH];
__constant__ int8_t c_mat3Di[SW_ALPH * SW_ALPH];

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }

// sign-extending pack of byte J of a (low half) and byte J of b (high half)
template <int J>
__device__ __forceinline__ uint32_t pack_scores(uint32_t a, uint32_t b) {
    constexpr uint32_t sel = (uint32_t) J | ((uint32_t) (J | 8) << 4) | ((uint32_t) (4 + J) << 8) | ((uint32_t) ((4 + J) | 8) << 12);
    // __byte_perm() masks the selector nibbles to 3 bits; the sign-replicate bit needs the PTX form
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// ---------------------------------------------------------------------------------------------------------
// Packed DPX kernel. R = rows per thread (even, >= segLen), L <= 16*R.
// Shared memory: combined profile [SW_COMBOS][LP] int8 with LP = 128*ceil(R/8); element (g, r) of a letter row is
// stored at (r/8)*128 + g*8 + (r%8): the 16 threads of a group read 128 consecutive bytes per LDS.64, and because
// LP is a multiple of 128 B the letter (which differs from thread to thread) does not change the bank.
// ---------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(SW_BLOCK, 1) sw_packed_kernel(SwArgs a) {
    constexpr int K = (R + 7) / 8;        // 8-byte chunks per thread
    constexpr int LP = 128 * K;           // bytes per letter row
    extern __shared__ __align__(16) uint8_t smem[];
    int8_t *prof = reinterpret_cast<int8_t *>(smem);
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_task;

    const int lane = threadIdx.x & 31;
    const int g = lane & 15;
    const int grp = lane >> 4;
    const uint32_t MGO = (uint32_t) ((-a.gapOpen) & 0xffff) * 0x10001u;
    const uint32_t MGE = (uint32_t) ((-a.gapExtend) & 0xffff) * 0x10001u;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_item = atomicAdd(a.itemCounter, 1u);
        }
        __syncthreads();
        const uint32_t itemIdx = s_item;
        if (itemIdx >= a.nItems) break;
        const SwItem item = a.items[itemIdx];
        const SwProfileDesc pd = a.profiles[item.profile];
        const int L = (int) pd.len;
        const int segLen = (L + SW_GROUP - 1) / SW_GROUP;
        const bool oddSeg = (segLen & 1) != 0;   // R == segLen + 1: the last row of every thread is padding
        const bool endRow = item.pad != 0;        // item mode: report the first row holding `want` in the last column

        // ---- build the combined int8 profile of this query in shared memory ----
        if (threadIdx.x == 0) s_task = item.taskBegin;
        {
            const uint8_t *qa = a.qAA + pd.off;
            const uint8_t *qs = a.q3Di + pd.off;
            const int8_t *ba = a.biasAA + pd.off;
            const int8_t *bs = a.biasSS + pd.off;
            // each thread fills 4 consecutive bytes (one 32-bit store) per iteration
            for (int w = threadIdx.x; w < SW_COMBOS * (LP / 4); w += SW_BLOCK) {
                const int X = w / (LP / 4);
                const int p4 = (w - X * (LP / 4)) * 4;
                const int k = p4 >> 7, gg = (p4 & 127) >> 3, j0 = p4 & 7;
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = 8 * k + j0 + j;
                    const int q = gg * segLen + r;
                    int v = -128;
                    if (X != SW_PADLETTER && r < segLen && q < L) {
                        const int aa = X / SW_ALPH, ss = X - aa * SW_ALPH;
                        v = (int) c_matAA[aa * SW_ALPH + qa[q]] + (int) ba[q] + (int) c_mat3Di[ss * SW_ALPH + qs[q]] + (int) bs[q];
                        v = max(-127, min(127, v));   // host guarantees the range; clamp keeps memory safe
                    }
                    packed |= ((uint32_t) (v & 0xff)) << (8 * j);
                }
                *reinterpret_cast<uint32_t *>(prof + (size_t) X * LP + p4) = packed;
            }
        }
        __syncthreads();

        // ---- warps pull tasks (4 alignments each) of this item ----
        for (;;) {
            uint32_t taskIdx = 0;
            if (lane == 0) taskIdx = atomicAdd(&s_task, 1u);
            taskIdx = __shfl_sync(0xffffffffu, taskIdx, 0);
            if (taskIdx >= item.taskEnd) break;
            const SwTask task = a.tasks[taskIdx];
            const uint32_t pA = task.pair[grp * 2 + 0], pB = task.pair[grp * 2 + 1];
            int LtA = 0, LtB = 0;
            uint32_t wantA = 0, wantB = 0;
            const uint8_t *aaA = nullptr, *ssA = nullptr, *aaB = nullptr, *ssB = nullptr;
            if (pA != 0xFFFFFFFFu) {
                const SwPair pr = a.pairs[pA];
                LtA = (int) pr.tlen; wantA = pr.want;
                aaA = a.tAA + a.tOff[pr.target]; ssA = a.t3Di + a.tOff[pr.target];
            }
            if (pB != 0xFFFFFFFFu) {
                const SwPair pr = a.pairs[pB];
                LtB = (int) pr.tlen; wantB = pr.want;
                aaB = a.tAA + a.tOff[pr.target]; ssB = a.t3Di + a.tOff[pr.target];
            }
            const int LtG = max(LtA, LtB);
            int LtW = max(LtG, __shfl_xor_sync(0xffffffffu, LtG, 16));
            const int nSteps = LtW + SW_GROUP - 1;

            uint32_t H[R], E[R];
#pragma unroll
            for (int r = 0; r < R; r++) { H[r] = 0; E[r] = 0; }
            uint32_t best = 0;            // packed running maximum of this thread's rows
            int colA = 0, colB = 0;       // first column where the thread maximum was reached
            int rowA = 0x7fffffff, rowB = 0x7fffffff;   // end-row mode
            uint32_t hl = 0, fo = 0;      // values handed to thread g+1: H of my last row, F leaving my last row
            uint32_t diag0 = 0;           // H(c-1, row0-1)
            uint32_t lettersBuf = 0, letters = 0;

            for (int s = 0; s < nSteps; s++) {
                // refill the letter buffer every 16 steps: thread g fetches column s+g
                if ((s & 15) == 0) {
                    const int cc = s + g;
                    uint32_t xa = SW_PADLETTER, xb = SW_PADLETTER;
                    if (cc < LtA) xa = (uint32_t) aaA[cc] * SW_ALPH + (uint32_t) ssA[cc];
                    if (cc < LtB) xb = (uint32_t) aaB[cc] * SW_ALPH + (uint32_t) ssB[cc];
                    lettersBuf = xa | (xb << 16);
                }
                const uint32_t fresh = __shfl_sync(0xffffffffu, lettersBuf, (s & 15) | (grp << 4));
                uint32_t lin = __shfl_up_sync(0xffffffffu, letters, 1, 16);
                uint32_t hin = __shfl_up_sync(0xffffffffu, hl, 1, 16);
                uint32_t fin = __shfl_up_sync(0xffffffffu, fo, 1, 16);
                if (g == 0) { lin = fresh; hin = 0; fin = 0; }
                letters = lin;
                const int c = s - g;
                if (c >= 0 && c < LtG) {
                    const int8_t *rowA_p = prof + (size_t) (letters & 0xffffu) * LP + g * 8;
                    const int8_t *rowB_p = prof + (size_t) (letters >> 16) * LP + g * 8;
                    uint2 sa[K], sb[K];
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        sa[k] = *reinterpret_cast<const uint2 *>(rowA_p + k * 128);
                        sb[k] = *reinterpret_cast<const uint2 *>(rowB_p + k * 128);
                    }
                    uint32_t hd = diag0;
                    uint32_t fs = 0;
                    uint32_t ff = fin;
                    uint32_t cm = 0;
                    uint32_t ffBeforeLast = fin;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t wa = ((r & 7) < 4) ? sa[r >> 3].x : sa[r >> 3].y;
                        const uint32_t wb = ((r & 7) < 4) ? sb[r >> 3].x : sb[r >> 3].y;
                        uint32_t sc;
                        switch (r & 3) {
                            case 0: sc = pack_scores<0>(wa, wb); break;
                            case 1: sc = pack_scores<1>(wa, wb); break;
                            case 2: sc = pack_scores<2>(wa, wb); break;
                            default: sc = pack_scores<3>(wa, wb); break;
                        }
                        if (r == R - 1) ffBeforeLast = ff;
                        const uint32_t t = __viaddmax_s16x2_relu(hd, sc, E[r]);
                        const uint32_t hp = vmax2(t, fs);
                        const uint32_t h = vmax2(hp, ff);
                        const uint32_t x = __vadd2(hp, MGO);
                        E[r] = __viaddmax_s16x2(E[r], MGE, x);
                        fs = __viaddmax_s16x2(fs, MGE, x);
                        ff = __viaddmax_s16x2(ff, MGE, x);
                        hd = H[r];
                        H[r] = h;
                        if (r & 1) cm = __vimax3_s16x2(cm, H[r - 1], h);   // one VIMNMX3 per two rows
                    }
                    // hand-off values: with an odd segLen the last register row is padding
                    hl = oddSeg ? H[R - 2] : H[R - 1];
                    fo = oddSeg ? ffBeforeLast : ff;
                    diag0 = hin;
                    const uint32_t nb = vmax2(best, cm);
                    if (nb != best) {
                        if ((nb & 0xffffu) != (best & 0xffffu)) colA = c;
                        if ((nb >> 16) != (best >> 16)) colB = c;
                        best = nb;
                    }
                    if (endRow) {
                        if (c == LtA - 1) {
#pragma unroll
                            for (int r = R - 1; r >= 0; r--) if ((H[r] & 0xffffu) == wantA && r < segLen) rowA = g * segLen + r;
                        }
                        if (c == LtB - 1) {
#pragma unroll
                            for (int r = R - 1; r >= 0; r--) if ((H[r] >> 16) == wantB && r < segLen) rowB = g * segLen + r;
                        }
                    }
                } else {
                    diag0 = hin;
                }
            }
#ifdef FSB_DEBUG
            if (grp == 0) printf("g %d best %08x colA %d H0 %08x E0 %08x hl %08x fo %08x segLen %d L %d LtA %d LtB %d nSteps %d\n", g, best, colA, H[0], E[0], hl, fo, segLen, L, LtA, LtB, nSteps);
#endif
            // ---- reduce over the 16 threads of the group ----
            int bA = (int) (best & 0xffffu), bB = (int) (best >> 16);
            int sA = bA, sB = bB;
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                sA = max(sA, __shfl_xor_sync(0xffffffffu, sA, m, 16));
                sB = max(sB, __shfl_xor_sync(0xffffffffu, sB, m, 16));
            }
            int eA = (bA == sA) ? colA : 0x7fffffff, eB = (bB == sB) ? colB : 0x7fffffff;
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                eA = min(eA, __shfl_xor_sync(0xffffffffu, eA, m, 16));
                eB = min(eB, __shfl_xor_sync(0xffffffffu, eB, m, 16));
                rowA = min(rowA, __shfl_xor_sync(0xffffffffu, rowA, m, 16));
                rowB = min(rowB, __shfl_xor_sync(0xffffffffu, rowB, m, 16));
            }
            if (g == 0) {
                if (pA != 0xFFFFFFFFu) {
                    SwOut o; o.score = (uint32_t) sA; o.dbEnd = (sA == 0) ? 0 : eA;
                    o.qEnd = (rowA == 0x7fffffff) ? -1 : rowA; o.flags = (sA >= (int) SW_OVERFLOW_GUARD) ? 1u : 0u;
                    a.out[pA] = o;
                }
                if (pB != 0xFFFFFFFFu) {
                    SwOut o; o.score = (uint32_t) sB; o.dbEnd = (sB == 0) ? 0 : eB;
                    o.qEnd = (rowB == 0x7fffffff) ? -1 : rowB; o.flags = (sB >= (int) SW_OVERFLOW_GUARD) ? 1u : 0u;
                    a.out[pB] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Generic kernel: any query length, int32 arithmetic, one warp per alignment.  Used for queries longer than the
// packed kernel supports and to redo alignments whose int16 score could have saturated: it emulates first the word
// kernel (16 lanes, saturating at 32767) and, when that returns 32767, the int kernel (8 lanes) exactly like
// alignScoreEndPos (SSW.cpp:295-336).  Thread t owns rows [t*RB, (t+1)*RB); H/E columns live in a global scratch.
// ---------------------------------------------------------------------------------------------------------
struct SwGenericArgs {
    SwArgs base;
    const uint32_t *pairList;   // indices into pairs
    uint32_t nPairs;
    int32_t *scratch;           // nWarps * 2 * scratchStride int32
    uint32_t scratchStride;     // >= 32 * ceil(maxLpad/32)
    uint32_t *pairCounter;
};

template <bool SAT, int LANES>
__device__ void sw_generic_pass(const SwArgs &a, const SwProfileDesc pd, const uint8_t *tAA, const uint8_t *t3Di, int Lt,
                                int32_t *Hs, int32_t *Es, int lane, int &outScore, int &outCol, int &outRow) {
    const int L = (int) pd.len;
    const int segLen = (L + LANES - 1) / LANES;
    const int RB = (L + 31) / 32;
    const uint8_t *qa = a.qAA + pd.off;
    const uint8_t *qs = a.q3Di + pd.off;
    const int8_t *ba = a.biasAA + pd.off;
    const int8_t *bs = a.biasSS + pd.off;
    const int go = a.gapOpen, ge = a.gapExtend;
    const int row0 = lane * RB;
    for (int r = 0; r < RB; r++) { Hs[r * 32 + lane] = 0; Es[r * 32 + lane] = 0; }
    int best = 0, bestCol = 0, bestRow = 0x7fffffff;
    int hl = 0, fo = 0, fso = 0, diag0 = 0;
    const int nSteps = Lt + 31;
    for (int s = 0; s < nSteps; s++) {
        int hin = __shfl_up_sync(0xffffffffu, hl, 1);
        int fin = __shfl_up_sync(0xffffffffu, fo, 1);
        int fsin = __shfl_up_sync(0xffffffffu, fso, 1);
        if (lane == 0) { hin = 0; fin = 0; fsin = 0; }
        const int c = s - lane;
        if (c >= 0 && c < Lt && row0 < L) {
            const int ta = tAA[c], ts = t3Di[c];
            int hd = diag0, ff = fin, fs = fsin;
            for (int r = 0; r < RB; r++) {
                const int q = row0 + r;
                if (q >= L) break;
                if (q % segLen == 0) fs = 0;
                const int sc = (int) c_matAA[ta * SW_ALPH + qa[q]] + (int) ba[q] + (int) c_mat3Di[ts * SW_ALPH + qs[q]] + (int) bs[q];
                int h = hd + sc;
                if (SAT) h = min(h, 32767);
                const int e = Es[r * 32 + lane];
                h = max(max(h, e), max(fs, 0));
                const int x = max(h - go, 0);
                const int hfull = max(h, ff);
                Es[r * 32 + lane] = max(max(e - ge, 0), x);
                fs = max(max(fs - ge, 0), x);
                ff = max(max(ff - ge, 0), max(hfull - go, 0));
                hd = Hs[r * 32 + lane];
                Hs[r * 32 + lane] = hfull;
                if (hfull > best) { best = hfull; bestCol = c; bestRow = q; }
                hl = hfull;
            }
            fo = ff; fso = fs;
        }
        diag0 = hin;
    }
    int sMax = best;
    for (int m = 16; m >= 1; m >>= 1) sMax = max(sMax, __shfl_xor_sync(0xffffffffu, sMax, m));
    int col = (best == sMax) ? bestCol : 0x7fffffff;
    for (int m = 16; m >= 1; m >>= 1) col = min(col, __shfl_xor_sync(0xffffffffu, col, m));
    int row = (best == sMax && bestCol == col) ? bestRow : 0x7fffffff;
    for (int m = 16; m >= 1; m >>= 1) row = min(row, __shfl_xor_sync(0xffffffffu, row, m));
    outScore = sMax; outCol = (sMax == 0) ? 0 : col; outRow = (sMax == 0) ? L - 1 : row;
}

__global__ void __launch_bounds__(128) sw_generic_kernel(SwGenericArgs ga) {
    const SwArgs &a = ga.base;
    const int lane = threadIdx.x & 31;
    const uint32_t warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int32_t *Hs = ga.scratch + (size_t) warpGlobal * 2 * ga.scratchStride;
    int32_t *Es = Hs + ga.scratchStride;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(ga.pairCounter, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= ga.nPairs) break;
        const uint32_t p = ga.pairList[idx];
        const SwPair pr = a.pairs[p];
        const SwProfileDesc pd = a.profiles[pr.profile];
        const uint8_t *tAA = a.tAA + a.tOff[pr.target];
        const uint8_t *t3Di = a.t3Di + a.tOff[pr.target];
        int sc, col, row;
        sw_generic_pass<true, 16>(a, pd, tAA, t3Di, (int) pr.tlen, Hs, Es, lane, sc, col, row);
        uint32_t flags = 4;
        if (sc == 32767) {
            sw_generic_pass<false, 8>(a, pd, tAA, t3Di, (int) pr.tlen, Hs, Es, lane, sc, col, row);
            flags = 6;
        }
        if (lane == 0) {
            SwOut o; o.score = (uint32_t) sc; o.dbEnd = col; o.qEnd = row; o.flags = flags;
            a.out[p] = o;
        }
    }
}

}  // namespace fsb
