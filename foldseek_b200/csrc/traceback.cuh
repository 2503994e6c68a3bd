// traceback.cuh — start positions + backtrace of accepted pairs on the GPU.
//
// Reference behaviour reproduced: StructureSmithWaterman::alignStartPosBacktraceBlock
// (src/commons/StructureSmithWaterman.cpp:369-537) which drives the block-aligner crate
// (lib/mmseqs/lib/block-aligner/src/scan_block.rs: align_core :120-640, place_block_3di :1302-1440,
// Trace::cigar_core :1844-2006; AVX2 build: 16 int16 lanes, ZERO = 1<<14, MIN = 0, STEP = 8).
// Ties between insertion and deletion depend on the orientation of the block a cell was computed in, so the block
// trajectory (grow / shift / shrink decisions, 32-bit offsets, saturating int16 deltas) is reproduced exactly.
//
// Mapping: 16 threads per pair, thread k = SIMD lane k of the reference's vectors; the scalar control state is replicated
// in every lane (all decisions derive from group-wide reductions, so the 16 lanes stay converged). Border rows/columns
// and their checkpoints live in a per-group global scratch (hot part a few hundred bytes, L1 resident); the trace is one
// byte per cell (4 bits used) in a per-group global buffer; the walk back is done by lane 0.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tb {

constexpr int TL = 16;                 // lanes per pair
constexpr int TB_THREADS = 128;        // 8 pairs in flight per CTA
constexpr int TB_MAXS = 4096;          // MAX_SIZE (StructureSmithWaterman.cpp:377)
constexpr int TB_STATE = 8 * TB_MAXS + 2 * TL;   // int16 per group: D_col C_col D_row R_row + checkpoints + 2 temp vectors
constexpr int TB_ZERO = 1 << 14;
constexpr int TB_MIN = 0;
constexpr int TB_STEP = 8;
constexpr int TB_NUL = 21;             // padding code: row/column of i8::MIN in the 27x32 AAMatrix (scores.rs:52-57)
constexpr int TB_STRIDE = 32;

struct TbPair {
    uint64_t qOff;        // forward query codes/bias offset (profile 2q)
    uint64_t tOff;        // target offset
    uint64_t btOff;       // where this pair's backtrace goes
    int32_t qEnd, dbEnd;  // inclusive end positions of the forward pass
    int32_t targetScore;  // forward score the block aligner has to reproduce
    uint32_t pad;
};

struct TbOut {
    int32_t qStart, dbStart;
    uint32_t identical, alnLen;
    int32_t ok;           // 1 = score reproduced, 0 = reference's "score mismatch" exit, -1 = trace capacity exceeded (retry)
    int32_t score;        // block-aligner score of the last attempt
};

struct TbArgs {
    const TbPair *pairs;
    const uint32_t *order;     // pair indices to process
    uint32_t nPairs;
    uint32_t *counter;
    const uint8_t *qAA, *q3Di;
    const int8_t *qBiasAA, *qBiasSS;
    const uint8_t *tAA, *t3Di;
    const int8_t *tabAA, *tab3Di;   // 32 x 32, padding row/col = -128
    int gapOpen, gapExtend;         // negative (Gaps, StructureSmithWaterman.cpp:380-381)
    int16_t *state;
    uint8_t *trace;
    uint64_t traceCapVec;           // vectors per group
    uint4 *blocks;
    uint32_t blockCap;
    char *bt;
    TbOut *out;
    // test hook (fsb200_test_block_align): one engine run with this block range and x-drop instead of the exponential search
    uint32_t testMinSize = 0, testMaxSize = 0;
    int testXDrop = 0;
};

__device__ __forceinline__ int sat16(int x) { return min(32767, max(-32768, x)); }

struct Group {
    unsigned mask;
    int k;
    __device__ __forceinline__ int up(int x, int n) const { return __shfl_up_sync(mask, x, n, TL); }
    __device__ __forceinline__ int idx(int x, int src) const { return __shfl_sync(mask, x, src, TL); }
    __device__ __forceinline__ unsigned long long idx64(unsigned long long x, int src) const { return __shfl_sync(mask, x, src, TL); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ int hmax(int x) const {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) x = max(x, __shfl_xor_sync(mask, x, o, TL));
        return x;
    }
    __device__ __forceinline__ unsigned long long hmax64(unsigned long long x) const {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) { unsigned long long y = __shfl_xor_sync(mask, x, o, TL); x = y > x ? y : x; }
        return x;
    }
    // _mm256_slli_si256 by n int16 inside each 128-bit half, zeros shifted in
    __device__ __forceinline__ int sllz(int x, int n) const { const int t = up(x, n); return (k & 7) >= n ? t : 0; }
};

struct Engine {
    Group g;
    const TbArgs *a;
    // pair
    const uint8_t *qa, *qs, *ta, *ts;     // forward arrays; block position p (1-based) maps to index end - (p - 1)
    const int8_t *qba, *qbs;
    int qEnd, dbEnd, qn, tn;
    const int8_t *tabAA, *tab3Di;         // shared memory
    // per-lane constants of the prefix scan (avx2.rs:296-310)
    int gOpen, gExt, gExtAll, gConsts, gExt2, gExt4;
    // scratch
    int16_t *Dcol, *Ccol, *Drow, *Rrow, *DcolK, *CcolK, *DrowK, *RrowK, *temp1, *temp2;
    uint8_t *trace;
    uint4 *blocks;
    unsigned long long traceIdx, ckptTraceIdx, traceCap;
    unsigned blockIdx_, ckptBlockIdx, blockCap;
    bool overflow;
    // result of one align()
    int resScore;
    unsigned resI, resJ;

    __device__ __forceinline__ void init_consts() {
        const int gap = gExt;
        int s1 = sat16(g.sllz(gap, 1) + gap);
        int s2 = sat16(g.sllz(s1, 2) + s1);
        int s4 = sat16(g.sllz(s2, 4) + s2);
        // words 0..3 of each half = word 7 of that half, words 4..7 = 0; then q0 = q1 = src q1, q2 = q3 = src q0
        const int w7lo = g.idx(s4, 7);
        const int p = g.k < 8 ? 0 : w7lo;
        gExtAll = sat16(p + s4);
        gConsts = s4;
        gExt2 = (int) (int16_t) ((uint16_t) (int16_t) gap << 1);
        gExt4 = (int) (int16_t) ((uint16_t) (int16_t) gap << 2);
    }

    __device__ __forceinline__ int prefix_scan(int R) const {   // simd_prefix_scan_i16 (avx2.rs:312-337)
        int s1 = max(R, sat16(g.sllz(R, 1) + gExt));
        int s2 = max(s1, sat16(g.sllz(s1, 2) + gExt2));
        int s4 = max(s2, sat16(g.sllz(s2, 4) + gExt4));
        int c = g.idx(s4, g.k < 8 ? (g.k & 3) : 7);
        c = sat16(c + gConsts);
        return max(s4, c);
    }

    __device__ __forceinline__ int score(bool right, unsigned vecPos, unsigned scalarPos) const {
        // vector position: this lane's padded index; scalar position: the other sequence (place_block_3di :1343-1351)
        const unsigned vp = vecPos + g.k;
        int va, vs, vb, ca, cs, sb;
        if (right) {
            const bool in = vp >= 1 && vp <= (unsigned) qn;
            const int ix = qEnd - (int) (vp - 1);
            va = in ? qa[ix] : TB_NUL; vs = in ? qs[ix] : TB_NUL; vb = in ? (int) qba[ix] + (int) qbs[ix] : 0;
            const bool in2 = scalarPos >= 1 && scalarPos <= (unsigned) tn;
            const int jx = dbEnd - (int) (scalarPos - 1);
            ca = in2 ? ta[jx] : TB_NUL; cs = in2 ? ts[jx] : TB_NUL; sb = 0;
        } else {
            const bool in = vp >= 1 && vp <= (unsigned) tn;
            const int jx = dbEnd - (int) (vp - 1);
            va = in ? ta[jx] : TB_NUL; vs = in ? ts[jx] : TB_NUL; vb = 0;
            const bool in2 = scalarPos >= 1 && scalarPos <= (unsigned) qn;
            const int ix = qEnd - (int) (scalarPos - 1);
            ca = in2 ? qa[ix] : TB_NUL; cs = in2 ? qs[ix] : TB_NUL; sb = in2 ? (int) qba[ix] + (int) qbs[ix] : 0;
        }
        const int sA = tabAA[ca * TB_STRIDE + va], sS = tab3Di[cs * TB_STRIDE + vs];
        const int pb = sat16(sb + vb);
        return sat16(sat16(sA + sS) + pb);
    }

    __device__ __forceinline__ void add_block(unsigned i, unsigned j, unsigned width, unsigned height, bool right) {
        if (blockIdx_ >= blockCap) { overflow = true; return; }
        if (g.k == 0) blocks[blockIdx_] = make_uint4(i, j, height | (width << 16), right ? 1u : 0u);
        blockIdx_++;
    }

    // place_block_3di (scan_block.rs:1302-1440), TRACE = X_DROP = true
    __device__ void place_block(unsigned startI, unsigned startJ, unsigned width, unsigned height, int16_t *Dc, int16_t *Cc,
                                int16_t *Dr, int16_t *Rr, int corner, bool right, int &dmax, int &ai, int &aj) {
        dmax = TB_MIN; ai = 0; aj = 0;
        if (width == 0 || height == 0 || overflow) return;
        const unsigned long long need = (unsigned long long) width * (height / TL);
        if (traceIdx + need > traceCap) { overflow = true; return; }
        const int gapOE = sat16(gOpen - gExt);
        for (unsigned j = 0; j < width; j++) {
            int R01top = TB_MIN, D11 = TB_MIN, R11 = TB_MIN, prevTraceRtop = 0;
            for (unsigned i = 0; i < height; i += TL) {
                const int D10 = Dc[i + g.k], C10 = Cc[i + g.k];
                const int upv = g.up(D10, 1);
                const int D00 = g.k == 0 ? corner : upv;
                corner = g.idx(D10, TL - 1);
                D11 = sat16(D00 + score(right, startI + i, startJ + j));
                if (startI + i == 0 && startJ + j == 0 && g.k == 0) D11 = TB_ZERO;
                const int C11open = sat16(D10 + gOpen);
                const int C11 = max(sat16(C10 + gExt), C11open);
                D11 = max(D11, C11);
                const int D11open = sat16(D11 + gapOE);
                R11 = prefix_scan(D11open);
                R11 = max(R11, sat16(R01top + gExtAll));
                D11 = max(D11, R11);
                R01top = g.idx(R11, TL - 1);
                const int tempTraceR = R11 == D11open;
                const int upT = g.up(tempTraceR, 1);
                const int traceR = g.k == 0 ? prevTraceRtop : upT;
                prevTraceRtop = g.idx(tempTraceR, TL - 1);
                trace[traceIdx * TL + g.k] = (uint8_t) ((D11 == C11) | ((D11 == R11) << 1) | ((C11 == C11open) << 2) | (traceR << 3));
                traceIdx++;
                if (D11 >= dmax) { dmax = D11; ai = (int) i; aj = (int) j; }
                Dc[i + g.k] = (int16_t) D11;
                Cc[i + g.k] = (int16_t) C11;
            }
            corner = TB_MIN;
            if (g.k == TL - 1) { Dr[j] = (int16_t) D11; Rr[j] = (int16_t) R11; }
        }
        g.sync();
    }

    __device__ __forceinline__ void just_offset(unsigned bs, int16_t *b1, int16_t *b2, int offAdd) {
        for (unsigned i = g.k; i < bs; i += TL) { b1[i] = (int16_t) sat16(b1[i] + offAdd); b2[i] = (int16_t) sat16(b2[i] + offAdd); }
        g.sync();
    }
    // shift_and_offset (scan_block.rs:1102-1125): drop the first STEP entries, append the first STEP of temp
    __device__ __forceinline__ int shift_and_offset(unsigned bs, int16_t *b1, int16_t *b2, const int16_t *t1, const int16_t *t2, int offAdd) {
        const int cornerOut = sat16(b1[TB_STEP - 1] + offAdd);
        g.sync();
        for (unsigned i = 0; i < bs; i += TL) {
            const unsigned src = i + g.k + TB_STEP;
            int v1, v2;
            if (src < bs) { v1 = sat16(b1[src] + offAdd); v2 = sat16(b2[src] + offAdd); }
            else { v1 = t1[src - bs]; v2 = t2[src - bs]; }
            g.sync();
            b1[i + g.k] = (int16_t) v1;
            b2[i + g.k] = (int16_t) v2;
            g.sync();
        }
        return cornerOut;
    }
    __device__ __forceinline__ int prefix_hmax8(const int16_t *p) const {
        int m = p[0];
#pragma unroll
        for (int x = 1; x < TB_STEP; x++) m = max(m, (int) p[x]);
        return m;
    }
    __device__ __forceinline__ int suffix_hmax(const int16_t *p, unsigned len) const { return max((int) p[len - 1], (int) p[len - 2]); }
    __device__ __forceinline__ void save(unsigned n) {
        for (unsigned i = g.k; i < n; i += TL) { DcolK[i] = Dcol[i]; CcolK[i] = Ccol[i]; DrowK[i] = Drow[i]; RrowK[i] = Rrow[i]; }
        g.sync();
    }
    __device__ __forceinline__ void restore(unsigned n) {
        for (unsigned i = g.k; i < n; i += TL) { Dcol[i] = DcolK[i]; Ccol[i] = CcolK[i]; Drow[i] = DrowK[i]; Rrow[i] = RrowK[i]; }
        g.sync();
    }

    // Block::align_3di + align_core (scan_block.rs:973-1014, :120-640)
    __device__ void align(unsigned minSize, unsigned maxSize, int xDrop) {
        {   // Allocated::clear (:1704-1723). A block only grows while it fits inside both sequences (:491-539), so entries
            // beyond max(minSize, 2 * min(qn, tn)) are never touched and need no reset.
            const unsigned lim = min(maxSize, max(minSize, 2u * (unsigned) min(qn, tn)));
            for (unsigned i = g.k; i < lim; i += TL) {
                Dcol[i] = TB_MIN; Ccol[i] = TB_MIN; Drow[i] = TB_MIN; Rrow[i] = TB_MIN;
                DcolK[i] = TB_MIN; CcolK[i] = TB_MIN; DrowK[i] = TB_MIN; RrowK[i] = TB_MIN;
            }
        }
        temp1[g.k] = TB_MIN; temp2[g.k] = TB_MIN;
        g.sync();
        traceIdx = ckptTraceIdx = 0; blockIdx_ = ckptBlockIdx = 0;
        const unsigned qLen = (unsigned) qn, rLen = (unsigned) tn;
        unsigned si = 0, sj = 0;
        int bestMax = 0;
        unsigned bestI = 0, bestJ = 0;
        int prevDir = 2, dir = 2;   // 0 right, 1 down, 2 grow
        unsigned prevSize = 0, bs = minSize;
        int off = 0, prevOff, offMax = 0;
        unsigned yDropIter = 0, xDropIter = 0;
        unsigned iCkpt = 0, jCkpt = 0;
        int offCkpt = 0;
        int corner = TB_MIN;
        for (;;) {
            prevOff = off;
            int gmax = TB_MIN, gai = 0, gaj = 0, dmax, ai, aj, rightMax, downMax;
            if (dir == 0) {
                off = offMax;
                const int offAdd = sat16(prevOff - off);
                add_block(si, sj + bs - TB_STEP, TB_STEP, bs, true);
                just_offset(bs, Dcol, Ccol, offAdd);
                place_block(si, sj + bs - TB_STEP, TB_STEP, bs, Dcol, Ccol, temp1, temp2,
                            prevDir == 1 ? sat16(corner + offAdd) : TB_MIN, true, dmax, ai, aj);
                rightMax = prefix_hmax8(Dcol);
                corner = shift_and_offset(bs, Drow, Rrow, temp1, temp2, offAdd);
                downMax = prefix_hmax8(Drow);
            } else if (dir == 1) {
                off = offMax;
                const int offAdd = sat16(prevOff - off);
                add_block(si + bs - TB_STEP, sj, bs, TB_STEP, false);
                just_offset(bs, Drow, Rrow, offAdd);
                place_block(sj, si + bs - TB_STEP, TB_STEP, bs, Drow, Rrow, temp1, temp2,
                            prevDir == 0 ? sat16(corner + offAdd) : TB_MIN, false, dmax, ai, aj);
                downMax = prefix_hmax8(Drow);
                corner = shift_and_offset(bs, Dcol, Ccol, temp1, temp2, offAdd);
                rightMax = prefix_hmax8(Dcol);
            } else {
                corner = TB_MIN;
                const unsigned growStep = bs - prevSize;
                add_block(si + prevSize, sj, prevSize, growStep, false);
                place_block(sj, si + prevSize, growStep, prevSize, Drow, Rrow, Dcol + prevSize, Ccol + prevSize, TB_MIN, false, gmax, gai, gaj);
                add_block(si, sj + prevSize, growStep, bs, true);
                place_block(si, sj + prevSize, growStep, bs, Dcol, Ccol, Drow + prevSize, Rrow + prevSize, TB_MIN, true, dmax, ai, aj);
                rightMax = prefix_hmax8(Dcol);
                downMax = prefix_hmax8(Drow);
                save(bs);
                ckptTraceIdx = traceIdx; ckptBlockIdx = blockIdx_;
            }
            if (overflow) return;

            prevDir = dir;
            const int DmaxMax = g.hmax(dmax), growMax = g.hmax(gmax);
            const int mx = max(DmaxMax, growMax);
            offMax = off + mx - TB_ZERO;
            yDropIter++;
            bool growNoMax = dir == 2;

            if (offMax > bestMax) {
                const bool useGrow = dir == 2 && DmaxMax < growMax;
                const int cm = useGrow ? growMax : DmaxMax;
                const int lv = useGrow ? gmax : dmax;
                const unsigned idxI = (unsigned) (useGrow ? gai : ai), idxJ = (unsigned) (useGrow ? gaj : aj);
                unsigned gi, gj;
                if (useGrow) { gi = si + prevSize + idxJ; gj = sj + idxI + g.k; }
                else if (dir == 0) { gi = si + idxI + g.k; gj = sj + (bs - TB_STEP) + idxJ; }
                else if (dir == 1) { gi = si + (bs - TB_STEP) + idxJ; gj = sj + idxI + g.k; }
                else { gi = si + idxI + g.k; gj = sj + prevSize + idxJ; }
                // larger j, then larger i, starting from (0, 0) (scan_block.rs:374-444)
                const unsigned long long key = lv == cm ? (((unsigned long long) gj << 32) | gi) : 0ull;
                const unsigned long long best = g.hmax64(key);
                bestI = (unsigned) (best & 0xFFFFFFFFull);
                bestJ = (unsigned) (best >> 32);
                if (bs < maxSize) {
                    iCkpt = si; jCkpt = sj; offCkpt = off;
                    save(bs);
                    ckptTraceIdx = traceIdx; ckptBlockIdx = blockIdx_;
                    growNoMax = false;
                }
                bestMax = offMax;
                yDropIter = 0;
            }

            if (offMax < bestMax - xDrop) {
                if (xDropIter < 1) xDropIter++;   // X_DROP_ITER = 2
                else break;
            } else {
                xDropIter = 0;
            }

            if (si + bs > qLen && sj + bs > rLen) break;
            if (sj + bs > rLen) { si += TB_STEP; dir = 1; continue; }
            if (si + bs > qLen) { sj += TB_STEP; dir = 0; continue; }

            const unsigned nextSize = bs * 2;
            if (nextSize <= maxSize) {
                if (yDropIter > bs / TB_STEP - 1 || growNoMax) {
                    prevSize = bs; bs = nextSize; dir = 2;
                    si = iCkpt; sj = jCkpt; off = offCkpt;
                    restore(prevSize);
                    traceIdx = ckptTraceIdx; blockIdx_ = ckptBlockIdx;
                    yDropIter = 0;
                    continue;
                }
            }

            if (bs > minSize && yDropIter == 0) {
                const int shrinkMax = max(suffix_hmax(Drow, bs), suffix_hmax(Dcol, bs));
                if (shrinkMax >= mx) {
                    prevDir = 2;
                    bs /= 2;
                    for (unsigned i = g.k; i < bs; i += TL) {
                        Dcol[i] = Dcol[i + bs]; Ccol[i] = Ccol[i + bs]; Drow[i] = Drow[i + bs]; Rrow[i] = Rrow[i + bs];
                    }
                    g.sync();
                    si += bs; sj += bs;
                    iCkpt = si; jCkpt = sj; offCkpt = off;
                    save(bs);
                    rightMax = prefix_hmax8(Dcol);
                    downMax = prefix_hmax8(Drow);
                    ckptTraceIdx = traceIdx; ckptBlockIdx = blockIdx_;
                    yDropIter = 0;
                }
            }

            if (downMax > rightMax) { si += TB_STEP; dir = 1; }
            else { sj += TB_STEP; dir = 0; }
        }
        resScore = bestMax; resI = bestI; resJ = bestJ;
    }

    // Trace::cigar_core (scan_block.rs:1844-2006) fused with the caller's CIGAR loop (StructureSmithWaterman.cpp:491-519).
    // Walking from the block-aligner end cell back to its origin visits the alignment from its start to its end in the
    // original orientation, which is the order of the final backtrace string. Lane 0 only.
    __device__ void walk(char *bt, unsigned &identical, unsigned &alnLen) const {
        unsigned i = resI, j = resJ;
        unsigned bIdx = blockIdx_;
        unsigned long long tIdx = traceIdx;
        unsigned bi = 0, bj = 0, bw = 0, bh = 0, right = 0;
        int table = 0;   // 0 D, 1 C, 2 R
        unsigned n = 0, ids = 0;
        while (i > 0 || j > 0) {
            for (;;) {
                bIdx--;
                const uint4 b = blocks[bIdx];
                bi = b.x; bj = b.y; bh = b.z & 0xFFFFu; bw = b.z >> 16; right = b.w;
                tIdx -= (unsigned long long) bw * bh / TL;
                if (i >= bi && j >= bj) break;
            }
            while (i >= bi && j >= bj && (i > 0 || j > 0)) {
                const unsigned ci = i - bi, cj = j - bj;
                const unsigned long long cell = right ? (tIdx + ci / TL + (unsigned long long) cj * (bh / TL)) * TL + (ci % TL)
                                                      : (tIdx + cj / TL + (unsigned long long) ci * (bw / TL)) * TL + (cj % TL);
                const unsigned t4 = trace[cell];
                const unsigned tLo = t4 & 1, t = t4 & 3, t2Lo = (t4 >> 2) & 1, t2Hi = (t4 >> 3) & 1;
                // first-listed gap table: C when the block was placed right, R when placed down (LUT :1862-1925)
                char op; int next;
                const int firstTab = right ? 1 : 2, secondTab = right ? 2 : 1;
                const char firstOp = right ? 'D' : 'I', secondOp = right ? 'I' : 'D';
                if (table == firstTab) { op = firstOp; next = t2Lo ? 0 : firstTab; }
                else if (table == secondTab) { op = secondOp; next = t2Hi ? 0 : secondTab; }
                else if (t == 0) { op = 'M'; next = 0; }
                else if (tLo) { op = firstOp; next = t2Lo ? 0 : firstTab; }
                else { op = secondOp; next = t2Hi ? 0 : secondTab; }
                if (op == 'M') {
                    if (qa[qEnd - (int) (i - 1)] == ta[dbEnd - (int) (j - 1)]) ids++;
                    i--; j--;
                } else if (op == 'I') i--;
                else j--;
                table = next;
                bt[n++] = op;
            }
        }
        identical = ids; alnLen = n;
    }
};

__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(TbArgs a) {
    __shared__ int8_t sAA[TB_STRIDE * TB_STRIDE], sS[TB_STRIDE * TB_STRIDE];
    for (int x = threadIdx.x; x < TB_STRIDE * TB_STRIDE; x += blockDim.x) { sAA[x] = a.tabAA[x]; sS[x] = a.tab3Di[x]; }
    __syncthreads();
    Engine e;
    e.g.k = threadIdx.x & (TL - 1);
    e.g.mask = 0xFFFFu << (threadIdx.x & 16);
    e.a = &a;
    e.tabAA = sAA; e.tab3Di = sS;
    e.gOpen = a.gapOpen; e.gExt = a.gapExtend;
    e.init_consts();
    const unsigned group = (blockIdx.x * blockDim.x + threadIdx.x) / TL;
    int16_t *st = a.state + (size_t) group * TB_STATE;
    e.Dcol = st; e.Ccol = st + TB_MAXS; e.Drow = st + 2 * TB_MAXS; e.Rrow = st + 3 * TB_MAXS;
    e.DcolK = st + 4 * TB_MAXS; e.CcolK = st + 5 * TB_MAXS; e.DrowK = st + 6 * TB_MAXS; e.RrowK = st + 7 * TB_MAXS;
    e.temp1 = st + 8 * TB_MAXS; e.temp2 = e.temp1 + TL;
    e.trace = a.trace + (size_t) group * a.traceCapVec * TL;
    e.traceCap = a.traceCapVec;
    e.blocks = a.blocks + (size_t) group * a.blockCap;
    e.blockCap = a.blockCap;
    for (;;) {
        unsigned slot = 0;
        if (e.g.k == 0) slot = atomicAdd(a.counter, 1u);
        slot = (unsigned) e.g.idx((int) slot, 0);
        if (slot >= a.nPairs) break;
        const unsigned pi = a.order[slot];
        const TbPair p = a.pairs[pi];
        e.qa = a.qAA + p.qOff; e.qs = a.q3Di + p.qOff; e.qba = a.qBiasAA + p.qOff; e.qbs = a.qBiasSS + p.qOff;
        e.ta = a.tAA + p.tOff; e.ts = a.t3Di + p.tOff;
        e.qEnd = p.qEnd; e.dbEnd = p.dbEnd; e.qn = p.qEnd + 1; e.tn = p.dbEnd + 1;
        e.overflow = false;
        e.resScore = -1000000000; e.resI = 0; e.resJ = 0;
        // exponential search on the minimum block size (StructureSmithWaterman.cpp:458-469)
        unsigned minSize = 32;
        if (a.testMinSize) e.align(a.testMinSize, a.testMaxSize, a.testXDrop);
        else while (minSize <= (unsigned) TB_MAXS && e.resScore < p.targetScore && !e.overflow) {
            const int xDrop = -((int) minSize * a.gapExtend + a.gapOpen);
            e.align(minSize, TB_MAXS, xDrop);
            minSize *= 2;
        }
        TbOut o;
        o.qStart = -1; o.dbStart = -1; o.identical = 0; o.alnLen = 0; o.score = e.resScore;
        if (e.overflow) o.ok = -1;
        else if (!a.testMinSize && e.resScore != p.targetScore && !(p.targetScore == 32767 && e.resScore >= p.targetScore)) o.ok = 0;
        else {
            o.ok = 1;
            e.g.sync();
            if (e.g.k == 0) {
                unsigned ids = 0, n = 0;
                e.walk(a.bt + p.btOff, ids, n);
                o.identical = ids; o.alnLen = n;
                o.qStart = p.qEnd + 1 - (int) e.resI;
                o.dbStart = p.dbEnd + 1 - (int) e.resJ;
            }
        }
        if (e.g.k == 0) a.out[pi] = o;
        e.g.sync();
    }
}

}  // namespace tb
