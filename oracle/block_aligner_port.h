// oracle/block_aligner_port.h — TEST INFRASTRUCTURE (CPU restatement, never linked into libfsb200.so).
//
// Scalar restatement of the block-aligner engine the reference links as a Rust crate
// (lib/mmseqs/lib/block-aligner, AVX2 feature: 16 int16 lanes, ZERO = 1<<14, MIN = 0, src/avx2.rs:11-16).
// There is no Rust toolchain in this image, so the crate cannot be built; this file restates its published
// algorithm lane by lane (every SIMD intrinsic is emulated on int16[16], including the 128-bit lane
// structure of the byte shifts) and is pinned by the crate's own unit-test vectors
// (src/scan_block.rs:2267-2413, tests/test_block_aligner_oracle.py).
//
//   align_core      src/scan_block.rs:120-640    block trajectory: grow / shift right / shift down / shrink,
//                                                32-bit offsets, checkpoints, x-drop termination
//   place_block     src/scan_block.rs:1145-1283  one rectangular region, column by column, prefix-scan R
//   helpers         src/scan_block.rs:1065-1125  just_offset, prefix_max, suffix_max, shift_and_offset
//   Trace/cigar     src/scan_block.rs:1726-2006  2+2 trace bits per cell, block stack, LUT walk
//   prefix scan     src/avx2.rs:296-337
#ifndef BLOCK_ALIGNER_PORT_H
#define BLOCK_ALIGNER_PORT_H
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace ba {

constexpr int L = 16;
constexpr int16_t ZERO = 1 << 14;
constexpr int16_t VMIN = 0;
constexpr size_t STEP = 8;                  // scan_block.rs:813
constexpr size_t X_DROP_ITER = 2;           // :814
constexpr size_t SHRINK_SUFFIX_LEN = STEP / 4;  // :816

struct V {
    int16_t v[L];
};

static inline int16_t sat16(int x) { return (int16_t) std::min(32767, std::max(-32768, x)); }
static inline V set1(int16_t x) { V r; for (int k = 0; k < L; k++) r.v[k] = x; return r; }
static inline V load(const int16_t *p) { V r; memcpy(r.v, p, sizeof(r.v)); return r; }
static inline void store(int16_t *p, const V &a) { memcpy(p, a.v, sizeof(a.v)); }
static inline V adds(const V &a, const V &b) { V r; for (int k = 0; k < L; k++) r.v[k] = sat16((int) a.v[k] + b.v[k]); return r; }
static inline V subs(const V &a, const V &b) { V r; for (int k = 0; k < L; k++) r.v[k] = sat16((int) a.v[k] - b.v[k]); return r; }
static inline V vmax(const V &a, const V &b) { V r; for (int k = 0; k < L; k++) r.v[k] = std::max(a.v[k], b.v[k]); return r; }
static inline V cmpeq(const V &a, const V &b) { V r; for (int k = 0; k < L; k++) r.v[k] = a.v[k] == b.v[k] ? -1 : 0; return r; }
// simd_sl_i16!(a, b, 1) (avx2.rs:103-118): lanes move up by one, lane 0 takes the top lane of b
static inline V sl1(const V &a, const V &b) { V r; r.v[0] = b.v[L - 1]; for (int k = 1; k < L; k++) r.v[k] = a.v[k - 1]; return r; }
// simd_step(a, b) = _mm256_permute2x128_si256(a, b, 0x03): low half = high half of b, high half = low half of a
static inline V step8(const V &a, const V &b) { V r; for (int k = 0; k < 8; k++) { r.v[k] = b.v[8 + k]; r.v[8 + k] = a.v[k]; } return r; }
// _mm256_slli_si256(a, 2n): shift inside each 128-bit half, zeros shifted in
static inline V sllz(const V &a, int n) {
    V r;
    for (int h = 0; h < 2; h++) for (int k = 0; k < 8; k++) r.v[8 * h + k] = k >= n ? a.v[8 * h + k - n] : 0;
    return r;
}
static inline V shl_bits(const V &a, int n) { V r; for (int k = 0; k < L; k++) r.v[k] = (int16_t) ((uint16_t) a.v[k] << n); return r; }
static inline V broadcasthi(const V &a) { return set1(a.v[L - 1]); }
static inline int16_t hmax(const V &a) { int16_t m = a.v[0]; for (int k = 1; k < L; k++) m = std::max(m, a.v[k]); return m; }

// get_prefix_scan_consts (avx2.rs:296-310)
static inline void prefix_scan_consts(const V &gap, V &gapAll, V &consts) {
    V s1 = adds(sllz(gap, 1), gap);
    V s2 = adds(sllz(s1, 2), s1);
    V s4 = adds(sllz(s2, 4), s2);
    // shufflehi(0xFF) -> words 4..7 of each half = word 7; srli_si256(8) -> words 0..3 = old words 4..7, words 4..7 = 0
    V c;
    for (int h = 0; h < 2; h++) for (int k = 0; k < 8; k++) c.v[8 * h + k] = k < 4 ? s4.v[8 * h + 7] : 0;
    // permute4x64(0b00000101): q0 = src q1, q1 = src q1, q2 = src q0, q3 = src q0
    V p;
    for (int k = 0; k < 4; k++) { p.v[k] = c.v[4 + k]; p.v[4 + k] = c.v[4 + k]; p.v[8 + k] = c.v[k]; p.v[12 + k] = c.v[k]; }
    gapAll = adds(p, s4);
    consts = s4;
}

// simd_prefix_scan_i16 (avx2.rs:312-337)
static inline V prefix_scan(const V &R, const V &gap, const V &consts) {
    V s1 = vmax(R, adds(sllz(R, 1), gap));
    V s2 = vmax(s1, adds(sllz(s1, 2), shl_bits(gap, 1)));
    V s4 = vmax(s2, adds(sllz(s2, 4), shl_bits(gap, 2)));
    // shufflehi(0xFF) then permute4x64(0b01010000): q0 = q1 = src q0, q2 = q3 = src q1 (= word 7 of the low half)
    V c;
    for (int k = 0; k < 4; k++) { c.v[k] = s4.v[k]; c.v[4 + k] = s4.v[k]; c.v[8 + k] = s4.v[7]; c.v[12 + k] = s4.v[7]; }
    c = adds(c, consts);
    return vmax(s4, c);
}

static inline int16_t prefix_hmax8(const int16_t *p) { int16_t m = p[0]; for (size_t k = 1; k < STEP; k++) m = std::max(m, p[k]); return m; }
static inline int16_t suffix_hmax(const int16_t *p, size_t len) {
    int16_t m = p[len - 1];
    for (size_t k = 1; k < SHRINK_SUFFIX_LEN; k++) m = std::max(m, p[len - 1 - k]);
    return m;
}
static inline int16_t clamp16(int32_t x) { return (int16_t) std::min<int32_t>(std::max<int32_t>(x, INT16_MIN), INT16_MAX); }

// What place_block needs from the sequences and the scoring scheme.  `right` = vectors run along the query;
// vecPos indexes the vector sequence (16 consecutive padded positions), scalarPos the other one.
struct Scorer {
    virtual ~Scorer() {}
    virtual size_t query_len() const = 0;
    virtual size_t reference_len() const = 0;
    virtual V scores(bool right, size_t vecPos, size_t scalarPos) const = 0;
};

enum Op : uint8_t { OP_SENTINEL = 0, OP_M = 1, OP_EQ = 2, OP_X = 3, OP_I = 4, OP_D = 5 };  // cigar.rs:10-31

struct Trace {
    std::vector<uint32_t> trace, trace2;
    std::vector<uint8_t> right;
    std::vector<uint32_t> blockStart;
    std::vector<uint16_t> blockSize;
    size_t traceIdx = 0, blockIdx = 0, ckptTraceIdx = 0, ckptBlockIdx = 0;
    size_t queryLen = 0, referenceLen = 0;

    void clear(size_t q, size_t r) { traceIdx = blockIdx = ckptTraceIdx = ckptBlockIdx = 0; queryLen = q; referenceLen = r; }
    void add_trace(uint32_t t, uint32_t t2) {
        if (traceIdx >= trace.size()) { size_t n = std::max<size_t>(1024, trace.size() * 2); trace.resize(n); trace2.resize(n); }
        trace[traceIdx] = t;
        trace2[traceIdx] = t2;
        traceIdx++;
    }
    void add_block(size_t i, size_t j, size_t width, size_t height, bool r) {  // scan_block.rs:1790-1806
        if (blockIdx * 2 + 1 >= blockStart.size()) { size_t n = std::max<size_t>(256, blockStart.size() * 2); blockStart.resize(n); blockSize.resize(n); right.resize(n); }
        blockStart[blockIdx * 2] = (uint32_t) i;
        blockStart[blockIdx * 2 + 1] = (uint32_t) j;
        blockSize[blockIdx * 2] = (uint16_t) height;
        blockSize[blockIdx * 2 + 1] = (uint16_t) width;
        right[blockIdx] = r;
        blockIdx++;
    }
    void add_trace_idx(size_t add) { traceIdx += add; if (traceIdx > trace.size()) { trace.resize(traceIdx); trace2.resize(traceIdx); } }
    void save_ckpt() { ckptTraceIdx = traceIdx; ckptBlockIdx = blockIdx; }
    void restore_ckpt() { traceIdx = ckptTraceIdx; blockIdx = ckptBlockIdx; }

    // one LUT entry of cigar_core (scan_block.rs:1862-1925): table 0 = D, 1 = C, 2 = R
    static void lut(bool right, unsigned t, unsigned t2, int table, Op &op, int &di, int &dj, int &next) {
        const bool t_lo = t & 1, t_hi = (t >> 1) & 1, t2_lo = t2 & 1, t2_hi = (t2 >> 1) & 1;
        if (right) {
            if (table == 1) { op = OP_D; di = 0; dj = 1; next = t2_lo ? 0 : 1; return; }
            if (table == 2) { op = OP_I; di = 1; dj = 0; next = t2_hi ? 0 : 2; return; }
            if (t == 0) { op = OP_M; di = 1; dj = 1; next = 0; return; }
            if (t_lo) { op = OP_D; di = 0; dj = 1; next = t2_lo ? 0 : 1; return; }
            (void) t_hi;
            op = OP_I; di = 1; dj = 0; next = t2_hi ? 0 : 2;
        } else {
            if (table == 2) { op = OP_I; di = 1; dj = 0; next = t2_lo ? 0 : 2; return; }
            if (table == 1) { op = OP_D; di = 0; dj = 1; next = t2_hi ? 0 : 1; return; }
            if (t == 0) { op = OP_M; di = 1; dj = 1; next = 0; return; }
            if (t_lo) { op = OP_I; di = 1; dj = 0; next = t2_lo ? 0 : 2; return; }
            op = OP_D; di = 0; dj = 1; next = t2_hi ? 0 : 1;
        }
    }

    // cigar_core (scan_block.rs:1844-2006); returns run-length encoded ops in forward order.
    // eqQuery/eqRef: padded byte strings for the =/X variant (may be NULL).
    std::vector<std::pair<Op, size_t>> cigar(size_t i, size_t j, const uint8_t *eqQuery = NULL, const uint8_t *eqRef = NULL) const {
        std::vector<std::pair<Op, size_t>> rev;
        size_t bIdx = blockIdx, tIdx = traceIdx;
        size_t bi = 0, bj = 0, bw = 0, bh = 0;
        bool r = false;
        int table = 0;
        while (i > 0 || j > 0) {
            for (;;) {
                bIdx--;
                bi = blockStart[bIdx * 2];
                bj = blockStart[bIdx * 2 + 1];
                bh = blockSize[bIdx * 2];
                bw = blockSize[bIdx * 2 + 1];
                tIdx -= bw * bh / L;
                if (i >= bi && j >= bj) { r = right[bIdx]; break; }
            }
            while (i >= bi && j >= bj && (i > 0 || j > 0)) {
                const size_t ci = i - bi, cj = j - bj;
                size_t idx, sh;
                if (r) { idx = tIdx + ci / L + cj * (bh / L); sh = (ci % L) * 2; }
                else   { idx = tIdx + cj / L + ci * (bw / L); sh = (cj % L) * 2; }
                const unsigned t = (trace[idx] >> sh) & 3, t2 = (trace2[idx] >> sh) & 3;
                Op op; int di, dj, next;
                lut(r, t, t2, table, op, di, dj, next);
                if (eqQuery != NULL && op == OP_M) op = eqQuery[i] == eqRef[j] ? OP_EQ : OP_X;
                i -= di;
                j -= dj;
                table = next;
                if (!rev.empty() && rev.back().first == op) rev.back().second++;
                else rev.push_back(std::make_pair(op, (size_t) 1));
            }
        }
        std::reverse(rev.begin(), rev.end());
        return rev;
    }
};

struct AlignResult {
    int32_t score;
    size_t query_idx, reference_idx;
};

class Block {
public:
    Block(bool traceFlag, bool xDrop) : TRACE(traceFlag), X_DROP(xDrop) { res.score = 0; res.query_idx = res.reference_idx = 0; }
    AlignResult res;
    Trace trace;

    // Block::align / align_3di (scan_block.rs:862-1014): gapOpen / gapExtend negative, open includes the first extension
    void align(const Scorer &sc, int gapOpen, int gapExtend, size_t minSize, size_t maxSize, int32_t xDrop) {
        minSize = std::max<size_t>(minSize, L);
        maxSize = std::max<size_t>(maxSize, L);
        gOpen = (int16_t) gapOpen;
        gExt = (int16_t) gapExtend;
        trace.clear(sc.query_len(), sc.reference_len());
        for (int a = 0; a < 8; a++) buf[a].assign(maxSize, VMIN);
        for (int k = 0; k < L; k++) temp1[k] = temp2[k] = VMIN;
        core(sc, minSize, maxSize, xDrop);
    }

private:
    enum Dir { RIGHT, DOWN, GROW };
    const bool TRACE, X_DROP;
    int16_t gOpen, gExt;
    std::vector<int16_t> buf[8];  // D_col C_col D_row R_row and their checkpoints
    int16_t temp1[L], temp2[L];

    static void just_offset(size_t bs, int16_t *b1, int16_t *b2, const V &offAdd) {
        for (size_t i = 0; i < bs; i += L) { store(b1 + i, adds(load(b1 + i), offAdd)); store(b2 + i, adds(load(b2 + i), offAdd)); }
    }
    static V shift_and_offset(size_t bs, int16_t *b1, int16_t *b2, const int16_t *t1, const int16_t *t2, const V &offAdd) {
        V curr1 = adds(load(b1), offAdd);
        V corner = set1(curr1.v[STEP - 1]);
        V curr2 = adds(load(b2), offAdd);
        size_t i = 0;
        while (i < bs - L) {
            V next1 = adds(load(b1 + i + L), offAdd);
            V next2 = adds(load(b2 + i + L), offAdd);
            store(b1 + i, step8(next1, curr1));
            store(b2 + i, step8(next2, curr2));
            curr1 = next1;
            curr2 = next2;
            i += L;
        }
        store(b1 + bs - L, step8(load(t1), curr1));
        store(b2 + bs - L, step8(load(t2), curr2));
        return corner;
    }

    struct Best { V dmax, ai, aj; };

    // place_block (scan_block.rs:1145-1283); qLen/rLen are the lengths of this call's vector / scalar sequence
    Best place_block(const Scorer &sc, size_t qLen, size_t rLen, size_t startI, size_t startJ, size_t width, size_t height,
                     int16_t *Dcol, int16_t *Ccol, int16_t *Drow, int16_t *Rrow, V Dcorner, bool right) {
        const V gapOpen = set1(gOpen), gapExtend = set1(gExt);
        V gapExtendAll, consts;
        prefix_scan_consts(gapExtend, gapExtendAll, consts);
        Best b;
        b.dmax = set1(VMIN);
        b.ai = set1(0);
        b.aj = set1(0);
        if (width == 0 || height == 0) return b;
        for (size_t j = 0; j < width; j++) {
            V R01 = set1(VMIN), D11 = set1(VMIN), R11 = set1(VMIN), prevTraceR = set1(0);
            for (size_t i = 0; i < height; i += L) {
                const V D10 = load(Dcol + i), C10 = load(Ccol + i);
                const V D00 = sl1(D10, Dcorner);
                Dcorner = D10;
                D11 = adds(D00, sc.scores(right, startI + i, startJ + j));
                if (startI + i == 0 && startJ + j == 0) D11.v[0] = ZERO;
                const V C11open = adds(D10, gapOpen);
                const V C11 = vmax(adds(C10, gapExtend), C11open);
                D11 = vmax(D11, C11);
                const V D11open = adds(D11, subs(gapOpen, gapExtend));
                R11 = prefix_scan(D11open, gapExtend, consts);
                R11 = vmax(R11, adds(broadcasthi(R01), gapExtendAll));
                D11 = vmax(D11, R11);
                R01 = R11;
                if (TRACE) {
                    const V tDC = cmpeq(D11, C11), tDR = cmpeq(D11, R11);
                    const V tempTraceR = cmpeq(R11, D11open);
                    const V traceR = sl1(tempTraceR, prevTraceR);
                    const V tCopen = cmpeq(C11, C11open);
                    uint32_t t = 0, t2 = 0;
                    for (int k = 0; k < L; k++) {
                        t |= (uint32_t) (tDC.v[k] != 0) << (2 * k) | (uint32_t) (tDR.v[k] != 0) << (2 * k + 1);
                        t2 |= (uint32_t) (tCopen.v[k] != 0) << (2 * k) | (uint32_t) (traceR.v[k] != 0) << (2 * k + 1);
                    }
                    prevTraceR = tempTraceR;
                    trace.add_trace(t, t2);
                }
                b.dmax = vmax(b.dmax, D11);
                if (X_DROP) {
                    for (int k = 0; k < L; k++) if (b.dmax.v[k] == D11.v[k]) { b.ai.v[k] = (int16_t) i; b.aj.v[k] = (int16_t) j; }
                }
                store(Dcol + i, D11);
                store(Ccol + i, C11);
            }
            Dcorner = set1(VMIN);
            Drow[j] = D11.v[L - 1];
            Rrow[j] = R11.v[L - 1];
            if (!X_DROP && startI + height > qLen && startJ + j >= rLen) {
                if (TRACE) trace.add_trace_idx((width - 1 - j) * (height / L));
                break;
            }
        }
        return b;
    }

    void core(const Scorer &sc, const size_t minSize, const size_t maxSize, const int32_t xDrop) {
        const size_t qLen = sc.query_len(), rLen = sc.reference_len();
        int16_t *Dcol = buf[0].data(), *Ccol = buf[1].data(), *Drow = buf[2].data(), *Rrow = buf[3].data();
        int16_t *DcolK = buf[4].data(), *CcolK = buf[5].data(), *DrowK = buf[6].data(), *RrowK = buf[7].data();
        auto save = [&](size_t n) { memcpy(DcolK, Dcol, n * 2); memcpy(CcolK, Ccol, n * 2); memcpy(DrowK, Drow, n * 2); memcpy(RrowK, Rrow, n * 2); };
        auto restore = [&](size_t n) { memcpy(Dcol, DcolK, n * 2); memcpy(Ccol, CcolK, n * 2); memcpy(Drow, DrowK, n * 2); memcpy(Rrow, RrowK, n * 2); };

        size_t si = 0, sj = 0;
        int32_t bestMax = 0;
        size_t bestI = 0, bestJ = 0;
        Dir prevDir = GROW, dir = GROW;
        size_t prevSize = 0, bs = minSize;
        int32_t off = 0, prevOff, offMax = 0;
        size_t yDropIter = 0, xDropIter = 0;
        size_t iCkpt = si, jCkpt = sj;
        int32_t offCkpt = 0;
        V Dcorner = set1(VMIN);

        for (;;) {
            prevOff = off;
            Best grow;
            grow.dmax = set1(VMIN);
            grow.ai = set1(0);
            grow.aj = set1(0);
            Best cur;
            int16_t rightMax, downMax;
            if (dir == RIGHT) {
                off = offMax;
                const V offAdd = set1(clamp16(prevOff - off));
                if (TRACE) trace.add_block(si, sj + bs - STEP, STEP, bs, true);
                just_offset(bs, Dcol, Ccol, offAdd);
                cur = place_block(sc, qLen, rLen, si, sj + bs - STEP, STEP, bs, Dcol, Ccol, temp1, temp2,
                                  prevDir == DOWN ? adds(Dcorner, offAdd) : set1(VMIN), true);
                rightMax = prefix_hmax8(Dcol);
                Dcorner = shift_and_offset(bs, Drow, Rrow, temp1, temp2, offAdd);
                downMax = prefix_hmax8(Drow);
            } else if (dir == DOWN) {
                off = offMax;
                const V offAdd = set1(clamp16(prevOff - off));
                if (TRACE) trace.add_block(si + bs - STEP, sj, bs, STEP, false);
                just_offset(bs, Drow, Rrow, offAdd);
                cur = place_block(sc, rLen, qLen, sj, si + bs - STEP, STEP, bs, Drow, Rrow, temp1, temp2,
                                  prevDir == RIGHT ? adds(Dcorner, offAdd) : set1(VMIN), false);
                downMax = prefix_hmax8(Drow);
                Dcorner = shift_and_offset(bs, Dcol, Ccol, temp1, temp2, offAdd);
                rightMax = prefix_hmax8(Dcol);
            } else {
                Dcorner = set1(VMIN);
                const size_t growStep = bs - prevSize;
                if (TRACE) trace.add_block(si + prevSize, sj, prevSize, growStep, false);
                grow = place_block(sc, rLen, qLen, sj, si + prevSize, growStep, prevSize, Drow, Rrow, Dcol + prevSize, Ccol + prevSize,
                                   set1(VMIN), false);
                if (TRACE) trace.add_block(si, sj + prevSize, growStep, bs, true);
                cur = place_block(sc, qLen, rLen, si, sj + prevSize, growStep, bs, Dcol, Ccol, Drow + prevSize, Rrow + prevSize,
                                  set1(VMIN), true);
                rightMax = prefix_hmax8(Dcol);
                downMax = prefix_hmax8(Drow);
                save(bs);
                if (TRACE) trace.save_ckpt();
            }

            prevDir = dir;
            const int16_t DmaxMax = hmax(cur.dmax), growMax = hmax(grow.dmax);
            const int16_t mx = std::max(DmaxMax, growMax);
            offMax = off + (int32_t) mx - (int32_t) ZERO;
            yDropIter++;
            bool growNoMax = dir == GROW;

            if (offMax > bestMax) {
                if (X_DROP) {
                    size_t bI = 0, bJ = 0;
                    const bool useGrow = dir == GROW && DmaxMax < growMax;
                    const Best &c = useGrow ? grow : cur;
                    const int16_t cm = useGrow ? growMax : DmaxMax;
                    for (int lane = 0; lane < L; lane++) {
                        if (c.dmax.v[lane] != cm) continue;
                        const size_t idxI = (size_t) c.ai.v[lane], idxJ = (size_t) c.aj.v[lane];
                        const size_t r = idxI + lane, cc = (bs - STEP) + idxJ;
                        size_t gi, gj;
                        if (useGrow) { gi = si + prevSize + idxJ; gj = sj + idxI + lane; }
                        else if (dir == RIGHT) { gi = si + r; gj = sj + cc; }
                        else if (dir == DOWN) { gi = si + cc; gj = sj + r; }
                        else { gi = si + idxI + lane; gj = sj + prevSize + idxJ; }
                        const bool better = gj != bJ ? gj > bJ : gi > bI;
                        if (better) { bI = gi; bJ = gj; }
                    }
                    bestI = bI;
                    bestJ = bJ;
                }
                if (bs < maxSize) {
                    iCkpt = si;
                    jCkpt = sj;
                    offCkpt = off;
                    save(bs);
                    if (TRACE) trace.save_ckpt();
                    growNoMax = false;
                }
                bestMax = offMax;
                yDropIter = 0;
            }

            if (X_DROP) {
                if (offMax < bestMax - xDrop) {
                    if (xDropIter < X_DROP_ITER - 1) xDropIter++;
                    else break;
                } else {
                    xDropIter = 0;
                }
            }

            if (si + bs > qLen && sj + bs > rLen) break;
            if (sj + bs > rLen) { si += STEP; dir = DOWN; continue; }
            if (si + bs > qLen) { sj += STEP; dir = RIGHT; continue; }

            const size_t nextSize = bs * 2;
            if (nextSize <= maxSize) {
                if (yDropIter > (bs / STEP) - 1 || growNoMax) {
                    prevSize = bs;
                    bs = nextSize;
                    dir = GROW;
                    si = iCkpt;
                    sj = jCkpt;
                    off = offCkpt;
                    restore(prevSize);
                    if (TRACE) trace.restore_ckpt();
                    yDropIter = 0;
                    continue;
                }
            }

            if (bs > minSize && yDropIter == 0) {  // SHRINK = true (:815)
                const int16_t shrinkMax = std::max(suffix_hmax(Drow, bs), suffix_hmax(Dcol, bs));
                if (shrinkMax >= mx) {
                    prevDir = GROW;
                    bs /= 2;
                    memmove(Dcol, Dcol + bs, bs * 2);
                    memmove(Ccol, Ccol + bs, bs * 2);
                    memmove(Drow, Drow + bs, bs * 2);
                    memmove(Rrow, Rrow + bs, bs * 2);
                    si += bs;
                    sj += bs;
                    iCkpt = si;
                    jCkpt = sj;
                    offCkpt = off;
                    save(bs);
                    rightMax = prefix_hmax8(Dcol);
                    downMax = prefix_hmax8(Drow);
                    if (TRACE) trace.save_ckpt();
                    yDropIter = 0;
                }
            }

            if (downMax > rightMax) { si += STEP; dir = DOWN; }
            else { sj += STEP; dir = RIGHT; }
        }

        if (X_DROP) {
            res.score = bestMax;
            res.query_idx = bestI;
            res.reference_idx = bestJ;
        } else {
            int32_t score;
            if (dir == RIGHT || dir == GROW) score = off + (int32_t) Dcol[qLen - si] - (int32_t) ZERO;
            else score = off + (int32_t) Drow[rLen - sj] - (int32_t) ZERO;
            res.score = score;
            res.query_idx = qLen;
            res.reference_idx = rLen;
        }
    }
};

// Sequence-to-sequence scoring with one int8 table over converted bytes (Matrix::get_scores, scores.rs:128-135):
// `table[a * stride + b]`, both sequences padded: [0] = NULL, 1..len = data, then >= maxSize NULLs.
struct MatrixScorer : Scorer {
    const uint8_t *q, *r;
    size_t ql, rl;
    const int8_t *table;
    int stride;
    size_t query_len() const { return ql; }
    size_t reference_len() const { return rl; }
    V scores(bool right, size_t vecPos, size_t scalarPos) const {
        const uint8_t *vs = right ? q : r;
        const uint8_t c = right ? r[scalarPos] : q[scalarPos];
        V o;
        for (int k = 0; k < L; k++) o.v[k] = table[(int) c * stride + vs[vecPos + k]];
        return o;
    }
};

// AA + 3Di + position bias (place_block_3di, scan_block.rs:1343-1351)
struct Scorer3di : Scorer {
    const uint8_t *qAA, *q3Di, *rAA, *r3Di;   // padded, codes 0..NULLCODE
    const int16_t *qBias, *rBias;              // padded: [0] = 0, 1..len, then zeros
    size_t ql, rl;
    const int8_t *tabAA, *tab3Di;              // stride x stride
    int stride;
    size_t query_len() const { return ql; }
    size_t reference_len() const { return rl; }
    V scores(bool right, size_t vecPos, size_t scalarPos) const {
        const uint8_t *va = right ? qAA : rAA, *vs = right ? q3Di : r3Di;
        const int16_t *vb = right ? qBias : rBias;
        const uint8_t ca = right ? rAA[scalarPos] : qAA[scalarPos], cs = right ? r3Di[scalarPos] : q3Di[scalarPos];
        const int16_t sb = right ? rBias[scalarPos] : qBias[scalarPos];
        V a, s, pb;
        for (int k = 0; k < L; k++) {
            a.v[k] = tabAA[(int) ca * stride + va[vecPos + k]];
            s.v[k] = tab3Di[(int) cs * stride + vs[vecPos + k]];
            pb.v[k] = sat16((int) sb + vb[vecPos + k]);
        }
        return adds(adds(a, s), pb);
    }
};

}  // namespace ba
#endif
