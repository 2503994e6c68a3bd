// Stage-1 kernels: 3Di k-mer prefilter (reference: lib/mmseqs/src/prefiltering, abbreviated M/.. below).
//
//   index build (once per target DB)      IndexBuilder::fillDatabase  M/IndexBuilder.cpp:56-271, IndexTable.h:135-175,350-403
//   similar k-mer enumeration             KmerGenerator::generateKmerList M/KmerGenerator.cpp:177-286
//   double-diagonal matching              QueryMatcher::match + CacheFriendlyOperations::findDuplicates
//                                         M/QueryMatcher.cpp:243-376, M/CacheFriendlyOperations.cpp:185-278
//   ungapped diagonal rescoring           UngappedAlignment::computeScores M/UngappedAlignment.cpp:315-346,45-58
//   best diagonal per target, threshold,  QueryMatcher::matchQuery tail M/QueryMatcher.cpp:143-240
//   top-N with the reference's tie order
//
// GPU decomposition (DESIGN.md "Prefilter"): match.cuh turns every query into its stream of index hits and detects the double
// diagonal hits with per-target state in shared memory (any DB size).  Double hits are appended to a global list keyed by
// (query, target, stream position); a radix sort brings each target's double hits together in stream order for the
// run-collapse, the ungapped rescoring and the per-target maximum; a second sort realises the reference's (score, bin, array
// position) order that decides which hits survive the --max-seqs cut.
#pragma once
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>
#include "../../include/fsb200.h"
#include "host_core.h"

namespace fsb {
namespace prefilter {

constexpr int KA = 20;                 // k-mer alphabet (X excluded)
constexpr int N3 = KA * KA * KA;       // 3-mers
constexpr int KSPAN = 10;              // spaced_seed_6 = 1101010011 (M/../commons/Sequence.h:25)
constexpr int PF_MAXCHUNK = 126;       // overflow chunks per query handled by the replay (2·max(1e6,N) hits each; the rank of a chunk has 7 bits)
constexpr int PF_MAXOV = 128;          // double hits of one target of an overflowed query

struct Tables {
    uint16_t *sIdx = nullptr;     // [N3][N3] 3-mer indices of each row, sorted by score desc (stable in permutation order)
    int16_t *sScore = nullptr;    // [N3][N3]
    uint16_t *cntGe = nullptr;    // [N3][NS]: number of row entries with score >= vmin + s
    uint16_t *colPerm = nullptr;  // [N3]: column of a second 3-mer in the list-head table / occupancy bitmap (similar letters -> neighbouring columns)
    int vmin = 0, NS = 0;
};

struct Elem {                      // one double-diagonal hit
    uint64_t key;                  // (query_local << (32+tBits)) | (target << 32) | stream position
    uint32_t diag;                 // 16-bit diagonal i - j
    uint32_t pad;                  // overflow chunk of the hit (0 unless the query exceeded maxDbMatches)
};

struct QueryDesc {
    uint64_t seqOff;               // into q3Di
    uint64_t thrOff;               // into thr (int16 per window, -1 = window contains X)
    uint64_t profOff;              // into ungapped profile (22 bytes per position)
    uint32_t len;
    uint32_t pad;
};

__device__ __forceinline__ int cnt_ge(const Tables &tb, int row, int v) {
    int s = v - tb.vmin;
    if (s <= 0) return N3;
    if (s >= tb.NS) return 0;
    return (int) tb.cntGe[(size_t) row * tb.NS + s];
}

// Final array position of a chunk's elements after the reference's overflow merges (DESIGN.md "overflow chunks"):
// every merge of an overflowed chunk (mergeDiagonalKeepScoredHitsDuplicates, CFO.cpp:121-152) reverses the array inside a bin,
// the last merge (mergeDiagonalDuplicates, :85-118) keeps it.  Returns for 1-based chunk c of n its block rank and direction.
__device__ inline void chunk_order(int n, int *rank, int *dir) {   // arrays of n+1 entries, index 1..n
    if (n <= 2) { for (int c = 1; c <= n; c++) { rank[c] = c - 1; dir[c] = 1; } return; }
    int ordC[PF_MAXCHUNK + 1], ordD[PF_MAXCHUNK + 1], len = 2;
    ordC[0] = 2; ordD[0] = -1; ordC[1] = 1; ordD[1] = -1;            // O_2 = reverse(S_1 S_2)
    for (int k = 3; k <= n - 1; k++) {                                 // O_k = [(k,-)] ++ reverse(O_{k-1})
        int tc[PF_MAXCHUNK + 1], td[PF_MAXCHUNK + 1];
        tc[0] = k; td[0] = -1;
        for (int j = 0; j < len; j++) { tc[1 + j] = ordC[len - 1 - j]; td[1 + j] = -ordD[len - 1 - j]; }
        len++;
        for (int j = 0; j < len; j++) { ordC[j] = tc[j]; ordD[j] = td[j]; }
    }
    ordC[len] = n; ordD[len] = 1; len++;                              // final = O_{n-1} ++ S_n
    for (int j = 0; j < len; j++) { rank[ordC[j]] = j; dir[ordC[j]] = ordD[j]; }
}

}  // namespace prefilter
}  // namespace fsb
#include "match.cuh"
namespace fsb {
namespace prefilter {

// ---- after the sort by (query, target, stream position): run-collapse + ungapped rescoring ----
struct ScoreArgs {
    const Elem *elems;             // sorted
    uint32_t n;
    uint32_t tBits;
    const QueryDesc *queries;
    const int8_t *profiles;        // per query [len][22]
    const uint8_t *tSeq;           // masked target 3Di (SequenceLookup)
    const uint64_t *tOff;
    const uint32_t *tLen;
    uint8_t *score;                // out: 0..255, per element; 0 for collapsed elements
    uint8_t *kept;
    unsigned long long *stats;     // [3] ungapped cells
};

// scalarDiagonalScoring / computeSingelSequenceScores (UngappedAlignment.cpp:45-58,423-438)
// The target bytes are fetched as aligned 32-bit words (targets are staged on 16-byte boundaries); prof may point to shared memory.
__device__ __forceinline__ int ungapped_diag(const int8_t *prof, int qLen, const uint8_t *db, int dbLen, uint32_t diag16, unsigned long long &cells) {
    const short diag = (short) diag16;
    const unsigned short d1 = (unsigned short) (0u - diag16), d2 = (unsigned short) diag16;
    const int minDist = (int) min(d1, d2);
    const int8_t *p; const uint8_t *d; int len;
    if (diag >= 0 && minDist < qLen) { p = prof + (size_t) minDist * 22; d = db; len = min(dbLen, qLen - minDist); }
    else if (diag < 0 && minDist < dbLen) { p = prof; d = db + minDist; len = min(dbLen - minDist, qLen); }
    else return 0;
    int sc = 0, mx = 0, k = 0;
    // targets are staged on 16-byte boundaries: bytes up to the next boundary, then 16 residues per load
    const int head = min(len, (int) ((16u - (uint32_t) (reinterpret_cast<uintptr_t>(d) & 15u)) & 15u));
    for (; k < head; k++) { sc = max(sc + p[k * 22 + d[k]], 0); mx = max(mx, sc); }
    for (; k + 16 <= len; k += 16) {
        const uint4 w4 = __ldg(reinterpret_cast<const uint4 *>(d + k));
        const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t w = ws[j];
            const int8_t *pk = p + (k + 4 * j) * 22;
            sc = max(sc + pk[w & 0xffu], 0); mx = max(mx, sc);
            sc = max(sc + pk[22 + ((w >> 8) & 0xffu)], 0); mx = max(mx, sc);
            sc = max(sc + pk[44 + ((w >> 16) & 0xffu)], 0); mx = max(mx, sc);
            sc = max(sc + pk[66 + (w >> 24)], 0); mx = max(mx, sc);
        }
    }
    for (; k < len; k++) { sc = max(sc + p[k * 22 + d[k]], 0); mx = max(mx, sc); }
    cells += (unsigned long long) len;
    return mx;
}

// (Measured alternative, git 2a29d90: diagonals regrouped per block by length and diagonal class, lanes time-skewed so that the target
// loads stay 16-byte aligned and the lookups of a 32-bit [letter][pos] profile fall into distinct banks, out-of-range cells on a padding
// letter instead of predicates.  Bit-identical, but 23.5 ms instead of 14.8 ms per 256 queries x 1M targets: a class with more than two
// diagonals per warp brings the 2-way conflicts back, every lane issues every lookup of its warp, and the regrouping costs three
// barriers per 256 diagonals.  profiles/r02_score_variants.txt.)
constexpr int PF_SPROF_LEN = 1024;   // longest query whose ungapped profile is staged in shared memory by score_kernel

__global__ void __launch_bounds__(256) score_kernel(ScoreArgs a) {
    __shared__ int8_t sprof[PF_SPROF_LEN * 22];
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    // elements are sorted by query: the block's first query almost always covers the whole block; stage its profile
    const uint32_t e0 = blockIdx.x * blockDim.x;
    const uint32_t q0 = (uint32_t) (a.elems[e0].key >> 32) >> a.tBits;
    const QueryDesc qd0 = a.queries[q0];
    const bool staged = qd0.len <= (uint32_t) PF_SPROF_LEN;
    if (staged) {
        const uint32_t words = (qd0.len * 22 + 3) / 4;
        const int8_t *src = a.profiles + qd0.profOff;
        if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
            for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) reinterpret_cast<uint32_t *>(sprof)[w] = reinterpret_cast<const uint32_t *>(src)[w];
        } else {
            for (uint32_t b = threadIdx.x; b < qd0.len * 22; b += blockDim.x) sprof[b] = src[b];
        }
    }
    __syncthreads();
    unsigned long long cells = 0;
    if (e < a.n) {
        const Elem el = a.elems[e];
        const uint32_t hiKey = (uint32_t) (el.key >> 32);
        bool keep = true;
        if (e > 0) {
            const Elem pv = a.elems[e - 1];
            // consecutive double hits of the same target on the same 8-bit diagonal collapse to the first
            // (CacheFriendlyOperations.cpp:247-267)
            if ((uint32_t) (pv.key >> 32) == hiKey && pv.pad == el.pad && (uint8_t) pv.diag == (uint8_t) el.diag) keep = false;
        }
        int sc = 0;
        if (keep) {
            const uint32_t q = hiKey >> a.tBits, t = hiKey & ((1u << a.tBits) - 1u);
            const QueryDesc qd = (q == q0) ? qd0 : a.queries[q];
            const int8_t *prof = (q == q0 && staged) ? sprof : a.profiles + qd.profOff;
            sc = ungapped_diag(prof, (int) qd.len, a.tSeq + a.tOff[t], (int) a.tLen[t], el.diag, cells);
            sc = min(sc, 255);
        }
        a.score[e] = (uint8_t) sc;
        a.kept[e] = keep ? 1 : 0;
    }
    for (int m = 16; m >= 1; m >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, m);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(a.stats + 3, cells);
}

// ---- best diagonal per (query, target): first element with the maximal score (keepMaxElement, CFO.cpp:354-385) ----
struct Cand {
    uint64_t key;                  // (query << 52) | ((255 - score) << 44) | (bin << 32) | stream position
    uint32_t target;
    uint32_t diagScore;            // diag16 | score << 16
};

struct BestArgs {
    const Elem *elems; const uint8_t *score; const uint8_t *kept; uint32_t n; uint32_t tBits;
    uint32_t binMask;              // BINSIZE - 1 of the reference's CacheFriendlyOperations on this host
    Cand *cands; uint32_t *candCount; uint32_t candCap;
    const uint32_t *qinfo; uint32_t *errFlag;
    const uint8_t *chunkOrd;       // per query [PF_MAXCHUNK + 2]: final block rank | direction << 7 of every 1-based chunk (chunk_order)
    const uint32_t *bounds;        // per query [PF_MAXCHUNK + 1]: first stream position of every 0-based chunk (chunk_bounds_kernel)
    uint32_t minScore;             // --min-ungapped-score
    uint32_t *hist;                // [queries of the sub-batch][256] number of candidates per 8-bit score (zeroed by the host)
};

struct OvElem { uint16_t diag; uint8_t count; uint8_t score; uint8_t chunk; uint32_t pos; };

// Replays, for ONE target, what QueryMatcher::match does with its double hits when the query overflowed (QueryMatcher.cpp:305-372):
// chunk 1 is kept unscored; every further overflowed chunk k: merge(keep scored) -> score -> keep max; last chunk: merge -> score ->
// keep max (matchQuery :129-181).  x[0..m) = the target's run-collapsed double hits in stream order with their final scores.
// Returns the index of the surviving best element or -1.
__device__ inline int overflow_replay(OvElem *x, int m, int nChunks) {
    OvElem A[PF_MAXOV], X[PF_MAXOV];
    int na = 0;
    for (int j = 0; j < m; j++) if (x[j].chunk == 1) { A[na] = x[j]; A[na].count = 0; na++; }
    for (int k = 2; k <= nChunks; k++) {
        int nx = 0;
        for (int j = 0; j < na; j++) X[nx++] = A[j];
        for (int j = 0; j < m; j++) if (x[j].chunk == k) { X[nx] = x[j]; X[nx].count = 0; nx++; }
        na = 0;
        if (nx == 0) continue;
        if (k < nChunks) {
            // mergeDiagonalKeepScoredHitsDuplicates: reverse pass, keeps scored elements and the last of every equal-diagonal run
            uint8_t dup = (uint8_t) ((uint8_t) X[nx - 1].diag + 1);
            for (int j = nx - 1; j >= 0; j--) {
                if (X[j].count != 0 || dup != (uint8_t) X[j].diag) A[na++] = X[j];
                dup = (uint8_t) X[j].diag;
            }
        } else {
            // mergeDiagonalDuplicates: forward pass, keeps the first of every equal-diagonal run
            uint8_t dup = (uint8_t) ((uint8_t) X[0].diag + 1);
            for (int j = 0; j < nx; j++) {
                if (dup != (uint8_t) X[j].diag) A[na++] = X[j];
                dup = (uint8_t) X[j].diag;
            }
        }
        // align: unscored elements get their diagonal score
        for (int j = 0; j < na; j++) if (A[j].count == 0) A[j].count = A[j].score;
        // keepMaxElement: first maximum, then every later element with count 0
        uint8_t mx = 0;
        for (int j = 0; j < na; j++) mx = max(mx, A[j].count);
        int nk = 0; uint8_t d = mx;
        for (int j = 0; j < na; j++) { if (d == A[j].count) { X[nk++] = A[j]; d = 0; } }
        for (int j = 0; j < nk; j++) A[j] = X[j];
        na = nk;
    }
    if (nChunks == 1) {
        for (int j = 0; j < na; j++) A[j].count = A[j].score;
        uint8_t mx = 0;
        for (int j = 0; j < na; j++) mx = max(mx, A[j].count);
        for (int j = 0; j < na; j++) if (A[j].count == mx) { x[0] = A[j]; return mx > 0 ? 0 : -1; }
        return -1;
    }
    // the first element of A with the maximal count is the survivor
    uint8_t mx = 0;
    for (int j = 0; j < na; j++) mx = max(mx, A[j].count);
    if (mx == 0) return -1;
    for (int j = 0; j < na; j++) if (A[j].count == mx) { x[0] = A[j]; return 0; }
    return -1;
}

// slow path of best_kernel: a target of an overflowed query with several double hits (chunk replay)
__device__ __noinline__ int best_overflowed(const BestArgs &a, uint32_t e, uint32_t hiKey, int nChunks, uint32_t &bestDiag, uint32_t &bestPos, uint32_t &bestChunk) {
    OvElem x[PF_MAXOV];
    int m = 0;
    for (uint32_t k = e; k < a.n && (uint32_t) (a.elems[k].key >> 32) == hiKey; k++) {
        if (!a.kept[k]) continue;
        if (m >= PF_MAXOV) { atomicExch(a.errFlag, 1u); return 0; }
        x[m].diag = (uint16_t) a.elems[k].diag; x[m].score = a.score[k]; x[m].count = 0; x[m].chunk = (uint8_t) (a.elems[k].pad + 1);
        x[m].pos = (uint32_t) a.elems[k].key; m++;
    }
    if (overflow_replay(x, m, nChunks) < 0) return 0;
    bestDiag = x[0].diag; bestPos = x[0].pos; bestChunk = x[0].chunk;
    return x[0].count;
}

__global__ void __launch_bounds__(256) best_kernel(BestArgs a) {
    // score histogram of the candidates per query, over every fourth block only: a count over a subset that reaches max-seqs at score L
    // proves that the full count reaches it at or above L, so L is a valid (lower) pruning threshold, and three quarters of the blocks
    // keep running without a block barrier (the barrier waits for the block's longest run of double hits).  The elements are sorted by
    // query: the block's first query gets a shared-memory histogram that is flushed once, the others (a block rarely spans two queries)
    // count in global memory.
    __shared__ uint32_t s_h[256];
    const bool sampled = (blockIdx.x & 3u) == 0u;
    uint32_t q0 = 0;
    if (sampled) {
        s_h[threadIdx.x] = 0;
        q0 = (uint32_t) (a.elems[blockIdx.x * blockDim.x].key >> 32) >> a.tBits;
        __syncthreads();
    }
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int best = 0; uint32_t bestDiag = 0, orderKey = 0, q = 0, t = 0;
    bool head = false;
    if (e < a.n) {
        const uint32_t hiKey = (uint32_t) (a.elems[e].key >> 32);
        head = !(e > 0 && (uint32_t) (a.elems[e - 1].key >> 32) == hiKey);
        if (head) {
            q = hiKey >> a.tBits; t = hiKey & ((1u << a.tBits) - 1u);
            const uint32_t qi = a.qinfo[q];
            const int nChunks = (int) (qi & 0x7fffffffu);
            if (qi & 0x80000000u) head = false;   // last chunk empty: the reference returns nothing for this query
            else if (nChunks <= 1) {
                uint32_t bestE = e;
                for (uint32_t k = e; k < a.n && (uint32_t) (a.elems[k].key >> 32) == hiKey; k++) {
                    if (a.kept[k] && (int) a.score[k] > best) { best = a.score[k]; bestE = k; }
                }
                bestDiag = a.elems[bestE].diag & 0xffffu; orderKey = (uint32_t) a.elems[bestE].key;
            } else if (nChunks > PF_MAXCHUNK) { atomicExch(a.errFlag, 1u); head = false; }
            else {
                // The replay only matters when it can change the outcome.  The surviving element is never better than the best kept
                // score, so targets below --min-ungapped-score are dropped at once; with distinct 8-bit diagonals no merge removes
                // anything and with a unique maximum the champion is that element whatever the (reversing) merge order is.
                uint32_t m = 0, mxK = e, mxCnt = 0, pos = 0, chunk = 0; int mx = -1; bool dupDiag = false;
                unsigned long long seen = 0;
                for (uint32_t k = e; k < a.n && (uint32_t) (a.elems[k].key >> 32) == hiKey; k++) {
                    if (!a.kept[k]) continue;
                    const int sc = a.score[k];
                    const uint32_t d8 = a.elems[k].diag & 0xffu;
                    for (uint32_t j = 0; j < min(m, 8u); j++) if (((seen >> (8 * j)) & 0xffull) == d8) dupDiag = true;
                    seen = (seen << 8) | d8;
                    if (sc > mx) { mx = sc; mxK = k; mxCnt = 1; } else if (sc == mx) mxCnt++;
                    m++;
                }
                if (mx <= 0 || (uint32_t) mx < a.minScore) best = 0;
                else if (m == 1 || (m <= 8 && mxCnt == 1 && !dupDiag)) {
                    best = mx; bestDiag = a.elems[mxK].diag & 0xffffu; pos = (uint32_t) a.elems[mxK].key; chunk = a.elems[mxK].pad + 1;
                } else best = best_overflowed(a, e, hiKey, nChunks, bestDiag, pos, chunk);
                if (best > 0) {
                    const uint8_t ord = a.chunkOrd[(size_t) q * (PF_MAXCHUNK + 2) + chunk];   // rank | direction << 7 (chunk_order)
                    // position inside the chunk: a chunk is shorter than 2·max(1e6, N) <= 2^23 hits (N <= 4M)
                    const uint32_t rel = (pos - a.bounds[(size_t) q * (PF_MAXCHUNK + 1) + chunk - 1]) & 0x7fffffu;
                    orderKey = ((uint32_t) (ord & 0x7fu) << 23) | ((ord & 0x80u) ? rel : (0x7fffffu - rel));
                }
            }
        }
    }
    // scores below --min-ungapped-score never pass the threshold max(minDiagScoreThr, .) (QueryMatcher.cpp:183-185) and the
    // threshold itself only depends on the counts at or above it: such candidates are dropped here
    const bool emit = head && best > 0 && (uint32_t) best >= a.minScore;
    const uint32_t em = __ballot_sync(0xffffffffu, emit);
    if (em != 0) {
        uint32_t slot = 0;
        if (lane == __ffs(em) - 1) slot = atomicAdd(a.candCount, (uint32_t) __popc(em));
        slot = __shfl_sync(0xffffffffu, slot, __ffs(em) - 1) + (uint32_t) __popc(em & ((1u << lane) - 1u));
        if (emit) {
            if (sampled) { if (q == q0) atomicAdd(&s_h[best], 1u); else atomicAdd(a.hist + (size_t) q * 256 + (uint32_t) best, 1u); }
            if (slot < a.candCap) {   // overflow is reported by the host (nCands > candCap)
                Cand c;
                c.key = ((uint64_t) q << 52) | ((uint64_t) (255 - best) << 44) | ((uint64_t) (t & a.binMask) << 32) | (uint64_t) orderKey;
                c.target = t;
                c.diagScore = bestDiag | ((uint32_t) best << 16);
                a.cands[slot] = c;
            }
        }
    }
    if (sampled) {
        __syncthreads();
        const uint32_t hv = s_h[threadIdx.x];
        if (hv) atomicAdd(a.hist + (size_t) q0 * 256 + threadIdx.x, hv);
    }
}

// A lower bound of computeScoreThreshold (QueryMatcher.h:211-221) per query from the (sampled) candidate histogram: the score at which
// the count from 255 downwards reaches maxHits, 0 if it never does.  Candidates below it can neither be selected nor move the
// threshold: they are left out of the candidate sort; select_kernel computes the exact threshold from the survivors, which include
// every candidate at or above it.
__global__ void cand_threshold_kernel(const uint32_t *hist, uint32_t nQueries, uint32_t maxHits, uint32_t *thr) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nQueries) return;
    const uint32_t *h = hist + (size_t) q * 256;
    uint32_t cum = 0, t = 0;
    for (int sc = 255; sc >= 1; sc--) { cum += h[sc]; if (cum >= maxHits) { t = (uint32_t) sc; break; } }
    thr[q] = t;
}

// ---- per query: threshold, cut at max_seqs in the reference's order, final (score desc, id asc) sort ----
struct SelectArgs {
    const Cand *cands;             // sorted by key
    const uint32_t *candOff;       // per query offsets into cands (exclusive scan of perQuery)
    const QueryDesc *queries;
    const int8_t *profiles;
    const uint8_t *qSeqs;          // query 3Di (self score of the rescoreHits path)
    const uint8_t *tSeq; const uint64_t *tOff; const uint32_t *tLen;
    uint32_t nQueries;
    uint32_t maxHits;              // min(max_seqs, nTargets)
    uint32_t minDiagScore;
    uint32_t queryBase;            // first query of this batch in the output arrays
    uint32_t outStride;
    fsb200_hit *hits;
    uint32_t *counts;
    const uint32_t *identity;      // per query: target that is the query itself (query DB == target DB), 0xFFFFFFFF = none; may be null
    uint8_t *scratch;              // one byte per candidate (rescoreHits: rescaled 8-bit score)
};

constexpr int SEL_MAX = 2048;

// exclusive count of `flag` over the block in thread order + block total (blockDim.x = 256)
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t *warpCnt, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();
    if (lane == 0) warpCnt[warp] = (uint32_t) __popc(bal);
    __syncthreads();
    uint32_t before = 0, tot = 0;
    for (int w = 0; w < 8; w++) { const uint32_t c = warpCnt[w]; if (w < warp) before += c; tot += c; }
    total = tot;
    return before + (uint32_t) __popc(bal & ((1u << lane) - 1u));
}

// QueryMatcher::matchQuery tail (QueryMatcher.cpp:183-240): threshold, radix order, rescoreHits when the cut falls into the
// saturated bucket (:189-199,563-586), getResult incl. the identity hit (:402-458), final sort (:233-239).
__global__ void __launch_bounds__(256) select_kernel(SelectArgs a) {
    __shared__ unsigned long long skey[SEL_MAX];   // (0xffffff - score) << 32 | target   (ascending == score desc, id asc)
    __shared__ uint32_t sdiag[SEL_MAX];
    __shared__ uint32_t s_thr0, s_lo, s_pI, s_n, s_hist[256], s_warp[8];
    __shared__ int s_cut; __shared__ uint32_t s_needEq, s_maxSelf;
    const uint32_t q = blockIdx.x;
    if (q >= a.nQueries) return;
    const uint32_t begin = a.candOff[q], end = a.candOff[q + 1];
    const uint32_t n = end - begin;
    const uint32_t ident = a.identity ? a.identity[q] : 0xFFFFFFFFu;
    const bool hasI = ident != 0xFFFFFFFFu;
    const uint32_t maxTake = a.maxHits - (hasI ? 1u : 0u);
    const QueryDesc qd = a.queries[q];
    if (threadIdx.x == 0) {
        // candidates are sorted by score descending, so computeScoreThreshold (QueryMatcher.h:211-221: the score at which the
        // cumulative count from 255 downwards reaches maxHits, 0 if it never does) is the score of the maxHits-th candidate
        uint32_t thr0 = 0;
        if (n >= a.maxHits && a.maxHits > 0) thr0 = 255u - (uint32_t) ((a.cands[begin + a.maxHits - 1].key >> 44) & 0xffu);
        const uint32_t thr = max(thr0, a.minDiagScore);
        // number of candidates with score >= thr (binary search in the descending list)
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; const uint32_t sc = 255u - (uint32_t) ((a.cands[begin + mid].key >> 44) & 0xffu); if (sc >= thr) lo = mid + 1; else hi = mid; }
        s_thr0 = thr; s_lo = lo; s_pI = 0xFFFFFFFFu; s_n = 0;
    }
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_hist[k] = 0;
    __syncthreads();
    const uint32_t thr = s_thr0, lo = s_lo;
    uint32_t take = 0;
    if (thr < 255u) {
        // cands are sorted by (score desc, bin asc, stream position asc): the first ones are the survivors (getResult, :402-458);
        // the query's own entry is skipped there and does not use up a slot of the loop
        const uint32_t m = min(lo, maxTake + 1u);
        if (hasI) for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) if (a.cands[begin + k].target == ident) atomicMin(&s_pI, k);
        __syncthreads();
        const uint32_t pI = s_pI;
        take = min(maxTake, m - (pI < m ? 1u : 0u));
        int npad = 1; while (npad < (int) take) npad <<= 1;
        for (int k = threadIdx.x; k < npad; k += blockDim.x) {
            if (k < (int) take) {
                const Cand c = a.cands[begin + (uint32_t) k + ((pI <= (uint32_t) k) ? 1u : 0u)];
                uint32_t sc = c.diagScore >> 16;
                if (sc >= 255) {   // the clamped score is replaced by the real one (QueryMatcher.cpp:441-444)
                    unsigned long long cells = 0;
                    sc = (uint32_t) ungapped_diag(a.profiles + qd.profOff, (int) qd.len, a.tSeq + a.tOff[c.target], (int) a.tLen[c.target], c.diagScore & 0xffffu, cells);
                }
                skey[k] = ((unsigned long long) (0xffffffu - sc) << 32) | c.target;
                sdiag[k] = c.diagScore & 0xffffu;
            } else { skey[k] = ~0ull; sdiag[k] = 0; }
        }
    } else {
        // ---- rescoreHits: at least max-seqs candidates saturate the 8-bit score ----
        if (threadIdx.x == 0) {
            unsigned long long cells = 0;
            int ms = ungapped_diag(a.profiles + qd.profOff, (int) qd.len, a.qSeqs + qd.seqOff, (int) qd.len, 0u, cells) - 255;
            ms = max(1, ms); ms = min(ms, 65535);
            s_maxSelf = (uint32_t) ms;
        }
        __syncthreads();
        const uint32_t maxSelf = s_maxSelf;
        const float fm = (float) (int) maxSelf;
        for (uint32_t k = threadIdx.x; k < lo; k += blockDim.x) {
            const Cand c = a.cands[begin + k];
            unsigned long long cells = 0;
            uint32_t ns = (uint32_t) ungapped_diag(a.profiles + qd.profOff, (int) qd.len, a.tSeq + a.tOff[c.target], (int) a.tLen[c.target], c.diagScore & 0xffffu, cells);
            ns -= 255u;
            const float sc = (float) min(ns, 65535u);
            const uint8_t cnt = (uint8_t) (int) ((double) __fmul_rn(__fdiv_rn(sc, fm), 255.0f) + 0.5);
            a.scratch[begin + k] = cnt;
            if (c.target != ident) atomicAdd(&s_hist[cnt], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cum = 0; int cut = -1; uint32_t needEq = 0;
            for (int c = 255; c >= 0; c--) { if (cum + s_hist[c] >= maxTake) { cut = c; needEq = maxTake - cum; break; } cum += s_hist[c]; }
            s_cut = cut; s_needEq = needEq;
        }
        __syncthreads();
        const int cut = s_cut; const uint32_t needEq = s_needEq;
        uint32_t eqBase = 0;
        for (uint32_t k0 = 0; k0 < lo; k0 += blockDim.x) {   // the counting sort is stable: of the candidates at the cut score the first ones survive
            const uint32_t k = k0 + threadIdx.x;
            bool valid = false, eq = false; Cand c; c.key = 0; c.target = 0; c.diagScore = 0; int cnt = 0;
            if (k < lo) { c = a.cands[begin + k]; cnt = a.scratch[begin + k]; valid = c.target != ident; eq = valid && cnt == cut; }
            uint32_t eqTot = 0;
            const uint32_t er = block_rank(eq, s_warp, eqTot);
            if (valid && (cnt > cut || (eq && eqBase + er < needEq))) {
                const uint32_t slot = atomicAdd(&s_n, 1u);
                if (slot < (uint32_t) SEL_MAX) {
                    const uint32_t sc = 255u + ((uint32_t) cnt * maxSelf / 255u);   // getResult, rescaleScore != 0 (:436-440)
                    skey[slot] = ((unsigned long long) (0xffffffu - sc) << 32) | c.target;
                    sdiag[slot] = c.diagScore & 0xffffu;
                }
            }
            eqBase += eqTot;
        }
        __syncthreads();
        take = min(s_n, (uint32_t) SEL_MAX);
        int npad = 1; while (npad < (int) take) npad <<= 1;
        for (int k = (int) take + (int) threadIdx.x; k < npad; k += blockDim.x) { skey[k] = ~0ull; sdiag[k] = 0; }
    }
    __syncthreads();
    int npad = 1; while (npad < (int) take) npad <<= 1;
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = threadIdx.x; k < npad; k += blockDim.x) {
                const int j = k ^ stride;
                if (j > k) {
                    const bool up = ((k & size) == 0);
                    const unsigned long long x = skey[k], y = skey[j];
                    if ((x > y) == up) { skey[k] = y; skey[j] = x; const uint32_t dx = sdiag[k]; sdiag[k] = sdiag[j]; sdiag[j] = dx; }
                }
            }
            __syncthreads();
        }
    }
    fsb200_hit *out = a.hits + (size_t) (a.queryBase + q) * a.outStride;
    const uint32_t shift = hasI ? 1u : 0u;
    if (hasI && threadIdx.x == 0 && a.maxHits > 0) { fsb200_hit hh; hh.target = ident; hh.score = 65535; hh.diagonal = 0; out[0] = hh; }   // USHRT_MAX, diagonal 0 (:409-421)
    for (int k = threadIdx.x; k < (int) take; k += blockDim.x) {
        fsb200_hit hh;
        hh.target = (uint32_t) skey[k];
        hh.score = (int32_t) (0xffffffu - (uint32_t) (skey[k] >> 32));
        hh.diagonal = (int32_t) (short) sdiag[k];
        out[k + shift] = hh;
    }
    if (threadIdx.x == 0) a.counts[a.queryBase + q] = take + ((hasI && a.maxHits > 0) ? 1u : 0u);
}

// (query, target) of every double hit.  diag_kernel appends the double hits of one (query, bucket) in stream order (one warp, its
// flushes reserve increasing ranges of the element list), and a target lives in one bucket: a STABLE sort by (query, target) leaves
// every target's hits in stream order.
__global__ void pf_extract_keys(const Elem *e, uint32_t n, uint32_t *k, uint32_t *v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = (uint32_t) (e[i].key >> 32); v[i] = i; }
}
__global__ void pf_gather_elems(const Elem *e, const uint32_t *v, uint32_t n, Elem *o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = e[v[i]];
}
// sort keys of the candidates that reach their query's threshold (compacted; the order of the survivors is restored by the sort, the
// keys are unique inside a query: they end with a stream position)
__global__ void pf_extract_cand_keys(const Cand *e, const uint32_t *candCount, uint32_t candCap, const uint32_t *thr, uint64_t *k, uint32_t *v, uint32_t *keepCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = min(*candCount, candCap);
    const int lane = threadIdx.x & 31;
    bool keep = false; uint64_t key = 0;
    if (i < n) { key = e[i].key; keep = 255u - (uint32_t) ((key >> 44) & 0xffu) >= thr[(uint32_t) (key >> 52)]; }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return;
    uint32_t slot = 0;
    if (lane == __ffs(m) - 1) slot = atomicAdd(keepCount, (uint32_t) __popc(m));
    slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1) + (uint32_t) __popc(m & ((1u << lane) - 1u));
    if (keep) { k[slot] = key; v[slot] = i; }
}
__global__ void pf_gather_cands(const Cand *e, const uint32_t *v, uint32_t n, Cand *o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = e[v[i]];
}

// first candidate of every query in the list sorted by (query, ...)
__global__ void cand_offsets_kernel(const Cand *c, uint32_t n, uint32_t nQueries, uint32_t *off) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nQueries) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint32_t) (c[mid].key >> 52) < q) lo = mid + 1; else hi = mid; }
    off[q] = lo;
}

// ---- index build ----
// key = kmer << 38 | target << 16 | position ; invalid positions get ~0
__global__ void index_keys_kernel(const uint8_t *seq, const uint64_t *tOff, const uint32_t *tLen, uint32_t nT, const int8_t *selfScore,
                                  int kmerThr, const uint64_t *posBase, uint64_t *keys) {
    const uint32_t t = blockIdx.x;
    if (t >= nT) return;
    const uint8_t *s = seq + tOff[t];
    const int L = (int) tLen[t];
    for (int pos = threadIdx.x; pos + KSPAN <= L; pos += blockDim.x) {
        const int c0 = s[pos], c1 = s[pos + 1], c2 = s[pos + 3], c3 = s[pos + 5], c4 = s[pos + 8], c5 = s[pos + 9];
        uint64_t key = ~0ull;
        if (c0 < KA && c1 < KA && c2 < KA && c3 < KA && c4 < KA && c5 < KA) {
            const int self = selfScore[c0] + selfScore[c1] + selfScore[c2] + selfScore[c3] + selfScore[c4] + selfScore[c5];
            if (!(kmerThr > 0 && self < kmerThr)) {   // IndexTable.h:147-155
                const uint64_t kmer = (uint64_t) c0 + KA * ((uint64_t) c1 + KA * ((uint64_t) c2 + KA * ((uint64_t) c3 + KA * ((uint64_t) c4 + KA * (uint64_t) c5))));
                key = (kmer << 38) | ((uint64_t) t << 16) | (uint64_t) pos;
            }
        }
        keys[posBase[t] + pos] = key;
    }
}

// one entry per distinct k-mer per target (smallest position, IndexTable.h:386-402): heads of (kmer, target) runs
__global__ void index_heads_kernel(const uint64_t *keys, uint64_t n, uint8_t *flag) {
    const uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t k = keys[e];
    bool head = (k != ~0ull) && (e == 0 || (keys[e - 1] >> 16) != (k >> 16));
    flag[e] = head ? 1 : 0;
}

__global__ void index_entries_kernel(const uint64_t *keys, const uint8_t *flag, const uint64_t *scan, uint64_t n, uint64_t *entries, uint32_t *kmerOf) {
    const uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || !flag[e]) return;
    const uint64_t k = keys[e];
    const uint64_t dst = scan[e];
    entries[dst] = ((k >> 16) & 0x3fffffull) | ((k & 0xffffull) << 32);
    kmerOf[dst] = (uint32_t) (k >> 38);
}

__global__ void index_offsets_kernel(const uint32_t *kmerOf, uint64_t nEntries, uint32_t tableSize, uint32_t *offsets) {
    const uint64_t e = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (e > nEntries) return;
    const uint32_t cur = (e < nEntries) ? kmerOf[e] : tableSize;
    const uint32_t prev = (e == 0) ? 0xFFFFFFFFu : kmerOf[e - 1];
    // offsets[k] = e for every k in (prev, cur]
    uint32_t k = (e == 0) ? 0 : prev + 1;
    for (; k <= cur; k++) offsets[k] = (uint32_t) e;
}

__global__ void klist_kernel(const uint32_t *offsets, const uint64_t *entries, uint32_t tableSize, const uint16_t *colPerm, uint2 *klist, uint32_t *kbits,
                             unsigned long long *nonEmpty) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t len0 = (k < tableSize) ? offsets[k + 1] - offsets[k] : 0u;
    const uint32_t bal = __ballot_sync(0xffffffffu, len0 != 0);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(nonEmpty, (unsigned long long) __popc(bal));
    if (k >= tableSize) return;
    const uint32_t o0 = offsets[k], len = len0;
    // list heads are stored [first 3-mer][second 3-mer] like the bitmap: the similar k-mers of one row of a query window share the first
    // 3-mer, so their heads fall into one 64 KB stretch and share DRAM bursts (k itself is first + 8000 * second, the reference's index);
    // inside the row the columns follow Tables::colPerm
    const uint32_t kl = (k % (uint32_t) N3) * (uint32_t) N3 + (uint32_t) colPerm[k / (uint32_t) N3];
    if (len) atomicOr(kbits + (kl >> 5), 1u << (kl & 31u));
    if (len == 1) {
        const uint64_t en = entries[o0];
        klist[kl] = make_uint2(0x80000000u | ((uint32_t) en & 0x3fffffu), 1u | ((uint32_t) (en >> 32) << 16));
    } else klist[kl] = make_uint2(o0, len);
}

// ---------------------------------------------------------------------------------------------------------
// stage timing with CUDA events on the launching stream (read once, after the final synchronise)
struct StageTimer {
    struct Rec { int stage; cudaEvent_t e0, e1; };
    std::vector<cudaEvent_t> pool;
    std::vector<Rec> recs;
    size_t used = 0;
    cudaEvent_t get() { if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); } return pool[used++]; }
    void begin(int stage, cudaStream_t s) { Rec r; r.stage = stage; r.e0 = get(); r.e1 = nullptr; cudaEventRecord(r.e0, s); recs.push_back(r); }
    void end(cudaStream_t s) { recs.back().e1 = get(); cudaEventRecord(recs.back().e1, s); }
    void collect(uint64_t *ns, int nStages) {   // caller has synchronised the stream
        for (const Rec &r : recs) { float ms = 0; if (r.e1 && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.stage < nStages) ns[r.stage] += (uint64_t) (ms * 1e6); }
        recs.clear(); used = 0;
    }
    void release() { for (cudaEvent_t e : pool) cudaEventDestroy(e); pool.clear(); recs.clear(); used = 0; }
};

enum { ST_COUNT = 0, ST_LISTS = 1, ST_EXPAND = 2, ST_DIAG = 3, ST_SORT = 4, ST_SCORE = 5, ST_SELECT = 6, ST_N = 8 };

struct State {
    uint64_t stats[8] = {0};
    uint64_t stageNs[ST_N] = {0};   // device time per pipeline stage of the last run (ST_*)
    StageTimer timer;
    bool tablesReady = false, indexReady = false;
    bool borrowedIndex = false;     // tables + index belong to another context on this GPU (fsb200_create_lane): never freed here
    int share = 1;                  // contexts that work on this GPU side by side: scratch budgets are divided by it
    uint64_t residentBytes = 0;     // device memory in use when the index was built (DB + index + tables + CUDA context)
    bool useBits = true;            // probe the occupancy bitmap before the list heads (sparse indexes only)
    uint64_t nonEmptyKmers = 0;
    Tables tb;
    uint32_t *d_offsets = nullptr;
    uint2 *d_klist = nullptr;
    uint32_t *d_kbits = nullptr;
    uint64_t *d_entries = nullptr;
    uint8_t *d_masked = nullptr;
    uint64_t nEntries = 0;
    int kmerThr = 0;
    struct Buf { void *p = nullptr; size_t bytes = 0; };
    Buf bufs[48];
    size_t scratchBytes() const { size_t t = 0; for (const Buf &b : bufs) t += b.bytes; return t; }
    // grow-only device scratch that survives across calls (cudaMalloc/cudaFree of GBs per call costs more than the kernels)
    template <typename T> cudaError_t ensure(int slot, T *&ptr, size_t count) {
        const size_t need = std::max<size_t>(16, count * sizeof(T));
        if (bufs[slot].bytes < need) {
            if (bufs[slot].p) cudaFree(bufs[slot].p);
            bufs[slot].p = nullptr; bufs[slot].bytes = 0;
            cudaError_t e = cudaMalloc(&bufs[slot].p, need + need / 8);
            if (e != cudaSuccess) return e;
            bufs[slot].bytes = need + need / 8;
        }
        ptr = reinterpret_cast<T *>(bufs[slot].p);
        return cudaSuccess;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    void drop_borrowed() {
        if (!borrowedIndex) return;
        tb = Tables(); d_offsets = nullptr; d_klist = nullptr; d_kbits = nullptr; d_entries = nullptr; d_masked = nullptr;
        tablesReady = indexReady = borrowedIndex = false;
    }
    void borrow_index(const State &o) {
        tb = o.tb; d_offsets = o.d_offsets; d_klist = o.d_klist; d_kbits = o.d_kbits; d_entries = o.d_entries; d_masked = o.d_masked;
        nEntries = o.nEntries; kmerThr = o.kmerThr; useBits = o.useBits; nonEmptyKmers = o.nonEmptyKmers;
        tablesReady = o.tablesReady; indexReady = o.indexReady; borrowedIndex = true; residentBytes = o.residentBytes;
    }
    void invalidate() { if (borrowedIndex) drop_borrowed(); indexReady = false; }
    void release() {
        drop_borrowed();
        cudaFree(tb.sIdx); cudaFree(tb.sScore); cudaFree(tb.cntGe); cudaFree(tb.colPerm); cudaFree(d_offsets); cudaFree(d_klist); cudaFree(d_kbits); cudaFree(d_entries); cudaFree(d_masked);
        tb = Tables(); d_offsets = nullptr; d_klist = nullptr; d_kbits = nullptr; d_entries = nullptr; d_masked = nullptr; tablesReady = indexReady = false;
        for (Buf &b : bufs) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
        if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); ev0 = ev1 = nullptr; }
        timer.release();
    }
};

#define PF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); return FSB200_E_CUDA; } \
    } while (0)

// ExtendedSubstitutionMatrix::calcScoreMatrix for 3-mers over the 20-letter alphabet (M/ExtendedSubstitutionMatrix.cpp:20-72)
inline int build_tables(State &st, const host::ScoreMatrix &mKmer, cudaStream_t stream, std::string &err) {
    std::vector<uint16_t> idx((size_t) N3 * N3);
    std::vector<int16_t> sc((size_t) N3 * N3);
    int vmin = 32767, vmax = -32768;
    for (int a = 0; a < KA; a++) for (int b = 0; b < KA; b++) { vmin = std::min(vmin, (int) mKmer.s[a][b]); vmax = std::max(vmax, (int) mKmer.s[a][b]); }
    vmin *= 3; vmax *= 3;
    const int NS = vmax - vmin + 2;
    std::vector<uint16_t> cge((size_t) N3 * NS);
#pragma omp parallel
    {
        std::vector<int> cnt(NS), start(NS);
#pragma omp for schedule(dynamic, 16)
        for (int row = 0; row < N3; row++) {
            const int i0 = row % KA, i1 = (row / KA) % KA, i2 = row / (KA * KA);
            // stable sort by score descending of the entries visited in permutation order (first letter outermost)
            // == counting sort with buckets filled in visiting order
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int j0 = 0; j0 < KA; j0++) for (int j1 = 0; j1 < KA; j1++) for (int j2 = 0; j2 < KA; j2++)
                cnt[mKmer.s[i0][j0] + mKmer.s[i1][j1] + mKmer.s[i2][j2] - vmin]++;
            int run = 0;
            for (int v = NS - 1; v >= 0; v--) { start[v] = run; run += cnt[v]; }
            for (int v = 0; v < NS; v++) cge[(size_t) row * NS + v] = (uint16_t) std::min(65535, start[v] + cnt[v]);   // #entries with score >= vmin+v (8000 max)
            for (int j0 = 0; j0 < KA; j0++) for (int j1 = 0; j1 < KA; j1++) for (int j2 = 0; j2 < KA; j2++) {
                const int s = mKmer.s[i0][j0] + mKmer.s[i1][j1] + mKmer.s[i2][j2];
                const int dst = start[s - vmin]++;
                idx[(size_t) row * N3 + dst] = (uint16_t) (j0 + j1 * KA + j2 * KA * KA);
                sc[(size_t) row * N3 + dst] = (int16_t) s;
            }
        }
    }
    // Column order of the list-head table: the similar k-mers of a window row differ from the query's second 3-mer mostly by one similar
    // letter.  Letters are chained greedily by substitution score (each next letter is the unplaced one most similar to the last), and a
    // 3-mer's column is its id in the renamed letters: the variants in the fastest letter fall into the same 32-byte sector (4 heads).
    std::vector<uint16_t> colPerm(N3);
    {
        int order[KA], rank[KA]; bool used[KA] = {false};
        int cur = 0, best = 32767;
        for (int a = 0; a < KA; a++) { int tot = 0; for (int b = 0; b < KA; b++) tot += mKmer.s[a][b]; if (tot < best) { best = tot; cur = a; } }
        for (int i = 0; i < KA; i++) {
            order[i] = cur; used[cur] = true;
            int nxt = -1;
            for (int b = 0; b < KA; b++) if (!used[b] && (nxt < 0 || mKmer.s[cur][b] > mKmer.s[cur][nxt])) nxt = b;
            cur = nxt;
        }
        for (int i = 0; i < KA; i++) rank[order[i]] = i;
        for (int b = 0; b < N3; b++) colPerm[b] = (uint16_t) (rank[b % KA] + KA * rank[(b / KA) % KA] + KA * KA * rank[b / (KA * KA)]);
    }
    PF_CUDA(cudaMalloc(&st.tb.colPerm, N3 * 2));
    PF_CUDA(cudaMemcpyAsync(st.tb.colPerm, colPerm.data(), N3 * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMalloc(&st.tb.sIdx, idx.size() * 2));
    PF_CUDA(cudaMalloc(&st.tb.sScore, sc.size() * 2));
    PF_CUDA(cudaMalloc(&st.tb.cntGe, cge.size() * 2));
    PF_CUDA(cudaMemcpyAsync(st.tb.sIdx, idx.data(), idx.size() * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(st.tb.sScore, sc.data(), sc.size() * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(st.tb.cntGe, cge.data(), cge.size() * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaStreamSynchronize(stream));
    st.tb.vmin = vmin; st.tb.NS = NS;
    st.tablesReady = true;
    return FSB200_OK;
}

// Masker::maskRepeats + finalizeMasking (M/../commons/Masker.cpp:85-127); Foldseek runs with tantan off, repeats > 6 masked
inline void mask_repeats(uint8_t *s, uint32_t len, int n) {
    uint32_t count = 0; int start = -1; int prev = 0;
    for (uint32_t pos = 0; pos < len; ++pos) {
        const int cur = s[pos];
        if (cur == prev) { count++; }
        else {
            if (count > (uint32_t) n) for (uint32_t k = (uint32_t) start; k < pos; ++k) s[k] = 20;
            count = 1; start = (int) pos; prev = cur;
        }
    }
    if (count > (uint32_t) n) for (uint32_t k = (uint32_t) start; k < len; ++k) s[k] = 20;
}

inline int build_index(State &st, const fsb200_config &cfg, const host::ScoreMatrix &mKmer, cudaStream_t stream,
                       const uint8_t *d_t3Di, size_t t3DiBytes, const uint64_t *d_tOff, const uint32_t *d_tLen,
                       const std::vector<uint64_t> &h_tOff, const std::vector<uint32_t> &h_tLen, uint32_t nT, std::string &err) {
    if (nT >= (1u << 22)) { err = "more than 4M targets need target splits (not implemented)"; return FSB200_E_RANGE; }
    // Prefiltering::getKmerThreshold, k = 6 sequence search (Prefiltering.cpp:1084-1086)
    if (cfg.kmer_size != 6) { err = "only -k 6 is implemented"; return FSB200_E_RANGE; }
    float base = 163.2;
    st.kmerThr = (int) (base - (cfg.sensitivity * 8.917));
    // masked copy of the target 3Di sequences = the reference's SequenceLookup
    std::vector<uint8_t> masked(t3DiBytes);
    PF_CUDA(cudaMemcpyAsync(masked.data(), d_t3Di, t3DiBytes, cudaMemcpyDeviceToHost, stream));
    PF_CUDA(cudaStreamSynchronize(stream));
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t t = 0; t < nT; t++) mask_repeats(masked.data() + h_tOff[t], h_tLen[t], 6);
    cudaFree(st.d_masked); st.d_masked = nullptr;
    PF_CUDA(cudaMalloc(&st.d_masked, t3DiBytes));
    PF_CUDA(cudaMemcpyAsync(st.d_masked, masked.data(), t3DiBytes, cudaMemcpyHostToDevice, stream));
    std::vector<uint64_t> posBase(nT + 1);
    uint64_t total = 0;
    for (uint32_t t = 0; t < nT; t++) { posBase[t] = total; total += (h_tLen[t] >= KSPAN) ? (h_tLen[t] - KSPAN + 1) : 0; }
    posBase[nT] = total;
    if (total == 0) { err = "no k-mer could be extracted from the target database"; return FSB200_E_RANGE; }
    int8_t self[32];
    for (int a = 0; a < 21; a++) self[a] = (int8_t) mKmer.s[a][a];   // getScoreLookup truncates to char (IndexBuilder.cpp:11-23)
    uint64_t *d_posBase = nullptr, *d_keys = nullptr, *d_keys2 = nullptr, *d_scan = nullptr;
    int8_t *d_self = nullptr; uint8_t *d_flag = nullptr; uint32_t *d_kmerOf = nullptr; void *d_tmp = nullptr;
    PF_CUDA(cudaMalloc(&d_posBase, (nT + 1) * 8)); PF_CUDA(cudaMalloc(&d_keys, total * 8)); PF_CUDA(cudaMalloc(&d_keys2, total * 8));
    PF_CUDA(cudaMalloc(&d_self, 32)); PF_CUDA(cudaMalloc(&d_flag, total)); PF_CUDA(cudaMalloc(&d_scan, total * 8));
    PF_CUDA(cudaMemcpyAsync(d_posBase, posBase.data(), (nT + 1) * 8, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_self, self, 32, cudaMemcpyHostToDevice, stream));
    index_keys_kernel<<<nT, 128, 0, stream>>>(st.d_masked, d_tOff, d_tLen, nT, d_self, st.kmerThr, d_posBase, d_keys);
    PF_CUDA(cudaGetLastError());
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, d_keys, d_keys2, total, 0, 64, stream);
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, (const uint8_t *) nullptr, (uint64_t *) nullptr, total, stream);
    tmpBytes = std::max(tmpBytes, tmp2);
    PF_CUDA(cudaMalloc(&d_tmp, tmpBytes));
    PF_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmpBytes, d_keys, d_keys2, total, 0, 64, stream));
    const unsigned blocks = (unsigned) ((total + 255) / 256);
    index_heads_kernel<<<blocks, 256, 0, stream>>>(d_keys2, total, d_flag);
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmpBytes, d_flag, d_scan, total, stream));
    uint64_t lastScan = 0; uint8_t lastFlag = 0;
    PF_CUDA(cudaMemcpyAsync(&lastScan, d_scan + total - 1, 8, cudaMemcpyDeviceToHost, stream));
    PF_CUDA(cudaMemcpyAsync(&lastFlag, d_flag + total - 1, 1, cudaMemcpyDeviceToHost, stream));
    PF_CUDA(cudaStreamSynchronize(stream));
    st.nEntries = lastScan + lastFlag;
    cudaFree(st.d_entries); cudaFree(st.d_offsets); st.d_entries = nullptr; st.d_offsets = nullptr;
    const uint32_t tableSize = 64000000u;
    PF_CUDA(cudaMalloc(&st.d_entries, std::max<uint64_t>(1, st.nEntries) * 8));
    PF_CUDA(cudaMalloc(&st.d_offsets, ((size_t) tableSize + 1) * 4));
    PF_CUDA(cudaMalloc(&d_kmerOf, std::max<uint64_t>(1, st.nEntries) * 4));
    index_entries_kernel<<<blocks, 256, 0, stream>>>(d_keys2, d_flag, d_scan, total, st.d_entries, d_kmerOf);
    PF_CUDA(cudaGetLastError());
    if (st.nEntries >= 0xFFFFFFFFull) { err = "index has more than 2^32 entries (target split needed)"; return FSB200_E_RANGE; }
    index_offsets_kernel<<<(unsigned) ((st.nEntries + 1 + 255) / 256), 256, 0, stream>>>(d_kmerOf, st.nEntries, tableSize, st.d_offsets);
    PF_CUDA(cudaGetLastError());
    cudaFree(st.d_klist); st.d_klist = nullptr; cudaFree(st.d_kbits); st.d_kbits = nullptr;
    PF_CUDA(cudaMalloc(&st.d_klist, (size_t) tableSize * 8));
    PF_CUDA(cudaMalloc(&st.d_kbits, (size_t) tableSize / 8 + 64));
    PF_CUDA(cudaMemsetAsync(st.d_kbits, 0, (size_t) tableSize / 8 + 64, stream));
    if (st.nEntries >= 0x7fffffffull) { err = "index has more than 2^31 entries (target split needed)"; return FSB200_E_RANGE; }
    unsigned long long *d_ne = nullptr;
    PF_CUDA(cudaMalloc(&d_ne, 8)); PF_CUDA(cudaMemsetAsync(d_ne, 0, 8, stream));
    klist_kernel<<<(tableSize + 255) / 256, 256, 0, stream>>>(st.d_offsets, st.d_entries, tableSize, st.tb.colPerm, st.d_klist, st.d_kbits, d_ne);
    { unsigned long long ne = 0; PF_CUDA(cudaMemcpyAsync(&ne, d_ne, 8, cudaMemcpyDeviceToHost, stream)); PF_CUDA(cudaStreamSynchronize(stream)); cudaFree(d_ne);
      st.nonEmptyKmers = ne;
      // with most k-mers present (large DBs) the bitmap probe only adds a dependent load in front of the list head
      st.useBits = ne < (unsigned long long) tableSize * 7 / 10;
      if (getenv("FSB200_PF_BITS")) st.useBits = atoi(getenv("FSB200_PF_BITS")) != 0; }
    {   // keep the 8 MB occupancy bitmap resident in L2 (persisting access window); best effort
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t) 16 << 20);
        cudaStreamAttrValue av; memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = st.d_kbits; av.accessPolicyWindow.num_bytes = (size_t) tableSize / 8;
        av.accessPolicyWindow.hitRatio = 1.0f; av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_posBase); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_self); cudaFree(d_flag); cudaFree(d_scan); cudaFree(d_kmerOf); cudaFree(d_tmp);
    st.indexReady = true;
    { size_t fb = 0, tb2 = 0; cudaMemGetInfo(&fb, &tb2); st.residentBytes = (tb2 - fb) > st.scratchBytes() ? (tb2 - fb) - st.scratchBytes() : 0; }
    return FSB200_OK;
}

// BINSIZE of the reference's CacheFriendlyOperations for this DB size on this host (QueryMatcher.cpp:460-487)
inline uint32_t reference_bin_size(size_t dbSize, int override_) {
    if (override_ > 0) return (uint32_t) override_;
    uint64_t l2 = 262144;
#if defined(_SC_LEVEL2_CACHE_SIZE)
    long v = sysconf(_SC_LEVEL2_CACHE_SIZE);
    if (v > 0) l2 = (uint64_t) v;
#endif
    for (uint32_t b = 2; b <= 1024; b <<= 1) if (dbSize / b < l2) return b;
    return 2048;
}

template <typename QF>
inline int run(State &st, const fsb200_config &cfg, const host::ScoreMatrix &mKmer, const host::ScoreMatrix &mUngapped,
               cudaStream_t stream, int smCount, const uint8_t *d_t3Di, size_t t3DiBytes, const uint64_t *d_tOff, const uint32_t *d_tLen,
               const std::vector<uint64_t> &h_tOff, const std::vector<uint32_t> &h_tLen, uint32_t nT,
               const std::vector<uint8_t> &h_p3Di, size_t nQ, QF queryAt, const uint32_t *identity, fsb200_hit *hits, uint32_t *counts, std::string &err) {
    memset(st.stats, 0, sizeof(st.stats));
    memset(st.stageNs, 0, sizeof(st.stageNs));
    const bool tmg = getenv("FSB200_TIMING") != NULL;
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto since = [&]() { struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); return 1e3 * (t1.tv_sec - ts0.tv_sec) + 1e-6 * (t1.tv_nsec - ts0.tv_nsec); };
    int rc;
    if (!st.tablesReady && (rc = build_tables(st, mKmer, stream, err)) != FSB200_OK) return rc;
    if (!st.indexReady && (rc = build_index(st, cfg, mKmer, stream, d_t3Di, t3DiBytes, d_tOff, d_tLen, h_tOff, h_tLen, nT, err)) != FSB200_OK) return rc;
    uint32_t tBits = 1; while ((1u << tBits) < nT) tBits++;
    const uint32_t qBits = std::min<uint32_t>(12, 32 - tBits);
    const size_t batchMax = (size_t) 1 << qBits;
    const uint32_t maxHits = std::min<uint32_t>((uint32_t) cfg.max_seqs, nT);
    if (maxHits > SEL_MAX) { err = "--max-seqs above 2048 is not implemented"; return FSB200_E_RANGE; }
    const uint32_t binSize = reference_bin_size(nT, cfg.diag_bin_count);
    const uint64_t maxDbMatches = std::max<uint64_t>(1000000, nT) * 2;
    auto envU = [](const char *name, uint64_t def) { const char *v = getenv(name); return (v && *v) ? (uint64_t) strtoull(v, NULL, 10) : def; };

    // ---- per query host preparation: thresholds per window and the ungapped profile ----
    std::vector<QueryDesc> qd(nQ);
    uint64_t seqTotal = 0, thrTotal = 0, profTotal = 0;
    std::vector<uint64_t> srcOff(nQ);
    for (size_t i = 0; i < nQ; i++) {
        uint64_t off; uint32_t len; queryAt(i, off, len);
        srcOff[i] = off;
        qd[i].seqOff = seqTotal; qd[i].thrOff = thrTotal; qd[i].profOff = profTotal; qd[i].len = len; qd[i].pad = 0;
        seqTotal += len; thrTotal += len; profTotal += (uint64_t) len * 22;
    }
    std::vector<uint8_t> seqs(seqTotal); std::vector<int16_t> thr(thrTotal); std::vector<int8_t> prof(profTotal);
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < nQ; i++) {
        const int L = (int) qd[i].len;
        const uint8_t *s = h_p3Di.data() + srcOff[i];
        memcpy(seqs.data() + qd[i].seqOff, s, L);
        std::vector<float> bias(L, 0.0f);
        if (cfg.comp_bias_corr) host::composition_bias(mKmer, s, L, cfg.comp_bias_scale_prefilter, bias.data());   // QueryMatcher.cpp:109-112
        int16_t *th = thr.data() + qd[i].thrOff;
        for (int w = 0; w < L; w++) th[w] = -1;
        for (int w = 0; w + KSPAN <= L; w++) {
            static const int P[6] = {0, 1, 3, 5, 8, 9};
            float bc = 0; bool hasX = false;
            for (int k = 0; k < 6; k++) { bc += bias[w + P[k]]; if (s[w + P[k]] >= KA) hasX = true; }
            if (hasX) continue;
            const short b = (short) ((bc < 0.0) ? bc - 0.5 : bc + 0.5);
            th[w] = (int16_t) std::max(st.kmerThr - b, 0);   // QueryMatcher.cpp:269-274
        }
        int8_t *p = prof.data() + qd[i].profOff;   // UngappedAlignment::createProfile (UngappedAlignment.cpp:388-421)
        for (int pos = 0; pos < L; pos++) {
            float b = bias[pos];
            b = (b < 0.0) ? b / 4 - 0.5 : b / 4 + 0.5;
            const int8_t corr = (int8_t) b;
            for (int a = 0; a < 22; a++) p[(size_t) pos * 22 + a] = (a < 21) ? (int8_t) (mUngapped.s[s[pos]][a] + corr) : 0;
        }
    }
    uint8_t *d_seqs = nullptr; int16_t *d_thr = nullptr; int8_t *d_prof = nullptr; QueryDesc *d_qd = nullptr; uint32_t *d_ident = nullptr;
    PF_CUDA(st.ensure(0, d_seqs, seqTotal + 16)); PF_CUDA(st.ensure(1, d_thr, thrTotal + 16)); PF_CUDA(st.ensure(2, d_prof, profTotal + 16));
    PF_CUDA(st.ensure(3, d_qd, nQ));
    PF_CUDA(cudaMemcpyAsync(d_seqs, seqs.data(), seqTotal, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_thr, thr.data(), thrTotal * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_prof, prof.data(), profTotal, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_qd, qd.data(), nQ * sizeof(QueryDesc), cudaMemcpyHostToDevice, stream));
    if (identity) {
        PF_CUDA(st.ensure(4, d_ident, std::max<size_t>(1, nQ)));
        PF_CUDA(cudaMemcpyAsync(d_ident, identity, nQ * 4, cudaMemcpyHostToDevice, stream));
    }
    if (tmg) fprintf(stderr, "prefilter: host prep+upload issued at %.1f ms\n", since());

    // ---- geometry of the matching pipeline (match.cuh) ----
    const uint32_t S = (uint32_t) std::min<uint64_t>(MT_SMAX, std::max<uint64_t>(32, envU("FSB200_PF_S", MT_SMAX)));
    uint32_t NB = 16;
    while (NB < (uint32_t) MT_MAXNB && (nT + NB - 1) / NB > 8192) NB <<= 1;
    { const uint64_t v = envU("FSB200_PF_NB", 0); if (v >= 16 && v <= (uint64_t) MT_MAXNB && (v & (v - 1)) == 0) NB = (uint32_t) v; }
    uint32_t nbShift = 0; while ((1u << nbShift) < NB) nbShift++;
    const uint32_t stateStride = (((nT + NB - 1) / NB) + 15u) & ~15u;
    const size_t diagSmem = (size_t) MT_CWARPS * ((size_t) stateStride + MT_STAGE * 16 + MT_MAP);
    if ((nT + NB - 1) / NB > (1u << MT_LBITS) || nT > (1u << 22)) { err = "more than 4M targets are not supported"; return FSB200_E_RANGE; }
    if (diagSmem > 227 * 1024 - 1024) { err = "per-target state does not fit in shared memory (more than 4M targets?)"; return FSB200_E_RANGE; }
    PF_CUDA(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) diagSmem));
    const size_t outTileBytes = (size_t) MT_R * 8 + (size_t) (NB + MT_TROW) * 2;
    size_t freeB = 0, totalB = 0;
    cudaMemGetInfo(&freeB, &totalB);
    // budgets: slot array (8 B per similar k-mer of a batch) and out-tile pool (8 B per index hit of a sub-batch)
    // The scratch of one context: its part (lanes: `share` contexts work on this GPU side by side) of 85 % of the device memory minus what
    // was resident when the index was built (DB, index, tables).  Derived from totals so that it does not move from call to call while
    // other lanes allocate; what is free right now only caps it.  Half of it for the pool, a quarter each for slots and double hits.
    const uint64_t share = (uint64_t) std::max(1, st.share);
    const uint64_t fromTotal = ((uint64_t) ((double) totalB * 0.85) > st.residentBytes ? (uint64_t) ((double) totalB * 0.85) - st.residentBytes : 0) / share;
    const uint64_t fromFree = freeB + st.scratchBytes() > ((uint64_t) 4 << 30) ? freeB + st.scratchBytes() - ((uint64_t) 4 << 30) : 0;
    const uint64_t scratchBudget = std::max<uint64_t>((uint64_t) 1 << 30, std::min(fromTotal, fromFree));
    const uint64_t slotBudget = std::min<uint64_t>(envU("FSB200_PF_SLOT_GB", 24) << 30, std::max<uint64_t>((uint64_t) 256 << 20, scratchBudget / 4));
    const uint64_t poolBudget = std::min<uint64_t>(envU("FSB200_PF_POOL_GB", 32) << 30, std::max<uint64_t>((uint64_t) 256 << 20, scratchBudget / 2));
    // double-hit buffers (~90 B per element over all arrays): sized per batch from its index hits (random data: ~3 % of the hits
    // repeat their diagonal), grown when a single query needs more
    const uint64_t elemCapMax = std::min<uint64_t>(envU("FSB200_PF_ELEMCAP", 400000000ull), std::max<uint64_t>(1u << 20, scratchBudget / 4 / 96));
    uint32_t elemCap = 0;
    if (tmg) fprintf(stderr, "prefilter: S %u, NB %u, state %u B/warp, diag smem %zu B, slot budget %.1f GB, pool budget %.1f GB, bitmap %d\n",
                     S, NB, stateStride, diagSmem, slotBudget / 1073741824.0, poolBudget / 1073741824.0, st.useBits ? 1 : 0);

    uint32_t *d_ctr = nullptr; unsigned long long *d_stats = nullptr;
    PF_CUDA(st.ensure(5, d_ctr, 16)); PF_CUDA(st.ensure(6, d_stats, 16));
    PF_CUDA(cudaMemsetAsync(d_stats, 0, 128, stream));
    Elem *d_elems = nullptr, *d_elems2 = nullptr; uint8_t *d_score = nullptr, *d_kept = nullptr; Cand *d_cands = nullptr, *d_cands2 = nullptr;
    uint64_t *d_k1 = nullptr, *d_k2 = nullptr; uint32_t *d_v1 = nullptr, *d_v2 = nullptr; uint8_t *d_tmp = nullptr; size_t tmpBytes = 0;
    uint32_t *d_candOff = nullptr;
    auto alloc_elems = [&](uint64_t want) -> cudaError_t {
        const uint32_t cap = (uint32_t) std::min<uint64_t>(elemCapMax, std::max<uint64_t>(1u << 20, want));
        if (cap <= elemCap) return cudaSuccess;
        cudaError_t e;
        if ((e = st.ensure(7, d_elems, cap)) != cudaSuccess || (e = st.ensure(8, d_elems2, cap)) != cudaSuccess ||
            (e = st.ensure(9, d_score, cap)) != cudaSuccess || (e = st.ensure(10, d_kept, cap)) != cudaSuccess ||
            (e = st.ensure(11, d_k1, cap)) != cudaSuccess || (e = st.ensure(12, d_k2, cap)) != cudaSuccess ||
            (e = st.ensure(13, d_v1, cap)) != cudaSuccess || (e = st.ensure(14, d_v2, cap)) != cudaSuccess ||
            (e = st.ensure(21, d_cands, cap)) != cudaSuccess || (e = st.ensure(22, d_cands2, cap)) != cudaSuccess) return e;
        elemCap = cap;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, d_k1, d_k2, d_v1, d_v2, (int) elemCap, 0, 64, stream);
        if (need > tmpBytes) tmpBytes = need;
        return st.ensure(15, d_tmp, tmpBytes);
    };
    PF_CUDA(st.ensure(18, d_candOff, batchMax + 1));
    uint32_t *d_hist = nullptr, *d_thrQ = nullptr;   // per query of a sub-batch: candidate score histogram, selection threshold
    PF_CUDA(st.ensure(38, d_hist, batchMax * 256)); PF_CUDA(st.ensure(39, d_thrQ, batchMax));
    fsb200_hit *d_hits = nullptr; uint32_t *d_counts = nullptr;
    PF_CUDA(st.ensure(19, d_hits, std::max<size_t>(1, nQ) * cfg.max_seqs)); PF_CUDA(st.ensure(20, d_counts, std::max<size_t>(1, nQ)));
    PF_CUDA(cudaMemsetAsync(d_counts, 0, std::max<size_t>(1, nQ) * 4, stream));
    int result = FSB200_OK;
    StageTimer &tmr = st.timer;

    // ---- batches: slots of all queries of a batch are produced at once; expansion / detection / selection run on sub-batches ----
    size_t curBatch = batchMax, subMax = batchMax;
    std::vector<uint32_t> h_qWin, h_qTile, h_qOut; std::vector<unsigned long long> h_qPos;
    for (size_t qb = 0; qb < nQ && result == FSB200_OK;) {
        uint32_t nb = (uint32_t) std::min(curBatch, nQ - qb);
        const uint64_t winBase = qd[qb].thrOff;
        h_qWin.resize(nb + 1);
        for (uint32_t i = 0; i <= nb; i++) h_qWin[i] = (uint32_t) ((qb + i < nQ ? qd[qb + i].thrOff : thrTotal) - winBase);
        const uint32_t nWin = h_qWin[nb];
        if (nWin == 0) { qb += nb; continue; }
        uint32_t *d_qWin = nullptr; unsigned long long *d_winPacked = nullptr, *d_winPre = nullptr;
        PF_CUDA(st.ensure(23, d_qWin, nb + 1)); PF_CUDA(st.ensure(24, d_winPacked, (size_t) nWin + 1)); PF_CUDA(st.ensure(25, d_winPre, (size_t) nWin + 1));
        PF_CUDA(cudaMemcpyAsync(d_qWin, h_qWin.data(), (nb + 1) * 4, cudaMemcpyHostToDevice, stream));
        tmr.begin(ST_COUNT, stream);
        win_count_kernel<<<(nWin + 1 + 255) / 256, 256, 0, stream>>>(st.tb, d_seqs + winBase, d_thr + winBase, nWin, S, d_winPacked);
        PF_CUDA(cudaGetLastError());
        { size_t need = 0; cub::DeviceScan::ExclusiveSum(nullptr, need, d_winPacked, d_winPre, (int) (nWin + 1), stream);
          if (need > tmpBytes) tmpBytes = need; }
        PF_CUDA(st.ensure(15, d_tmp, tmpBytes));
        PF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmpBytes, d_winPacked, d_winPre, (int) (nWin + 1), stream));
        tmr.end(stream);
        unsigned long long winTot = 0;
        PF_CUDA(cudaMemcpyAsync(&winTot, d_winPre + nWin, 8, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        const uint64_t totalK = winTot & MT_LOW40; const uint32_t nTiles = (uint32_t) (winTot >> 40);
        if (totalK * 8 > slotBudget && nb > 1) {   // too many similar k-mers for the slot array: smaller batch
            curBatch = std::max<size_t>(1, (size_t) ((double) nb * (double) slotBudget / (double) (totalK * 8) * 0.95));
            if (curBatch >= nb) curBatch = nb - 1;
            continue;
        }
        if (nTiles == 0) { qb += nb; continue; }
        uint2 *d_slots = nullptr; TileRec *d_tiles = nullptr; unsigned long long *d_tilePacked = nullptr, *d_tilePre = nullptr;
        uint32_t *d_qTile = nullptr, *d_qOut = nullptr, *d_qinfo = nullptr, *d_bounds = nullptr; unsigned long long *d_qPos = nullptr;
        PF_CUDA(st.ensure(26, d_slots, std::max<uint64_t>(1, totalK))); PF_CUDA(st.ensure(27, d_tiles, nTiles));
        PF_CUDA(st.ensure(28, d_tilePacked, (size_t) nTiles + 1)); PF_CUDA(st.ensure(29, d_tilePre, (size_t) nTiles + 1));
        PF_CUDA(st.ensure(30, d_qTile, nb + 1)); PF_CUDA(st.ensure(31, d_qOut, nb + 1)); PF_CUDA(st.ensure(32, d_qPos, nb + 1));
        PF_CUDA(st.ensure(33, d_qinfo, nb + 1)); PF_CUDA(st.ensure(34, d_bounds, (size_t) nb * (PF_MAXCHUNK + 1)));
        uint8_t *d_chunkOrd = nullptr;
        PF_CUDA(st.ensure(37, d_chunkOrd, (size_t) nb * (PF_MAXCHUNK + 2)));
        PF_CUDA(cudaMemsetAsync(d_ctr, 0, 64, stream));
        ListArgs la;
        la.tb = st.tb; la.klist = st.d_klist; la.kbits = st.d_kbits; la.useBits = st.useBits ? 1 : 0; la.seq = d_seqs + winBase; la.thr = d_thr + winBase;
        la.qWin = d_qWin; la.nQueries = nb; la.nWin = nWin; la.S = S; la.winPre = d_winPre; la.slots = d_slots; la.tiles = d_tiles;
        la.counter = d_ctr; la.stats = d_stats;
        tmr.begin(ST_LISTS, stream);
        kmer_lists_kernel<<<(unsigned) smCount * 8, 256, 0, stream>>>(la);
        PF_CUDA(cudaGetLastError());
        tmr.end(stream);
        tmr.begin(ST_COUNT, stream);
        tile_pack_kernel<<<(nTiles + 1 + 255) / 256, 256, 0, stream>>>(d_tiles, nTiles, d_tilePacked);
        { size_t need = 0; cub::DeviceScan::ExclusiveSum(nullptr, need, d_tilePacked, d_tilePre, (int) (nTiles + 1), stream);
          if (need > tmpBytes) { tmpBytes = need; PF_CUDA(st.ensure(15, d_tmp, tmpBytes)); } }
        PF_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmpBytes, d_tilePacked, d_tilePre, (int) (nTiles + 1), stream));
        query_meta_kernel<<<(nb + 1 + 255) / 256, 256, 0, stream>>>(d_qWin, nb, d_winPre, d_tilePre, d_qTile, d_qOut, d_qPos);
        PF_CUDA(cudaGetLastError());
        ChunkArgs ca;
        ca.qWin = d_qWin; ca.winPre = d_winPre; ca.tilePre = d_tilePre; ca.tiles = d_tiles; ca.slots = d_slots; ca.qTile = d_qTile; ca.qPos = d_qPos;
        ca.nQueries = nb; ca.S = S; ca.maxDbMatches = maxDbMatches; ca.qinfo = d_qinfo; ca.bounds = d_bounds; ca.chunkOrd = d_chunkOrd; ca.errFlag = d_ctr + 5; ca.stats = d_stats;
        chunk_bounds_kernel<<<(nb * 32 + 127) / 128, 128, 0, stream>>>(ca);
        PF_CUDA(cudaGetLastError());
        tmr.end(stream);
        h_qTile.resize(nb + 1); h_qOut.resize(nb + 1); h_qPos.resize(nb + 1);
        PF_CUDA(cudaMemcpyAsync(h_qTile.data(), d_qTile, (nb + 1) * 4, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(h_qOut.data(), d_qOut, (nb + 1) * 4, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(h_qPos.data(), d_qPos, (nb + 1) * 8, cudaMemcpyDeviceToHost, stream));
        uint32_t ctr[8];
        PF_CUDA(cudaMemcpyAsync(ctr, d_ctr, 32, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        st.stats[4] += 7;
        if (ctr[5]) { err = ctr[5] == 1 ? "more than 126 overflow chunks for one query" : "internal error: chunk boundary not found"; result = FSB200_E_RANGE; break; }
        uint32_t maxOut = 0;
        for (uint32_t i = 0; i < nb; i++) {
            if (h_qPos[i + 1] - h_qPos[i] >= (1ull << 32)) { err = "a query has more than 2^32 index hits"; result = FSB200_E_RANGE; break; }
            maxOut = std::max(maxOut, h_qOut[i + 1] - h_qOut[i]);
        }
        if (result != FSB200_OK) break;
        if (tmg) fprintf(stderr, "prefilter: batch of %u queries: %llu similar k-mers, %u tiles, %llu index hits, %u out tiles; lists done at %.1f ms\n",
                         nb, (unsigned long long) totalK, nTiles, (unsigned long long) (h_qPos[nb] - h_qPos[0]), h_qOut[nb] - h_qOut[0], since());
        PF_CUDA(alloc_elems((h_qPos[nb] - h_qPos[0]) / 6 / std::max<uint64_t>(1, (h_qOut[nb] - h_qOut[0] + poolBudget / outTileBytes - 1) / (poolBudget / outTileBytes))));
        const uint64_t poolTiles = std::max<uint64_t>(maxOut, std::min<uint64_t>(h_qOut[nb] - h_qOut[0], poolBudget / outTileBytes));
        unsigned long long *d_pool = nullptr; uint16_t *d_tables = nullptr;
        PF_CUDA(st.ensure(35, d_pool, std::max<uint64_t>(1, poolTiles) * MT_R)); PF_CUDA(st.ensure(36, d_tables, std::max<uint64_t>(1, poolTiles) * (NB + MT_TROW)));

        // sub-batches of about equal size (a greedy fill leaves a last sub-batch of a few queries that cannot fill the GPU)
        const uint64_t allOut = h_qOut[nb] - h_qOut[0];
        const uint64_t nSubWant = std::max<uint64_t>(1, (allOut + poolTiles - 1) / std::max<uint64_t>(1, poolTiles));
        const uint64_t subTiles = std::min<uint64_t>(poolTiles, std::max<uint64_t>(maxOut, (allOut + nSubWant - 1) / nSubWant + maxOut));
        for (uint32_t s0 = 0; s0 < nb && result == FSB200_OK;) {
            uint32_t s1 = s0;
            while (s1 < nb && s1 - s0 < subMax && (uint64_t) (h_qOut[s1 + 1] - h_qOut[s0]) <= subTiles) s1++;
            if (s1 == s0) s1 = s0 + 1;   // a single query always fits (poolTiles >= maxOut)
            const uint32_t ns = s1 - s0;
            const uint32_t tileLo = h_qTile[s0], nTileSub = h_qTile[s1] - tileLo, outLo = h_qOut[s0];
            PF_CUDA(cudaMemsetAsync(d_ctr + 1, 0, 4, stream));
            if (nTileSub > 0) {
                ExpandArgs ea;
                ea.tiles = d_tiles; ea.tilePre = d_tilePre; ea.winPre = d_winPre; ea.qWin = d_qWin; ea.qPos = d_qPos; ea.slots = d_slots; ea.entries = st.d_entries;
                ea.tileLo = tileLo; ea.nTiles = nTileSub; ea.outLo = outLo; ea.S = S; ea.NB = NB; ea.pool = d_pool; ea.tables = d_tables;
                tmr.begin(ST_EXPAND, stream);
                switch (nbShift) {
                    case 4: expand_kernel<4><<<nTileSub, MT_BTHREADS, 0, stream>>>(ea); break;
                    case 5: expand_kernel<5><<<nTileSub, MT_BTHREADS, 0, stream>>>(ea); break;
                    case 6: expand_kernel<6><<<nTileSub, MT_BTHREADS, 0, stream>>>(ea); break;
                    case 7: expand_kernel<7><<<nTileSub, MT_BTHREADS, 0, stream>>>(ea); break;
                    default: expand_kernel<8><<<nTileSub, MT_BTHREADS, 0, stream>>>(ea); break;
                }
                PF_CUDA(cudaGetLastError());
                tmr.end(stream);
                DiagArgs da;
                da.pool = d_pool; da.tables = d_tables; da.qOut = d_qOut; da.outLo = outLo; da.qinfo = d_qinfo; da.bounds = d_bounds;
                da.qLo = s0; da.nQueries = ns; da.NB = NB; da.nbShift = nbShift; da.tBits = tBits; da.stateStride = stateStride;
                da.elems = d_elems; da.elemCount = d_ctr + 1; da.elemCap = elemCap;
                tmr.begin(ST_DIAG, stream);
                diag_kernel<<<(unsigned) (((uint64_t) ns * NB + MT_CWARPS - 1) / MT_CWARPS), MT_CWARPS * 32, diagSmem, stream>>>(da);
                PF_CUDA(cudaGetLastError());
                tmr.end(stream);
                st.stats[4] += 2; st.stats[7] += 1;
            }
            uint32_t nElems = 0;
            PF_CUDA(cudaMemcpyAsync(&nElems, d_ctr + 1, 4, cudaMemcpyDeviceToHost, stream));
            PF_CUDA(cudaStreamSynchronize(stream));
            if (tmg) fprintf(stderr, "prefilter: sub-batch %u..%u: %u tiles, %u double hits; diag done at %.1f ms\n", s0, s1, nTileSub, nElems, since());
            if (nElems > elemCap) {
                if (ns <= 1) {   // one query alone: larger buffers, same query again
                    if (elemCap >= elemCapMax) { err = "more double hits for a single query than the element buffers hold (FSB200_PF_ELEMCAP)"; result = FSB200_E_RANGE; break; }
                    PF_CUDA(alloc_elems((uint64_t) nElems + nElems / 8));
                    continue;
                }
                subMax = std::max<size_t>(1, ns / 2);
                continue;   // same queries again, in a smaller pass
            }
            st.stats[2] += nElems;
            const size_t q0 = qb + s0;   // first query of the sub-batch in the call's numbering
            uint32_t nCands = 0, nKeep = 0;
            if (nElems > 0) {
                // sort by key (query, target, stream position): sort (key, index) pairs, then gather
                tmr.begin(ST_SORT, stream);
                uint32_t *k32a = reinterpret_cast<uint32_t *>(d_k1), *k32b = reinterpret_cast<uint32_t *>(d_k2);
                pf_extract_keys<<<(nElems + 255) / 256, 256, 0, stream>>>(d_elems, nElems, k32a, d_v1);
                { size_t need = 0; cub::DeviceRadixSort::SortPairs(nullptr, need, k32a, k32b, d_v1, d_v2, (int) nElems, 0, 32, stream);
                  if (need > tmpBytes) { tmpBytes = need; PF_CUDA(st.ensure(15, d_tmp, tmpBytes)); } }
                PF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, k32a, k32b, d_v1, d_v2, (int) nElems, 0, (int) (tBits + qBits), stream));
                pf_gather_elems<<<(nElems + 255) / 256, 256, 0, stream>>>(d_elems, d_v2, nElems, d_elems2);
                tmr.end(stream);
                ScoreArgs sa;
                sa.elems = d_elems2; sa.n = nElems; sa.tBits = tBits; sa.queries = d_qd + q0; sa.profiles = d_prof; sa.tSeq = st.d_masked;
                sa.tOff = d_tOff; sa.tLen = d_tLen; sa.score = d_score; sa.kept = d_kept; sa.stats = d_stats;
                tmr.begin(ST_SCORE, stream);
                score_kernel<<<(nElems + 255) / 256, 256, 0, stream>>>(sa);
                PF_CUDA(cudaGetLastError());
                tmr.end(stream);
                PF_CUDA(cudaMemsetAsync(d_ctr + 3, 0, 4, stream));
                PF_CUDA(cudaMemsetAsync(d_ctr + 7, 0, 4, stream));
                PF_CUDA(cudaMemsetAsync(d_hist, 0, (size_t) ns * 256 * 4, stream));
                BestArgs ba;
                ba.elems = d_elems2; ba.score = d_score; ba.kept = d_kept; ba.n = nElems; ba.tBits = tBits; ba.binMask = binSize - 1;
                ba.cands = d_cands; ba.candCount = d_ctr + 3; ba.candCap = elemCap; ba.qinfo = d_qinfo + s0; ba.errFlag = d_ctr + 6;
                ba.chunkOrd = d_chunkOrd + (size_t) s0 * (PF_MAXCHUNK + 2); ba.bounds = d_bounds + (size_t) s0 * (PF_MAXCHUNK + 1); ba.minScore = (uint32_t) std::max(1, cfg.min_diag_score);
                ba.hist = d_hist;
                tmr.begin(ST_SELECT, stream);
                best_kernel<<<(nElems + 255) / 256, 256, 0, stream>>>(ba);
                PF_CUDA(cudaGetLastError());
                // only the candidates that reach their query's selection threshold are sorted (there are at most nElems candidates)
                cand_threshold_kernel<<<(ns + 127) / 128, 128, 0, stream>>>(d_hist, ns, maxHits, d_thrQ);
                pf_extract_cand_keys<<<(nElems + 255) / 256, 256, 0, stream>>>(d_cands, d_ctr + 3, elemCap, d_thrQ, d_k1, d_v1, d_ctr + 7);
                PF_CUDA(cudaGetLastError());
                tmr.end(stream);
                uint32_t c2[2] = {0, 0};
                PF_CUDA(cudaMemcpyAsync(&nCands, d_ctr + 3, 4, cudaMemcpyDeviceToHost, stream));
                PF_CUDA(cudaMemcpyAsync(&nKeep, d_ctr + 7, 4, cudaMemcpyDeviceToHost, stream));
                PF_CUDA(cudaMemcpyAsync(&c2[0], d_ctr + 6, 4, cudaMemcpyDeviceToHost, stream));
                PF_CUDA(cudaStreamSynchronize(stream));
                if (nCands > elemCap) { err = "candidate buffer overflow"; result = FSB200_E_RANGE; break; }
                if (c2[0]) { err = "overflow replay limits exceeded (more than 126 chunks or 128 double hits of one target)"; result = FSB200_E_RANGE; break; }
            }
            tmr.begin(ST_SELECT, stream);
            if (nKeep > 0) {
                PF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, d_k1, d_k2, d_v1, d_v2, (int) nKeep, 0, 64, stream));
                pf_gather_cands<<<(nKeep + 255) / 256, 256, 0, stream>>>(d_cands, d_v2, nKeep, d_cands2);
            }
            cand_offsets_kernel<<<(ns + 1 + 255) / 256, 256, 0, stream>>>(d_cands2, nKeep, ns, d_candOff);
            PF_CUDA(cudaGetLastError());
            SelectArgs se;
            se.cands = d_cands2; se.candOff = d_candOff; se.queries = d_qd + q0; se.profiles = d_prof; se.tSeq = st.d_masked;
            se.tOff = d_tOff; se.tLen = d_tLen; se.nQueries = ns; se.maxHits = maxHits; se.minDiagScore = (uint32_t) cfg.min_diag_score;
            se.queryBase = (uint32_t) q0; se.outStride = (uint32_t) cfg.max_seqs; se.hits = d_hits; se.counts = d_counts; se.qSeqs = d_seqs; se.scratch = d_score;
            se.identity = identity ? d_ident + q0 : nullptr;
            select_kernel<<<ns, 256, 0, stream>>>(se);
            PF_CUDA(cudaGetLastError());
            tmr.end(stream);
            st.stats[4] += 10;
            s0 = s1;
        }
        qb += nb;
    }
    if (result == FSB200_OK) {
        PF_CUDA(cudaMemcpyAsync(hits, d_hits, nQ * (size_t) cfg.max_seqs * sizeof(fsb200_hit), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(counts, d_counts, nQ * 4, cudaMemcpyDeviceToHost, stream));
        unsigned long long hs[16];
        PF_CUDA(cudaMemcpyAsync(hs, d_stats, 128, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        st.stats[0] = hs[0]; st.stats[1] = hs[1]; st.stats[3] = hs[3]; st.stats[5] = hs[2];
        tmr.collect(st.stageNs, ST_N);
        st.stats[6] = st.stageNs[ST_LISTS] + st.stageNs[ST_EXPAND] + st.stageNs[ST_DIAG];
        if (tmg) fprintf(stderr, "prefilter: results on host at %.1f ms; kernels: count+scan %.2f, lists %.2f, expand %.2f, diag %.2f, sort %.2f, score %.2f, best+select %.2f ms\n",
                         since(), st.stageNs[0] / 1e6, st.stageNs[1] / 1e6, st.stageNs[2] / 1e6, st.stageNs[3] / 1e6, st.stageNs[4] / 1e6, st.stageNs[5] / 1e6, st.stageNs[6] / 1e6);
    } else {
        cudaStreamSynchronize(stream);
        tmr.collect(st.stageNs, ST_N);
    }
    return result;
}

}  // namespace prefilter
}  // namespace fsb
