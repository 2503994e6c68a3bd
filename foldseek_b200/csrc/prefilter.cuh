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
// GPU decomposition (DESIGN.md "Prefilter"): one persistent CTA per query walks the query windows in order.  For a
// window the CTA enumerates the similar k-mers in the reference's list order, gathers their index lists into a
// shared-memory hit FIFO *in stream order* (positions from block scans), and then consumes the FIFO: warp w owns the
// targets with (id mod 16 == w), reads the FIFO 32 hits at a time and resolves same-target hits inside a 32-hit group
// with __match_any_sync, so every target sees its hits exactly in the reference's order.  The per-target "previous
// diagonal" byte (the reference's duplicateBitArray) lives in shared memory (or L2-resident global memory when the
// DB has more targets than fit).  Double hits are appended to a global list keyed by (query, target, stream position);
// a radix sort brings each target's double hits together in stream order for the run-collapse, the ungapped
// rescoring and the per-target maximum; a second sort realises the reference's (score, bin, array position) order
// that decides which hits survive the --max-seqs cut.
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
#ifndef PF_FORCE_GLOBAL
#define PF_FORCE_GLOBAL 0  // 1: keep the per-target state in global memory even when it fits in shared memory (tuning experiments)
#endif
#ifndef PF_BLOCK_THREADS
#define PF_BLOCK_THREADS 512
#endif
#ifndef PF_CTAS_PER_SM
#define PF_CTAS_PER_SM 1
#endif
constexpr int PF_BLOCK = PF_BLOCK_THREADS;   // measured: 2 x 256-thread CTAs per SM (per-target state then in global memory) is 7x slower
constexpr int PF_CTAS = PF_CTAS_PER_SM;
constexpr int PF_WARPS = PF_BLOCK / 32;
#ifndef PF_KPT_V
#define PF_KPT_V 8
#endif
constexpr int PF_KPT = PF_KPT_V;              // k-mers per thread per batch
#ifndef PF_PCAP_V
#define PF_PCAP_V 512
#endif
constexpr int PF_PCAP = PF_PCAP_V;           // capacity of one per-warp partition FIFO
constexpr int PF_FIFO = PF_WARPS * PF_PCAP;   // hits per FIFO fill (8192)
#ifndef PF_MAXROW_V
#define PF_MAXROW_V 1790
#endif
constexpr int PF_MAXROW = PF_MAXROW_V;        // rows of the first 3-mer handled per segment (pre[] capacity)
constexpr int PF_NLIST = PF_BLOCK * PF_KPT;   // k-mers per batch
#ifndef PF_HPT_V
#define PF_HPT_V 2
#endif
constexpr int PF_HPT = PF_HPT_V;              // stream positions per thread per round
constexpr int PF_MAXCHUNK = 30;        // overflow chunks per query handled by the replay (2M hits each)
constexpr int PF_MAXOV = 48;           // double hits of one target of an overflowed query

struct Tables {
    uint16_t *sIdx = nullptr;     // [N3][N3] 3-mer indices of each row, sorted by score desc (stable in permutation order)
    int16_t *sScore = nullptr;    // [N3][N3]
    uint16_t *cntGe = nullptr;    // [N3][NS]: number of row entries with score >= vmin + s
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

struct PArgs {
    Tables tb;
    const uint32_t *offsets;       // 20^6 + 1
    const uint2 *klist;            // 20^6: (first entry, list length) of every k-mer; singleton lists are stored inline (bit 31 of x)
    const uint32_t *kbits;         // 20^6 bits, "list not empty", addressed [first 3-mer][second 3-mer]: the probes of one warp share the
                                   // first 3-mer, so they fall into one 1000-byte row that stays in L1 — empty k-mers never leave the SM
    const uint64_t *entries;       // id | pos << 32
    const uint8_t *q3Di;
    const int16_t *thr;
    const QueryDesc *queries;
    uint32_t nQueries;
    uint32_t nTargets;
    uint32_t tBits;
    uint8_t *globalState;          // nCTA * stateBytes when the state does not fit in shared memory, else null
    uint32_t stateStride;          // bytes per warp partition: target t lives at (t % PF_WARPS) * stateStride + t / PF_WARPS, so the bytes a
                                   // warp touches never share a 128-byte line with another warp's (in global memory the L1 orders a load
                                   // behind every pending store to its line: with interleaved bytes a drain load took ~17 000 cycles)
    uint32_t stateBytes;           // PF_WARPS * stateStride
    uint32_t *queryCounter;
    Elem *elems;
    uint32_t *elemCount;
    uint32_t elemCap;
    uint64_t maxDbMatches;         // 2 * max(1e6, N) (QueryMatcher.cpp:45)
    unsigned long long *stats;     // [0] similar k-mers, [1] index hits, [2] overflow flushes
    uint32_t *overflowFlag;        // set when a single k-mer list reaches maxDbMatches (the reference abandons the query)
    uint32_t *qinfo;               // per query: number of chunks | bit31 = last chunk empty
};

__device__ __forceinline__ int cnt_ge(const Tables &tb, int row, int v) {
    int s = v - tb.vmin;
    if (s <= 0) return N3;
    if (s >= tb.NS) return 0;
    return (int) tb.cntGe[(size_t) row * tb.NS + s];
}

// One 32-hit group of a FIFO: every lane with `mine` set applies its hit to the per-target state in lane order.
// key = target; returns true if the lane's hit repeats the previous 8-bit diagonal of its target.
__device__ __forceinline__ bool apply_group(bool mine, uint32_t t, uint32_t d16, uint8_t *state, uint32_t stateStride) {
    // MATCH.ANY iterates over the distinct keys of the warp: all foreign lanes share one dummy key
    const uint32_t key = mine ? t : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int lane = threadIdx.x & 31;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int maxRank = __reduce_max_sync(0xffffffffu, mine ? rank : 0);
    bool emit = false;
    for (int r = 0; r <= maxRank; r++) {
        if (mine && rank == r) {
            const uint8_t d8 = (uint8_t) d16;
            uint8_t *p = state + (size_t) (t & (PF_WARPS - 1)) * stateStride + (t / PF_WARPS);
            emit = (*p == d8);                // prev diagonal (0 initially: CacheFriendlyOperations.cpp:186,205)
            *p = d8;
        }
        __syncwarp();
    }
    return emit;
}

__device__ __forceinline__ void emit_group(bool emit, uint32_t t, uint32_t d16, uint32_t streamPos, uint64_t qKey, uint32_t chunkId,
                                            Elem *elems, uint32_t *elemCount, uint32_t elemCap) {
    const int lane = threadIdx.x & 31;
    const uint32_t em = __ballot_sync(0xffffffffu, emit);
    if (em) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(elemCount, (uint32_t) __popc(em));
        slot = __shfl_sync(0xffffffffu, slot, 0) + (uint32_t) __popc(em & ((1u << lane) - 1u));
        if (emit && slot < elemCap) {
            Elem e; e.key = qKey | ((uint64_t) t << 32) | (uint64_t) streamPos; e.diag = d16; e.pad = chunkId;
            elems[slot] = e;
        }
    }
}

// Partitioned consumer: warp w drains its own FIFO; entry = target (22 bits) | diag16 << 22 | stream position << 38.
__device__ __forceinline__ void consume_partition(const uint64_t *pf, int n, uint8_t *state, uint32_t stateStride, uint64_t qKey, uint32_t chunkId,
                                                   Elem *elems, uint32_t *elemCount, uint32_t elemCap) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n; base += 32) {
        const int pos = base + lane;
        const bool valid = pos < n;
        const uint64_t v = valid ? pf[pos] : 0ull;
        const uint32_t t = (uint32_t) v & 0x3fffffu;
        const uint32_t d16 = (uint32_t) (v >> 22) & 0xffffu;
        const bool emit = apply_group(valid, t, d16, state, stateStride);
        emit_group(emit, t, d16, (uint32_t) (v >> 38), qKey, chunkId, elems, elemCount, elemCap);
    }
}

__global__ void __launch_bounds__(PF_BLOCK, PF_CTAS) match_kernel(PArgs a, uint32_t smemStateBytes) {
    extern __shared__ __align__(16) uint8_t dynsmem[];
    uint64_t *pfifo = reinterpret_cast<uint64_t *>(dynsmem);                              // PF_WARPS * PF_PCAP
    uint2 *lists = reinterpret_cast<uint2 *>(dynsmem + (size_t) PF_FIFO * 8);             // PF_NLIST + 1: (first entry, first stream position in batch)
    uint32_t *pre = reinterpret_cast<uint32_t *>(lists + PF_NLIST + 2);                   // PF_MAXROW + 2
    uint16_t *rowIdx = reinterpret_cast<uint16_t *>(pre + PF_MAXROW + 2);                 // PF_MAXROW + 2: 3-mer index of every row of the segment
    uint16_t *colIdx = rowIdx + PF_MAXROW + 2;                                            // N3: 3-mer indices of the second row (prefix used by this window)
    uint8_t *state = (smemStateBytes > 0) ? reinterpret_cast<uint8_t *>(colIdx + N3)
                                          : a.globalState + (size_t) blockIdx.x * a.stateBytes;
    typedef cub::BlockScan<uint32_t, PF_BLOCK> Scan;
    typedef cub::BlockScan<unsigned long long, PF_BLOCK> Scan64;
    typedef cub::BlockReduce<uint32_t, PF_BLOCK> Reduce;
    __shared__ union { typename Scan::TempStorage s32; typename Scan64::TempStorage s64; typename Reduce::TempStorage red; } collTmp;   // used one at a time, barriers in between
    __shared__ uint32_t s_query;
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_ptot[PF_WARPS];                 // fill level of each partition FIFO
    __shared__ uint16_t s_wc[PF_HPT][PF_WARPS][PF_WARPS + 1];     // [sub-step][warp][partition] hits of the current round
    __shared__ uint16_t s_off[PF_HPT][PF_WARPS][PF_WARPS + 1];    // [sub-step][warp][partition] write offset of the current round
    __shared__ uint32_t s_need[PF_WARPS], s_oldtot[PF_WARPS];
    __shared__ uint32_t s_drain, s_skew;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long nKmers = 0, nHitsTotal = 0;
    long long tm[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tlast = clock64();
#define PF_TICK(slot) do { if (tid == 0) { const long long now_ = clock64(); tm[slot] += now_ - tlast; tlast = now_; } } while (0)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_query = atomicAdd(a.queryCounter, 1u);
        __syncthreads();
        const uint32_t qi = s_query;
        if (qi >= a.nQueries) break;
        const QueryDesc qd = a.queries[qi];
        const uint8_t *seq = a.q3Di + qd.seqOff;
        const int16_t *thrArr = a.thr + qd.thrOff;
        const int L = (int) qd.len;
        const uint64_t qKey = (uint64_t) qi << (32 + a.tBits);
        for (uint32_t k = tid; k < a.stateBytes / 4; k += PF_BLOCK) reinterpret_cast<uint32_t *>(state)[k] = 0;
        if (tid < PF_WARPS) s_ptot[tid] = 0;
        if (tid == 0) { s_drain = 0; s_skew = 0; }
        uint32_t streamBase = 0;
        uint64_t chunkHits = 0;
        uint32_t chunkId = 0;
        __syncthreads();

        for (int i = 0; i + KSPAN <= L; i++) {
            const int thr = thrArr[i];
            if (thr < 0) continue;   // window contains X (QueryMatcher.cpp:264-268)
            const int r0 = seq[i] + KA * seq[i + 1] + KA * KA * seq[i + 3];
            const int r1 = seq[i + 5] + KA * seq[i + 8] + KA * KA * seq[i + 9];
            const int16_t *s0 = a.tb.sScore + (size_t) r0 * N3;
            const uint16_t *i0 = a.tb.sIdx + (size_t) r0 * N3;
            const uint16_t *i1 = a.tb.sIdx + (size_t) r1 * N3;
            const int s1top = a.tb.sScore[(size_t) r1 * N3];
            PF_TICK(5);
            const int n0all = cnt_ge(a.tb, r0, (int) (short) (thr - s1top));
            if (n0all > 0) {
                const int maxC = cnt_ge(a.tb, r1, (int) (short) (thr - (int) s0[0]));   // rows are sorted by score: row 0 pairs with the most columns
                for (int c = tid; c < maxC; c += PF_BLOCK) colIdx[c] = i1[c];
            }
            // the rows of the first 3-mer are walked in segments of PF_MAXROW so that pre[] stays small
            for (int seg0 = 0; seg0 < n0all; seg0 += PF_MAXROW) {
                const int n0 = min(PF_MAXROW, n0all - seg0);
                // pre[x] = number of similar k-mers generated by the first x rows of this segment
                if (tid == 0) s_carry = 0;
                __syncthreads();
                for (int base = 0; base < n0; base += PF_BLOCK) {
                    const int x = base + tid;
                    uint32_t c = 0;
                    if (x < n0) { c = (uint32_t) cnt_ge(a.tb, r1, (int) (short) (thr - (int) s0[seg0 + x])); rowIdx[x] = i0[seg0 + x]; }
                    uint32_t exv, tot;
                    Scan(collTmp.s32).ExclusiveSum(c, exv, tot);
                    const uint32_t carry = s_carry;
                    if (x < n0) pre[x] = carry + exv;
                    __syncthreads();
                    if (tid == 0) s_carry = carry + tot;
                    __syncthreads();
                }
                const uint32_t K = s_carry;
                if (tid == 0) pre[n0] = K;
                __syncthreads();
                nKmers += (tid == 0) ? K : 0;
                PF_TICK(0);

                for (uint32_t kb = 0; kb < K; kb += PF_NLIST) {
                    // ---- A: each thread resolves PF_KPT consecutive k-mer ranks (thread-major order == list order) ----
                    const uint32_t x0 = kb + (uint32_t) tid * PF_KPT;
                    uint32_t myHits = 0;
                    uint2 my[PF_KPT];
#pragma unroll
                    for (int k = 0; k < PF_KPT; k++) my[k] = make_uint2(0, 0);
                    if (x0 < K) {
                        // three separate passes so that the global loads of all PF_KPT k-mers are in flight together:
                        // (1) rows / column ranks from shared memory, (2) second 3-mer indices, (3) index offsets
                        int lo = 0, hi = n0;   // largest row with pre[row] <= x0
                        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= x0) lo = mid; else hi = mid; }
                        int arow = lo;
                        uint32_t nextPre = pre[arow + 1];
                        uint32_t kmA[PF_KPT], cIdx[PF_KPT];
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            const uint32_t x = x0 + k;
                            kmA[k] = 0; cIdx[k] = 0;
                            if (x < K) {
                                while (x >= nextPre) { arow++; nextPre = pre[arow + 1]; }
                                kmA[k] = rowIdx[arow];
                                cIdx[k] = x - pre[arow];
                            }
                        }
                        uint32_t kmer[PF_KPT];
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) kmer[k] = (x0 + k < K) ? kmA[k] + (uint32_t) colIdx[cIdx[k]] * (uint32_t) N3 : 0u;
                        PF_TICK(1);
                        uint32_t l8[PF_KPT];
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            const uint32_t b = kmA[k] * (uint32_t) N3 + (uint32_t) colIdx[cIdx[k]];
                            l8[k] = (x0 + k < K) ? ((__ldg(a.kbits + (b >> 5)) >> (b & 31u)) & 1u) : 0u;
                        }
                        uint2 ol[PF_KPT];
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) ol[k] = l8[k] ? __ldg(a.klist + kmer[k]) : make_uint2(0, 0);
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            // inline singleton: x = 0x80000000 | target, y = 1 | position << 16
                            if (l8[k]) { my[k] = ol[k]; myHits += (ol[k].x & 0x80000000u) ? 1u : ol[k].y; }
                        }
                    }
                    PF_TICK(4);
                    // one scan for both prefix sums: hits (low 40 bits) and non-empty lists (high 24 bits)
                    uint32_t myNE = 0;
#pragma unroll
                    for (int k = 0; k < PF_KPT; k++) myNE += ((my[k].x & 0x80000000u) || my[k].y) ? 1u : 0u;
                    unsigned long long exPacked, totPacked;
                    Scan64(collTmp.s64).ExclusiveSum((unsigned long long) myHits | ((unsigned long long) myNE << 40), exPacked, totPacked);
                    const uint32_t exv = (uint32_t) (exPacked & 0xffffffffffull), tot = (uint32_t) (totPacked & 0xffffffffffull);
                    const uint32_t nNE = (uint32_t) (totPacked >> 40);
                    {
                        // lists[] holds only the non-empty lists, in list order: (first entry | inline hit, first stream position)
                        uint32_t run = exv, w = (uint32_t) (exPacked >> 40);
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            const bool inl = (my[k].x & 0x80000000u) != 0;
                            if (inl || my[k].y) {
                                // inline singletons carry their position in the spare bits: x = flag | pos[8:0] << 22 | target, y = pos[15:9] << 25 | start
                                const uint32_t pos = my[k].y >> 16;
                                lists[w++] = inl ? make_uint2(my[k].x | ((pos & 0x1ffu) << 22), run | ((pos >> 9) << 25)) : make_uint2(my[k].x, run);
                            }
                            run += inl ? 1u : my[k].y;
                        }
                    }
                    if (tid == 0) { lists[nNE] = make_uint2(0, tot); if (tot >= (1u << 25)) atomicExch(a.overflowFlag, 1u); }
                    __syncthreads();
                    PF_TICK(1);
                    nHitsTotal += (tid == 0) ? tot : 0;
                    // ---- overflow chunks (QueryMatcher.cpp:305-345): before the list of a k-mer is appended the reference checks
                    // hitsInChunk + listLen >= maxDbMatches and, if so, flushes the chunk (findDuplicates on what it has, state reset).
                    // rangeBegin..rangeEnd = stream positions of this batch that belong to the current chunk.
                    uint32_t rangeBegin = 0;
                    for (;;) {
                    uint32_t rangeEnd = tot;
                    bool flushAfter = false;
                    if (chunkHits + (uint64_t) (tot - rangeBegin) >= a.maxDbMatches) {
                        // first k-mer (list order) at or after rangeBegin whose list would overflow the chunk
                        uint32_t cand = 0xFFFFFFFFu;
                        uint32_t run = exv;
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            const uint32_t len_k = (my[k].x & 0x80000000u) ? 1u : my[k].y;
                            if (run >= rangeBegin && cand == 0xFFFFFFFFu &&
                                chunkHits + (uint64_t) (run - rangeBegin) + len_k >= a.maxDbMatches) cand = run;
                            run += len_k;
                        }
                        const uint32_t first = Reduce(collTmp.red).Reduce(cand, cub::Min());
                        if (tid == 0) s_carry = first;
                        __syncthreads();
                        const uint32_t cut = s_carry;
                        __syncthreads();
                        if (cut != 0xFFFFFFFFu) { rangeEnd = cut; flushAfter = true; }
                    }
                    // ---- B: load-balanced expansion, PF_HPT * PF_BLOCK consecutive stream positions per round ----
                    // thread tid takes positions h0 + u * PF_BLOCK + tid (u < PF_HPT): all loads of a round are in flight together
                    // software pipeline: the entry loads of round r+1 are issued before round r is routed
                    uint64_t nEn[PF_HPT]; uint2 nLi[PF_HPT]; bool nV[PF_HPT];
                    auto fetch = [&](uint32_t h0f) {
#pragma unroll
                        for (int u = 0; u < PF_HPT; u++) {
                            const uint32_t h = h0f + (uint32_t) u * PF_BLOCK + (uint32_t) tid;
                            nV[u] = h < rangeEnd; nEn[u] = 0; nLi[u] = make_uint2(0, 0);
                            // the 32 lanes of a warp hold 32 consecutive stream positions, so their lists are among the 32 non-empty
                            // lists that follow the owner of the first position: one uniform binary search, one coalesced load, then a
                            // 5-step search over the lanes' start values with shuffles
                            const uint32_t hFirst = h - (uint32_t) lane;
                            if (hFirst < rangeEnd) {
                                int lo = 0, hi = (int) nNE + 1;   // last list whose first position is <= hFirst (same address in all lanes)
                                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((lists[mid].y & 0x1ffffffu) <= hFirst) lo = mid; else hi = mid; }
                                const uint32_t jm = min((uint32_t) lo + (uint32_t) lane, nNE);
                                const uint2 mine = lists[jm];
                                const uint32_t st = ((uint32_t) lo + (uint32_t) lane <= nNE) ? (mine.y & 0x1ffffffu) : 0xFFFFFFFFu;
                                int sel = 0;
#pragma unroll
                                for (int step = 16; step >= 1; step >>= 1) {
                                    const uint32_t v = __shfl_sync(0xffffffffu, st, (sel + step) & 31);
                                    if (sel + step < 32 && v <= h) sel += step;
                                }
                                nLi[u].x = __shfl_sync(0xffffffffu, mine.x, sel);
                                nLi[u].y = __shfl_sync(0xffffffffu, mine.y, sel);
                            }
                            if (nV[u] && !(nLi[u].x & 0x80000000u)) nEn[u] = __ldg(a.entries + nLi[u].x + (h - (nLi[u].y & 0x1ffffffu)));
                        }
                    };
                    if (rangeBegin < rangeEnd) fetch(rangeBegin);
                    for (uint32_t h0 = rangeBegin; h0 < rangeEnd; h0 += PF_BLOCK * PF_HPT) {
                        uint32_t ht[PF_HPT], hd[PF_HPT], hp[PF_HPT], hr[PF_HPT];
                        bool hv[PF_HPT];
                        uint64_t cEn[PF_HPT]; uint2 cLi[PF_HPT];
#pragma unroll
                        for (int u = 0; u < PF_HPT; u++) { cEn[u] = nEn[u]; cLi[u] = nLi[u]; hv[u] = nV[u]; }
                        if (h0 + PF_BLOCK * PF_HPT < rangeEnd) fetch(h0 + PF_BLOCK * PF_HPT);
#pragma unroll
                        for (int u = 0; u < PF_HPT; u++) {
                            ht[u] = 0; hd[u] = 0; hp[u] = PF_WARPS;
                            if (hv[u]) {
                                if (cLi[u].x & 0x80000000u) {   // singleton list stored inline in the k-mer table
                                    ht[u] = cLi[u].x & 0x3fffffu;
                                    hd[u] = (uint32_t) (i - (int) (((cLi[u].x >> 22) & 0x1ffu) | ((cLi[u].y >> 25) << 9))) & 0xffffu;
                                } else {
                                    ht[u] = (uint32_t) cEn[u];
                                    hd[u] = (uint32_t) (i - (int) (uint32_t) (cEn[u] >> 32)) & 0xffffu;
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < PF_HPT; u++) if (hv[u]) hp[u] = ht[u] & (PF_WARPS - 1);
                        PF_TICK(6);
                        // a round normally goes through in one piece; if one partition would receive more than its FIFO holds
                        // (skewed target ids) it is replayed in slices of PF_PCAP positions, each of which always fits
                        int nSub = 1;
                        for (int sub = 0; sub < nSub; sub++) {
                            bool act[PF_HPT];
#pragma unroll
                            for (int u = 0; u < PF_HPT; u++) {
                                act[u] = hv[u] && (nSub == 1 || (int) ((u * PF_BLOCK + tid) / PF_PCAP) == sub);
                                const uint32_t peers = __match_any_sync(0xffffffffu, act[u] ? hp[u] : (uint32_t) PF_WARPS);
                                hr[u] = (uint32_t) __popc(peers & ((1u << lane) - 1u));
                                for (int p = lane; p <= PF_WARPS; p += 32) s_wc[u][warp][p] = 0;
                                __syncwarp();
                                if (act[u] && hr[u] == 0) s_wc[u][warp][hp[u]] = (uint16_t) __popc(peers);
                            }
                            __syncthreads();
                            PF_TICK(7);
                            if (tid < PF_WARPS) {
                                // offsets in stream order: sub-step u of all warps precedes sub-step u+1
                                uint32_t basep = s_ptot[tid];
                                bool skew = false;
#pragma unroll
                                for (int u = 0; u < PF_HPT; u++)
                                    for (int w = 0; w < PF_WARPS; w++) { s_off[u][w][tid] = (uint16_t) basep; basep += s_wc[u][w][tid]; }
                                const uint32_t need = basep - s_ptot[tid];
                                if (need > PF_PCAP) { s_skew = 1; skew = true; }
                                else if (basep > PF_PCAP) s_drain = 1;
                                if (!skew) s_need[tid] = need;
                            }
                            __syncthreads();
                            if (s_skew) {   // only possible with nSub == 1
                                __syncthreads();
                                if (tid == 0) { s_skew = 0; s_drain = 0; }
                                nSub = (PF_BLOCK * PF_HPT) / PF_PCAP; sub = -1;
                                __syncthreads();
                                continue;
                            }
                            PF_TICK(8);
                            uint32_t rebase = 0;   // subtract the old fill level when the FIFOs are drained first
                            if (s_drain) {
                                PF_TICK(2);
                                consume_partition(pfifo + warp * PF_PCAP, (int) s_ptot[warp], state, a.stateStride, qKey, chunkId, a.elems, a.elemCount, a.elemCap);
                                rebase = 1;
                                __syncthreads();
                                PF_TICK(3);
                                if (tid < PF_WARPS) s_oldtot[tid] = s_ptot[tid];
                                __syncthreads();
                                if (tid < PF_WARPS) s_ptot[tid] = 0;
                                if (tid == 0) s_drain = 0;
                            }
#pragma unroll
                            for (int u = 0; u < PF_HPT; u++) {
                                if (act[u]) {
                                    const uint32_t dst = (uint32_t) s_off[u][warp][hp[u]] + hr[u] - (rebase ? s_oldtot[hp[u]] : 0u);
                                    const uint32_t h = h0 + (uint32_t) u * PF_BLOCK + (uint32_t) tid;
                                    pfifo[hp[u] * PF_PCAP + dst] = (uint64_t) ht[u] | ((uint64_t) hd[u] << 22) | ((uint64_t) (streamBase + h) << 38);
                                }
                            }
                            __syncthreads();
                            PF_TICK(9);
                            if (tid < PF_WARPS) s_ptot[tid] += s_need[tid];
                            // (the next round's first barrier orders this update before its readers)
                        }
                    }
                    PF_TICK(2);
                    chunkHits += (uint64_t) (rangeEnd - rangeBegin);
                    if (!flushAfter) break;
                    // ---- flush: the chunk ends before the k-mer at stream position rangeEnd ----
                    __syncthreads();
                    consume_partition(pfifo + warp * PF_PCAP, (int) s_ptot[warp], state, a.stateStride, qKey, chunkId, a.elems, a.elemCount, a.elemCap);
                    __syncthreads();
                    if (tid < PF_WARPS) s_ptot[tid] = 0;
                    for (uint32_t k = tid; k < a.stateBytes / 4; k += PF_BLOCK) reinterpret_cast<uint32_t *>(state)[k] = 0;
                    chunkId++;
                    chunkHits = 0;
                    rangeBegin = rangeEnd;
                    {   // a single list that alone reaches maxDbMatches makes the reference abandon the query (QueryMatcher.cpp:341-343)
                        uint32_t run = exv; bool huge = false;
#pragma unroll
                        for (int k = 0; k < PF_KPT; k++) {
                            const uint32_t len_k = (my[k].x & 0x80000000u) ? 1u : my[k].y;
                            if (run == rangeBegin && len_k >= a.maxDbMatches) huge = true;
                            run += len_k;
                        }
                        if (huge) atomicExch(a.overflowFlag, 1u);
                    }
                    __syncthreads();
                    }   // chunk loop
                    streamBase += tot;
                }
            }   // row segments
        }
        // drain what is left of this query
        __syncthreads();
        consume_partition(pfifo + warp * PF_PCAP, (int) s_ptot[warp], state, a.stateStride, qKey, chunkId, a.elems, a.elemCount, a.elemCap);
        if (tid == 0) {
            // chunks of this query and whether the last one is empty (then the reference returns no hit at all, QueryMatcher.cpp:358-366)
            a.qinfo[qi] = (chunkId + 1) | ((chunkId > 0 && chunkHits == 0) ? 0x80000000u : 0u);
            if (chunkId > 0) atomicAdd(a.stats + 2, (unsigned long long) chunkId);
        }
    }
    if (tid == 0) { atomicAdd(a.stats + 0, nKmers); atomicAdd(a.stats + 1, nHitsTotal);
        for (int k = 0; k < 10; k++) atomicAdd(a.stats + 4 + k, (unsigned long long) tm[k]); }
}

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
    const int head = min(len, (int) ((4u - (uint32_t) (reinterpret_cast<uintptr_t>(d) & 3u)) & 3u));
    for (; k < head; k++) { sc = max(sc + p[k * 22 + d[k]], 0); mx = max(mx, sc); }
    for (; k + 4 <= len; k += 4) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(d + k));
        const int8_t *pk = p + k * 22;
        sc = max(sc + pk[w & 0xffu], 0); mx = max(mx, sc);
        sc = max(sc + pk[22 + ((w >> 8) & 0xffu)], 0); mx = max(mx, sc);
        sc = max(sc + pk[44 + ((w >> 16) & 0xffu)], 0); mx = max(mx, sc);
        sc = max(sc + pk[66 + (w >> 24)], 0); mx = max(mx, sc);
    }
    for (; k < len; k++) { sc = max(sc + p[k * 22 + d[k]], 0); mx = max(mx, sc); }
    cells += (unsigned long long) len;
    return mx;
}

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
};

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

__global__ void best_kernel(BestArgs a) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.n) return;
    const uint32_t hiKey = (uint32_t) (a.elems[e].key >> 32);
    if (e > 0 && (uint32_t) (a.elems[e - 1].key >> 32) == hiKey) return;   // not a segment head
    const uint32_t q = hiKey >> a.tBits, t = hiKey & ((1u << a.tBits) - 1u);
    const uint32_t qi = a.qinfo[q];
    const int nChunks = (int) (qi & 0x7fffffffu);
    if (qi & 0x80000000u) return;   // last chunk empty: the reference returns nothing for this query
    int best = 0; uint32_t bestDiag = 0, orderKey = 0;
    if (nChunks <= 1) {
        uint32_t bestE = e;
        for (uint32_t k = e; k < a.n && (uint32_t) (a.elems[k].key >> 32) == hiKey; k++) {
            if (a.kept[k] && (int) a.score[k] > best) { best = a.score[k]; bestE = k; }
        }
        bestDiag = a.elems[bestE].diag & 0xffffu; orderKey = (uint32_t) a.elems[bestE].key;
    } else {
        OvElem x[PF_MAXOV];
        int m = 0;
        for (uint32_t k = e; k < a.n && (uint32_t) (a.elems[k].key >> 32) == hiKey; k++) {
            if (!a.kept[k]) continue;
            if (m >= PF_MAXOV) { atomicExch(a.errFlag, 1u); return; }
            x[m].diag = (uint16_t) a.elems[k].diag; x[m].score = a.score[k]; x[m].count = 0; x[m].chunk = (uint8_t) (a.elems[k].pad + 1);
            x[m].pos = (uint32_t) a.elems[k].key; m++;
        }
        if (nChunks > PF_MAXCHUNK) { atomicExch(a.errFlag, 1u); return; }
        if (overflow_replay(x, m, nChunks) < 0) return;
        best = x[0].count; bestDiag = x[0].diag;
        int rank[PF_MAXCHUNK + 2], dir[PF_MAXCHUNK + 2];
        chunk_order(nChunks, rank, dir);
        const uint32_t pos26 = x[0].pos & 0x3ffffffu;
        orderKey = ((uint32_t) rank[x[0].chunk] << 26) | (dir[x[0].chunk] > 0 ? pos26 : (0x3ffffffu - pos26));
    }
    if (best <= 0) return;   // zero-score elements never pass the threshold and do not touch bins >= 1
    const uint32_t slot = atomicAdd(a.candCount, 1u);
    if (slot >= a.candCap) return;   // reported by the host (nCands > candCap)
    Cand c;
    c.key = ((uint64_t) q << 52) | ((uint64_t) (255 - best) << 44) | ((uint64_t) (t & a.binMask) << 32) | (uint64_t) orderKey;
    c.target = t;
    c.diagScore = bestDiag | ((uint32_t) best << 16);
    a.cands[slot] = c;
}

// ---- per query: threshold, cut at max_seqs in the reference's order, final (score desc, id asc) sort ----
struct SelectArgs {
    const Cand *cands;             // sorted by key
    const uint32_t *candOff;       // per query offsets into cands (exclusive scan of perQuery)
    const QueryDesc *queries;
    const int8_t *profiles;
    const uint8_t *tSeq; const uint64_t *tOff; const uint32_t *tLen;
    uint32_t nQueries;
    uint32_t maxHits;              // min(max_seqs, nTargets)
    uint32_t minDiagScore;
    uint32_t queryBase;            // first query of this batch in the output arrays
    uint32_t outStride;
    fsb200_hit *hits;
    uint32_t *counts;
    uint32_t *truncFlag;
};

constexpr int SEL_MAX = 2048;

__global__ void __launch_bounds__(256) select_kernel(SelectArgs a) {
    __shared__ unsigned long long skey[SEL_MAX];   // (65535 - score) << 32 | target   (ascending == score desc, id asc)
    __shared__ uint32_t sdiag[SEL_MAX];
    __shared__ uint32_t s_thr, s_take;
    const uint32_t q = blockIdx.x;
    if (q >= a.nQueries) return;
    const uint32_t begin = a.candOff[q], end = a.candOff[q + 1];
    if (threadIdx.x == 0) {
        // candidates are sorted by score descending, so computeScoreThreshold (QueryMatcher.h:211-221: the score at which the
        // cumulative count from 255 downwards reaches maxHits, 0 if it never does) is the score of the maxHits-th candidate
        const uint32_t n = end - begin;
        uint32_t thr0 = 0;
        if (n >= a.maxHits && a.maxHits > 0) thr0 = 255u - (uint32_t) ((a.cands[begin + a.maxHits - 1].key >> 44) & 0xffu);
        const uint32_t thr = max(thr0, a.minDiagScore);
        if (thr0 >= 255) atomicExch(a.truncFlag, 1u);   // rescoreHits path (QueryMatcher.cpp:192-203) not supported
        // number of candidates with score >= thr (binary search in the descending list)
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; const uint32_t sc = 255u - (uint32_t) ((a.cands[begin + mid].key >> 44) & 0xffu); if (sc >= thr) lo = mid + 1; else hi = mid; }
        s_thr = thr; s_take = min(lo, a.maxHits);
    }
    __syncthreads();
    const uint32_t take = s_take;
    // cands are sorted by (score desc, bin asc, stream position asc): the first `take` are the survivors (getResult, :402-458)
    int npad = 1; while (npad < (int) take) npad <<= 1;
    for (int k = threadIdx.x; k < npad; k += blockDim.x) {
        if (k < (int) take) {
            const Cand c = a.cands[begin + k];
            uint32_t sc = c.diagScore >> 16;
            if (sc >= 255) {   // the clamped score is replaced by the real one (QueryMatcher.cpp:441-444)
                unsigned long long cells = 0;
                const QueryDesc qd = a.queries[q];
                sc = (uint32_t) ungapped_diag(a.profiles + qd.profOff, (int) qd.len, a.tSeq + a.tOff[c.target], (int) a.tLen[c.target], c.diagScore & 0xffffu, cells);
            }
            skey[k] = ((unsigned long long) (0xffffffu - sc) << 32) | c.target;
            sdiag[k] = c.diagScore & 0xffffu;
        } else { skey[k] = ~0ull; sdiag[k] = 0; }
    }
    __syncthreads();
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
    for (int k = threadIdx.x; k < (int) take; k += blockDim.x) {
        fsb200_hit hh;
        hh.target = (uint32_t) skey[k];
        hh.score = (int32_t) (0xffffffu - (uint32_t) (skey[k] >> 32));
        hh.diagonal = (int32_t) (short) sdiag[k];
        out[k] = hh;
    }
    if (threadIdx.x == 0) a.counts[a.queryBase + q] = take;
}

__global__ void pf_extract_keys(const Elem *e, uint32_t n, uint64_t *k, uint32_t *v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = e[i].key; v[i] = i; }
}
__global__ void pf_gather_elems(const Elem *e, const uint32_t *v, uint32_t n, Elem *o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = e[v[i]];
}
__global__ void pf_extract_cand_keys(const Cand *e, uint32_t n, uint64_t *k, uint32_t *v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = e[i].key; v[i] = i; }
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

__global__ void klist_kernel(const uint32_t *offsets, const uint64_t *entries, uint32_t tableSize, uint2 *klist, uint32_t *kbits) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tableSize) return;
    const uint32_t o0 = offsets[k], len = offsets[k + 1] - o0;
    if (len) { const uint32_t b = (k % (uint32_t) N3) * (uint32_t) N3 + k / (uint32_t) N3; atomicOr(kbits + (b >> 5), 1u << (b & 31u)); }
    if (len == 1) {
        const uint64_t en = entries[o0];
        klist[k] = make_uint2(0x80000000u | ((uint32_t) en & 0x3fffffu), 1u | ((uint32_t) (en >> 32) << 16));
    } else klist[k] = make_uint2(o0, len);
}

// ---------------------------------------------------------------------------------------------------------
struct State {
    uint64_t stats[8] = {0};
    bool tablesReady = false, indexReady = false;
    Tables tb;
    uint32_t *d_offsets = nullptr;
    uint2 *d_klist = nullptr;
    uint32_t *d_kbits = nullptr;
    uint64_t *d_entries = nullptr;
    uint8_t *d_masked = nullptr;
    uint64_t nEntries = 0;
    int kmerThr = 0;
    struct Buf { void *p = nullptr; size_t bytes = 0; };
    Buf bufs[32];
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

    void invalidate() { indexReady = false; }
    void release() {
        cudaFree(tb.sIdx); cudaFree(tb.sScore); cudaFree(tb.cntGe); cudaFree(d_offsets); cudaFree(d_klist); cudaFree(d_kbits); cudaFree(d_entries); cudaFree(d_masked);
        tb = Tables(); d_offsets = nullptr; d_klist = nullptr; d_kbits = nullptr; d_entries = nullptr; d_masked = nullptr; tablesReady = indexReady = false;
        for (Buf &b : bufs) { if (b.p) cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
        if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); ev0 = ev1 = nullptr; }
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
    klist_kernel<<<(tableSize + 255) / 256, 256, 0, stream>>>(st.d_offsets, st.d_entries, tableSize, st.d_klist, st.d_kbits);
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
               const std::vector<uint8_t> &h_p3Di, size_t nQ, QF queryAt, fsb200_hit *hits, uint32_t *counts, std::string &err) {
    memset(st.stats, 0, sizeof(st.stats));
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
    uint8_t *d_seqs = nullptr; int16_t *d_thr = nullptr; int8_t *d_prof = nullptr; QueryDesc *d_qd = nullptr;
    PF_CUDA(st.ensure(0, d_seqs, seqTotal + 16)); PF_CUDA(st.ensure(1, d_thr, thrTotal + 16)); PF_CUDA(st.ensure(2, d_prof, profTotal + 16));
    PF_CUDA(st.ensure(3, d_qd, nQ));
    PF_CUDA(cudaMemcpyAsync(d_seqs, seqs.data(), seqTotal, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_thr, thr.data(), thrTotal * 2, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_prof, prof.data(), profTotal, cudaMemcpyHostToDevice, stream));
    PF_CUDA(cudaMemcpyAsync(d_qd, qd.data(), nQ * sizeof(QueryDesc), cudaMemcpyHostToDevice, stream));

    if (tmg) fprintf(stderr, "prefilter: host prep+upload issued at %.1f ms\n", since());
    // ---- device buffers ----
    const size_t fixedSmem = (size_t) PF_FIFO * 8 + (size_t) (PF_NLIST + 2) * 8 + (PF_MAXROW + 2) * 4 + (PF_MAXROW + 2) * 2 + (size_t) N3 * 2;
    cudaFuncAttributes fa; PF_CUDA(cudaFuncGetAttributes(&fa, match_kernel));
    const size_t smemMax = (227 * 1024) / PF_CTAS - (PF_CTAS > 1 ? 1024 : 0) - fa.sharedSizeBytes - 256;
    const uint32_t stateStride = (uint32_t) ((((size_t) nT + PF_WARPS - 1) / PF_WARPS + 127) & ~(size_t) 127);
    const uint32_t stateBytes = stateStride * PF_WARPS;
    const bool stateInSmem = !PF_FORCE_GLOBAL && fixedSmem + stateBytes <= smemMax;
    const size_t smemBytes = fixedSmem + (stateInSmem ? stateBytes : 0);
    if (tmg) fprintf(stderr, "prefilter: match kernel smem %zu B dynamic + %zu B static, per-target state in %s memory\n", smemBytes,
                     (size_t) fa.sharedSizeBytes, stateInSmem ? "shared" : "global");
    const unsigned nCta = (unsigned) std::min<size_t>((size_t) smCount * PF_CTAS, std::max<size_t>(1, nQ));
    uint8_t *d_state = nullptr;
    if (!stateInSmem) PF_CUDA(st.ensure(4, d_state, (size_t) nCta * stateBytes));
    if (!stateInSmem) {
        // The per-target state is hit once per index hit with single-byte read-modify-writes while ~1 GB of index data streams
        // through L2 at random: without protection its lines are evicted between two touches and every touch becomes a DRAM
        // sector read + partial write. Pin it with the persisting access window (best effort, clipped to the device limits).
        int dev = 0, maxPersist = 0, maxWin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t bytes = (size_t) nCta * stateBytes;
        const size_t setAside = std::min<size_t>((size_t) maxPersist, bytes + (bytes >> 3));
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside);
        cudaStreamAttrValue av; memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = d_state;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t) maxWin);
        av.accessPolicyWindow.hitRatio = (float) std::min(1.0, (double) setAside / (double) std::max<size_t>(1, av.accessPolicyWindow.num_bytes));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
        if (tmg) fprintf(stderr, "prefilter: per-target state %zu B, persisting L2 set-aside %zu B (device max %d), window max %d, hit ratio %.2f\n",
                         bytes, setAside, maxPersist, maxWin, av.accessPolicyWindow.hitRatio);
    }
    uint32_t *d_ctr = nullptr; unsigned long long *d_stats = nullptr;
    PF_CUDA(st.ensure(5, d_ctr, 16)); PF_CUDA(st.ensure(6, d_stats, 16));
    uint32_t *d_qinfo = nullptr;
    PF_CUDA(st.ensure(23, d_qinfo, batchMax + 1));
    PF_CUDA(cudaMemsetAsync(d_stats, 0, 128, stream));
    PF_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemBytes));
    if (!st.ev0) { cudaEventCreate(&st.ev0); cudaEventCreate(&st.ev1); }
    cudaEvent_t ev0 = st.ev0, ev1 = st.ev1;
    uint32_t elemCap = (uint32_t) std::min<uint64_t>(400000000ull, std::max<uint64_t>(1u << 20, (uint64_t) std::min(nQ, batchMax) * 65536ull));
    Elem *d_elems = nullptr, *d_elems2 = nullptr; uint8_t *d_score = nullptr, *d_kept = nullptr; Cand *d_cands = nullptr, *d_cands2 = nullptr;
    uint64_t *d_k1 = nullptr, *d_k2 = nullptr; uint32_t *d_v1 = nullptr, *d_v2 = nullptr; uint8_t *d_tmp = nullptr; size_t tmpBytes = 0;
    uint32_t *d_perQ = nullptr, *d_hist = nullptr, *d_candOff = nullptr;
    PF_CUDA(st.ensure(7, d_elems, elemCap)); PF_CUDA(st.ensure(8, d_elems2, elemCap));
    PF_CUDA(st.ensure(9, d_score, elemCap)); PF_CUDA(st.ensure(10, d_kept, elemCap));
    PF_CUDA(st.ensure(11, d_k1, elemCap)); PF_CUDA(st.ensure(12, d_k2, elemCap));
    PF_CUDA(st.ensure(13, d_v1, elemCap)); PF_CUDA(st.ensure(14, d_v2, elemCap));
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, d_k1, d_k2, d_v1, d_v2, (int) elemCap, 0, 64, stream);
    PF_CUDA(st.ensure(15, d_tmp, tmpBytes));
    PF_CUDA(st.ensure(16, d_perQ, batchMax + 1)); PF_CUDA(st.ensure(17, d_hist, batchMax * 256)); PF_CUDA(st.ensure(18, d_candOff, batchMax + 1));
    fsb200_hit *d_hits = nullptr; uint32_t *d_counts = nullptr;
    PF_CUDA(st.ensure(19, d_hits, std::max<size_t>(1, nQ) * cfg.max_seqs)); PF_CUDA(st.ensure(20, d_counts, std::max<size_t>(1, nQ)));
    PF_CUDA(st.ensure(21, d_cands, elemCap)); PF_CUDA(st.ensure(22, d_cands2, elemCap));
    const uint32_t candCap = elemCap;
    PF_CUDA(cudaMemsetAsync(d_counts, 0, std::max<size_t>(1, nQ) * 4, stream));
    int result = FSB200_OK;

    // queries per pass: limited by the key width and, adaptively, by the double-hit buffer (halved and retried on overflow)
    size_t curBatch = batchMax;
    for (size_t qb = 0; qb < nQ && result == FSB200_OK;) {
        const uint32_t nb = (uint32_t) std::min(curBatch, nQ - qb);
        PF_CUDA(cudaMemsetAsync(d_ctr, 0, 64, stream));
        PArgs pa;
        pa.tb = st.tb; pa.offsets = st.d_offsets; pa.klist = st.d_klist; pa.kbits = st.d_kbits; pa.entries = st.d_entries; pa.q3Di = d_seqs; pa.thr = d_thr; pa.queries = d_qd + qb;
        pa.nQueries = nb; pa.nTargets = nT; pa.tBits = tBits; pa.globalState = d_state; pa.stateStride = stateStride; pa.stateBytes = stateBytes; pa.queryCounter = d_ctr;
        pa.elems = d_elems; pa.elemCount = d_ctr + 1; pa.elemCap = elemCap; pa.maxDbMatches = maxDbMatches;
        pa.stats = d_stats; pa.overflowFlag = d_ctr + 2; pa.qinfo = d_qinfo;
        PF_CUDA(cudaEventRecord(ev0, stream));
        match_kernel<<<std::min<unsigned>(nCta, nb), PF_BLOCK, smemBytes, stream>>>(pa, stateInSmem ? 1u : 0u);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaEventRecord(ev1, stream));
        uint32_t ctr[4];
        PF_CUDA(cudaMemcpyAsync(ctr, d_ctr, 16, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
        st.stats[6] += (uint64_t) (ms * 1e6); st.stats[7] += 1; st.stats[4] += 1;
        const uint32_t nElems = ctr[1];
        if (tmg) fprintf(stderr, "prefilter: match kernel done at %.1f ms\n", since());
        if (ctr[2]) { err = "a single k-mer list reaches maxDbMatches (2*max(1e6,N)); the reference abandons such a query, not reproduced"; result = FSB200_E_RANGE; break; }
        if (nElems > elemCap) {
            if (nb <= 1) { err = "double-hit buffer overflow for a single query"; result = FSB200_E_RANGE; break; }
            curBatch = std::max<size_t>(1, nb / 2);
            continue;   // same queries again, in a smaller pass
        }
        st.stats[2] += nElems;
        if (nElems == 0) { qb += nb; continue; }
        // sort by key (query, target, stream position): sort (key, index) pairs, then gather
        pf_extract_keys<<<(nElems + 255) / 256, 256, 0, stream>>>(d_elems, nElems, d_k1, d_v1);
        PF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, d_k1, d_k2, d_v1, d_v2, (int) nElems, 0, 32 + tBits + qBits, stream));
        pf_gather_elems<<<(nElems + 255) / 256, 256, 0, stream>>>(d_elems, d_v2, nElems, d_elems2);
        ScoreArgs sa;
        sa.elems = d_elems2; sa.n = nElems; sa.tBits = tBits; sa.queries = d_qd + qb; sa.profiles = d_prof; sa.tSeq = st.d_masked;
        sa.tOff = d_tOff; sa.tLen = d_tLen; sa.score = d_score; sa.kept = d_kept; sa.stats = d_stats;
        score_kernel<<<(nElems + 255) / 256, 256, 0, stream>>>(sa);
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemsetAsync(d_ctr + 3, 0, 4, stream));
        BestArgs ba;
        ba.elems = d_elems2; ba.score = d_score; ba.kept = d_kept; ba.n = nElems; ba.tBits = tBits; ba.binMask = binSize - 1;
        ba.cands = d_cands; ba.candCount = d_ctr + 3; ba.candCap = candCap; ba.qinfo = d_qinfo; ba.errFlag = d_ctr + 5;
        best_kernel<<<(nElems + 255) / 256, 256, 0, stream>>>(ba);
        PF_CUDA(cudaGetLastError());
        uint32_t nCands = 0;
        PF_CUDA(cudaMemcpyAsync(&nCands, d_ctr + 3, 4, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        if (nCands > candCap) { err = "candidate buffer overflow"; result = FSB200_E_RANGE; break; }
        { uint32_t ef = 0; PF_CUDA(cudaMemcpy(&ef, d_ctr + 5, 4, cudaMemcpyDeviceToHost));
          if (ef) { err = "overflow replay limits exceeded (more than 30 chunks or 48 double hits of one target)"; result = FSB200_E_RANGE; break; } }
        if (nCands > 0) {
            pf_extract_cand_keys<<<(nCands + 255) / 256, 256, 0, stream>>>(d_cands, nCands, d_k1, d_v1);
            PF_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, d_k1, d_k2, d_v1, d_v2, (int) nCands, 0, 64, stream));
            pf_gather_cands<<<(nCands + 255) / 256, 256, 0, stream>>>(d_cands, d_v2, nCands, d_cands2);
        }
        cand_offsets_kernel<<<(nb + 1 + 255) / 256, 256, 0, stream>>>(d_cands2, nCands, nb, d_candOff);
        PF_CUDA(cudaGetLastError());
        SelectArgs se;
        se.cands = d_cands2; se.candOff = d_candOff; se.queries = d_qd + qb; se.profiles = d_prof; se.tSeq = st.d_masked;
        se.tOff = d_tOff; se.tLen = d_tLen; se.nQueries = nb; se.maxHits = maxHits; se.minDiagScore = (uint32_t) cfg.min_diag_score;
        se.queryBase = (uint32_t) qb; se.outStride = (uint32_t) cfg.max_seqs; se.hits = d_hits; se.counts = d_counts; se.truncFlag = d_ctr + 4;
        select_kernel<<<nb, 256, 0, stream>>>(se);
        PF_CUDA(cudaGetLastError());
        st.stats[4] += 8;
        uint32_t trunc = 0;
        PF_CUDA(cudaMemcpyAsync(&trunc, d_ctr + 4, 4, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        if (tmg) fprintf(stderr, "prefilter: select done at %.1f ms\n", since());
        if (trunc) { err = "more than max-seqs hits saturate the 8-bit diagonal score: rescoreHits path not implemented"; result = FSB200_E_RANGE; }
        qb += nb;
    }
    if (result == FSB200_OK) {
        PF_CUDA(cudaMemcpyAsync(hits, d_hits, nQ * (size_t) cfg.max_seqs * sizeof(fsb200_hit), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(counts, d_counts, nQ * 4, cudaMemcpyDeviceToHost, stream));
        unsigned long long hs[16];
        PF_CUDA(cudaMemcpyAsync(hs, d_stats, 128, cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        st.stats[0] = hs[0]; st.stats[1] = hs[1]; st.stats[3] = hs[3]; st.stats[5] = hs[2];
        if (tmg) fprintf(stderr, "prefilter: results on host at %.1f ms\n", since());
        if (getenv("FSB200_PF_TIMING")) fprintf(stderr, "match_kernel cycles (sum over CTAs): rows/pre %llu  kmer->lists %llu  expand+route %llu  drain %llu  other %llu\n", hs[4], hs[5], hs[6], hs[7], hs[9]); if (getenv("FSB200_PF_TIMING")) fprintf(stderr, "   (kmer->lists split: search+ranks %llu (in slot1)  loads %llu)  expand+route: search+entries %llu  match+count %llu  offsets %llu  fifo write %llu\n", hs[5], hs[8], hs[10], hs[11], hs[12], hs[13]);
    }
    return result;
}

}  // namespace prefilter
}  // namespace fsb
