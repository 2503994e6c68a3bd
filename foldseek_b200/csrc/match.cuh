// Stage-1 matching pipeline: similar k-mers -> index lists -> hit stream -> double-diagonal detection.
// Reference: QueryMatcher::match (lib/mmseqs/src/prefiltering/QueryMatcher.cpp:243-376), KmerGenerator::generateKmerList
// (KmerGenerator.cpp:177-286), CacheFriendlyOperations::findDuplicates (CacheFriendlyOperations.cpp:185-278,341-351).
//
// The reference walks one query at a time: windows in order, per window the similar k-mers in list order, per k-mer the index
// list (sorted by target); that "stream" of (target, diagonal) hits is scattered into L2-sized bins by target id and every
// bin is scanned with one "previous diagonal" byte per target.  Same scheme here, with HBM in the role of the binned buffer
// and shared memory in the role of the L2-resident byte array:
//
//   win_count_kernel    one thread per query window: number of similar k-mers K (closed form over the score levels of the two
//                       sorted 3-mer rows) -> prefix sums give every window a slot region and a range of slot tiles
//   kmer_lists_kernel   one warp per window: enumerates the similar k-mers in the reference's list order, probes the occupancy
//                       bitmap / reads the 8-byte list head of each and writes the non-empty heads, compacted and in order,
//                       into the window's slot region; per tile of S slots it records the number of index hits
//   (prefix sum over the tiles: every tile knows its first stream position and its output tiles)
//   chunk_bounds_kernel one warp per query: the stream positions at which the reference flushes an overflowed chunk
//                       (QueryMatcher.cpp:305-345), found from the tile prefix sums and the slots of one tile per flush
//   expand_kernel       one CTA per slot tile: expands the lists into hits (stream order = position p), MT_R positions per
//                       round; every round is counting-sorted (stable) by bucket = target mod NB in shared memory and written
//                       out as one contiguous "out tile" + its bucket offsets
//   diag_kernel         one warp per (query, bucket): walks the query's out tiles in order, reads only its bucket's segment of
//                       each, and applies the hits in stream order to the bucket's previous-diagonal bytes in shared memory
//                       (nTargets / NB bytes per warp at any DB size); double hits are staged per warp and appended to the
//                       global element list, tagged with their overflow chunk.
// No kernel needs block-wide ordering: the order that matters (per target, stream order) is carried by the data layout.
#pragma once

namespace fsb {
namespace prefilter {

constexpr int MT_R = 2048;                 // stream positions per out tile
constexpr int MT_BTHREADS = 256;           // expand_kernel block
constexpr int MT_BWARPS = MT_BTHREADS / 32;
constexpr int MT_PPT = MT_R / MT_BTHREADS; // positions per thread per round
constexpr int MT_SMAX = 1024;              // slots per tile (upper bound of the runtime value S)
constexpr int MT_MAXNB = 256;              // buckets (upper bound of the runtime value NB)
constexpr int MT_CWARPS = 7;               // warps per diag_kernel block (3 blocks of 7 x ~10 KB fit one SM at 1M targets)
constexpr int MT_STAGE = 96;               // double hits staged per warp (flushed in groups of 64)
constexpr int MT_MAP = 768;                // diag_kernel: entries of one batch of 32 tiles whose pool offsets are mapped in shared memory
constexpr int MT_LBITS = 15;               // bits of the bucket-local target (target / NB) in an out-tile entry; 32 bits remain for the position
constexpr int MT_TROW = 3;                 // extra uint16 per table row behind the NB bucket offsets: entry count, first stream position (lo, hi)
constexpr unsigned long long MT_LOW40 = (1ull << 40) - 1ull;

struct TileRec {
    uint32_t win;                          // query (batch-local) << 16 | window position i
    uint32_t nslots;                       // non-empty lists in this tile (0: unused capacity tile)
    unsigned long long hits;               // index hits of the tile
};

// number of similar k-mers of a window: rows of the first 3-mer are sorted by score, so all rows of one score level v pair with
// the same number of columns cnt_ge(r1, thr - v); the list order is (row ascending, column ascending) (KmerGenerator.cpp:256-286)
__device__ __forceinline__ uint32_t count_similar_kmers(const Tables &tb, int r0, int r1, int thr) {
    const int s0top = tb.sScore[(size_t) r0 * N3], s1top = tb.sScore[(size_t) r1 * N3];
    const int vlow = max((int) (short) (thr - s1top), tb.vmin);
    uint32_t K = 0;
    int prev = 0;
    for (int v = s0top; v >= vlow; v--) {
        const int cnt = cnt_ge(tb, r0, v);
        if (cnt != prev) K += (uint32_t) (cnt - prev) * (uint32_t) cnt_ge(tb, r1, (int) (short) (thr - v));
        prev = cnt;
    }
    return K;
}

__global__ void win_count_kernel(Tables tb, const uint8_t *seq, const int16_t *thr, uint32_t nWin, uint32_t S, unsigned long long *packed) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nWin) return;
    uint32_t K = 0;
    if (g < nWin) {
        const int t = thr[g];
        if (t >= 0) {   // -1: window contains X or runs over the end of its query (QueryMatcher.cpp:264-268)
            const int r0 = seq[g] + KA * seq[g + 1] + KA * KA * seq[g + 3];
            const int r1 = seq[g + 5] + KA * seq[g + 8] + KA * KA * seq[g + 9];
            K = count_similar_kmers(tb, r0, r1, t);
        }
    }
    packed[g] = ((unsigned long long) ((K + S - 1) / S) << 40) | (unsigned long long) K;   // [nWin] = 0: the scan's total
}

struct ListArgs {
    Tables tb;
    const uint2 *klist;                    // list heads, addressed [first 3-mer][second 3-mer]
    const uint32_t *kbits;
    int useBits;
    const uint8_t *seq;                    // concatenated query 3Di of the batch
    const int16_t *thr;                    // per window threshold, -1 = no k-mer
    const uint32_t *qWin;                  // [nQueries + 1] first window of every query of the batch
    uint32_t nQueries;
    uint32_t nWin;
    uint32_t S;
    const unsigned long long *winPre;      // [nWin + 1] exclusive prefix: tiles << 40 | slots
    uint2 *slots;
    TileRec *tiles;
    uint32_t *counter;
    unsigned long long *stats;             // [0] similar k-mers, [1] index hits
};

__global__ void __launch_bounds__(256) kmer_lists_kernel(ListArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t ltMask = (1u << lane) - 1u;
    unsigned long long nKmers = 0, nHits = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= a.nWin) break;
        const unsigned long long p0 = a.winPre[w], p1 = a.winPre[w + 1];
        const uint32_t K = (uint32_t) ((p1 & MT_LOW40) - (p0 & MT_LOW40));
        if (K == 0) continue;
        const unsigned long long slotBase = p0 & MT_LOW40;
        const uint32_t tile0 = (uint32_t) (p0 >> 40), tileCap = (uint32_t) (p1 >> 40) - tile0;
        // query of this window
        uint32_t lo = 0, hi = a.nQueries;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.qWin[mid] <= w) lo = mid; else hi = mid; }
        const uint32_t winField = (lo << 16) | (w - a.qWin[lo]);
        const int thr = a.thr[w];
        const uint8_t *s = a.seq + w;
        const int r0 = s[0] + KA * s[1] + KA * KA * s[3];
        const int r1 = s[5] + KA * s[8] + KA * KA * s[9];
        const uint16_t *i0 = a.tb.sIdx + (size_t) r0 * N3;
        const uint16_t *i1 = a.tb.sIdx + (size_t) r1 * N3;
        const int s0top = a.tb.sScore[(size_t) r0 * N3], s1top = a.tb.sScore[(size_t) r1 * N3];
        const int vlow = max((int) (short) (thr - s1top), a.tb.vmin);
        uint32_t cnt = 0;                  // non-empty lists written so far (compacted index inside the window)
        unsigned long long acc = 0;        // hits of the tile being filled
        int prevRows = 0;
        for (int v = s0top; v >= vlow; v--) {
            const int rowsEnd = cnt_ge(a.tb, r0, v);
            const uint32_t nRows = (uint32_t) (rowsEnd - prevRows);
            const uint32_t xb = (uint32_t) prevRows;
            prevRows = rowsEnd;
            if (nRows == 0) continue;
            const uint32_t C = (uint32_t) cnt_ge(a.tb, r1, (int) (short) (thr - v));
            const uint32_t total = nRows * C;
            constexpr int U = 4;
            for (uint32_t e0 = 0; e0 < total; e0 += 32 * U) {
                uint2 ol[U]; bool ne[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t e = e0 + (uint32_t) u * 32u + (uint32_t) lane;
                    ne[u] = false; ol[u] = make_uint2(0, 0);
                    if (e < total) {
                        const uint32_t dx = e / C, c = e - dx * C;
                        const uint32_t kmA = __ldg(i0 + xb + dx), colB = __ldg(a.tb.colPerm + __ldg(i1 + c));
                        const uint32_t b = kmA * (uint32_t) N3 + colB;
                        bool occ = true;
                        if (a.useBits) occ = (__ldg(a.kbits + (b >> 5)) >> (b & 31u)) & 1u;
                        if (occ) { ol[u] = __ldg(a.klist + b); ne[u] = (ol[u].x & 0x80000000u) || ol[u].y; }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t bal = __ballot_sync(0xffffffffu, ne[u]);
                    if (bal == 0) continue;
                    const uint32_t idx = cnt + (uint32_t) __popc(bal & ltMask);
                    if (ne[u]) a.slots[slotBase + idx] = ol[u];
                    const uint32_t h = ne[u] ? ((ol[u].x & 0x80000000u) ? 1u : ol[u].y) : 0u;   // inline singleton: x = flag | target, y = 1 | pos << 16
                    const uint32_t boundary = (cnt / a.S + 1u) * a.S;
                    const bool inLo = idx < boundary;
                    const uint32_t sumLo = __reduce_add_sync(0xffffffffu, inLo ? h : 0u);
                    const uint32_t sumAll = __reduce_add_sync(0xffffffffu, h);
                    acc += sumLo;
                    const uint32_t n = (uint32_t) __popc(bal);
                    if (cnt + n >= boundary) {
                        if (lane == 0) { TileRec r; r.win = winField; r.nslots = a.S; r.hits = acc; a.tiles[tile0 + cnt / a.S] = r; }
                        nHits += acc;
                        acc = sumAll - sumLo;
                    }
                    cnt += n;
                }
            }
        }
        nKmers += K;
        uint32_t tIdx = cnt / a.S;
        if (cnt % a.S) {
            if (lane == 0) { TileRec r; r.win = winField; r.nslots = cnt % a.S; r.hits = acc; a.tiles[tile0 + tIdx] = r; }
            nHits += acc;
            tIdx++;
        }
        for (uint32_t t = tIdx + (uint32_t) lane; t < tileCap; t += 32) { TileRec r; r.win = winField; r.nslots = 0; r.hits = 0; a.tiles[tile0 + t] = r; }
    }
    if (lane == 0) { atomicAdd(a.stats + 0, nKmers); atomicAdd(a.stats + 1, nHits); }
}

__global__ void tile_pack_kernel(const TileRec *tiles, uint32_t nTiles, unsigned long long *packed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nTiles) return;
    unsigned long long v = 0;
    if (t < nTiles) { const unsigned long long h = tiles[t].hits; v = (((h + MT_R - 1) / MT_R) << 40) | h; }
    packed[t] = v;
}

// per query of the batch: first tile, first out tile, first stream position (absolute in the batch); entry [nQueries] = totals
__global__ void query_meta_kernel(const uint32_t *qWin, uint32_t nQueries, const unsigned long long *winPre, const unsigned long long *tilePre,
                                  uint32_t *qTile, uint32_t *qOut, unsigned long long *qPos) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nQueries) return;
    const uint32_t t = (uint32_t) (winPre[qWin[q]] >> 40);
    qTile[q] = t; qOut[q] = (uint32_t) (tilePre[t] >> 40); qPos[q] = tilePre[t] & MT_LOW40;
}

struct ChunkArgs {
    const uint32_t *qWin; const unsigned long long *winPre; const unsigned long long *tilePre; const TileRec *tiles;
    const uint2 *slots; const uint32_t *qTile; const unsigned long long *qPos;
    uint32_t nQueries; uint32_t S;
    unsigned long long maxDbMatches;       // 2 * max(1e6, N) (QueryMatcher.cpp:45)
    uint32_t *qinfo;                       // chunks | bit31 = the reference returns nothing for this query
    uint32_t *bounds;                      // [nQueries][PF_MAXCHUNK + 1] first stream position (query-relative) of every chunk
    uint8_t *chunkOrd;                     // [nQueries][PF_MAXCHUNK + 2] chunk_order() of the query's chunks: rank | (direction > 0) << 7
    uint32_t *errFlag;
    unsigned long long *stats;             // [2] overflow flushes
};

// Before the list of a k-mer is appended the reference checks hitsInChunk + listLen >= maxDbMatches and, if so, flushes the chunk
// (QueryMatcher.cpp:305-345).  With S(k) the stream position of list k and p the start of the chunk this is S(k+1) >= p + max:
// the first such list is the one that contains position X = p + max - 1, and the next chunk starts at S(k).
__global__ void chunk_bounds_kernel(ChunkArgs a) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= a.nQueries) return;
    const unsigned long long base = a.qPos[q], H = a.qPos[q + 1] - base;
    const uint32_t t0 = a.qTile[q], t1 = a.qTile[q + 1];
    uint32_t *bd = a.bounds + (size_t) q * (PF_MAXCHUNK + 1);
    uint32_t nChunks = 1; bool abandoned = false;
    unsigned long long p = 0;
    if (lane == 0) bd[0] = 0;
    while (p + a.maxDbMatches <= H) {
        const unsigned long long X = base + p + a.maxDbMatches - 1ull;
        uint32_t lo = t0, hi = t1;         // largest tile whose first position is <= X
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if ((a.tilePre[mid] & MT_LOW40) <= X) lo = mid; else hi = mid; }
        const TileRec rec = a.tiles[lo];
        const uint32_t w = a.qWin[rec.win >> 16] + (rec.win & 0xffffu);
        const unsigned long long wp = a.winPre[w];
        const uint2 *sl = a.slots + (wp & MT_LOW40) + (unsigned long long) (lo - (uint32_t) (wp >> 40)) * a.S;
        unsigned long long run = a.tilePre[lo] & MT_LOW40;   // stream position of the tile's first list
        unsigned long long listStart = 0; uint32_t listLen = 0; bool found = false;
        for (uint32_t sb = 0; sb < rec.nslots && !found; sb += 32) {
            const uint32_t j = sb + (uint32_t) lane;
            uint32_t len = 0;
            if (j < rec.nslots) { const uint2 v = sl[j]; len = (v.x & 0x80000000u) ? 1u : v.y; }
            uint32_t inc = len;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            const unsigned long long st = run + (unsigned long long) (inc - len);
            const bool has = len > 0 && st <= X && X < st + len;
            const uint32_t m = __ballot_sync(0xffffffffu, has);
            if (m) {
                const int src = __ffs(m) - 1;
                listStart = __shfl_sync(0xffffffffu, st, src); listLen = __shfl_sync(0xffffffffu, len, src); found = true;
            }
            run += (unsigned long long) __shfl_sync(0xffffffffu, inc, 31);
        }
        if (!found) { if (lane == 0) atomicExch(a.errFlag, 2u); break; }
        if (listLen >= a.maxDbMatches) { abandoned = true; break; }   // QueryMatcher.cpp:341-343: the query is abandoned, no hits
        if (nChunks >= (uint32_t) PF_MAXCHUNK) { if (lane == 0) atomicExch(a.errFlag, 1u); break; }
        p = listStart - base;
        if (lane == 0) bd[nChunks] = (uint32_t) p;
        nChunks++;
    }
    if (lane == 0) {
        a.qinfo[q] = nChunks | (abandoned ? 0x80000000u : 0u);
        if (nChunks > 1) atomicAdd(a.stats + 2, (unsigned long long) (nChunks - 1));
        if (nChunks <= (uint32_t) PF_MAXCHUNK) {
            int rank[PF_MAXCHUNK + 2], dir[PF_MAXCHUNK + 2];
            chunk_order((int) nChunks, rank, dir);
            uint8_t *co = a.chunkOrd + (size_t) q * (PF_MAXCHUNK + 2);
            for (uint32_t c = 1; c <= nChunks; c++) co[c] = (uint8_t) (rank[c] | (dir[c] > 0 ? 0x80 : 0));
        }
    }
}

struct ExpandArgs {
    const TileRec *tiles; const unsigned long long *tilePre; const unsigned long long *winPre; const uint32_t *qWin;
    const unsigned long long *qPos;
    const uint2 *slots; const uint64_t *entries;
    uint32_t tileLo, nTiles;               // tiles [tileLo, tileLo + nTiles) of the batch
    uint32_t outLo;                        // first out tile of the sub-batch (pool slot = out tile - outLo)
    uint32_t S, NB;
    unsigned long long *pool;              // out tiles: MT_R entries each: target / NB | diag16 << MT_LBITS | stream position << (MT_LBITS + 16)
    uint16_t *tables;                      // per out tile: NB bucket offsets, entry count, first stream position (NB + MT_TROW uint16)
};

template <int NBBITS>
__global__ void __launch_bounds__(MT_BTHREADS, 4) expand_kernel(ExpandArgs a) {
    static_assert(MT_R == 8 * MT_BTHREADS && MT_R == 32 * MT_PPT * MT_BWARPS, "one 16-byte row of s_list per thread, MT_PPT steps per warp");
    __shared__ uint2 s_slots[MT_SMAX];
    __shared__ uint32_t s_lstart[MT_SMAX + 1];
    __shared__ __align__(16) unsigned long long s_sorted[MT_R];
    __shared__ uint16_t s_cnt[MT_BWARPS][MT_MAXNB];
    __shared__ __align__(16) uint16_t s_list[MT_R];     // position of the round -> list
    __shared__ uint32_t s_scan[MT_BWARPS], s_scanL[MT_BWARPS];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t t = a.tileLo + blockIdx.x;
    const TileRec rec = a.tiles[t];
    if (rec.nslots == 0) return;
    const uint32_t q = rec.win >> 16, wi = rec.win & 0xffffu;
    const unsigned long long wp = a.winPre[a.qWin[q] + wi];
    const uint2 *sl = a.slots + (wp & MT_LOW40) + (unsigned long long) (t - (uint32_t) (wp >> 40)) * a.S;
    const unsigned long long tp = a.tilePre[t];
    const uint32_t posBase = (uint32_t) ((tp & MT_LOW40) - a.qPos[q]);
    const uint32_t out0 = (uint32_t) (tp >> 40) - a.outLo;
    const uint32_t H = (uint32_t) rec.hits;
    // list lengths -> exclusive prefix (block scan: 4 consecutive slots per thread)
    {
        reinterpret_cast<uint4 *>(s_list)[tid] = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < MT_BWARPS * MT_MAXNB / 2; k += MT_BTHREADS) reinterpret_cast<uint32_t *>(&s_cnt[0][0])[k] = 0;
        if (tid == 0) s_carry = 0;
        uint32_t len[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = (uint32_t) tid * 4u + (uint32_t) k;
            len[k] = 0;
            if (j < rec.nslots) { const uint2 v = sl[j]; s_slots[j] = v; len[k] = (v.x & 0x80000000u) ? 1u : v.y; }
            sum += len[k];
        }
        uint32_t inc = sum;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int w2 = 0; w2 < warp; w2++) wb += s_scan[w2];
        uint32_t run = wb + inc - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t j = (uint32_t) tid * 4u + (uint32_t) k; if (j < rec.nslots) s_lstart[j] = run; run += len[k]; }
        if (tid == 0) s_lstart[rec.nslots] = H;   // sentinel: larger than every valid position
        __syncthreads();
    }
    constexpr uint32_t nbMask = (1u << NBBITS) - 1u;
    // ---- position -> list of a round (s_list): the lists that start inside the round mark their first position (s_list is zero
    // before), an inclusive max-scan spreads the marks (lists start in ascending order), the list that covers the round's first position
    // is carried over from the previous round (its list at position MT_R - 1).  Three steps separated by block barriers; for the rounds
    // after the first they ride on the barriers of the preceding round's sort.
    auto list_marks = [&](uint32_t rs, bool first, uint32_t carry) {
        for (uint32_t j = (first ? 0u : carry + 1u) + (uint32_t) tid; j < rec.nslots; j += MT_BTHREADS) {
            const uint32_t rel = s_lstart[j] - rs;           // lists behind the carried one start at or after rs
            if (rel >= (uint32_t) MT_R) break;
            s_list[rel] = (uint16_t) j;
        }
    };
    auto list_scan1 = [&](uint32_t carry) -> uint32_t {      // returns the maximum in front of the thread's 8 positions, inside its warp
        const uint4 raw = reinterpret_cast<const uint4 *>(s_list)[tid];
        uint32_t inc = max(max(max(raw.x & 0xffffu, raw.x >> 16), max(raw.y & 0xffffu, raw.y >> 16)),
                           max(max(raw.z & 0xffffu, raw.z >> 16), max(raw.w & 0xffffu, raw.w >> 16)));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = max(inc, o); }
        if (lane == 31) s_scanL[warp] = inc;
        const uint32_t ex = __shfl_up_sync(0xffffffffu, inc, 1);
        return lane == 0 ? carry : max(ex, carry);           // positions in front of the round's first mark continue the carried list
    };
    auto list_scan2 = [&](uint32_t base) {
#pragma unroll
        for (int w2 = 0; w2 < MT_BWARPS - 1; w2++) if (w2 < warp) base = max(base, s_scanL[w2]);
        uint4 raw = reinterpret_cast<const uint4 *>(s_list)[tid];     // 8 consecutive positions per thread
        uint32_t v[8] = {raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16, raw.z & 0xffffu, raw.z >> 16, raw.w & 0xffffu, raw.w >> 16};
        v[0] = max(v[0], base);
#pragma unroll
        for (int i = 1; i < 8; i++) v[i] = max(v[i], v[i - 1]);
        raw.x = v[0] | (v[1] << 16); raw.y = v[2] | (v[3] << 16); raw.z = v[4] | (v[5] << 16); raw.w = v[6] | (v[7] << 16);
        reinterpret_cast<uint4 *>(s_list)[tid] = raw;
        if (tid == MT_BTHREADS - 1) s_carry = v[7];
    };
    list_marks(0u, true, 0u);
    __syncthreads();
    {
        const uint32_t base0 = list_scan1(0u);
        __syncthreads();
        list_scan2(base0);
    }
    __syncthreads();
    for (uint32_t r0 = 0, round = 0; r0 < H; r0 += MT_R, round++) {
        const uint32_t n = min((uint32_t) MT_R, H - r0);
        const bool hasNext = r0 + (uint32_t) MT_R < H;
        const uint32_t carry = s_carry;                      // list of this round's last position (rewritten behind the third barrier below)
        // ---- expand: position p = r0 + warp * 256 + k * 32 + lane; the 8 entry loads of a thread are in flight together
        uint64_t en[MT_PPT]; uint32_t valid = 0;
        const uint32_t pw = (uint32_t) warp * (32u * MT_PPT) + (uint32_t) lane;
#pragma unroll
        for (int k = 0; k < MT_PPT; k++) {
            const uint32_t pr = pw + (uint32_t) k * 32u;
            en[k] = 0;
            if (pr < n) {
                const uint32_t j = s_list[pr];
                const uint2 li = s_slots[j];
                if (li.x & 0x80000000u) en[k] = (uint64_t) (li.x & 0x3fffffu) | ((uint64_t) (li.y >> 16) << 32);
                else en[k] = __ldg(a.entries + li.x + (r0 + pr - s_lstart[j]));
                valid |= 1u << k;
            }
        }
        // ---- rank inside (warp, bucket) in stream order: steps in order, lanes in order.  The matches do not depend on the counters and
        // are issued back to back (MATCH.ANY has a long latency)
        uint32_t peers[MT_PPT];
#pragma unroll
        for (int k = 0; k < MT_PPT; k++) {
            // lanes with the same bucket: MATCH.ANY on the three low bucket bits (at most 8 distinct keys: its walk over the distinct keys
            // stays short) refined by one ballot per remaining bit.  Measured per 256 queries x 1M targets: MATCH.ANY on the full bucket
            // 34.2 ms, ballots only 28.5 ms, this hybrid 26.2 ms.
            const bool ok = (valid >> k) & 1u;
            const uint32_t b = (uint32_t) en[k] & nbMask;
            uint32_t m = __match_any_sync(0xffffffffu, ok ? (b & 7u) : 8u);
#pragma unroll
            for (int i = 3; i < NBBITS; i++) {
                // lanes whose bit i equals mine; spelled in PTX so that the bit test feeds the vote as a predicate (4 instead of 6 instructions)
                uint32_t eq;
                asm volatile("{ .reg .pred p; .reg .b32 t, bal, s;\n\t"
                             "and.b32 t, %1, %2; setp.ne.u32 p, t, 0;\n\t"
                             "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
                             "selp.b32 s, 0xffffffff, 0, p; xor.b32 bal, bal, s; not.b32 %0, bal; }"
                             : "=r"(eq) : "r"(b), "r"(1u << i));
                m &= eq;
            }
            peers[k] = m;
        }
        // (s_cnt was zeroed at the end of the previous round / in the prologue)
        uint32_t lr[MT_PPT / 4] = {0};   // rank of every position inside its (warp, bucket), one byte each (< 256)
#pragma unroll
        for (int k = 0; k < MT_PPT; k++) {
            const uint32_t b = (uint32_t) en[k] & nbMask;
            const bool ok = (valid >> k) & 1u;
            const int leader = __ffs(peers[k]) - 1;
            uint32_t old = 0;
            if (ok && lane == leader) { old = s_cnt[warp][b]; s_cnt[warp][b] = (uint16_t) (old + (uint32_t) __popc(peers[k])); }
            old = __shfl_sync(0xffffffffu, old, leader);
            lr[k >> 2] |= ((old + (uint32_t) __popc(peers[k] & ltMask)) & 0xffu) << (8 * (k & 3));
            __syncwarp();
        }
        __syncthreads();
        reinterpret_cast<uint4 *>(s_list)[tid] = make_uint4(0, 0, 0, 0);   // every warp is past its lookups: clean marks for the next round
        // ---- offsets: bucket-major, warps in order inside a bucket
        {
            uint32_t tot = 0;
            if (tid < (int) a.NB) {
#pragma unroll
                for (int w2 = 0; w2 < MT_BWARPS; w2++) { const uint32_t c = s_cnt[w2][tid]; s_cnt[w2][tid] = (uint16_t) tot; tot += c; }
            }
            uint32_t inc = tot;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            if (hasNext) list_marks(r0 + (uint32_t) MT_R, false, carry);     // s_list was cleaned in front of this barrier
            uint32_t wb = 0;
            for (int w2 = 0; w2 < warp; w2++) wb += s_scan[w2];
            const uint32_t basep = wb + inc - tot;
            if (tid < (int) a.NB) {
#pragma unroll
                for (int w2 = 0; w2 < MT_BWARPS; w2++) s_cnt[w2][tid] = (uint16_t) (s_cnt[w2][tid] + basep);
                a.tables[(size_t) (out0 + round) * (a.NB + MT_TROW) + tid] = (uint16_t) basep;
            }
            if (tid == 0) {
                uint16_t *row = a.tables + (size_t) (out0 + round) * (a.NB + MT_TROW) + a.NB;
                row[0] = (uint16_t) n; row[1] = (uint16_t) ((posBase + r0) & 0xffffu); row[2] = (uint16_t) ((posBase + r0) >> 16);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < MT_PPT; k++) {
            if ((valid >> k) & 1u) {
                const uint32_t tg = (uint32_t) en[k] & 0x3fffffu;
                const uint32_t d16 = (wi - (uint32_t) (en[k] >> 32)) & 0xffffu;
                const uint32_t p = r0 + pw + (uint32_t) k * 32u;
                s_sorted[(uint32_t) s_cnt[warp][tg & nbMask] + ((lr[k >> 2] >> (8 * (k & 3))) & 0xffu)] =
                    (uint64_t) (tg >> NBBITS) | ((uint64_t) d16 << MT_LBITS) | ((uint64_t) (posBase + p) << (MT_LBITS + 16));
            }
        }
        uint32_t listBase = 0;
        if (hasNext) listBase = list_scan1(carry);
        __syncthreads();
        // out tiles are 16 KB apart: 16-byte stores, the odd entry of the last pair is padding inside the tile
        uint4 *dst = reinterpret_cast<uint4 *>(a.pool + (size_t) (out0 + round) * MT_R);
        const uint4 *src = reinterpret_cast<const uint4 *>(s_sorted);
        for (uint32_t k = tid; k < (n + 1) / 2; k += MT_BTHREADS) dst[k] = src[k];
        for (int k = tid; k < MT_BWARPS * (int) (nbMask + 1u) / 2; k += MT_BTHREADS)      // counters of the buckets in use, for the next round
            reinterpret_cast<uint32_t *>(&s_cnt[k / (int) ((nbMask + 1u) / 2)][0])[k % (int) ((nbMask + 1u) / 2)] = 0;
        if (hasNext) list_scan2(listBase);
        __syncthreads();
    }
}

struct DiagArgs {
    const unsigned long long *pool; const uint16_t *tables;
    const uint32_t *qOut;                  // [batch queries + 1] first out tile of every query (absolute; pool slot = - outLo)
    uint32_t outLo;
    const uint32_t *qinfo; const uint32_t *bounds;
    uint32_t qLo, nQueries;                // queries [qLo, qLo + nQueries) of the batch; element keys use q - qLo
    uint32_t NB, nbShift, tBits;
    uint32_t stateStride;                  // bytes of per-target state per warp (nTargets / NB, rounded up to 16)
    Elem *elems; uint32_t *elemCount; uint32_t elemCap;
};

__device__ __forceinline__ void diag_flush(uint4 *stage, uint32_t cnt, Elem *elems, uint32_t *elemCount, uint32_t elemCap, int lane) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(elemCount, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t k = (uint32_t) lane; k < cnt; k += 32) if (base + k < elemCap) reinterpret_cast<uint4 *>(elems)[base + k] = stage[k];
}

// One 32-entry group: the hits of one target must be applied in stream (= lane) order.  Two hits of one target inside a group are
// rare: every lane reads the previous diagonal, claims its state byte with its lane number and checks the claim; only if a claim
// was lost the group is resolved with MATCH.ANY.  Returns true if the lane's hit repeats the previous 8-bit diagonal of its target.
__device__ __forceinline__ bool diag_apply(bool mine, uint32_t lt, uint8_t d8, uint8_t *state, int lane, uint32_t ltMask) {
    uint8_t prev = 0;
    bool emit = false;
    if (mine) prev = state[lt];
    __syncwarp();
    if (mine) state[lt] = (uint8_t) lane;
    __syncwarp();
    const bool lost = mine && state[lt] != (uint8_t) lane;
    __syncwarp();
    if (mine) state[lt] = d8;                              // previous diagonal, 0 initially (CacheFriendlyOperations.cpp:186,205)
    if (__ballot_sync(0xffffffffu, lost) == 0) {
        emit = mine && prev == d8;
    } else {
        const uint32_t peers = __match_any_sync(0xffffffffu, mine ? lt : (0x80000000u | (uint32_t) lane));
        const uint32_t below = peers & ltMask;
        const int before = below ? 31 - __clz((int) below) : lane;   // previous hit of the same target in this group
        const uint32_t dPrev = __shfl_sync(0xffffffffu, (uint32_t) d8, before);
        if (mine) emit = below ? (dPrev == (uint32_t) d8) : (prev == d8);
        __syncwarp();
        if (mine && lane == 31 - __clz((int) peers)) state[lt] = d8;   // the last hit of the target stays
    }
    __syncwarp();
    return emit;
}

__device__ __forceinline__ void diag_emit(bool emit, uint64_t key, uint32_t d16, uint32_t chunk, uint4 *stage, uint32_t &nStage, const DiagArgs &a,
                                          int lane, uint32_t ltMask) {
    const uint32_t em = __ballot_sync(0xffffffffu, emit);
    if (em == 0) return;
    if (emit) stage[nStage + (uint32_t) __popc(em & ltMask)] = make_uint4((uint32_t) key, (uint32_t) (key >> 32), d16, chunk);
    nStage += (uint32_t) __popc(em);
    __syncwarp();
    if (nStage >= 64) {
        diag_flush(stage, 64, a.elems, a.elemCount, a.elemCap, lane);
        __syncwarp();
        uint4 tmp = make_uint4(0, 0, 0, 0);
        if ((uint32_t) lane < nStage - 64) tmp = stage[64 + lane];
        __syncwarp();
        if ((uint32_t) lane < nStage - 64) stage[lane] = tmp;
        nStage -= 64;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MT_CWARPS * 32, 3) diag_kernel(DiagArgs a) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t item = blockIdx.x * MT_CWARPS + (uint32_t) warp;
    if (item >= a.nQueries * a.NB) return;          // no block-wide barrier below
    const uint32_t ql = item / a.NB, b = item - ql * a.NB, q = a.qLo + ql;
    uint8_t *state = dsm + (size_t) warp * (a.stateStride + MT_STAGE * 16 + MT_MAP);
    uint4 *stage = reinterpret_cast<uint4 *>(state + a.stateStride);
    uint8_t *omap = state + a.stateStride + MT_STAGE * 16;   // entry of the batch -> tile (lane) that holds it
    const uint32_t qi = a.qinfo[q];
    if (qi & 0x80000000u) return;                   // abandoned query
    const uint32_t nChunks = qi & 0x7fffffffu;
    const uint32_t *bd = a.bounds + (size_t) q * (PF_MAXCHUNK + 1);
    for (uint32_t k = (uint32_t) lane; k < a.stateStride / 16; k += 32) reinterpret_cast<uint4 *>(state)[k] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const uint64_t qKey = (uint64_t) ql << (32 + a.tBits);
    uint32_t chunk = 0, nextBound = (nChunks > 1) ? bd[1] : 0xffffffffu;
    uint32_t nStage = 0;
    const uint32_t o0 = a.qOut[q] - a.outLo, o1 = a.qOut[q + 1] - a.outLo;
    const uint32_t rowLen = a.NB + MT_TROW;
    constexpr int U = 8;
    // table row of tile ob + lane: this bucket's segment and the tile's first stream position; the next batch is fetched while the
    // current one is processed
    uint32_t nSegStart = 0, nLen = 0, nFirst = 0xffffffffu, nEnd = 0xffffffffu;
    auto fetch_tables = [&](uint32_t obf) {
        const uint32_t o = obf + (uint32_t) lane;
        nSegStart = 0; nLen = 0; nFirst = 0xffffffffu; nEnd = 0xffffffffu;
        if (o < o1) {
            const uint16_t *row = a.tables + (size_t) o * rowLen;
            nSegStart = row[b]; nLen = (uint32_t) row[b + 1] - nSegStart;
            nFirst = (uint32_t) row[a.NB + 1] | ((uint32_t) row[a.NB + 2] << 16);
            nEnd = nFirst + (uint32_t) row[a.NB];          // out tiles of a query are consecutive in the stream
        }
    };
    fetch_tables(o0);
    for (uint32_t ob = o0; ob < o1;) {
        uint32_t segStart = nSegStart, len = nLen;
        const uint32_t tileEnd = nEnd;
        // batch = up to 32 tiles whose entries of this bucket fit the offset map; normally all 32
        uint32_t nTiles = min(32u, o1 - ob);
        uint32_t inc = len;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
        if (__shfl_sync(0xffffffffu, inc, (int) nTiles - 1) > (uint32_t) MT_MAP) {
            const uint32_t fit = (uint32_t) __popc(__ballot_sync(0xffffffffu, inc <= (uint32_t) MT_MAP));   // prefix sums are monotone
            nTiles = max(1u, min(nTiles, fit));
        }
        if ((uint32_t) lane >= nTiles) len = 0;
        const uint32_t pre = ((uint32_t) lane < nTiles) ? inc - len : 0u;
        const uint32_t T = __shfl_sync(0xffffffffu, inc, (int) nTiles - 1);
        const uint32_t batchEnd = __shfl_sync(0xffffffffu, tileEnd, (int) nTiles - 1);      // first stream position behind the batch
        const uint32_t obNext = ob + nTiles;
        const unsigned long long *tile0 = a.pool + (size_t) ob * MT_R;
        if (nTiles == 32 || obNext >= o1) fetch_tables(obNext);     // (a shortened batch refetches after processing)
        const bool refetch = !(nTiles == 32 || obNext >= o1);
        if (T > (uint32_t) MT_MAP) {
            // one tile alone exceeds the map (extremely skewed target ids): walk its segment directly
            for (uint32_t p0 = 0; p0 < T; p0 += 32) {
                const uint32_t idx = p0 + (uint32_t) lane;
                const bool val = idx < T;
                const uint32_t s0 = __shfl_sync(0xffffffffu, segStart, 0);
                const unsigned long long e = val ? __ldg(tile0 + s0 + idx) : 0ull;
                const uint32_t lt = (uint32_t) e & ((1u << MT_LBITS) - 1u), d16 = (uint32_t) (e >> MT_LBITS) & 0xffffu, pos = (uint32_t) (e >> (MT_LBITS + 16));
                const uint32_t tg = (lt << a.nbShift) | b;
                uint32_t pending = __ballot_sync(0xffffffffu, val);
                for (;;) {
                    const uint32_t over = __ballot_sync(0xffffffffu, ((pending >> lane) & 1u) && pos >= nextBound);
                    const bool mine = ((pending >> lane) & 1u) && !((over >> lane) & 1u);
                    const bool emit = diag_apply(mine, lt, (uint8_t) d16, state, lane, ltMask);
                    diag_emit(emit, qKey | ((uint64_t) tg << 32) | (uint64_t) pos, d16, chunk, stage, nStage, a, lane, ltMask);
                    pending = over;
                    if (over == 0) break;
                    for (uint32_t k = (uint32_t) lane; k < a.stateStride / 16; k += 32) reinterpret_cast<uint4 *>(state)[k] = make_uint4(0, 0, 0, 0);
                    __syncwarp();
                    chunk++;
                    nextBound = (chunk + 1 < nChunks) ? bd[chunk + 1] : 0xffffffffu;
                }
            }
        } else if (T > 0) {
            // ---- tile map: entry i of the batch lives at tile0[delta(omap[i]) + i], delta = lane * MT_R + segStart - pre of that tile
            const uint32_t delta = (uint32_t) lane * (uint32_t) MT_R + segStart - pre;
            const uint32_t maxLen = __reduce_max_sync(0xffffffffu, len);
            for (uint32_t j = 0; j < maxLen; j++) if (j < len) omap[pre + j] = (uint8_t) lane;
            __syncwarp();
            // does a flush position of the reference fall into this batch?  (QueryMatcher.cpp:335-337; rare: every ~2M hits)
            const bool boundary = nextBound < batchEnd;
            for (uint32_t p0 = 0; p0 < T; p0 += 32 * U) {
                unsigned long long e[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t idx = p0 + (uint32_t) u * 32u + (uint32_t) lane;
                    e[u] = 0;
                    const uint32_t dl = __shfl_sync(0xffffffffu, delta, idx < T ? (int) omap[idx] : 0);
                    if (idx < T) e[u] = __ldg(tile0 + (dl + idx));
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t g0 = p0 + (uint32_t) u * 32u;
                    if (g0 >= T) break;
                    const bool val = g0 + (uint32_t) lane < T;
                    const uint32_t lo32 = (uint32_t) e[u];
                    const uint32_t lt = lo32 & ((1u << MT_LBITS) - 1u), pos = (uint32_t) (e[u] >> (MT_LBITS + 16));
                    const uint32_t d16 = (lo32 >> MT_LBITS) & 0xffffu, tg = (lt << a.nbShift) | b;
                    if (!boundary) {
                        const bool emit = diag_apply(val, lt, (uint8_t) d16, state, lane, ltMask);
                        diag_emit(emit, qKey | ((uint64_t) tg << 32) | (uint64_t) pos, d16, chunk, stage, nStage, a, lane, ltMask);
                    } else {
                        // overflow chunks: hits at or after the next flush position see a zeroed state
                        uint32_t pending = __ballot_sync(0xffffffffu, val);
                        for (;;) {
                            const uint32_t over = __ballot_sync(0xffffffffu, ((pending >> lane) & 1u) && pos >= nextBound);
                            const bool mine = ((pending >> lane) & 1u) && !((over >> lane) & 1u);
                            const bool emit = diag_apply(mine, lt, (uint8_t) d16, state, lane, ltMask);
                            diag_emit(emit, qKey | ((uint64_t) tg << 32) | (uint64_t) pos, d16, chunk, stage, nStage, a, lane, ltMask);
                            pending = over;
                            if (over == 0) break;
                            for (uint32_t k = (uint32_t) lane; k < a.stateStride / 16; k += 32) reinterpret_cast<uint4 *>(state)[k] = make_uint4(0, 0, 0, 0);
                            __syncwarp();
                            chunk++;
                            nextBound = (chunk + 1 < nChunks) ? bd[chunk + 1] : 0xffffffffu;
                        }
                    }
                }
            }
            __syncwarp();
        }
        ob = obNext;
        if (refetch) fetch_tables(ob);
    }
    if (nStage) diag_flush(stage, nStage, a.elems, a.elemCount, a.elemCap, lane);
}

}  // namespace prefilter
}  // namespace fsb
