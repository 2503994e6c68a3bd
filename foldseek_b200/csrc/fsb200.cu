// libfsb200.so — C ABI implementation (include/fsb200.h). Host orchestration + kernel launches.
#include "../../include/fsb200.h"
#include "host_core.h"
#include "sw_kernels.cuh"
#include <new>
#include <omp.h>
#include "prefilter.cuh"
#include "traceback.cuh"
#include "structscore.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>
#include <string>
#include <vector>
#include <time.h>

using namespace fsb;

namespace {

template <typename T>
struct DevBuf {   // grow-only device buffer
    T *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;   // a lane's view of its parent's array (fsb200_create_lane): never freed here
    cudaError_t reserve(size_t n) {
        if (n <= cap && !borrowed) return cudaSuccess;
        if (p && !borrowed) cudaFree(p);
        p = nullptr; cap = 0; borrowed = false;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void borrow(const DevBuf &o) { release(); p = o.p; cap = o.cap; borrowed = true; }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
};

// Host array in page-locked memory (cudaMemcpyAsync from pageable vectors is staged through a driver bounce buffer:
// ~3 ms per 16 MB each way at 1e6 pairs per step). Grow-only, geometric; only what the SW driver needs of std::vector.
template <typename T>
struct PinnedVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    ~PinnedVec() { release(p, pinned); }
    static void release(T *q, bool wasPinned) { if (q) { if (wasPinned) cudaFreeHost(q); else free(q); } }
    void reserve(size_t want) {
        if (want <= cap) return;
        const size_t c = std::max<size_t>(want, cap + cap / 2 + 16);
        T *q = nullptr;
        bool pin = cudaMallocHost(&q, c * sizeof(T)) == cudaSuccess;
        if (!pin) { cudaGetLastError(); q = static_cast<T *>(malloc(c * sizeof(T))); }   // pageable memory still works, only slower
        if (q == nullptr) { fprintf(stderr, "Error: out of host memory\n"); abort(); }
        if (p) memcpy(q, p, n * sizeof(T));
        release(p, pinned);
        p = q; cap = c; pinned = pin;
    }
    void resize(size_t m) { reserve(m); n = m; }
    void assign(size_t m, const T &v) { resize(m); for (size_t i = 0; i < m; i++) p[i] = v; }
    void push_back(const T &v) { if (n == cap) reserve(n + 1); p[n++] = v; }
    void clear() { n = 0; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    void swap(PinnedVec &o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); std::swap(pinned, o.pinned); }
};

struct QueryHost {
    uint32_t len;
    uint64_t off;            // into h_qAA/h_q3Di (forward order)
    double mu, lambda;       // predictMuLambda .first/.second
    int minScore;            // smallest score with evalue <= thr
    bool packedOk;           // combined int8 profile representable, L <= 384
};

}  // namespace

struct fsb200_ctx {
    fsb200_config cfg;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int smCount = 148;
    int8_t tinyAA[SW_ALPH * SW_ALPH], tiny3Di[SW_ALPH * SW_ALPH];   // tinySubMat of this context (structurealign.cpp:272-281)
    ~fsb200_ctx() {   // also the clean-up of a half-built context on the error paths of fsb200_create
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }

    host::ScoreMatrix mAA, m3Di, mKmer, mUngapped;
    host::EvalueModel nn;
    bool hasNN = false;

    // target
    uint32_t nT = 0;
    uint64_t tResidues = 0, aaDbResidues = 0;
    size_t tBytes = 0;
    std::vector<uint64_t> h_tOff;
    std::vector<uint32_t> h_tLen;
    DevBuf<uint8_t> d_tAA, d_t3Di;
    DevBuf<uint64_t> d_tOff;
    DevBuf<uint32_t> d_tLen;

    // queries: profile 2q = forward, 2q+1 = reversed query
    uint32_t nQ = 0;
    std::vector<QueryHost> q;
    std::vector<uint8_t> h_pAA, h_p3Di;      // profile sequences (fwd, rev interleaved per query)
    std::vector<int8_t> h_pBiasAA, h_pBiasSS;
    std::vector<SwProfileDesc> h_prof;
    DevBuf<uint8_t> d_pAA, d_p3Di;
    DevBuf<int8_t> d_pBiasAA, d_pBiasSS;
    DevBuf<SwProfileDesc> d_prof;

    // align
    std::vector<uint64_t> candOff;
    std::vector<uint32_t> candTarget;
    std::vector<uint32_t> candQuery;
    std::vector<fsb200_aln> results;
    PinnedVec<SwPair> h_pairs;
    std::vector<SwTask> h_tasks;
    std::vector<SwItem> h_items[SW_MAXR / 2 + 1];   // by R/2
    std::vector<uint32_t> h_generic;
    PinnedVec<SwOut> h_out, h_fwd;
    std::vector<uint32_t> h_idx;
    DevBuf<SwPair> d_pairs;
    DevBuf<SwTask> d_tasks;
    DevBuf<SwItem> d_items;
    DevBuf<SwOut> d_out;
    DevBuf<uint32_t> d_generic;
    DevBuf<int32_t> d_scratch;
    DevBuf<uint32_t> d_counter;
    uint64_t alnStats[8] = {0};
    PinnedVec<fsb200_hit> pfHits;            // page-locked: the hit lists come back with one cudaMemcpyAsync per prefilter call
    std::vector<uint32_t> pfCounts;
    std::vector<uint32_t> identity;          // per query: its own target id (query DB == target DB) or 0xFFFFFFFF; empty = none
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;

    prefilter::State pf;
    // lanes: contexts on the same GPU that share this context's staged target, C-alpha and k-mer index (fsb200_create_lane)
    std::shared_ptr<int> laneGroup;          // number of contexts in the group (this one included); shared by all of them
    bool isLane = false;
    int hostThreads = 0;                     // OpenMP team of this context's host-side loops (0 = the runtime's default)

    // traceback of accepted pairs
    DevBuf<int8_t> d_tbTabs;
    DevBuf<tb::TbPair> d_tbPairs;
    DevBuf<uint32_t> d_tbOrder;
    DevBuf<int16_t> d_tbState;
    DevBuf<uint8_t> d_tbTrace;
    DevBuf<uint4> d_tbBlocks;
    DevBuf<char> d_bt;
    DevBuf<tb::TbOut> d_tbOut;
    // traceback / structure scores live in the context as compact lists over the accepted pairs (a handful per query); the dense
    // per-candidate arrays of the ABI are materialised by the download calls
    std::vector<uint32_t> accIdx;            // accepted pairs (fsb200_aln.status == 0), ascending candidate index
    std::vector<fsb200_bt> accMeta;
    std::vector<fsb200_sscore> accSs;
    bool btValid = false, ssValid = false;
    std::vector<char> btBuf;
    uint64_t btStats[8] = {0};
    uint64_t btDeviceBytes = 0;      // backtraces of the last fsb200_backtrace_run are still in d_bt

    // C-alpha coordinates (x / y / z arrays, tight residue offsets) and structure scores
    bool haveTCa = false, haveQCa = false;
    std::vector<uint64_t> h_tCaOff, h_qCaOff;
    DevBuf<float> d_tCa, d_qCa, d_qNorm;
    DevBuf<uint64_t> d_qCaOff;
    DevBuf<uint32_t> d_qCaLen;
    uint64_t tCaTotal = 0, qCaTotal = 0;
    DevBuf<ss::SsPair> d_ssPairs;
    DevBuf<ss::SsOut> d_ssOut;
    DevBuf<float> d_ssScratch;
    uint64_t ssStats[8] = {0};
};

namespace {

// large host-to-host copies of result arrays (tens of MB per step) with all host threads
void par_memcpy(void *dst, const void *src, size_t bytes) {
    const size_t chunk = (size_t) 1 << 20;
    const long n = (long) ((bytes + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) memcpy((char *) dst + (size_t) i * chunk, (const char *) src + (size_t) i * chunk, std::min(chunk, bytes - (size_t) i * chunk));
}

double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int fail(fsb200_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

// every entry point that touches the device or runs host loops: the calling thread works on this context's GPU with this
// context's OpenMP team (the team size is an ICV of the calling thread: lanes driven by different host threads do not interfere)
void enter(const fsb200_ctx *c) {
    cudaSetDevice(c->cfg.device);
    if (c->hostThreads > 0) omp_set_num_threads(c->hostThreads);
}

#define CUDA_OK(c, call)                                                                                   \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) return fail(c, FSB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// c_matAA / c_mat3Di are symbols of the module, i.e. shared by all contexts of a device: every align_run re-sends its own matrices
// in stream order (two contexts with different --alignment-type on one device would otherwise score with each other's matrix)
static cudaError_t upload_matrices(fsb200_ctx *c) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_matAA, c->tinyAA, sizeof(c->tinyAA), 0, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_mat3Di, c->tiny3Di, sizeof(c->tiny3Di), 0, cudaMemcpyHostToDevice, c->stream);
}

template <int R>
cudaError_t launch_packed(fsb200_ctx *c, const SwArgs &a) {
    constexpr int K = (R + 7) / 8;
    const size_t smem = (size_t) SW_COMBOS * 128 * K;
    cudaError_t e = cudaFuncSetAttribute(sw_packed_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = std::min<unsigned>(a.nItems, (unsigned) c->smCount);
    sw_packed_kernel<R><<<grid, SW_BLOCK, smem, c->stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_packed_r(fsb200_ctx *c, int R, const SwArgs &a) {
    switch (R) {
        case 2: return launch_packed<2>(c, a);
        case 4: return launch_packed<4>(c, a);
        case 6: return launch_packed<6>(c, a);
        case 8: return launch_packed<8>(c, a);
        case 10: return launch_packed<10>(c, a);
        case 12: return launch_packed<12>(c, a);
        case 14: return launch_packed<14>(c, a);
        case 16: return launch_packed<16>(c, a);
        case 18: return launch_packed<18>(c, a);
        case 20: return launch_packed<20>(c, a);
        case 22: return launch_packed<22>(c, a);
        case 24: return launch_packed<24>(c, a);
        default: return cudaErrorInvalidValue;
    }
}

int seg_r(uint32_t len) {   // rows per thread of the packed kernel for a query of this length
    int seg = ((int) len + SW_GROUP - 1) / SW_GROUP;
    return (seg + 1) & ~1;
}

// Build tasks/items for the pairs listed in `idx` (all pairs in h_pairs), grouped by profile.
// endRow marks the items as end-row mode.  Appends to h_tasks / h_items / h_generic.
void build_tasks(fsb200_ctx *c, std::vector<uint32_t> &idx, bool endRow) {
    // group by profile (callers pass the pairs in query order, so this is normally already the case)
    bool grouped = true;
    for (size_t k = 1; k < idx.size(); k++) if (c->h_pairs[idx[k]].profile < c->h_pairs[idx[k - 1]].profile) { grouped = false; break; }
    if (!grouped) std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return c->h_pairs[x].profile < c->h_pairs[y].profile; });
    std::vector<size_t> starts;
    {   // group boundaries, found by all host threads (1e6 pairs per 1k queries)
        const long n = (long) idx.size();
        std::vector<std::vector<size_t>> part(omp_get_max_threads());
#pragma omp parallel
        {
            std::vector<size_t> &mine = part[omp_get_thread_num()];
#pragma omp for schedule(static)
            for (long k = 0; k < n; k++) if (k == 0 || c->h_pairs[idx[k]].profile != c->h_pairs[idx[k - 1]].profile) mine.push_back((size_t) k);
        }
        for (auto &v : part) starts.insert(starts.end(), v.begin(), v.end());
    }
    starts.push_back(idx.size());
    // inside a group: longest targets first, so that the 4 alignments of a warp have similar lengths
    const long nGroups = (long) starts.size() - 1;
#pragma omp parallel for schedule(dynamic, 8)
    for (long g = 0; g < nGroups; g++) {
        bool sorted = true;   // equal-length DBs (and pre-sorted lists) need no sort
        for (size_t k = starts[g] + 1; k < starts[g + 1]; k++) {
            const SwPair &a = c->h_pairs[idx[k - 1]], &b = c->h_pairs[idx[k]];
            if (a.tlen < b.tlen || (a.tlen == b.tlen && idx[k - 1] > idx[k])) { sorted = false; break; }
        }
        if (!sorted) std::sort(idx.begin() + starts[g], idx.begin() + starts[g + 1], [&](uint32_t x, uint32_t y) {
            const SwPair &a = c->h_pairs[x], &b = c->h_pairs[y];
            if (a.tlen != b.tlen) return a.tlen > b.tlen;
            return x < y;
        });
    }
    // work items of the persistent SW kernel: at most 64 tasks (256 alignments share one profile build); with few queries
    // (a rank's shard of a strong-scaled batch) smaller items keep the last wave of the 148 CTAs short
    const uint32_t ITEM_TASKS = (uint32_t) std::min<size_t>(64, std::max<size_t>(8, (idx.size() / 4 + 1) / ((size_t) c->smCount * 8) + 1));
    // task ranges per group first (serial, one entry per query), then the tasks themselves in parallel
    std::vector<uint32_t> taskBase(nGroups + 1, 0);
    const uint32_t task0 = (uint32_t) c->h_tasks.size();
    for (long g = 0; g < nGroups; g++) {
        const QueryHost &qh = c->q[c->h_pairs[idx[starts[g]]].profile >> 1];
        taskBase[g + 1] = taskBase[g] + (qh.packedOk ? (uint32_t) ((starts[g + 1] - starts[g] + 3) / 4) : 0u);
    }
    c->h_tasks.resize((size_t) task0 + taskBase[nGroups]);
#pragma omp parallel for schedule(static, 16)
    for (long g = 0; g < nGroups; g++) {
        const size_t i = starts[g], j = starts[g + 1];
        if (!c->q[c->h_pairs[idx[i]].profile >> 1].packedOk) continue;
        SwTask *out = c->h_tasks.data() + task0 + taskBase[g];
        for (size_t k = i; k < j; k += 4) {
            SwTask t;
            for (int m = 0; m < 4; m++) t.pair[m] = (k + m < j) ? idx[k + m] : 0xFFFFFFFFu;
            *out++ = t;
        }
    }
    for (long g = 0; g < nGroups; g++) {
        const size_t i = starts[g], j = starts[g + 1];
        const uint32_t prof = c->h_pairs[idx[i]].profile;
        const QueryHost &qh = c->q[prof >> 1];
        if (!qh.packedOk) {
            for (size_t k = i; k < j; k++) c->h_generic.push_back(idx[k]);
            continue;
        }
        const int R = seg_r(qh.len);
        const uint32_t taskBegin = task0 + taskBase[g], taskEnd = task0 + taskBase[g + 1];
        for (uint32_t tb = taskBegin; tb < taskEnd; tb += ITEM_TASKS) {
            SwItem it; it.profile = prof; it.taskBegin = tb; it.taskEnd = std::min(taskEnd, tb + ITEM_TASKS);
            it.pad = endRow ? 1u : 0u;
            c->h_items[R / 2].push_back(it);
        }
    }
}

// Uploads h_pairs/h_tasks/h_items, launches every R class + the generic kernel, downloads h_out.
int run_sw(fsb200_ctx *c) {
    const size_t nPairs = c->h_pairs.size();
    c->h_out.resize(nPairs);   // overwritten by the download of d_out (zero-initialised on the device)
    if (nPairs == 0) return FSB200_OK;
    CUDA_OK(c, c->d_pairs.reserve(nPairs));
    CUDA_OK(c, c->d_out.reserve(nPairs));
    CUDA_OK(c, c->d_tasks.reserve(std::max<size_t>(1, c->h_tasks.size())));
    size_t nItems = 0;
    for (auto &v : c->h_items) nItems += v.size();
    CUDA_OK(c, c->d_items.reserve(std::max<size_t>(1, nItems)));
    CUDA_OK(c, c->d_counter.reserve(64));
    CUDA_OK(c, cudaMemcpyAsync(c->d_pairs.p, c->h_pairs.data(), nPairs * sizeof(SwPair), cudaMemcpyHostToDevice, c->stream));
    if (!c->h_tasks.empty())
        CUDA_OK(c, cudaMemcpyAsync(c->d_tasks.p, c->h_tasks.data(), c->h_tasks.size() * sizeof(SwTask), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemsetAsync(c->d_counter.p, 0, 64 * sizeof(uint32_t), c->stream));
    CUDA_OK(c, cudaMemsetAsync(c->d_out.p, 0, nPairs * sizeof(SwOut), c->stream));
    SwArgs a;
    a.tAA = c->d_tAA.p; a.t3Di = c->d_t3Di.p; a.tOff = c->d_tOff.p;
    a.qAA = c->d_pAA.p; a.q3Di = c->d_p3Di.p; a.biasAA = c->d_pBiasAA.p; a.biasSS = c->d_pBiasSS.p;
    a.profiles = c->d_prof.p; a.pairs = c->d_pairs.p; a.tasks = c->d_tasks.p;
    a.out = c->d_out.p; a.gapOpen = c->cfg.gap_open; a.gapExtend = c->cfg.gap_extend;
    size_t itemOff = 0;
    int cls = 0;
    for (int r2 = 1; r2 <= SW_MAXR / 2; r2++) {
        auto &v = c->h_items[r2];
        if (v.empty()) continue;
        CUDA_OK(c, cudaMemcpyAsync(c->d_items.p + itemOff, v.data(), v.size() * sizeof(SwItem), cudaMemcpyHostToDevice, c->stream));
        a.items = c->d_items.p + itemOff;
        a.nItems = (uint32_t) v.size();
        a.itemCounter = c->d_counter.p + (cls % 32);
        CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
        CUDA_OK(c, launch_packed_r(c, r2 * 2, a));
        CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
        CUDA_OK(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        CUDA_OK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->alnStats[6] += (uint64_t) (ms * 1e6);
        c->alnStats[7] += 1;
        c->alnStats[4] += 1;
        itemOff += v.size();
        cls++;
    }
    auto run_generic = [&](const std::vector<uint32_t> &list) -> int {
        if (list.empty()) return FSB200_OK;
        uint32_t maxL = 1;
        for (uint32_t p : list) maxL = std::max(maxL, c->q[c->h_pairs[p].profile >> 1].len);
        const uint32_t stride = 32 * ((maxL + 31) / 32);
        const unsigned blocks = std::min<unsigned>((unsigned) ((list.size() + 3) / 4), (unsigned) c->smCount * 4);
        CUDA_OK(c, c->d_generic.reserve(list.size()));
        CUDA_OK(c, c->d_scratch.reserve((size_t) blocks * 4 * 2 * stride));
        CUDA_OK(c, cudaMemcpyAsync(c->d_generic.p, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(c, cudaMemsetAsync(c->d_counter.p + 40, 0, sizeof(uint32_t), c->stream));
        SwGenericArgs ga;
        ga.base = a; ga.pairList = c->d_generic.p; ga.nPairs = (uint32_t) list.size();
        ga.scratch = c->d_scratch.p; ga.scratchStride = stride; ga.pairCounter = c->d_counter.p + 40;
        sw_generic_kernel<<<blocks, 128, 0, c->stream>>>(ga);
        CUDA_OK(c, cudaGetLastError());
        c->alnStats[4] += 1;
        c->alnStats[5] += list.size();
        return FSB200_OK;
    };
    int rc = run_generic(c->h_generic);
    if (rc != FSB200_OK) return rc;
    CUDA_OK(c, cudaMemcpyAsync(c->h_out.data(), c->d_out.p, nPairs * sizeof(SwOut), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    // redo alignments whose packed int16 score may have wrapped
    std::vector<uint32_t> redo;
    for (size_t i = 0; i < nPairs; i++) if (c->h_out[i].flags & 1u) redo.push_back((uint32_t) i);
    if (!redo.empty()) {
        rc = run_generic(redo);
        if (rc != FSB200_OK) return rc;
        CUDA_OK(c, cudaMemcpyAsync(c->h_out.data(), c->d_out.p, nPairs * sizeof(SwOut), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
    }
    return FSB200_OK;
}

// results of earlier batches must not be indexed with the ids of new inputs
void invalidate_results(fsb200_ctx *c, bool targetChanged) {
    c->results.clear(); c->candOff.clear(); c->candTarget.clear(); c->candQuery.clear();
    c->accIdx.clear(); c->accMeta.clear(); c->accSs.clear(); c->btValid = c->ssValid = false; c->btBuf.clear(); c->btDeviceBytes = 0;
    c->pfHits.clear(); c->pfCounts.clear(); c->identity.clear();
    if (targetChanged) c->haveTCa = false; else c->haveQCa = false;
}

void clear_work(fsb200_ctx *c) {
    c->h_pairs.clear(); c->h_tasks.clear(); c->h_generic.clear();
    for (auto &v : c->h_items) v.clear();
}

}  // namespace

extern "C" {

#ifndef FSB200_BUILD_ID
#define FSB200_BUILD_ID "unknown"
#endif
static const char fsb200_build_marker[] = "FSB200_BUILD_ID=" FSB200_BUILD_ID;   // also findable in the file without loading it
const char *fsb200_build_id(void) { return fsb200_build_marker + 16; }

void fsb200_default_config(fsb200_config *cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->gap_open = 10; cfg->gap_extend = 1; cfg->comp_bias_corr = 1;
    cfg->comp_bias_scale_align = 0.5f; cfg->comp_bias_scale_prefilter = 0.15f;
    cfg->sensitivity = 9.5f; cfg->kmer_size = 6; cfg->max_seqs = 1000; cfg->min_diag_score = 30;
    cfg->evalue_thr = 10.0; cfg->aa_factor = 1.4f; cfg->diag_bin_count = 0;
}

int fsb200_create(const fsb200_config *cfg, fsb200_ctx **out) {
    if (cfg == NULL || out == NULL) return FSB200_E_ARG;
    *out = NULL;
    std::unique_ptr<fsb200_ctx> c(new fsb200_ctx());
    c->cfg = *cfg;
    if (!c->m3Di.parse(cfg->mat3di_text, 2.1, 0.0) || !c->mAA.parse(cfg->mataa_text, cfg->aa_factor, 0.0) ||
        !c->mKmer.parse(cfg->mat3di_text, 8.0, -0.2f) || !c->mUngapped.parse(cfg->mat3di_text, 2.0, -0.2f))
        return FSB200_E_MATRIX;
    if (cfg->evalue_nn != NULL) {
        if (!c->nn.load(cfg->evalue_nn, cfg->evalue_nn_len)) return FSB200_E_ARG;
        c->hasNN = true;
    }
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= cfg->device) return FSB200_E_CUDA;   // no CPU fallback
    if (cudaSetDevice(cfg->device) != cudaSuccess) return FSB200_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return FSB200_E_CUDA;
    c->smCount = prop.multiProcessorCount;
    if (cudaStreamCreate(&c->stream) != cudaSuccess) return FSB200_E_CUDA;
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    for (int i = 0; i < SW_ALPH; i++) for (int j = 0; j < SW_ALPH; j++) {
        c->tinyAA[i * SW_ALPH + j] = (int8_t) c->mAA.s[i][j];     // tinySubMat is int8 (structurealign.cpp:272-281)
        c->tiny3Di[i * SW_ALPH + j] = (int8_t) c->m3Di.s[i][j];
    }
    if (upload_matrices(c.get()) != cudaSuccess) return FSB200_E_CUDA;
    *out = c.release();
    return FSB200_OK;
}

void fsb200_destroy(fsb200_ctx *c) {
    if (c == NULL) return;
    enter(c);
    c->d_tAA.release(); c->d_t3Di.release(); c->d_tOff.release(); c->d_tLen.release();
    c->d_pAA.release(); c->d_p3Di.release(); c->d_pBiasAA.release(); c->d_pBiasSS.release(); c->d_prof.release();
    c->d_pairs.release(); c->d_tasks.release(); c->d_items.release(); c->d_out.release(); c->d_generic.release();
    c->d_scratch.release(); c->d_counter.release();
    c->d_tbTabs.release(); c->d_tbPairs.release(); c->d_tbOrder.release(); c->d_tbState.release(); c->d_tbTrace.release();
    c->d_tbBlocks.release(); c->d_bt.release(); c->d_tbOut.release();
    c->d_tCa.release(); c->d_qCa.release(); c->d_qNorm.release(); c->d_qCaOff.release(); c->d_qCaLen.release();
    c->d_ssPairs.release(); c->d_ssOut.release(); c->d_ssScratch.release();
    c->pf.release();
    if (c->laneGroup && *c->laneGroup > 1) *c->laneGroup -= 1;
    if (c->tm0) { cudaEventDestroy(c->tm0); cudaEventDestroy(c->tm1); }
    // stream and events: ~fsb200_ctx
    delete c;
}

const char *fsb200_last_error(const fsb200_ctx *c) { return c ? c->err.c_str() : "null context"; }

int fsb200_get_matrix(const fsb200_ctx *c, int32_t which, int16_t *out441) {
    if (c == NULL || out441 == NULL || which < 0 || which > 3) return FSB200_E_ARG;
    const host::ScoreMatrix *m = which == 0 ? &c->mAA : which == 1 ? &c->m3Di : which == 2 ? &c->mKmer : &c->mUngapped;
    for (int i = 0; i < 21; i++) for (int j = 0; j < 21; j++) out441[i * 21 + j] = m->s[i][j];
    return FSB200_OK;
}

int fsb200_load_target(fsb200_ctx *c, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len,
                       uint32_t n, uint64_t aaDbResidues) {
    if (c == NULL || aa == NULL || tdi == NULL || off == NULL || len == NULL) return FSB200_E_ARG;
    if (!c->isLane && c->laneGroup && *c->laneGroup > 1) return fail(c, FSB200_E_STATE, "lanes share this context's target: destroy them before loading another");
    enter(c);
    // repack with each target padded to a multiple of 16 bytes (aligned, over-read safe); inputs are validated before any
    // context state changes
    std::vector<uint64_t> newOff(n);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (len[i] > 65535) return fail(c, FSB200_E_RANGE, "target longer than 65535 (Parameters.h:271 MAX_SEQ_LEN)");
        newOff[i] = total; total += ((uint64_t) len[i] + 16 + 15) & ~15ull;
    }
    std::vector<uint8_t> pa(total + 16, 20), ps(total + 16, 20);
    uint64_t res = 0;
    bool badCode = false;
#pragma omp parallel for schedule(static, 1024) reduction(+ : res) reduction(|| : badCode)
    for (uint32_t i = 0; i < n; i++) {
        memcpy(pa.data() + newOff[i], aa + off[i], len[i]);
        memcpy(ps.data() + newOff[i], tdi + off[i], len[i]);
        for (uint32_t k = 0; k < len[i]; k++) if (aa[off[i] + k] > 20 || tdi[off[i] + k] > 20) badCode = true;
        res += len[i];
    }
    if (badCode) return fail(c, FSB200_E_ARG, "residue code > 20");
    c->h_tOff.swap(newOff); c->h_tLen.assign(len, len + n);
    invalidate_results(c, true);
    c->nT = n; c->tResidues = res; c->aaDbResidues = aaDbResidues; c->tBytes = total + 16;
    CUDA_OK(c, c->d_tAA.reserve(total + 16)); CUDA_OK(c, c->d_t3Di.reserve(total + 16));
    CUDA_OK(c, c->d_tOff.reserve(std::max<uint32_t>(1, n))); CUDA_OK(c, c->d_tLen.reserve(std::max<uint32_t>(1, n)));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tAA.p, pa.data(), total + 16, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_t3Di.p, ps.data(), total + 16, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tOff.p, c->h_tOff.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tLen.p, c->h_tLen.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->nn.logDbResidues = log(static_cast<double>(aaDbResidues));
    c->pf.invalidate();
    return FSB200_OK;
}

int fsb200_set_queries(fsb200_ctx *c, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len, uint32_t nq) {
    if (c == NULL || aa == NULL || tdi == NULL || off == NULL || len == NULL) return FSB200_E_ARG;
    enter(c);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nq; i++) {
        if (len[i] == 0 || len[i] > 65535) return fail(c, FSB200_E_RANGE, "query length must be in 1..65535");
        total += 2ull * len[i];
    }
    invalidate_results(c, false);
    c->nQ = nq;
    c->q.resize(nq);
    c->h_pAA.resize(total); c->h_p3Di.resize(total); c->h_pBiasAA.resize(total); c->h_pBiasSS.resize(total);
    c->h_prof.resize(2ull * nq);
    uint64_t o = 0;
    for (uint32_t i = 0; i < nq; i++) {
        c->h_prof[2 * i] = SwProfileDesc{o, len[i], 0};
        c->h_prof[2 * i + 1] = SwProfileDesc{o + len[i], len[i], 0};
        c->q[i].len = len[i]; c->q[i].off = o;
        o += 2ull * len[i];
    }
    int maxMatAA = -128, maxMat3 = -128, minMatAA = 127, minMat3 = 127;
    for (int i = 0; i < 21; i++) for (int j = 0; j < 21; j++) {
        maxMatAA = std::max(maxMatAA, (int) (int8_t) c->mAA.s[i][j]); minMatAA = std::min(minMatAA, (int) (int8_t) c->mAA.s[i][j]);
        maxMat3 = std::max(maxMat3, (int) (int8_t) c->m3Di.s[i][j]); minMat3 = std::min(minMat3, (int) (int8_t) c->m3Di.s[i][j]);
    }
    bool bad = false;
#pragma omp parallel for schedule(dynamic, 16)
    for (uint32_t i = 0; i < nq; i++) {
        const uint32_t L = len[i];
        uint8_t *fa = c->h_pAA.data() + c->q[i].off, *fs = c->h_p3Di.data() + c->q[i].off;
        uint8_t *ra = fa + L, *rs = fs + L;
        for (uint32_t k = 0; k < L; k++) {
            const uint8_t a = aa[off[i] + k], s = tdi[off[i] + k];
            if (a > 20 || s > 20) { bad = true; }
            fa[k] = std::min<uint8_t>(a, 20); fs[k] = std::min<uint8_t>(s, 20);
            ra[L - 1 - k] = fa[k]; rs[L - 1 - k] = fs[k];
        }
        int8_t *bfa = c->h_pBiasAA.data() + c->q[i].off, *bfs = c->h_pBiasSS.data() + c->q[i].off;
        int lo = 0, hi = 0;
        if (c->cfg.comp_bias_corr) {
            std::vector<float> tmp(L);
            // ssw_init (SSW.cpp:1563-1575): both corrections use the AA matrix object, scales 1.0 and compBiasCorrectionScale;
            // the reverse profile recomputes them on the reversed sequences (structurealign.cpp:345-347).
            for (int dir = 0; dir < 2; dir++) {
                const uint8_t *sa = dir ? ra : fa, *ss = dir ? rs : fs;
                int8_t *ba = bfa + dir * L, *bs = bfs + dir * L;
                host::composition_bias(c->mAA, sa, (int) L, 1.0f, tmp.data());
                for (uint32_t k = 0; k < L; k++) ba[k] = host::round_bias(tmp[k]);
                host::composition_bias(c->mAA, ss, (int) L, c->cfg.comp_bias_scale_align, tmp.data());
                for (uint32_t k = 0; k < L; k++) bs[k] = host::round_bias(tmp[k]);
                for (uint32_t k = 0; k < L; k++) { lo = std::min(lo, (int) ba[k] + (int) bs[k]); hi = std::max(hi, (int) ba[k] + (int) bs[k]); }
            }
        } else {
            memset(bfa, 0, 2 * L); memset(bfs, 0, 2 * L);
        }
        c->q[i].packedOk = (L <= (uint32_t) (SW_GROUP * SW_MAXR)) && (maxMatAA + maxMat3 + hi <= 127) && (minMatAA + minMat3 + lo >= -127);
        c->q[i].mu = 0; c->q[i].lambda = 0; c->q[i].minScore = 0;
        if (c->hasNN) {
            c->nn.mu_lambda(fs, (int) L, c->q[i].mu, c->q[i].lambda);
            c->q[i].minScore = c->nn.min_passing_score(c->q[i].mu, c->q[i].lambda, c->cfg.evalue_thr);
        }
    }
    if (bad) return fail(c, FSB200_E_ARG, "query residue code > 20");
    CUDA_OK(c, c->d_pAA.reserve(total + 16)); CUDA_OK(c, c->d_p3Di.reserve(total + 16));
    CUDA_OK(c, c->d_pBiasAA.reserve(total + 16)); CUDA_OK(c, c->d_pBiasSS.reserve(total + 16));
    CUDA_OK(c, c->d_prof.reserve(2ull * nq + 1));
    CUDA_OK(c, cudaMemcpyAsync(c->d_pAA.p, c->h_pAA.data(), total, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_p3Di.p, c->h_p3Di.data(), total, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_pBiasAA.p, c->h_pBiasAA.data(), total, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_pBiasSS.p, c->h_pBiasSS.data(), total, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_prof.p, c->h_prof.data(), 2ull * nq * sizeof(SwProfileDesc), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return FSB200_OK;
}

int fsb200_query_info(fsb200_ctx *c, uint32_t qi, double *ml, int8_t *ba, int8_t *bs) {
    if (c == NULL || qi >= c->nQ) return FSB200_E_ARG;
    if (ml) { ml[0] = c->q[qi].mu; ml[1] = c->q[qi].lambda; }
    if (ba) memcpy(ba, c->h_pBiasAA.data() + c->q[qi].off, c->q[qi].len);
    if (bs) memcpy(bs, c->h_pBiasSS.data() + c->q[qi].off, c->q[qi].len);
    return FSB200_OK;
}

int fsb200_align_upload(fsb200_ctx *c, const uint64_t *candOff, const uint32_t *candTarget) {
    if (c == NULL || candOff == NULL) return FSB200_E_ARG;
    if (c->nT == 0 || c->nQ == 0) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    const uint64_t n = candOff[c->nQ];
    if (n > 0 && candTarget == NULL) return FSB200_E_ARG;
    c->candOff.assign(candOff, candOff + c->nQ + 1);
    c->candTarget.assign(candTarget, candTarget + n);
    c->candQuery.resize(n);
    for (uint32_t qi = 0; qi < c->nQ; qi++) {
        if (candOff[qi + 1] < candOff[qi]) return FSB200_E_ARG;
        for (uint64_t k = candOff[qi]; k < candOff[qi + 1]; k++) {
            if (candTarget[k] >= c->nT) return fail(c, FSB200_E_ARG, "candidate target id out of range");
            c->candQuery[k] = qi;
        }
    }
    return FSB200_OK;
}

int fsb200_align_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    if (c->nT == 0 || c->nQ == 0) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    if (!c->hasNN) return fail(c, FSB200_E_STATE, "E-value network not configured");
    enter(c);
    CUDA_OK(c, upload_matrices(c));
    memset(c->alnStats, 0, sizeof(c->alnStats));
    const bool tmg = getenv("FSB200_TIMING") != NULL;
    double tA = now_s();
    const size_t n = c->candTarget.size();
    c->results.resize(n);   // filled by the parallel loop below (40 B x ~1e6 pairs: a serial assign costs milliseconds)
    c->alnStats[0] = n;
    // ---- phase A: forward pass, score + dbEnd for every candidate ----
    clear_work(c);
    c->h_pairs.resize(n);
    std::vector<uint32_t> &idx = c->h_idx; idx.resize(n);
    uint64_t cellsA = 0;
#pragma omp parallel for schedule(static, 8192) reduction(+ : cellsA)
    for (size_t i = 0; i < n; i++) {
        const uint32_t t = c->candTarget[i];
        c->h_pairs[i] = SwPair{2 * c->candQuery[i], t, c->h_tLen[t], 0};
        idx[i] = (uint32_t) i;
        cellsA += (uint64_t) c->q[c->candQuery[i]].len * c->h_tLen[t];
    }
    c->alnStats[1] = cellsA;
    const double tB = now_s();
    build_tasks(c, idx, false);
    const double tC = now_s();
    int rc = run_sw(c);
    const double tD = now_s();
    if (rc != FSB200_OK) return rc;
    c->h_fwd.swap(c->h_out);   // both are members: their capacity survives across calls (no 16 MB page-fault storm per step)
    PinnedVec<SwOut> &fwd = c->h_fwd;
    // ---- gate 1 (structurealign.cpp:55-59) ----
    std::vector<uint32_t> surv;
    {
        std::vector<std::vector<uint32_t>> part(omp_get_max_threads());
#pragma omp parallel
        {
            std::vector<uint32_t> &mine = part[omp_get_thread_num()];
#pragma omp for schedule(static)
            for (size_t i = 0; i < n; i++) {
                const QueryHost &qh = c->q[c->candQuery[i]];
                // every field is written here (rejected after the forward pass: status 1, no reverse score, no qEnd)
                c->results[i] = fsb200_aln{fwd[i].score, 0, -1, fwd[i].dbEnd, 1, (int32_t) fwd[i].score, 0.0};
                if ((int) fwd[i].score >= qh.minScore) mine.push_back((uint32_t) i);
            }
        }
        for (auto &v : part) surv.insert(surv.end(), v.begin(), v.end());   // static schedule: ascending candidate order
    }
    // ---- phase B: reverse-query score (all survivors) and qEnd (survivors scored by the packed kernel) ----
    clear_work(c);
    const size_t ns = surv.size();
    std::vector<uint32_t> idxRev(ns), idxRow, rowOf(ns, 0xFFFFFFFFu);
    c->h_pairs.resize(ns);
    for (size_t k = 0; k < ns; k++) {
        const uint32_t i = surv[k];
        const uint32_t t = c->candTarget[i];
        const uint32_t qi = c->candQuery[i];
        c->h_pairs[k] = SwPair{2 * qi + 1, t, c->h_tLen[t], 0};
        idxRev[k] = (uint32_t) k;
        c->alnStats[2] += (uint64_t) c->q[qi].len * c->h_tLen[t];
    }
    for (size_t k = 0; k < ns; k++) {
        const uint32_t i = surv[k];
        if ((fwd[i].flags & 4u) || fwd[i].score == 0) continue;   // the generic kernel already reported the row
        const uint32_t t = c->candTarget[i];
        const uint32_t qi = c->candQuery[i];
        const uint32_t tl = (uint32_t) (fwd[i].dbEnd + 1);
        rowOf[k] = (uint32_t) c->h_pairs.size();
        idxRow.push_back(rowOf[k]);
        c->h_pairs.push_back(SwPair{2 * qi, t, tl, fwd[i].score});
        c->alnStats[3] += (uint64_t) c->q[qi].len * tl;
    }
    const double tE = now_s();
    build_tasks(c, idxRev, false);
    build_tasks(c, idxRow, true);
    rc = run_sw(c);
    const double tF = now_s();
    if (rc != FSB200_OK) return rc;
    for (size_t k = 0; k < ns; k++) {
        const uint32_t i = surv[k];
        const QueryHost &qh = c->q[c->candQuery[i]];
        fsb200_aln &r = c->results[i];
        r.rev_score = c->h_out[k].score;
        if (fwd[i].score == 0) r.q_end = (int32_t) qh.len - 1;     // SSW.cpp:1112: end_read stays query_length-1
        else if (rowOf[k] == 0xFFFFFFFFu) r.q_end = fwd[i].qEnd;
        else r.q_end = c->h_out[rowOf[k]].qEnd;
        r.score = (int32_t) r.fwd_score - (int32_t) r.rev_score;
        r.status = (r.score >= qh.minScore) ? 0 : 2;
        r.evalue = c->nn.evalue((double) r.score, qh.mu, qh.lambda);
    }
    if (tmg) fprintf(stderr, "align_run: pairs %.1f ms, tasks %.1f ms, phaseA run_sw %.1f ms, gate+phaseB prep %.1f ms, phaseB run_sw %.1f ms, finish %.1f ms\n",
                     1e3 * (tB - tA), 1e3 * (tC - tB), 1e3 * (tD - tC), 1e3 * (tE - tD), 1e3 * (tF - tE), 1e3 * (now_s() - tF));
    return FSB200_OK;
}

// E-value of a pair rejected after the forward pass (fsb200_aln.status == 1): the gate itself is an integer comparison on the device
// (min_passing_score), the value is only reported.  computeEvalueCorr costs five libm calls and depends on (query, integer score)
// alone: one table per query and thread, filled on demand, gives bit-identical values for a tenth of the calls.
namespace {
struct EvalueMemo {
    static constexpr int LO = -64, N = 1088;
    double v[N]; uint8_t have[N];
    void reset() { memset(have, 0, sizeof(have)); }
    double get(const host::EvalueModel &nn, int score, double mu, double lambda) {
        const int k = score - LO;
        if (k < 0 || k >= N) return nn.evalue((double) score, mu, lambda);
        if (!have[k]) { v[k] = nn.evalue((double) score, mu, lambda); have[k] = 1; }
        return v[k];
    }
};
}  // namespace

int fsb200_align_download(fsb200_ctx *c, fsb200_aln *out) {
    if (c == NULL || (out == NULL && !c->results.empty())) return FSB200_E_ARG;
    enter(c);
    if (c->results.empty() || c->candOff.size() != (size_t) c->nQ + 1) return FSB200_OK;
#pragma omp parallel
    {
        EvalueMemo memo;
#pragma omp for schedule(dynamic, 8)
        for (uint32_t q = 0; q < c->nQ; q++) {
            const QueryHost &qh = c->q[q];
            memo.reset();
            for (uint64_t i = c->candOff[q]; i < c->candOff[q + 1]; i++) {
                out[i] = c->results[i];
                if (out[i].status == 1) out[i].evalue = memo.get(c->nn, out[i].fwd_score, qh.mu, qh.lambda);
            }
        }
    }
    return FSB200_OK;
}

int fsb200_align(fsb200_ctx *c, const uint64_t *candOff, const uint32_t *candTarget, fsb200_aln *out) {
    int rc = fsb200_align_upload(c, candOff, candTarget);
    if (rc != FSB200_OK) return rc;
    rc = fsb200_align_run(c);
    if (rc != FSB200_OK) return rc;
    return fsb200_align_download(c, out);
}

int fsb200_align_stats(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->alnStats, sizeof(c->alnStats));
    return FSB200_OK;
}

// ---- traceback of the accepted pairs (alignStartPosBacktraceBlock, StructureSmithWaterman.cpp:369-537) ----
int fsb200_backtrace_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    if (c->nT == 0 || c->nQ == 0) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    enter(c);
    memset(c->btStats, 0, sizeof(c->btStats));
    const bool tmgBt = getenv("FSB200_TIMING") != NULL;
    const double tBt0 = now_s();
    const size_t n = c->results.size();
    std::vector<uint32_t> &acc = c->accIdx;
    acc.clear();
    c->btValid = c->ssValid = false;
    {   // one status per candidate pair (~1e6 per 1k queries): scan with all host threads
        std::vector<std::vector<uint32_t>> part(omp_get_max_threads());
#pragma omp parallel
        {
            std::vector<uint32_t> &mine = part[omp_get_thread_num()];
#pragma omp for schedule(static)
            for (size_t i = 0; i < n; i++) if (c->results[i].status == 0) mine.push_back((uint32_t) i);
        }
        for (auto &v : part) acc.insert(acc.end(), v.begin(), v.end());   // static schedule: already in ascending order
    }
    c->accMeta.assign(acc.size(), fsb200_bt{-1, -1, 0, 0, 0, -1, 0});
    c->btBuf.clear();
    c->btDeviceBytes = 0;
    std::vector<tb::TbPair> pairs;
    uint64_t btTotal = 0;
    uint32_t maxSum = 0;
    for (size_t ai = 0; ai < acc.size(); ai++) {
        const uint32_t i = acc[ai];
        const fsb200_aln &r = c->results[i];
        const QueryHost &qh = c->q[c->candQuery[i]];
        tb::TbPair p;
        p.qOff = qh.off; p.tOff = c->h_tOff[c->candTarget[i]]; p.btOff = btTotal;
        p.qEnd = r.q_end; p.dbEnd = r.db_end; p.targetScore = (int32_t) r.fwd_score; p.pad = 0;
        if (p.qEnd < -1 || p.dbEnd < -1) return fail(c, FSB200_E_STATE, "accepted pair without end positions");
        const uint32_t sum = (uint32_t) (p.qEnd + 1) + (uint32_t) (p.dbEnd + 1);
        c->accMeta[ai].bt_off = btTotal;
        btTotal += sum;
        maxSum = std::max(maxSum, sum);
        pairs.push_back(p);
    }
    const size_t ns = pairs.size();
    c->btStats[0] = ns;
    c->btBuf.assign(btTotal, 0);
    c->btValid = true;
    if (ns == 0) return FSB200_OK;
    // AAMatrix::set writes [a][b] and [b][a]; the caller loops aa1 outer, aa2 inner (StructureSmithWaterman.cpp:428-446), so the
    // pair {a, b} ends up with subMatrix[max(a,b)][min(a,b)]. Padding row/column keeps i8::MIN (scores.rs:52-57).
    std::vector<int8_t> tabs(2 * tb::TB_STRIDE * tb::TB_STRIDE, INT8_MIN);
    for (int a = 0; a < 21; a++) for (int b = 0; b < 21; b++) {
        tabs[a * tb::TB_STRIDE + b] = (int8_t) c->mAA.s[std::max(a, b)][std::min(a, b)];
        tabs[tb::TB_STRIDE * tb::TB_STRIDE + a * tb::TB_STRIDE + b] = (int8_t) c->m3Di.s[std::max(a, b)][std::min(a, b)];
    }
    CUDA_OK(c, c->d_tbTabs.reserve(tabs.size()));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tbTabs.p, tabs.data(), tabs.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, c->d_tbPairs.reserve(ns)); CUDA_OK(c, c->d_tbOrder.reserve(ns)); CUDA_OK(c, c->d_tbOut.reserve(ns));
    CUDA_OK(c, c->d_bt.reserve(btTotal + 16));
    CUDA_OK(c, c->d_counter.reserve(64));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tbPairs.p, pairs.data(), ns * sizeof(tb::TbPair), cudaMemcpyHostToDevice, c->stream));
    // longest first: the persistent groups finish together
    std::vector<uint32_t> order(ns);
    for (size_t k = 0; k < ns; k++) order[k] = (uint32_t) k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return pairs[x].qEnd + pairs[x].dbEnd > pairs[y].qEnd + pairs[y].dbEnd;
    });
    std::vector<tb::TbOut> out(ns);
    const uint32_t blockCap = maxSum + 64;
    // pass 0: 32 K vectors of trace per pair in flight; pairs whose block path needs more are redone with a larger buffer
    uint64_t capVec = 32768;
    size_t groupsMax = (size_t) c->smCount * 2 * (tb::TB_THREADS / tb::TL);
    const uint64_t worstVec = (uint64_t) (tb::TB_MAXS / tb::TL) * ((uint64_t) maxSum + 2 * tb::TB_MAXS);   // Trace::new (scan_block.rs:1742-1746)
    for (int pass = 0; !order.empty(); pass++) {
        if (pass > 3) return fail(c, FSB200_E_RANGE, "traceback: trace buffer exhausted");
        const size_t m = order.size();
        const size_t groups = std::min(m, groupsMax);
        const unsigned ctas = (unsigned) ((groups + tb::TB_THREADS / tb::TL - 1) / (tb::TB_THREADS / tb::TL));
        const size_t g = (size_t) ctas * (tb::TB_THREADS / tb::TL);
        CUDA_OK(c, c->d_tbState.reserve(g * tb::TB_STATE));
        CUDA_OK(c, c->d_tbTrace.reserve(g * capVec * tb::TL));
        CUDA_OK(c, c->d_tbBlocks.reserve(g * blockCap));
        CUDA_OK(c, cudaMemcpyAsync(c->d_tbOrder.p, order.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(c, cudaMemsetAsync(c->d_counter.p + 48, 0, sizeof(uint32_t), c->stream));
        tb::TbArgs a;
        a.pairs = c->d_tbPairs.p; a.order = c->d_tbOrder.p; a.nPairs = (uint32_t) m; a.counter = c->d_counter.p + 48;
        a.qAA = c->d_pAA.p; a.q3Di = c->d_p3Di.p; a.qBiasAA = c->d_pBiasAA.p; a.qBiasSS = c->d_pBiasSS.p;
        a.tAA = c->d_tAA.p; a.t3Di = c->d_t3Di.p;
        a.tabAA = c->d_tbTabs.p; a.tab3Di = c->d_tbTabs.p + tb::TB_STRIDE * tb::TB_STRIDE;
        a.gapOpen = -c->cfg.gap_open; a.gapExtend = -c->cfg.gap_extend;
        a.state = c->d_tbState.p; a.trace = c->d_tbTrace.p; a.traceCapVec = capVec;
        a.blocks = c->d_tbBlocks.p; a.blockCap = blockCap; a.bt = c->d_bt.p; a.out = c->d_tbOut.p;
        CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
        tb::traceback_kernel<<<ctas, tb::TB_THREADS, 0, c->stream>>>(a);
        CUDA_OK(c, cudaGetLastError());
        CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
        CUDA_OK(c, cudaMemcpyAsync(out.data(), c->d_tbOut.p, ns * sizeof(tb::TbOut), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_OK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->btStats[1] += 1;
        c->btStats[2] += (uint64_t) (ms * 1e6);
        std::vector<uint32_t> redo;
        for (uint32_t k : order) if (out[k].ok < 0) redo.push_back(k);
        c->btStats[3] += redo.size();
        order.swap(redo);
        if (pass == 0) { capVec = std::min<uint64_t>(worstVec, 1u << 20); groupsMax = 64; }
        else { capVec = worstVec; groupsMax = 8; }
    }
    CUDA_OK(c, cudaMemcpyAsync(c->btBuf.data(), c->d_bt.p, btTotal, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->btDeviceBytes = btTotal;
    for (size_t k = 0; k < ns; k++) {
        fsb200_bt &m = c->accMeta[k];
        m.status = out[k].ok; m.block_score = out[k].score;
        if (out[k].ok == 1) { m.q_start = out[k].qStart; m.db_start = out[k].dbStart; m.identical_aa = out[k].identical; m.aln_len = out[k].alnLen; }
        else c->btStats[4] += 1;
    }
    if (tmgBt) fprintf(stderr, "backtrace_run: %zu pairs, %.1f ms (kernel %.2f ms)\n", ns, 1e3 * (now_s() - tBt0), c->btStats[2] / 1e6);
    return FSB200_OK;
}

// Test hook: one run of the CUDA block-aligner engine (traceback.cuh) on caller-supplied data, see include/fsb200.h.
int fsb200_test_block_align(fsb200_ctx *c, const uint8_t *qAA, const uint8_t *q3Di, const int8_t *qBiasAA, const int8_t *qBiasSS, uint32_t qLen,
                            const uint8_t *tAA, const uint8_t *t3Di, uint32_t tLen, const int8_t *tabAA, const int8_t *tab3Di,
                            int32_t gapOpen, int32_t gapExtend, uint32_t minSize, uint32_t maxSize, int32_t xDrop,
                            int32_t *out3, char *bt, uint32_t btCap) {
    if (c == NULL || qAA == NULL || q3Di == NULL || tAA == NULL || t3Di == NULL || tabAA == NULL || tab3Di == NULL || out3 == NULL) return FSB200_E_ARG;
    if (qLen == 0 || tLen == 0 || qLen > 65535 || tLen > 65535 || minSize < (uint32_t) tb::TL || maxSize > (uint32_t) tb::TB_MAXS || minSize > maxSize ||
        (minSize & (minSize - 1)) || (maxSize & (maxSize - 1))) return fail(c, FSB200_E_RANGE, "test_block_align: lengths 1..65535, block sizes powers of two in 16..4096");
    enter(c);
    // the engine reads the reversed prefixes that end at (qEnd, dbEnd): hand it the sequences reversed so that it sees them as given
    const size_t qPad = ((size_t) qLen + 16 + 15) & ~(size_t) 15, tPad = ((size_t) tLen + 16 + 15) & ~(size_t) 15;
    std::vector<uint8_t> hq(2 * qPad, 20), ht(2 * tPad, 20);
    std::vector<int8_t> hb(2 * qPad, 0);
    for (uint32_t i = 0; i < qLen; i++) {
        if (qAA[i] > 20 || q3Di[i] > 20) return fail(c, FSB200_E_ARG, "residue code > 20");
        hq[qLen - 1 - i] = qAA[i]; hq[qPad + qLen - 1 - i] = q3Di[i];
        hb[qLen - 1 - i] = qBiasAA ? qBiasAA[i] : 0; hb[qPad + qLen - 1 - i] = qBiasSS ? qBiasSS[i] : 0;
    }
    for (uint32_t i = 0; i < tLen; i++) {
        if (tAA[i] > 20 || t3Di[i] > 20) return fail(c, FSB200_E_ARG, "residue code > 20");
        ht[tLen - 1 - i] = tAA[i]; ht[tPad + tLen - 1 - i] = t3Di[i];
    }
    DevBuf<uint8_t> dq, dt; DevBuf<int8_t> db, dtab; DevBuf<tb::TbPair> dp; DevBuf<uint32_t> dord, dctr; DevBuf<int16_t> dstate; DevBuf<uint8_t> dtrace;
    DevBuf<uint4> dblocks; DevBuf<char> dbt; DevBuf<tb::TbOut> dout;
    struct Rel { std::vector<std::function<void()>> f; ~Rel() { for (auto &g : f) g(); } } rel;
    rel.f = {[&] { dq.release(); dt.release(); db.release(); dtab.release(); dp.release(); dord.release(); dctr.release(); dstate.release();
                   dtrace.release(); dblocks.release(); dbt.release(); dout.release(); }};
    const uint32_t sum = qLen + tLen;
    const uint64_t capVec = (uint64_t) (maxSize / tb::TL) * ((uint64_t) sum + 2 * maxSize) + 64;   // Trace::new (scan_block.rs:1742-1746)
    const size_t groups = tb::TB_THREADS / tb::TL;
    CUDA_OK(c, dq.reserve(hq.size())); CUDA_OK(c, dt.reserve(ht.size())); CUDA_OK(c, db.reserve(hb.size()));
    CUDA_OK(c, dtab.reserve(2 * tb::TB_STRIDE * tb::TB_STRIDE)); CUDA_OK(c, dp.reserve(1)); CUDA_OK(c, dord.reserve(1)); CUDA_OK(c, dctr.reserve(4));
    CUDA_OK(c, dstate.reserve(groups * tb::TB_STATE)); CUDA_OK(c, dtrace.reserve(groups * capVec * tb::TL));
    CUDA_OK(c, dblocks.reserve(groups * (sum + 64))); CUDA_OK(c, dbt.reserve(sum + 16)); CUDA_OK(c, dout.reserve(1));
    tb::TbPair p; p.qOff = 0; p.tOff = 0; p.btOff = 0; p.qEnd = (int32_t) qLen - 1; p.dbEnd = (int32_t) tLen - 1; p.targetScore = 0; p.pad = 0;
    const uint32_t zero = 0;
    CUDA_OK(c, cudaMemcpyAsync(dq.p, hq.data(), hq.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dt.p, ht.data(), ht.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(db.p, hb.data(), hb.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dtab.p, tabAA, tb::TB_STRIDE * tb::TB_STRIDE, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dtab.p + tb::TB_STRIDE * tb::TB_STRIDE, tab3Di, tb::TB_STRIDE * tb::TB_STRIDE, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dp.p, &p, sizeof(p), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dord.p, &zero, 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemsetAsync(dctr.p, 0, 16, c->stream));
    CUDA_OK(c, cudaMemsetAsync(dbt.p, 0, sum + 16, c->stream));
    tb::TbArgs a;
    a.pairs = dp.p; a.order = dord.p; a.nPairs = 1; a.counter = dctr.p;
    a.qAA = dq.p; a.q3Di = dq.p + qPad; a.qBiasAA = db.p; a.qBiasSS = db.p + qPad; a.tAA = dt.p; a.t3Di = dt.p + tPad;
    a.tabAA = dtab.p; a.tab3Di = dtab.p + tb::TB_STRIDE * tb::TB_STRIDE; a.gapOpen = gapOpen; a.gapExtend = gapExtend;
    a.state = dstate.p; a.trace = dtrace.p; a.traceCapVec = capVec; a.blocks = dblocks.p; a.blockCap = sum + 64; a.bt = dbt.p; a.out = dout.p;
    a.testMinSize = minSize; a.testMaxSize = maxSize; a.testXDrop = xDrop;
    tb::traceback_kernel<<<1, tb::TB_THREADS, 0, c->stream>>>(a);
    CUDA_OK(c, cudaGetLastError());
    tb::TbOut o;
    std::vector<char> hbt(sum + 16, 0);
    CUDA_OK(c, cudaMemcpyAsync(&o, dout.p, sizeof(o), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(hbt.data(), dbt.p, sum + 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (o.ok != 1) return fail(c, FSB200_E_RANGE, "test_block_align: trace capacity exceeded");
    out3[0] = o.score; out3[1] = (int32_t) qLen - o.qStart; out3[2] = (int32_t) tLen - o.dbStart;   // query_idx / reference_idx of the engine
    if (bt != NULL && btCap > 0) {
        const uint32_t n = std::min<uint32_t>(o.alnLen, btCap - 1);
        memcpy(bt, hbt.data(), n); bt[n] = 0;
    }
    return FSB200_OK;
}

int fsb200_backtrace_size(const fsb200_ctx *c, uint64_t *bytes) {
    if (c == NULL || bytes == NULL) return FSB200_E_ARG;
    *bytes = c->btBuf.size();
    return FSB200_OK;
}

int fsb200_backtrace_download(fsb200_ctx *c, fsb200_bt *meta, char *bt, uint64_t cap) {
    if (c == NULL || (meta == NULL && !c->results.empty())) return FSB200_E_ARG;
    if (!c->btValid) return fail(c, FSB200_E_STATE, "fsb200_backtrace_run first");
    enter(c);
    if (cap < c->btBuf.size() || (bt == NULL && !c->btBuf.empty())) return fail(c, FSB200_E_RANGE, "backtrace buffer too small, see fsb200_backtrace_size");
    const size_t n = c->results.size();
    const fsb200_bt none{-1, -1, 0, 0, 0, -1, 0};
#pragma omp parallel for schedule(static, 16384)
    for (size_t i = 0; i < n; i++) meta[i] = none;
    for (size_t k = 0; k < c->accIdx.size(); k++) meta[c->accIdx[k]] = c->accMeta[k];
    if (!c->btBuf.empty()) par_memcpy(bt, c->btBuf.data(), c->btBuf.size());
    return FSB200_OK;
}

int fsb200_backtrace_stats(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->btStats, sizeof(c->btStats));
    return FSB200_OK;
}

// ---- C-alpha coordinates + LDDT / TM-score of the traced pairs (structurealign.cpp:376-412) ----
namespace {
int stage_ca(fsb200_ctx *c, const float *x, const float *y, const float *z, const uint64_t *off, const uint32_t *len, uint32_t n,
             std::vector<uint64_t> &outOff, DevBuf<float> &dev, uint64_t &total) {
    outOff.resize(n);
    total = 0;
    for (uint32_t i = 0; i < n; i++) { outOff[i] = total; total += len[i]; }
    std::vector<float> h(3 * total + 3);
    for (uint32_t i = 0; i < n; i++) {
        memcpy(h.data() + outOff[i], x + off[i], len[i] * sizeof(float));
        memcpy(h.data() + total + outOff[i], y + off[i], len[i] * sizeof(float));
        memcpy(h.data() + 2 * total + outOff[i], z + off[i], len[i] * sizeof(float));
    }
    CUDA_OK(c, dev.reserve(3 * total + 3));
    CUDA_OK(c, cudaMemcpyAsync(dev.p, h.data(), 3 * total * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return FSB200_OK;
}
}  // namespace

int fsb200_load_target_ca(fsb200_ctx *c, const float *x, const float *y, const float *z, const uint64_t *off) {
    if (c == NULL || x == NULL || y == NULL || z == NULL || off == NULL) return FSB200_E_ARG;
    if (c->nT == 0) return fail(c, FSB200_E_STATE, "load_target first");
    if (!c->isLane && c->laneGroup && *c->laneGroup > 1) return fail(c, FSB200_E_STATE, "lanes share this context's target: destroy them before loading another");
    enter(c);
    const int rc = stage_ca(c, x, y, z, off, c->h_tLen.data(), c->nT, c->h_tCaOff, c->d_tCa, c->tCaTotal);
    c->haveTCa = rc == FSB200_OK;
    return rc;
}

int fsb200_set_queries_ca(fsb200_ctx *c, const float *x, const float *y, const float *z, const uint64_t *off) {
    if (c == NULL || x == NULL || y == NULL || z == NULL || off == NULL) return FSB200_E_ARG;
    if (c->nQ == 0) return fail(c, FSB200_E_STATE, "set_queries first");
    enter(c);
    std::vector<uint32_t> len(c->nQ);
    for (uint32_t i = 0; i < c->nQ; i++) len[i] = c->q[i].len;
    int rc = stage_ca(c, x, y, z, off, len.data(), c->nQ, c->h_qCaOff, c->d_qCa, c->qCaTotal);
    if (rc != FSB200_OK) return rc;
    // LDDTCalculator::initQuery: 1 / (number of residues closer than 15 A) per query residue
    CUDA_OK(c, c->d_qNorm.reserve(c->qCaTotal + 1));
    CUDA_OK(c, c->d_qCaOff.reserve(c->nQ)); CUDA_OK(c, c->d_qCaLen.reserve(c->nQ));
    CUDA_OK(c, cudaMemcpyAsync(c->d_qCaOff.p, c->h_qCaOff.data(), c->nQ * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_qCaLen.p, len.data(), c->nQ * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    ss::lddt_norm_kernel<<<c->nQ, 128, 0, c->stream>>>(c->d_qCa.p, c->d_qCa.p + c->qCaTotal, c->d_qCa.p + 2 * c->qCaTotal, c->d_qCaOff.p,
                                                        c->d_qCaLen.p, c->nQ, c->d_qNorm.p);
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->haveQCa = true;
    return FSB200_OK;
}

int fsb200_structure_scores_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    if (!c->haveTCa || !c->haveQCa) return fail(c, FSB200_E_STATE, "load_target_ca and set_queries_ca first");
    if (!c->btValid) return fail(c, FSB200_E_STATE, "fsb200_backtrace_run first");
    enter(c);
    memset(c->ssStats, 0, sizeof(c->ssStats));
    c->accSs.assign(c->accIdx.size(), fsb200_sscore{0.0, 0.0f, -1, 0});
    c->ssValid = true;
    std::vector<ss::SsPair> pairs;
    std::vector<uint32_t> where;
    uint32_t maxCols = 0;
    for (size_t ai = 0; ai < c->accIdx.size(); ai++) {
        const size_t i = c->accIdx[ai];
        const fsb200_bt &m = c->accMeta[ai];
        if (m.status != 1 || m.aln_len == 0) continue;
        const fsb200_aln &r = c->results[i];
        const uint32_t qi = c->candQuery[i], t = c->candTarget[i];
        ss::SsPair p;
        p.qCaOff = c->h_qCaOff[qi]; p.tCaOff = c->h_tCaOff[t]; p.btOff = m.bt_off; p.qLen = c->q[qi].len; p.tLen = c->h_tLen[t];
        p.qStart = m.q_start; p.tStart = m.db_start; p.btLen = m.aln_len; p.qNormOff = c->h_qCaOff[qi]; p.pad = 0;
        uint32_t nM = 0;
        const char *b = c->btBuf.data() + m.bt_off;
        for (uint32_t k = 0; k < m.aln_len; k++) nM += b[k] == 'M';
        p.nAligned = nM;
        if (nM == 0) continue;
        // TMaligner::normalization(par.tmScoreThrMode, min(qEnd - qStart, dbEnd - dbStart), qLen, dbLen) (structurealign.cpp:391,
        // TMaligner.cpp:330-341)
        const int mode = c->cfg.tmscore_norm_mode;
        const int normLen = mode == 1 ? (int) p.qLen : mode == 2 ? (int) p.tLen : mode == 3 ? (int) std::min(p.qLen, p.tLen)
                                                                                             : std::min(r.q_end - m.q_start, r.db_end - m.db_start);
        // parameter_set4search (lib/tmalign/TMalign.cpp:25-47), standard_TMscore (:1371-1377)
        float Lnorm = (float) normLen, d0, D0_MIN, d0_search, score_d8, sD0;
        if (Lnorm <= 19) d0 = 0.168; else d0 = (1.24 * pow((Lnorm * 1.0 - 15), 1.0 / 3) - 1.8);
        D0_MIN = d0 + 0.8;
        d0 = D0_MIN;
        d0_search = d0;
        if (d0_search > 8) d0_search = 8;
        if (d0_search < 4.5) d0_search = 4.5;
        score_d8 = 1.5 * pow(Lnorm * 1.0, 0.3) + 3.5;
        { float dmin = 0.5; if (Lnorm > 21) sD0 = (1.24 * pow((Lnorm * 1.0 - 15), 1.0 / 3) - 1.8); else sD0 = dmin; if (sD0 < dmin) sD0 = dmin; }
        p.Lnorm = Lnorm; p.scoreD8 = score_d8; p.d0Search = d0; p.d0SearchClipped = d0_search; p.d0Std = sD0;
        maxCols = std::max(maxCols, nM);
        pairs.push_back(p);
        where.push_back((uint32_t) ai);
    }
    const size_t ns = pairs.size();
    c->ssStats[0] = ns;
    if (ns == 0) return FSB200_OK;
    if (c->btDeviceBytes == 0) return fail(c, FSB200_E_STATE, "backtraces are no longer resident on the device");
    // longest alignments first
    std::vector<uint32_t> order(ns);
    for (size_t k = 0; k < ns; k++) order[k] = (uint32_t) k;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return pairs[x].nAligned > pairs[y].nAligned; });
    std::vector<ss::SsPair> sorted(ns);
    for (size_t k = 0; k < ns; k++) sorted[k] = pairs[order[k]];
    const uint32_t cols = (maxCols + 31) & ~31u;
    const size_t perWarp = (size_t) cols * ss::SS_FLOATS_PER_COL * sizeof(float);
    const size_t smemLimit = 200 * 1024;
    if (perWarp > smemLimit) return fail(c, FSB200_E_RANGE, "alignment with more than 3900 aligned columns: structure scores need a shared-memory tile per pair");
    const int warps = (int) std::max<size_t>(1, std::min<size_t>(8, smemLimit / perWarp));
    const size_t smem = perWarp * warps;
    CUDA_OK(c, cudaFuncSetAttribute(ss::structscore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CUDA_OK(c, c->d_ssPairs.reserve(ns)); CUDA_OK(c, c->d_ssOut.reserve(ns)); CUDA_OK(c, c->d_counter.reserve(64));
    CUDA_OK(c, cudaMemcpyAsync(c->d_ssPairs.p, sorted.data(), ns * sizeof(ss::SsPair), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemsetAsync(c->d_counter.p + 52, 0, sizeof(uint32_t), c->stream));
    ss::SsArgs a;
    a.pairs = c->d_ssPairs.p; a.nPairs = (uint32_t) ns; a.counter = c->d_counter.p + 52;
    a.qx = c->d_qCa.p; a.qy = c->d_qCa.p + c->qCaTotal; a.qz = c->d_qCa.p + 2 * c->qCaTotal;
    a.tx = c->d_tCa.p; a.ty = c->d_tCa.p + c->tCaTotal; a.tz = c->d_tCa.p + 2 * c->tCaTotal;
    a.qNorm = c->d_qNorm.p; a.bt = c->d_bt.p; a.out = c->d_ssOut.p; a.colsPerWarp = cols;
    // few pairs (the usual case: a handful of accepted hits per query): several warps share a pair, each runs a share of the
    // independent search starts and LDDT rows; many pairs: one warp per pair fills the GPU anyway
    const size_t warpSlots = (size_t) c->smCount * 2 * warps;
    a.jobsPerPair = (uint32_t) std::max<size_t>(1, std::min<size_t>(16, warpSlots / std::max<size_t>(1, ns)));
    CUDA_OK(c, c->d_ssScratch.reserve(ns * ((size_t) cols + 8)));
    CUDA_OK(c, cudaMemsetAsync(c->d_ssScratch.p, 0, ns * ((size_t) cols + 8) * sizeof(float), c->stream));
    a.scratch = c->d_ssScratch.p;
    const unsigned ctas = (unsigned) std::min<size_t>((ns * a.jobsPerPair + warps - 1) / warps, (size_t) c->smCount * 2);
    CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
    ss::structscore_kernel<<<ctas, warps * 32, smem, c->stream>>>(a);
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
    std::vector<ss::SsOut> out(ns);
    CUDA_OK(c, cudaMemcpyAsync(out.data(), c->d_ssOut.p, ns * sizeof(ss::SsOut), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_OK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->ssStats[1] = 1; c->ssStats[2] = (uint64_t) (ms * 1e6);
    for (size_t k = 0; k < ns; k++) {
        const size_t ai = where[order[k]], i = c->accIdx[ai];
        fsb200_sscore &o = c->accSs[ai];
        o.tm = out[k].tm; o.lddt = out[k].lddt; o.status = (out[k].flags & 1) ? 0 : 1;
        // res.score = res.score * sqrt(avgLddtScore * tmscore) (structurealign.cpp:406-410)
        o.score_bits = (int32_t) (c->results[i].score * sqrt((double) o.lddt * o.tm));
        if (out[k].flags & 1) c->ssStats[3] += 1;
    }
    return FSB200_OK;
}

int fsb200_structure_scores_download(fsb200_ctx *c, fsb200_sscore *out) {
    if (c == NULL || (out == NULL && !c->results.empty())) return FSB200_E_ARG;
    if (!c->ssValid) return fail(c, FSB200_E_STATE, "fsb200_structure_scores_run first");
    enter(c);
    const size_t n = c->results.size();
    const fsb200_sscore none{0.0, 0.0f, -1, 0};
#pragma omp parallel for schedule(static, 16384)
    for (size_t i = 0; i < n; i++) out[i] = none;
    for (size_t k = 0; k < c->accIdx.size(); k++) out[c->accIdx[k]] = c->accSs[k];
    return FSB200_OK;
}

int fsb200_structure_scores_stats(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->ssStats, sizeof(c->ssStats));
    return FSB200_OK;
}

int fsb200_prefilter(fsb200_ctx *c, fsb200_hit *hits, uint32_t *counts) {
    if (c == NULL || hits == NULL || counts == NULL) return FSB200_E_ARG;
    if (c->nT == 0 || c->nQ == 0) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    enter(c);
    std::string err;
    c->pf.share = c->laneGroup ? std::max(1, *c->laneGroup) : 1;
    int rc = prefilter::run(c->pf, c->cfg, c->mKmer, c->mUngapped, c->stream, c->smCount,
                            c->d_t3Di.p, c->tBytes, c->d_tOff.p, c->d_tLen.p, c->h_tOff, c->h_tLen, c->nT,
                            c->h_p3Di, c->q.size(), [&](size_t i, uint64_t &off, uint32_t &len) { off = c->q[i].off; len = c->q[i].len; },
                            c->identity.empty() ? nullptr : c->identity.data(), hits, counts, err);
    if (rc != FSB200_OK) return fail(c, rc, err);
    return FSB200_OK;
}

int fsb200_prefilter_stats(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->pf.stats, sizeof(c->pf.stats));
    return FSB200_OK;
}

int fsb200_prefilter_stage_ns(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->pf.stageNs, 8 * sizeof(uint64_t));
    return FSB200_OK;
}

int fsb200_set_identity(fsb200_ctx *c, const uint32_t *identity) {
    if (c == NULL) return FSB200_E_ARG;
    if (identity == NULL) { c->identity.clear(); return FSB200_OK; }
    if (c->nQ == 0 || c->nT == 0) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    for (uint32_t i = 0; i < c->nQ; i++) if (identity[i] != 0xFFFFFFFFu && identity[i] >= c->nT) return fail(c, FSB200_E_ARG, "identity target id out of range");
    c->identity.assign(identity, identity + c->nQ);
    return FSB200_OK;
}


int fsb200_search_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    enter(c);
    const bool tm = getenv("FSB200_TIMING") != NULL;
    const double t0 = now_s();
    const size_t ms = (size_t) c->cfg.max_seqs;
    c->pfHits.resize((size_t) c->nQ * ms);
    c->pfCounts.assign(c->nQ, 0);
    int rc = fsb200_prefilter(c, c->pfHits.data(), c->pfCounts.data());
    if (rc != FSB200_OK) return rc;
    std::vector<uint64_t> off(c->nQ + 1, 0);
    for (uint32_t q = 0; q < c->nQ; q++) off[q + 1] = off[q] + c->pfCounts[q];
    // candidate lists straight into the context (what fsb200_align_upload does for a caller's CSR; the ids come from select_kernel)
    c->candOff.swap(off);
    c->candTarget.resize(c->candOff[c->nQ]); c->candQuery.resize(c->candOff[c->nQ]);
#pragma omp parallel for schedule(static, 16)
    for (uint32_t q = 0; q < c->nQ; q++) {
        const fsb200_hit *h = c->pfHits.data() + (size_t) q * ms;
        uint32_t *tg = c->candTarget.data() + c->candOff[q], *cq = c->candQuery.data() + c->candOff[q];
        for (uint32_t k = 0; k < c->pfCounts[q]; k++) { tg[k] = h[k].target; cq[k] = q; }
    }
    const double t1 = now_s();
    const double t2 = now_s();
    rc = fsb200_align_run(c);
    if (tm) fprintf(stderr, "search_run: prefilter %.1f ms (match kernel %.1f), csr+upload %.1f ms, align_run %.1f ms (sw kernels %.1f)\n",
                    1e3 * (t1 - t0), c->pf.stats[6] / 1e6, 1e3 * (t2 - t1), 1e3 * (now_s() - t2), c->alnStats[6] / 1e6);
    return rc;
}

int fsb200_search_download(fsb200_ctx *c, fsb200_hit *hits, uint32_t *counts, fsb200_aln *aln) {
    if (c == NULL || hits == NULL || counts == NULL || aln == NULL) return FSB200_E_ARG;
    enter(c);
    const size_t ms = (size_t) c->cfg.max_seqs;
    memcpy(counts, c->pfCounts.data(), c->pfCounts.size() * sizeof(uint32_t));
#pragma omp parallel
    {
        EvalueMemo memo;
#pragma omp for schedule(dynamic, 8)
        for (uint32_t q = 0; q < c->nQ; q++) {
            memcpy(hits + (size_t) q * ms, c->pfHits.data() + (size_t) q * ms, ms * sizeof(fsb200_hit));
            const QueryHost &qh = c->q[q];
            memo.reset();
            for (uint32_t k = 0; k < c->pfCounts[q]; k++) {
                fsb200_aln r = c->results[c->candOff[q] + k];
                if (r.status == 1) r.evalue = memo.get(c->nn, r.fwd_score, qh.mu, qh.lambda);   // as in fsb200_align_download
                aln[(size_t) q * ms + k] = r;
            }
        }
    }
    return FSB200_OK;
}

int fsb200_search(fsb200_ctx *c, fsb200_hit *hits, uint32_t *counts, fsb200_aln *aln) {
    int rc = fsb200_search_run(c);
    if (rc != FSB200_OK) return rc;
    return fsb200_search_download(c, hits, counts, aln);
}

/* ---- index build as its own call, lanes ---- */
namespace {
int ensure_index(fsb200_ctx *c) {
    if (c->nT == 0) return fail(c, FSB200_E_STATE, "load_target first");
    std::string err;
    int rc;
    if (!c->pf.tablesReady && (rc = prefilter::build_tables(c->pf, c->mKmer, c->stream, err)) != FSB200_OK) return fail(c, rc, err);
    if (!c->pf.indexReady && (rc = prefilter::build_index(c->pf, c->cfg, c->mKmer, c->stream, c->d_t3Di.p, c->tBytes, c->d_tOff.p, c->d_tLen.p,
                                                           c->h_tOff, c->h_tLen, c->nT, err)) != FSB200_OK) return fail(c, rc, err);
    return FSB200_OK;
}
}  // namespace

int fsb200_build_index(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    enter(c);
    return ensure_index(c);
}

int fsb200_set_host_threads(fsb200_ctx *c, int32_t n) {
    if (c == NULL || n < 0) return FSB200_E_ARG;
    c->hostThreads = n;
    return FSB200_OK;
}

int fsb200_create_lane(fsb200_ctx *parent, fsb200_ctx **out) {
    if (parent == NULL || out == NULL) return FSB200_E_ARG;
    *out = NULL;
    if (parent->isLane) return fail(parent, FSB200_E_STATE, "lanes are created from the context that staged the target");
    if (parent->nT == 0) return fail(parent, FSB200_E_STATE, "load_target first");
    enter(parent);
    std::unique_ptr<fsb200_ctx> c(new fsb200_ctx());
    c->cfg = parent->cfg;
    c->mAA = parent->mAA; c->m3Di = parent->m3Di; c->mKmer = parent->mKmer; c->mUngapped = parent->mUngapped;
    c->nn = parent->nn; c->hasNN = parent->hasNN;
    c->smCount = parent->smCount;
    memcpy(c->tinyAA, parent->tinyAA, sizeof(c->tinyAA)); memcpy(c->tiny3Di, parent->tiny3Di, sizeof(c->tiny3Di));
    if (cudaStreamCreate(&c->stream) != cudaSuccess) return fail(parent, FSB200_E_CUDA, "cudaStreamCreate failed");
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    c->isLane = true; c->hostThreads = parent->hostThreads;
    // read-only device data of the parent: staged target, C-alpha, 3-mer tables and k-mer index
    c->nT = parent->nT; c->tResidues = parent->tResidues; c->aaDbResidues = parent->aaDbResidues; c->tBytes = parent->tBytes;
    c->h_tOff = parent->h_tOff; c->h_tLen = parent->h_tLen;
    c->d_tAA.borrow(parent->d_tAA); c->d_t3Di.borrow(parent->d_t3Di); c->d_tOff.borrow(parent->d_tOff); c->d_tLen.borrow(parent->d_tLen);
    if (parent->haveTCa) { c->haveTCa = true; c->h_tCaOff = parent->h_tCaOff; c->tCaTotal = parent->tCaTotal; c->d_tCa.borrow(parent->d_tCa); }
    if (parent->pf.tablesReady && parent->pf.indexReady) c->pf.borrow_index(parent->pf);   // else: a lane that prefilters builds its own
    if (!parent->laneGroup) parent->laneGroup = std::make_shared<int>(1);
    c->laneGroup = parent->laneGroup;
    *c->laneGroup += 1;
    *out = c.release();
    return FSB200_OK;
}

int fsb200_timer_start(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    enter(c);
    if (!c->tm0) { cudaEventCreate(&c->tm0); cudaEventCreate(&c->tm1); }
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    CUDA_OK(c, cudaEventRecord(c->tm0, c->stream));
    return FSB200_OK;
}

int fsb200_timer_stop(fsb200_ctx *c, double *ms) {
    if (c == NULL || ms == NULL || !c->tm0) return FSB200_E_ARG;
    CUDA_OK(c, cudaEventRecord(c->tm1, c->stream));
    CUDA_OK(c, cudaEventSynchronize(c->tm1));
    float f = 0;
    CUDA_OK(c, cudaEventElapsedTime(&f, c->tm0, c->tm1));
    *ms = (double) f;
    return FSB200_OK;
}

int fsb200_staged_target_size(const fsb200_ctx *c, uint64_t *bytes) {
    if (c == NULL || bytes == NULL) return FSB200_E_ARG;
    *bytes = c->tBytes;
    return FSB200_OK;
}

int fsb200_export_staged_target(const fsb200_ctx *c, void **dAA, void **dTdi, uint64_t *offOut, uint32_t *lenOut) {
    if (c == NULL || dAA == NULL || dTdi == NULL || c->nT == 0) return FSB200_E_ARG;
    *dAA = c->d_tAA.p; *dTdi = c->d_t3Di.p;
    if (offOut) memcpy(offOut, c->h_tOff.data(), c->nT * sizeof(uint64_t));
    if (lenOut) memcpy(lenOut, c->h_tLen.data(), c->nT * sizeof(uint32_t));
    return FSB200_OK;
}

int fsb200_load_target_device(fsb200_ctx *c, const void *dAA, const void *dTdi, uint64_t bytes, const uint64_t *off,
                              const uint32_t *len, uint32_t n, uint64_t aaDbResidues) {
    if (c == NULL || dAA == NULL || dTdi == NULL || off == NULL || len == NULL) return FSB200_E_ARG;
    enter(c);
    uint64_t res = 0;
    for (uint32_t i = 0; i < n; i++) { if (len[i] > 65535 || off[i] + len[i] > bytes) return FSB200_E_RANGE; res += len[i]; }
    c->h_tOff.assign(off, off + n); c->h_tLen.assign(len, len + n);
    invalidate_results(c, true);
    c->nT = n; c->tResidues = res; c->aaDbResidues = aaDbResidues; c->tBytes = bytes;
    CUDA_OK(c, c->d_tAA.reserve(bytes)); CUDA_OK(c, c->d_t3Di.reserve(bytes));
    CUDA_OK(c, c->d_tOff.reserve(std::max<uint32_t>(1, n))); CUDA_OK(c, c->d_tLen.reserve(std::max<uint32_t>(1, n)));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tAA.p, dAA, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_t3Di.p, dTdi, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tOff.p, c->h_tOff.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_tLen.p, c->h_tLen.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->nn.logDbResidues = log(static_cast<double>(aaDbResidues));
    c->pf.invalidate();
    return FSB200_OK;
}

}  // extern "C"
