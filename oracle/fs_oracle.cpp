// oracle/fs_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked or imported by the product path).
//
// A plain scalar CPU restatement of the reference's search hot path, written from the reference's
// *behaviour*; every function cites the reference file:line it follows (paths relative to
// /root/reference, M/ = lib/mmseqs/).  It is pinned against the reference itself: oracle/_ref/refseek
// (the reference's own sources compiled by oracle/Makefile) is run on the same inputs by
// tests/test_oracle_pin.py and by tests/golden/make_golden.py, which also wrote the committed fixtures.
//
// Parity status: stage 1 (prefilter hit lists) and stage 2 forward/reverse SW scores, end positions,
// (mu,lambda) and E-values are pinned against refseek.  Traceback (block-aligner, Rust) is UNPINNED:
// the Rust crate cannot be built in this image.
//
// Build: g++ -O2 -shared -fPIC (see oracle/build.sh).  C ABI, loaded by oracle/oracle.py via ctypes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sstream>
#include "block_aligner_port.h"

namespace {

const int A = 21;  // alphabet incl. X (M/src/commons/Parameters.cpp:2492-2559 alphabetSize=21)

// ---------------------------------------------------------------------------------------------
// Substitution matrices.  M/src/commons/SubstitutionMatrix.cpp:12-57 (ctor), :327-417 (readProbMatrix),
// M/src/commons/BaseMatrix.cpp:108-162 (computeBackground, generateSubMatrix).
// ---------------------------------------------------------------------------------------------
struct Matrix {
    short sub[A][A];
    double pBack[A];   // member pBack after readProbMatrix (used by calcLocalAaBiasCorrection)
    double lambda;
};

static size_t words_of_line(const std::string &line, std::vector<std::string> &w) {
    w.clear();
    std::istringstream is(line);
    std::string t;
    while (is >> t) w.push_back(t);
    return w.size();
}

// text = contents of a ".out" matrix file whose alphabet header is "A C D ... Y X".
static bool build_matrix(const char *text, double bitFactor, double scoreBias, Matrix &m) {
    std::string data(text);
    std::istringstream in(data);
    std::string line;
    std::vector<std::string> w;
    double prob[A][A];
    memset(prob, 0, sizeof(prob));
    for (int i = 0; i < A; i++) m.pBack[i] = 0.0;
    int order[256];
    for (int i = 0; i < 256; i++) order[i] = -1;
    bool started = false, hasLambda = false, hasBg = false;
    m.lambda = 0;
    while (std::getline(in, line)) {
        size_t n = words_of_line(line, w);
        if (!line.empty() && line[0] == '#') {
            // SubstitutionMatrix.cpp:340-352
            if (line.find("# Background (precomputed optional):") == 0) {
                for (size_t i = 4; i < n && i - 4 < (size_t) A; i++) m.pBack[i - 4] = strtod(w[i].c_str(), NULL);
                hasBg = true;
            }
            if (line.find("# Lambda     (precomputed optional):") == 0) {
                m.lambda = strtod(w[4].c_str(), NULL);
                hasLambda = true;
            }
            continue;
        }
        if (n > 1 && !started) {  // alphabet header (SubstitutionMatrix.cpp:419-452)
            for (size_t i = 0; i < n; i++) order[toupper(w[i][0])] = (int) i;
            started = true;
            continue;
        }
        if (n > 1 && started) {
            int aa = order[toupper(w[0][0])];
            if (aa < 0) return false;
            for (int i = 0; i < A && (size_t) (i + 1) < n; i++) prob[aa][i] = strtod(w[i + 1].c_str(), NULL);
        }
    }
    if (!hasLambda || !hasBg) return false;  // lambda estimation (LambdaCalculation.cpp) not restated: both shipped matrices carry them
    // X is never positive in the shipped matrices -> SubstitutionMatrix.cpp:404-408
    const int X = A - 1;
    bool xPositive = false;
    for (int j = 0; j < A; j++) if (prob[X][j] > 0 || prob[j][X] > 0) { xPositive = true; break; }
    if (!xPositive) for (int i = 0; i < A - 1; i++) m.pBack[i] = m.pBack[i] * (1.0 - m.pBack[X]);
    // SubstitutionMatrix.cpp:410-416
    for (int i = 0; i < A; i++) for (int j = 0; j < A; j++) prob[i][j] = std::exp(m.lambda * prob[i][j]) * m.pBack[i] * m.pBack[j];
    // BaseMatrix.cpp:108-119 (row sums, X forced to ANY_BACK) and :142-162
    double pb[A];
    for (int i = 0; i < A; i++) { pb[i] = 0; for (int j = 0; j < A; j++) pb[i] += prob[i][j]; }
    pb[A - 1] = 1E-5;
    for (int i = 0; i < A; i++) for (int j = 0; j < A; j++) {
        double s = std::log2(prob[i][j] / (pb[i] * pb[j]));
        double v = bitFactor * s + scoreBias;
        m.sub[i][j] = (short) ((v < 0.0) ? v - 0.5 : v + 0.5);
    }
    return true;
}

// M/src/commons/SubstitutionMatrix.cpp:79-109
static void comp_bias(const Matrix &m, const uint8_t *seq, int N, float scale, float *out) {
    const int windowSize = 40;
    for (int i = 0; i < N; i++) {
        const int minPos = std::max(0, i - windowSize / 2);
        const int maxPos = std::min(N, i + windowSize / 2);
        const int windowLength = maxPos - minPos;
        int sum = 0;
        const short *row = m.sub[seq[i]];
        for (int j = minPos; j < maxPos; j++) sum += row[seq[j]];
        sum -= row[seq[i]];
        float d = (float) sum;
        d /= -1.0 * static_cast<float>(windowLength);
        for (int a = 0; a < A; a++) d += m.pBack[a] * static_cast<float>(row[a]);
        out[i] = scale * d;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 2: query profile + striped-SW semantics.
// src/commons/StructureSmithWaterman.cpp:1556-1722 (ssw_init), :206-240 (createQueryProfile),
// :1094-1322 (sw_sse2_word), :1326-1554 (sw_sse2_int), :264-362 (alignScoreEndPos).
// ---------------------------------------------------------------------------------------------
struct QueryProfile {
    int L;
    std::vector<int8_t> biasAA, biasSS;
    std::vector<int16_t> profAA, prof3Di;  // [letter][pos], row stride L
};

static void build_profile(const Matrix &mAA, const Matrix &m3Di, const uint8_t *qAA, const uint8_t *q3Di, int L,
                          bool biasCorr, float biasScale, QueryProfile &p) {
    p.L = L;
    p.biasAA.assign(L, 0);
    p.biasSS.assign(L, 0);
    if (biasCorr) {
        std::vector<float> tmp(L);
        // NB: both corrections use the AA matrix object (structurealign.cpp:344 passes &subMatAA; SSW.cpp:1565,1570)
        comp_bias(mAA, qAA, L, 1.0f, tmp.data());
        for (int i = 0; i < L; i++) p.biasAA[i] = (int8_t) ((tmp[i] < 0.0) ? tmp[i] - 0.5 : tmp[i] + 0.5);
        comp_bias(mAA, q3Di, L, biasScale, tmp.data());
        for (int i = 0; i < L; i++) p.biasSS[i] = (int8_t) ((tmp[i] < 0.0) ? tmp[i] - 0.5 : tmp[i] + 0.5);
    }
    p.profAA.assign((size_t) A * L, 0);
    p.prof3Di.assign((size_t) A * L, 0);
    for (int t = 0; t < A; t++) for (int i = 0; i < L; i++) {
        // tinySubMat is int8 (structurealign.cpp:272-281); profile word = mat + bias (SSW.cpp:222)
        p.profAA[(size_t) t * L + i] = (int16_t) ((int8_t) mAA.sub[t][qAA[i]] + p.biasAA[i]);
        p.prof3Di[(size_t) t * L + i] = (int16_t) ((int8_t) m3Di.sub[t][q3Di[i]] + p.biasSS[i]);
    }
}

struct SwResult { uint32_t score; int32_t qEnd, dbEnd; int word; };

// Exact semantics of the striped kernels, written row-major.  LANES = 16 (word) or 8 (int), SAT = int16 saturation.
// H(i,q) = max(Hdiag+s, E, F) with the standard F, *except* that E is opened from the H value that only contains
// the F contributions of the same SIMD stripe (the lazy-F loop corrects H but not E; SSW.cpp:1230-1259).
template <int LANES, bool SAT>
static SwResult sw_striped(const QueryProfile &p, const uint8_t *tAA, const uint8_t *t3Di, int Lt, int go, int ge) {
    const int L = p.L;
    const int segLen = (L + LANES - 1) / LANES;
    const int Lpad = segLen * LANES;
    std::vector<int32_t> Hprev(Lpad, 0), Hcur(Lpad, 0), E(Lpad, 0), Hmax(Lpad, 0), Hp(Lpad, 0);
    int32_t best = 0;
    int32_t endRef = 0;       // SSW.cpp:1113 (0 when nothing aligned)
    int32_t endRead = L - 1;  // SSW.cpp:1112
    std::vector<int32_t> slotMax(LANES, 0);  // vMaxScore: per SIMD slot running max
    for (int i = 0; i < Lt; i++) {
        const int16_t *pa = &p.profAA[(size_t) tAA[i] * L];
        const int16_t *ps = &p.prof3Di[(size_t) t3Di[i] * L];
        int32_t Fs = 0;
        for (int q = 0; q < Lpad; q++) {
            if (q % segLen == 0) Fs = 0;  // vF starts at zero in every lane (SSW.cpp:1159)
            int32_t diag = (q == 0) ? 0 : Hprev[q - 1];
            int32_t s = (q < L) ? (int32_t) pa[q] + (int32_t) ps[q] : 0;  // padding rows carry bias 0 (SSW.cpp:222)
            int32_t h = diag + s;
            if (SAT) { if (h > 32767) h = 32767; if (h < -32768) h = -32768; }
            h = std::max(h, E[q]);
            h = std::max(h, Fs);
            Hp[q] = h;
            int32_t x = std::max(h - go, 0);          // unsigned saturating subtract
            E[q] = std::max(std::max(E[q] - ge, 0), x);
            Fs = std::max(std::max(Fs - ge, 0), x);
        }
        // lazy-F loop == full F over the whole column (argument in DESIGN.md §SW semantics)
        int32_t Ff = 0;
        bool changed = false;
        for (int q = 0; q < Lpad; q++) {
            int32_t h = std::max(Hp[q], Ff);
            Hcur[q] = h;
            Ff = std::max(std::max(Ff - ge, 0), std::max(h - go, 0));
            int slot = q / segLen;
            if (h > slotMax[slot]) { slotMax[slot] = h; changed = true; }
        }
        if (changed) {  // SSW.cpp:1263-1276
            int32_t temp = 0;
            for (int l = 0; l < LANES; l++) temp = std::max(temp, slotMax[l]);
            if (temp > best) {
                best = temp;
                endRef = i;
                Hmax = Hcur;
            }
        }
        std::swap(Hprev, Hcur);
    }
    // SSW.cpp:1283-1292: smallest query index holding the maximum in the saved column
    for (int q = 0; q < Lpad; q++) if (Hmax[q] == best && q < endRead) endRead = q;
    SwResult r;
    r.score = (uint32_t) best; r.qEnd = endRead; r.dbEnd = endRef; r.word = 1;
    return r;
}

// alignScoreEndPos (SSW.cpp:264-362): word kernel, int kernel when the word score saturates.
static SwResult align_score_endpos(const QueryProfile &p, const uint8_t *tAA, const uint8_t *t3Di, int Lt, int go, int ge) {
    SwResult r = sw_striped<16, true>(p, tAA, t3Di, Lt, go, ge);
    if (r.score == 32767) {
        r = sw_striped<8, false>(p, tAA, t3Di, Lt, go, ge);
        r.word = 2;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// E-value network.  src/strucclustutils/EvalueNeuralNet.cpp:9-34, EvalueNeuralNet.h:43-79,
// lib/kerasify/keras_model.cpp:157-211 (dense), :99-155 (activation), :632-705 (model).
// ---------------------------------------------------------------------------------------------
struct Dense { int rows, cols; std::vector<float> w, b; int act; };
struct EvalueNN {
    std::vector<Dense> layers;
    double logDbRes;
};

static bool load_nn(const uint8_t *blob, size_t len, EvalueNN &nn) {
    size_t off = 0;
    auto rd = [&](uint32_t &v) { if (off + 4 > len) return false; memcpy(&v, blob + off, 4); off += 4; return true; };
    uint32_t n;
    if (!rd(n)) return false;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t type, r, c, b, act;
        if (!rd(type) || type != 1) return false;  // only dense layers occur in evalue_nn.kerasify
        if (!rd(r) || !rd(c) || !rd(b)) return false;
        Dense d; d.rows = r; d.cols = c;
        d.w.resize((size_t) r * c); d.b.resize(b);
        if (off + 4 * ((size_t) r * c + b) > len) return false;
        memcpy(d.w.data(), blob + off, 4 * (size_t) r * c); off += 4 * (size_t) r * c;
        memcpy(d.b.data(), blob + off, 4 * (size_t) b); off += 4 * (size_t) b;
        if (!rd(act)) return false;
        d.act = act;
        nn.layers.push_back(d);
    }
    return true;
}

static void predict_mu_lambda(const EvalueNN &nn, const uint8_t *q3Di, int L, double &first, double &second) {
    std::vector<float> in(A + 1, 0.0f);
    for (int i = 0; i < L; i++) in[q3Di[i]]++;
    in[A] = L;
    for (const Dense &d : nn.layers) {
        std::vector<float> tmp(d.cols, 0.0f);
        for (int i = 0; i < d.rows; i++) for (int j = 0; j < d.cols; j++) tmp[j] += in[i] * d.w[(size_t) i * d.cols + j];
        for (size_t i = 0; i < d.b.size(); i++) tmp[i] += d.b[i];
        if (d.act == 2) for (float &x : tmp) if (x < 0.0) x = 0.0;  // kRelu
        in = tmp;
    }
    first = in[0] * 0.03260331312698818 + 0.17518475184751847;
    second = in[1] * 0.4353169278257701 + -2.5569312493124934;
}

static double evalue_corr(const EvalueNN &nn, double score, double lambda_, double mu) {
    double h = lambda_ * (score - mu);
    double logP;
    if (h > 10) logP = -h;
    else if (h < -2.5) logP = -exp(-exp(-h));
    else logP = log((1.0 - exp(-exp(-h))));
    return pow(exp(logP + nn.logDbRes), 0.32);
}


// ---------------------------------------------------------------------------------------------
// Stage 1: k-mer prefilter.
// M/src/prefiltering/{IndexBuilder.cpp:56-271, IndexTable.h:135-175,350-403, ExtendedSubstitutionMatrix.cpp:20-72,
// KmerGenerator.cpp:177-286, QueryMatcher.cpp:103-376, CacheFriendlyOperations.cpp, UngappedAlignment.cpp}
// ---------------------------------------------------------------------------------------------
const int KA = 20;                       // k-mer alphabet (X excluded; Prefiltering.cpp:203 setupSplit(alphabetSize-1))
const int KSPAN = 10;
const int KPOS[6] = {0, 1, 3, 5, 8, 9};  // spaced_seed_6 = 1101010011 (M/src/commons/Sequence.h:25)
const int N3 = KA * KA * KA;

struct Entry { uint32_t seqId; uint16_t pos; };

struct CounterResult { uint32_t id; uint16_t diagonal; uint8_t count; };

struct PrefIndex {
    std::vector<uint64_t> offsets;       // 20^6 + 1
    std::vector<Entry> entries;
    std::vector<uint8_t> seqs;           // masked targets (SequenceLookup)
    std::vector<uint64_t> seqOff;
    std::vector<uint32_t> seqLen;
    // 3-mer neighbour table: for each 3-mer row all 8000 (score, index), stable-sorted by score descending
    std::vector<int16_t> score3;
    std::vector<uint32_t> index3;
    int kmerThr;
};

// Masker::maskRepeats + finalizeMasking (M/src/commons/Masker.cpp:85-127); tantan is off for Foldseek (maskMode=0)
static void mask_repeats(uint8_t *seq, unsigned int len, int n) {
    unsigned int repeatCount = 0;
    int start = -1;
    int prev = -1;   // reference starts with '\0' == code 0: a leading run of code 0 ... see note below
    // NB: previousChar is initialised to '\0' (== code of 'A'); a sequence starting with A therefore counts from
    // repeatCount 0 with startOfRepeat -1; reproduced literally:
    prev = 0;
    for (unsigned int pos = 0; pos < len; ++pos) {
        int cur = seq[pos];
        if (cur == prev) {
            repeatCount++;
        } else {
            if (repeatCount > (unsigned int) n) for (unsigned int i = (unsigned int) start; i < pos; ++i) seq[i] = 20;
            repeatCount = 1; start = (int) pos; prev = cur;
        }
    }
    if (repeatCount > (unsigned int) n) for (unsigned int i = (unsigned int) start; i < len; ++i) seq[i] = 20;
}

static void build_3mer_table(const Matrix &m, PrefIndex &ix) {
    ix.score3.assign((size_t) N3 * N3, 0);
    ix.index3.assign((size_t) N3 * N3, 0);
    std::vector<std::pair<short, unsigned int>> tmp(N3);
    for (int i0 = 0; i0 < KA; i0++) for (int i1 = 0; i1 < KA; i1++) for (int i2 = 0; i2 < KA; i2++) {
        const unsigned int iIdx = i0 + i1 * KA + i2 * KA * KA;
        int j = 0;
        // permutation order of createCartesianProduct: first letter outermost (ExtendedSubstitutionMatrix.cpp:112-137)
        for (int j0 = 0; j0 < KA; j0++) for (int j1 = 0; j1 < KA; j1++) for (int j2 = 0; j2 < KA; j2++) {
            tmp[j].first = (short) (m.sub[i0][j0] + m.sub[i1][j1] + m.sub[i2][j2]);
            tmp[j].second = j0 + j1 * KA + j2 * KA * KA;
            j++;
        }
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<short, unsigned int> &a, const std::pair<short, unsigned int> &b) { return a.first > b.first; });
        for (int z = 0; z < N3; z++) { ix.score3[(size_t) iIdx * N3 + z] = tmp[z].first; ix.index3[(size_t) iIdx * N3 + z] = tmp[z].second; }
    }
}

static void build_index(const Matrix &mKmer, const uint8_t *t3Di, const uint64_t *tOff, const uint32_t *tLen, uint32_t n,
                        int kmerThr, int maskNrepeats, PrefIndex &ix) {
    ix.kmerThr = kmerThr;
    ix.seqOff.resize(n); ix.seqLen.assign(tLen, tLen + n);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) { ix.seqOff[i] = total; total += tLen[i]; }
    ix.seqs.resize(total);
    size_t tableSize = 1;
    for (int i = 0; i < 6; i++) tableSize *= KA;
    ix.offsets.assign(tableSize + 1, 0);
    std::vector<std::vector<std::pair<uint32_t, uint16_t>>> perSeq(n);
    for (uint32_t id = 0; id < n; id++) {
        uint8_t *s = ix.seqs.data() + ix.seqOff[id];
        memcpy(s, t3Di + tOff[id], tLen[id]);
        if (maskNrepeats > 0) mask_repeats(s, tLen[id], maskNrepeats);
        std::vector<std::pair<uint32_t, uint16_t>> &v = perSeq[id];
        for (int pos = 0; pos + KSPAN <= (int) tLen[id]; pos++) {
            bool hasX = false; int self = 0; uint32_t idx = 0, pw = 1;
            for (int k = 0; k < 6; k++) {
                const int c = s[pos + KPOS[k]];
                if (c >= KA) hasX = true;
                else { self += (signed char) mKmer.sub[c][c]; idx += c * pw; }   // getScoreLookup truncates to char (IndexBuilder.cpp:11-23)
                pw *= KA;
            }
            if (hasX) continue;
            if (kmerThr > 0 && self < kmerThr) continue;   // IndexTable.h:147-155
            v.push_back(std::make_pair(idx, (uint16_t) pos));
        }
        std::sort(v.begin(), v.end());
        // one entry per distinct k-mer per sequence, smallest position (IndexTable.h:386-402)
        size_t w = 0;
        for (size_t k = 0; k < v.size(); k++) if (k == 0 || v[k].first != v[k - 1].first) v[w++] = v[k];
        v.resize(w);
        for (auto &e : v) ix.offsets[e.first + 1]++;
    }
    for (size_t k = 0; k < tableSize; k++) ix.offsets[k + 1] += ix.offsets[k];
    ix.entries.resize(ix.offsets[tableSize]);
    std::vector<uint64_t> cur(ix.offsets.begin(), ix.offsets.end() - 1);
    for (uint32_t id = 0; id < n; id++) for (auto &e : perSeq[id]) { Entry en; en.seqId = id; en.pos = e.second; ix.entries[cur[e.first]++] = en; }
    // lists come out sorted by (seqId, pos) because ids are visited in order (IndexTable::sortDBSeqLists)
}

struct PrefHit { uint32_t seqId; int score; uint16_t diagonal; };

// CacheFriendlyOperations<BINSIZE> restated with explicit bins (bin = id & (B-1), order inside a bin = insertion order)
struct DiagMatcher {
    unsigned int B;
    std::vector<uint8_t> dup;   // indexed by id >> log2(B)
    unsigned int shift;
    explicit DiagMatcher(unsigned int bins, size_t maxElement) : B(bins) {
        shift = 0; while ((1u << shift) < B) shift++;
        size_t size = (size_t) pow(2, ceil(log((double) maxElement) / log(2)));
        size = std::max(size >> shift, (size_t) 1);
        dup.assign(size + 1, 0);
    }
    void binned(const std::vector<CounterResult> &in, std::vector<std::vector<CounterResult>> &bins) const {
        bins.assign(B, std::vector<CounterResult>());
        for (const CounterResult &e : in) bins[e.id & (B - 1)].push_back(e);
    }
    // findDuplicates (CacheFriendlyOperations.cpp:38-49,185-278), computeTotalScore == false
    void find_duplicates(const std::vector<std::vector<CounterResult>> &bins, std::vector<CounterResult> &out) {
        for (unsigned int b = 0; b < B; b++) {
            std::vector<CounterResult> tmp;
            for (const CounterResult &e : bins[b]) {
                const size_t h = e.id >> shift;
                const uint8_t cur = (uint8_t) e.diagonal;
                const uint8_t prev = dup[h];
                if (cur == prev) tmp.push_back(e);
                dup[h] = cur;
            }
            for (size_t n = tmp.size(); n-- > 0;) dup[tmp[n].id >> shift] = (uint8_t) ((uint8_t) tmp[n].diagonal + 1);
            for (const CounterResult &e : tmp) {
                const size_t h = e.id >> shift;
                CounterResult o = e; o.count = 0;
                if (dup[h] != (uint8_t) e.diagonal) out.push_back(o);
                dup[h] = (uint8_t) e.diagonal;
            }
            for (const CounterResult &e : bins[b]) dup[e.id >> shift] = 0;
        }
    }
    // mergeDiagonalDuplicates (:85-118)
    void merge_diagonal(std::vector<CounterResult> &arr) {
        std::vector<std::vector<CounterResult>> bins; binned(arr, bins);
        std::vector<CounterResult> out;
        for (unsigned int b = 0; b < B; b++) {
            const std::vector<CounterResult> &v = bins[b];
            for (size_t n = v.size(); n-- > 0;) dup[v[n].id >> shift] = (uint8_t) ((uint8_t) v[n].diagonal + 1);
            for (const CounterResult &e : v) {
                const size_t h = e.id >> shift;
                if (dup[h] != (uint8_t) e.diagonal) out.push_back(e);
                dup[h] = (uint8_t) e.diagonal;
            }
        }
        // the reference leaves duplicateBitArray dirty here; the next user (keepMaxElement / findDuplicates) memsets it
        std::fill(dup.begin(), dup.end(), 0);
        arr.swap(out);
    }
    // mergeDiagonalKeepScoredHitsDuplicates (:121-152)
    void merge_diagonal_keep_scored(std::vector<CounterResult> &arr) {
        std::vector<std::vector<CounterResult>> bins; binned(arr, bins);
        std::vector<CounterResult> out;
        for (unsigned int b = 0; b < B; b++) {
            const std::vector<CounterResult> &v = bins[b];
            for (const CounterResult &e : v) dup[e.id >> shift] = (uint8_t) ((uint8_t) e.diagonal + 1);
            for (size_t n = v.size(); n-- > 0;) {
                const CounterResult &e = v[n];
                const size_t h = e.id >> shift;
                if (e.count != 0 || dup[h] != (uint8_t) e.diagonal) out.push_back(e);
                dup[h] = (uint8_t) e.diagonal;
            }
        }
        std::fill(dup.begin(), dup.end(), 0);
        arr.swap(out);
    }
    // keepMaxElement (:354-385)
    void keep_max(std::vector<CounterResult> &arr) {
        std::vector<std::vector<CounterResult>> bins; binned(arr, bins);
        std::vector<CounterResult> out;
        std::fill(dup.begin(), dup.end(), 0);
        for (unsigned int b = 0; b < B; b++) {
            const std::vector<CounterResult> &v = bins[b];
            for (const CounterResult &e : v) { const size_t h = e.id >> shift; dup[h] = std::max(dup[h], e.count); }
            for (const CounterResult &e : v) {
                const size_t h = e.id >> shift;
                const bool found = dup[h] == e.count;
                if (found) out.push_back(e);
                dup[h] = (uint8_t) (dup[h] * (1 - (int) found));
            }
        }
        std::fill(dup.begin(), dup.end(), 0);
        arr.swap(out);
    }
};

// UngappedAlignment::computeSingelSequenceScores + scalarDiagonalScoring (UngappedAlignment.cpp:45-58,423-438), clamp 255
static int ungapped_score(const std::vector<int8_t> &profile, int qLen, const uint8_t *db, int dbLen, uint16_t diagonal) {
    const short diag = (short) diagonal;
    const unsigned short minDist = (unsigned short) std::min((unsigned short) (0 - diagonal), (unsigned short) (diagonal - 0));
    int max = 0;
    const int8_t *p; const uint8_t *d; int len;
    if (diag >= 0 && (int) minDist < qLen) { p = profile.data() + (size_t) minDist * 22; d = db; len = std::min(dbLen, qLen - (int) minDist); }
    else if (diag < 0 && (int) minDist < dbLen) { p = profile.data(); d = db + minDist; len = std::min(dbLen - (int) minDist, qLen); }
    else return 0;
    int score = 0;
    for (int pos = 0; pos < len; pos++) {
        score += p[(size_t) pos * 22 + d[pos]];
        score = (score < 0) ? 0 : score;
        max = (score > max) ? score : max;
    }
    return max;
}

struct PrefStats { uint64_t kmers, dbMatches, overflows, doubleHits; };

// QueryMatcher::matchQuery for one query (diagonalScoring, no aux alphabet, amino-acid type).
static void prefilter_query(const Matrix &mKmer, const Matrix &mUngapped, const PrefIndex &ix, uint32_t nTargets,
                            const uint8_t *q, int L, float biasScale, bool biasCorr, size_t maxHitsPerQueryIn, unsigned int minDiagScoreThr,
                            unsigned int bins, uint32_t identityId, std::vector<PrefHit> &res, PrefStats &st) {
    const size_t dbSize = nTargets;
    const size_t maxDbMatches = std::max((size_t) 1000000, dbSize) * 2;       // QueryMatcher.cpp:45
    const size_t foundDiagonalsSize = std::max((size_t) 1000000, dbSize);     // :44
    const size_t maxHitsPerQuery = std::min(maxHitsPerQueryIn, dbSize);       // :47
    std::vector<float> bias(L, 0.0f);
    if (biasCorr) comp_bias(mKmer, q, L, biasScale, bias.data());
    // UngappedAlignment::createProfile (UngappedAlignment.cpp:388-421): 22 bytes per position
    std::vector<int8_t> profile((size_t) L * 22, 0);
    for (int pos = 0; pos < L; pos++) {
        float b = bias[pos];
        b = (b < 0.0) ? b / 4 - 0.5 : b / 4 + 0.5;
        const int8_t corr = (int8_t) b;
        for (int a = 0; a < A; a++) profile[(size_t) pos * 22 + a] = (int8_t) (mUngapped.sub[q[pos]][a] + corr);
    }
    DiagMatcher dm(bins, dbSize);
    std::vector<CounterResult> found;        // foundDiagonals[0 .. overflowHitCount)
    std::vector<CounterResult> chunk;        // databaseHits of the current chunk as (id, diagonal)
    size_t overflowHitCount = 0;
    bool anyOverflow = false;
    st.kmers = st.dbMatches = st.overflows = st.doubleHits = 0;
    auto score_unscored = [&](std::vector<CounterResult> &v) {
        for (CounterResult &e : v) if (e.count == 0) {
            const int sc = ungapped_score(profile, L, ix.seqs.data() + ix.seqOff[e.id], (int) ix.seqLen[e.id], e.diagonal);
            e.count = (uint8_t) std::min(255, sc);
        }
    };
    auto flush_chunk = [&](std::vector<CounterResult> &dst) {
        std::vector<std::vector<CounterResult>> b; dm.binned(chunk, b);
        dm.find_duplicates(b, dst);
        chunk.clear();
    };
    bool aborted = false;
    for (int i = 0; i + KSPAN <= L && !aborted; i++) {
        float bc = 0;
        for (int k = 0; k < 6; k++) bc += bias[i + KPOS[k]];
        bool hasX = false; uint8_t km[6];
        for (int k = 0; k < 6; k++) { km[k] = q[i + KPOS[k]]; if (km[k] >= KA) hasX = true; }
        if (hasX) continue;
        const short b = (short) ((bc < 0.0) ? bc - 0.5 : bc + 0.5);
        const short thr = (short) std::max(ix.kmerThr - b, 0);
        // KmerGenerator::generateKmerList for k = 6 -> two 3-mers (KmerGenerator.cpp:177-286)
        const unsigned int r0 = km[0] + km[1] * KA + km[2] * KA * KA;
        const unsigned int r1 = km[3] + km[4] * KA + km[5] * KA * KA;
        const int16_t *s0 = &ix.score3[(size_t) r0 * N3]; const uint32_t *i0 = &ix.index3[(size_t) r0 * N3];
        const int16_t *s1 = &ix.score3[(size_t) r1 * N3]; const uint32_t *i1 = &ix.index3[(size_t) r1 * N3];
        const short possibleRest = s1[0];
        const short cutoff1 = (short) (thr - possibleRest);
        for (int a = 0; a < N3 && !aborted; a++) {
            if (s0[a] < cutoff1) break;
            const short cutoff2 = (short) (thr - s0[a]);
            for (int c = 0; c < N3 && s1[c] >= cutoff2; c++) {
                const size_t kmer = (size_t) i0[a] + (size_t) i1[c] * N3;
                st.kmers++;
                const uint64_t o0 = ix.offsets[kmer], o1 = ix.offsets[kmer + 1];
                const size_t listLen = o1 - o0;
                if (chunk.size() + listLen >= maxDbMatches) {   // QueryMatcher.cpp:305
                    anyOverflow = true; st.overflows++;
                    std::vector<CounterResult> d;
                    flush_chunk(d);
                    if (overflowHitCount != 0) {
                        found.insert(found.end(), d.begin(), d.end());
                        dm.merge_diagonal_keep_scored(found);
                        score_unscored(found);
                        dm.keep_max(found);
                    } else {
                        found = d;
                    }
                    overflowHitCount = found.size();
                    if (listLen >= maxDbMatches) { aborted = true; break; }
                }
                for (uint64_t e = o0; e < o1; e++) {
                    CounterResult cr; cr.id = ix.entries[e].seqId; cr.diagonal = (uint16_t) (i - ix.entries[e].pos); cr.count = 0;
                    chunk.push_back(cr);
                }
                st.dbMatches += listLen;
            }
        }
    }
    if (!chunk.empty()) {
        std::vector<CounterResult> d;
        flush_chunk(d);
        found.insert(found.end(), d.begin(), d.end());
        if (overflowHitCount != 0) dm.merge_diagonal(found);
    } else {
        found.clear();   // QueryMatcher.cpp:358-366: hitCount stays 0 when the last chunk is empty (earlier chunks are dropped)
    }
    (void) anyOverflow; (void) foundDiagonalsSize;
    st.doubleHits = found.size();
    // matchQuery tail (QueryMatcher.cpp:129-240)
    score_unscored(found);
    dm.keep_max(found);
    unsigned int scoreSizes[256];
    memset(scoreSizes, 0, sizeof(scoreSizes));
    for (const CounterResult &e : found) scoreSizes[e.count]++;
    unsigned int thr;
    { size_t fh = 0; size_t t; for (t = 255; t > 0; t--) { fh += scoreSizes[t]; if (fh >= maxHitsPerQuery) break; } thr = (unsigned int) t; }
    thr = std::max(minDiagScoreThr, thr);
    // radixSortByScoreSize: stable counting sort, score descending (QueryMatcher.cpp:536-561)
    std::vector<CounterResult> sorted;
    for (int sc = 255; sc >= (int) thr; sc--) for (const CounterResult &e : found) if (e.count == sc) sorted.push_back(e);
    // score truncation path (>= 255 - queryBias): not reachable with the matrices of this path unless > maxHits hits saturate;
    // reported as unsupported by returning a marker score
    res.clear();
    if (identityId != 0xFFFFFFFFu) { PrefHit h; h.seqId = identityId; h.score = 65535; h.diagonal = 0; res.push_back(h); }
    for (size_t k = 0; k < sorted.size() && res.size() < maxHitsPerQuery; k++) {
        const CounterResult &e = sorted[k];
        if (e.count >= thr && e.id != identityId) {
            PrefHit h; h.seqId = e.id; h.score = e.count; h.diagonal = e.diagonal;
            if (e.count >= 255) h.score = ungapped_score(profile, L, ix.seqs.data() + ix.seqOff[e.id], (int) ix.seqLen[e.id], e.diagonal);   // getResult: real score (QueryMatcher.cpp:441-444), unclamped
            res.push_back(h);
        }
    }
    auto cmp = [](const PrefHit &a, const PrefHit &b) {
        if (abs(a.score) > abs(b.score)) return true;
        if (abs(b.score) > abs(a.score)) return false;
        return a.seqId < b.seqId;
    };
    if (res.size() > 1) {
        if (identityId != 0xFFFFFFFFu) std::sort(res.begin() + 1, res.end(), cmp); else std::sort(res.begin(), res.end(), cmp);
    }
}

struct Oracle {
    Matrix mAA, m3Di;        // SW matrices: blosum62 x 1.4, 3di x 2.1 (structurealign.cpp:252-265)
    Matrix mKmer, mUngapped; // prefilter: 3di x 8.0 bias -0.2, 3di x 2.0 bias -0.2 (Prefiltering.cpp:68-69,1022)
    EvalueNN nn;
    bool hasNN = false;
};

}  // namespace

extern "C" {

void *fso_create(const char *mat3di_text, const char *blosum_text, const uint8_t *nn_blob, uint64_t nn_len) {
    Oracle *o = new Oracle();
    if (!build_matrix(mat3di_text, 2.1, 0.0, o->m3Di) || !build_matrix(blosum_text, 1.4, 0.0, o->mAA) ||
        !build_matrix(mat3di_text, 8.0, -0.2f, o->mKmer) || !build_matrix(mat3di_text, 2.0, -0.2f, o->mUngapped)) {
        delete o;
        return NULL;
    }
    if (nn_blob != NULL) o->hasNN = load_nn(nn_blob, nn_len, o->nn);
    return o;
}

void fso_destroy(void *h) { delete (Oracle *) h; }

// which: 0 = SW AA (1.4), 1 = SW 3Di (2.1), 2 = k-mer 3Di (8.0), 3 = ungapped 3Di (2.0)
void fso_get_matrix(void *h, int which, int16_t *out441, double *pback21) {
    Oracle *o = (Oracle *) h;
    const Matrix *m = which == 0 ? &o->mAA : which == 1 ? &o->m3Di : which == 2 ? &o->mKmer : &o->mUngapped;
    for (int i = 0; i < A; i++) for (int j = 0; j < A; j++) out441[i * A + j] = m->sub[i][j];
    for (int i = 0; i < A; i++) pback21[i] = m->pBack[i];
}

void fso_query_bias(void *h, const uint8_t *qAA, const uint8_t *q3Di, int L, float scale, int8_t *biasAA, int8_t *biasSS) {
    Oracle *o = (Oracle *) h;
    QueryProfile p;
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, true, scale, p);
    memcpy(biasAA, p.biasAA.data(), L);
    memcpy(biasSS, p.biasSS.data(), L);
}

void fso_mu_lambda(void *h, uint64_t dbResidues, const uint8_t *q3Di, int L, double *first, double *second) {
    Oracle *o = (Oracle *) h;
    o->nn.logDbRes = log(static_cast<double>(dbResidues));
    predict_mu_lambda(o->nn, q3Di, L, *first, *second);
}

// One query against n targets: forward SW (score, qEnd, dbEnd), reverse-query SW score, the two E-values of
// alignStructure (structurealign.cpp:48-74).  Targets are given as concatenated code arrays + offsets.
// out per target: fwdScore, qEnd, dbEnd, revScore (int32 x4); evalues: e(fwd), e(fwd-rev) (double x2).
void fso_align_query(void *h, uint64_t dbResidues, const uint8_t *qAA, const uint8_t *q3Di, int L, float biasScale,
                     const uint8_t *tAA, const uint8_t *t3Di, const uint64_t *tOff, const uint32_t *tLen, int n,
                     int go, int ge, int32_t *out4, double *evals2) {
    Oracle *o = (Oracle *) h;
    QueryProfile fwd, rev;
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, true, biasScale, fwd);
    std::vector<uint8_t> rAA(qAA, qAA + L), r3Di(q3Di, q3Di + L);
    std::reverse(rAA.begin(), rAA.end());
    std::reverse(r3Di.begin(), r3Di.end());
    // the reverse profile recomputes the bias on the reversed sequence (structurealign.cpp:345-347)
    build_profile(o->mAA, o->m3Di, rAA.data(), r3Di.data(), L, true, biasScale, rev);
    double mu = 0, lam = 0;
    if (o->hasNN) {
        o->nn.logDbRes = log(static_cast<double>(dbResidues));
        predict_mu_lambda(o->nn, q3Di, L, mu, lam);
    }
    for (int k = 0; k < n; k++) {
        const uint8_t *a = tAA + tOff[k];
        const uint8_t *s = t3Di + tOff[k];
        SwResult f = align_score_endpos(fwd, a, s, (int) tLen[k], go, ge);
        SwResult r = align_score_endpos(rev, a, s, (int) tLen[k], go, ge);
        out4[4 * k + 0] = (int32_t) f.score;
        out4[4 * k + 1] = f.qEnd;
        out4[4 * k + 2] = f.dbEnd;
        out4[4 * k + 3] = (int32_t) r.score;
        if (o->hasNN) {
            evals2[2 * k + 0] = evalue_corr(o->nn, (double) f.score, mu, lam);
            evals2[2 * k + 1] = evalue_corr(o->nn, (double) ((int32_t) f.score - (int32_t) r.score), mu, lam);
        }
    }
}


// ---- stage 2 traceback: block-aligner (restated in block_aligner_port.h) ----

// Generic sequence-to-sequence entry used to pin the engine against the crate's own unit-test vectors
// (scan_block.rs:2267-2413). q/r: converted bytes (Matrix::convert_char), table[a*stride+b] int8, nullCode = padding byte.
// flags: bit0 TRACE, bit1 X_DROP, bit2 cigar with =/X. out3: score, query_idx, reference_idx. cigarOut: e.g. "3=2X1=".
void fso_ba_align(const uint8_t *q, int qLen, const uint8_t *r, int rLen, const int8_t *table, int stride, int nullCode,
                  int gapOpen, int gapExtend, int minSize, int maxSize, int xDrop, int flags, int32_t *out3,
                  char *cigarOut, int cigarCap) {
    const bool traceFlag = flags & 1, xdrop = flags & 2, eq = flags & 4;
    std::vector<uint8_t> pq(1 + qLen + maxSize + ba::L, (uint8_t) nullCode), pr(1 + rLen + maxSize + ba::L, (uint8_t) nullCode);
    memcpy(pq.data() + 1, q, qLen);
    memcpy(pr.data() + 1, r, rLen);
    ba::MatrixScorer sc;
    sc.q = pq.data(); sc.r = pr.data(); sc.ql = qLen; sc.rl = rLen; sc.table = table; sc.stride = stride;
    ba::Block b(traceFlag, xdrop);
    b.align(sc, gapOpen, gapExtend, minSize, maxSize, xDrop);
    out3[0] = b.res.score; out3[1] = (int32_t) b.res.query_idx; out3[2] = (int32_t) b.res.reference_idx;
    if (cigarOut != NULL && cigarCap > 0) {
        std::string s;
        if (traceFlag) {
            std::vector<std::pair<ba::Op, size_t>> c = b.trace.cigar(b.res.query_idx, b.res.reference_idx, eq ? pq.data() : NULL, eq ? pr.data() : NULL);
            const char names[] = {'?', 'M', '=', 'X', 'I', 'D'};
            for (size_t k = 0; k < c.size(); k++) { s += std::to_string(c[k].second); s += names[c[k].first]; }
        }
        snprintf(cigarOut, cigarCap, "%s", s.c_str());
    }
}

// alignStartPosBacktraceBlock (src/commons/StructureSmithWaterman.cpp:369-537) for one (query, target) pair whose
// forward pass gave (targetScore, qEnd, dbEnd).  Returns 1 when the block aligner reproduced targetScore (backtrace,
// start positions and identicalAACnt valid), 0 for the reference's "score mismatch" exit (:474-477).
// out4: qStart, dbStart, identicalAACnt, block-aligner score of the last attempt.
int fso_backtrace(void *h, const uint8_t *qAA, const uint8_t *q3Di, int L, float biasScale, const uint8_t *tAA, const uint8_t *t3Di,
                  int tLen, int go, int ge, int qEnd, int dbEnd, int32_t targetScore, int32_t *out4, char *btOut, int btCap) {
    Oracle *o = (Oracle *) h;
    (void) tLen;
    const int MAXS = 4096;  // #define MAX_SIZE 4096 (:377)
    const int NUL = A;      // padding row/col of the 27x32 AAMatrix: every score i8::MIN (scores.rs:52-57,88)
    const int ST = 32;
    QueryProfile p;
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, true, biasScale, p);
    // AAMatrix::set writes both [a][b] and [b][a] (scores.rs:93-103); the caller loops aa1 outer, aa2 inner (:429-446),
    // so the surviving value of the pair {a,b} is subMatrix[max][min]
    std::vector<int8_t> tabAA(ST * ST, INT8_MIN), tab3Di(ST * ST, INT8_MIN);
    for (int a = 0; a < A; a++) for (int b = 0; b < A; b++) {
        tabAA[a * ST + b] = (int8_t) o->mAA.sub[std::max(a, b)][std::min(a, b)];
        tab3Di[a * ST + b] = (int8_t) o->m3Di.sub[std::max(a, b)][std::min(a, b)];
    }
    const int qn = qEnd + 1, tn = dbEnd + 1;
    std::vector<uint8_t> pqa(1 + qn + MAXS + ba::L, NUL), pqs(pqa), pta(1 + tn + MAXS + ba::L, NUL), pts(pta);
    std::vector<int16_t> qb(1 + qn + MAXS + ba::L, 0), tb(1 + tn + MAXS + ba::L, 0);
    for (int i = 0; i < qn; i++) {   // reversed prefix of the query (:394-405)
        pqa[1 + i] = qAA[qEnd - i];
        pqs[1 + i] = q3Di[qEnd - i];
        qb[1 + i] = (int16_t) (p.biasAA[qEnd - i] + p.biasSS[qEnd - i]);
    }
    for (int i = 0; i < tn; i++) {   // reversed prefix of the target (:417-420), bias 0 (:423-425)
        pta[1 + i] = tAA[dbEnd - i];
        pts[1 + i] = t3Di[dbEnd - i];
    }
    ba::Scorer3di sc;
    sc.qAA = pqa.data(); sc.q3Di = pqs.data(); sc.rAA = pta.data(); sc.r3Di = pts.data();
    sc.qBias = qb.data(); sc.rBias = tb.data(); sc.ql = qn; sc.rl = tn;
    sc.tabAA = tabAA.data(); sc.tab3Di = tab3Di.data(); sc.stride = ST;
    ba::Block blk(true, true);
    ba::AlignResult res;
    res.score = -1000000000; res.query_idx = res.reference_idx = (size_t) -1;
    size_t minSize = 32;
    const int gOpen = -go, gExt = -ge;  // (:380-381)
    while (minSize <= (size_t) MAXS && res.score < targetScore) {   // (:458-469)
        const int32_t xDrop = -((int32_t) minSize * gExt + gOpen);
        blk.align(sc, gOpen, gExt, minSize, MAXS, xDrop);
        res = blk.res;
        minSize *= 2;
    }
    out4[0] = -1; out4[1] = -1; out4[2] = 0; out4[3] = res.score;
    if (btCap > 0) btOut[0] = 0;
    if (res.score != targetScore && !(targetScore == INT16_MAX && res.score >= targetScore)) return 0;
    std::vector<std::pair<ba::Op, size_t>> c = blk.trace.cigar(res.query_idx, res.reference_idx);
    std::string bt;
    size_t qPos = 0, tPos = 0;
    int ids = 0;
    for (size_t k = 0; k < c.size(); k++) {   // (:491-514)
        const size_t len = c[k].second;
        if (c[k].first == ba::OP_M) {
            for (size_t j = 0; j < len; j++) if (pqa[1 + qPos + j] == pta[1 + tPos + j]) ids++;
            qPos += len; tPos += len; bt.append(len, 'M');
        } else if (c[k].first == ba::OP_I) { qPos += len; bt.append(len, 'I'); }
        else if (c[k].first == ba::OP_D) { tPos += len; bt.append(len, 'D'); }
    }
    std::reverse(bt.begin(), bt.end());
    out4[0] = (qEnd + 1) - (int) qPos;
    out4[1] = (dbEnd + 1) - (int) tPos;
    out4[2] = ids;
    snprintf(btOut, btCap, "%s", bt.c_str());
    return 1;
}

// ---- LDDT of a backtrace (src/commons/LDDT.cpp:87-215, LDDT.h:102-119) ----
// The reference walks a 15 A grid (14 neighbour-box directions) to enumerate every unordered pair of aligned query residues
// closer than CUTOFF exactly once; the per-residue sums are multiples of 0.25 and therefore exact in float whatever the
// order, so the restatement enumerates the pairs directly. dist() is the fused chain the reference binary executes
// (gcc -O3 -mfma contracts D2 += d*d: fma(dz,dz,fma(dy,dy,dx*dx)), then sqrtf).
static inline float lddt_dist(const float *x, const float *y, const float *z, int a, int b) {
    const float dx = x[a] - x[b], dy = y[a] - y[b], dz = z[a] - z[b];
    float d2 = dx * dx;
    d2 = __builtin_fmaf(dy, dy, d2);
    d2 = __builtin_fmaf(dz, dz, d2);
    return sqrtf(d2);
}

// qca / tca: x[0..L), y[0..L), z[0..L). Returns avgLddtScore; perRes (may be NULL) gets alignLength floats.
double fso_lddt(const float *qca, int qLen, const float *tca, int tLen, int qStart, int tStart, const char *bt, float *perRes, int *alnLenOut) {
    const float CUTOFF = 15.0f;
    const float *qx = qca, *qy = qca + qLen, *qz = qca + 2 * qLen;
    const float *tx = tca, *ty = tca + tLen, *tz = tca + 2 * tLen;
    std::vector<float> norm(qLen, 0.0f);
    for (int col = 0; col < qLen; col++) {      // initQuery (:87-117)
        float cnt = 0.0f;
        for (int row = 0; row < qLen; row++) cnt += (float) ((col != row) && (lddt_dist(qx, qy, qz, row, col) < CUTOFF));
        norm[col] = cnt != 0 ? 1 / cnt : INFINITY;
    }
    std::vector<int> a2q, a2t;                  // constructAlignHashes (:135-156)
    int qi = qStart, ti = tStart;
    for (const char *c = bt; *c; c++) {
        if (*c == 'M') { a2q.push_back(qi); a2t.push_back(ti); qi++; ti++; }
        else if (*c == 'D') ti++;
        else if (*c == 'I') qi++;
    }
    const int n = (int) a2q.size();
    std::vector<float> red(n, 0.0f);
    for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) {   // calculateLddtScores (:158-215)
        const float distance = lddt_dist(qx, qy, qz, a2q[a], a2q[b]);
        if (!(distance < CUTOFF)) continue;
        const float distSub = lddt_dist(tx, ty, tz, a2t[a], a2t[b]);
        const float dl = std::abs(distance - distSub);
        const float sc = 0.25 * ((dl < 0.5) + (dl < 1.0) + (dl < 2.0) + (dl < 4.0));
        red[a] += sc; red[b] += sc;
    }
    for (int a = 0; a < n; a++) red[a] *= norm[a2q[a]];
    float sum = 0.0f;                           // LDDTScoreResult ctor (LDDT.h:102-119)
    int len = n;
    for (int a = 0; a < n; a++) {
        if (std::isnan(red[a])) { len--; if (perRes) perRes[a] = 0; }
        else { sum += red[a]; if (perRes) perRes[a] = red[a]; }
    }
    if (alnLenOut) *alnLenOut = n;
    return (double) (sum / (float) len);
}

// ---- stage 1 ----
void *fso_index_create(void *h, const uint8_t *t3Di, const uint64_t *tOff, const uint32_t *tLen, uint32_t n, float sensitivity) {
    Oracle *o = (Oracle *) h;
    PrefIndex *ix = new PrefIndex();
    // Prefiltering::getKmerThreshold for k = 6, sequence search (Prefiltering.cpp:1084-1086): 163.2 - 8.917 * s
    float base = 163.2;
    const int kmerThr = (int) (base - (sensitivity * 8.917));
    build_3mer_table(o->mKmer, *ix);
    build_index(o->mKmer, t3Di, tOff, tLen, n, kmerThr, 6, *ix);
    return ix;
}
void fso_index_destroy(void *ix) { delete (PrefIndex *) ix; }
int fso_index_kmer_thr(void *ix) { return ((PrefIndex *) ix)->kmerThr; }
uint64_t fso_index_entries(void *ix) { return ((PrefIndex *) ix)->entries.size(); }

// out: up to maxSeqs triples (target, score, diagonal as short); returns the hit count. stats4: kmers, dbMatches, overflows, doubleHits
uint32_t fso_prefilter_query(void *h, void *ixp, uint32_t nTargets, const uint8_t *q3Di, int L, float biasScale, int maxSeqs,
                             int minDiagScore, int bins, uint32_t identityId, int32_t *out3, uint64_t *stats4) {
    Oracle *o = (Oracle *) h;
    PrefIndex *ix = (PrefIndex *) ixp;
    std::vector<PrefHit> res;
    PrefStats st;
    prefilter_query(o->mKmer, o->mUngapped, *ix, nTargets, q3Di, L, biasScale, true, (size_t) maxSeqs, (unsigned int) minDiagScore,
                    (unsigned int) bins, identityId, res, st);
    for (size_t k = 0; k < res.size(); k++) { out3[3 * k] = (int32_t) res[k].seqId; out3[3 * k + 1] = res[k].score; out3[3 * k + 2] = (short) res[k].diagonal; }
    if (stats4) { stats4[0] = st.kmers; stats4[1] = st.dbMatches; stats4[2] = st.overflows; stats4[3] = st.doubleHits; }
    return (uint32_t) res.size();
}

}  // extern "C"
