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
    res.clear();
    if (identityId != 0xFFFFFFFFu) { PrefHit h; h.seqId = identityId; h.score = 65535; h.diagonal = 0; res.push_back(h); }
    // score truncation (QueryMatcher.cpp:186-199): when the max-seqs cut falls inside the saturated bucket (thr >= UCHAR_MAX -
    // getQueryBias(), the bias is 0), the saturated hits are rescored, rescaled to 8 bits relative to the query's self score
    // (rescoreHits :563-589), sorted again and reported as 255 + count * (selfScore - 255) / 255 (getResult :436-440)
    const unsigned int maxDiagonalScoreThr = 255;
    if (thr >= maxDiagonalScoreThr) {
        int maxSelf = ungapped_score(profile, L, q, L, 0) - (int) maxDiagonalScoreThr;
        maxSelf = std::max(1, maxSelf);
        maxSelf = std::min(maxSelf, 65535);
        const float fltMaxSelf = static_cast<float>(maxSelf);
        std::vector<CounterResult> resc;
        for (size_t k = 0; k < sorted.size() && sorted[k].count >= maxDiagonalScoreThr; k++) {
            CounterResult e = sorted[k];
            unsigned int ns = (unsigned int) ungapped_score(profile, L, ix.seqs.data() + ix.seqOff[e.id], (int) ix.seqLen[e.id], e.diagonal);
            ns -= maxDiagonalScoreThr;
            const float sc = static_cast<float>(std::min(ns, 65535u));
            e.count = static_cast<unsigned char>((sc / fltMaxSelf) * static_cast<float>(255) + 0.5);
            resc.push_back(e);
        }
        std::vector<CounterResult> resorted;
        for (int sc = 255; sc >= 0; sc--) for (const CounterResult &e : resc) if (e.count == sc) resorted.push_back(e);
        for (size_t k = 0; k < resorted.size() && res.size() < maxHitsPerQuery; k++) {
            const CounterResult &e = resorted[k];
            if (e.id != identityId) {
                PrefHit h; h.seqId = e.id; h.diagonal = e.diagonal;
                h.score = (int) (255u + ((unsigned int) e.count * (unsigned int) maxSelf / 255u));
                res.push_back(h);
            }
        }
    } else
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

// aaFactor: bit factor of the AA matrix, 1.4 for --alignment-type 2 (3Di+AA) and 0.0 for --alignment-type 0 (structurealign.cpp:264-265)
void *fso_create2(const char *mat3di_text, const char *blosum_text, float aaFactor, const uint8_t *nn_blob, uint64_t nn_len) {
    Oracle *o = new Oracle();
    if (!build_matrix(mat3di_text, 2.1, 0.0, o->m3Di) || !build_matrix(blosum_text, aaFactor, 0.0, o->mAA) ||
        !build_matrix(mat3di_text, 8.0, -0.2f, o->mKmer) || !build_matrix(mat3di_text, 2.0, -0.2f, o->mUngapped)) {
        delete o;
        return NULL;
    }
    if (nn_blob != NULL) o->hasNN = load_nn(nn_blob, nn_len, o->nn);
    return o;
}

void *fso_create(const char *mat3di_text, const char *blosum_text, const uint8_t *nn_blob, uint64_t nn_len) {
    return fso_create2(mat3di_text, blosum_text, 1.4f, nn_blob, nn_len);
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
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, scale >= 0.0f, scale, p);
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
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, biasScale >= 0.0f, biasScale, fwd);   // a negative scale stands for --comp-bias-corr 0
    std::vector<uint8_t> rAA(qAA, qAA + L), r3Di(q3Di, q3Di + L);
    std::reverse(rAA.begin(), rAA.end());
    std::reverse(r3Di.begin(), r3Di.end());
    // the reverse profile recomputes the bias on the reversed sequence (structurealign.cpp:345-347)
    build_profile(o->mAA, o->m3Di, rAA.data(), r3Di.data(), L, biasScale >= 0.0f, biasScale, rev);
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

// The 3Di+AA entry of the engine (block_align_3di_aa_trace_xdrop -> align_3di -> place_block_3di, scan_block.rs:973-1013,
// 1302-1448) on converted byte strings with caller-supplied tables and position biases.  Exists so that the crate's own unit-test
// vectors can be pushed through the 3Di+AA code path: with one of the two tables all-zero and zero biases the cell score
// adds(adds(sAA, s3Di), bias) (scan_block.rs:1343-1351) degenerates to the single-matrix score of place_block, so the
// crate's expected values apply unchanged (tests/test_block_aligner_oracle.py).
void fso_ba_align_3di(const uint8_t *qAA, const uint8_t *q3Di, const int16_t *qBias, int qLen, const uint8_t *rAA, const uint8_t *r3Di,
                      const int16_t *rBias, int rLen, const int8_t *tabAA, const int8_t *tab3Di, int stride, int nullCode,
                      int gapOpen, int gapExtend, int minSize, int maxSize, int xDrop, int flags, int32_t *out3, char *cigarOut, int cigarCap) {
    const bool traceFlag = flags & 1, xdrop = flags & 2, eq = flags & 4;
    std::vector<uint8_t> pqa(1 + qLen + maxSize + ba::L, (uint8_t) nullCode), pqs(pqa), pra(1 + rLen + maxSize + ba::L, (uint8_t) nullCode), prs(pra);
    std::vector<int16_t> qb(1 + qLen + maxSize + ba::L, 0), rb(1 + rLen + maxSize + ba::L, 0);
    memcpy(pqa.data() + 1, qAA, qLen); memcpy(pqs.data() + 1, q3Di, qLen);
    memcpy(pra.data() + 1, rAA, rLen); memcpy(prs.data() + 1, r3Di, rLen);
    for (int i = 0; i < qLen; i++) qb[1 + i] = qBias ? qBias[i] : 0;
    for (int i = 0; i < rLen; i++) rb[1 + i] = rBias ? rBias[i] : 0;
    ba::Scorer3di sc;
    sc.qAA = pqa.data(); sc.q3Di = pqs.data(); sc.rAA = pra.data(); sc.r3Di = prs.data();
    sc.qBias = qb.data(); sc.rBias = rb.data(); sc.ql = qLen; sc.rl = rLen;
    sc.tabAA = tabAA; sc.tab3Di = tab3Di; sc.stride = stride;
    ba::Block b(traceFlag, xdrop);
    b.align(sc, gapOpen, gapExtend, minSize, maxSize, xDrop);
    out3[0] = b.res.score; out3[1] = (int32_t) b.res.query_idx; out3[2] = (int32_t) b.res.reference_idx;
    if (cigarOut != NULL && cigarCap > 0) {
        std::string s;
        if (traceFlag) {
            std::vector<std::pair<ba::Op, size_t>> c = b.trace.cigar(b.res.query_idx, b.res.reference_idx, eq ? pqa.data() : NULL, eq ? pra.data() : NULL);
            const char names[] = {'?', 'M', '=', 'X', 'I', 'D'};
            for (size_t k = 0; k < c.size(); k++) { s += std::to_string(c[k].second); s += names[c[k].first]; }
        }
        snprintf(cigarOut, cigarCap, "%s", s.c_str());
    }
}

// simd_prefix_scan_i16 of the AVX2 build on one vector (avx2.rs:312-337), exposed for the crate's test_prefix_scan vectors
void fso_ba_prefix_scan(const int16_t *in16, int gap, int16_t *out16) {
    ba::V r = ba::load(in16), g = ba::set1((int16_t) gap), all, consts;
    ba::prefix_scan_consts(g, all, consts);
    ba::store(out16, ba::prefix_scan(r, g, consts));
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
    build_profile(o->mAA, o->m3Di, qAA, q3Di, L, biasScale >= 0.0f, biasScale, p);
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

// ---- TM-score of a backtrace: TMaligner::computeAppoximateTMscore (src/commons/TMaligner.cpp:50-104) ----
// Float operation order follows the reference *binary* built by oracle/Makefile (gcc -O3 -mavx2 -mfma contracts mul+add of the
// vector intrinsics): do_rotation is t + fma(u2,z, fma(u0,x, u1*y)) (lib/tmalign/basic_fun.h:73-122), score_fun8's squared
// distance is fma(dz,dz, fma(dx,dx, dy*dy)) (lib/tmalign/TMalign.cpp:76-141), the covariance correction of the quaternion Kabsch
// is fma(-mean1, sum2, crossSum) (lib/tmalign/Kabsch.h:474-722). Sums of the AVX Kabsch are kept in its 4-lane / hadd order.
namespace tmo {
// a float product that the compiler must not contract with a following add/sub (the reference computes these products with
// vector multiplies whose results pass through a permute or hadd before the addition, so they are rounded on their own)
static inline float mul(float a, float b) { float r = a * b; __asm__ volatile("" : "+x"(r)); return r; }
struct Xyz { std::vector<float> x, y, z; void resize(size_t n) { x.assign(n, 0.f); y.assign(n, 0.f); z.assign(n, 0.f); } };

// Double section of rmsd_uncentered_avx (Kabsch.h:661-705: eigenvalues of the 3x3 Gram matrix in closed form) and rmatrix<double>
// (Kabsch.h:80-120: rotation from the adjugate of K - ev*I as a quaternion), with the fused multiply-adds spelled out exactly as
// gcc contracted them in the reference binary (disassembly of rmsd_uncentered_avx / rmatrix<double> in oracle/_ref/obj/.../TMalign.o;
// verified bit-identical to the compiler-contracted source form on 1.8 M rotation entries). frr6 = r0r0 r0r1 r1r1 r0r2 r1r2 r2r2.
__attribute__((optimize("-ffp-contract=off")))
static void eig_rot(const float r[9], float detf, const float frr6[6], double ud[3][3]) {
    const double det = detf;
    const double rr0 = frr6[0], rr1 = frr6[1], rr2 = frr6[2], rr3 = frr6[3], rr4 = frr6[4], rr5 = frr6[5];
    const double inv3 = 1.0 / 3.0, SQRT3 = 1.732050807568877;
    const double spur = ((rr0 + rr2) + rr5) * inv3;
    double t = rr2 * rr5;
    t = fma(-rr4, rr4, t);
    t = fma(rr5, rr0, t);
    t = fma(-rr3, rr3, t);
    t = fma(rr0, rr2, t);
    t = fma(-rr1, rr1, t);
    const double cof = t * inv3;
    double e0 = spur, e1 = spur, e2 = spur;
    const double h = (spur > 0) ? fma(spur, spur, -cof) : -1.0;
    if (h > 0) {
        const double detsq = det * det;
        double g = fma(cof, spur, -detsq);
        g = fma(g, 0.5, -(spur * h));
        const double sqrth = sqrt(h);
        double d1 = fma(h * h, h, -(g * g));
        d1 = (d1 < 0) ? atan2(0, -g) * inv3 : atan2(sqrt(d1), -g) * inv3;
        const double c = cos(d1), sn = sin(d1);
        const double cth = sqrth * c, k3 = sqrth * SQRT3;
        e0 = fma(cth, 2.0, spur);
        e1 = fma(k3, sn, -cth) + spur;
        e2 = (-fma(k3, sn, cth)) + spur;
    }
    e0 = (e0 < 0) ? 0 : sqrt(e0); e1 = (e1 < 0) ? 0 : sqrt(e1); e2 = (e2 < 0) ? 0 : sqrt(e2);
    const double ev = (det < 0) ? (e0 + e1) - e2 : e2 + (e0 + e1);
    const double r00 = r[0], r01 = r[1], r02 = r[2], r10 = r[3], r11 = r[4], r12 = r[5], r20 = r[6], r21 = r[7], r22 = r[8];
    const double a01 = r12 - r21, a02 = r20 - r02, a03 = r01 - r10, a12 = r01 + r10, a13 = r20 + r02, a23 = r12 + r21;
    const double a11 = ((r00 - r11) - r22) - ev, a22 = ((r11 - r00) - r22) - ev, a33 = ((-r00 - r11) + r22) - ev;
    const double p2233 = fma(a22, a33, -(a23 * a23));
    const double p1233 = fma(a12, a33, -(a13 * a23));
    const double p1223 = fma(a12, a23, -(a13 * a22));
    const double p0232 = fma(a02, a23, -(a03 * a22));
    const double p0233 = fma(a33, a02, -(a03 * a23));
    const double p0231 = fma(a02, a13, -(a03 * a12));
    const double q0 = fma(a13, p1223, fma(a11, p2233, -(a12 * p1233)));
    const double q1 = fma(-a13, p0232, fma(-p2233, a01, a12 * p0233));
    const double q2 = fma(a13, p0231, fma(p1233, a01, -(a11 * p0233)));
    const double q3 = fma(-a12, p0231, fma(-a01, p1223, a11 * p0232));
    const double q00 = q0 * q0, q11 = q1 * q1, q22 = q2 * q2, q33 = q3 * q3;
    const double il = 1.0 / (((q00 + q11) + q22) + q33);
    const double q12 = q1 * q2, q13 = q1 * q3, q23 = q2 * q3;
    const double xj = q11 * il, az = (q0 * q3) * il, ay = (q0 * q2) * il, ax = (q0 * q1) * il;
    const double ajmxj = fma(q00, il, -xj);
    ud[0][0] = fma(-q33, il, fma(-q22, il, fma(q00, il, xj)));
    ud[0][1] = 2.0 * fma(q12, il, az);
    ud[0][2] = 2.0 * fma(q13, il, -ay);
    ud[1][0] = 2.0 * fma(q12, il, -az);
    ud[1][1] = fma(-q33, il, fma(q22, il, ajmxj));
    ud[1][2] = 2.0 * fma(q23, il, ax);
    ud[2][0] = 2.0 * fma(q13, il, ay);
    ud[2][1] = 2.0 * fma(q23, il, -ax);
    ud[2][2] = fma(q33, il, fma(-q22, il, ajmxj));
}

// kabsch_quat_soa_avx + rmsd_uncentered_avx (Kabsch.h:474-758) + KabschFast (TMalign.cpp:191-222): rotation u, translation t that
// superpose set 1 (a) onto set 2 (b), n points. Returns false when the rotation is NaN (the reference then falls back to Kabsch()).
static bool kabsch_fast(const Xyz &a, const Xyz &b, int n, float t[3], float u[3][3]) {
    const int up = (n % 8) ? (n / 8) * 8 + 8 : n;
    // 16 accumulators of 4 lanes: index by name
    float A[16][4];
    memset(A, 0, sizeof(A));
    enum { S1X, S2X, SXX, SSQ, S1Y, S2Y, S1Z, S2Z, SXY, SYZ, SXZ, SZX, SYX, SZY, SYY, SZZ };
    auto at = [&](const std::vector<float> &v, int i) { return i < n ? v[i] : 0.0f; };
    for (int i = 0; i < up; i += 8) {
        for (int half = 0; half < 2; half++) for (int j = 0; j < 4; j++) {
            const int k = i + 4 * half + j;
            const float x1 = at(a.x, k), y1 = at(a.y, k), z1 = at(a.z, k), x2 = at(b.x, k), y2 = at(b.y, k), z2 = at(b.z, k);
            A[S1X][j] += x1; A[S2X][j] += x2; A[SXX][j] += mul(x1, x2); A[SSQ][j] += mul(x1, x1);
            A[SXY][j] += mul(x1, y2); A[SYZ][j] += mul(y1, z2); A[SXZ][j] += mul(x1, z2); A[SZX][j] += mul(z1, x2);
            A[S1Y][j] += y1; A[S2Y][j] += y2; A[S1Z][j] += z1; A[S2Z][j] += z2;
            A[SYX][j] += mul(y1, x2); A[SZY][j] += mul(z1, y2); A[SYY][j] += mul(y1, y2); A[SZZ][j] += mul(z1, z2);
        }
        // (the remaining terms of ssq are added after both halves; ssq only feeds the rmsd, which the TM-score search ignores)
    }
    auto h4 = [&](int k) { return (A[k][0] + A[k][1]) + (A[k][2] + A[k][3]); };
    const float fn = (float) n, inv = 1.0f / fn;
    const float sxx = h4(SXX), s1x = h4(S1X), s1y = h4(S1Y), s1z = h4(S1Z), s2x = h4(S2X), s2y = h4(S2Y), s2z = h4(S2Z);
    const float sxy = h4(SXY), sxz = h4(SXZ), syx = h4(SYX), syy = h4(SYY), syz = h4(SYZ), szx = h4(SZX), szy = h4(SZY), szz = h4(SZZ);
    const float m1x = mul(s1x, inv), m1y = mul(s1y, inv), m1z = mul(s1z, inv), m2x = mul(s2x, inv), m2y = mul(s2y, inv), m2z = mul(s2z, inv);
    float r[9];
    r[0] = sxx - mul(m1x, s2x);                   // plain mul, permute, sub
    r[1] = __builtin_fmaf(-m1x, s2y, sxy);        // vfnmadd132ps
    r[2] = __builtin_fmaf(-m1x, s2z, sxz);
    r[3] = __builtin_fmaf(-m1y, s2x, syx);
    r[4] = __builtin_fmaf(-m1y, s2y, syy);
    r[5] = __builtin_fmaf(-m2z, s1y, syz);
    r[6] = __builtin_fmaf(-m2x, s1z, szx);
    r[7] = __builtin_fmaf(-m2y, s1z, szy);
    r[8] = __builtin_fmaf(-m2z, s1z, szz);
    // frr = r0r0, r0r1, r1r1, det | r1r2, r0r2, r2r2 with the hadd tree (0 + a) + (b + c)
    auto dot3 = [](float a0, float a1, float a2, float b0, float b1, float b2) { return (0.0f + mul(a0, b0)) + (mul(a1, b1) + mul(a2, b2)); };
    const float r0r0 = dot3(r[0], r[1], r[2], r[0], r[1], r[2]), r0r1 = dot3(r[0], r[1], r[2], r[3], r[4], r[5]);
    const float r1r1 = dot3(r[3], r[4], r[5], r[3], r[4], r[5]), r1r2 = dot3(r[3], r[4], r[5], r[6], r[7], r[8]);
    const float r0r2 = dot3(r[0], r[1], r[2], r[6], r[7], r[8]), r2r2 = dot3(r[6], r[7], r[8], r[6], r[7], r[8]);
    const float c1 = mul(r[4], r[8]) - mul(r[7], r[5]), c2 = mul(r[5], r[6]) - mul(r[8], r[3]), c3 = mul(r[3], r[7]) - mul(r[6], r[4]);
    const float detf = (0.0f + mul(r[0], c1)) + (mul(r[1], c2) + mul(r[2], c3));
    const float frr6[6] = {r0r0, r0r1, r1r1, r0r2, r1r2, r2r2};
    double ud[3][3];
    eig_rot(r, detf, frr6, ud);
    // rm = transposed rotation + translation w = c2 - U c1 through R34v4_sse3 (products, then (p0 + p1) + (p2 + c2))
    for (int i = 0; i < 3; i++) {
        const float n0 = static_cast<float>(-ud[0][i]), n1 = static_cast<float>(-ud[1][i]), n2 = static_cast<float>(-ud[2][i]);
        const float c2i = i == 0 ? m2x : i == 1 ? m2y : m2z;
        t[i] = (mul(n0, m1x) + mul(n1, m1y)) + (mul(n2, m1z) + mul(c2i, 1.0f));
        u[i][0] = (float) ud[0][i]; u[i][1] = (float) ud[1][i]; u[i][2] = (float) ud[2][i];
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) if (std::isnan(u[i][j])) return false;
    return true;
}

static inline void rotate(const Xyz &x, Xyz &y, int len, const float t[3], const float u[3][3]) {   // BasicFunction::do_rotation
    for (int i = 0; i < len; i++) {
        y.x[i] = t[0] + __builtin_fmaf(u[0][2], x.z[i], __builtin_fmaf(u[0][0], x.x[i], u[0][1] * x.y[i]));
        y.y[i] = t[1] + __builtin_fmaf(u[1][2], x.z[i], __builtin_fmaf(u[1][0], x.x[i], u[1][1] * x.y[i]));
        y.z[i] = t[2] + __builtin_fmaf(u[2][2], x.z[i], __builtin_fmaf(u[2][0], x.x[i], u[2][1] * x.y[i]));
    }
}

// score_fun8 (TMalign.cpp:76-141)
static int score_fun8(const Xyz &xa, const Xyz &ya, int n, float d, std::vector<int> &iAli, float &score1, float Lnorm, float score_d8, float d0) {
    float dTmp = d * d;
    const float d02 = d0 * d0, cut = score_d8 * score_d8;
    int nCut, inc = 0;
    float sum;
    for (;;) {
        nCut = 0; sum = 0;
        for (int i = 0; i < n; i++) {
            const float dx = xa.x[i] - ya.x[i], dy = xa.y[i] - ya.y[i], dz = xa.z[i] - ya.z[i];
            const float di = __builtin_fmaf(dz, dz, __builtin_fmaf(dx, dx, dy * dy));
            const float term = (di < cut) ? 1.0f / (1.0f + di / d02) : 0.0f;
            iAli[nCut] = i;
            nCut += (di < dTmp);
            sum += term;
        }
        if (nCut < 3 && n > 3) { inc++; const double dinc = (d + inc * 0.5); dTmp = dinc * dinc; }
        else break;
    }
    score1 = sum / Lnorm;
    return nCut;
}

// TMscore8_search_standard (TMalign.cpp:394-554)
static double search_standard(const Xyz &xtm, const Xyz &ytm, int Lali, float t0[3], float u0[3][3], int simplifyStep, float localD0Search,
                              float score_d8, float d0, bool &nanSeen) {
    Xyz r1, r2, xt;
    r1.resize(Lali + 8); r2.resize(Lali + 8); xt.resize(Lali + 8);
    std::vector<int> kAli(Lali + 1), iAli(Lali + 1);
    float t[3], u[3][3];
    const int nIt = 20, nInitMax = 6;
    int Lini[6], LiniMin = 4;
    if (Lali < LiniMin) LiniMin = Lali;
    int nInit = 0, i;
    for (i = 0; i < nInitMax - 1; i++) {
        nInit++;
        Lini[i] = (int) (Lali / pow(2.0, (double) i));
        if (Lini[i] <= LiniMin) { Lini[i] = LiniMin; break; }
    }
    if (i == nInitMax - 1) { nInit++; Lini[i] = LiniMin; }
    float scoreMax = -1, score;
    for (int iInit = 0; iInit < nInit; iInit++) {
        const int Lfrag = Lini[iInit], iLmax = Lali - Lfrag;
        i = 0;
        for (;;) {
            int ka = 0;
            for (int k = 0; k < Lfrag; k++) {
                const int kk = k + i;
                r1.x[k] = xtm.x[kk]; r1.y[k] = xtm.y[kk]; r1.z[k] = xtm.z[kk];
                r2.x[k] = ytm.x[kk]; r2.y[k] = ytm.y[kk]; r2.z[k] = ytm.z[kk];
                kAli[ka++] = kk;
            }
            if (!kabsch_fast(r1, r2, Lfrag, t, u)) nanSeen = true;
            rotate(xtm, xt, Lali, t, u);
            float d = localD0Search - 1;
            int nCut = score_fun8(xt, ytm, Lali, d, iAli, score, (float) Lali, score_d8, d0);
            if (score > scoreMax) { scoreMax = score; memcpy(t0, t, sizeof(t)); memcpy(u0, u, sizeof(u)); }
            d = localD0Search + 1;
            for (int it = 0; it < nIt; it++) {
                ka = 0;
                for (int k = 0; k < nCut; k++) {
                    const int m = iAli[k];
                    r1.x[k] = xtm.x[m]; r1.y[k] = xtm.y[m]; r1.z[k] = xtm.z[m];
                    r2.x[k] = ytm.x[m]; r2.y[k] = ytm.y[m]; r2.z[k] = ytm.z[m];
                    kAli[ka++] = m;
                }
                if (!kabsch_fast(r1, r2, nCut, t, u)) nanSeen = true;
                rotate(xtm, xt, Lali, t, u);
                nCut = score_fun8(xt, ytm, Lali, d, iAli, score, (float) Lali, score_d8, d0);
                if (score > scoreMax) { scoreMax = score; memcpy(t0, t, sizeof(t)); memcpy(u0, u, sizeof(u)); }
                if (nCut == ka) {
                    int k = 0;
                    for (; k < nCut; k++) if (iAli[k] != kAli[k]) break;
                    if (k == nCut) break;
                }
            }
            if (i < iLmax) { i += simplifyStep; if (i > iLmax) i = iLmax; }
            else if (i >= iLmax) break;
        }
    }
    return scoreMax;
}
}  // namespace tmo

// qca / tca: x[0..L), y[0..L), z[0..L). flags (may be NULL): bit 0 set when a Kabsch rotation was NaN (reference fallback not restated).
double fso_tmscore(const float *qca, int qLen, const float *tca, int tLen, int qStart, int tStart, const char *bt, int normLen, int *flags) {
    using namespace tmo;
    std::vector<int> invmap(qLen, -1);
    int qi = qStart, ti = tStart;
    for (const char *c = bt; *c; c++) {
        if (*c == 'M') { invmap[qi] = ti; qi++; ti++; }
        else if (*c == 'I') qi++;
        else ti++;
    }
    // parameter_set4search(normLen, normLen, ...) (TMalign.cpp:25-47)
    float D0_MIN, Lnorm, score_d8, d0, d0_search;
    Lnorm = (float) normLen;
    if (Lnorm <= 19) d0 = 0.168; else d0 = (1.24 * pow((Lnorm * 1.0 - 15), 1.0 / 3) - 1.8);
    D0_MIN = d0 + 0.8;
    d0 = D0_MIN;
    d0_search = d0;
    if (d0_search > 8) d0_search = 8;
    if (d0_search < 4.5) d0_search = 4.5;
    score_d8 = 1.5 * pow(Lnorm * 1.0, 0.3) + 3.5;
    const double prevd0 = d0;
    const int prevLnorm = Lnorm;
    const double localD0Search = d0_search;
    // aligned coordinates: x = target, y = query (computeAppoximateTMscore passes target as x)
    Xyz xtm, ytm;
    int nAl = 0;
    for (int j = 0; j < qLen; j++) if (invmap[j] >= 0) nAl++;
    xtm.resize(nAl + 8); ytm.resize(nAl + 8);
    nAl = 0;
    for (int j = 0; j < qLen; j++) {
        const int i = invmap[j];
        if (i < 0) continue;
        xtm.x[nAl] = tca[i]; xtm.y[nAl] = tca[tLen + i]; xtm.z[nAl] = tca[2 * tLen + i];
        ytm.x[nAl] = qca[j]; ytm.y[nAl] = qca[qLen + j]; ytm.z[nAl] = qca[2 * qLen + j];
        nAl++;
    }
    bool nanSeen = false;
    float t[3], u[3][3];
    // standard_TMscore (TMalign.cpp:1366-1427)
    float sD0;
    {
        float dmin = 0.5;
        if (Lnorm > 21) sD0 = (1.24 * pow((Lnorm * 1.0 - 15), 1.0 / 3) - 1.8); else sD0 = dmin;
        if (sD0 < dmin) sD0 = dmin;
    }
    const double d0Input = sD0;
    {
        Xyz r1 = xtm, r2 = ytm;
        if (!kabsch_fast(r1, r2, nAl, t, u)) nanSeen = true;
    }
    double tmscore = search_standard(xtm, ytm, nAl, t, u, 40, (float) d0Input, score_d8, sD0, nanSeen);
    const double TMaln = tmscore * nAl / (1.0 * Lnorm);
    // detailed_search_standard (TMalign.cpp:589-626) with the search parameters restored (TMaligner.cpp:96-100)
    d0 = (float) prevd0; Lnorm = (float) prevLnorm;
    float tm2 = (float) search_standard(xtm, ytm, nAl, t, u, 40, (float) localD0Search, score_d8, d0, nanSeen);
    tm2 = tm2 * nAl / Lnorm;
    if (flags) *flags = nanSeen ? 1 : 0;
    return std::max((double) tm2, TMaln);
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
    prefilter_query(o->mKmer, o->mUngapped, *ix, nTargets, q3Di, L, biasScale, biasScale >= 0.0f, (size_t) maxSeqs, (unsigned int) minDiagScore,
                    (unsigned int) bins, identityId, res, st);
    for (size_t k = 0; k < res.size(); k++) { out3[3 * k] = (int32_t) res[k].seqId; out3[3 * k + 1] = res[k].score; out3[3 * k + 2] = (short) res[k].diagonal; }
    if (stats4) { stats4[0] = st.kmers; stats4[1] = st.dbMatches; stats4[2] = st.overflows; stats4[3] = st.doubleHits; }
    return (uint32_t) res.size();
}

}  // extern "C"
