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

}  // extern "C"
