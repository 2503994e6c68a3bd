// Host-side pieces of the hot path that the reference also computes once per run / once per query on the CPU:
// substitution matrices, composition bias, the E-value network.  Double/float operation order follows the
// reference so that the rounded integer matrices, int8 biases and gate decisions are identical.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace fsb {
namespace host {

constexpr int ALPH = 21;

// Integer score matrix + background of one ".out" matrix.
// lib/mmseqs/src/commons/SubstitutionMatrix.cpp:12-57,327-417; BaseMatrix.cpp:108-162.
struct ScoreMatrix {
    short s[ALPH][ALPH];
    double background[ALPH];
    double lambda;

    bool parse(const char *text, double bitFactor, double scoreBias) {
        if (text == NULL) return false;
        std::istringstream in((std::string(text)));
        std::string line;
        double p[ALPH][ALPH];
        int col[256];
        for (int i = 0; i < 256; i++) col[i] = -1;
        for (int i = 0; i < ALPH; i++) { background[i] = 0.0; for (int j = 0; j < ALPH; j++) p[i][j] = 0.0; }
        bool header = false, gotLambda = false, gotBg = false;
        int nLetters = 0;
        lambda = 0.0;
        while (std::getline(in, line)) {
            std::vector<std::string> tok;
            { std::istringstream ls(line); std::string t; while (ls >> t) tok.push_back(t); }
            if (!line.empty() && line[0] == '#') {
                if (line.compare(0, 36, "# Background (precomputed optional):") == 0) {
                    for (size_t i = 4; i < tok.size() && i - 4 < (size_t) ALPH; i++) background[i - 4] = strtod(tok[i].c_str(), NULL);
                    gotBg = true;
                } else if (line.compare(0, 36, "# Lambda     (precomputed optional):") == 0 && tok.size() > 4) {
                    lambda = strtod(tok[4].c_str(), NULL);
                    gotLambda = true;
                }
                continue;
            }
            if (tok.size() < 2) continue;
            if (!header) {
                nLetters = (int) tok.size();
                if (nLetters != ALPH) return false;   // the hot path is defined for the 21-letter alphabets
                for (int i = 0; i < nLetters; i++) col[toupper((unsigned char) tok[i][0])] = i;
                header = true;
                continue;
            }
            const int row = col[toupper((unsigned char) tok[0][0])];
            if (row < 0) return false;
            for (int i = 0; i < ALPH && i + 1 < (int) tok.size(); i++) p[row][i] = strtod(tok[i + 1].c_str(), NULL);
        }
        if (!header || !gotLambda || !gotBg) return false;
        const int X = ALPH - 1;
        bool xPos = false;
        for (int j = 0; j < ALPH; j++) if (p[X][j] > 0 || p[j][X] > 0) { xPos = true; break; }
        if (!xPos) for (int i = 0; i < ALPH - 1; i++) background[i] = background[i] * (1.0 - background[X]);
        for (int i = 0; i < ALPH; i++) for (int j = 0; j < ALPH; j++) p[i][j] = std::exp(lambda * p[i][j]) * background[i] * background[j];
        double rowSum[ALPH];
        for (int i = 0; i < ALPH; i++) { rowSum[i] = 0; for (int j = 0; j < ALPH; j++) rowSum[i] += p[i][j]; }
        rowSum[X] = 1E-5;
        for (int i = 0; i < ALPH; i++) for (int j = 0; j < ALPH; j++) {
            const double bits = std::log2(p[i][j] / (rowSum[i] * rowSum[j]));
            const double v = bitFactor * bits + scoreBias;
            s[i][j] = (short) ((v < 0.0) ? v - 0.5 : v + 0.5);
        }
        return true;
    }
};

// SubstitutionMatrix::calcLocalAaBiasCorrection (SubstitutionMatrix.cpp:79-109): window [i-20, i+20)
inline void composition_bias(const ScoreMatrix &m, const uint8_t *seq, int n, float scale, float *out) {
    for (int i = 0; i < n; i++) {
        const int lo = std::max(0, i - 20), hi = std::min(n, i + 20);
        const short *row = m.s[seq[i]];
        int acc = 0;
        for (int j = lo; j < hi; j++) acc += row[seq[j]];
        acc -= row[seq[i]];
        float d = (float) acc;
        d /= -1.0 * static_cast<float>(hi - lo);
        for (int a = 0; a < ALPH; a++) d += m.background[a] * static_cast<float>(row[a]);
        out[i] = scale * d;
    }
}

inline int8_t round_bias(float v) { return (int8_t) ((v < 0.0) ? v - 0.5 : v + 0.5); }

// (mu, lambda) network + Gumbel E-value: src/strucclustutils/EvalueNeuralNet.{h,cpp}, lib/kerasify/keras_model.cpp:157-211
struct EvalueModel {
    struct Layer { int in, out, act; std::vector<float> w, b; };
    std::vector<Layer> layers;
    double logDbResidues = 0.0;

    bool load(const uint8_t *blob, uint64_t len) {
        layers.clear();
        uint64_t off = 0;
        auto u32 = [&](uint32_t &v) { if (off + 4 > len) return false; memcpy(&v, blob + off, 4); off += 4; return true; };
        uint32_t n = 0;
        if (!u32(n)) return false;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t type, rows, cols, nb, act;
            if (!u32(type) || type != 1 || !u32(rows) || !u32(cols) || !u32(nb)) return false;
            Layer l; l.in = (int) rows; l.out = (int) cols;
            const uint64_t bytes = 4ull * ((uint64_t) rows * cols + nb);
            if (off + bytes > len) return false;
            l.w.resize((size_t) rows * cols); l.b.resize(nb);
            memcpy(l.w.data(), blob + off, 4ull * rows * cols); off += 4ull * rows * cols;
            memcpy(l.b.data(), blob + off, 4ull * nb); off += 4ull * nb;
            if (!u32(act)) return false;
            l.act = (int) act;
            layers.push_back(l);
        }
        return !layers.empty() && layers[0].in == ALPH + 1 && layers.back().out == 2;
    }

    void mu_lambda(const uint8_t *tdi, int n, double &first, double &second) const {
        std::vector<float> x(ALPH + 1, 0.0f);
        for (int i = 0; i < n; i++) x[tdi[i]]++;
        x[ALPH] = n;
        for (const Layer &l : layers) {
            std::vector<float> y(l.out, 0.0f);
            for (int i = 0; i < l.in; i++) for (int j = 0; j < l.out; j++) y[j] += x[i] * l.w[(size_t) i * l.out + j];
            for (size_t i = 0; i < l.b.size(); i++) y[i] += l.b[i];
            if (l.act == 2) for (float &v : y) if (v < 0.0) v = 0.0;
            x.swap(y);
        }
        first = x[0] * 0.03260331312698818 + 0.17518475184751847;
        second = x[1] * 0.4353169278257701 + -2.5569312493124934;
    }

    // computeEvalueCorr(score, lambda_ = first, mu = second)  (EvalueNeuralNet.h:73-79)
    double evalue(double score, double first, double second) const {
        const double h = first * (score - second);
        double logP;
        if (h > 10) logP = -h;
        else if (h < -2.5) logP = -exp(-exp(-h));
        else logP = log((1.0 - exp(-exp(-h))));
        return pow(exp(logP + logDbResidues), 0.32);
    }

    // Smallest integer score whose E-value passes `evalue <= thr`. The E-value is monotone in the integer score
    // (DESIGN.md "E-value gate"), so the two gates of alignStructure become integer comparisons on the device side.
    int min_passing_score(double first, double second, double thr) const {
        int lo = -70000, hi = 70000;   // evalue(lo) > thr expected, evalue(hi) <= thr expected
        if (!(evalue(hi, first, second) <= thr)) return hi + 1;
        if (evalue(lo, first, second) <= thr) return lo;
        while (hi - lo > 1) {
            const int mid = lo + (hi - lo) / 2;
            if (evalue(mid, first, second) <= thr) hi = mid; else lo = mid;
        }
        return hi;
    }
};

}  // namespace host
}  // namespace fsb
