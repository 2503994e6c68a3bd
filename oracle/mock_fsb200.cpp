// TEST INFRASTRUCTURE — never shipped, never linked into libfsb200.so or foldseek-b200.
//
// A CPU stand-in for the C ABI of include/fsb200.h, implemented with the oracle restatement (fs_oracle.cpp).  It exists so that the
// HOST MODULE (foldseek_b200/csrc/cli_main.cpp: parameter handling, DB reading/encoding, coverage filters, checkCriteria, result sorting
// and serialisation) can be run and compared with the reference's own modules (oracle/_ref/refseek) in this container, which has no
// GPU: oracle/build.sh links cli_main.cpp against this file into oracle/_mock/foldseek-b200-mock, tests/test_host_module_cpu.py runs it.
// The numbers it returns are the oracle's (pinned against the reference); the CUDA library is compared with the
// same oracle in the -m gpu tests.  Only what cli_main.cpp calls is implemented; multi-GPU staging and lanes are served by plain copies.
#include "../include/fsb200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

extern "C" {
void *fso_create2(const char *mat3di_text, const char *blosum_text, float aaFactor, const uint8_t *nn_blob, uint64_t nn_len);
void fso_destroy(void *h);
void fso_align_query(void *h, uint64_t dbResidues, const uint8_t *qAA, const uint8_t *q3Di, int L, float biasScale, const uint8_t *tAA,
                     const uint8_t *t3Di, const uint64_t *tOff, const uint32_t *tLen, int n, int go, int ge, int32_t *out4, double *evals2);
int fso_backtrace(void *h, const uint8_t *qAA, const uint8_t *q3Di, int L, float biasScale, const uint8_t *tAA, const uint8_t *t3Di, int tLen,
                  int go, int ge, int qEnd, int dbEnd, int32_t targetScore, int32_t *out4, char *btOut, int btCap);
double fso_lddt(const float *qca, int qLen, const float *tca, int tLen, int qStart, int tStart, const char *bt, float *perRes, int *alnLenOut);
double fso_tmscore(const float *qca, int qLen, const float *tca, int tLen, int qStart, int tStart, const char *bt, int normLen, int *flags);
void *fso_index_create(void *h, const uint8_t *t3Di, const uint64_t *tOff, const uint32_t *tLen, uint32_t n, float sensitivity);
void fso_index_destroy(void *ix);
uint32_t fso_prefilter_query(void *h, void *ixp, uint32_t nTargets, const uint8_t *q3Di, int L, float biasScale, int maxSeqs, int minDiagScore,
                             int bins, uint32_t identityId, int32_t *out3, uint64_t *stats4);
}

struct fsb200_ctx {
    fsb200_config cfg;
    std::string err;
    void *o = nullptr, *ix = nullptr;
    bool ownsOracle = true;
    std::vector<uint8_t> tAA, t3Di, qAA, q3Di;
    std::vector<uint64_t> tOff, qOff;
    std::vector<uint32_t> tLen, qLen, identity;
    uint64_t aaDbResidues = 0;
    std::vector<float> tCa, qCa;                 // per entry x[L] y[L] z[L]
    std::vector<uint64_t> tCaOff, qCaOff;
    bool haveTCa = false, haveQCa = false;
    // last align
    std::vector<uint64_t> candOff; std::vector<uint32_t> candTarget, candQuery;
    std::vector<fsb200_aln> aln;
    std::vector<fsb200_bt> meta; std::string bt; bool btValid = false;
    std::vector<fsb200_sscore> ss; bool ssValid = false;
    uint64_t pfStats[8] = {0};
};

namespace {
int fail(fsb200_ctx *c, int rc, const char *msg) { if (c) c->err = msg; return rc; }

uint32_t reference_bin_size(size_t dbSize, int override_) {   // QueryMatcher.cpp:460-487, as in the library (prefilter.cuh)
    if (override_ > 0) return (uint32_t) override_;
    uint64_t l2 = 262144;
#if defined(_SC_LEVEL2_CACHE_SIZE)
    long v = sysconf(_SC_LEVEL2_CACHE_SIZE);
    if (v > 0) l2 = (uint64_t) v;
#endif
    for (uint32_t b = 2; b <= 1024; b <<= 1) if (dbSize / b < l2) return b;
    return 2048;
}

void stage_ca(const float *x, const float *y, const float *z, const uint64_t *off, const std::vector<uint32_t> &len, std::vector<float> &out, std::vector<uint64_t> &outOff) {
    outOff.resize(len.size());
    uint64_t total = 0;
    for (size_t i = 0; i < len.size(); i++) { outOff[i] = total; total += 3ull * len[i]; }
    out.resize(total);
    for (size_t i = 0; i < len.size(); i++) {
        memcpy(out.data() + outOff[i], x + off[i], len[i] * sizeof(float));
        memcpy(out.data() + outOff[i] + len[i], y + off[i], len[i] * sizeof(float));
        memcpy(out.data() + outOff[i] + 2ull * len[i], z + off[i], len[i] * sizeof(float));
    }
}
}  // namespace

extern "C" {

const char *fsb200_build_id(void) { return "mock"; }

void fsb200_default_config(fsb200_config *cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->gap_open = 10; cfg->gap_extend = 1; cfg->comp_bias_corr = 1;
    cfg->comp_bias_scale_align = 0.5f; cfg->comp_bias_scale_prefilter = 0.15f;
    cfg->sensitivity = 9.5f; cfg->kmer_size = 6; cfg->max_seqs = 1000; cfg->min_diag_score = 30;
    cfg->evalue_thr = 10.0; cfg->aa_factor = 1.4f; cfg->diag_bin_count = 0;
}

int fsb200_create(const fsb200_config *cfg, fsb200_ctx **out) {
    if (cfg == NULL || out == NULL) return FSB200_E_ARG;
    *out = NULL;
    fsb200_ctx *c = new fsb200_ctx();
    c->cfg = *cfg;
    if (!cfg->comp_bias_corr) { c->cfg.comp_bias_scale_align = -1.0f; c->cfg.comp_bias_scale_prefilter = -1.0f; }   // the oracle's convention for --comp-bias-corr 0
    c->o = fso_create2(cfg->mat3di_text, cfg->mataa_text, cfg->aa_factor, cfg->evalue_nn, cfg->evalue_nn_len);
    if (c->o == NULL) { delete c; return FSB200_E_MATRIX; }
    *out = c;
    return FSB200_OK;
}

void fsb200_destroy(fsb200_ctx *c) {
    if (c == NULL) return;
    if (c->ix) fso_index_destroy(c->ix);
    if (c->o && c->ownsOracle) fso_destroy(c->o);
    delete c;
}

const char *fsb200_last_error(const fsb200_ctx *c) { return c ? c->err.c_str() : "null context"; }

int fsb200_load_target(fsb200_ctx *c, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len, uint32_t n, uint64_t aaDbResidues) {
    if (c == NULL || aa == NULL || tdi == NULL || off == NULL || len == NULL) return FSB200_E_ARG;
    c->tOff.resize(n); c->tLen.assign(len, len + n);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) { c->tOff[i] = total; total += len[i]; }
    c->tAA.resize(total + 1); c->t3Di.resize(total + 1);
    for (uint32_t i = 0; i < n; i++) {
        for (uint32_t k = 0; k < len[i]; k++) if (aa[off[i] + k] > 20 || tdi[off[i] + k] > 20) return fail(c, FSB200_E_ARG, "residue code > 20");
        memcpy(c->tAA.data() + c->tOff[i], aa + off[i], len[i]);
        memcpy(c->t3Di.data() + c->tOff[i], tdi + off[i], len[i]);
    }
    c->aaDbResidues = aaDbResidues;
    if (c->ix) { fso_index_destroy(c->ix); c->ix = nullptr; }
    c->haveTCa = false;
    return FSB200_OK;
}

int fsb200_set_queries(fsb200_ctx *c, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len, uint32_t nq) {
    if (c == NULL || aa == NULL || tdi == NULL || off == NULL || len == NULL) return FSB200_E_ARG;
    c->qOff.resize(nq); c->qLen.assign(len, len + nq);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nq; i++) {
        if (len[i] == 0 || len[i] > 65535) return fail(c, FSB200_E_RANGE, "query length must be in 1..65535");
        c->qOff[i] = total; total += len[i];
    }
    c->qAA.resize(total + 1); c->q3Di.resize(total + 1);
    for (uint32_t i = 0; i < nq; i++) {
        for (uint32_t k = 0; k < len[i]; k++) {
            if (aa[off[i] + k] > 20 || tdi[off[i] + k] > 20) return fail(c, FSB200_E_ARG, "query residue code > 20");
            c->qAA[c->qOff[i] + k] = aa[off[i] + k]; c->q3Di[c->qOff[i] + k] = tdi[off[i] + k];
        }
    }
    c->identity.clear(); c->haveQCa = false; c->btValid = c->ssValid = false; c->aln.clear();
    return FSB200_OK;
}

int fsb200_set_identity(fsb200_ctx *c, const uint32_t *identity) {
    if (c == NULL) return FSB200_E_ARG;
    if (identity == NULL) { c->identity.clear(); return FSB200_OK; }
    c->identity.assign(identity, identity + c->qLen.size());
    return FSB200_OK;
}

int fsb200_build_index(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    if (c->tLen.empty()) return fail(c, FSB200_E_STATE, "load_target first");
    if (c->cfg.kmer_size != 6) return fail(c, FSB200_E_RANGE, "only -k 6 is implemented");
    if (!c->ix) c->ix = fso_index_create(c->o, c->t3Di.data(), c->tOff.data(), c->tLen.data(), (uint32_t) c->tLen.size(), c->cfg.sensitivity);
    return FSB200_OK;
}

int fsb200_prefilter(fsb200_ctx *c, fsb200_hit *hits, uint32_t *counts) {
    if (c == NULL || hits == NULL || counts == NULL) return FSB200_E_ARG;
    if (c->tLen.empty() || c->qLen.empty()) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    int rc = fsb200_build_index(c);
    if (rc != FSB200_OK) return rc;
    const uint32_t nT = (uint32_t) c->tLen.size();
    const int bins = (int) reference_bin_size(nT, c->cfg.diag_bin_count);
    const int ms = c->cfg.max_seqs;
    memset(c->pfStats, 0, sizeof(c->pfStats));
    std::vector<int32_t> out((size_t) std::max(ms, 1) * 3);
    for (size_t q = 0; q < c->qLen.size(); q++) {
        uint64_t st[4] = {0, 0, 0, 0};
        const uint32_t ident = c->identity.empty() ? 0xFFFFFFFFu : c->identity[q];
        const uint32_t n = fso_prefilter_query(c->o, c->ix, nT, c->q3Di.data() + c->qOff[q], (int) c->qLen[q], c->cfg.comp_bias_scale_prefilter, ms,
                                               c->cfg.min_diag_score, bins, ident, out.data(), st);
        counts[q] = n;
        for (uint32_t k = 0; k < n; k++) { fsb200_hit h; h.target = (uint32_t) out[3 * k]; h.score = out[3 * k + 1]; h.diagonal = out[3 * k + 2]; hits[q * (size_t) ms + k] = h; }
        c->pfStats[0] += st[0]; c->pfStats[1] += st[1]; c->pfStats[2] += st[3]; c->pfStats[5] += st[2];
    }
    return FSB200_OK;
}

int fsb200_prefilter_stats(const fsb200_ctx *c, uint64_t *out8) {
    if (c == NULL || out8 == NULL) return FSB200_E_ARG;
    memcpy(out8, c->pfStats, sizeof(c->pfStats));
    return FSB200_OK;
}

// alignStructure up to the traceback (structurealign.cpp:37-74), fields as the library fills them (fsb200.cu: fsb200_align_run)
int fsb200_align(fsb200_ctx *c, const uint64_t *candOff, const uint32_t *candTarget, fsb200_aln *out) {
    if (c == NULL || candOff == NULL) return FSB200_E_ARG;
    if (c->tLen.empty() || c->qLen.empty()) return fail(c, FSB200_E_STATE, "load_target and set_queries first");
    if (c->cfg.evalue_nn == NULL) return fail(c, FSB200_E_STATE, "E-value network not configured");
    const size_t nq = c->qLen.size();
    const uint64_t n = candOff[nq];
    c->candOff.assign(candOff, candOff + nq + 1); c->candTarget.assign(candTarget, candTarget + n); c->candQuery.resize(n);
    c->aln.resize(n);
    for (size_t q = 0; q < nq; q++) {
        const uint64_t b = candOff[q], e = candOff[q + 1];
        const int m = (int) (e - b);
        if (m == 0) continue;
        std::vector<uint64_t> off(m); std::vector<uint32_t> len(m);
        for (int k = 0; k < m; k++) {
            const uint32_t t = candTarget[b + k];
            if (t >= c->tLen.size()) return fail(c, FSB200_E_ARG, "candidate target id out of range");
            off[k] = c->tOff[t]; len[k] = c->tLen[t]; c->candQuery[b + k] = (uint32_t) q;
        }
        std::vector<int32_t> o4((size_t) m * 4); std::vector<double> ev((size_t) m * 2);
        fso_align_query(c->o, c->aaDbResidues, c->qAA.data() + c->qOff[q], c->q3Di.data() + c->qOff[q], (int) c->qLen[q], c->cfg.comp_bias_scale_align,
                        c->tAA.data(), c->t3Di.data(), off.data(), len.data(), m, c->cfg.gap_open, c->cfg.gap_extend, o4.data(), ev.data());
        for (int k = 0; k < m; k++) {
            fsb200_aln r;
            const int32_t fwd = o4[4 * k], qEnd = o4[4 * k + 1], dbEnd = o4[4 * k + 2], rev = o4[4 * k + 3];
            if (ev[2 * k] > c->cfg.evalue_thr) r = fsb200_aln{(uint32_t) fwd, 0, -1, dbEnd, 1, fwd, ev[2 * k]};
            else {
                r.fwd_score = (uint32_t) fwd; r.rev_score = (uint32_t) rev; r.q_end = qEnd; r.db_end = dbEnd;
                r.score = fwd - rev; r.evalue = ev[2 * k + 1]; r.status = (ev[2 * k + 1] <= c->cfg.evalue_thr) ? 0 : 2;
            }
            c->aln[b + k] = r;
        }
    }
    c->btValid = c->ssValid = false;
    if (out) memcpy(out, c->aln.data(), n * sizeof(fsb200_aln));
    return FSB200_OK;
}

int fsb200_backtrace_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    const fsb200_bt none{-1, -1, 0, 0, 0, -1, 0};
    c->meta.assign(c->aln.size(), none);
    c->bt.clear();
    for (size_t i = 0; i < c->aln.size(); i++) {
        const fsb200_aln &r = c->aln[i];
        if (r.status != 0) continue;
        const uint32_t q = c->candQuery[i], t = c->candTarget[i];
        int32_t o4[4] = {0, 0, 0, 0};
        std::vector<char> buf(c->qLen[q] + c->tLen[t] + 16, 0);
        const int ok = fso_backtrace(c->o, c->qAA.data() + c->qOff[q], c->q3Di.data() + c->qOff[q], (int) c->qLen[q], c->cfg.comp_bias_scale_align,
                                     c->tAA.data() + c->tOff[t], c->t3Di.data() + c->tOff[t], (int) c->tLen[t], c->cfg.gap_open, c->cfg.gap_extend,
                                     r.q_end, r.db_end, (int32_t) r.fwd_score, o4, buf.data(), (int) buf.size());
        fsb200_bt m = none;
        m.status = ok ? 1 : 0; m.block_score = o4[3];
        if (ok) {
            m.q_start = o4[0]; m.db_start = o4[1]; m.identical_aa = (uint32_t) o4[2]; m.aln_len = (uint32_t) strlen(buf.data());
            m.bt_off = c->bt.size();
            c->bt.append(buf.data(), m.aln_len);
        }
        c->meta[i] = m;
    }
    c->btValid = true; c->ssValid = false;
    return FSB200_OK;
}

int fsb200_backtrace_size(const fsb200_ctx *c, uint64_t *bytes) {
    if (c == NULL || bytes == NULL) return FSB200_E_ARG;
    *bytes = c->bt.size();
    return FSB200_OK;
}

int fsb200_backtrace_download(fsb200_ctx *c, fsb200_bt *meta, char *bt, uint64_t cap) {
    if (c == NULL || !c->btValid) return c ? fail(c, FSB200_E_STATE, "fsb200_backtrace_run first") : FSB200_E_ARG;
    if (cap < c->bt.size()) return fail(c, FSB200_E_RANGE, "backtrace buffer too small");
    if (!c->meta.empty()) memcpy(meta, c->meta.data(), c->meta.size() * sizeof(fsb200_bt));
    if (!c->bt.empty()) memcpy(bt, c->bt.data(), c->bt.size());
    return FSB200_OK;
}

int fsb200_load_target_ca(fsb200_ctx *c, const float *x, const float *y, const float *z, const uint64_t *off) {
    if (c == NULL || x == NULL || y == NULL || z == NULL || off == NULL) return FSB200_E_ARG;
    stage_ca(x, y, z, off, c->tLen, c->tCa, c->tCaOff);
    c->haveTCa = true;
    return FSB200_OK;
}

int fsb200_set_queries_ca(fsb200_ctx *c, const float *x, const float *y, const float *z, const uint64_t *off) {
    if (c == NULL || x == NULL || y == NULL || z == NULL || off == NULL) return FSB200_E_ARG;
    stage_ca(x, y, z, off, c->qLen, c->qCa, c->qCaOff);
    c->haveQCa = true;
    return FSB200_OK;
}

// structurealign.cpp:376-412 for every traced pair; untraced / empty backtraces stay at status -1 like in the library
int fsb200_structure_scores_run(fsb200_ctx *c) {
    if (c == NULL) return FSB200_E_ARG;
    if (!c->haveTCa || !c->haveQCa) return fail(c, FSB200_E_STATE, "load_target_ca and set_queries_ca first");
    if (!c->btValid) return fail(c, FSB200_E_STATE, "fsb200_backtrace_run first");
    const fsb200_sscore none{0.0, 0.0f, -1, 0};
    c->ss.assign(c->aln.size(), none);
    for (size_t i = 0; i < c->aln.size(); i++) {
        const fsb200_bt &m = c->meta[i];
        if (m.status != 1 || m.aln_len == 0) continue;
        const uint32_t q = c->candQuery[i], t = c->candTarget[i];
        const std::string bt = c->bt.substr(m.bt_off, m.aln_len);
        const float *qca = c->qCa.data() + c->qCaOff[q], *tca = c->tCa.data() + c->tCaOff[t];
        int alnLen = 0, flags = 0;
        fsb200_sscore s;
        s.lddt = (float) fso_lddt(qca, (int) c->qLen[q], tca, (int) c->tLen[t], m.q_start, m.db_start, bt.c_str(), NULL, &alnLen);
        const int mode = c->cfg.tmscore_norm_mode;
        const int normLen = mode == 1 ? (int) c->qLen[q] : mode == 2 ? (int) c->tLen[t] : mode == 3 ? (int) std::min(c->qLen[q], c->tLen[t])
                                                                                                       : std::min(c->aln[i].q_end - m.q_start, c->aln[i].db_end - m.db_start);
        s.tm = fso_tmscore(qca, (int) c->qLen[q], tca, (int) c->tLen[t], m.q_start, m.db_start, bt.c_str(), normLen, &flags);
        s.status = (flags & 1) ? 0 : 1;
        s.score_bits = (int32_t) (c->aln[i].score * sqrt((double) s.lddt * s.tm));
        c->ss[i] = s;
    }
    c->ssValid = true;
    return FSB200_OK;
}

int fsb200_structure_scores_download(fsb200_ctx *c, fsb200_sscore *out) {
    if (c == NULL || !c->ssValid) return c ? fail(c, FSB200_E_STATE, "fsb200_structure_scores_run first") : FSB200_E_ARG;
    if (!c->ss.empty()) memcpy(out, c->ss.data(), c->ss.size() * sizeof(fsb200_sscore));
    return FSB200_OK;
}

// ---- multi-context plumbing of the host module, served by plain host copies ----
int fsb200_staged_target_size(const fsb200_ctx *c, uint64_t *bytes) { if (c == NULL || bytes == NULL) return FSB200_E_ARG; *bytes = c->tAA.size(); return FSB200_OK; }
int fsb200_export_staged_target(const fsb200_ctx *c, void **dAA, void **dTdi, uint64_t *offOut, uint32_t *lenOut) {
    if (c == NULL || dAA == NULL || dTdi == NULL) return FSB200_E_ARG;
    *dAA = (void *) c->tAA.data(); *dTdi = (void *) c->t3Di.data();
    if (offOut) memcpy(offOut, c->tOff.data(), c->tOff.size() * sizeof(uint64_t));
    if (lenOut) memcpy(lenOut, c->tLen.data(), c->tLen.size() * sizeof(uint32_t));
    return FSB200_OK;
}
int fsb200_load_target_device(fsb200_ctx *c, const void *dAA, const void *dTdi, uint64_t bytes, const uint64_t *off, const uint32_t *len, uint32_t n, uint64_t aaDbResidues) {
    (void) bytes;
    return fsb200_load_target(c, (const uint8_t *) dAA, (const uint8_t *) dTdi, off, len, n, aaDbResidues);
}
int fsb200_set_host_threads(fsb200_ctx *c, int32_t n) { return (c == NULL || n < 0) ? FSB200_E_ARG : FSB200_OK; }

int fsb200_create_lane(fsb200_ctx *parent, fsb200_ctx **out) {
    if (parent == NULL || out == NULL) return FSB200_E_ARG;
    fsb200_ctx *c = new fsb200_ctx();
    c->cfg = parent->cfg;
    // an oracle of its own: the oracle object keeps per-call state (log of the DB size), and lanes are driven by different host threads
    c->o = fso_create2(c->cfg.mat3di_text, c->cfg.mataa_text, c->cfg.aa_factor, c->cfg.evalue_nn, c->cfg.evalue_nn_len); c->ownsOracle = true;
    if (c->o == NULL) { delete c; return FSB200_E_MATRIX; }
    c->tAA = parent->tAA; c->t3Di = parent->t3Di; c->tOff = parent->tOff; c->tLen = parent->tLen; c->aaDbResidues = parent->aaDbResidues;
    c->tCa = parent->tCa; c->tCaOff = parent->tCaOff; c->haveTCa = parent->haveTCa;
    *out = c;
    return FSB200_OK;
}

}  // extern "C"
