// foldseek-b200 — host side of the two hot-path modules above the C ABI (include/fsb200.h), in C++ like the reference.
//
//   foldseek-b200 prefilter      <queryDB_ss> <targetDB_ss> <prefilterDB> [options]   (lib/mmseqs/src/prefiltering/Main.cpp:13)
//   foldseek-b200 structurealign <queryDB> <targetDB> <prefilterDB> <alnDB> [options] (src/strucclustutils/structurealign.cpp:141)
//
// Same positional arguments and the same MMseqs2 on-disk formats as the reference modules:
//   data file + ".index" (key \t offset \t length, length includes the trailing "\n\0" of sequences / "\0" of results,
//   lib/mmseqs/src/commons/DBWriter.cpp:418-428) + ".dbtype" (int32, written last so that the workflow's
//   `notExists "<out>.dbtype"` resume guard stays valid, data/structuresearch.sh:41,119).
// prefilter writes byte-identical entries ("targetKey\tscore\tdiagonal\n", QueryMatcher.h:120-132, dbtype 7).
// structurealign writes the 10 (+1 with -a) columns of Matcher::resultToBuffer (lib/mmseqs/src/alignment/Matcher.cpp:282-329,
// dbtype 5) with start positions, seqId and backtrace from the GPU traceback and, for the default --sort-by-structure-bits 1,
// the score rescaled by sqrt(LDDT * TM-score) of the alignment (structurealign.cpp:376-412) from the GPU structure-score
// kernels; without C-alpha databases the rescoring is disabled with the reference's warning (structurealign.cpp:176-196).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <libgen.h>
#include <unistd.h>
#include "fsb200.h"

namespace {

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "Error: %s\n", msg.c_str());   // Debug(Debug::ERROR) << ...; EXIT(EXIT_FAILURE)
    exit(EXIT_FAILURE);
}

std::string slurp(const std::string &path, bool required = true) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { if (required) die("cannot read " + path); return std::string(); }
    std::ostringstream ss; ss << in.rdbuf();
    return ss.str();
}

struct Db {   // read-only MMseqs2 DB (plain, uncompressed; per-thread shards <name>.0, .1, ... are concatenated)
    std::string data;
    std::vector<uint32_t> key; std::vector<uint64_t> off, len;
    int dbtype = 0;
    void open(const std::string &name) {
        std::ifstream f(name, std::ios::binary);
        if (f) { std::ostringstream ss; ss << f.rdbuf(); data = ss.str(); }
        else {
            for (int i = 0;; i++) {
                std::ifstream s(name + "." + std::to_string(i), std::ios::binary);
                if (!s) { if (i == 0) die("cannot open database " + name); break; }
                std::ostringstream ss; ss << s.rdbuf(); data += ss.str();
            }
        }
        std::ifstream idx(name + ".index");
        if (!idx) die("cannot open " + name + ".index");
        unsigned long long k, o, l;
        while (idx >> k >> o >> l) { key.push_back((uint32_t) k); off.push_back(o); len.push_back(l); }
        std::string t = slurp(name + ".dbtype");
        if (t.size() >= 4) memcpy(&dbtype, t.data(), 4);
        if (dbtype & (1 << 31)) die(name + ": compressed databases are not supported");
    }
    size_t size() const { return key.size(); }
};

void write_db(const std::string &name, const std::vector<uint32_t> &keys, const std::vector<std::string> &entries, int dbtype) {
    std::ofstream d(name, std::ios::binary), i(name + ".index");
    if (!d || !i) die("cannot write " + name);
    uint64_t off = 0;
    for (size_t k = 0; k < keys.size(); k++) {
        d.write(entries[k].data(), entries[k].size()); d.put('\0');
        i << keys[k] << '\t' << off << '\t' << (entries[k].size() + 1) << '\n';
        off += entries[k].size() + 1;
    }
    d.close(); i.close();
    std::ofstream t(name + ".dbtype", std::ios::binary);   // last: resume guard
    t.write(reinterpret_cast<const char *>(&dbtype), 4);
}

// letters -> codes (BaseMatrix alphabet A C D E F G H I K L M N P Q R S T V W Y X; unknown -> X; lower == upper)
void encode(const Db &db, bool maskLower, std::vector<uint8_t> &codes, std::vector<uint64_t> &off, std::vector<uint32_t> &len) {
    static const char *alph = "ACDEFGHIKLMNPQRSTVWYX";
    uint8_t map[256];
    for (int c = 0; c < 256; c++) map[c] = 20;
    for (int a = 0; a < 21; a++) { map[(unsigned char) alph[a]] = (uint8_t) a; map[(unsigned char) tolower(alph[a])] = (uint8_t) a; }
    off.resize(db.size()); len.resize(db.size());
    uint64_t total = 0;
    for (size_t i = 0; i < db.size(); i++) { len[i] = (uint32_t) (db.len[i] >= 2 ? db.len[i] - 2 : 0); off[i] = total; total += len[i]; }   // length includes "\n\0"
    codes.resize(total + 1);
    for (size_t i = 0; i < db.size(); i++) {
        const char *s = db.data.data() + db.off[i];
        for (uint32_t k = 0; k < len[i]; k++) {
            const unsigned char c = (unsigned char) s[k];
            // --mask-lower-case 1 (Foldseek default, LocalParameters.cpp:382-385): lower-case target residues are masked
            codes[off[i] + k] = (maskLower && islower(c)) ? 20 : map[c];
        }
    }
}

struct Args {
    std::vector<std::string> pos;
    fsb200_config cfg;
    int verbosity = 3;
    int addBacktrace = 0;           // -a
    int sortByStructureBits = 1;    // LocalParameters.cpp: default 1
};

Args parse(int argc, char **argv, int first) {
    Args a;
    fsb200_default_config(&a.cfg);
    for (int i = first; i < argc; i++) {
        std::string s = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) die("missing value for " + s); return argv[++i]; };
        if (s == "-s") a.cfg.sensitivity = (float) atof(val());
        else if (s == "-k") a.cfg.kmer_size = atoi(val());
        else if (s == "--max-seqs") a.cfg.max_seqs = atoi(val());
        else if (s == "-e") a.cfg.evalue_thr = atof(val());
        else if (s == "--comp-bias-corr") a.cfg.comp_bias_corr = atoi(val());
        else if (s == "--comp-bias-corr-scale") { float v = (float) atof(val()); a.cfg.comp_bias_scale_prefilter = v; a.cfg.comp_bias_scale_align = v; }
        else if (s == "--min-ungapped-score") a.cfg.min_diag_score = atoi(val());
        else if (s == "--gap-open") a.cfg.gap_open = atoi(val());
        else if (s == "--gap-extend") a.cfg.gap_extend = atoi(val());
        else if (s == "--alignment-type") a.cfg.aa_factor = (atoi(val()) == 2) ? 1.4f : 0.0f;
        else if (s == "--gpu") a.cfg.device = std::max(0, atoi(val()) - 1);
        else if (s == "-v") a.verbosity = atoi(val());
        else if (s == "-a") a.addBacktrace = atoi(val());
        else if (s == "--sort-by-structure-bits") a.sortByStructureBits = atoi(val());
        else if (s == "--threads" || s == "--alignment-mode" || s == "--db-load-mode" ||
                 s == "--split" || s == "--split-mode" || s == "--split-memory-limit" || s == "--compressed" || s == "--mask" ||
                 s == "--mask-lower-case" || s == "--spaced-kmer-mode" || s == "--exact-kmer-matching" || s == "--cov-mode" || s == "-c") (void) val();
        else if (!s.empty() && s[0] == '-') die("unrecognized parameter " + s);
        else a.pos.push_back(s);
    }
    return a;
}

std::string data_dir(const char *argv0) {
    if (const char *e = getenv("FSB200_DATA")) return e;
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    std::string exe = (n > 0) ? std::string(buf, n) : std::string(argv0);
    std::vector<char> tmp(exe.begin(), exe.end()); tmp.push_back('\0');
    return std::string(dirname(tmp.data())) + "/../data";
}

fsb200_ctx *make_ctx(Args &a, const char *argv0, std::string &m3, std::string &mA, std::string &nn) {
    const std::string d = data_dir(argv0);
    m3 = slurp(d + "/mat3di.out"); mA = slurp(d + "/blosum62.out"); nn = slurp(d + "/evalue_nn.kerasify");
    a.cfg.mat3di_text = m3.c_str(); a.cfg.mataa_text = mA.c_str();
    a.cfg.evalue_nn = reinterpret_cast<const uint8_t *>(nn.data()); a.cfg.evalue_nn_len = nn.size();
    fsb200_ctx *ctx = NULL;
    const int rc = fsb200_create(&a.cfg, &ctx);
    if (rc != FSB200_OK) die("fsb200_create failed with " + std::to_string(rc) + " (a CUDA device is required, there is no CPU path)");
    return ctx;
}

void check(fsb200_ctx *ctx, int rc, const char *what) {
    if (rc != FSB200_OK) die(std::string(what) + " failed (" + std::to_string(rc) + "): " + fsb200_last_error(ctx));
}

int prefilter_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 3) die("usage: foldseek-b200 prefilter <i:queryDB> <i:targetDB> <o:prefilterDB> [options]");
    Db q, t; q.open(a.pos[0]); t.open(a.pos[1]);
    std::string m3, mA, nn;
    fsb200_ctx *ctx = make_ctx(a, argv[0], m3, mA, nn);
    std::vector<uint8_t> tc, qc; std::vector<uint64_t> to, qo; std::vector<uint32_t> tl, ql;
    encode(t, true, tc, to, tl); encode(q, false, qc, qo, ql);
    uint64_t res = 0; for (uint32_t l : tl) res += l;
    check(ctx, fsb200_load_target(ctx, tc.data(), tc.data(), to.data(), tl.data(), (uint32_t) t.size(), res), "load_target");
    // queries shorter than one window produce empty lists; zero-length entries are given a single X
    for (size_t i = 0; i < ql.size(); i++) if (ql[i] == 0) die("empty query sequence");
    check(ctx, fsb200_set_queries(ctx, qc.data(), qc.data(), qo.data(), ql.data(), (uint32_t) q.size()), "set_queries");
    std::vector<fsb200_hit> hits((size_t) q.size() * a.cfg.max_seqs); std::vector<uint32_t> counts(q.size());
    check(ctx, fsb200_prefilter(ctx, hits.data(), counts.data()), "prefilter");
    std::vector<std::string> out(q.size());
    for (size_t i = 0; i < q.size(); i++) {
        std::string &s = out[i];
        char buf[64];
        for (uint32_t k = 0; k < counts[i]; k++) {
            const fsb200_hit &h = hits[i * (size_t) a.cfg.max_seqs + k];
            snprintf(buf, sizeof(buf), "%u\t%d\t%d\n", t.key[h.target], h.score, h.diagonal);
            s += buf;
        }
    }
    write_db(a.pos[2], q.key, out, 7 /* DBTYPE_PREFILTER_RES */);
    if (a.verbosity >= 3) {
        uint64_t st[8]; fsb200_prefilter_stats(ctx, st);
        printf("%.1f k-mers per position\n%llu DB matches per sequence\n", (double) st[0] / std::max<uint64_t>(1, q.size()) / 341.0,
               (unsigned long long) (st[1] / std::max<uint64_t>(1, q.size())));
    }
    fsb200_destroy(ctx);
    return EXIT_SUCCESS;
}

// Coordinate16::read (src/commons/Coordinate16.h:15-57): raw float32 x[L] y[L] z[L], or int32 start + int16 differences per axis
void decode_ca(const Db &db, size_t entry, uint32_t L, std::vector<float> &x, std::vector<float> &y, std::vector<float> &z) {
    const char *mem = db.data.data() + db.off[entry];
    const size_t entryLen = db.len[entry];
    const size_t base = x.size();
    x.resize(base + L); y.resize(base + L); z.resize(base + L);
    float *out[3] = {x.data() + base, y.data() + base, z.data() + base};
    if (entryLen >= (size_t) L * 3 * sizeof(float)) {
        for (int d = 0; d < 3; d++) memcpy(out[d], mem + (size_t) d * L * sizeof(float), (size_t) L * sizeof(float));
        return;
    }
    const char *p = mem;
    for (int d = 0; d < 3; d++) {
        int32_t start, diffSum = 0;
        memcpy(&start, p, 4); p += 4;
        if (L > 0) out[d][0] = start / 1000.0f;
        for (uint32_t i = 1; i < L; i++) {
            int16_t diff;
            memcpy(&diff, p, 2); p += 2;
            diffSum += diff;
            out[d][i] = (start + diffSum) / 1000.0f;
        }
    }
}

int structurealign_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: foldseek-b200 structurealign <i:queryDB> <i:targetDB> <i:prefilterDB> <o:alnDB> [options]");
    Db qa, qs, ta, ts, pref;
    qa.open(a.pos[0]); qs.open(a.pos[0] + "_ss"); ta.open(a.pos[1]); ts.open(a.pos[1] + "_ss"); pref.open(a.pos[2]);
    Db qCa, tCa;
    if (a.sortByStructureBits) {   // structurealign.cpp:176-196
        const bool ca1 = !slurp(a.pos[0] + "_ca.dbtype", false).empty(), ca2 = !slurp(a.pos[1] + "_ca.dbtype", false).empty();
        if (ca1 && ca2) { qCa.open(a.pos[0] + "_ca"); tCa.open(a.pos[1] + "_ca"); }
        else {
            fprintf(stderr, "Cannot find query or target C-alpha database\nDisabling --sort-by-structure-bits\n");
            a.sortByStructureBits = 0;
        }
    }
    std::string m3, mA, nn;
    fsb200_ctx *ctx = make_ctx(a, argv[0], m3, mA, nn);
    std::vector<uint8_t> tca, tcs, qca, qcs; std::vector<uint64_t> to, qo, o2; std::vector<uint32_t> tl, ql, l2;
    encode(ta, false, tca, to, tl); encode(ts, false, tcs, o2, l2);
    if (l2 != tl) die("target AA and 3Di databases differ in sequence lengths");
    encode(qa, false, qca, qo, ql); encode(qs, false, qcs, o2, l2);
    if (l2 != ql) die("query AA and 3Di databases differ in sequence lengths");
    // DBReader::getAminoAcidDBSize() = data size - 2 * entries (lib/mmseqs/src/commons/DBReader.cpp:403-412)
    const uint64_t aaRes = ta.data.size() - 2 * ta.size();
    check(ctx, fsb200_load_target(ctx, tca.data(), tcs.data(), to.data(), tl.data(), (uint32_t) ta.size(), aaRes), "load_target");
    if (a.sortByStructureBits) {
        std::vector<uint32_t> keyToCa(tCa.size() ? *std::max_element(tCa.key.begin(), tCa.key.end()) + 1 : 0, 0xFFFFFFFFu);
        for (size_t i = 0; i < tCa.size(); i++) keyToCa[tCa.key[i]] = (uint32_t) i;
        std::vector<float> x, y, z; std::vector<uint64_t> co(ta.size());
        for (size_t i = 0; i < ta.size(); i++) {
            if (ta.key[i] >= keyToCa.size() || keyToCa[ta.key[i]] == 0xFFFFFFFFu) die("no C-alpha entry for a target key");
            co[i] = x.size();
            decode_ca(tCa, keyToCa[ta.key[i]], tl[i], x, y, z);
        }
        check(ctx, fsb200_load_target_ca(ctx, x.data(), y.data(), z.data(), co.data()), "load_target_ca");
    }
    // one query batch = the queries that have a prefilter entry, in prefilter-DB order
    std::vector<uint32_t> keyToT(ta.size() ? *std::max_element(ta.key.begin(), ta.key.end()) + 1 : 0, 0xFFFFFFFFu);
    for (size_t i = 0; i < ta.size(); i++) keyToT[ta.key[i]] = (uint32_t) i;
    std::vector<uint32_t> keyToQ(qa.size() ? *std::max_element(qa.key.begin(), qa.key.end()) + 1 : 0, 0xFFFFFFFFu);
    for (size_t i = 0; i < qa.size(); i++) keyToQ[qa.key[i]] = (uint32_t) i;
    std::vector<uint8_t> bqa, bqs; std::vector<uint64_t> bo; std::vector<uint32_t> bl;
    std::vector<uint64_t> candOff(1, 0); std::vector<uint32_t> cand;
    for (size_t p = 0; p < pref.size(); p++) {
        if (pref.key[p] >= keyToQ.size() || keyToQ[pref.key[p]] == 0xFFFFFFFFu) die("prefilter entry for unknown query key");
        const uint32_t qi = keyToQ[pref.key[p]];
        bo.push_back(bqa.size()); bl.push_back(ql[qi]);
        bqa.insert(bqa.end(), qca.begin() + qo[qi], qca.begin() + qo[qi] + ql[qi]);
        bqs.insert(bqs.end(), qcs.begin() + qo[qi], qcs.begin() + qo[qi] + ql[qi]);
        const char *s = pref.data.data() + pref.off[p];
        while (*s != '\0') {
            const unsigned long k = strtoul(s, NULL, 10);
            if (k >= keyToT.size() || keyToT[k] == 0xFFFFFFFFu) die("prefilter hit for unknown target key");
            cand.push_back(keyToT[k]);
            while (*s != '\n' && *s != '\0') s++;
            if (*s == '\n') s++;
        }
        candOff.push_back(cand.size());
    }
    std::vector<std::string> out(pref.size());
    if (!bl.empty()) {
        check(ctx, fsb200_set_queries(ctx, bqa.data(), bqs.data(), bo.data(), bl.data(), (uint32_t) bl.size()), "set_queries");
        if (a.sortByStructureBits) {
            std::vector<uint32_t> keyToCa(qCa.size() ? *std::max_element(qCa.key.begin(), qCa.key.end()) + 1 : 0, 0xFFFFFFFFu);
            for (size_t i = 0; i < qCa.size(); i++) keyToCa[qCa.key[i]] = (uint32_t) i;
            std::vector<float> x, y, z; std::vector<uint64_t> co(pref.size());
            for (size_t p = 0; p < pref.size(); p++) {
                if (pref.key[p] >= keyToCa.size() || keyToCa[pref.key[p]] == 0xFFFFFFFFu) die("no C-alpha entry for a query key");
                co[p] = x.size();
                decode_ca(qCa, keyToCa[pref.key[p]], bl[p], x, y, z);
            }
            check(ctx, fsb200_set_queries_ca(ctx, x.data(), y.data(), z.data(), co.data()), "set_queries_ca");
        }
        std::vector<fsb200_aln> res(cand.size());
        check(ctx, fsb200_align(ctx, candOff.data(), cand.data(), res.data()), "align");
        check(ctx, fsb200_backtrace_run(ctx), "backtrace");
        uint64_t btBytes = 0;
        check(ctx, fsb200_backtrace_size(ctx, &btBytes), "backtrace_size");
        std::vector<fsb200_bt> meta(cand.size());
        std::vector<char> bt(btBytes + 1);
        check(ctx, fsb200_backtrace_download(ctx, meta.data(), bt.data(), bt.size()), "backtrace_download");
        std::vector<fsb200_sscore> sscore;
        if (a.sortByStructureBits) {
            sscore.resize(cand.size());
            check(ctx, fsb200_structure_scores_run(ctx), "structure_scores");
            check(ctx, fsb200_structure_scores_download(ctx, sscore.data()), "structure_scores_download");
        }
        struct Row { uint32_t dbKey; int score; float seqId; double eval; int qStart, qEnd, dbStart, dbEnd; uint32_t dbLen; std::string cigar; };
        for (size_t p = 0; p < pref.size(); p++) {
            std::vector<Row> rows;
            for (uint64_t k = candOff[p]; k < candOff[p + 1]; k++) {
                if (res[k].status != 0) continue;
                Row r;
                r.dbKey = ta.key[cand[k]]; r.score = res[k].score; r.eval = res[k].evalue;
                r.qEnd = res[k].q_end; r.dbEnd = res[k].db_end; r.dbLen = tl[cand[k]];
                r.qStart = -1; r.dbStart = -1; r.seqId = 0.0f;   // alignScoreEndPos leaves -1 (StructureSmithWaterman.cpp:274-275)
                int alnLen;
                if (meta[k].status == 1 && meta[k].aln_len > 0) {
                    r.qStart = meta[k].q_start; r.dbStart = meta[k].db_start;
                    alnLen = (int) meta[k].aln_len;
                    r.seqId = (float) meta[k].identical_aa / (float) alnLen;   // Util::computeSeqId, SEQ_ID_ALN_LEN
                    // Matcher::compressAlignment (Matcher.cpp:168-187)
                    const char *b = bt.data() + meta[k].bt_off;
                    char state = 'M'; size_t cnt = 0;
                    for (uint32_t x = 0; x < meta[k].aln_len; x++) {
                        if (b[x] != state) { r.cigar += std::to_string(cnt); r.cigar.push_back(state); state = b[x]; cnt = 1; }
                        else cnt++;
                    }
                    r.cigar += std::to_string(cnt); r.cigar.push_back(state);
                } else {
                    if (meta[k].status == 1) { r.qStart = meta[k].q_start; r.dbStart = meta[k].db_start; }
                    alnLen = std::max(abs(r.qEnd - r.qStart), abs(r.dbEnd - r.dbStart)) + 1;   // Matcher::computeAlnLength
                    r.cigar = "0M";                                                           // compressAlignment("")
                }
                // Alignment::checkCriteria with the module defaults (--min-seq-id 0, -c 0, --min-aln-len 0)
                if (!(r.eval <= a.cfg.evalue_thr && r.seqId >= 0.0f && alnLen >= 0)) continue;
                if (a.sortByStructureBits) {   // structurealign.cpp:376-412 (thresholds --tmscore-threshold / --lddt-threshold default 0)
                    if (sscore[k].status != 1) die("structure scores: degenerate superposition or missing backtrace for target key " + std::to_string(r.dbKey));
                    r.score = sscore[k].score_bits;
                }
                rows.push_back(r);
            }
            // compareHitsByStructureBits (structurealign.cpp:21-29) or Matcher::compareHits (Matcher.h:161-172)
            const bool bits = a.sortByStructureBits != 0;
            std::sort(rows.begin(), rows.end(), [bits](const Row &x, const Row &y) {
                if (!bits && x.eval != y.eval) return x.eval < y.eval;
                if (x.score != y.score) return x.score > y.score;
                if (x.dbLen != y.dbLen) return x.dbLen < y.dbLen;
                return x.dbKey < y.dbKey;
            });
            char buf[256];
            for (const Row &r : rows) {
                // Util::fastSeqIdToBuffer (Util.cpp:251-280)
                char sid[16];
                if (r.seqId == 1.0f) snprintf(sid, sizeof(sid), "1.000");
                else snprintf(sid, sizeof(sid), "0.%s%s%d", r.seqId < 0.10 ? "0" : "", r.seqId < 0.01 ? "0" : "", (int) (r.seqId * 1000));
                snprintf(buf, sizeof(buf), "%u\t%d\t%s\t%.3E\t%d\t%d\t%u\t%d\t%d\t%u", r.dbKey, r.score, sid, r.eval,
                         r.qStart, r.qEnd, bl[p], r.dbStart, r.dbEnd, r.dbLen);
                out[p] += buf;
                if (a.addBacktrace) { out[p] += '\t'; out[p] += r.cigar; }
                out[p] += '\n';
            }
        }
    }
    write_db(a.pos[3], pref.key, out, 5 /* DBTYPE_ALIGNMENT_RES */);
    fsb200_destroy(ctx);
    return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) die("usage: foldseek-b200 <prefilter|structurealign> ...");
    if (strcmp(argv[1], "prefilter") == 0) return prefilter_main(argc, argv);
    if (strcmp(argv[1], "structurealign") == 0) return structurealign_main(argc, argv);
    die(std::string("unknown module ") + argv[1]);
}
