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
//
//   foldseek-b200 gpuserver <targetDB> [options]                                      (lib/mmseqs/src/util/gpuserver.cpp:24)
//
// keeps the target DB of both modules resident on the GPU(s) - staged sequences, C-alpha coordinates and the k-mer index - and
// serves `prefilter` / `structurealign` calls that carry `--gpu-server 1` (lib/mmseqs/src/prefiltering/ungappedprefilter.cpp:72-141):
// such a call hands its argument list to the server of its target DB and the server runs the module on its resident contexts, so
// the per-invocation staging (3.4 s for 1M structures) and index build (0.5 s) are paid once.  Where the reference exchanges one
// query at a time through a shared-memory block named by a hash of the DB path (GPUSharedMemory, GpuUtil.h), a whole module call
// travels here over a Unix-domain socket named the same way; results are the DBs the module writes itself.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <csignal>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include <libgen.h>
#include <poll.h>
#include <sched.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>
#include "fsb200.h"

namespace {

#include "cli_db.h"
#include "cli_args.h"
#include "cli_contexts.h"

// Util::canBeCovered (lib/mmseqs/src/commons/Util.cpp:542-559); --cov-mode: 0 bidirectional, 1 TARGET, 2 QUERY, 3 length of the query,
// 4 length of the target, 5 shorter sequence (Parameters.h:284-289)
bool can_be_covered(float covThr, int covMode, float q, float t) {
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;   // COV_MODE_TARGET
        case 2: return (t / q) >= covThr;   // COV_MODE_QUERY
        case 3: return (t / q) >= covThr && (t / q) <= 1.0;
        case 4: return (q / t) >= covThr && (q / t) <= 1.0;
        case 5: return (std::min(t, q) / std::max(t, q)) >= covThr;
        default: return true;
    }
}
// Util::hasCoverage (Util.cpp:561-576)
bool has_coverage(float covThr, int covMode, float qCov, float tCov) {
    switch (covMode) {
        case 0: return qCov >= covThr && tCov >= covThr;
        case 1: return tCov >= covThr;      // COV_MODE_TARGET
        case 2: return qCov >= covThr;      // COV_MODE_QUERY
        default: return true;
    }
}
// StructureSmithWaterman::computeCov (src/commons/StructureSmithWaterman.cpp:2028-2030), unsigned arithmetic included
float compute_cov(unsigned int startPos, unsigned int endPos, unsigned int len) {
    return (std::min(len, std::max(startPos, endPos)) - std::min(startPos, endPos) + 1) / (float) len;
}


// The prefilter module behind its arguments: the hit list of every query of qPath (a 3Di DB) against tPath (a 3Di DB), after the
// canBeCovered filter and the target-split merge.  onContexts != NULL: contexts that already hold this target, staged and indexed (search).
struct PH { uint32_t key; uint32_t tIdx; int score; int diag; };   // target key, target DBReader id, ungapped score, diagonal
std::vector<std::vector<PH>> prefilter_core(Args &a, const char *argv0, const std::string &qPath, const std::string &tPath, Db &q, PrefTarget *onContexts, bool &compressed) {
    const bool sameDB = qPath == tPath;   // Prefiltering::isSameQTDB (Prefiltering.cpp:585-598)
    q.open(qPath);
    std::vector<uint8_t> qc; std::vector<uint64_t> qo; std::vector<uint32_t> ql;
    encode(q, false, qc, qo, ql);
    for (size_t i = 0; i < ql.size(); i++) if (ql[i] == 0) die("empty query sequence");
    // --split N --split-mode 0 (TARGET_DB_SPLIT): the target DB is searched in N entry ranges of about equal size, each with its own index
    // and a shorter result list (max-seqs / N + 4 standard deviations, Prefiltering.cpp:388-394), the lists of a query are concatenated
    // in split order and sorted by score and target key (mergeTargetSplits, Prefiltering.cpp:412-526).  Query split (mode 1) and the
    // automatic mode on a DB that fits never change a result: the queries are independent.
    const bool targetSplit = a.split > 1 && a.splitMode == 0;
    if (a.split > 1 && a.splitMode == 2) die("parameter --split " + std::to_string(a.split) + " --split-mode 2 is not implemented by foldseek-b200 (supported: --split-mode 0 target split, 1 query split)");
    std::vector<std::vector<PH>> lists(q.size());
    const bool covFilter = a.covThr > 0.0f && (a.covMode == 0 || a.covMode == 2 || a.covMode == 5);   // bidirectional, QUERY, shorter: Prefiltering.cpp:886-894
    uint64_t windows = 0; for (uint32_t l : ql) windows += l;
    // one pass of all queries over one staged + indexed target range; dbFrom = DBReader id of its first target
    auto pass = [&](PrefTarget &T, const std::vector<uint32_t> &ident, int maxSeqs, size_t dbFrom, std::vector<std::vector<PH>> &dst) {
        std::vector<fsb200_ctx *> &ctxs = T.ctxs;
        const size_t ms = (size_t) maxSeqs;
        std::vector<fsb200_hit> hits(q.size() * ms); std::vector<uint32_t> counts(q.size());
        std::vector<uint64_t> w(ql.begin(), ql.end());
        const std::vector<size_t> bounds = split_by_weight(w, ctxs.size());
        std::vector<char> ran(ctxs.size(), 0);
        if (g_resident) g_resident->touched = true;
        run_workers(ctxs.size(), [&](size_t g) {
            const size_t b = bounds[g], e = bounds[g + 1];
            if (e <= b) return;
            std::vector<uint64_t> off(qo.begin() + b, qo.begin() + e);
            check(ctxs[g], fsb200_set_queries(ctxs[g], qc.data(), qc.data(), off.data(), ql.data() + b, (uint32_t) (e - b)), "set_queries");
            if (!ident.empty()) check(ctxs[g], fsb200_set_identity(ctxs[g], ident.data() + b), "set_identity");
            check(ctxs[g], fsb200_prefilter(ctxs[g], hits.data() + b * ms, counts.data() + b), "prefilter");
            ran[g] = 1;
        });
        for (size_t i = 0; i < q.size(); i++)
            for (uint32_t k = 0; k < counts[i]; k++) {
                const fsb200_hit &h = hits[i * ms + k];
                if (covFilter && !can_be_covered(a.covThr, a.covMode, (float) ql[i], (float) T.tl[h.target])) continue;
                dst[i].push_back(PH{T.tkey[h.target], (uint32_t) (dbFrom + h.target), h.score, h.diagonal});
            }
        if (a.verbosity >= 3) {
            uint64_t st[8], sum[8] = {0};
            for (size_t g = 0; g < ctxs.size(); g++) if (ran[g]) { fsb200_prefilter_stats(ctxs[g], st); for (int k = 0; k < 8; k++) sum[k] += st[k]; }
            char line[160];
            snprintf(line, sizeof(line), "%.1f k-mers per position\n%llu DB matches per sequence\n", (double) sum[0] / std::max<uint64_t>(1, windows),
                     (unsigned long long) (sum[1] / std::max<uint64_t>(1, q.size())));
            say_out(line);
        }
    };
    // identity: the query's own entry in the target DB (Prefiltering.cpp:856-868); with target splits only the split that holds it
    auto identity = [&](const std::vector<uint32_t> &tkey) {
        std::vector<uint32_t> ident;
        if (!(sameDB || a.includeIdentity)) return ident;
        std::vector<uint32_t> keyToT(tkey.size() ? *std::max_element(tkey.begin(), tkey.end()) + 1 : 0, 0xFFFFFFFFu);
        for (size_t i = 0; i < tkey.size(); i++) keyToT[tkey[i]] = (uint32_t) i;
        ident.resize(q.size());
        for (size_t i = 0; i < q.size(); i++) ident[i] = (q.key[i] < keyToT.size()) ? keyToT[q.key[i]] : 0xFFFFFFFFu;
        return ident;
    };
    compressed = a.compressed != 0;
    if (!targetSplit && onContexts != NULL) {
        set_threads(a, onContexts->ctxs);
        pass(*onContexts, identity(onContexts->tkey), a.cfg.max_seqs, 0, lists);
    } else if (!targetSplit) {
        std::shared_ptr<PrefTarget> T = pref_target(a, argv0, tPath, g_resident ? (size_t) -1 : ql.size());
        pass(*T, identity(T->tkey), a.cfg.max_seqs, 0, lists);
        T->served++;
        if (g_resident && a.verbosity >= 3) say_out("gpuserver: target DB resident, served " + std::to_string(T->served) + " prefilter call(s) since it was staged\n");
    } else {
        if (compressed) {   // Prefiltering.cpp:710-714
            say_err("The output of the prefilter cannot be compressed during target split mode. Prefilter result will not be compressed.\n");
            compressed = false;
        }
        Db t; t.open(tPath);
        if ((size_t) a.split > t.size()) die("split was set to " + std::to_string(a.split) + " but the db to split has only " + std::to_string(t.size()) + " sequences. Please run with default paramerters");
        std::vector<uint8_t> tc; std::vector<uint64_t> to; std::vector<uint32_t> tl;
        encode(t, a.maskLowerCase != 0, tc, to, tl);
        const size_t four = (size_t) (4 * sqrt((double) a.cfg.max_seqs / (double) a.split));
        const int maxSeqs = (int) std::max<size_t>(1, (size_t) a.cfg.max_seqs / (size_t) a.split + four);
        Args as = a; as.cfg.max_seqs = maxSeqs;
        const std::vector<std::pair<size_t, size_t>> ranges = decompose_by_residues(t, (size_t) a.split);
        // One GPU: the ranges one after the other.  Several GPUs: GPU g takes the ranges g, g + G, ... and runs all queries against each -
        // the target (and its k-mer index, the large part) is sharded over the GPUs instead of replicated, so a DB whose index does not fit
        // one GPU's memory is searched by G of them; the merge below does not depend on the order the ranges finish in.
        const size_t G = a.gpus.size();
        std::vector<std::vector<std::vector<PH>>> perRange(ranges.size());
        auto one_range = [&](Args &ar, size_t sp) {
            if (ranges[sp].second == 0) return;
            if (a.verbosity >= 3) say_out("Process prefiltering step " + std::to_string(sp + 1) + " of " + std::to_string(a.split) + "\nTarget db start " +
                                          std::to_string(ranges[sp].first + 1) + " to " + std::to_string(ranges[sp].first + ranges[sp].second) + "\n");
            perRange[sp].assign(q.size(), std::vector<PH>());
            std::shared_ptr<PrefTarget> T = build_pref_target(ar, argv0, t.key, tc, to, tl, ranges[sp].first, ranges[sp].second, ql.size());
            pass(*T, identity(T->tkey), maxSeqs, ranges[sp].first, perRange[sp]);
        };
        if (G <= 1) for (size_t sp = 0; sp < ranges.size(); sp++) one_range(as, sp);
        else {
            const int hw = a.threads > 0 ? a.threads : usable_cpus();
            run_workers(G, [&](size_t g) {
                Args ag = as;
                ag.gpus.assign(1, a.gpus[g]); ag.cfg.device = a.gpus[g]; ag.threads = std::max(1, hw / (int) G);
                for (size_t sp = g; sp < ranges.size(); sp += G) one_range(ag, sp);
            });
        }
        for (size_t sp = 0; sp < ranges.size(); sp++)   // concatenated in split order (mergeTargetSplits reads the split files in order)
            for (size_t i = 0; i < perRange[sp].size(); i++) lists[i].insert(lists[i].end(), perRange[sp][i].begin(), perRange[sp][i].end());
        for (std::vector<PH> &l : lists)   // hit_t::compareHitsByScoreAndId (QueryMatcher.h:38-48)
            if (l.size() > 1) std::sort(l.begin(), l.end(), [](const PH &x, const PH &y) { return abs(x.score) != abs(y.score) ? abs(x.score) > abs(y.score) : x.key < y.key; });
    }
    return lists;
}

int prefilter_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 3) die("usage: foldseek-b200 prefilter <i:queryDB> <i:targetDB> <o:prefilterDB> [options]");
    if (a.gpuServer && g_resident == NULL) return forward_to_server(argc, argv, a.pos[1], a);
    use_server_layout(a);
    Db q; bool compressed = false;
    const std::vector<std::vector<PH>> lists = prefilter_core(a, argv[0], a.pos[0], a.pos[1], q, NULL, compressed);
    std::vector<std::string> out(q.size());
    for (size_t i = 0; i < q.size(); i++) {
        char buf[64];
        for (const PH &h : lists[i]) { snprintf(buf, sizeof(buf), "%u\t%d\t%d\n", h.key, h.score, h.diag); out[i] += buf; }
    }
    write_db(a.pos[2], q.key, out, 7 /* DBTYPE_PREFILTER_RES */, compressed);
    return EXIT_SUCCESS;
}

// The structurealign module behind its arguments.  The entries to align are given by `fill`: entryKey[p] = query key of result entry p,
// cand[candOff[p] .. candOff[p + 1]) = its candidate targets (DBReader ids of the target DB) in prefilter order - read from a prefilter DB
// (structurealign) or produced by the prefilter stage on the same contexts (search).  `units` = number of entries, known before the target
// is staged (it bounds the lanes of a one-off run).  Returns the alignment DB entries, one per entryKey.
struct AlignInput { std::vector<uint32_t> entryKey; std::vector<uint64_t> candOff; std::vector<uint32_t> cand; };
std::vector<std::string> structurealign_core(Args &a, const char *argv0, const std::string &qPath, const std::string &tPath, size_t units, bool withIndex,
                                             const std::function<void(AlnTarget &, const Db &, AlignInput &)> &fill, AlignInput &in) {
    const bool sameDB = qPath == tPath;   // structurealign.cpp:165-169
    Db qa, qs;
    qa.open(qPath); qs.open(qPath + "_ss");
    Db qCa;
    const bool ca1 = !slurp(qPath + "_ca.dbtype", false).empty(), ca2 = !slurp(tPath + "_ca.dbtype", false).empty();
    if (a.sortByStructureBits && !(ca1 && ca2)) {   // structurealign.cpp:176-196
        say_err("Cannot find query or target C-alpha database\nDisabling --sort-by-structure-bits\n");
        a.sortByStructureBits = 0;
    }
    if (a.sortByStructureBits && (a.alignmentMode == 1 || a.alignmentMode == 2)) {   // structurealign.cpp:188-196
        say_err("Cannot use --sort-by-structure-bits 1 with --alignment-mode 1 or 2\nDisabling --sort-by-structure-bits\n");
        a.sortByStructureBits = 0;
    }
    bool needTM = a.tmScoreThr > 0, needLDDT = a.lddtThr > 0;   // structurealign.cpp:207-226
    if (a.sortByStructureBits) { needTM = true; needLDDT = true; }
    else {
        if (needTM && !(ca1 && ca2)) { say_err("Cannot use --tmscore-threshold with --sort-by-structure-bits 0\nDisabling --tmscore-threshold\n"); needTM = false; }
        if (needLDDT && !(ca1 && ca2)) { say_err("Cannot use --lddt-threshold with --sort-by-structure-bits 0\nDisabling --lddt-threshold\n"); needLDDT = false; }
    }
    const bool needCa = needTM || needLDDT;
    if (needCa) qCa.open(qPath + "_ca");
    Resources res;
    std::vector<uint8_t> qca, qcs; std::vector<uint64_t> qo, o2; std::vector<uint32_t> ql, l2;
    encode(qa, false, qca, qo, ql); encode(qs, false, qcs, o2, l2);
    if (l2 != ql) die("query AA and 3Di databases differ in sequence lengths");
    std::shared_ptr<AlnTarget> T = aln_target(a, argv0, tPath, g_resident ? (size_t) -1 : units, needCa, withIndex);
    std::vector<fsb200_ctx *> &ctxs = T->ctxs;
    const std::vector<uint8_t> &tca = T->tca, &tcs = T->tcs;
    const std::vector<uint64_t> &to = T->to;
    const std::vector<uint32_t> &tl = T->tl;
    const uint64_t aaRes = T->aaRes;
    const std::vector<uint32_t> &tkeyA = T->tkey;   // target keys by DBReader id
    fill(*T, qa, in);
    struct { const std::vector<uint32_t> &key; size_t size() const { return key.size(); } } pref{in.entryKey};
    const std::vector<uint64_t> &candOff = in.candOff; const std::vector<uint32_t> &cand = in.cand;
    std::vector<uint32_t> keyToQ(qa.size() ? *std::max_element(qa.key.begin(), qa.key.end()) + 1 : 0, 0xFFFFFFFFu);
    for (size_t i = 0; i < qa.size(); i++) keyToQ[qa.key[i]] = (uint32_t) i;
    std::vector<uint32_t> keyToQCa;
    if (needCa) { keyToQCa.assign(qCa.size() ? *std::max_element(qCa.key.begin(), qCa.key.end()) + 1 : 0, 0xFFFFFFFFu); for (size_t i = 0; i < qCa.size(); i++) keyToQCa[qCa.key[i]] = (uint32_t) i; }
    std::vector<uint32_t> prefQ(pref.size());
    for (size_t p = 0; p < pref.size(); p++) {
        if (pref.key[p] >= keyToQ.size() || keyToQ[pref.key[p]] == 0xFFFFFFFFu) die("prefilter entry for unknown query key");
        prefQ[p] = keyToQ[pref.key[p]];
    }
    std::vector<std::string> out(pref.size());
    std::vector<uint64_t> w(pref.size());
    for (size_t p = 0; p < pref.size(); p++) w[p] = (candOff[p + 1] - candOff[p]) * (uint64_t) ql[prefQ[p]] + 1;
    const std::vector<size_t> bounds = split_by_weight(w, ctxs.size());
    if (g_resident) g_resident->touched = true;
    // A context's share of the entries is processed in batches of at most FSB200_ALIGN_BATCH candidate pairs (default 32 M: 40 B of
    // results per pair on the host, the same again in the library): an exhaustive search or very long prefilter lists against a large
    // DB must not need memory in proportion to queries x targets.  Batches never split an entry.
    const uint64_t batchPairs = getenv("FSB200_ALIGN_BATCH") ? std::max<uint64_t>(1, strtoull(getenv("FSB200_ALIGN_BATCH"), NULL, 10)) : (32ull << 20);
    run_workers(ctxs.size(), [&](size_t g) {
        fsb200_ctx *ctx = ctxs[g];
        for (size_t pb = bounds[g], pe; pb < bounds[g + 1]; pb = pe) {
            pe = pb + 1;
            while (pe < bounds[g + 1] && candOff[pe + 1] - candOff[pb] <= batchPairs) pe++;
            // one query batch = the queries of these prefilter entries, in prefilter-DB order
            std::vector<uint8_t> bqa, bqs; std::vector<uint64_t> bo; std::vector<uint32_t> bl;
            std::vector<uint64_t> cOff(1, 0);
            for (size_t p = pb; p < pe; p++) {
                const uint32_t qi = prefQ[p];
                bo.push_back(bqa.size()); bl.push_back(ql[qi]);
                bqa.insert(bqa.end(), qca.begin() + qo[qi], qca.begin() + qo[qi] + ql[qi]);
                bqs.insert(bqs.end(), qcs.begin() + qo[qi], qcs.begin() + qo[qi] + ql[qi]);
                cOff.push_back(candOff[p + 1] - candOff[pb]);
            }
            const uint32_t *cnd = cand.data() + candOff[pb];
            const size_t nc = (size_t) cOff.back();
            check(ctx, fsb200_set_queries(ctx, bqa.data(), bqs.data(), bo.data(), bl.data(), (uint32_t) bl.size()), "set_queries");
            if (needCa) {
                std::vector<float> x, y, z; std::vector<uint64_t> co(pe - pb);
                for (size_t p = pb; p < pe; p++) {
                    if (pref.key[p] >= keyToQCa.size() || keyToQCa[pref.key[p]] == 0xFFFFFFFFu) die("no C-alpha entry for a query key");
                    co[p - pb] = x.size();
                    decode_ca(qCa, keyToQCa[pref.key[p]], bl[p - pb], x, y, z);
                }
                check(ctx, fsb200_set_queries_ca(ctx, x.data(), y.data(), z.data(), co.data()), "set_queries_ca");
            }
            std::vector<fsb200_aln> aln(nc);
            check(ctx, fsb200_align(ctx, cOff.data(), cnd, aln.data()), "align");
            check(ctx, fsb200_backtrace_run(ctx), "backtrace");
            uint64_t btBytes = 0;
            check(ctx, fsb200_backtrace_size(ctx, &btBytes), "backtrace_size");
            std::vector<fsb200_bt> meta(nc);
            std::vector<char> bt(btBytes + 1);
            check(ctx, fsb200_backtrace_download(ctx, meta.data(), bt.data(), bt.size()), "backtrace_download");
            std::vector<fsb200_sscore> sscore;
            if (needCa) {
                sscore.resize(nc);
                check(ctx, fsb200_structure_scores_run(ctx), "structure_scores");
                check(ctx, fsb200_structure_scores_download(ctx, sscore.data()), "structure_scores_download");
            }
            struct Row { uint32_t dbKey; int score; float seqId; double eval; int qStart, qEnd, dbStart, dbEnd; uint32_t dbLen; std::string cigar; };
            // What alignStructure (structurealign.cpp:37-112) leaves in `res` for one pair, from the library's answers: false where it returns -1
            // (coverage of the end positions, the two E-value gates).  ok = Alignment::checkCriteria of the finished row.
            auto build_row = [&](const fsb200_aln &al, const fsb200_bt &m, const char *btBuf, uint32_t dbKey, uint32_t dbLen, uint32_t qLen, bool isIdentity, Row &r, bool &ok) -> bool {
                if (al.status != 0) return false;
                if (!has_coverage(a.covThr, a.covMode, compute_cov(0, (unsigned) al.q_end, qLen), compute_cov(0, (unsigned) al.db_end, dbLen))) return false;
                r.dbKey = dbKey; r.score = al.score; r.eval = al.evalue;
                r.qEnd = al.q_end; r.dbEnd = al.db_end; r.dbLen = dbLen;
                r.qStart = -1; r.dbStart = -1; r.seqId = 0.0f;   // alignScoreEndPos leaves -1 (StructureSmithWaterman.cpp:274-275)
                r.cigar.clear();
                float qCov = compute_cov(0, (unsigned) r.qEnd, qLen), tCov = compute_cov(0, (unsigned) r.dbEnd, r.dbLen);
                int alnLen;
                const bool traced = m.status == 1;
                if (traced) {
                    r.qStart = m.q_start; r.dbStart = m.db_start;
                    qCov = compute_cov((unsigned) r.qStart, (unsigned) r.qEnd, qLen); tCov = compute_cov((unsigned) r.dbStart, (unsigned) r.dbEnd, r.dbLen);   // :521-522
                }
                if (traced && m.aln_len > 0) {
                    alnLen = (int) m.aln_len;
                    // Util::computeSeqId (Util.cpp:597-607): SEQ_ID_ALN_LEN 0, SEQ_ID_SHORT 1, SEQ_ID_LONG 2
                    const int denom = a.seqIdMode == 1 ? (int) std::min(qLen, r.dbLen) : a.seqIdMode == 2 ? (int) std::max(qLen, r.dbLen) : alnLen;
                    r.seqId = (float) (int) m.identical_aa / (float) denom;
                    // Matcher::compressAlignment (Matcher.cpp:168-187)
                    const char *b = btBuf + m.bt_off;
                    char state = 'M'; size_t cnt = 0;
                    for (uint32_t x = 0; x < m.aln_len; x++) {
                        if (b[x] != state) { r.cigar += std::to_string(cnt); r.cigar.push_back(state); state = b[x]; cnt = 1; }
                        else cnt++;
                    }
                    r.cigar += std::to_string(cnt); r.cigar.push_back(state);
                } else {
                    alnLen = std::max(abs(r.qEnd - r.qStart), abs(r.dbEnd - r.dbStart)) + 1;   // Matcher::computeAlnLength
                    r.cigar = "0M";                                                           // compressAlignment("")
                }
                // Alignment::checkCriteria (lib/mmseqs/src/alignment/Alignment.cpp:548-567)
                ok = isIdentity || (r.eval <= a.cfg.evalue_thr && r.seqId >= a.seqIdThr && has_coverage(a.covThr, a.covMode, qCov, tCov) && alnLen >= a.alnLenThr);
                return true;
            };
            // rows[p - pb]: the accepted rows of one prefilter entry in the order the reference pushes them (every accepted pair, then its
            // alternative alignments); chains: the accepted pairs --alt-ali continues with
            struct Chain { size_t p; size_t at; uint32_t tIdx; int dbStart, dbEnd; std::vector<uint8_t> aa, ss; };
            std::vector<std::vector<std::vector<Row>>> rows(pe - pb);
            std::vector<Chain> chains;
            for (size_t p = pb; p < pe; p++) {
                std::vector<std::vector<Row>> &mine = rows[p - pb];
                const uint32_t qLen = bl[p - pb];
                int passedNum = 0, rejected = 0;   // structurealign.cpp:348-350
                for (uint64_t k = cOff[p - pb]; k < cOff[p - pb + 1] && passedNum < a.maxAccept && rejected < a.maxRejected; k++) {
                    const uint32_t tIdx = cnd[k];
                    const bool isIdentity = (prefQ[p] == tIdx) && (a.includeIdentity || sameDB);
                    if (!can_be_covered(a.covThr, a.covMode, (float) qLen, (float) tl[tIdx])) { rejected++; continue; }   // :364-367
                    Row r; bool ok = false;
                    if (!build_row(aln[k], meta[k], bt.data(), tkeyA[tIdx], tl[tIdx], qLen, isIdentity, r, ok) || !ok) { rejected++; continue; }
                    if (needCa) {   // structurealign.cpp:376-412
                        if (sscore[k].status != 1) {
                            // no backtrace or a degenerate superposition: the reference scores an empty alignment / falls back to the classic
                            // Kabsch (lib/tmalign/Kabsch.h:761).  Neither is reproduced: the run goes on like the reference's, the row keeps its
                            // Smith-Waterman bit score and the thresholds are not applied to it; FSB200_STRICT=1 stops instead.
                            const std::string msg = "structure scores not computed for query key " + std::to_string(pref.key[p]) + " target key " + std::to_string(r.dbKey) +
                                                    " (degenerate superposition or empty backtrace): row written with its unrescored score";
                            if (getenv("FSB200_STRICT")) die(msg);
                            say_err("foldseek-b200: WARNING: " + msg + "\n");
                        } else {
                            if (needTM && sscore[k].tm < a.tmScoreThr) continue;    // neither accepted nor rejected (:393-395)
                            if (needLDDT && sscore[k].lddt < a.lddtThr) continue;
                            if (a.sortByStructureBits) r.score = sscore[k].score_bits;
                        }
                    }
                    if (a.altAli > 0) chains.push_back(Chain{p, mine.size(), tIdx, r.dbStart, r.dbEnd, {}, {}});
                    mine.emplace_back(1, r);
                    passedNum++; rejected = 0;
                }
            }
            // --alt-ali N (computeAlternativeAlignment, structurealign.cpp:115-138; loop :415-429): the target residues of the last alignment,
            // [dbStartPos, dbEndPos), become X in both sequences (cumulatively), the pair is aligned again by alignStructure, and the result is
            // kept if it meets checkCriteria; up to N times per accepted pair, the first failure ends the pair's series.  The rows keep their
            // Smith-Waterman score (the reference does not rescore them).  Here: all pairs of a round at once - their masked targets are staged
            // as a small target DB of a second context on the same GPU and aligned with the library's usual calls.
            if (!chains.empty()) {
                struct CtxGuard { fsb200_ctx *c; ~CtxGuard() { fsb200_destroy(c); } } altGuard{make_ctx(a, argv0, res, a.gpus[g % a.gpus.size()])};
                fsb200_ctx *alt = altGuard.c;
                check(alt, fsb200_set_queries(alt, bqa.data(), bqs.data(), bo.data(), bl.data(), (uint32_t) bl.size()), "set_queries (alt-ali)");
                for (Chain &c : chains) {
                    c.aa.assign(tca.begin() + to[c.tIdx], tca.begin() + to[c.tIdx] + tl[c.tIdx]);
                    c.ss.assign(tcs.begin() + to[c.tIdx], tcs.begin() + to[c.tIdx] + tl[c.tIdx]);
                }
                for (int round = 0; round < a.altAli && !chains.empty(); round++) {
                    std::vector<uint8_t> maa, mss; std::vector<uint64_t> mo; std::vector<uint32_t> ml, mc;
                    std::vector<uint64_t> mOff(pe - pb + 1, 0);
                    std::vector<Chain> live;
                    for (Chain &c : chains) {
                        if (c.dbStart < 0) {   // no start position (block aligner did not reproduce the score; the reference's fallback a16 is not built)
                            const std::string msg = "--alt-ali: no start position for query key " + std::to_string(pref.key[c.p]) + " target key " + std::to_string(tkeyA[c.tIdx]) + ": no alternative alignments for this pair";
                            if (getenv("FSB200_STRICT")) die(msg);
                            say_err("foldseek-b200: WARNING: " + msg + "\n");
                            continue;
                        }
                        for (int pos = c.dbStart; pos < c.dbEnd; pos++) { c.aa[pos] = 20; c.ss[pos] = 20; }
                        live.push_back(std::move(c));
                    }
                    chains.swap(live);
                    if (chains.empty()) break;
                    for (const Chain &c : chains) {   // chains stay ordered by prefilter entry: the candidate CSR is contiguous per query
                        mo.push_back(maa.size()); ml.push_back(tl[c.tIdx]); mc.push_back((uint32_t) mc.size());
                        maa.insert(maa.end(), c.aa.begin(), c.aa.end()); mss.insert(mss.end(), c.ss.begin(), c.ss.end());
                        mOff[c.p - pb + 1]++;
                    }
                    for (size_t i = 1; i < mOff.size(); i++) mOff[i] += mOff[i - 1];
                    check(alt, fsb200_load_target(alt, maa.data(), mss.data(), mo.data(), ml.data(), (uint32_t) ml.size(), aaRes), "load_target (alt-ali)");
                    std::vector<fsb200_aln> aaln(mc.size());
                    check(alt, fsb200_align(alt, mOff.data(), mc.data(), aaln.data()), "align (alt-ali)");
                    check(alt, fsb200_backtrace_run(alt), "backtrace (alt-ali)");
                    uint64_t abytes = 0;
                    check(alt, fsb200_backtrace_size(alt, &abytes), "backtrace_size (alt-ali)");
                    std::vector<fsb200_bt> ameta(mc.size()); std::vector<char> abt(abytes + 1);
                    check(alt, fsb200_backtrace_download(alt, ameta.data(), abt.data(), abt.size()), "backtrace_download (alt-ali)");
                    live.clear();
                    for (size_t i = 0; i < chains.size(); i++) {
                        Chain &c = chains[i];
                        Row r; bool ok = false;
                        if (!build_row(aaln[i], ameta[i], abt.data(), tkeyA[c.tIdx], tl[c.tIdx], bl[c.p - pb], false, r, ok) || !ok) continue;
                        rows[c.p - pb][c.at].push_back(r);
                        c.dbStart = r.dbStart; c.dbEnd = r.dbEnd;
                        live.push_back(std::move(c));
                    }
                    chains.swap(live);
                }
            }
            for (size_t p = pb; p < pe; p++) {
                const uint32_t qLen = bl[p - pb];
                std::vector<Row> flat;
                for (std::vector<Row> &grp : rows[p - pb]) for (Row &r : grp) flat.push_back(std::move(r));
                // compareHitsByStructureBits (structurealign.cpp:21-29) or Matcher::compareHits (Matcher.h:161-172)
                const bool bits = a.sortByStructureBits != 0;
                std::sort(flat.begin(), flat.end(), [bits](const Row &x, const Row &y) {
                    if (!bits && x.eval != y.eval) return x.eval < y.eval;
                    if (x.score != y.score) return x.score > y.score;
                    if (x.dbLen != y.dbLen) return x.dbLen < y.dbLen;
                    return x.dbKey < y.dbKey;
                });
                char buf[256];
                for (const Row &r : flat) {
                    // Util::fastSeqIdToBuffer (Util.cpp:251-280); for seqId == 1.0 it returns the position OF its terminator, so the tab
                    // that Matcher::resultToBuffer writes at tmpBuff - 1 (Matcher.cpp:288-289) replaces the last digit: "1.00"
                    char sid[16];
                    if (r.seqId == 1.0f) snprintf(sid, sizeof(sid), "1.00");
                    else snprintf(sid, sizeof(sid), "0.%s%s%d", r.seqId < 0.10 ? "0" : "", r.seqId < 0.01 ? "0" : "", (int) (r.seqId * 1000));
                    snprintf(buf, sizeof(buf), "%u\t%d\t%s\t%.3E\t%d\t%d\t%u\t%d\t%d\t%u", r.dbKey, r.score, sid, r.eval,
                             r.qStart, r.qEnd, qLen, r.dbStart, r.dbEnd, r.dbLen);
                    out[p] += buf;
                    if (a.addBacktrace) { out[p] += '\t'; out[p] += r.cigar; }
                    out[p] += '\n';
                }
            }
        }
    });
    T->served++;
    if (g_resident && a.verbosity >= 3) say_out("gpuserver: target DB resident, served " + std::to_string(T->served) + " structurealign call(s) since it was staged\n");
    return out;
}

int structurealign_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: foldseek-b200 structurealign <i:queryDB> <i:targetDB> <i:prefilterDB> <o:alnDB> [options]");
    if (a.gpuServer && g_resident == NULL) return forward_to_server(argc, argv, a.pos[1], a);
    use_server_layout(a);
    Db pref; pref.open(a.pos[2]);
    AlignInput in;
    in.entryKey = pref.key;
    // candidate lists: the hits of every prefilter entry, in prefilter order
    const std::vector<std::string> out = structurealign_core(a, argv[0], a.pos[0], a.pos[1], pref.size(), false, [&](AlnTarget &T, const Db &, AlignInput &x) {
        x.candOff.assign(1, 0);
        for (size_t p = 0; p < pref.size(); p++) {
            const char *sp = pref.data.data() + pref.off[p];
            while (*sp != '\0') {
                const unsigned long k = strtoul(sp, NULL, 10);
                if (k >= T.keyToT.size() || T.keyToT[k] == 0xFFFFFFFFu) die("prefilter hit for unknown target key");
                x.cand.push_back(T.keyToT[k]);
                while (*sp != '\n' && *sp != '\0') sp++;
                if (*sp == '\n') sp++;
            }
            x.candOff.push_back(x.cand.size());
        }
    }, in);
    write_db(a.pos[3], pref.key, out, 5 /* DBTYPE_ALIGNMENT_RES */, a.compressed != 0);
    return EXIT_SUCCESS;
}


// foldseek-b200 search <queryDB> <targetDB> <alnDB> <tmpDir>: the search workflow's two stages (src/workflow/StructureSearch.cpp,
// data/structuresearch.sh:50,121: `prefilter` on the _ss DBs, then `structurealign` on its result) in one process.  The hit lists never
// become a prefilter DB on disk: they are the candidate lists of the alignment stage, and both stages run on ONE staged copy of the
// target (AA + unmasked 3Di + C-alpha + k-mer index) - what bench.py's end-to-end number measures through fsb200_search.  Only when the
// target's 3Di DB has masked (lower-case) residues and --mask-lower-case 1 asks the prefilter to see them as X, or with a target split,
// does the prefilter stage its own copy, as the separate modules do.  The alignment DB is the one the two modules write when chained.
int search_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: foldseek-b200 search <i:queryDB> <i:targetDB> <o:alignmentDB> <tmpDir> [options]");
    if (a.gpuServer && g_resident == NULL) return forward_to_server(argc, argv, a.pos[1], a);
    use_server_layout(a);
    // the workflow fixes the composition-bias scale per module, whatever the command line says (StructureSearch.cpp:101-107)
    a.cfg.comp_bias_scale_prefilter = 0.15f; a.cfg.comp_bias_scale_align = 0.5f;
    const bool targetSplit = a.split > 1 && a.splitMode == 0;
    AlignInput in;
    bool compressedPref = false;
    const std::vector<std::string> out = structurealign_core(a, argv[0], a.pos[0], a.pos[1], (size_t) -1, !targetSplit && !a.exhaustive, [&](AlnTarget &T, const Db &qa, AlignInput &x) {
        if (a.exhaustive) {   // PREFMODE=EXHAUSTIVE (data/structuresearch.sh:27-41 fake_pref): the candidate list of every query is the target DB's index, in its order
            x.entryKey = qa.key;
            x.candOff.assign(1, 0);
            for (size_t i = 0; i < qa.size(); i++) {
                for (uint32_t t = 0; t < (uint32_t) T.tkey.size(); t++) x.cand.push_back(t);
                x.candOff.push_back(x.cand.size());
            }
            return;
        }
        Db q;
        std::vector<std::vector<PH>> lists;
        if (!targetSplit && T.hasIndex && !(a.maskLowerCase != 0 && T.masked3Di)) {
            PrefTarget view; view.owned = false; view.tkey = T.tkey; view.tl = T.tl; view.ctxs = T.ctxs;
            lists = prefilter_core(a, argv[0], a.pos[0] + "_ss", a.pos[1] + "_ss", q, &view, compressedPref);
        } else lists = prefilter_core(a, argv[0], a.pos[0] + "_ss", a.pos[1] + "_ss", q, NULL, compressedPref);
        if (q.key != qa.key) die("query AA and 3Di databases differ in their keys");
        x.entryKey = q.key;
        x.candOff.assign(1, 0);
        for (const std::vector<PH> &l : lists) {
            for (const PH &h : l) x.cand.push_back(h.tIdx);
            x.candOff.push_back(x.cand.size());
        }
    }, in);
    write_db(a.pos[2], in.entryKey, out, 5 /* DBTYPE_ALIGNMENT_RES */, a.compressed != 0);
    return EXIT_SUCCESS;
}

#include "cli_server.h"
#include "cli_convertalis.h"

// foldseek-b200 easy-search <queryDB> <targetDB> <alignmentFile> <tmpDir>: data/easystructuresearch.sh for inputs that are databases
// already (the script skips createdb when "<input>.dbtype" exists, :14-33): search into <tmpDir>/result, then convertalis.  Structure
// files as input need the reference's createdb (PDB / mmCIF parsing and the 3Di encoder are outside this path) and are refused by name.
int easy_search_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: foldseek-b200 easy-search <i:queryDB> <i:targetDB> <o:alignmentFile> <tmpDir> [options]");
    for (int k = 0; k < 2; k++)
        if (slurp(a.pos[k] + ".dbtype", false).empty())
            die(a.pos[k] + " is not a database: foldseek-b200 easy-search takes the databases `foldseek createdb` writes (structure files are converted by the reference's createdb)");
    mkdir(a.pos[3].c_str(), 0777);
    const std::string result = a.pos[3] + "/result";
    std::vector<std::string> flags;
    bool removeTmp = true;   // the easy workflows remove their temporary files by default
    for (int i = 2; i < argc; i++) {
        if (std::find(a.posIdx.begin(), a.posIdx.end(), i) != a.posIdx.end()) continue;
        if (strcmp(argv[i], "--remove-tmp-files") == 0 && i + 1 < argc) removeTmp = atoi(argv[i + 1]) != 0;
        flags.push_back(argv[i]);
    }
    auto call = [&](const char *module, std::vector<std::string> pos) {
        std::vector<std::string> av{argv[0], module};
        av.insert(av.end(), pos.begin(), pos.end()); av.insert(av.end(), flags.begin(), flags.end());
        std::vector<char *> cv; for (std::string &t : av) cv.push_back(&t[0]);
        return run_module((int) cv.size(), cv.data());
    };
    int rc = call("search", {a.pos[0], a.pos[1], result, a.pos[3]});
    if (rc != EXIT_SUCCESS) return rc;
    rc = call("convertalis", {a.pos[0], a.pos[1], result, a.pos[2]});
    if (removeTmp) for (const char *ext : {"", ".index", ".dbtype"}) unlink((result + ext).c_str());
    return rc;
}

const char *HELP =
    "foldseek-b200 - B200-native Foldseek search hot path (3Di k-mer prefilter + structurealign), host modules on MMseqs2 databases\n\n"
    "  easy-search    <queryDB> <targetDB> <aln.m8> <tmpDir>        search + convertalis (inputs: databases written by `foldseek createdb`)\n"
    "  search         <queryDB> <targetDB> <alnDB> <tmpDir>         both stages in one process on one staged copy of the target\n"
    "  prefilter      <queryDB_ss> <targetDB_ss> <prefilterDB>      stage 1, the reference module's result DB\n"
    "  structurealign <queryDB> <targetDB> <prefilterDB> <alnDB>    stage 2, the reference module's result DB\n"
    "  convertalis    <queryDB> <targetDB> <alnDB> <aln.m8>         default 12-column BLAST-tab output\n"
    "  gpuserver      <targetDB>                                    keep the target resident; serve calls that carry --gpu-server 1\n\n"
    "Every parameter the reference workflow serialises for these modules is known. Implemented with the reference's results:\n"
    "  -s  --max-seqs  -k 6  --comp-bias-corr(-scale)  --min-ungapped-score  --mask-lower-case  -c  --cov-mode 0-5  --add-self-matches\n"
    "  --split N --split-mode 0|1  -e  --min-seq-id  --seq-id-mode  --min-aln-len  --max-accept  --max-rejected  -a  --alt-ali\n"
    "  --alignment-type 0|2  --alignment-mode  --gap-open  --gap-extend  --sort-by-structure-bits  --tmscore-threshold(-mode)\n"
    "  --lddt-threshold  --compressed  --threads  -v\n"
    "This tool: --gpu N | --gpu-list a,b,..  (GPUs of this process)   --lanes N  (contexts per GPU sharing the staged target)\n"
    "           --gpu-server 1  --gpu-server-wait-timeout S  (run the call on the gpuserver of the target DB)\n"
    "A value that asks for behaviour that is not built ends the run with an error naming the parameter; nothing is ignored silently.\n"
    "A CUDA device is required: there is no CPU path.\n";

int run_module(int argc, char **argv) {
    for (int i = 1; i < argc; i++)
        if (strcmp(argv[i], "-h") == 0 || strcmp(argv[i], "--help") == 0 || (i == 1 && strcmp(argv[i], "help") == 0)) { say_out(HELP); return EXIT_SUCCESS; }
    if (strcmp(argv[1], "easy-search") == 0) return easy_search_main(argc, argv);
    if (strcmp(argv[1], "search") == 0) return search_main(argc, argv);
    if (strcmp(argv[1], "prefilter") == 0) return prefilter_main(argc, argv);
    if (strcmp(argv[1], "structurealign") == 0) return structurealign_main(argc, argv);
    if (strcmp(argv[1], "gpuserver") == 0) return gpuserver_main(argc, argv);
    if (strcmp(argv[1], "convertalis") == 0) return convertalis_main(argc, argv);
    die(std::string("unknown module ") + argv[1]);
}

}  // namespace

int main(int argc, char **argv) {
    try {
        if (argc < 2) die("usage: foldseek-b200 <easy-search|search|prefilter|structurealign|gpuserver|convertalis> ...");
        return run_module(argc, argv);
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return EXIT_FAILURE;
    }
}
