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

// Debug(Debug::ERROR) << ...; EXIT(EXIT_FAILURE): a module run ends with this message and a non-zero exit code.  Thrown instead of
// exiting on the spot because the same module code runs inside the gpuserver, where the error goes back to the calling process.
struct Fatal : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] void die(const std::string &msg) { throw Fatal(msg); }

// stdout / stderr of a module run: the process's own streams, or (inside the gpuserver) the text returned to the client
struct Console {
    std::mutex m; bool capture = false; std::string out, err;
    void put(bool toErr, const std::string &t) {
        std::lock_guard<std::mutex> g(m);
        if (capture) (toErr ? err : out) += t;
        else { fputs(t.c_str(), toErr ? stderr : stdout); fflush(toErr ? stderr : stdout); }
    }
} g_console;
void say_err(const std::string &t) { g_console.put(true, t); }
void say_out(const std::string &t) { g_console.put(false, t); }

// one host thread per context; the first error of any of them ends the run like an error on the main thread
template <class F> void run_workers(size_t n, F fn) {
    std::vector<std::thread> th; std::mutex m; std::string first; bool failed = false;
    for (size_t g = 0; g < n; g++) th.emplace_back([&, g]() {
        try { fn(g); } catch (const std::exception &e) { std::lock_guard<std::mutex> l(m); if (!failed) { failed = true; first = e.what(); } }
    });
    for (std::thread &t : th) t.join();
    if (failed) throw Fatal(first);
}

std::string slurp(const std::string &path, bool required = true) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { if (required) die("cannot read " + path); return std::string(); }
    std::ostringstream ss; ss << in.rdbuf();
    return ss.str();
}

// zstd frames of a compressed DB are inflated with the system's libzstd, looked up at run time (the image has the library but no
// headers; the two entry points used are part of zstd's stable ABI)
size_t zstd_inflate(void *dst, size_t cap, const void *src, size_t n) {
    typedef size_t (*dec_t)(void *, size_t, const void *, size_t);
    typedef unsigned (*err_t)(size_t);
    static dec_t dec = NULL; static err_t isErr = NULL;
    if (dec == NULL) {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (h == NULL) h = dlopen("libzstd.so", RTLD_NOW);
        if (h == NULL) die("compressed database: libzstd.so.1 not found on this system");
        dec = (dec_t) dlsym(h, "ZSTD_decompress"); isErr = (err_t) dlsym(h, "ZSTD_isError");
        if (dec == NULL || isErr == NULL) die("compressed database: libzstd lacks ZSTD_decompress");
    }
    const size_t r = dec(dst, cap, src, n);
    if (isErr(r)) die("compressed database: zstd frame cannot be decompressed");
    return r;
}

struct Db {   // read-only MMseqs2 DB (per-thread shards <name>.0, .1, ... are concatenated; compressed DBs are inflated on open)
    std::string data;
    uint64_t fileBytes = 0;        // size of the data file(s) as stored: DBReader::getAminoAcidDBSize() counts these bytes
    std::vector<uint32_t> key; std::vector<uint64_t> off, len;
    int dbtype = 0;
    bool padded = false;           // DBTYPE_EXTENDED_GPU (lib/mmseqs/src/util/makepaddedseqdb.cpp:27): numeric codes, no "\n\0"
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
        std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> rows;
        while (idx >> k >> o >> l) rows.push_back({(uint32_t) k, {o, l}});
        // DBReader ids follow the key order of the index (DBReader::open sorts an unsorted index by key)
        std::stable_sort(rows.begin(), rows.end(), [](const std::pair<uint32_t, std::pair<uint64_t, uint64_t>> &x, const std::pair<uint32_t, std::pair<uint64_t, uint64_t>> &y) { return x.first < y.first; });
        for (auto &r : rows) { key.push_back(r.first); off.push_back(r.second.first); len.push_back(r.second.second); }
        std::string t = slurp(name + ".dbtype");
        if (t.size() >= 4) memcpy(&dbtype, t.data(), 4);
        fileBytes = data.size();
        if (dbtype & (1 << 31)) {
            // DBWriter::writeEnd (lib/mmseqs/src/commons/DBWriter.cpp:331-398): u32 stored size, the bytes, then '\0' behind a zstd frame or
            // 0xFF behind an entry stored as is; the index keeps the ORIGINAL length (incl. the terminator).  DBReader::getDataCompressed
            // (DBReader.cpp:374-401).
            std::string plain;
            std::vector<uint64_t> noff(key.size());
            for (size_t i = 0; i < key.size(); i++) {
                if (off[i] + 4 > data.size()) die(name + ": truncated compressed entry");
                uint32_t cSize; memcpy(&cSize, data.data() + off[i], 4);
                const char *body = data.data() + off[i] + 4;
                if (off[i] + 4 + (uint64_t) cSize + 1 > data.size()) die(name + ": truncated compressed entry");
                noff[i] = plain.size();
                const uint64_t orig = len[i] > 0 ? len[i] - 1 : 0;
                if (body[cSize] == 0) {
                    std::string buf(orig + 64, '\0');
                    const size_t got = zstd_inflate(&buf[0], buf.size(), body, cSize);
                    if (got != orig) die(name + ": compressed entry does not have the length of its index line");
                    plain.append(buf.data(), got);
                } else {
                    if (cSize != orig) die(name + ": stored entry does not have the length of its index line");
                    plain.append(body, cSize);
                }
                plain.push_back('\0');
            }
            data.swap(plain); off.swap(noff);
            dbtype &= ~(1 << 31);
        }
        // extended flags live in bits 16..30 (DBReader.h:408-415); DBTYPE_EXTENDED_GPU = 8 (Parameters.h:94)
        padded = (((uint32_t) dbtype >> 16) & 0x7FFEu & 8u) != 0;
    }
    size_t size() const { return key.size(); }
};

// one zstd frame per entry for --compressed 1 (DBWriter.cpp:289-398; any valid frame is read back by DBReader::getDataCompressed)
std::string zstd_deflate(const std::string &in) {
    typedef size_t (*cmp_t)(void *, size_t, const void *, size_t, int);
    typedef size_t (*bnd_t)(size_t);
    typedef unsigned (*err_t)(size_t);
    static cmp_t cmp = NULL; static bnd_t bound = NULL; static err_t isErr = NULL;
    if (cmp == NULL) {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (h == NULL) h = dlopen("libzstd.so", RTLD_NOW);
        if (h == NULL) die("--compressed 1: libzstd.so.1 not found on this system");
        cmp = (cmp_t) dlsym(h, "ZSTD_compress"); bound = (bnd_t) dlsym(h, "ZSTD_compressBound"); isErr = (err_t) dlsym(h, "ZSTD_isError");
        if (cmp == NULL || bound == NULL || isErr == NULL) die("--compressed 1: libzstd lacks ZSTD_compress");
    }
    std::string out(bound(in.size()), '\0');
    const size_t n = cmp(&out[0], out.size(), in.data(), in.size(), 3);
    if (isErr(n)) die("--compressed 1: zstd compression failed");
    out.resize(n);
    return out;
}

void write_db(const std::string &name, const std::vector<uint32_t> &keys, const std::vector<std::string> &entries, int dbtype, bool compressed = false) {
    std::ofstream d(name, std::ios::binary), i(name + ".index");
    if (!d || !i) die("cannot write " + name);
    uint64_t off = 0;
    for (size_t k = 0; k < keys.size(); k++) {
        if (compressed) {   // u32 frame size, frame, '\0'; the index keeps the original length (DBWriter.cpp:359-396)
            const std::string frame = zstd_deflate(entries[k]);
            const uint32_t n = (uint32_t) frame.size();
            d.write(reinterpret_cast<const char *>(&n), 4); d.write(frame.data(), frame.size()); d.put('\0');
            i << keys[k] << '\t' << off << '\t' << (entries[k].size() + 1) << '\n';
            off += 4 + frame.size() + 1;
            continue;
        }
        d.write(entries[k].data(), entries[k].size()); d.put('\0');
        i << keys[k] << '\t' << off << '\t' << (entries[k].size() + 1) << '\n';
        off += entries[k].size() + 1;
    }
    if (compressed) dbtype |= (int) (1u << 31);
    d.close(); i.close();
    std::ofstream t(name + ".dbtype", std::ios::binary);   // last: resume guard
    t.write(reinterpret_cast<const char *>(&dbtype), 4);
}

// letters -> codes (BaseMatrix alphabet A C D E F G H I K L M N P Q R S T V W Y X; unknown -> X; lower == upper).
// Padded DBs (makepaddedseqdb.cpp:66-98) already hold the codes, +32 for lower-case/masked residues (DBReader.cpp:349-372).
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
            if (db.padded) {
                const unsigned char base = (c >= 32) ? c - 32 : c;
                codes[off[i] + k] = (maskLower && c >= 32) ? 20 : (base > 20 ? 20 : base);
            } else {
                // --mask-lower-case 1 (Foldseek default, LocalParameters.cpp:382-385): lower-case target residues are masked
                codes[off[i] + k] = (maskLower && islower(c)) ? 20 : map[c];
            }
        }
    }
}

struct Args {
    std::vector<std::string> pos;
    fsb200_config cfg;
    int verbosity = 3;
    int addBacktrace = 0;           // -a
    int sortByStructureBits = 1;    // LocalParameters.cpp: default 1
    float covThr = 0.0f; int covMode = 0;           // -c / --cov-mode
    float seqIdThr = 0.0f; int alnLenThr = 0;       // --min-seq-id / --min-aln-len
    int maxAccept = INT_MAX, maxRejected = INT_MAX; // --max-accept / --max-rejected
    float tmScoreThr = 0.0f, lddtThr = 0.0f;        // --tmscore-threshold / --lddt-threshold
    int maskLowerCase = 1;                          // --mask-lower-case
    int includeIdentity = 0;                        // --add-self-matches
    int compressed = 0;                             // --compressed: zstd frames in the result DB (and read from any input DB)
    int seqIdMode = 0;                              // --seq-id-mode: 0 alignment length, 1 shorter, 2 longer sequence (Util::computeSeqId)
    int alignmentMode = 0;                          // --alignment-mode: 1 / 2 only switch the structure-bits rescoring off (structurealign.cpp:188-191)
    int altAli = 0;                                 // --alt-ali: alternative alignments per accepted pair (structurealign.cpp:115-138,415-429)
    std::vector<int> gpus;                          // --gpu-list a,b,.. (this tool) or --gpu N (first N devices)
    int lanes = 1;                                  // --lanes N (this tool): contexts per GPU that share the staged DB and index
    int threads = 0;                                // --threads: host threads, split over the contexts
    int split = 1, splitMode = 2;                   // --split N --split-mode 0: target DB split (Prefiltering.cpp:384-397,412-526); mode 1 (query split) never changes results
    int gpuServer = 0;                              // --gpu-server 1: hand the call to the gpuserver of the target DB
    int gpuServerWait = 600;                        // --gpu-server-wait-timeout: 0 don't wait, -1 no limit, > 0 seconds (Parameters.cpp:2669)
};

// Every parameter the reference serialises for the two modules (createParameterString of par.prefilter / par.structurealign,
// lib/mmseqs/src/commons/Parameters.cpp:472-505, src/commons/LocalParameters.cpp:160-167 + the MMseqs2 align list) is known
// here.  A parameter is either implemented, or accepted only with the value the implemented behaviour corresponds to; any other
// value, and any unknown flag, ends the run with an error that names it (nothing is swallowed).
Args parse(int argc, char **argv, int first) {
    Args a;
    fsb200_default_config(&a.cfg);
    auto typed = [](const char *v) -> std::string {   // "aa:10,nucl:10" -> "10"; "seq:..,prof:.." -> first value
        std::string s(v); const size_t c = s.find(':');
        if (c == std::string::npos) return s;
        const size_t e = s.find(',', c);
        return s.substr(c + 1, e == std::string::npos ? std::string::npos : e - c - 1);
    };
    for (int i = first; i < argc; i++) {
        std::string s = argv[i];
        auto val = [&]() -> const char * { if (i + 1 >= argc) die("missing value for " + s); return argv[++i]; };
        auto only = [&](const std::vector<std::string> &ok) {   // accepted only with these values
            const std::string v = val();
            for (const std::string &o : ok) if (v == o || typed(v.c_str()) == o) return;
            bool num = false; double d = 0; { char *e = NULL; d = strtod(typed(v.c_str()).c_str(), &e); num = e && *e == 0; }
            if (num) for (const std::string &o : ok) { char *e = NULL; const double od = strtod(o.c_str(), &e); if (e && *e == 0 && od == d) return; }
            die("parameter " + s + " " + v + " is not implemented by foldseek-b200 (supported: " + ok[0] + ")");
        };
        if (s == "-s") a.cfg.sensitivity = (float) atof(val());
        else if (s == "-k") { const int k = atoi(val()); if (k != 0 && k != 6) die("parameter -k " + std::to_string(k) + " is not implemented by foldseek-b200 (supported: 0 = automatic, 6)"); a.cfg.kmer_size = 6; }
        else if (s == "--max-seqs") a.cfg.max_seqs = atoi(val());
        else if (s == "-e") a.cfg.evalue_thr = atof(val());
        else if (s == "--comp-bias-corr") a.cfg.comp_bias_corr = atoi(val());
        else if (s == "--comp-bias-corr-scale") { float v = (float) atof(val()); a.cfg.comp_bias_scale_prefilter = v; a.cfg.comp_bias_scale_align = v; }
        else if (s == "--min-ungapped-score") a.cfg.min_diag_score = atoi(val());
        else if (s == "--gap-open") a.cfg.gap_open = atoi(typed(val()).c_str());
        else if (s == "--gap-extend") a.cfg.gap_extend = atoi(typed(val()).c_str());
        else if (s == "--alignment-type") { const int t = atoi(val()); if (t != 0 && t != 2) die("parameter --alignment-type " + std::to_string(t) + " is not implemented by foldseek-b200 structurealign (supported: 0, 2)"); a.cfg.aa_factor = (t == 2) ? 1.4f : 0.0f; }
        else if (s == "--gpu") { const int n = atoi(val()); a.gpus.clear(); for (int g = 0; g < std::max(1, n); g++) a.gpus.push_back(g); }
        else if (s == "--gpu-list") { a.gpus.clear(); std::stringstream ss(val()); std::string tok; while (std::getline(ss, tok, ',')) a.gpus.push_back(atoi(tok.c_str())); }
        else if (s == "--lanes") { a.lanes = atoi(val()); if (a.lanes < 1 || a.lanes > 8) die("--lanes must be in 1..8"); }
        else if (s == "--threads") a.threads = atoi(val());
        else if (s == "--gpu-server") { a.gpuServer = atoi(val()); if (a.gpuServer != 0 && a.gpuServer != 1) die("invalid --gpu-server"); }
        else if (s == "--gpu-server-wait-timeout") a.gpuServerWait = atoi(val());
        else if (s == "-v") a.verbosity = atoi(val());
        else if (s == "-a") a.addBacktrace = atoi(val());
        else if (s == "--sort-by-structure-bits") a.sortByStructureBits = atoi(val());
        else if (s == "-c") a.covThr = (float) atof(val());
        else if (s == "--cov-mode") { a.covMode = atoi(val()); if (a.covMode < 0 || a.covMode > 5) die("invalid --cov-mode"); }
        else if (s == "--min-seq-id") a.seqIdThr = (float) atof(val());
        else if (s == "--min-aln-len") a.alnLenThr = atoi(val());
        else if (s == "--max-accept") a.maxAccept = atoi(val());
        else if (s == "--max-rejected") a.maxRejected = atoi(val());
        else if (s == "--tmscore-threshold") a.tmScoreThr = (float) atof(val());
        else if (s == "--lddt-threshold") a.lddtThr = (float) atof(val());
        else if (s == "--mask-lower-case") a.maskLowerCase = atoi(val());
        else if (s == "--add-self-matches") a.includeIdentity = atoi(val());
        // parameters without influence on the result or on this implementation
        else if (s == "--db-load-mode" || s == "--local-tmp" || s == "--split-memory-limit" || s == "--pca" || s == "--pcb" ||
                 s == "--max-seq-len" || s == "--zdrop" || s == "--realign-score-bias" || s == "--realign-max-seqs" || s == "--mask-prob" ||
                 s == "--corr-score-weight") (void) val();
        // parameters accepted only with the value that matches what is implemented
        else if (s == "--sub-mat" || s == "--seed-sub-mat") only({"b64:YWE6M2RpLm91dCxudWNsOjNkaS5vdXQ=", "3di.out", "aa:3di.out,nucl:3di.out"});
        else if (s == "--target-search-mode") only({"0"});
        else if (s == "--k-score") only({"2147483647"});
        else if (s == "--alph-size") only({"21"});
        else if (s == "--split") { a.split = atoi(val()); if (a.split < 0) die("invalid --split"); if (a.split == 0) a.split = 1; }   // 0 = automatic: the DB fits
        else if (s == "--split-mode") { a.splitMode = atoi(val()); if (a.splitMode < 0 || a.splitMode > 2) die("invalid --split-mode"); }
        else if (s == "--diag-score") only({"1"});
        else if (s == "--exact-kmer-matching") only({"0"});
        else if (s == "--mask") only({"0"});
        else if (s == "--mask-n-repeat") only({"6"});
        else if (s == "--spaced-kmer-mode") only({"1"});
        else if (s == "--spaced-kmer-pattern") only({""});
        else if (s == "--taxon-list") only({""});
        else if (s == "--compressed") { a.compressed = atoi(val()); if (a.compressed != 0 && a.compressed != 1) die("invalid --compressed"); }
        else if (s == "--alignment-mode") { a.alignmentMode = atoi(val()); if (a.alignmentMode < 0 || a.alignmentMode > 3) die("invalid --alignment-mode"); }
        else if (s == "--alignment-output-mode") only({"0"});
        else if (s == "--wrapped-scoring") only({"0"});
        else if (s == "--seq-id-mode") { a.seqIdMode = atoi(val()); if (a.seqIdMode < 0 || a.seqIdMode > 2) die("invalid --seq-id-mode"); }
        // parameters of the search workflow itself (src/commons/LocalParameters.cpp:293-304): the single pass over both stages is what is built
        else if (s == "--num-iterations") only({"1"});
        else if (s == "--exhaustive-search") only({"0"});
        else if (s == "--cluster-search") only({"0"});
        else if (s == "--prefilter-mode") only({"0"});
        else if (s == "--remove-tmp-files" || s == "--force-reuse") (void) val();
        else if (s == "--mpi-runner") only({""});
        else if (s == "--alt-ali") { a.altAli = atoi(val()); if (a.altAli < 0) die("invalid --alt-ali"); }
        else if (s == "--score-bias") only({"0"});
        else if (s == "--realign") only({"0"});
        else if (s == "--tmscore-threshold-mode") { a.cfg.tmscore_norm_mode = atoi(val()); if (a.cfg.tmscore_norm_mode < 0 || a.cfg.tmscore_norm_mode > 2) die("invalid --tmscore-threshold-mode"); }
        else if (s == "--exact-tmscore") only({"0"});
        else if (!s.empty() && s[0] == '-' && s.size() > 1 && !isdigit((unsigned char) s[1])) die("unrecognized parameter " + s);
        else a.pos.push_back(s);
    }
    if (a.gpus.empty()) a.gpus.push_back(a.cfg.device);
    a.cfg.device = a.gpus[0];
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

struct Resources { std::string m3, mA, nn; };

fsb200_ctx *make_ctx(Args &a, const char *argv0, Resources &r, int device) {
    if (r.m3.empty()) {
        const std::string d = data_dir(argv0);
        r.m3 = slurp(d + "/mat3di.out"); r.mA = slurp(d + "/blosum62.out"); r.nn = slurp(d + "/evalue_nn.kerasify");
    }
    fsb200_config cfg = a.cfg;
    cfg.device = device;
    cfg.mat3di_text = r.m3.c_str(); cfg.mataa_text = r.mA.c_str();
    cfg.evalue_nn = reinterpret_cast<const uint8_t *>(r.nn.data()); cfg.evalue_nn_len = r.nn.size();
    fsb200_ctx *ctx = NULL;
    const int rc = fsb200_create(&cfg, &ctx);
    if (rc != FSB200_OK) die("fsb200_create failed with " + std::to_string(rc) + " on device " + std::to_string(device) + " (a CUDA device is required, there is no CPU path)");
    return ctx;
}

void check(fsb200_ctx *ctx, int rc, const char *what) {
    if (rc != FSB200_OK) die(std::string(what) + " failed (" + std::to_string(rc) + "): " + fsb200_last_error(ctx));
}

// One context per GPU of --gpu / --gpu-list: the first stages the padded target DB from the host, the others receive the staged
// arrays device-to-device (NVLink peer copy); queries are split in contiguous ranges balanced by `weight` (the reference's
// query split, Prefiltering.cpp:781-786) and every range is processed by its own host thread; one writer at the end.
// off[i] / len[i] (i < n) locate the targets inside the code arrays taa / tss
std::vector<fsb200_ctx *> make_contexts(Args &a, const char *argv0, Resources &r, const uint8_t *taa, const uint8_t *tss,
                                        const uint64_t *off, const uint32_t *len, size_t n, uint64_t aaRes) {
    std::vector<fsb200_ctx *> ctxs;
    try {
        for (size_t g = 0; g < a.gpus.size(); g++) {
            fsb200_ctx *ctx = make_ctx(a, argv0, r, a.gpus[g]);
            ctxs.push_back(ctx);
            if (g == 0) check(ctx, fsb200_load_target(ctx, taa, tss, off, len, (uint32_t) n, aaRes), "load_target");
            else {
                void *dA = NULL, *dS = NULL; uint64_t bytes = 0;
                std::vector<uint64_t> so(n); std::vector<uint32_t> sl(n);
                check(ctxs[0], fsb200_staged_target_size(ctxs[0], &bytes), "staged_target_size");
                check(ctxs[0], fsb200_export_staged_target(ctxs[0], &dA, &dS, so.data(), sl.data()), "export_staged_target");
                check(ctx, fsb200_load_target_device(ctx, dA, dS, bytes, so.data(), sl.data(), (uint32_t) n, aaRes), "load_target_device");
            }
        }
    } catch (...) {
        for (size_t i = ctxs.size(); i-- > 0;) fsb200_destroy(ctxs[i]);
        throw;
    }
    return ctxs;
}

// CPUs this process may really use: affinity mask and cgroup quota (a container often shows all cores of the box but grants a few)
int usable_cpus() {
    int n = (int) std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n, CPU_COUNT(&set));
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[64]; long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) n = std::min(n, (int) ((atol(q) + period / 2) / period));
        fclose(f);
    }
    return std::max(1, n);
}

// --lanes: every GPU context gets lanes - 1 more contexts that share its staged target, C-alpha and k-mer index (fsb200_create_lane).
// One host thread drives each context, so one lane's kernels run while another lane is in a host-side part of its batch.  Called once
// the parents hold everything a lane shares; the lanes are appended behind the parents (destroy in reverse order).
void add_lanes(const Args &a, std::vector<fsb200_ctx *> &ctxs, size_t units) {
    const size_t parents = ctxs.size();
    const size_t pp = std::max<size_t>(1, parents);
    const size_t lanes = std::max<size_t>(1, std::min<size_t>((size_t) a.lanes, units / pp + (units % pp != 0)));   // units may be SIZE_MAX (resident state)
    for (size_t l = 1; l < lanes; l++)
        for (size_t g = 0; g < parents; g++) {
            fsb200_ctx *lane = NULL;
            check(ctxs[g], fsb200_create_lane(ctxs[g], &lane), "create_lane");
            ctxs.push_back(lane);
        }
}

void set_threads(const Args &a, std::vector<fsb200_ctx *> &ctxs) {
    int hw = a.threads > 0 ? a.threads : usable_cpus();
    const int per = std::max(1, hw / (int) std::max<size_t>(1, ctxs.size()));
    for (fsb200_ctx *c : ctxs) fsb200_set_host_threads(c, per);
}

void destroy_contexts(std::vector<fsb200_ctx *> &ctxs) {
    for (size_t i = ctxs.size(); i-- > 0;) fsb200_destroy(ctxs[i]);   // lanes before their parents
    ctxs.clear();
}

std::vector<size_t> split_by_weight(const std::vector<uint64_t> &w, size_t parts) {
    std::vector<size_t> b(parts + 1, w.size());
    uint64_t total = 0; for (uint64_t x : w) total += x;
    b[0] = 0;
    uint64_t run = 0; size_t i = 0;
    for (size_t p = 1; p < parts; p++) {
        const uint64_t target = total * p / parts;
        while (i < w.size() && run < target) run += w[i++];
        b[p] = i;
    }
    return b;
}

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

// ---- what a module keeps of its target DB: built per run, or kept by the gpuserver between the calls it serves ----
struct PrefTarget {      // prefilter: masked 3Di sequences staged + k-mer index built on every GPU, lanes created
    std::string key;
    std::vector<uint32_t> tkey, tl;
    std::vector<fsb200_ctx *> ctxs;
    uint64_t served = 0;
    bool owned = true;       // false: a view on the contexts of an AlnTarget (search: both stages on one staged DB)
    ~PrefTarget() { if (owned) destroy_contexts(ctxs); }
};
struct AlnTarget {       // structurealign: AA + 3Di sequences (and C-alpha coordinates once a call needs them) staged on every GPU
    std::string key, path;
    std::string prefKey;          // with hasIndex: the prefilter parameters the contexts were created with
    std::vector<uint32_t> tkey, tl; std::vector<uint64_t> to;
    std::vector<uint8_t> tca, tcs;          // unmasked codes: --alt-ali masks copies of them
    std::vector<uint32_t> keyToT;
    uint64_t aaRes = 0;
    bool hasCa = false;
    bool hasIndex = false;        // k-mer index built on the parents before the lanes were created (search runs the prefilter on these contexts)
    bool masked3Di = false;       // the 3Di DB has lower-case (masked) residues: the prefilter of --mask-lower-case 1 needs its own staged copy
    std::vector<fsb200_ctx *> ctxs;
    uint64_t served = 0;
    ~AlnTarget() { destroy_contexts(ctxs); }
};
struct Resident {
    std::shared_ptr<PrefTarget> pref; std::shared_ptr<AlnTarget> aln;
    std::vector<int> gpus; int lanes = 1, threads = 0;   // the server's own --gpu-list / --lanes / --threads: the GPUs are the server's, not the caller's
    bool touched = false;                                // the running call has started work on the resident contexts
};
Resident *g_resident = NULL;   // non-NULL inside the gpuserver
void use_server_layout(Args &a) {
    if (g_resident == NULL || g_resident->gpus.empty()) return;
    a.gpus = g_resident->gpus; a.lanes = g_resident->lanes; a.cfg.device = a.gpus[0];
    if (g_resident->threads > 0) a.threads = g_resident->threads;
}

std::string file_stamp(const std::string &path) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return path + ":-";
    return path + ":" + std::to_string((unsigned long long) st.st_size) + ":" + std::to_string((long long) st.st_mtime) + "." + std::to_string((long) st.st_mtim.tv_nsec);
}
std::string real_path(const std::string &path) {
    char buf[PATH_MAX];
    return realpath(path.c_str(), buf) ? std::string(buf) : path;
}
std::string ctx_layout(const Args &a) {
    std::string k = "gpus";
    for (int g : a.gpus) k += "," + std::to_string(g);
    return k + "|lanes" + std::to_string(a.lanes);
}

// contexts with the targets [from, from + n) of an encoded DB staged and indexed
std::shared_ptr<PrefTarget> build_pref_target(Args &a, const char *argv0, const std::vector<uint32_t> &keys, const std::vector<uint8_t> &tc,
                                              const std::vector<uint64_t> &to, const std::vector<uint32_t> &tl, size_t from, size_t n, size_t units) {
    if (n >= (1u << 22))   // one k-mer index holds fewer than 4M targets (22-bit target ids in the hit stream)
        die("target range of " + std::to_string(n) + " sequences: more than 4194303 per index need a target split, --split " +
            std::to_string(n / ((1u << 22) - 1) + 1) + " --split-mode 0 (or more)");
    std::shared_ptr<PrefTarget> T(new PrefTarget());
    Resources res;
    T->tkey.assign(keys.begin() + from, keys.begin() + from + n);
    T->tl.assign(tl.begin() + from, tl.begin() + from + n);
    uint64_t resid = 0; for (uint32_t l : T->tl) resid += l;
    T->ctxs = make_contexts(a, argv0, res, tc.data(), tc.data(), to.data() + from, tl.data() + from, n, resid);
    for (fsb200_ctx *c : T->ctxs) check(c, fsb200_build_index(c), "build_index");   // before the lanes: they share it
    add_lanes(a, T->ctxs, units);
    set_threads(a, T->ctxs);
    return T;
}

// the parameters a context is created with that matter to the prefilter stage
std::string pref_params(const Args &a) {
    char num[256];
    snprintf(num, sizeof(num), "|s%.9g|k%d|ms%d|mds%d|cb%d|cbs%.9g|bins%d|mlc%d|", a.cfg.sensitivity, a.cfg.kmer_size, a.cfg.max_seqs, a.cfg.min_diag_score,
             a.cfg.comp_bias_corr, a.cfg.comp_bias_scale_prefilter, a.cfg.diag_bin_count, a.maskLowerCase);
    return num;
}

std::shared_ptr<PrefTarget> pref_target(Args &a, const char *argv0, const std::string &path, size_t units) {
    const std::string key = "prefilter|" + file_stamp(real_path(path)) + "|" + file_stamp(real_path(path + ".index")) + pref_params(a) + ctx_layout(a);
    if (g_resident && g_resident->pref && g_resident->pref->key == key) { set_threads(a, g_resident->pref->ctxs); return g_resident->pref; }
    if (g_resident) g_resident->pref.reset();   // frees the device memory of the state that is being replaced
    Db t; t.open(path);
    std::vector<uint8_t> tc; std::vector<uint64_t> to; std::vector<uint32_t> tl;
    encode(t, a.maskLowerCase != 0, tc, to, tl);
    std::shared_ptr<PrefTarget> T = build_pref_target(a, argv0, t.key, tc, to, tl, 0, tl.size(), units);
    T->key = key;
    if (g_resident) g_resident->pref = T;
    return T;
}

void load_target_ca(AlnTarget &T) {
    if (T.hasCa) return;
    Db tCa; tCa.open(T.path + "_ca");
    std::vector<uint32_t> keyToCa(tCa.size() ? *std::max_element(tCa.key.begin(), tCa.key.end()) + 1 : 0, 0xFFFFFFFFu);
    for (size_t i = 0; i < tCa.size(); i++) keyToCa[tCa.key[i]] = (uint32_t) i;
    std::vector<float> x, y, z; std::vector<uint64_t> co(T.tkey.size());
    for (size_t i = 0; i < T.tkey.size(); i++) {
        if (T.tkey[i] >= keyToCa.size() || keyToCa[T.tkey[i]] == 0xFFFFFFFFu) die("no C-alpha entry for a target key");
        co[i] = x.size();
        decode_ca(tCa, keyToCa[T.tkey[i]], T.tl[i], x, y, z);
    }
    for (fsb200_ctx *c : T.ctxs) check(c, fsb200_load_target_ca(c, x.data(), y.data(), z.data(), co.data()), "load_target_ca");
    T.hasCa = true;
}

std::shared_ptr<AlnTarget> aln_target(Args &a, const char *argv0, const std::string &path, size_t units, bool needCa, bool withIndex = false) {
    char num[256];
    snprintf(num, sizeof(num), "|go%d|ge%d|cb%d|cbs%.9g|e%.17g|aaf%.9g|tmn%d|", a.cfg.gap_open, a.cfg.gap_extend, a.cfg.comp_bias_corr, a.cfg.comp_bias_scale_align,
             a.cfg.evalue_thr, a.cfg.aa_factor, a.cfg.tmscore_norm_mode);
    const std::string rp = real_path(path);
    const std::string key = "structurealign|" + file_stamp(rp) + "|" + file_stamp(rp + ".index") + "|" + file_stamp(rp + "_ss") + "|" + file_stamp(rp + "_ca") + num + ctx_layout(a);
    std::shared_ptr<AlnTarget> T;
    // a resident state serves the call if it was created with the same alignment parameters and, where the call runs the prefilter on
    // it (search), holds an index and was created with the same prefilter parameters
    if (g_resident && g_resident->aln && g_resident->aln->key == key && (!withIndex || !g_resident->aln->hasIndex || g_resident->aln->prefKey == pref_params(a)))
        T = g_resident->aln;
    else {
        if (g_resident) g_resident->aln.reset();
        T.reset(new AlnTarget());
        T->key = key; T->path = path;
        Db ta, ts; ta.open(path); ts.open(path + "_ss");
        Resources res;
        std::vector<uint64_t> o2; std::vector<uint32_t> l2;
        encode(ta, false, T->tca, T->to, T->tl); encode(ts, false, T->tcs, o2, l2);
        if (l2 != T->tl) die("target AA and 3Di databases differ in sequence lengths");
        for (size_t i = 0; i < ts.size() && !T->masked3Di; i++) {
            const char *sq = ts.data.data() + ts.off[i];
            for (uint32_t k = 0; k < T->tl[i]; k++) if (ts.padded ? (unsigned char) sq[k] >= 32 : islower((unsigned char) sq[k]) != 0) { T->masked3Di = true; break; }
        }
        T->tkey = ta.key;
        // DBReader::getAminoAcidDBSize() = data size - 2 * entries (lib/mmseqs/src/commons/DBReader.cpp:403-412)
        T->aaRes = ta.fileBytes - 2 * ta.size();
        T->ctxs = make_contexts(a, argv0, res, T->tca.data(), T->tcs.data(), T->to.data(), T->tl.data(), T->tl.size(), T->aaRes);
        T->keyToT.assign(ta.size() ? *std::max_element(ta.key.begin(), ta.key.end()) + 1 : 0, 0xFFFFFFFFu);
        for (size_t i = 0; i < ta.size(); i++) T->keyToT[ta.key[i]] = (uint32_t) i;
        if (needCa) load_target_ca(*T);   // before the lanes: they share the coordinates the parent holds when they are created
        if (withIndex) { for (fsb200_ctx *c : T->ctxs) check(c, fsb200_build_index(c), "build_index"); T->hasIndex = true; T->prefKey = pref_params(a); }   // likewise
        add_lanes(a, T->ctxs, units);
        if (g_resident) g_resident->aln = T;
    }
    // a resident state staged by a call that did not need the coordinates or the index: lanes take over what their parent holds when
    // they are created, and an index must be built by contexts created with this call's prefilter parameters - restage in both cases
    if ((withIndex && !T->hasIndex) || (needCa && !T->hasCa && T->ctxs.size() > a.gpus.size())) {
        const bool ca = needCa || T->hasCa, ix = withIndex || T->hasIndex;
        if (g_resident) g_resident->aln.reset();
        T.reset();
        return aln_target(a, argv0, path, units, ca, ix);
    }
    if (needCa) load_target_ca(*T);
    set_threads(a, T->ctxs);
    return T;
}

int forward_to_server(int argc, char **argv, const std::string &targetPath, const Args &a);

// DBReader::decomposeDomainByAminoAcid (lib/mmseqs/src/commons/DBReader.cpp:1108-1150): entry ranges of `parts` chunks of about equal bytes
std::vector<std::pair<size_t, size_t>> decompose_by_residues(const Db &db, size_t parts) {
    std::vector<std::pair<size_t, size_t>> r(parts, {0, 0});
    const size_t entries = db.size();
    if (parts <= 1) { r.assign(1, {0, entries}); return r; }
    if (entries <= parts) { for (size_t p = 0; p < parts; p++) r[p] = {p < entries ? p : 0, p < entries ? 1 : 0}; return r; }
    const size_t chunk = (size_t) ceil((double) db.fileBytes / (double) parts);
    std::vector<size_t> per(parts, 0);
    size_t rank = 0, sum = 0;
    for (size_t i = 0; i < entries; i++) {
        if (sum >= chunk) { sum = 0; rank++; }
        sum += db.len[i];
        per[rank]++;
    }
    size_t start = 0;
    for (size_t p = 0; p < parts; p++) { r[p] = {start, per[p]}; start += per[p]; }
    return r;
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
    auto pass = [&](PrefTarget &T, const std::vector<uint32_t> &ident, int maxSeqs, size_t dbFrom) {
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
                lists[i].push_back(PH{T.tkey[h.target], (uint32_t) (dbFrom + h.target), h.score, h.diagonal});
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
        pass(*onContexts, identity(onContexts->tkey), a.cfg.max_seqs, 0);
    } else if (!targetSplit) {
        std::shared_ptr<PrefTarget> T = pref_target(a, argv0, tPath, g_resident ? (size_t) -1 : ql.size());
        pass(*T, identity(T->tkey), a.cfg.max_seqs, 0);
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
        for (size_t sp = 0; sp < ranges.size(); sp++) {
            if (ranges[sp].second == 0) continue;
            if (a.verbosity >= 3) say_out("Process prefiltering step " + std::to_string(sp + 1) + " of " + std::to_string(a.split) + "\nTarget db start " +
                                          std::to_string(ranges[sp].first + 1) + " to " + std::to_string(ranges[sp].first + ranges[sp].second) + "\n");
            std::shared_ptr<PrefTarget> T = build_pref_target(as, argv0, t.key, tc, to, tl, ranges[sp].first, ranges[sp].second, ql.size());
            pass(*T, identity(T->tkey), maxSeqs, ranges[sp].first);
        }
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
    const std::vector<uint32_t> &tl = T->tl, &keyToT = T->keyToT;
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
    run_workers(ctxs.size(), [&](size_t g) {
        fsb200_ctx *ctx = ctxs[g];
        const size_t pb = bounds[g], pe = bounds[g + 1];
        if (pe <= pb) return;
        // one query batch = the queries of this shard's prefilter entries, in prefilter-DB order
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
    const std::vector<std::string> out = structurealign_core(a, argv[0], a.pos[0], a.pos[1], (size_t) -1, !targetSplit, [&](AlnTarget &T, const Db &qa, AlignInput &x) {
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

// ---- gpuserver: resident target DB, module calls over a Unix-domain socket ----
// GPUSharedMemory::getShmHash (lib/mmseqs/src/commons/GpuUtil.cpp:18-34): real path of the DB + CUDA_VISIBLE_DEVICES + version
std::string server_socket_path(const std::string &targetDb) {
    std::string id = real_path(targetDb);
    if (const char *v = getenv("CUDA_VISIBLE_DEVICES")) id += v;
    id += fsb200_build_id();
    uint64_t h = 1469598103934665603ull;   // FNV-1a
    for (unsigned char c : id) { h ^= c; h *= 1099511628211ull; }
    const char *dir = getenv("FSB200_SOCKET_DIR");
    char name[64]; snprintf(name, sizeof(name), "/fsb200-gpuserver-%016llx.sock", (unsigned long long) h);
    return std::string(dir ? dir : "/tmp") + name;
}

bool send_all(int fd, const void *buf, size_t n) {
    const char *p = static_cast<const char *>(buf);
    while (n > 0) { const ssize_t w = send(fd, p, n, MSG_NOSIGNAL); if (w <= 0) return false; p += w; n -= (size_t) w; }
    return true;
}
bool recv_all(int fd, void *buf, size_t n) {
    char *p = static_cast<char *>(buf);
    while (n > 0) { const ssize_t r = recv(fd, p, n, 0); if (r <= 0) return false; p += r; n -= (size_t) r; }
    return true;
}
bool send_str(int fd, const std::string &t) { const uint32_t n = (uint32_t) t.size(); return send_all(fd, &n, 4) && send_all(fd, t.data(), n); }
bool recv_str(int fd, std::string &t) {
    uint32_t n = 0;
    if (!recv_all(fd, &n, 4) || n > (1u << 28)) return false;
    t.resize(n);
    return n == 0 || recv_all(fd, &t[0], n);
}

int connect_to(const std::string &path) {
    const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    sockaddr_un addr; memset(&addr, 0, sizeof(addr)); addr.sun_family = AF_UNIX;
    if (path.size() >= sizeof(addr.sun_path)) { close(fd); die("socket path too long: " + path); }
    strcpy(addr.sun_path, path.c_str());
    if (connect(fd, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0) { close(fd); return -1; }
    return fd;
}

// client side of --gpu-server 1 (ungappedprefilter.cpp:72-124: wait for the server of the target DB, then use it).  The prefilter's
// target is <targetDB>_ss while the server is started on <targetDB>: both names lead to the same server.
int forward_to_server(int argc, char **argv, const std::string &targetPath, const Args &a) {
    std::vector<std::string> names(1, targetPath);
    if (targetPath.size() > 3 && targetPath.compare(targetPath.size() - 3, 3, "_ss") == 0) names.push_back(targetPath.substr(0, targetPath.size() - 3));
    int fd = -1;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool told = false;
    for (;;) {
        for (const std::string &n : names) { fd = connect_to(server_socket_path(n)); if (fd >= 0) break; }
        if (fd >= 0) break;
        const long elapsed = (long) std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
        if (a.gpuServerWait == 0 || (a.gpuServerWait > 0 && elapsed >= a.gpuServerWait))
            die("gpuserver for database " + targetPath + " not found" + (a.gpuServerWait > 0 ? " after " + std::to_string(elapsed) + " seconds" : std::string()) +
                ".\nPlease start `foldseek-b200 gpuserver` on it with the same CUDA_VISIBLE_DEVICES");
        if (!told) { say_err("Waiting for `gpuserver`\n"); told = true; }
        std::this_thread::sleep_for(std::chrono::milliseconds(200));
    }
    char cwd[PATH_MAX];
    if (getcwd(cwd, sizeof(cwd)) == NULL) die("getcwd failed");
    bool ok = send_str(fd, cwd);
    const uint32_t n = (uint32_t) argc;
    ok = ok && send_all(fd, &n, 4);
    for (int i = 0; ok && i < argc; i++) ok = send_str(fd, argv[i]);
    int32_t code = EXIT_FAILURE; std::string err, out;
    ok = ok && recv_all(fd, &code, 4) && recv_str(fd, err) && recv_str(fd, out);
    close(fd);
    if (!ok) die("connection to the gpuserver for database " + targetPath + " broke");
    fputs(out.c_str(), stdout); fputs(err.c_str(), stderr);
    return code;
}

volatile sig_atomic_t g_keepRunning = 1;
void stop_handler(int) { g_keepRunning = 0; }

int run_module(int argc, char **argv);

int gpuserver_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 1) die("usage: foldseek-b200 gpuserver <i:targetDB> [options]   (serves prefilter / structurealign calls that carry --gpu-server 1)");
    const std::string sock = server_socket_path(a.pos[0]);
    { const int probe = connect_to(sock); if (probe >= 0) { close(probe); die("a gpuserver for database " + a.pos[0] + " is already running (" + sock + ")"); } }
    unlink(sock.c_str());   // left behind by a server that was killed
    Resident resident;
    resident.gpus = a.gpus; resident.lanes = a.lanes; resident.threads = a.threads;
    g_resident = &resident;
    // Preload with the parameters of this command line (defaults where none are given): a later call whose parameters differ in a way
    // that matters for the staged state replaces the state (the staging is then paid once more).  --comp-bias-corr-scale is per module
    // in the workflow (0.15 prefilter, 0.5 structurealign, StructureSearch.cpp:101-107), the defaults of fsb200_config.
    const bool haveSs = !slurp(a.pos[0] + "_ss.dbtype", false).empty();
    const bool haveCa = !slurp(a.pos[0] + "_ca.dbtype", false).empty();
    if (haveSs) {
        pref_target(a, argv[0], a.pos[0] + "_ss", (size_t) -1);
        aln_target(a, argv[0], a.pos[0], (size_t) -1, haveCa, true);   // with the index: serves structurealign and search calls
    } else say_err("gpuserver: " + a.pos[0] + "_ss not found, nothing preloaded\n");
    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) die("socket() failed");
    sockaddr_un addr; memset(&addr, 0, sizeof(addr)); addr.sun_family = AF_UNIX;
    if (sock.size() >= sizeof(addr.sun_path)) die("socket path too long: " + sock);
    strcpy(addr.sun_path, sock.c_str());
    if (bind(ls, reinterpret_cast<sockaddr *>(&addr), sizeof(addr)) != 0 || listen(ls, 16) != 0) die("cannot listen on " + sock);
    struct sigaction act; memset(&act, 0, sizeof(act)); act.sa_handler = stop_handler;
    sigaction(SIGINT, &act, NULL); sigaction(SIGTERM, &act, NULL);
    say_out(sock + "\n");   // the reference prints the name of its shared-memory block once the DB is loaded (gpuserver.cpp:81)
    while (g_keepRunning) {
        pollfd pf{ls, POLLIN, 0};
        if (poll(&pf, 1, 200) <= 0) continue;
        const int fd = accept(ls, NULL, NULL);
        if (fd < 0) continue;
        std::string cwd; uint32_t n = 0; std::vector<std::string> args;
        bool ok = recv_str(fd, cwd) && recv_all(fd, &n, 4) && n >= 2 && n < 4096;
        for (uint32_t i = 0; ok && i < n; i++) { std::string t; ok = recv_str(fd, t); args.push_back(t); }
        if (!ok) { close(fd); continue; }
        int32_t code = EXIT_FAILURE;
        { std::lock_guard<std::mutex> g(g_console.m); g_console.capture = true; g_console.out.clear(); g_console.err.clear(); }
        resident.touched = false;
        try {
            if (chdir(cwd.c_str()) != 0) die("gpuserver: cannot change to " + cwd);
            if (args[1] != "prefilter" && args[1] != "structurealign" && args[1] != "search") die("gpuserver: module " + args[1] + " is not served");
            std::vector<char *> av; for (std::string &t : args) av.push_back(&t[0]);
            av[0] = argv[0];
            code = run_module((int) av.size(), av.data());
        } catch (const std::exception &e) {
            say_err(std::string("Error: ") + e.what() + "\n");
            code = EXIT_FAILURE;
            if (resident.touched) { resident.pref.reset(); resident.aln.reset(); }   // a context may be mid-batch: restage at the next call
        }
        std::string err, out;
        { std::lock_guard<std::mutex> g(g_console.m); g_console.capture = false; err.swap(g_console.err); out.swap(g_console.out); }
        if (!(send_all(fd, &code, 4) && send_str(fd, err) && send_str(fd, out))) say_err("gpuserver: client went away\n");
        close(fd);
    }
    close(ls);
    unlink(sock.c_str());
    g_resident = NULL;
    return EXIT_SUCCESS;
}

int run_module(int argc, char **argv) {
    if (strcmp(argv[1], "search") == 0) return search_main(argc, argv);
    if (strcmp(argv[1], "prefilter") == 0) return prefilter_main(argc, argv);
    if (strcmp(argv[1], "structurealign") == 0) return structurealign_main(argc, argv);
    if (strcmp(argv[1], "gpuserver") == 0) return gpuserver_main(argc, argv);
    die(std::string("unknown module ") + argv[1]);
}

}  // namespace

int main(int argc, char **argv) {
    try {
        if (argc < 2) die("usage: foldseek-b200 <search|prefilter|structurealign|gpuserver> ...");
        return run_module(argc, argv);
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return EXIT_FAILURE;
    }
}
