// cli_contexts.h - contexts per GPU and lane, and what a module keeps of its target DB (per run, or resident inside the gpuserver)
// Part of the host module (foldseek-b200), included by cli_main.cpp inside its anonymous namespace.
#ifndef CLI_CONTEXTS_H
#define CLI_CONTEXTS_H

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
    std::vector<uint64_t> co(T.tkey.size());
    uint64_t total = 0;
    for (size_t i = 0; i < T.tkey.size(); i++) {
        if (T.tkey[i] >= keyToCa.size() || keyToCa[T.tkey[i]] == 0xFFFFFFFFu) die("no C-alpha entry for a target key");
        co[i] = total; total += T.tl[i];
    }
    std::vector<float> x(total), y(total), z(total);
    const int64_t n = (int64_t) T.tkey.size();
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) decode_ca_into(tCa, keyToCa[T.tkey[i]], T.tl[i], x.data() + co[i], y.data() + co[i], z.data() + co[i]);
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
        bool masked = false;
        const int64_t nts = (int64_t) ts.size();
#pragma omp parallel for schedule(static, 4096) reduction(|| : masked)
        for (int64_t i = 0; i < nts; i++) {
            const unsigned char *sq = reinterpret_cast<const unsigned char *>(ts.data.data()) + ts.off[i];
            bool m = false;
            for (uint32_t k = 0; k < T->tl[i]; k++) m = m || (ts.padded ? sq[k] >= 32 : (sq[k] >= 'a' && sq[k] <= 'z'));
            masked = masked || m;
        }
        T->masked3Di = masked;
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

#endif
