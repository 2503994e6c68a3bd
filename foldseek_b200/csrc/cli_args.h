// cli_args.h - the parameters of the modules (every flag the workflow serialises is known; nothing is swallowed)
// Part of the host module (foldseek-b200), included by cli_main.cpp inside its anonymous namespace.
#ifndef CLI_ARGS_H
#define CLI_ARGS_H

struct Args {
    std::vector<std::string> pos;
    std::vector<int> posIdx;                        // argv index of every positional
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
    int exhaustive = 0;                             // --exhaustive-search 1 (search): no prefilter, every target is a candidate of every query
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
    bool lanesSet = false;
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
        else if (s == "--lanes") { a.lanes = atoi(val()); lanesSet = true; if (a.lanes < 1 || a.lanes > 8) die("--lanes must be in 1..8"); }
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
        // MMseqs2 align parameters that structurealign never reads (src/strucclustutils/structurealign.cpp uses none of par.realign,
        // par.wrappedScoring, par.alignmentOutputMode): any value gives the reference's result, checked against refseek
        else if (s == "--alignment-output-mode" || s == "--wrapped-scoring" || s == "--realign") (void) val();
        else if (s == "--seq-id-mode") { a.seqIdMode = atoi(val()); if (a.seqIdMode < 0 || a.seqIdMode > 2) die("invalid --seq-id-mode"); }
        // parameters of the search workflow itself (src/commons/LocalParameters.cpp:293-304): the single pass over both stages is what is built
        else if (s == "--num-iterations") only({"1"});
        else if (s == "--exhaustive-search") { a.exhaustive = atoi(val()); if (a.exhaustive != 0 && a.exhaustive != 1) die("invalid --exhaustive-search"); }
        else if (s == "--cluster-search") only({"0"});
        else if (s == "--prefilter-mode") only({"0"});
        else if (s == "--remove-tmp-files" || s == "--force-reuse") (void) val();
        else if (s == "--mpi-runner") only({""});
        // convertalis (lib/mmseqs/src/commons/Parameters.cpp:661-672): the default BLAST-tab output is what is built
        else if (s == "--format-mode") only({"0"});
        else if (s == "--format-output") only({"query,target,fident,alnlen,mismatch,gapopen,qstart,qend,tstart,tend,evalue,bits"});
        else if (s == "--db-output") only({"0"});
        else if (s == "--translation-table" || s == "--search-type") (void) val();
        else if (s == "--alt-ali") { a.altAli = atoi(val()); if (a.altAli < 0) die("invalid --alt-ali"); }
        else if (s == "--score-bias") only({"0"});
        else if (s == "--tmscore-threshold-mode") { a.cfg.tmscore_norm_mode = atoi(val()); if (a.cfg.tmscore_norm_mode < 0 || a.cfg.tmscore_norm_mode > 2) die("invalid --tmscore-threshold-mode"); }
        else if (s == "--exact-tmscore") only({"0"});
        else if (!s.empty() && s[0] == '-' && s.size() > 1 && !isdigit((unsigned char) s[1])) die("unrecognized parameter " + s);
        else { a.pos.push_back(s); a.posIdx.push_back(i); }
    }
    // a target given as "<db>.idx" (the workflow appends ${INDEXEXT} when `createindex` was run, data/structuresearch.sh:50,121) means the
    // DB it was built from (PrefilteringIndexReader::dbPathWithoutIndex, PrefilteringIndexReader.cpp:613-626): the k-mer index is built on
    // the GPU in any case, nothing of the precomputed one is read
    for (std::string &p : a.pos) {
        for (const char *ext : {".idx", ".linidx"}) {
            const size_t n = strlen(ext);
            if (p.size() > n && p.compare(p.size() - n, n, ext) == 0) p.resize(p.size() - n);
        }
    }
    // GPUs and lanes for module calls that come from the reference's workflow scripts, whose parameter strings cannot carry this tool's own
    // flags (the prefilter / structurealign lists have no --gpu): FSB200_GPUS=0,1,2,3  FSB200_LANES=2
    if (a.gpus.empty()) if (const char *e = getenv("FSB200_GPUS")) { std::stringstream ss(e); std::string tok; while (std::getline(ss, tok, ',')) if (!tok.empty()) a.gpus.push_back(atoi(tok.c_str())); }
    if (!lanesSet) if (const char *e = getenv("FSB200_LANES")) { a.lanes = atoi(e); if (a.lanes < 1 || a.lanes > 8) die("FSB200_LANES must be in 1..8"); }
    if (a.gpus.empty()) a.gpus.push_back(a.cfg.device);
    a.cfg.device = a.gpus[0];
    return a;
}

#endif
