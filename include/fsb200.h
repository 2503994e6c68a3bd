/* fsb200.h — C ABI of the B200-native Foldseek search hot path (libfsb200.so).
 *
 * The two reference modules this library sits under are invoked as separate processes with on-disk MMseqs2
 * databases (data/structuresearch.sh:50,121):
 *     prefilter      <queryDB_ss> <targetDB_ss> <prefDB>          lib/mmseqs/src/prefiltering/Main.cpp:13
 *     structurealign <queryDB> <targetDB> <prefDB> <alnDB>        src/strucclustutils/structurealign.cpp:141
 * After parameter parsing and DB opening those modules run an OpenMP loop over queries
 * (Prefiltering.cpp:820-900, structurealign.cpp:284-452).  The entry points below replace exactly that loop:
 * plain pointers and sizes in, plain arrays out, every function returns 0 on success or a negative FSB200_E_*
 * code (the host maps != 0 to Debug(ERROR)+EXIT(EXIT_FAILURE) like the reference).
 * The caller owns all host buffers; the library owns device memory. One context drives one GPU.
 * There is no CPU fallback: without a usable CUDA device fsb200_create fails.
 */
#ifndef FSB200_H
#define FSB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB200_OK 0
#define FSB200_E_ARG (-1)      /* invalid argument */
#define FSB200_E_CUDA (-2)     /* CUDA runtime error, see fsb200_last_error */
#define FSB200_E_STATE (-3)    /* call order violated (e.g. align before load_target) */
#define FSB200_E_MATRIX (-4)   /* substitution matrix could not be parsed */
#define FSB200_E_RANGE (-5)    /* input exceeds a documented limit */

typedef struct fsb200_ctx fsb200_ctx;

/* Parameters that reach the two modules; defaults = `foldseek search` defaults. */
typedef struct fsb200_config {
    int32_t device;                 /* CUDA device ordinal */
    const char *mat3di_text;        /* contents of the 3Di matrix file (reference data/mat3di.out; LocalParameters.cpp:94-96) */
    const char *mataa_text;         /* contents of the AA matrix file (reference blosum62.out; structurealign.cpp:253-262) */
    const uint8_t *evalue_nn;       /* contents of evalue_nn.kerasify (EvalueNeuralNet.cpp:9-16) */
    uint64_t evalue_nn_len;
    int32_t gap_open;               /* --gap-open   10 (LocalParameters.cpp:404-418) */
    int32_t gap_extend;             /* --gap-extend  1 */
    int32_t comp_bias_corr;         /* --comp-bias-corr 1 */
    float comp_bias_scale_align;    /* 0.5  (StructureSearch.cpp:107) */
    float comp_bias_scale_prefilter;/* 0.15 (StructureSearch.cpp:101-102) */
    float sensitivity;              /* -s 9.5 */
    int32_t kmer_size;              /* -k 6 */
    int32_t max_seqs;               /* --max-seqs 1000 */
    int32_t min_diag_score;         /* --min-ungapped-score 30 */
    double evalue_thr;              /* -e 10 (code default, LocalParameters.cpp:409) */
    float aa_factor;                /* 1.4 for --alignment-type 2 (3Di+AA), 0 for 3Di only (structurealign.cpp:264) */
    int32_t diag_bin_count;         /* BINCOUNT the reference's CacheFriendlyOperations would use on this host
                                       (QueryMatcher.cpp:460-487); 0 = derive from db size and host L2 like the reference */
    int32_t tmscore_norm_mode;      /* --tmscore-threshold-mode: length the TM-score of fsb200_structure_scores_run is normalised by
                                       (TMaligner::normalization, src/commons/TMaligner.cpp:330-341): 0 = min(qEnd-qStart, dbEnd-dbStart)
                                       of the alignment (default), 1 = query length, 2 = target length, 3 = the shorter of the two */
} fsb200_config;

/* Hash of the sources (csrc + this header) the library was compiled from, set by the build (__graft_entry__.build);
 * tests/test_abi_exports.py checks it against the tree, so a stale binary cannot pass for the current sources. */
const char *fsb200_build_id(void);

/* Fills *cfg with the defaults above (pointers NULL). */
void fsb200_default_config(fsb200_config *cfg);

int fsb200_create(const fsb200_config *cfg, fsb200_ctx **out);
void fsb200_destroy(fsb200_ctx *ctx);
const char *fsb200_last_error(const fsb200_ctx *ctx);

/* 21x21 score matrices the context derived (row-major int16): which = 0 SW-AA, 1 SW-3Di, 2 k-mer 3Di (x8), 3 ungapped 3Di (x2).
 * Mirrors SubstitutionMatrix::subMatrix (lib/mmseqs/src/commons/BaseMatrix.cpp:142-162). */
int fsb200_get_matrix(const fsb200_ctx *ctx, int32_t which, int16_t *out441);

/* Stage the target DB in HBM. aa/tdi: residue codes 0..20 (A C D E F G H I K L M N P Q R S T V W Y X), concatenated;
 * off[i]/len[i] locate target i. aa_db_residues = DBReader::getAminoAcidDBSize() of the target AA DB
 * (lib/mmseqs/src/commons/DBReader.cpp:403-412), it feeds the E-value (EvalueNeuralNet.cpp:10). */
int fsb200_load_target(fsb200_ctx *ctx, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len,
                       uint32_t n, uint64_t aa_db_residues);

/* Query batch (same encoding). Replaces the per-query mapSequence/ssw_init calls (structurealign.cpp:326-347). */
int fsb200_set_queries(fsb200_ctx *ctx, const uint8_t *aa, const uint8_t *tdi, const uint64_t *off, const uint32_t *len,
                       uint32_t nq);

/* Per query values the reference computes on the host: (mu, lambda) of EvalueNeuralNet::predictMuLambda and the
 * composition bias of ssw_init; exposed for parity tests. bias arrays may be NULL. */
int fsb200_query_info(fsb200_ctx *ctx, uint32_t q, double *mu_lambda2, int8_t *bias_aa, int8_t *bias_ss);

/* Result of alignStructure() up to the traceback (structurealign.cpp:37-74) for one (query, candidate) pair. */
typedef struct fsb200_aln {
    uint32_t fwd_score;   /* align.score1 of the forward pass */
    uint32_t rev_score;   /* revAlign.score1 (0 when the pair was rejected before the reverse pass) */
    int32_t q_end;        /* qEndPos1  (-1 when rejected before it was needed) */
    int32_t db_end;       /* dbEndPos1 */
    int32_t status;       /* 0 = passed both E-value gates, 1 = rejected after forward pass, 2 = rejected after reverse pass */
    int32_t score;        /* fwd_score - rev_score for status 0/2 */
    double evalue;        /* E-value of the last gate evaluated */
} fsb200_aln;

/* Stage 2 score phase for all queries of the batch. cand_off[nq+1] is a CSR into cand_target[] (target ids in
 * prefilter order). out[] has cand_off[nq] entries in the same order. */
int fsb200_align(fsb200_ctx *ctx, const uint64_t *cand_off, const uint32_t *cand_target, fsb200_aln *out);

/* Same computation split for measurement: upload candidates / run on resident data / download. */
int fsb200_align_upload(fsb200_ctx *ctx, const uint64_t *cand_off, const uint32_t *cand_target);
int fsb200_align_run(fsb200_ctx *ctx);
int fsb200_align_download(fsb200_ctx *ctx, fsb200_aln *out);

/* Counters of the last fsb200_align_run: [0] pairs, [1] forward DP cells, [2] reverse DP cells, [3] end-row DP cells,
 * [4] kernel launches, [5] pairs redone by the int32 kernel, [6] dominant-kernel time in ns (CUDA events),
 * [7] dominant-kernel launches. */
int fsb200_align_stats(const fsb200_ctx *ctx, uint64_t *out8);

/* Start positions and backtrace of one accepted pair: what StructureSmithWaterman::alignStartPosBacktraceBlock
 * (src/commons/StructureSmithWaterman.cpp:369-537, block-aligner engine lib/mmseqs/lib/block-aligner/src/scan_block.rs)
 * adds to the s_align of alignStructure (structurealign.cpp:77-87). */
typedef struct fsb200_bt {
    int32_t q_start, db_start;   /* qStartPos1 / dbStartPos1; -1 when status != 1 */
    uint32_t identical_aa;       /* identicalAACnt (StructureSmithWaterman.cpp:498-503) */
    uint32_t aln_len;            /* length of the backtrace string */
    uint64_t bt_off;             /* the pair's backtrace is bt[bt_off .. bt_off + aln_len), letters M / I / D */
    int32_t status;              /* 1 traced; 0 the block aligner did not reproduce the forward score (the reference then
                                    keeps an empty backtrace, :474-477); -1 pair was not accepted (fsb200_aln.status != 0) */
    int32_t block_score;         /* score of the last block-aligner attempt */
} fsb200_bt;

/* Traceback for every accepted pair (fsb200_aln.status == 0) of the last fsb200_align_run / fsb200_search_run.
 * meta[] is indexed like the out[] of fsb200_align (candidate CSR order; after fsb200_search: the prefix sums of counts[]).
 * fsb200_backtrace_size reports the bytes fsb200_backtrace_download will write to bt. */
int fsb200_backtrace_run(fsb200_ctx *ctx);
int fsb200_backtrace_size(const fsb200_ctx *ctx, uint64_t *bytes);
int fsb200_backtrace_download(fsb200_ctx *ctx, fsb200_bt *meta, char *bt, uint64_t bt_cap);
/* [0] pairs traced, [1] kernel launches, [2] kernel ns (CUDA events), [3] pairs redone with a larger trace buffer,
 * [4] pairs whose block-aligner score did not match. */
int fsb200_backtrace_stats(const fsb200_ctx *ctx, uint64_t *out8);

/* TEST HOOK, not part of the reference-facing surface: ONE run of the CUDA block-aligner engine (the code fsb200_backtrace_run
 * uses) in trace + x-drop mode on caller-supplied sequences (codes 0..20, engine orientation = as given), per-position query
 * biases, the two 32x32 int8 score tables (row/column 21 = padding, i8::MIN), gap costs (negative, block-aligner convention),
 * block range and x-drop -- the parameters of block_align_3di_aa_trace_xdrop (lib/mmseqs/lib/block-aligner/c/block_aligner.h:635).
 * out3 = {score, query_idx, reference_idx} of AlignResult; bt = the M/I/D path from the origin to that cell.  Lets the crate's own
 * unit-test vectors (scan_block.rs:2267-2430) run through the CUDA engine (tests/test_traceback_gpu.py). */
int fsb200_test_block_align(fsb200_ctx *ctx, const uint8_t *q_aa, const uint8_t *q_3di, const int8_t *q_bias_aa, const int8_t *q_bias_ss,
                            uint32_t q_len, const uint8_t *t_aa, const uint8_t *t_3di, uint32_t t_len, const int8_t *tab_aa, const int8_t *tab_3di,
                            int32_t gap_open, int32_t gap_extend, uint32_t min_size, uint32_t max_size, int32_t x_drop,
                            int32_t *out3, char *bt, uint32_t bt_cap);

/* C-alpha coordinates (what Coordinate16::read returns, src/commons/Coordinate16.h:15-57): x[], y[], z[] hold the residues of
 * entry i at off[i] .. off[i] + len[i]; the lengths are those of fsb200_load_target / fsb200_set_queries. */
int fsb200_load_target_ca(fsb200_ctx *ctx, const float *x, const float *y, const float *z, const uint64_t *off);
int fsb200_set_queries_ca(fsb200_ctx *ctx, const float *x, const float *y, const float *z, const uint64_t *off);

/* LDDT and TM-score of one traced alignment and the structure-bits score derived from them (structurealign.cpp:376-412:
 * TMaligner::computeTMscore with normalisation min(qEnd-qStart, dbEnd-dbStart), LDDTCalculator::computeLDDTScore,
 * score = int(score * sqrt(lddt * tm))). */
typedef struct fsb200_sscore {
    double tm;            /* TMscoreResult::tmscore */
    float lddt;           /* LDDTScoreResult::avgLddtScore (also the dbcov column of the reference's result) */
    int32_t status;       /* 1 scored; 0 a degenerate superposition (NaN rotation; the reference's classic-Kabsch fallback is not
                             reproduced, tm is not meaningful); -1 pair has no backtrace */
    int32_t score_bits;   /* res.score after the rescoring */
} fsb200_sscore;

/* Scores every traced pair of the last fsb200_backtrace_run (its backtraces are still on the device). out[] is indexed like the
 * meta[] of fsb200_backtrace_download. Stats: [0] pairs scored, [1] launches, [2] kernel ns, [3] degenerate superpositions. */
int fsb200_structure_scores_run(fsb200_ctx *ctx);
int fsb200_structure_scores_download(fsb200_ctx *ctx, fsb200_sscore *out);
int fsb200_structure_scores_stats(const fsb200_ctx *ctx, uint64_t *out8);

/* One prefilter hit (QueryMatcher.h:31-48 hit_t). */
typedef struct fsb200_hit {
    uint32_t target;
    int32_t score;
    int32_t diagonal;     /* (short)(i - j) as written by prefilterHitToBuffer (QueryMatcher.h:120-132) */
} fsb200_hit;

/* Stage 1 for all queries of the batch: builds the k-mer index of the loaded target on first use.
 * hits must hold nq*max_seqs entries; query q writes hits[q*max_seqs .. +counts[q]). */
int fsb200_prefilter(fsb200_ctx *ctx, fsb200_hit *hits, uint32_t *counts);

/* Counters of the last fsb200_prefilter: [0] similar k-mers, [1] index hits, [2] double-diagonal
 * candidates, [3] ungapped cells, [4] kernel launches, [5] overflow flushes, [6] ns of the matching kernels (lists + expand +
 * diag), [7] launches of the dominant kernel. */
int fsb200_prefilter_stats(const fsb200_ctx *ctx, uint64_t *out8);

/* Device time (CUDA events on the context's stream) of the last fsb200_prefilter per pipeline stage, in ns:
 * [0] window counts + prefix sums + chunk bounds, [1] kmer_lists_kernel (similar k-mers -> index list heads),
 * [2] expand_kernel (lists -> bucketed hit stream), [3] diag_kernel (double-diagonal detection), [4] double-hit sort,
 * [5] score_kernel (ungapped rescoring), [6] best/candidate sort/select, [7] unused. */
int fsb200_prefilter_stage_ns(const fsb200_ctx *ctx, uint64_t *out8);

/* Query DB == target DB (Prefiltering.cpp:856-868 passes the query's own target id as identityId when the two DBs are the
 * same): identity[q] = index of the target that is query q itself, 0xFFFFFFFF for none.  QueryMatcher::getResult
 * (QueryMatcher.cpp:409-421) then reports that target first with score USHRT_MAX and diagonal 0 and skips it in the hit list.
 * NULL clears the setting.  Applies to the current query batch; fsb200_set_queries resets it. */
int fsb200_set_identity(fsb200_ctx *ctx, const uint32_t *identity);

/* Both stages back to back, the way data/structuresearch.sh:50,121 chains the two modules: the prefilter hit list of
 * every query (in prefilter order) is the candidate list of its alignments. fsb200_search_run keeps everything inside
 * the context; fsb200_search_download copies hits[nq*max_seqs], counts[nq] and aln[nq*max_seqs] (entry q*max_seqs+k
 * belongs to hit k of query q) to the caller. fsb200_search = run + download. */
int fsb200_search_run(fsb200_ctx *ctx);
int fsb200_search_download(fsb200_ctx *ctx, fsb200_hit *hits, uint32_t *counts, fsb200_aln *aln);
int fsb200_search(fsb200_ctx *ctx, fsb200_hit *hits, uint32_t *counts, fsb200_aln *aln);

/* Builds the 3-mer neighbour tables and the k-mer index of the staged target now instead of at the first fsb200_prefilter
 * (IndexBuilder::fillDatabase, lib/mmseqs/src/prefiltering/IndexBuilder.cpp:60-260; Prefiltering.cpp:255-268). */
int fsb200_build_index(fsb200_ctx *ctx);

/* Lanes: a second (third, ...) context on the SAME GPU that shares the parent's read-only device data - staged target, target
 * C-alpha and, if the parent has built them (fsb200_build_index or an earlier prefilter call), the 3-mer tables and the k-mer index -
 * and owns everything else (stream, queries, scratch, results).  One host thread per lane, each
 * with its own query range, keeps the GPU busy while another lane is in a host-side part of its step (the reference's analogue is
 * the OpenMP team over queries, Prefiltering.cpp:820-900 / structurealign.cpp:284-452).  The parent must outlive its lanes and must
 * not reload its target while lanes exist (FSB200_E_STATE).  Scratch budgets of the contexts of a group are divided by the group size. */
int fsb200_create_lane(fsb200_ctx *parent, fsb200_ctx **out);

/* OpenMP team size of this context's host-side loops (0 = runtime default).  Lanes driven by different host threads each use
 * their own team. */
int fsb200_set_host_threads(fsb200_ctx *ctx, int32_t n);

/* CUDA-event stopwatch on the context's stream (bench.py times the steps on the device with it). */
int fsb200_timer_start(fsb200_ctx *ctx);
int fsb200_timer_stop(fsb200_ctx *ctx, double *elapsed_ms);

/* Target DB already resident in device memory in the library's staging layout (produced by another rank's context and
 * received over NCCL): aa/tdi are device pointers to `bytes` bytes each, off/len are host arrays. */
int fsb200_staged_target_size(const fsb200_ctx *ctx, uint64_t *bytes);
int fsb200_export_staged_target(const fsb200_ctx *ctx, void **d_aa, void **d_tdi, uint64_t *off_out, uint32_t *len_out);
int fsb200_load_target_device(fsb200_ctx *ctx, const void *d_aa, const void *d_tdi, uint64_t bytes, const uint64_t *off,
                              const uint32_t *len, uint32_t n, uint64_t aa_db_residues);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
