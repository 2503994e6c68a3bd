// oracle/blockaligner_stub.cpp — TEST INFRASTRUCTURE.
// The reference links the Rust crate lib/mmseqs/lib/block-aligner through the C ABI in
// lib/mmseqs/lib/block-aligner/c/block_aligner.h. There is no Rust toolchain in this image, so
// this stub provides the 18 symbols StructureSmithWaterman.cpp uses (:129,:200,:387-:533) and always
// reports an alignment score that cannot match (so alignStartPosBacktraceBlock takes its
// "score mismatch" exit, StructureSmithWaterman.cpp:474-477). With this stub the reference binary is
// authoritative for: prefilter hit lists, forward/reverse SW scores, end positions and E-values.
// Start positions / backtrace / seqId / TM / LDDT columns of its output are NOT authoritative.
#include <cstdint>
#include <cstdlib>
#include "block_aligner.h"
extern "C" {
struct AAMatrix { int dummy; };
struct PaddedBytes { int dummy; };
struct PosBias { int dummy; };
struct Cigar { int dummy; };
BlockHandle block_new_aa_trace_xdrop(uintptr_t, uintptr_t, uintptr_t) { return malloc(8); }
void block_free_aa_trace_xdrop(BlockHandle b) { free(b); }
PaddedBytes *block_new_padded_aa(uintptr_t, uintptr_t) { return new PaddedBytes(); }
void block_set_bytes_padded_aa(PaddedBytes *, const uint8_t *, uintptr_t, uintptr_t) {}
void block_free_padded_aa(PaddedBytes *p) { delete p; }
PosBias *block_new_pos_bias(uintptr_t, uintptr_t) { return new PosBias(); }
void block_set_pos_bias(PosBias *, const int16_t *, uintptr_t) {}
void block_free_pos_bias(PosBias *p) { delete p; }
AAMatrix *block_new_simple_aamatrix(int8_t, int8_t) { return new AAMatrix(); }
void block_set_aamatrix(AAMatrix *, uint8_t, uint8_t, int8_t) {}
void block_free_aamatrix(AAMatrix *m) { delete m; }
Cigar *block_new_cigar(uintptr_t, uintptr_t) { return new Cigar(); }
OpLen block_get_cigar(const Cigar *, uintptr_t) { OpLen o; o.op = (Operation)0; o.len = 0; return o; }
uintptr_t block_len_cigar(const Cigar *) { return 0; }
void block_free_cigar(Cigar *c) { delete c; }
void block_align_3di_aa_trace_xdrop(BlockHandle, const PaddedBytes *, const PaddedBytes *, const PosBias *,
                                    const PaddedBytes *, const PaddedBytes *, const PosBias *,
                                    const AAMatrix *, const AAMatrix *, Gaps, SizeRange, int32_t) {}
AlignResult block_res_aa_trace_xdrop(BlockHandle) { AlignResult r; r.score = -1000000000; r.query_idx = 0; r.reference_idx = 0; return r; }
void block_cigar_aa_trace_xdrop(BlockHandle, uintptr_t, uintptr_t, Cigar *) {}
}
