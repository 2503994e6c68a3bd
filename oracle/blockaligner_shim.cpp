// oracle/blockaligner_shim.cpp — TEST INFRASTRUCTURE.
// The reference links the Rust crate lib/mmseqs/lib/block-aligner through the C ABI in
// lib/mmseqs/lib/block-aligner/c/block_aligner.h. There is no Rust toolchain in this image, so this shim provides the
// 18 symbols StructureSmithWaterman.cpp uses (:129,:200,:387-:533) on top of the scalar restatement of the engine in
// block_aligner_port.h (pinned by the crate's own unit-test vectors, tests/test_block_aligner_oracle.py).
// With this shim the reference binary is authoritative for: prefilter hit lists, forward/reverse SW scores, end
// positions, E-values and everything the reference's own code derives from a given backtrace (TM-score, LDDT, seqId);
// the backtrace / start positions themselves come from the restated engine ("parity unpinned" beyond the unit vectors).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "block_aligner.h"
#include "block_aligner_port.h"

extern "C" {
struct AAMatrix { int8_t s[27 * 32]; };                 // scores.rs:47-49
struct PaddedBytes { std::vector<uint8_t> s; size_t len; };
struct PosBias { std::vector<int16_t> b; size_t len; };
struct Cigar { std::vector<std::pair<ba::Op, size_t>> ops; };
}
namespace {
struct Handle { ba::Block blk; Handle() : blk(true, true) {} };
inline uint8_t conv(uint8_t c) { if (c >= 'a' && c <= 'z') c -= 32; return (uint8_t) (c - 'A'); }  // AAMatrix::convert_char
}

extern "C" {
BlockHandle block_new_aa_trace_xdrop(uintptr_t, uintptr_t, uintptr_t) { return new Handle(); }
void block_free_aa_trace_xdrop(BlockHandle b) { delete (Handle *) b; }
PaddedBytes *block_new_padded_aa(uintptr_t len, uintptr_t maxSize) {
    PaddedBytes *p = new PaddedBytes();
    p->s.assign(1 + len + maxSize + ba::L, 26);
    p->len = len;
    return p;
}
void block_set_bytes_padded_aa(PaddedBytes *p, const uint8_t *s, uintptr_t len, uintptr_t maxSize) {   // PaddedBytes::set_bytes
    if (p->s.size() < 1 + len + maxSize + ba::L) p->s.resize(1 + len + maxSize + ba::L);
    std::fill(p->s.begin(), p->s.end(), 26);
    for (size_t i = 0; i < len; i++) p->s[1 + i] = conv(s[i]);
    p->len = len;
}
void block_free_padded_aa(PaddedBytes *p) { delete p; }
PosBias *block_new_pos_bias(uintptr_t len, uintptr_t maxSize) {
    PosBias *p = new PosBias();
    p->b.assign(len + maxSize + 1 + ba::L, 0);
    p->len = len;
    return p;
}
void block_set_pos_bias(PosBias *p, const int16_t *b, uintptr_t len) {   // PosBias::set_biases
    std::fill(p->b.begin(), p->b.end(), 0);
    if (p->b.size() < len + 1) p->b.resize(len + 1 + 4096 + ba::L, 0);
    memcpy(p->b.data() + 1, b, len * sizeof(int16_t));
    p->len = len;
}
void block_free_pos_bias(PosBias *p) { delete p; }
AAMatrix *block_new_simple_aamatrix(int8_t match, int8_t mismatch) {     // AAMatrix::new_simple
    AAMatrix *m = new AAMatrix();
    memset(m->s, 0x80, sizeof(m->s));
    for (int i = 0; i < 26; i++) for (int j = 0; j < 26; j++) m->s[i * 32 + j] = i == j ? match : mismatch;
    return m;
}
void block_set_aamatrix(AAMatrix *m, uint8_t a, uint8_t b, int8_t score) {   // AAMatrix::set
    const int x = conv(a), y = conv(b);
    m->s[x * 32 + y] = score;
    m->s[y * 32 + x] = score;
}
void block_free_aamatrix(AAMatrix *m) { delete m; }
Cigar *block_new_cigar(uintptr_t, uintptr_t) { return new Cigar(); }
OpLen block_get_cigar(const Cigar *c, uintptr_t i) { OpLen o; o.op = (Operation) c->ops[i].first; o.len = c->ops[i].second; return o; }
uintptr_t block_len_cigar(const Cigar *c) { return c->ops.size(); }
void block_free_cigar(Cigar *c) { delete c; }
void block_align_3di_aa_trace_xdrop(BlockHandle h, const PaddedBytes *q, const PaddedBytes *q3, const PosBias *qb,
                                    const PaddedBytes *r, const PaddedBytes *r3, const PosBias *rb,
                                    const AAMatrix *m, const AAMatrix *m3, Gaps g, SizeRange s, int32_t x) {
    ba::Scorer3di sc;
    sc.qAA = q->s.data(); sc.q3Di = q3->s.data(); sc.rAA = r->s.data(); sc.r3Di = r3->s.data();
    sc.qBias = qb->b.data(); sc.rBias = rb->b.data(); sc.ql = q->len; sc.rl = r->len;
    sc.tabAA = m->s; sc.tab3Di = m3->s; sc.stride = 32;
    ((Handle *) h)->blk.align(sc, g.open, g.extend, s.min, s.max, x);
}
AlignResult block_res_aa_trace_xdrop(BlockHandle h) {
    const ba::AlignResult &a = ((Handle *) h)->blk.res;
    AlignResult r; r.score = a.score; r.query_idx = a.query_idx; r.reference_idx = a.reference_idx;
    return r;
}
void block_cigar_aa_trace_xdrop(BlockHandle h, uintptr_t qi, uintptr_t ri, Cigar *c) {
    c->ops = ((Handle *) h)->blk.trace.cigar(qi, ri);
}
}
