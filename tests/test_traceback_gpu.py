"""Parity of the CUDA traceback (fsb200_backtrace_*) against the oracle's block-aligner restatement: start positions,
identical-residue counts and backtrace strings must be identical, byte for byte."""
import numpy as np
import pytest

from foldseek_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, q, db, cands):
    dbres = int(sum(len(a) for a in db.aa))
    ctx.load_target(db.aa, db.tdi, dbres)
    ctx.set_queries(q.aa, q.tdi)
    off, flat = api.Context.csr(cands)
    res = ctx.align(off, flat)
    meta, bt = ctx.backtrace(len(res))
    n = 0
    for qi in range(len(q)):
        for k in range(int(off[qi]), int(off[qi + 1])):
            r, m = res[k], meta[k]
            if r["status"] != 0:
                assert m["status"] == -1
                continue
            t = int(flat[k])
            ok, qs, ds, ids, bscore, s = oracle.backtrace(q.aa[qi], q.tdi[qi], db.aa[t], db.tdi[t], r["q_end"], r["db_end"],
                                                          r["fwd_score"])
            assert m["status"] == (1 if ok else 0), (qi, t, m, ok, bscore)
            assert m["block_score"] == bscore, (qi, t)
            if ok:
                got = bytes(bt[int(m["bt_off"]):int(m["bt_off"]) + int(m["aln_len"])]).decode()
                assert got == s, (qi, t, got[:60], s[:60])
                assert (m["q_start"], m["db_start"], m["identical_aa"]) == (qs, ds, ids), (qi, t)
            n += 1
    return n


def test_planted_pairs_varied_lengths(ctx, oracle):
    db = synth.gen_db(120, 420, 41, min_len=25)
    q = synth.gen_queries(db, 48, 300, 42, planted_frac=1.0, sub_prob=0.35, indel_prob=0.08)
    cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
    assert _check(ctx, oracle, q, db, cands) > 40


def test_long_gappy_pairs_grow_blocks(ctx, oracle):
    """long sequences with long indels force block growth, checkpoints restores and shrinking"""
    rng = np.random.default_rng(5)
    db = synth.gen_db(16, 1500, 51, min_len=700)
    q = synth.gen_queries(db, 16, 1500, 52, planted_frac=1.0, sub_prob=0.3, indel_prob=0.04)
    for i in range(len(q)):      # cut a 40..120 residue piece out of every query and insert a random piece elsewhere
        a, s = q.aa[i], q.tdi[i]
        c0 = int(rng.integers(100, len(a) - 300)); cl = int(rng.integers(40, 120))
        a = np.concatenate([a[:c0], a[c0 + cl:]]); s = np.concatenate([s[:c0], s[c0 + cl:]])
        p0 = int(rng.integers(50, len(a) - 50)); pl = int(rng.integers(30, 90))
        ins = synth.gen_db(1, pl, 1000 + i)
        a = np.concatenate([a[:p0], ins.aa[0], a[p0:]]); s = np.concatenate([s[:p0], ins.tdi[0], s[p0:]])
        q.aa[i], q.tdi[i] = a.astype(np.uint8), s.astype(np.uint8)
    cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
    assert _check(ctx, oracle, q, db, cands) >= 16
    st = ctx.backtrace_stats()
    assert st[0] >= 16 and st[1] >= 1


# ---- the CUDA engine itself against crate-held vectors and the pinned single-matrix engine (test hook fsb200_test_block_align) ----
def _tabs(pairs, fill=-128):
    """32 x 32 int8 table over codes 0..20, row/column 21.. = padding (i8::MIN)"""
    t = np.full((32, 32), -128, np.int8)
    t[:21, :21] = fill
    for (x, y), v in pairs.items():
        t[x, y] = t[y, x] = v
    return t


def _expand(cigar):
    out, n = "", ""
    for ch in cigar:
        if ch.isdigit():
            n += ch
        else:
            out += ch * int(n); n = ""
    return out


def test_cuda_engine_runs_the_crates_xdrop_vectors(ctx):
    """scan_block.rs:2338-2370 test_x_drop (BLOSUM62 A/A 4, A/R -1, R/R 5, gaps -11/-1, block 16..16, x-drop 1), through the CUDA
    engine's 3Di+AA cell score with (a) the 3Di table all-zero and (b) the AA table all-zero"""
    A, R = 0, 14
    blosum = _tabs({(A, A): 4, (A, R): -1, (R, R): 5}, fill=-4)
    zero = _tabs({}, fill=0)
    conv = lambda s: np.array([A if c == ord("A") else R for c in s], np.uint8)
    cases = [(b"AAAAAA", b"AAARRA", (14, 6, 6)),
             (b"A" * 44, b"AAAAAAAAAAAAAAARRRRRRRRRRRRRRRRAAAAAAAAAAAAA", (60, 15, 15))]
    for q, r, want in cases:
        junk_q, junk_r = conv(q)[::-1].copy() % 3, conv(r)[::-1].copy() % 3
        assert ctx.test_block_align(conv(q), junk_q, conv(r), junk_r, blosum, zero, -11, -1, 16, 16, 1)[:3] == want
        assert ctx.test_block_align(junk_q, conv(q), junk_r, conv(r), zero, blosum, -11, -1, 16, 16, 1)[:3] == want
    long = np.zeros(2048, np.uint8)
    assert ctx.test_block_align(long, long, long, long, blosum, zero, -11, -1, 2048, 2048, 100)[:3] == (8192, 2048, 2048)


def test_cuda_engine_equals_pinned_single_matrix_engine(ctx):
    """random sequences, block ranges and x-drops: CUDA 3Di+AA engine with one table zeroed == the restatement's single-matrix path
    (ba::MatrixScorer, the path the crate's own vectors pin), score, end cell and path; and with both tables and biases ==
    the restatement's 3Di+AA path"""
    from oracle.oracle import ba_align, ba_align_3di
    rng = np.random.default_rng(17)
    m = rng.integers(-6, 9, size=(21, 21)); m = np.minimum(m, m.T)
    np.fill_diagonal(m, rng.integers(3, 10, size=21))
    tab = np.full((32, 32), -128, np.int8); tab[:21, :21] = m
    m2 = rng.integers(-4, 7, size=(21, 21)); m2 = np.minimum(m2, m2.T)
    tab2 = np.full((32, 32), -128, np.int8); tab2[:21, :21] = m2
    zero = np.full((32, 32), -128, np.int8); zero[:21, :21] = 0
    n = 0
    for it in range(60):
        ql, rl = int(rng.integers(5, 400)), int(rng.integers(5, 400))
        r = rng.integers(0, 20, size=rl).astype(np.uint8)
        q = r[:ql].copy() if (it % 2 == 0 and rl >= ql) else rng.integers(0, 20, size=ql).astype(np.uint8)
        mut = rng.random(len(q)) < 0.25
        q[mut] = rng.integers(0, 20, size=int(mut.sum()))
        mn = int(2 ** rng.integers(4, 8)); mx = int(mn * 2 ** rng.integers(0, 4))
        xd = int(rng.integers(5, 60))
        go, ge = -int(rng.integers(4, 12)), -int(rng.integers(1, 3))
        want = ba_align(bytes(q), bytes(r), tab, 32, 21, go, ge, mn, mx, x_drop=xd, trace=True, xdrop=True)
        got = ctx.test_block_align(q, q, r, r, tab, zero, go, ge, mn, mx, xd)
        assert got[:3] == want[:3], (it, got[:3], want[:3])
        assert got[3] == _expand(want[3])[::-1], it      # the kernel writes the path from the alignment start to the end cell
        bias_a = rng.integers(-3, 4, size=ql).astype(np.int8); bias_s = rng.integers(-2, 3, size=ql).astype(np.int8)
        q2 = rng.integers(0, 20, size=ql).astype(np.uint8); r2 = rng.integers(0, 20, size=rl).astype(np.uint8)
        want2 = ba_align_3di(bytes(q), bytes(q2), bytes(r), bytes(r2), tab, tab2, 32, 21, go, ge, mn, mx, x_drop=xd, trace=True, xdrop=True,
                             q_bias=bias_a.astype(np.int16) + bias_s.astype(np.int16))
        got2 = ctx.test_block_align(q, q2, r, r2, tab, tab2, go, ge, mn, mx, xd, q_bias_aa=bias_a, q_bias_ss=bias_s)
        assert got2[:3] == want2[:3] and got2[3] == _expand(want2[3])[::-1], it
        n += 1
    assert n == 60
