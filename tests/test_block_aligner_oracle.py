"""CPU tests: pin the block-aligner restatement (oracle/block_aligner_port.h) against the known answers of the crate's own
unit tests (lib/mmseqs/lib/block-aligner/src/scan_block.rs:2267-2413: test_no_x_drop, test_x_drop, test_trace).
The Rust crate cannot be built in this image (no Rust toolchain), so these vectors are the only authority available for
the engine; the 3Di+AA scoring variant differs from the pinned path only in how the per-cell score is summed."""
import numpy as np
import pytest

from oracle.oracle import ba_align


def aa_table():
    """AAMatrix layout (scores.rs:47-49): 27 rows x 32 columns of int8 indexed by letter - 'A', default i8::MIN.
    Only the BLOSUM62 entries the vectors touch are filled (matrices/BLOSUM62: A/A 4, A/R -1, R/R 5)."""
    t = np.full((32, 32), -128, np.int8)
    a, r = 0, ord("R") - 65
    t[a, a] = 4
    t[a, r] = t[r, a] = -1
    t[r, r] = 5
    return t


def nuc_table(match=1, mismatch=-1):
    """NucMatrix (scores.rs:150-168): 8 rows x 16 columns indexed by (a & 7, b & 15) of the upper-case ASCII byte; expanded
    here to a 128 x 128 table over raw bytes so the engine can index it directly."""
    small = np.full((8, 16), -128, np.int8)
    for x in b"ATCGN":
        for y in b"ATCGN":
            small[x & 7, y & 15] = match if x == y else mismatch
    t = np.zeros((128, 128), np.int8)
    for a in range(128):
        for b in range(128):
            t[a, b] = small[a & 7, b & 15]
    return t


def aa(s):
    return bytes(c - 65 for c in s)


AA = dict(table=aa_table(), stride=32, null_code=26)
NUC = dict(table=nuc_table(), stride=128, null_code=ord("Z"))


@pytest.mark.parametrize("q,r,score", [
    (b"AARA", b"AAAA", 11), (b"AARAAAA", b"AAAAAAAA", 12), (b"AAAA", b"AAAA", 16), (b"RRRR", b"AAAA", -4), (b"AAA", b"AAAA", 1)])
def test_global_aa(q, r, score):      # test_no_x_drop, gaps -11/-1, block 16..16
    assert ba_align(aa(q), aa(r), gap_open=-11, gap_extend=-1, min_size=16, max_size=16, **AA)[0] == score


@pytest.mark.parametrize("q,r,score", [
    (b"ATAA", b"AAAN", 0), (b"A" * 32, b"A" * 32, 32), (b"T" * 32, b"A" * 32, -32), (b"TA" * 16, b"A" * 32, 0),
    (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", 7), (b"C", b"AAAA", -5), (b"AAAA", b"C", -5)])
def test_global_nuc(q, r, score):     # test_no_x_drop, gaps -2/-1
    assert ba_align(q, r, gap_open=-2, gap_extend=-1, min_size=16, max_size=16, **NUC)[0] == score


def test_x_drop():                     # test_x_drop
    assert ba_align(aa(b"AAAAAA"), aa(b"AAARRA"), gap_open=-11, gap_extend=-1, min_size=16, max_size=16, x_drop=1, xdrop=True,
                    **AA)[:3] == (14, 6, 6)
    r = b"AAAAAAAAAAAAAAARRRRRRRRRRRRRRRRAAAAAAAAAAAAA"
    assert ba_align(aa(b"A" * 44), aa(r), gap_open=-11, gap_extend=-1, min_size=16, max_size=16, x_drop=1, xdrop=True,
                    **AA)[:3] == (60, 15, 15)
    long = aa(b"A" * 2048)
    assert ba_align(long, long, gap_open=-11, gap_extend=-1, min_size=2048, max_size=2048, x_drop=100, trace=True, xdrop=True,
                    **AA)[:3] == (8192, 2048, 2048)


def test_trace():                      # test_trace
    assert ba_align(aa(b"AAAAAA"), aa(b"AAARRA"), gap_open=-11, gap_extend=-1, min_size=16, max_size=16, trace=True, eq=True,
                    **AA) == (14, 6, 6, "3=2X1=")
    assert ba_align(aa(b"AAA"), aa(b"AAAA"), gap_open=-11, gap_extend=-1, min_size=16, max_size=16, trace=True, **AA) \
        == (1, 3, 4, "3M1D")
    assert ba_align(b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", gap_open=-2, gap_extend=-1, min_size=16, max_size=16,
                    trace=True, **NUC) == (7, 24, 21, "2M6I16M3D")
    q, r = b"AAAAAAAAATTGCGCT", b"AAAAAAAAAGCGC"
    assert ba_align(q, r, gap_open=-2, gap_extend=-1, min_size=32, max_size=32, trace=True, eq=True, **NUC) \
        == (8, 16, 13, "9=2I4=1I")
    assert ba_align(q, r, nuc_table(2, -1), 128, ord("Z"), gap_open=-5, gap_extend=-2, min_size=32, max_size=32, trace=True,
                    eq=True) == (14, 16, 13, "9=2I4=1I")


def test_3di_aa_backtraces_rescore_to_the_forward_score():
    """The 3Di+AA scoring variant of the engine cannot be pinned against the crate here; this checks its internal consistency on
    the stage-2 oracle: for every accepted pair the block aligner reproduces the forward Smith-Waterman score, the path ends at
    (qEnd, dbEnd), and re-scoring the M/I/D string with the SW matrices, biases and gap costs gives exactly that score."""
    from foldseek_b200 import synth
    from oracle.oracle import Oracle
    o = Oracle()
    db = synth.gen_db(120, 300, seed=3, min_len=60)
    qs = synth.gen_queries(db, 10, 300, seed=4, planted_frac=1.0, sub_prob=0.35, indel_prob=0.06)
    m_aa, _ = o.matrix(0)
    m_3di, _ = o.matrix(1)
    tot = sum(len(t) for t in db.aa)
    n = 0
    for qi in range(len(qs)):
        out, _ = o.align_query(tot, qs.aa[qi], qs.tdi[qi], db.aa, db.tdi)
        b_aa, b_ss = o.query_bias(qs.aa[qi], qs.tdi[qi])
        for ti in range(len(db)):
            s, qe, de, _ = (int(v) for v in out[ti])
            if s < 25:
                continue
            ok, q0, d0, ids, bscore, bt = o.backtrace(qs.aa[qi], qs.tdi[qi], db.aa[ti], db.tdi[ti], qe, de, s)
            assert ok and bscore == s, (qi, ti, s, bscore)
            i, j, sc, prev, same = q0, d0, 0, "M", 0
            for c in bt:
                if c == "M":
                    sc += int(m_aa[db.aa[ti][j], qs.aa[qi][i]]) + int(m_3di[db.tdi[ti][j], qs.tdi[qi][i]]) + int(b_aa[i]) + int(b_ss[i])
                    same += int(db.aa[ti][j] == qs.aa[qi][i])
                    i += 1; j += 1
                else:
                    sc -= 1 if prev == c else 10
                    if c == "I":
                        i += 1
                    else:
                        j += 1
                prev = c
            assert (sc, i, j, same) == (s, qe + 1, de + 1, ids), (qi, ti, sc, s)
            n += 1
    assert n > 200


def test_prefix_scan_vectors():        # avx2.rs:470-490 test_prefix_scan
    from oracle.oracle import ba_prefix_scan
    v = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 12, 13, 14, 11]
    assert ba_prefix_scan(v, 0) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 15, 15, 15, 15]
    assert ba_prefix_scan(v, -1) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 14, 13, 14, 13]


def test_readme_example_block_range():   # README.md / lib.rs doc example: block sizes 32..=256
    assert ba_align(b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", gap_open=-2, gap_extend=-1, min_size=32, max_size=256,
                    trace=True, eq=True, **NUC) == (7, 24, 21, "2=6I16=3D")


def test_byte_matrix():                  # test_bytes: ByteMatrix(1, -1), padding byte 0
    t = np.full((256, 256), -1, np.int8)
    np.fill_diagonal(t, 1)
    kw = dict(table=t, stride=256, null_code=0, gap_open=-2, gap_extend=-1, min_size=16, max_size=16)
    assert ba_align(b"AAAAAA", b"AAAaaA", **kw)[0] == 2
    assert ba_align(b"abdefg", b"abcdefg", **kw)[0] == 4


def _gotoh_global(q, r, tab, go, ge):
    """Plain affine-gap global alignment score (gap of length n costs go + (n-1)*ge, both negative), O(|q||r|)."""
    NEG = -10 ** 9
    n, m = len(q), len(r)
    D = [[NEG] * (m + 1) for _ in range(n + 1)]
    R = [[NEG] * (m + 1) for _ in range(n + 1)]
    C = [[NEG] * (m + 1) for _ in range(n + 1)]
    D[0][0] = 0
    for i in range(n + 1):
        for j in range(m + 1):
            if i == 0 and j == 0:
                continue
            if i > 0:
                R[i][j] = max(R[i - 1][j] + ge, D[i - 1][j] + go)
            if j > 0:
                C[i][j] = max(C[i][j - 1] + ge, D[i][j - 1] + go)
            best = max(R[i][j], C[i][j])
            if i > 0 and j > 0:
                best = max(best, D[i - 1][j - 1] + int(tab[q[i - 1], r[j - 1]]))
            D[i][j] = best
    return D[n][m]


def test_block_covering_both_sequences_equals_full_dp():
    """When one block covers the whole DP matrix the engine must return the exact affine-gap global score
    (the recurrence documented at scan_block.rs:27-33)."""
    rng = np.random.default_rng(3)
    tab = np.full((32, 32), -128, np.int8)
    sub = rng.integers(-4, 7, size=(20, 20))
    sub = np.minimum(sub, sub.T)
    np.fill_diagonal(sub, rng.integers(3, 9, size=20))
    tab[:20, :20] = sub
    for _ in range(60):
        n, m = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        q = rng.integers(0, 20, size=n).astype(np.uint8)
        r = q.copy()[:m] if rng.random() < 0.5 and m <= n else rng.integers(0, 20, size=m).astype(np.uint8)
        go, ge = -int(rng.integers(2, 12)), -1
        got = ba_align(bytes(q), bytes(r), tab, 32, 26, gap_open=go, gap_extend=ge, min_size=64, max_size=64)[0]
        assert got == _gotoh_global(q, r, tab, go, ge), (n, m, go)


def test_xdrop_block_covering_both_sequences_finds_best_prefix_pair():
    """X-drop mode with one covering block and a huge threshold: score = max over all cells of the global-start DP, end cell = the
    maximal cell with the largest reference index, then the largest query index (scan_block.rs:373-444)."""
    rng = np.random.default_rng(4)
    tab = np.full((32, 32), -128, np.int8)
    sub = rng.integers(-5, 5, size=(20, 20))
    sub = np.minimum(sub, sub.T)
    np.fill_diagonal(sub, rng.integers(2, 7, size=20))
    tab[:20, :20] = sub
    NEG = -10 ** 9
    for _ in range(40):
        n, m = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        q = rng.integers(0, 20, size=n).astype(np.uint8)
        r = rng.integers(0, 20, size=m).astype(np.uint8)
        k = min(n, m) // 2
        r[:k] = q[:k]                      # a common prefix so that the best cell is not the origin
        go, ge = -int(rng.integers(3, 12)), -1
        D = [[NEG] * (m + 1) for _ in range(n + 1)]; R = [[NEG] * (m + 1) for _ in range(n + 1)]; C = [[NEG] * (m + 1) for _ in range(n + 1)]
        D[0][0] = 0
        best, bi, bj = 0, 0, 0
        for j in range(m + 1):
            for i in range(n + 1):
                if i or j:
                    if i: R[i][j] = max(R[i - 1][j] + ge, D[i - 1][j] + go)
                    if j: C[i][j] = max(C[i][j - 1] + ge, D[i][j - 1] + go)
                    v = max(R[i][j], C[i][j])
                    if i and j: v = max(v, D[i - 1][j - 1] + int(tab[q[i - 1], r[j - 1]]))
                    D[i][j] = v
                if D[i][j] > best or (D[i][j] == best and best > 0 and (j, i) > (bj, bi)):
                    best, bi, bj = D[i][j], i, j
        got = ba_align(bytes(q), bytes(r), tab, 32, 26, gap_open=go, gap_extend=ge, min_size=64, max_size=64, x_drop=100000, xdrop=True)
        assert got[:3] == (best, bi, bj), (n, m, got, best, bi, bj)


def test_global_cigar_rescoring():
    """The CIGAR of a global alignment consumes both sequences and re-scores to the reported score (small and growing blocks)."""
    import re
    rng = np.random.default_rng(5)
    tab = np.full((32, 32), -128, np.int8)
    sub = rng.integers(-4, 5, size=(20, 20))
    sub = np.minimum(sub, sub.T)
    np.fill_diagonal(sub, rng.integers(3, 8, size=20))
    tab[:20, :20] = sub
    for it in range(60):
        n = int(rng.integers(5, 200))
        q = rng.integers(0, 20, size=n).astype(np.uint8)
        r = [x for x in q if rng.random() > 0.05]                      # deletions
        r = [int(x) if rng.random() > 0.2 else int(rng.integers(0, 20)) for x in r]   # substitutions
        if rng.random() < 0.5 and len(r) > 10:                             # one long insertion
            p = int(rng.integers(0, len(r)))
            r[p:p] = [int(v) for v in rng.integers(0, 20, size=int(rng.integers(5, 40)))]
        r = np.array(r, np.uint8)
        go, ge = -int(rng.integers(3, 12)), -1
        lo = 16 if it % 2 else 32
        score, qi, ri, cigar = ba_align(bytes(q), bytes(r), tab, 32, 26, gap_open=go, gap_extend=ge, min_size=lo, max_size=512, trace=True)
        assert (qi, ri) == (len(q), len(r))
        i = j = sc = 0
        for cnt, op in re.findall(r"(\d+)([MID])", cigar):
            cnt = int(cnt)
            if op == "M":
                sc += sum(int(tab[q[i + k], r[j + k]]) for k in range(cnt)); i += cnt; j += cnt
            else:
                sc += go + (cnt - 1) * ge
                if op == "I":
                    i += cnt
                else:
                    j += cnt
        assert (i, j, sc) == (len(q), len(r), score), (it, cigar, sc, score)


# ---- the 3Di+AA entry (place_block_3di) pinned by the crate's vectors through algebra ----
# The 3Di+AA variant differs from place_block in one expression: the cell score adds(adds(sAA, s3Di), bias_r + bias_q)
# (scan_block.rs:1343-1351 vs :1202).  (a) with an all-zero 3Di table and zero biases it equals sAA, (b) with an all-zero AA table it
# equals s3Di, so every crate vector must give the crate's expected value through the 3Di+AA entry both ways; (c) a constant
# c added to every query position bias and subtracted from every table entry leaves all cell scores, hence everything, unchanged.
from oracle.oracle import ba_align_3di

VECTORS = [
    # (query, reference, matrix kind, gaps, block, kwargs, expected (score, query_idx, reference_idx, cigar) prefix)
    (b"AARA", b"AAAA", "aa", (-11, -1), (16, 16), {}, (11,)),
    (b"AARAAAA", b"AAAAAAAA", "aa", (-11, -1), (16, 16), {}, (12,)),
    (b"AAAA", b"AAAA", "aa", (-11, -1), (16, 16), {}, (16,)),
    (b"RRRR", b"AAAA", "aa", (-11, -1), (16, 16), {}, (-4,)),
    (b"AAA", b"AAAA", "aa", (-11, -1), (16, 16), {}, (1,)),
    (b"ATAA", b"AAAN", "nuc", (-2, -1), (16, 16), {}, (0,)),
    (b"T" * 32, b"A" * 32, "nuc", (-2, -1), (16, 16), {}, (-32,)),
    (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", "nuc", (-2, -1), (16, 16), {}, (7,)),
    (b"C", b"AAAA", "nuc", (-2, -1), (16, 16), {}, (-5,)),
    (b"AAAAAA", b"AAARRA", "aa", (-11, -1), (16, 16), dict(x_drop=1, xdrop=True), (14, 6, 6)),
    (b"A" * 44, b"AAAAAAAAAAAAAAARRRRRRRRRRRRRRRRAAAAAAAAAAAAA", "aa", (-11, -1), (16, 16), dict(x_drop=1, xdrop=True), (60, 15, 15)),
    (b"AAAAAA", b"AAARRA", "aa", (-11, -1), (16, 16), dict(trace=True, eq=True), (14, 6, 6, "3=2X1=")),
    (b"AAA", b"AAAA", "aa", (-11, -1), (16, 16), dict(trace=True), (1, 3, 4, "3M1D")),
    (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", "nuc", (-2, -1), (16, 16), dict(trace=True), (7, 24, 21, "2M6I16M3D")),
    (b"AAAAAAAAATTGCGCT", b"AAAAAAAAAGCGC", "nuc", (-2, -1), (32, 32), dict(trace=True, eq=True), (8, 16, 13, "9=2I4=1I")),
    (b"TTTTTTTTAAAAAAATTTTTTTTT", b"TTAAAAAAATTTTTTTTTTTT", "nuc", (-2, -1), (32, 256), dict(trace=True, eq=True), (7, 24, 21, "2=6I16=3D")),
]


def _conv(kind, s):
    return aa(s) if kind == "aa" else bytes(s)


@pytest.mark.parametrize("mode", ["3di_table_zero", "aa_table_zero"])
@pytest.mark.parametrize("vec", VECTORS)
def test_crate_vectors_through_the_3di_aa_entry(vec, mode):
    q, r, kind, (go, ge), (mn, mx), kw, want = vec
    m = AA if kind == "aa" else NUC
    table, stride, null = m["table"], m["stride"], m["null_code"]
    zero = np.zeros_like(table)
    cq, cr = _conv(kind, q), _conv(kind, r)
    # the channel with the all-zero table carries unrelated letters: they must not matter (eq-mode CIGARs compare the AA channel,
    # so that channel keeps the vector when it is the scoring one)
    junk_q, junk_r = bytes((x * 7 + 3) % 5 if kind == "aa" else x for x in cq), bytes((x * 5 + 1) % 5 if kind == "aa" else x for x in cr)
    if mode == "3di_table_zero":
        got = ba_align_3di(cq, junk_q, cr, junk_r, table, zero, stride, null, go, ge, mn, mx, **kw)
    else:
        if kw.get("eq"):
            kw = dict(kw); kw["eq"] = False
            want = want[:3]      # =/X ops are derived from the AA channel, which is junk in this mode
        got = ba_align_3di(junk_q, cq, junk_r, cr, zero, table, stride, null, go, ge, mn, mx, **kw)
    assert got[:len(want)] == want


@pytest.mark.parametrize("shift", [1, -2, 3])
@pytest.mark.parametrize("vec", [v for v in VECTORS if v[2] == "aa"])
def test_position_bias_shift_identity(vec, shift):
    """query bias +c at every position and every finite table entry -c: identical cell scores, identical result"""
    q, r, kind, (go, ge), (mn, mx), kw, want = vec
    table = AA["table"].astype(np.int16)
    shifted = np.where(table > -128, table - shift, table).astype(np.int8)
    cq, cr = aa(q), aa(r)
    zero = np.zeros_like(AA["table"])
    base = ba_align_3di(cq, cq, cr, cr, AA["table"], zero, 32, 26, go, ge, mn, mx, **kw)
    got = ba_align_3di(cq, cq, cr, cr, shifted, zero, 32, 26, go, ge, mn, mx, q_bias=np.full(len(cq), shift, np.int16), **kw)
    assert got == base and base[:len(want)] == want
    # the same shift split between the two tables and both biases
    half = np.where(table > -128, table - 1, table).astype(np.int8)
    z1 = np.full_like(AA["table"], 1 - shift)
    got2 = ba_align_3di(cq, cq, cr, cr, half, z1, 32, 26, go, ge, mn, mx, q_bias=np.full(len(cq), shift - 1, np.int16),
                        r_bias=np.full(len(cr), 1, np.int16), **kw)
    assert got2[:3] == base[:3]
