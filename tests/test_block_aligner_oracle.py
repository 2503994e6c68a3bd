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
