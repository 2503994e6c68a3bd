"""CPU tests: the oracle's LDDT restatement against values computed by the reference's own LDDTCalculator
(tests/golden/ref_struct.npz, generator tests/golden/make_golden_struct.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(HERE, "golden", "ref_struct.npz"))
    def split(flat, lens):
        off = np.concatenate([[0], np.cumsum(3 * lens)])
        return [flat[off[i]:off[i + 1]].reshape(3, -1) for i in range(len(lens))]
    qca = dict(zip(g["q_keys"].tolist(), split(g["q_ca"], g["q_len"])))
    tca = dict(zip(g["t_keys"].tolist(), split(g["t_ca"], g["t_len"])))
    return g, qca, tca


def test_lddt_matches_reference_bit_exact(gold):
    g, qca, tca = gold
    assert len(g["rows"]) >= 20
    for r, bt, ref in zip(g["rows"], g["bt"], g["lddt"]):
        got = orc.lddt(qca[int(r[0])], tca[int(r[1])], int(r[3]), int(r[5]), str(bt))
        assert got == ref, (r, got, ref)


def test_tmscore_matches_reference_bit_exact(gold):
    """TM-score of the reference's TMaligner (approximate mode). The restatement issues the float operations of the reference
    binary (fused where gcc fused the vector intrinsics, unfused where a permute/hadd separates product and sum), so the values
    are bit-identical (north star tolerance: 1e-5)."""
    g, qca, tca = gold
    for r, bt, ref in zip(g["rows"], g["bt"], g["tm"]):
        norm = min(int(r[4]) - int(r[3]), int(r[6]) - int(r[5]))      # structurealign.cpp:391, TMSCORE_THRESHOLD_MODE_ALIGNMENT
        got = orc.tmscore(qca[int(r[0])], tca[int(r[1])], int(r[3]), int(r[5]), str(bt), norm)
        assert got == ref, (r, got, ref)


def test_structure_bits_score(gold):
    """res.score * sqrt(lddt * tm) truncated to int (structurealign.cpp:406-410)."""
    g, qca, tca = gold
    for r, bt in zip(g["rows"], g["bt"]):
        norm = min(int(r[4]) - int(r[3]), int(r[6]) - int(r[5]))
        tm = orc.tmscore(qca[int(r[0])], tca[int(r[1])], int(r[3]), int(r[5]), str(bt), norm)
        ld = orc.lddt(qca[int(r[0])], tca[int(r[1])], int(r[3]), int(r[5]), str(bt))
        assert int(int(r[2]) * np.sqrt(ld * tm)) == int(r[7]), r
