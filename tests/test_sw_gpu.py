"""Parity of the CUDA stage-2 path (through the C ABI) against the oracle: bit-exact integer results."""
import numpy as np
import pytest

from foldseek_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, q, db, cands, dbres=None):
    dbres = int(sum(len(a) for a in db.aa)) if dbres is None else dbres
    ctx.load_target(db.aa, db.tdi, dbres)
    ctx.set_queries(q.aa, q.tdi)
    off, flat = api.Context.csr(cands)
    res = ctx.align(off, flat)
    n_pass = 0
    for qi in range(len(q)):
        ts = cands[qi]
        ref, ev = oracle.align_query(dbres, q.aa[qi], q.tdi[qi], [db.aa[t] for t in ts], [db.tdi[t] for t in ts])
        got = res[int(off[qi]):int(off[qi + 1])]
        ml, ba, bs = ctx.query_info(qi)
        rml = oracle.mu_lambda(dbres, q.tdi[qi])
        assert tuple(ml) == rml
        oba, obs = oracle.query_bias(q.aa[qi], q.tdi[qi])
        assert (ba == oba).all() and (bs == obs).all()
        assert (got["fwd_score"] == ref[:, 0]).all(), (qi, got["fwd_score"][:8], ref[:8, 0])
        assert (got["db_end"] == ref[:, 2]).all(), qi
        thr = 10.0
        st = np.where(ev[:, 0] > thr, 1, np.where(ev[:, 1] > thr, 2, 0))
        assert (got["status"] == st).all(), qi
        m = st != 1
        assert (got["rev_score"][m] == ref[m, 3]).all(), qi
        assert (got["q_end"][m] == ref[m, 1]).all(), (qi, got["q_end"][m][:8], ref[m, 1][:8])
        assert np.array_equal(got["evalue"][m], np.where(st[m] == 0, ev[m, 1], ev[m, 1])), qi
        assert np.array_equal(got["evalue"][~m], ev[~m, 0]), qi
        n_pass += int(m.sum())
    return n_pass


def test_random_and_planted_all_row_classes(ctx, oracle):
    """query lengths sweep every rows-per-thread class of the packed kernel, odd and even stripe lengths"""
    db = synth.gen_db(96, 420, 21, min_len=25)
    qs = synth.gen_db(40, 384, 22, min_len=17)
    # make some queries mutated copies of targets so that both E-value gates are passed
    planted = synth.gen_queries(db, 24, 300, 23, planted_frac=1.0)
    planted.aa = [a[:384] for a in planted.aa]; planted.tdi = [a[:384] for a in planted.tdi]
    q = synth.SynthSet(qs.aa + planted.aa, qs.tdi + planted.tdi, qs.ca + planted.ca)
    cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
    n_pass = _check(ctx, oracle, q, db, cands)
    assert n_pass > 20


def test_long_queries_use_generic_kernel(ctx, oracle):
    db = synth.gen_db(24, 900, 31, min_len=100)
    q = synth.gen_queries(db, 6, 900, 32, planted_frac=0.5)
    q.aa = [a for a in q.aa]
    cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
    _check(ctx, oracle, q, db, cands)


def test_low_complexity_exercises_gap_quirk(ctx, oracle):
    """repeats + shifted copies: many equal-score paths, gaps right after gaps (E opened from Hp, SSW.cpp:1230-1259)"""
    rng = np.random.default_rng(5)
    def lowc(n):
        unit = rng.integers(0, 20, size=rng.integers(2, 6))
        s = np.tile(unit, n // len(unit) + 1)[:n].astype(np.uint8)
        mut = rng.random(n) < 0.15
        s[mut] = rng.integers(0, 20, size=int(mut.sum()))
        return s
    taa = [lowc(int(rng.integers(40, 380))) for _ in range(64)]
    tss = [lowc(len(a)) for a in taa]
    qaa, qss = [], []
    for i in range(24):
        t = int(rng.integers(0, 64))
        a, s = taa[t].copy(), tss[t].copy()
        # delete and insert blocks to force gap-after-gap paths
        cut = int(rng.integers(5, max(6, len(a) // 2)))
        a = np.concatenate([a[:cut], a[cut + 3:], lowc(7)]); s = np.concatenate([s[:cut], s[cut + 3:], lowc(7)])
        qaa.append(a[:384]); qss.append(s[:384])
    db = synth.SynthSet(taa, tss, [None] * 64)
    q = synth.SynthSet(qaa, qss, [None] * 24)
    cands = [np.arange(64, dtype=np.uint32) for _ in range(24)]
    _check(ctx, oracle, q, db, cands)


def test_empty_and_ragged_candidate_lists(ctx, oracle):
    db = synth.gen_db(10, 100, 41, min_len=10)
    q = synth.gen_db(5, 64, 42, min_len=8)
    cands = [np.array([], np.uint32), np.array([3], np.uint32), np.array([9, 0, 9, 2, 2], np.uint32),
             np.array([], np.uint32), np.arange(10, dtype=np.uint32)]
    _check(ctx, oracle, q, db, cands)


def test_golden_vectors_from_reference(ctx):
    """committed outputs of the reference itself (tests/golden/make_golden.py) reproduced through the C ABI"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_small.npz"))
    def split(flat, lens):
        off = np.concatenate([[0], np.cumsum(lens)])
        return [flat[off[i]:off[i + 1]] for i in range(len(lens))]
    t_aa, t_ss = split(g["t_aa"], g["t_len"]), split(g["t_3di"], g["t_len"])
    q_aa, q_ss = split(g["q_aa"], g["q_len"]), split(g["q_3di"], g["q_len"])
    ctx.load_target(t_aa, t_ss, int(g["db_residues"]))
    ctx.set_queries(q_aa, q_ss)
    hits = ctx.prefilter()
    for qi in range(len(q_aa)):
        ref = g["pref_hits"][g["pref_query"] == qi]
        assert len(hits[qi]) == len(ref)
        assert (hits[qi]["target"] == ref[:, 0]).all() and (hits[qi]["score"] == ref[:, 1]).all() and (hits[qi]["diagonal"] == ref[:, 2]).all()
    sw = g["sw_int"]
    cands = [sw[sw[:, 0] == qi, 1].astype(np.uint32) for qi in range(len(q_aa))]
    off, flat = api.Context.csr(cands)
    res = ctx.align(off, flat)
    rows = np.concatenate([np.nonzero(sw[:, 0] == qi)[0] for qi in range(len(q_aa))])
    assert (res["fwd_score"] == sw[rows, 2]).all() and (res["db_end"] == sw[rows, 4]).all()
    ev = g["sw_eval"][rows]
    st = np.where(ev[:, 0] > 10.0, 1, np.where(ev[:, 1] > 10.0, 2, 0))
    assert (res["status"] == st).all()
    m = st != 1
    assert (res["q_end"][m] == sw[rows, 3][m]).all() and (res["rev_score"][m] == sw[rows, 5][m]).all()
    assert np.array_equal(res["evalue"][m], ev[m, 1])
