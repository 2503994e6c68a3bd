"""Parity of the CUDA stage-1 path (k-mer prefilter, through the C ABI) against the oracle: identical hit lists."""
import os

import numpy as np
import pytest

from foldseek_b200 import api, synth
from oracle.oracle import PrefilterIndex

pytestmark = pytest.mark.gpu


def ref_bins(n_targets):
    """BINSIZE the reference would pick on this host (QueryMatcher.cpp:460-487, Util.cpp:346-361)."""
    try:
        l2 = os.sysconf("SC_LEVEL2_CACHE_SIZE")
    except (ValueError, OSError):
        l2 = 0
    if l2 <= 0:
        l2 = 262144
    b = 2
    while b <= 1024:
        if n_targets // b < l2:
            return b
        b *= 2
    return 2048


def _check(ctx, oracle, db, q, max_seqs=1000):
    ctx.load_target(db.aa, db.tdi)
    ctx.set_queries(q.aa, q.tdi)
    got = ctx.prefilter()
    ix = PrefilterIndex(oracle, db.tdi)
    bins = ref_bins(len(db))
    tot = 0
    for qi in range(len(q)):
        ref, st = ix.query(q.tdi[qi], max_seqs=max_seqs, bins=bins)
        g = got[qi]
        assert len(g) == len(ref), (qi, len(g), len(ref))
        assert (g["target"] == ref[:, 0]).all(), qi
        assert (g["score"] == ref[:, 1]).all(), qi
        assert (g["diagonal"] == ref[:, 2]).all(), qi
        tot += len(g)
    return tot


def test_small_db(ctx, oracle):
    db = synth.gen_db(3000, 350, 51, min_len=40)
    q = synth.gen_queries(db, 24, 350, 52, planted_frac=0.3)
    q.aa[3] = q.aa[3][:9]; q.tdi[3] = q.tdi[3][:9]          # shorter than one k-mer window: no hits
    q.tdi[4] = q.tdi[4].copy(); q.tdi[4][10:40] = 20        # X run in the query: windows skipped
    assert _check(ctx, oracle, db, q) > 100


def test_low_complexity_targets_masked(ctx, oracle):
    rng = np.random.default_rng(7)
    db = synth.gen_db(1500, 300, 53, min_len=30)
    for t in range(0, 1500, 7):      # homopolymer runs > 6 are masked in the index and in the diagonal rescoring
        s = db.tdi[t].copy(); p = int(rng.integers(0, max(1, len(s) - 12))); s[p:p + int(rng.integers(5, 12))] = s[p]; db.tdi[t] = s
    q = synth.gen_queries(db, 16, 300, 54, planted_frac=1.0)
    _check(ctx, oracle, db, q)


def test_tie_cut_at_max_seqs(ctx, oracle):
    """more candidates than --max-seqs at the threshold score: the survivors follow the reference's bin/array order"""
    db = synth.gen_db(60000, 350, 55)
    q = synth.gen_queries(db, 8, 350, 56, planted_frac=0.25)
    tot = _check(ctx, oracle, db, q)
    assert tot == 8 * 1000


def test_overflow_chunk_merging(ctx, oracle):
    """> 2*max(1e6,N) index hits per query: the reference flushes chunks and merges them (QueryMatcher.cpp:305-372);
    also exercises the global-memory diagonal state (more targets than fit in shared memory)"""
    rng = np.random.default_rng(5)
    nt = 230000
    tss = synth.random_codes(rng, nt, 350, synth.BG_3DI)
    nq = 20      # 20 x ~75k double hits exceed the double-hit buffer of one pass: also covers the halve-and-retry path
    qss = synth.random_codes(rng, nq, 350, synth.BG_3DI)
    for k in range(2):
        s = tss[1000 * k + 7].copy(); m = rng.random(350) < 0.3; s[m] = qss[k][m]; qss[k] = s
    db = synth.SynthSet([tss[i] for i in range(nt)], [tss[i] for i in range(nt)], None)
    q = synth.SynthSet([qss[i] for i in range(nq)], [qss[i] for i in range(nq)], None)
    _check(ctx, oracle, db, q)
    assert int(ctx.prefilter_stats()[5]) >= nq     # every query flushed at least once


def test_rescore_hits_path_matches_reference(ctx):
    """> max-seqs targets saturate the 8-bit diagonal score: rescoreHits (QueryMatcher.cpp:186-199, 563-589), against the hit
    list the reference binary produced (tests/golden/ref_rescore.npz, generator tests/golden/make_golden_rescore.py)."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_rescore", os.path.join(here, "golden", "make_golden_rescore.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    db, q = mod.make_case()
    ref = np.load(os.path.join(here, "golden", "ref_rescore.npz"))["hits"]
    ctx.load_target(db.aa, db.tdi)
    ctx.set_queries(q.aa, q.tdi)
    g = ctx.prefilter()[0]
    assert len(g) == len(ref)
    assert (g["target"] == ref[:, 0]).all() and (g["score"] == ref[:, 1]).all() and (g["diagonal"] == ref[:, 2]).all()


def test_identity_self_hits_all_vs_all(ctx, oracle):
    """query DB == target DB: the query's own entry comes first with score 65535 and diagonal 0 and is skipped in the hit list
    (QueryMatcher.cpp:409-421, Prefiltering.cpp:856-868); also a family larger than max-seqs so that the self hit meets rescoreHits"""
    rng = np.random.default_rng(21)
    db = synth.gen_db(4000, 260, 61, min_len=50)
    base = db.tdi[0]
    for t in range(1, 1300):       # 1300 near-copies of target 0: its query saturates more than max-seqs diagonal scores
        s = base.copy(); m = rng.random(len(s)) < rng.uniform(0.02, 0.25)
        if m.sum():
            s[m] = synth.random_codes(rng, 1, int(m.sum()), synth.BG_3DI)[0]
        db.tdi[t] = s; db.aa[t] = db.aa[0].copy()
    ids = np.array([0, 5, 17, 1500, 2222, 3999, 7, 3000], np.uint32)
    q = synth.SynthSet([db.aa[i] for i in ids], [db.tdi[i] for i in ids], None)
    ctx.load_target(db.aa, db.tdi)
    ctx.set_queries(q.aa, q.tdi)
    ident = ids.copy(); ident[6] = 0xFFFFFFFF      # one query without identity for contrast
    ctx.set_identity(ident)
    got = ctx.prefilter()
    ix = PrefilterIndex(oracle, db.tdi)
    bins = ref_bins(len(db))
    for qi in range(len(q)):
        ref, _ = ix.query(q.tdi[qi], max_seqs=1000, bins=bins, identity=int(ident[qi]))
        g = got[qi]
        assert len(g) == len(ref), (qi, len(g), len(ref))
        assert (g["target"] == ref[:, 0]).all() and (g["score"] == ref[:, 1]).all() and (g["diagonal"] == ref[:, 2]).all(), qi
        if ident[qi] != 0xFFFFFFFF:
            assert g["target"][0] == ident[qi] and g["score"][0] == 65535 and g["diagonal"][0] == 0
    assert got[0]["score"][1] > 255     # rescaled scores: the rescoreHits path was taken for the family query
    ctx.set_identity(None)


def test_long_query_more_than_30_chunks(ctx, oracle):
    """a 9000-residue query against 230k targets: ~70M index hits = more than 2^26 stream positions and ~35 flushed chunks
    (the limits of the first implementation); the candidate order inside the cut uses chunk-relative positions"""
    rng = np.random.default_rng(9)
    nt = 230000
    tss = synth.random_codes(rng, nt, 350, synth.BG_3DI)
    q0 = synth.random_codes(rng, 1, 9000, synth.BG_3DI)[0]
    q1 = synth.random_codes(rng, 1, 2600, synth.BG_3DI)[0]
    for k in range(6):       # a few targets share segments with the long query
        s = tss[5000 * k + 11].copy(); p = int(rng.integers(0, 9000 - 350)); m = rng.random(350) < 0.6; s[m] = q0[p:p + 350][m]; tss[5000 * k + 11] = s
    db = synth.SynthSet([tss[i] for i in range(nt)], [tss[i] for i in range(nt)], None)
    q = synth.SynthSet([q0, q1], [q0, q1], None)
    _check(ctx, oracle, db, q)
    st = ctx.prefilter_stats()
    assert int(st[1]) > (1 << 26) and int(st[5]) > 30      # index hits of the two queries, flushes
