"""Lanes (fsb200_create_lane): contexts on one GPU that share the parent's staged target, C-alpha and k-mer index.  The query batch split
over two lanes, each driven by its own host thread at the same time, must give exactly what one context gives for the whole batch:
hit lists, alignments, backtraces and structure scores (queries are independent units: Prefiltering.cpp:820-900,
structurealign.cpp:284-452)."""
import threading

import numpy as np
import pytest

from foldseek_b200 import api, synth

pytestmark = pytest.mark.gpu


def _ca(rng, lens):
    return [synth.random_walk_ca(rng, 1, int(n))[0] for n in lens]


def _run(c, q, qca):
    c.set_queries(q.aa, q.tdi)
    c.set_queries_ca(qca)
    hits, counts, aln = c.search()
    n = int(counts.sum())
    meta, bt = c.backtrace(n)
    ss = c.structure_scores(n)
    return hits.copy(), counts.copy(), aln.copy(), meta.copy(), bytes(bt), ss.copy()


def test_two_lanes_equal_one_context():
    db = synth.gen_db(6000, 180, 11, min_len=60)
    q = synth.gen_queries(db, 24, 180, 12, planted_frac=0.5)
    rng = np.random.default_rng(5)
    tca = _ca(rng, [len(a) for a in db.aa])
    qca = _ca(rng, [len(a) for a in q.aa])
    parent = api.Context(0, max_seqs=300)
    try:
        parent.load_target(db.aa, db.tdi)
        parent.load_target_ca(tca)
        whole = _run(parent, q, qca)
        lane = parent.create_lane()
        parent.set_host_threads(2)
        lane.set_host_threads(2)
        with pytest.raises(RuntimeError):          # the lanes borrow the parent's arrays: no reload underneath them
            parent.load_target(db.aa, db.tdi)
        half = len(q) // 2

        class Part:
            def __init__(self, b, e):
                self.aa, self.tdi = q.aa[b:e], q.tdi[b:e]
        parts = [(parent, Part(0, half), qca[:half]), (lane, Part(half, len(q)), qca[half:])]
        out, errs = [None, None], []

        def work(k):
            try:
                for _ in range(2):                   # twice: scratch of the first pass is reused
                    out[k] = _run(*parts[k])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        hits = np.concatenate([out[0][0], out[1][0]]); counts = np.concatenate([out[0][1], out[1][1]]); aln = np.concatenate([out[0][2], out[1][2]])
        assert (counts == whole[1]).all()
        m = np.arange(hits.shape[1])[None, :] < counts[:, None]
        for f in ("target", "score", "diagonal"):
            assert (hits[f][m] == whole[0][f][m]).all(), f
        for f in ("fwd_score", "rev_score", "q_end", "db_end", "status", "score", "evalue"):
            assert (aln[f][m] == whole[2][f][m]).all(), f
        assert int((aln["status"][m] == 0).sum()) >= 6
        # backtraces / structure scores are dense over the candidate pairs of each call, in query order
        meta = np.concatenate([out[0][3], out[1][3]]); ss = np.concatenate([out[0][5], out[1][5]])
        wm, ws = whole[3], whole[5]
        assert len(meta) == len(wm)
        for f in ("q_start", "db_start", "identical_aa", "aln_len", "status", "block_score"):
            assert (meta[f] == wm[f]).all(), f
        for f in ("tm", "lddt", "status", "score_bits"):
            assert (ss[f] == ws[f]).all(), f
        # the strings themselves: offsets restart per call
        for k, (mm, btb) in enumerate(((out[0][3], out[0][4]), (out[1][3], out[1][4]))):
            base = 0 if k == 0 else len(out[0][3])
            for i in np.nonzero(mm["status"] == 1)[0][:50]:
                a = btb[int(mm["bt_off"][i]):int(mm["bt_off"][i]) + int(mm["aln_len"][i])]
                w = whole[4][int(wm["bt_off"][base + i]):int(wm["bt_off"][base + i]) + int(wm["aln_len"][base + i])]
                assert a == w
    finally:
        parent.close()
