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
