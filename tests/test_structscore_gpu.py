"""Parity of the CUDA LDDT / TM-score kernels (fsb200_structure_scores_*) against the oracle restatements, which are pinned to
the reference's own LDDTCalculator / TMaligner (tests/test_struct_oracle.py). LDDT: bit-exact. TM-score: 1e-5 (north star tolerance).
The float path is issued in the reference's operation order and the double eigen-solver spells out the reference binary's fused
operations (structscore.cuh, SS_EXPLICIT_DOUBLE); what can still differ is CUDA's cos / atan2 against glibc's.  Measured on a B200
(tests/ss_compare.py): 93 of 93 values bit-identical; the test demands 90 % bit-identical and all within 1e-5."""
import numpy as np
import pytest

from foldseek_b200 import api, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_lddt_tm_and_structure_bits(ctx, oracle):
    db = synth.gen_db(150, 420, 61, min_len=30)
    q = synth.gen_queries(db, 40, 300, 62, planted_frac=1.0, sub_prob=0.35, indel_prob=0.08)
    dbres = int(sum(len(a) for a in db.aa))
    ctx.load_target(db.aa, db.tdi, dbres)
    ctx.load_target_ca(db.ca)
    ctx.set_queries(q.aa, q.tdi)
    ctx.set_queries_ca(q.ca)
    cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
    off, flat = api.Context.csr(cands)
    res = ctx.align(off, flat)
    meta, bt = ctx.backtrace(len(res))
    sc = ctx.structure_scores(len(res))
    n = exact_tm = 0
    for qi in range(len(q)):
        for k in range(int(off[qi]), int(off[qi + 1])):
            m = meta[k]
            if m["status"] != 1 or m["aln_len"] == 0:
                assert sc[k]["status"] == -1
                continue
            t = int(flat[k])
            s = bytes(bt[int(m["bt_off"]):int(m["bt_off"]) + int(m["aln_len"])]).decode()
            if "M" not in s:
                continue
            ld = orc.lddt(q.ca[qi], db.ca[t], m["q_start"], m["db_start"], s)
            norm = min(int(res[k]["q_end"]) - int(m["q_start"]), int(res[k]["db_end"]) - int(m["db_start"]))
            tm = orc.tmscore(q.ca[qi], db.ca[t], m["q_start"], m["db_start"], s, norm)
            assert sc[k]["status"] == 1, (qi, t)
            assert np.float32(ld) == sc[k]["lddt"] or (np.isnan(ld) and np.isnan(sc[k]["lddt"])), (qi, t, ld, sc[k]["lddt"])
            assert abs(tm - sc[k]["tm"]) <= 1e-5, (qi, t, tm, sc[k]["tm"])
            exact_tm += tm == sc[k]["tm"]
            if abs(tm - sc[k]["tm"]) == 0 and not np.isnan(ld):
                assert sc[k]["score_bits"] == int(int(res[k]["score"]) * np.sqrt(float(np.float32(ld)) * tm)), (qi, t)
            n += 1
    assert n >= 40
    assert exact_tm >= 0.9 * n
