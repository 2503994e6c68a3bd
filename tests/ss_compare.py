"""TM-score parity of a libfsb200 build against the oracle on the pairs of tests/test_structscore_gpu.py:
   FSB200_LIB=foldseek_b200/libfsb200_ssx.so python tests/ss_compare.py
   prints pairs, bit-identical TM values, largest difference (used to decide the -DSS_EXPLICIT_DOUBLE default)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foldseek_b200 import api, synth
from oracle import oracle as orc

ctx = api.Context(0)
db = synth.gen_db(150, 420, 61, min_len=30)
q = synth.gen_queries(db, 40, 300, 62, planted_frac=1.0, sub_prob=0.35, indel_prob=0.08)
ctx.load_target(db.aa, db.tdi, int(sum(len(a) for a in db.aa)))
ctx.load_target_ca(db.ca)
ctx.set_queries(q.aa, q.tdi)
ctx.set_queries_ca(q.ca)
cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(q))]
off, flat = api.Context.csr(cands)
res = ctx.align(off, flat)
meta, bt = ctx.backtrace(len(res))
sc = ctx.structure_scores(len(res))
n = exact = 0
worst = 0.0
for qi in range(len(q)):
    for k in range(int(off[qi]), int(off[qi + 1])):
        m = meta[k]
        if m["status"] != 1 or m["aln_len"] == 0:
            continue
        s = bytes(bt[int(m["bt_off"]):int(m["bt_off"]) + int(m["aln_len"])]).decode()
        if "M" not in s:
            continue
        t = int(flat[k])
        norm = min(int(res[k]["q_end"]) - int(m["q_start"]), int(res[k]["db_end"]) - int(m["db_start"]))
        tm = orc.tmscore(q.ca[qi], db.ca[t], m["q_start"], m["db_start"], s, norm)
        n += 1
        exact += tm == sc[k]["tm"]
        worst = max(worst, abs(tm - float(sc[k]["tm"])))
print("lib %s: %d pairs, %d TM values bit-identical to the oracle, largest difference %.3g" % (os.path.basename(api.LIB_PATH), n, exact, worst))
