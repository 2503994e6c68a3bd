import sys, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
from oracle.oracle import Oracle
db = synth.gen_db(96, 420, 21, min_len=25)
qs = synth.gen_db(40, 384, 22, min_len=17)
o = Oracle(); ctx = api.Context(0)
dbres = int(sum(len(a) for a in db.aa))
ctx.load_target(db.aa, db.tdi, dbres); ctx.set_queries(qs.aa, qs.tdi)
cands = [np.arange(len(db), dtype=np.uint32) for _ in range(len(qs))]
off, flat = api.Context.csr(cands)
res = ctx.align(off, flat)
for qi in range(len(qs)):
    ref, ev = o.align_query(dbres, qs.aa[qi], qs.tdi[qi], db.aa, db.tdi)
    got = res[int(off[qi]):int(off[qi+1])]
    bad = np.nonzero((got["fwd_score"] != ref[:, 0]) | (got["db_end"] != ref[:, 2]))[0]
    L = len(qs.aa[qi])
    print("q", qi, "L", L, "seg", (L+15)//16, "nbad", len(bad), [(int(t), len(db.aa[t]), int(got["fwd_score"][t]), int(ref[t,0]), int(got["db_end"][t]), int(ref[t,2])) for t in bad[:6]])
