import sys, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
from oracle.oracle import Oracle
db = synth.gen_db(96, 420, 21, min_len=25)
qs = synth.gen_db(40, 384, 22, min_len=17)
sel = [11]
qa = [qs.aa[i] for i in sel]; qt = [qs.tdi[i] for i in sel]
o = Oracle(); ctx = api.Context(0)
dbres = int(sum(len(a) for a in db.aa))
ctx.load_target(db.aa, db.tdi, dbres); ctx.set_queries(qa, qt)
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cands = [np.arange(nt, dtype=np.uint32) for _ in sel]
off, flat = api.Context.csr(cands)
res = ctx.align(off, flat)
for k, qi in enumerate(sel):
    ref, ev = o.align_query(dbres, qa[k], qt[k], db.aa[:nt], db.tdi[:nt])
    got = res[int(off[k]):int(off[k+1])]
    print("L", len(qa[k]), "got", got["fwd_score"], "ref", ref[:, 0], "tlen", [len(a) for a in db.aa[:nt]])
