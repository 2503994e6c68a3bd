"""Times the traceback and structure-score kernels on a batch with many accepted pairs (planted queries)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
db = synth.gen_db(nt, 350, 7, min_len=200)
q = synth.gen_queries(db, nq, 350, 8, planted_frac=1.0, sub_prob=0.3, indel_prob=0.05)
ctx = api.Context(0)
ctx.load_target(db.aa, db.tdi, int(sum(len(a) for a in db.aa)))
ctx.load_target_ca(db.ca)
ctx.set_queries(q.aa, q.tdi)
ctx.set_queries_ca(q.ca)
cands = [np.arange(nt, dtype=np.uint32) for _ in range(nq)]
off, flat = api.Context.csr(cands)
res = ctx.align(off, flat)
for it in range(3):
    t0 = time.time(); meta, bt = ctx.backtrace(len(res)); t1 = time.time()
    sc = ctx.structure_scores(len(res)); t2 = time.time()
    bs, ss = ctx.backtrace_stats(), ctx.structure_scores_stats()
    print("accepted %d | backtrace %.1f ms (kernel %.2f ms, %d launches, %d redone, %d mismatch) | scores %.1f ms (kernel %.2f ms, %d degenerate) | mean tm %.3f lddt %.3f" % (
        bs[0], 1e3 * (t1 - t0), bs[2] / 1e6, bs[1], bs[3], bs[4], 1e3 * (t2 - t1), ss[2] / 1e6, ss[3],
        float(np.mean(sc["tm"][sc["status"] == 1])), float(np.nanmean(sc["lddt"][sc["status"] == 1]))))
