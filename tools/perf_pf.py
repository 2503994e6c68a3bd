import sys, time, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
rng = np.random.default_rng(0)
taa = synth.random_codes(rng, nt, 350, synth.BG_AA); tss = synth.random_codes(rng, nt, 350, synth.BG_3DI)
qaa = synth.random_codes(rng, nq, 350, synth.BG_AA); qss = synth.random_codes(rng, nq, 350, synth.BG_3DI)
ctx = api.Context(0)
off = (np.arange(nt, dtype=np.uint64) * 350); ln = np.full(nt, 350, np.uint32)
t0 = time.time()
ctx.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * 350)
qoff = (np.arange(nq, dtype=np.uint64) * 350); qln = np.full(nq, 350, np.uint32)
ctx.set_queries_flat(np.ascontiguousarray(qaa.reshape(-1)), np.ascontiguousarray(qss.reshape(-1)), qoff, qln)
t1 = time.time()
print("load %.2fs" % (t1 - t0))
for it in range(3):
    t2 = time.time(); hits = ctx.prefilter(); t3 = time.time()
    st = ctx.prefilter_stats()
    print("prefilter %.3fs | match kernel %.3f ms | kmers %.3e hits %.3e double %.3e cells %.3e | hits/query %.1f" % (
        t3 - t2, st[6] / 1e6, st[0], st[1], st[2], st[3], np.mean([len(h) for h in hits])))
