import sys, time, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
rng = np.random.default_rng(0)
t0 = time.time()
taa = synth.random_codes(rng, nt, 350, synth.BG_AA); tss = synth.random_codes(rng, nt, 350, synth.BG_3DI)
qaa = synth.random_codes(rng, nq, 350, synth.BG_AA); qss = synth.random_codes(rng, nq, 350, synth.BG_3DI)
ctx = api.Context(0)
off = (np.arange(nt, dtype=np.uint64) * 350); ln = np.full(nt, 350, np.uint32)
ctx.load_target_flat(np.ascontiguousarray(taa.reshape(-1)), np.ascontiguousarray(tss.reshape(-1)), off, ln, nt * 350)
qoff = (np.arange(nq, dtype=np.uint64) * 350); qln = np.full(nq, 350, np.uint32)
t1 = time.time()
ctx.set_queries_flat(np.ascontiguousarray(qaa.reshape(-1)), np.ascontiguousarray(qss.reshape(-1)), qoff, qln)
t2 = time.time()
cand = rng.integers(0, nt, size=(nq, nc)).astype(np.uint32)
coff = (np.arange(nq + 1, dtype=np.uint64) * nc)
print("gen %.2fs set_queries %.3fs" % (t1 - t0, t2 - t1))
for it in range(3):
    t3 = time.time(); ctx.align_upload(coff, np.ascontiguousarray(cand.reshape(-1))); t4 = time.time()
    ctx.align_run(); t5 = time.time(); res = ctx.align_download(); t6 = time.time()
    st = ctx.align_stats()
    cells = int(st[1] + st[2] + st[3])
    print("upload %.3f run %.3f download %.3f | kernel %.3f ms launches %d | cells %.3e -> %.1f GCUPS(kernel) %.1f GCUPS(run) survivors %d" % (
        t4 - t3, t5 - t4, t6 - t5, st[6] / 1e6, st[7], cells, cells / (st[6] / 1e9) / 1e9, cells / (t5 - t4) / 1e9, int((res["status"] != 1).sum())))
