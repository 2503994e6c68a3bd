"""One-off GPU check of the overflow-chunk path at 1M targets (6-7 chunks per query) against hit lists produced by
the REFERENCE binary (tests/golden/ref_1m_overflow.npz, inputs regenerated from the seed used in /tmp/ovf1m.py)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import api, synth
rng = np.random.default_rng(6)
nt = 1000000
tss = synth.random_codes(rng, nt, 350, synth.BG_3DI)
qss = synth.random_codes(rng, 3, 350, synth.BG_3DI)
for k in range(2):
    s = tss[1000 * k + 7].copy(); m = rng.random(350) < 0.3; s[m] = qss[k][m]; qss[k] = s
exp = np.load('tests/golden/ref_1m_overflow.npz')['hits']
ctx = api.Context(0)
off = np.arange(nt, dtype=np.uint64) * 350; ln = np.full(nt, 350, np.uint32)
flat = np.ascontiguousarray(tss.reshape(-1))
t = time.time(); ctx.load_target_flat(flat, flat, off, ln, nt * 350)
ctx.set_queries([qss[i] for i in range(3)], [qss[i] for i in range(3)])
hits = ctx.prefilter(); print("prefilter incl. index build %.1fs" % (time.time() - t), ctx.prefilter_stats())
t = time.time(); hits = ctx.prefilter(); print("prefilter %.3fs" % (time.time() - t))
ok = True
for qi in range(3):
    e = exp[exp[:, 0] == qi][:, 1:]
    g = hits[qi]
    same = len(g) == len(e) and (g["target"] == e[:, 0]).all() and (g["score"] == e[:, 1]).all() and (g["diagonal"] == e[:, 2]).all()
    print(qi, "MATCH" if same else "DIFF", len(g), len(e))
    ok &= same
print("ALL OK" if ok else "MISMATCH")
