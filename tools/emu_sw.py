# scalar emulation of sw_packed_kernel's algorithm (one target) to separate logic bugs from codegen bugs
import sys, numpy as np
sys.path.insert(0, '.')
from foldseek_b200 import synth
from oracle.oracle import Oracle
db = synth.gen_db(96, 420, 21, min_len=25)
qs = synth.gen_db(40, 384, 22, min_len=17)
o = Oracle()
qi, ti = 11, 0
qa, qt = qs.aa[qi], qs.tdi[qi]; ta, tt = db.aa[ti], db.tdi[ti]
L = len(qa); Lt = len(ta)
mAA, _ = o.matrix(0); m3, _ = o.matrix(1)
ba, bs = o.query_bias(qa, qt)
seg = (L + 15) // 16; R = (seg + 1) & ~1; odd = seg & 1
def score(X, g, r):
    if X == 441 or r >= seg or g * seg + r >= L: return -128
    a, s = X // 21, X % 21; q = g * seg + r
    return int(np.int8(mAA[a, qa[q]])) + int(ba[q]) + int(np.int8(m3[s, qt[q]])) + int(bs[q])
H = np.zeros((16, R), int); E = np.zeros((16, R), int)
best = np.zeros(16, int); col = np.zeros(16, int)
hl = np.zeros(16, int); fo = np.zeros(16, int); diag0 = np.zeros(16, int); letters = np.zeros(16, int)
for s in range(Lt + 15):
    hin = np.concatenate([[0], hl[:-1]]); fin = np.concatenate([[0], fo[:-1]]); lin = np.concatenate([[0], letters[:-1]])
    lin[0] = (int(ta[s]) * 21 + int(tt[s])) if s < Lt else 441
    letters = lin.copy()
    for g in range(16):
        c = s - g
        if 0 <= c < Lt:
            hd = diag0[g]; fs = 0; ff = fin[g]; cm = 0; ffb = fin[g]
            for r in range(R):
                sc = score(letters[g], g, r)
                if r == R - 1: ffb = ff
                t = max(hd + sc, E[g, r], 0); hp = max(t, fs); h = max(hp, ff); x = hp - 10
                E[g, r] = max(E[g, r] - 1, x); fs = max(fs - 1, x); ff = max(ff - 1, x)
                hd = H[g, r]; H[g, r] = h; cm = max(cm, h)
            hl[g] = H[g, R - 2] if odd else H[g, R - 1]; fo[g] = ffb if odd else ff
            if cm > best[g]: best[g] = cm; col[g] = c
        diag0[g] = hin[g]
S = best.max(); print("emu score", S, "col", min(col[g] for g in range(16) if best[g] == S))
ref, ev = o.align_query(1000000, qa, qt, [ta], [tt]); print("oracle", ref)
