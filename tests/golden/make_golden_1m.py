#!/usr/bin/env python3
"""Generates tests/golden/ref_1m_search.npz by running the REFERENCE ITSELF (oracle/_ref/refseek) on the headline-size case:
32 queries (len 350, 4 planted) against a 1M-structure synthetic DB (BASELINE.json configs[2] inputs, bench.py make_workload
seeds).  Pinned: the prefilter hit lists (6-7 overflow chunks per query at this size) and what alignStructure computes before the
traceback for every (query, hit) pair (swdump).  Run in the build container:  python tests/golden/make_golden_1m.py
(needs /root/reference-built oracle/_ref/refseek, ~6 GB RAM, several minutes)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from foldseek_b200 import synth  # noqa: E402

REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")
NT, NQ, LEN = 1000000, 32, 350


def make_case(nt=NT, nq=NQ):
    rng = np.random.default_rng(20240922)
    taa = synth.random_codes(rng, nt, LEN, synth.BG_AA)
    tss = synth.random_codes(rng, nt, LEN, synth.BG_3DI)
    qrng = np.random.default_rng(4242)
    qaa = synth.random_codes(qrng, nq, LEN, synth.BG_AA)
    qss = synth.random_codes(qrng, nq, LEN, synth.BG_3DI)
    src = qrng.integers(0, nt, size=4)
    for k in range(4):
        a, s = taa[src[k]].copy(), tss[src[k]].copy()
        m = qrng.random(LEN) < 0.3; a[m] = qaa[k][m]
        m = qrng.random(LEN) < 0.3; s[m] = qss[k][m]
        qaa[k], qss[k] = a, s
    return taa, tss, qaa, qss


def main():
    taa, tss, qaa, qss = make_case()
    d = tempfile.mkdtemp(prefix="golden_1m_", dir=os.environ.get("TMPDIR", "/tmp"))
    synth.write_seq_db_fixed(os.path.join(d, "t"), taa)
    synth.write_seq_db_fixed(os.path.join(d, "t_ss"), tss)
    synth.write_seq_db_fixed(os.path.join(d, "q"), qaa)
    synth.write_seq_db_fixed(os.path.join(d, "q_ss"), qss)
    run = lambda *a: subprocess.run([REFSEEK] + list(a), cwd=d, check=True, capture_output=True, text=True).stdout
    thr = str(os.cpu_count())
    run("prefilter", "q_ss", "t_ss", "pref", "-s", "9.5", "--max-seqs", "1000", "--comp-bias-corr-scale", "0.15", "--threads", thr)
    pref = synth.read_result_db(os.path.join(d, "pref"))
    sw = run("swdump", "q", "t", "pref").split("\n")
    dbres = int(sw[0].split()[1])
    ml, pairs = {}, []
    for l in sw[1:]:
        f = l.split()
        if not f:
            continue
        if f[0] == "Q":
            ml[int(f[1])] = (float(f[3]), float(f[4]))
        elif f[0] == "P":
            pairs.append([int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6])] + [float(f[7]), float(f[8])])
    out = {
        "db_residues": np.int64(dbres),
        "mu_lambda": np.array([ml.get(i, (0.0, 0.0)) for i in range(NQ)], np.float64),
        "pref_query": np.array([k for k in sorted(pref) for _ in pref[k]], np.int32),
        "pref_hits": np.array([[int(x) for x in h] for k in sorted(pref) for h in pref[k]], np.int32).reshape(-1, 3),
        "sw_int": np.array([p[:6] for p in pairs], np.int32), "sw_eval": np.array([p[6:] for p in pairs], np.float64),
    }
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_1m_search.npz"), **out)
    print("wrote ref_1m_search.npz: %d prefilter hits, %d pairs" % (len(out["pref_hits"]), len(pairs)))
    import shutil
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
