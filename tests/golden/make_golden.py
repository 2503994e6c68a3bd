#!/usr/bin/env python3
"""Generates tests/golden/ref_small.npz by running the REFERENCE ITSELF (oracle/_ref/refseek = the reference's own
sources compiled by oracle/Makefile) on a small seeded synthetic DB.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Outputs pinned: the four substitution matrices (matdump), the prefilter hit lists (prefilter module), and for every
(query, hit) pair what alignStructure computes before the traceback (swdump): forward score, qEnd, dbEnd, reverse
score, both E-values, plus (mu, lambda) per query.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from foldseek_b200 import synth  # noqa: E402

REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")


def main():
    db = synth.gen_db(2500, 220, 9001, min_len=30)
    q = synth.gen_queries(db, 16, 220, 9002, planted_frac=0.5)
    q.aa[5] = q.aa[5][:24]; q.tdi[5] = q.tdi[5][:24]; q.ca[5] = q.ca[5][:, :24]
    d = tempfile.mkdtemp(prefix="golden_")
    synth.write_foldseek_db(os.path.join(d, "t"), db, "t")
    synth.write_foldseek_db(os.path.join(d, "q"), q, "q")
    run = lambda *a: subprocess.run([REFSEEK] + list(a), cwd=d, check=True, capture_output=True, text=True).stdout
    mat = run("matdump").split("\n")
    mats = {}
    i = 0
    while i < len(mat):
        if mat[i].startswith("MAT"):
            mats[mat[i].split()[1]] = np.array([[int(x) for x in mat[i + 1 + r].split()] for r in range(21)], np.int16)
            i += 23
        else:
            i += 1
    run("prefilter", "q_ss", "t_ss", "pref", "-s", "9.5", "--max-seqs", "1000", "--comp-bias-corr-scale", "0.15", "--threads", "1")
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
        "t_len": np.array([len(a) for a in db.aa], np.int32), "t_aa": np.concatenate(db.aa), "t_3di": np.concatenate(db.tdi),
        "q_len": np.array([len(a) for a in q.aa], np.int32), "q_aa": np.concatenate(q.aa), "q_3di": np.concatenate(q.tdi),
        "mu_lambda": np.array([ml.get(i, (0.0, 0.0)) for i in range(len(q))], np.float64),
        "pref_query": np.array([k for k in sorted(pref) for _ in pref[k]], np.int32),
        "pref_hits": np.array([[int(x) for x in h] for k in sorted(pref) for h in pref[k]], np.int32).reshape(-1, 3),
        "sw_int": np.array([p[:6] for p in pairs], np.int64), "sw_eval": np.array([p[6:] for p in pairs], np.float64),
    }
    for k, v in mats.items():
        out["mat_" + k] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_small.npz"), **out)
    print("wrote ref_small.npz: %d prefilter hits, %d pairs" % (len(out["pref_hits"]), len(pairs)))


if __name__ == "__main__":
    main()
