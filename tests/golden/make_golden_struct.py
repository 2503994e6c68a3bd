"""Generates tests/golden/ref_struct.npz: LDDT and TM-score of reference backtraces, computed by the reference's own
LDDTCalculator / TMaligner (oracle/_ref/refseek sdump, see oracle/refseek_main.cpp) on a small synthetic search.
Only the structures that take part in an alignment are stored. Run in the build container (needs /root/reference built into
oracle/_ref): python tests/golden/make_golden_struct.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from foldseek_b200 import synth  # noqa: E402

REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")


def uncompress(cigar):
    out, num = [], ""
    for c in cigar:
        if c.isdigit():
            num += c
        else:
            out.append(c * int(num)); num = ""
    return "".join(out)


def main():
    db = synth.gen_db(1500, 260, 93, min_len=40)
    q = synth.gen_queries(db, 40, 260, 94, planted_frac=0.75, sub_prob=0.35, indel_prob=0.06)
    d = tempfile.mkdtemp(prefix="golden_struct_")
    synth.write_foldseek_db(os.path.join(d, "t"), db, "t")
    synth.write_foldseek_db(os.path.join(d, "q"), q, "q")
    run = lambda *a: subprocess.run([REFSEEK] + list(a), cwd=d, check=True, capture_output=True, text=True).stdout
    run("prefilter", "q_ss", "t_ss", "pref", "--threads", "4")
    run("structurealign", "q", "t", "pref", "aln", "-a", "1", "--sort-by-structure-bits", "0", "-e", "10", "--threads", "4")
    run("structurealign", "q", "t", "pref", "aln_bits", "-a", "1", "-e", "10", "--threads", "4")
    aln = synth.read_result_db(os.path.join(d, "aln"))
    bits = synth.read_result_db(os.path.join(d, "aln_bits"))
    sd = {}
    for l in run("sdump", "q", "t", "aln").split("\n"):
        f = l.split()
        if f and f[0] == "S":
            sd[(int(f[1]), int(f[2]))] = (float(f[3]), float(f[4]), int(f[5]))
    rows, bts = [], []
    for qk in sorted(aln):
        for h in aln[qk]:
            tk = int(h[0])
            lddt, tm, final = sd[(qk, tk)]
            bit_line = [b for b in bits[qk] if int(b[0]) == tk][0]
            assert int(bit_line[1]) == final
            rows.append([qk, tk, int(h[1]), int(h[4]), int(h[5]), int(h[7]), int(h[8]), final])
            bts.append(uncompress(h[10]))
    rows = np.array(rows, np.int64)
    used_q = sorted(set(rows[:, 0].tolist())); used_t = sorted(set(rows[:, 1].tolist()))
    out = {
        "rows": rows,  # qKey, tKey, score, qStart, qEnd, dbStart, dbEnd, final score of the default (structure-bits) mode
        "lddt": np.array([sd[(int(r[0]), int(r[1]))][0] for r in rows], np.float64),
        "tm": np.array([sd[(int(r[0]), int(r[1]))][1] for r in rows], np.float64),
        "bt": np.array(bts),
        "q_keys": np.array(used_q, np.int64), "t_keys": np.array(used_t, np.int64),
        "q_len": np.array([len(q.aa[k]) for k in used_q], np.int32), "t_len": np.array([len(db.aa[k]) for k in used_t], np.int32),
        "q_ca": np.concatenate([q.ca[k].reshape(-1) for k in used_q]).astype(np.float32),
        "t_ca": np.concatenate([db.ca[k].reshape(-1) for k in used_t]).astype(np.float32),
    }
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_struct.npz"), **out)
    print("wrote ref_struct.npz: %d alignments, %d queries, %d targets" % (len(rows), len(used_q), len(used_t)))


if __name__ == "__main__":
    main()
