"""Generates tests/golden/ref_rescore.npz: a prefilter case in which more than --max-seqs targets saturate the 8-bit diagonal
score, so the reference takes its rescoreHits path (QueryMatcher.cpp:186-199, 563-589). Reference output = oracle/_ref/refseek."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from foldseek_b200 import synth  # noqa: E402

REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")


def make_case():
    rng = np.random.default_rng(11)
    base = synth.gen_db(1, 300, 77)
    aa, tdi, ca = [], [], []
    for _ in range(1600):          # 1600 copies of one structure with 2-30 % of the 3Di letters replaced
        t = base.tdi[0].copy()
        m = rng.random(len(t)) < rng.uniform(0.02, 0.3)
        if m.sum():
            t[m] = synth.random_codes(rng, 1, int(m.sum()), synth.BG_3DI)[0]
        aa.append(base.aa[0].copy()); tdi.append(t); ca.append(base.ca[0])
    rnd = synth.gen_db(400, 300, 78)
    db = synth.SynthSet(aa + rnd.aa, tdi + rnd.tdi, ca + rnd.ca)
    q = synth.SynthSet([base.aa[0]], [base.tdi[0]], [base.ca[0]])
    return db, q


def main():
    db, q = make_case()
    d = tempfile.mkdtemp(prefix="golden_rescore_")
    synth.write_foldseek_db(os.path.join(d, "t"), db, "t")
    synth.write_foldseek_db(os.path.join(d, "q"), q, "q")
    subprocess.run([REFSEEK, "prefilter", "q_ss", "t_ss", "pref", "-s", "9.5", "--max-seqs", "1000", "--comp-bias-corr-scale", "0.15",
                    "--threads", "1"], cwd=d, check=True, capture_output=True)
    ref = synth.read_result_db(os.path.join(d, "pref"))[0]
    hits = np.array([[int(x) for x in h] for h in ref], np.int32)
    assert hits[:, 1].min() > 255          # every reported score is a rescaled one
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_rescore.npz"), hits=hits)
    print("wrote ref_rescore.npz: %d hits, scores %d..%d" % (len(hits), hits[:, 1].min(), hits[:, 1].max()))


if __name__ == "__main__":
    main()
