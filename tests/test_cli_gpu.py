"""The C++ host module `foldseek-b200 prefilter` writes the same prefilter DB entries as the reference module
(byte-identical per key) on MMseqs2 on-disk databases."""
import os
import subprocess

import numpy as np
import pytest

from foldseek_b200 import synth
from oracle.oracle import REFSEEK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "foldseek_b200", "bin", "foldseek-b200")


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="reference driver not present")
def test_prefilter_module_matches_reference_module(tmp_path):
    db = synth.gen_db(4000, 300, 91, min_len=40)
    q = synth.gen_queries(db, 12, 300, 92, planted_frac=0.5)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    synth.write_foldseek_db(str(tmp_path / "q"), q, "q")
    args = ["q_ss", "t_ss"]
    flags = ["-s", "9.5", "--max-seqs", "1000", "--comp-bias-corr-scale", "0.15"]
    subprocess.run([REFSEEK, "prefilter"] + args + ["pref_ref"] + flags + ["--threads", "4"], cwd=str(tmp_path), check=True, capture_output=True)
    r = subprocess.run([CLI, "prefilter"] + args + ["pref_gpu"] + flags, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = synth.read_result_db(str(tmp_path / "pref_ref"))
    got = synth.read_result_db(str(tmp_path / "pref_gpu"))
    assert sorted(ref) == sorted(got)
    for k in ref:
        assert ref[k] == got[k], k
    assert open(tmp_path / "pref_gpu.dbtype", "rb").read() == open(tmp_path / "pref_ref.dbtype", "rb").read()


def test_structurealign_module_runs_and_reports_scores(tmp_path):
    db = synth.gen_db(500, 200, 93, min_len=40)
    q = synth.gen_queries(db, 6, 200, 94, planted_frac=1.0)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    synth.write_foldseek_db(str(tmp_path / "q"), q, "q")
    subprocess.run([CLI, "prefilter", "q_ss", "t_ss", "pref"], cwd=str(tmp_path), check=True, capture_output=True)
    r = subprocess.run([CLI, "structurealign", "q", "t", "pref", "aln"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    aln = synth.read_result_db(str(tmp_path / "aln"))
    assert len(aln) == 6 and sum(len(v) for v in aln.values()) >= 6      # every planted query finds its source
    assert all(len(line) == 10 for v in aln.values() for line in v)


def test_cli_fails_loudly_on_bad_input(tmp_path):
    r = subprocess.run([CLI, "prefilter", "nope", "nope2", "out"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "Error" in r.stderr
