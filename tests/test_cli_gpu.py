"""The C++ host modules `foldseek-b200 prefilter` / `structurealign` write the same result DB entries as the reference modules
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


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="reference driver not present")
def test_structurealign_module_matches_reference_module(tmp_path):
    """--sort-by-structure-bits 0 -a 1: every output line (scores, E-values, start/end positions, seqId, CIGAR, order) must be
    identical to the reference module's (whose block-aligner is the restated engine, oracle/blockaligner_shim.cpp)."""
    db = synth.gen_db(1500, 260, 93, min_len=40)
    q = synth.gen_queries(db, 24, 260, 94, planted_frac=0.75, sub_prob=0.35, indel_prob=0.06)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    synth.write_foldseek_db(str(tmp_path / "q"), q, "q")
    subprocess.run([CLI, "prefilter", "q_ss", "t_ss", "pref"], cwd=str(tmp_path), check=True, capture_output=True)
    flags = ["-a", "1", "--sort-by-structure-bits", "0", "-e", "10"]
    subprocess.run([REFSEEK, "structurealign", "q", "t", "pref", "aln_ref"] + flags + ["--threads", "4"], cwd=str(tmp_path), check=True,
                   capture_output=True)
    r = subprocess.run([CLI, "structurealign", "q", "t", "pref", "aln_gpu"] + flags, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = synth.read_result_db(str(tmp_path / "aln_ref"))
    got = synth.read_result_db(str(tmp_path / "aln_gpu"))
    assert sorted(ref) == sorted(got)
    n = 0
    for k in ref:
        assert ref[k] == got[k], (k, ref[k][:3], got[k][:3])
        n += len(ref[k])
    assert n >= 10
    assert open(tmp_path / "aln_gpu.dbtype", "rb").read() == open(tmp_path / "aln_ref.dbtype", "rb").read()


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="reference driver not present")
def test_structurealign_default_mode_matches_reference_module(tmp_path):
    """default --sort-by-structure-bits 1: score = int(score * sqrt(LDDT * TM)). Every column must match; the rescored score may
    differ by 1 where the TM-score differs in its last bits (tolerance 1e-5), in which case the order of tied lines may differ."""
    db = synth.gen_db(1500, 260, 93, min_len=40)
    q = synth.gen_queries(db, 24, 260, 94, planted_frac=0.75, sub_prob=0.35, indel_prob=0.06)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    synth.write_foldseek_db(str(tmp_path / "q"), q, "q")
    subprocess.run([CLI, "prefilter", "q_ss", "t_ss", "pref"], cwd=str(tmp_path), check=True, capture_output=True)
    flags = ["-a", "1", "-e", "10"]
    subprocess.run([REFSEEK, "structurealign", "q", "t", "pref", "aln_ref"] + flags + ["--threads", "4"], cwd=str(tmp_path), check=True,
                   capture_output=True)
    r = subprocess.run([CLI, "structurealign", "q", "t", "pref", "aln_gpu"] + flags, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = synth.read_result_db(str(tmp_path / "aln_ref"))
    got = synth.read_result_db(str(tmp_path / "aln_gpu"))
    assert sorted(ref) == sorted(got)
    n = identical = 0
    for k in ref:
        rk = {l[0]: l for l in ref[k]}
        gk = {l[0]: l for l in got[k]}
        assert sorted(rk) == sorted(gk), k
        for t in rk:
            assert rk[t][2:] == gk[t][2:], (k, t, rk[t], gk[t])
            assert abs(int(rk[t][1]) - int(gk[t][1])) <= 1, (k, t, rk[t][1], gk[t][1])
            n += 1
        identical += ref[k] == got[k]
    assert n >= 10 and identical >= 0.9 * len(ref)


def test_structurealign_without_calpha_disables_structure_bits(tmp_path):
    db = synth.gen_db(50, 120, 95, min_len=40)
    q = synth.gen_queries(db, 2, 120, 96, planted_frac=1.0)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    synth.write_foldseek_db(str(tmp_path / "q"), q, "q")
    for f in os.listdir(tmp_path):
        if "_ca" in f:
            os.remove(tmp_path / f)
    subprocess.run([CLI, "prefilter", "q_ss", "t_ss", "pref"], cwd=str(tmp_path), check=True, capture_output=True)
    r = subprocess.run([CLI, "structurealign", "q", "t", "pref", "aln"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0 and "Disabling --sort-by-structure-bits" in r.stderr
    assert len(synth.read_result_db(str(tmp_path / "aln"))) == 2


def test_cli_fails_loudly_on_bad_input(tmp_path):
    r = subprocess.run([CLI, "prefilter", "nope", "nope2", "out"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode != 0 and "Error" in r.stderr
