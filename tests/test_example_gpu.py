"""BASELINE.json configs[0]: the reference's own example/ proteins.  `foldseek-b200 prefilter | structurealign`, called with the
full parameter strings the search workflow serialises for the two modules (tests/golden/example/params.txt), must write the same
result DB entries as the reference modules did on the same DBs (tests/golden/example/ref_*, produced by the reference's own
createdb / prefilter / structurealign, generator tests/golden/make_golden_example.py).  Real proteins exercise what i.i.d.
synthetic data does not: homologous families, saturated diagonal scores, identity hits (query DB == target DB), ragged lengths,
Coordinate16-encoded C-alpha traces, default --sort-by-structure-bits 1."""
import os
import shutil
import subprocess

import pytest

from foldseek_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "foldseek_b200", "bin", "foldseek-b200")
EX = os.path.join(ROOT, "tests", "golden", "example")


def _params():
    p = {}
    for line in open(os.path.join(EX, "params.txt")):
        mod, _, rest = line.strip().partition(" ")
        p[mod] = rest.split()
    return p


@pytest.fixture()
def work(tmp_path):
    d = tmp_path / "ex"
    shutil.copytree(EX, str(d))
    return d


def _run(work, *args):
    r = subprocess.run([CLI] + list(args), cwd=str(work), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    return r


def _same(work, got, ref):
    g, e = synth.read_result_db(str(work / got)), synth.read_result_db(str(work / ref))
    assert sorted(g) == sorted(e)
    n = 0
    for k in e:
        assert g[k] == e[k], (k, g[k][:4], e[k][:4])
        n += len(e[k])
    assert open(work / (got + ".dbtype"), "rb").read() == open(work / (ref + ".dbtype"), "rb").read()
    return n


def test_d1asha_vs_example_default_flags(work):
    p = _params()
    _run(work, "prefilter", "q_ss", "db_ss", "pref_q", *p["prefilter"])
    assert _same(work, "pref_q", "ref_pref_q") >= 20
    _run(work, "structurealign", "q", "db", "ref_pref_q", "aln_q", *p["structurealign"])
    assert _same(work, "aln_q", "ref_aln_q") >= 8


def test_padded_target_db(work):
    """target DB in the padded layout of `makepaddeddb` (DBTYPE_EXTENDED_GPU: numeric codes, 4-byte padded, re-keyed by length)"""
    p = _params()
    _run(work, "prefilter", "q_ss", "pad_ss", "pref_q_pad", *p["prefilter"])
    assert _same(work, "pref_q_pad", "ref_pref_q_pad") >= 20


def test_all_vs_all_identity_hits(work):
    """query DB == target DB: every query's own entry first with score 65535 (prefilter) and accepted as identity (align)"""
    p = _params()
    _run(work, "prefilter", "db_ss", "db_ss", "pref_all", *p["prefilter"])
    assert _same(work, "pref_all", "ref_pref_all") >= 30 * 20
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_all", *p["structurealign"])
    assert _same(work, "aln_all", "ref_aln_all") >= 100


def test_all_vs_all_two_lanes(work):
    """--lanes 2: two contexts on the GPU share the staged DB (prefilter: also the index), each with its own query range and host
    thread; the result DBs are the ones a single context writes"""
    p = _params()
    _run(work, "prefilter", "db_ss", "db_ss", "pref_all2", "--lanes", "2", "--threads", "4", *p["prefilter"])
    assert _same(work, "pref_all2", "ref_pref_all") >= 30 * 20
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_all2", "--lanes", "2", *p["structurealign"])
    assert _same(work, "aln_all2", "ref_aln_all") >= 100


def test_coverage_and_seqid_criteria(work):
    p = list(_params()["structurealign"])
    for flag, v in (("-c", "0.8"), ("--cov-mode", "0"), ("--min-seq-id", "0.2"), ("-a", "1")):
        p[p.index(flag) + 1] = v
    _run(work, "structurealign", "q", "db", "ref_pref_q", "aln_cov", *p)
    assert _same(work, "aln_cov", "ref_aln_q_cov") >= 1


@pytest.mark.parametrize("mode", ["1", "2"])
def test_coverage_modes_target_and_query(work, mode):
    """-c 0.7 --cov-mode 1 (TARGET) / 2 (QUERY), Parameters.h:284-289: the prefilter filters hits by length ratio only for modes 0, 2
    and 5 (Prefiltering.cpp:886-894, Util::canBeCovered), the aligner applies canBeCovered, hasCoverage after the forward pass and
    checkCriteria on the traced alignment (structurealign.cpp:364-367, :50-53, Alignment.cpp:548-567)"""
    pf, al = list(_params()["prefilter"]), list(_params()["structurealign"])
    for lst in (pf, al):
        lst[lst.index("-c") + 1] = "0.7"
        lst[lst.index("--cov-mode") + 1] = mode
    _run(work, "prefilter", "db_ss", "db_ss", "pref_cov", *pf)
    assert _same(work, "pref_cov", "ref_pref_all_cov" + mode) >= 30 * 20
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_cov", *al)
    assert _same(work, "aln_cov", "ref_aln_all_cov" + mode) >= 100


def test_unimplemented_values_fail_loudly(work):
    """a result-changing option the path does not implement must end the run with its name, never be swallowed"""
    for flags in (["--score-bias", "1"], ["--exact-tmscore", "1"], ["--no-such-flag", "1"]):
        r = subprocess.run([CLI, "structurealign", "q", "db", "ref_pref_q", "aln_x"] + flags, cwd=str(work), capture_output=True, text=True)
        assert r.returncode != 0 and flags[0] in r.stderr
    r = subprocess.run([CLI, "prefilter", "q_ss", "db_ss", "pref_x", "--exact-kmer-matching", "1"], cwd=str(work), capture_output=True, text=True)
    assert r.returncode != 0 and "--exact-kmer-matching" in r.stderr
