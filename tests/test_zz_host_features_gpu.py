"""Host-module features added after the last GPU session of round 2 (no GPU minutes were left): they were developed against the
reference's modules on the oracle-backed stand-in for the C ABI (tests/test_host_module_cpu.py) and use only library calls the other
-m gpu tests cover.  The file sorts last so that a surprise here cannot hide the results of the established tests under `pytest -x`."""
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
    r = subprocess.run([CLI] + list(args), cwd=str(work), capture_output=True, text=True, timeout=900)
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


def test_alternative_alignments(work):
    """--alt-ali 2 (computeAlternativeAlignment, structurealign.cpp:115-138,415-429): after every accepted pair the aligned target range
    is masked with X and the pair is aligned again, twice; the masked targets of a round are staged as a small target DB of a second
    context on the same GPU.  The reference wrote 1005 rows for this run against 834 without alternatives."""
    al = list(_params()["structurealign"])
    for flag, v in (("--alt-ali", "2"), ("-a", "1")):
        al[al.index(flag) + 1] = v
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_alt", *al)
    assert _same(work, "aln_alt", "ref_aln_all_alt2") >= 950
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_alt2", "--lanes", "2", *al)
    assert _same(work, "aln_alt2", "ref_aln_all_alt2") >= 950


def test_tmscore_normalised_by_query_length(work):
    """--tmscore-threshold-mode 1 --tmscore-threshold 0.3 (TMaligner::normalization, TMaligner.cpp:330-341): the normalisation length
    of the TM-score kernel's d0 / score_d8 / Lnorm is the query length instead of the aligned span; it changes which rows pass the
    threshold and the structure-bits score of every row"""
    al = list(_params()["structurealign"])
    for flag, v in (("--tmscore-threshold-mode", "1"), ("--tmscore-threshold", "0.3")):
        al[al.index(flag) + 1] = v
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_tm1", *al)
    assert _same(work, "aln_tm1", "ref_aln_all_tmmode1") >= 50


def test_target_split(work):
    """--split 3 --split-mode 0 --max-seqs 10 (TARGET_DB_SPLIT, Prefiltering.cpp:384-397,412-526): the 30 targets are searched as three
    ranges of about equal residues, each staged and indexed on its own with a result list of 10/3 + 4 sigma = 10 hits, the lists of a
    query merged by score and target key - up to 30 hits per query where the unsplit run keeps 10"""
    pf = list(_params()["prefilter"])
    for flag, v in (("--split", "3"), ("--split-mode", "0"), ("--max-seqs", "10")):
        pf[pf.index(flag) + 1] = v
    _run(work, "prefilter", "db_ss", "db_ss", "pref_split", *pf)
    assert _same(work, "pref_split", "ref_pref_all_split3") >= 400


def test_search_module_both_stages_on_one_staged_target(work):
    """`foldseek-b200 search`: prefilter and structurealign in one process on one staged copy of the target (AA + 3Di + C-alpha + k-mer
    index), hit lists handed over in memory; the alignment DB equals the chained reference modules' (CPU twin: tests/test_search_module_cpu.py)"""
    _run(work, "search", "q", "db", "res_q", "tmp")
    assert _same(work, "res_q", "ref_aln_q") >= 8
    _run(work, "search", "db", "db", "res_all", "tmp", "--lanes", "2")
    assert _same(work, "res_all", "ref_aln_all") >= 100
    # BASELINE configs[0] to its last byte: search + convertalis = the .m8 file `foldseek easy-search example/d1asha_ example/ aln.m8 tmp`
    # ends with (the reference's createdb -> search -> convertalis, tests/golden/example/ref_aln_q.m8)
    _run(work, "convertalis", "q", "db", "res_q", "aln.m8")
    assert open(work / "aln.m8", "rb").read() == open(work / "ref_aln_q.m8", "rb").read()


def test_padded_target_layout_through_structurealign_and_search(work):
    """the GPU-padded target of `foldseek makepaddeddb`: pad_ss in the padded numeric layout, pad / pad_ca / pad_h re-keyed in descending
    length order (derived here the way renamedbkeys does it); structurealign and search write the reference's rows (ref_aln_q_pad)"""
    synth.rekey_like_padded(str(work))
    p = _params()
    _run(work, "structurealign", "q", "pad", "ref_pref_q_pad", "aln_pad", *p["structurealign"])
    assert _same(work, "aln_pad", "ref_aln_q_pad") >= 8
    _run(work, "search", "q", "pad", "res_pad", "tmp")
    assert _same(work, "res_pad", "ref_aln_q_pad") >= 8


def test_without_composition_bias_correction(work):
    """--comp-bias-corr 0 in both modules (fsb200_config.comp_bias_corr: no query bias in the k-mer / ungapped scores, none in the SW
    profiles and the block aligner); the cluster workflow's early steps run this way"""
    pf, al = list(_params()["prefilter"]), list(_params()["structurealign"])
    for flag, v in (("--comp-bias-corr", "0"), ("-s", "4"), ("--max-seqs", "20")):
        pf[pf.index(flag) + 1] = v
    for flag, v in (("--comp-bias-corr", "0"), ("-e", "0.001")):
        al[al.index(flag) + 1] = v
    _run(work, "prefilter", "db_ss", "db_ss", "pref_nb", *pf)
    assert _same(work, "pref_nb", "ref_pref_all_nobias") >= 100
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_nb", *al)
    assert _same(work, "aln_nb", "ref_aln_all_nobias") >= 50


def test_alignment_type_3di_only(work):
    """--alignment-type 0: the AA matrix has bit factor 0 (structurealign.cpp:264-265), so scores, traceback and - because both
    composition-bias corrections are computed with the AA matrix object - the biases come from the 3Di matrix alone"""
    al = list(_params()["structurealign"])
    for flag, v in (("--alignment-type", "0"), ("-a", "1")):
        al[al.index(flag) + 1] = v
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_3di", *al)
    assert _same(work, "aln_3di", "ref_aln_all_3di") >= 100


def test_scoring_parameters_away_from_their_defaults(work):
    """-s 7.5 --min-ungapped-score 50 --max-seqs 20 in the prefilter (k-mer threshold, diagonal score cut, list length) and
    --gap-open 8 --gap-extend 2 -e 0.001 in the aligner (SW kernels, block aligner, E-value gates): values the library reads from
    fsb200_config, which the other GPU tests leave at their defaults"""
    pf, al = list(_params()["prefilter"]), list(_params()["structurealign"])
    for flag, v in (("-s", "7.5"), ("--min-ungapped-score", "50"), ("--max-seqs", "20")):
        pf[pf.index(flag) + 1] = v
    for flag, v in (("--gap-open", "aa:8,nucl:8"), ("--gap-extend", "aa:2,nucl:2"), ("-e", "0.001"), ("-a", "1")):
        al[al.index(flag) + 1] = v
    _run(work, "prefilter", "db_ss", "db_ss", "pref_s75", *pf)
    assert _same(work, "pref_s75", "ref_pref_all_s75") >= 100
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_gap", *al)
    assert _same(work, "aln_gap", "ref_aln_all_gap82") >= 50


def test_masked_3di_target(work):
    """lower-case (masked) stretches in the target's 3Di DB under the default --mask-lower-case 1: X for the k-mer index and the ungapped
    scores, letters for the aligner; `search` stages a masked copy for its prefilter stage next to the aligner's"""
    synth.mask_lower_case(str(work / "db_ss"))
    p = _params()
    _run(work, "prefilter", "db_ss", "db_ss", "pref_m", *p["prefilter"])
    assert _same(work, "pref_m", "ref_pref_all_masked") >= 600
    _run(work, "structurealign", "db", "db", "ref_pref_all_masked", "aln_m", *p["structurealign"])
    assert _same(work, "aln_m", "ref_aln_all_masked") >= 100
    _run(work, "search", "db", "db", "res_m", "tmp")
    assert _same(work, "res_m", "ref_aln_all_masked") >= 100


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "refseek")), reason="reference driver not present")
def test_ragged_lengths_and_x_residues_against_the_reference_modules(tmp_path):
    """targets of 3..900 residues (below the k-mer span, above the packed SW kernel's 384) with X residues, long planted queries, the
    workflow's parameter strings: `search -a 1` against the reference's prefilter + structurealign run live on this box"""
    refseek = os.path.join(ROOT, "oracle", "_ref", "refseek")
    synth.gen_edge_case_dbs(str(tmp_path))
    p = _params()
    al = list(p["structurealign"]); al[al.index("-a") + 1] = "1"
    for args in (["prefilter", "q_ss", "t_ss", "pref_ref"] + p["prefilter"], ["structurealign", "q", "t", "pref_ref", "aln_ref"] + al):
        subprocess.run([refseek] + args, cwd=str(tmp_path), check=True, capture_output=True, timeout=900)
    _run(tmp_path, "prefilter", "q_ss", "t_ss", "pref_got", *p["prefilter"])
    assert _same(tmp_path, "pref_got", "pref_ref") >= 1000
    _run(tmp_path, "search", "q", "t", "res", "tmp", "-a", "1")
    assert _same(tmp_path, "res", "aln_ref") >= 10


def test_two_contexts_with_device_to_device_staging(work):
    """--gpu-list 0,0: the multi-GPU path of the host module on one device - the second context receives the staged target from the
    first device-to-device (fsb200_export_staged_target -> fsb200_load_target_device), the queries are split between the two by residues,
    one writer; with a target split the ranges are sharded over the two contexts instead.  Results as on one context."""
    p = _params()
    _run(work, "prefilter", "db_ss", "db_ss", "pref_2", "--gpu-list", "0,0", *p["prefilter"])
    assert _same(work, "pref_2", "ref_pref_all") >= 600
    _run(work, "structurealign", "db", "db", "ref_pref_all", "aln_2", "--gpu-list", "0,0", *p["structurealign"])
    assert _same(work, "aln_2", "ref_aln_all") >= 100
    pf = list(p["prefilter"])
    for flag, v in (("--split", "3"), ("--split-mode", "0"), ("--max-seqs", "10")):
        pf[pf.index(flag) + 1] = v
    _run(work, "prefilter", "db_ss", "db_ss", "pref_split2", "--gpu-list", "0,0", *pf)
    assert _same(work, "pref_split2", "ref_pref_all_split3") >= 400


def test_gpuserver_serves_module_calls_from_resident_contexts(work, sockdir):
    """`foldseek-b200 gpuserver db --lanes 2` stages the target DB of both modules once (sequences, C-alpha, k-mer index, two lanes);
    prefilter / structurealign calls with --gpu-server 1 are run by it on the resident contexts and write the reference's result DBs
    (CPU twin on the stand-in: tests/test_gpuserver_cpu.py)"""
    import signal
    env = dict(os.environ, FSB200_SOCKET_DIR=sockdir)
    srv = subprocess.Popen([CLI, "gpuserver", "db", "--lanes", "2"], cwd=str(work), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        p = _params()

        def call(*args):
            return subprocess.run([CLI] + list(args) + ["--gpu-server", "1", "--gpu-server-wait-timeout", "300"], cwd=str(work), env=env, capture_output=True, text=True,
                                  timeout=600)
        r = call("prefilter", "db_ss", "db_ss", "pref_all", *p["prefilter"])
        assert r.returncode == 0 and "served 1 prefilter call" in r.stdout, r.stdout + r.stderr
        assert _same(work, "pref_all", "ref_pref_all") >= 600
        r = call("prefilter", "q_ss", "db_ss", "pref_q", *p["prefilter"])
        assert r.returncode == 0 and "served 2 prefilter call" in r.stdout, r.stdout + r.stderr
        assert _same(work, "pref_q", "ref_pref_q") >= 20
        r = call("structurealign", "db", "db", "ref_pref_all", "aln_all", *p["structurealign"])
        assert r.returncode == 0 and "served 1 structurealign call" in r.stdout, r.stdout + r.stderr
        assert _same(work, "aln_all", "ref_aln_all") >= 100
        r = call("structurealign", "q", "db", "ref_pref_q", "aln_q", *p["structurealign"])
        assert r.returncode == 0 and "served 2 structurealign call" in r.stdout, r.stdout + r.stderr
        assert _same(work, "aln_q", "ref_aln_q") >= 8
        r = call("structurealign", "q", "db", "no_such_pref", "aln_x", *p["structurealign"])
        assert r.returncode != 0 and "no_such_pref" in r.stderr
        r = call("search", "q", "db", "res_q", "tmp")
        assert r.returncode == 0 and "served 3 structurealign call" in r.stdout, r.stdout + r.stderr
        assert _same(work, "res_q", "ref_aln_q") >= 8
        assert srv.poll() is None
    finally:
        if srv.poll() is None:
            srv.send_signal(signal.SIGTERM)
            try:
                srv.wait(timeout=60)
            except subprocess.TimeoutExpired:
                srv.kill()
