"""The host module (foldseek_b200/csrc/cli_main.cpp: parameter handling, DB reading and encoding, coverage filters, checkCriteria,
accept/reject counters, result sorting and serialisation) against the reference's own modules, WITHOUT a GPU: oracle/build.sh links
cli_main.cpp against oracle/mock_fsb200.cpp, a stand-in for the C ABI that answers with the oracle restatement (test infrastructure,
never shipped).  What the CUDA library returns through the same ABI is compared with the same oracle in the -m gpu tests; here the
host-side logic around it is swept over the module parameters and compared per result-DB key with
  * the committed reference outputs on the reference's example/ proteins (tests/golden/example, always), and
  * live runs of oracle/_ref/refseek (= the reference's sources compiled here) where it exists.
This sweep found that --cov-mode 1 / 2 were swapped in the host module (1 is TARGET, 2 is QUERY: Parameters.h:284-289)."""
import os
import shutil
import subprocess

import pytest

from foldseek_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "oracle", "_mock", "foldseek-b200-mock")
REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")
EX = os.path.join(ROOT, "tests", "golden", "example")
ENV = dict(os.environ, FSB200_DATA=os.path.join(ROOT, "foldseek_b200", "data"))

pytestmark = pytest.mark.skipif(not os.path.exists(MOCK), reason="oracle/_mock/foldseek-b200-mock not built (sh oracle/build.sh)")


def _params():
    p = {}
    for line in open(os.path.join(EX, "params.txt")):
        mod, _, rest = line.strip().partition(" ")
        p[mod] = rest.split()
    return p


def _with(flags, kv):
    flags = list(flags)
    for k, v in kv:
        if k in flags:
            flags[flags.index(k) + 1] = v
        else:
            flags += [k, v]
    return flags


@pytest.fixture()
def work(tmp_path):
    d = tmp_path / "ex"
    shutil.copytree(EX, str(d))
    return d


def _run(binary, work, *args, env=None):
    r = subprocess.run([binary] + list(args), cwd=str(work), capture_output=True, text=True, env=env or ENV)
    assert r.returncode == 0, (args, r.stderr[-600:], r.stdout[-300:])
    return r


def _same(work, got, ref):
    g, e = synth.read_result_db(str(work / got)), synth.read_result_db(str(work / ref))
    assert sorted(g) == sorted(e)
    for k in e:
        assert g[k] == e[k], (got, k, [x for x in e[k] if x not in g[k]][:2], [x for x in g[k] if x not in e[k]][:2])
    assert open(work / (got + ".dbtype"), "rb").read() == open(work / (ref + ".dbtype"), "rb").read()
    return sum(len(v) for v in e.values())


# ---- committed reference outputs (the scenarios of tests/test_example_gpu.py, on the CPU stand-in) ----
def test_example_scenarios_against_committed_reference_outputs(work):
    p = _params()
    _run(MOCK, work, "prefilter", "q_ss", "db_ss", "pref_q", *p["prefilter"])
    assert _same(work, "pref_q", "ref_pref_q") >= 20
    _run(MOCK, work, "structurealign", "q", "db", "ref_pref_q", "aln_q", *p["structurealign"])
    assert _same(work, "aln_q", "ref_aln_q") >= 8
    _run(MOCK, work, "prefilter", "q_ss", "pad_ss", "pref_q_pad", *p["prefilter"])                      # padded (GPU) DB layout
    assert _same(work, "pref_q_pad", "ref_pref_q_pad") >= 20
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_all", "--lanes", "2", *p["prefilter"])         # identity hits; two lanes
    assert _same(work, "pref_all", "ref_pref_all") >= 600
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_all", "--gpu-list", "0,0", "--lanes", "2", *p["structurealign"])
    assert _same(work, "aln_all", "ref_aln_all") >= 100
    cov = _with(p["structurealign"], [("-c", "0.8"), ("--cov-mode", "0"), ("--min-seq-id", "0.2"), ("-a", "1")])
    _run(MOCK, work, "structurealign", "q", "db", "ref_pref_q", "aln_cov", *cov)
    assert _same(work, "aln_cov", "ref_aln_q_cov") >= 1
    alt = _with(p["structurealign"], [("--alt-ali", "2"), ("-a", "1")])
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_alt", "--lanes", "2", *alt)           # alternative alignments
    assert _same(work, "aln_alt", "ref_aln_all_alt2") >= 950
    sp = _with(p["prefilter"], [("--split", "3"), ("--split-mode", "0"), ("--max-seqs", "10")])
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_split", "--lanes", "2", *sp)                         # target split
    assert _same(work, "pref_split", "ref_pref_all_split3") >= 400
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_split2", "--gpu-list", "0,0", *sp)                  # the three ranges sharded over two GPUs
    assert _same(work, "pref_split2", "ref_pref_all_split3") >= 400
    nb = [("--comp-bias-corr", "0")]
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_nb", *_with(p["prefilter"], nb + [("-s", "4"), ("--max-seqs", "20")]))   # no composition-bias correction
    assert _same(work, "pref_nb", "ref_pref_all_nobias") >= 100
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_nb", *_with(p["structurealign"], nb + [("-e", "0.001")]))
    assert _same(work, "aln_nb", "ref_aln_all_nobias") >= 50
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_3di", *_with(p["structurealign"], [("--alignment-type", "0"), ("-a", "1")]))   # 3Di only
    assert _same(work, "aln_3di", "ref_aln_all_3di") >= 100
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_s75", *_with(p["prefilter"], [("-s", "7.5"), ("--min-ungapped-score", "50"), ("--max-seqs", "20")]))
    assert _same(work, "pref_s75", "ref_pref_all_s75") >= 100
    gap = [("--gap-open", "aa:8,nucl:8"), ("--gap-extend", "aa:2,nucl:2"), ("-e", "0.001"), ("-a", "1")]
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_gap", *_with(p["structurealign"], gap))
    assert _same(work, "aln_gap", "ref_aln_all_gap82") >= 50
    tm1 = _with(p["structurealign"], [("--tmscore-threshold-mode", "1"), ("--tmscore-threshold", "0.3")])
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_tm1", *tm1)                            # TM-score by query length
    assert _same(work, "aln_tm1", "ref_aln_all_tmmode1") >= 50
    for mode in ("1", "2"):
        kv = [("-c", "0.7"), ("--cov-mode", mode)]
        _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_c", *_with(p["prefilter"], kv))
        assert _same(work, "pref_c", "ref_pref_all_cov" + mode) >= 600
        _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_c", *_with(p["structurealign"], kv))
        assert _same(work, "aln_c", "ref_aln_all_cov" + mode) >= 100


def test_aligner_batches_of_candidate_pairs(work):
    """a context's entries are aligned in batches bounded by candidate pairs (FSB200_ALIGN_BATCH, default 32 M) so that exhaustive searches
    and very long prefilter lists do not need host memory in proportion to queries x targets; any batch size writes the same DB"""
    alt = _with(_params()["structurealign"], [("--alt-ali", "2"), ("-a", "1")])
    for batch in ("1", "700"):
        _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "aln_b", "--lanes", "2", *alt, env=dict(ENV, FSB200_ALIGN_BATCH=batch))
        assert _same(work, "aln_b", "ref_aln_all_alt2") >= 950


# ---- live sweep against the reference's modules ----
ALN_SWEEP = [[("-c", "0.5"), ("--cov-mode", m)] for m in "012"] + [[("-c", "0.8"), ("--cov-mode", m)] for m in "012345"] + [
    [("--min-seq-id", "0.3")], [("--min-aln-len", "50")], [("--max-accept", "3")], [("--max-rejected", "2")],
    [("--max-rejected", "1"), ("--max-accept", "2")], [("-e", "0.001")], [("-e", "1e-10")], [("-e", "1000")],
    [("--sort-by-structure-bits", "0")], [("--sort-by-structure-bits", "0"), ("-a", "1")], [("-a", "1")],
    [("--tmscore-threshold", "0.5")], [("--lddt-threshold", "0.6")], [("--tmscore-threshold", "0.3"), ("--sort-by-structure-bits", "0")],
    [("--lddt-threshold", "0.5"), ("--sort-by-structure-bits", "0")], [("--comp-bias-corr-scale", "0.2")],
    [("--gap-open", "aa:8,nucl:8"), ("--gap-extend", "aa:2,nucl:2")], [("--add-self-matches", "1")],
    [("-c", "0.6"), ("--cov-mode", "2"), ("--min-seq-id", "0.15"), ("--max-rejected", "5"), ("-a", "1")],
    [("--seq-id-mode", "1")], [("--seq-id-mode", "2"), ("--min-seq-id", "0.2")], [("--alignment-mode", "1")],
    [("--alignment-mode", "2"), ("--tmscore-threshold", "0.5")], [("--alignment-mode", "3"), ("-a", "1")],
    [("--alt-ali", "1")], [("--alt-ali", "4"), ("-a", "1")], [("--alt-ali", "2"), ("--sort-by-structure-bits", "0"), ("-c", "0.3"), ("--cov-mode", "2")],
    [("--alt-ali", "3"), ("--min-aln-len", "20"), ("--max-accept", "5"), ("-e", "1")],
    [("--realign", "1"), ("--wrapped-scoring", "1"), ("--alignment-output-mode", "1"), ("--zdrop", "10")],   # read by no code of structurealign
    [("--tmscore-threshold-mode", "1")], [("--tmscore-threshold-mode", "2"), ("--tmscore-threshold", "0.4")],
    [("--tmscore-threshold-mode", "1"), ("--tmscore-threshold", "0.3"), ("--sort-by-structure-bits", "0")],
]
PREF_SWEEP = [[("-c", "0.8"), ("--cov-mode", m)] for m in "02"] + [      # 1, 3 and 4 do not filter in the prefilter (Prefiltering.cpp:886-894)
    [("--max-seqs", "5")], [("-s", "7.5")], [("--min-ungapped-score", "50")], [("--comp-bias-corr-scale", "0.5")], [("--add-self-matches", "1")],
    [("--mask-lower-case", "0")],
    # target split (Prefiltering.cpp:384-397,412-526): own index and a shorter list per target range, merged by score and key
    [("--split", "3"), ("--split-mode", "0")], [("--split", "4"), ("--split-mode", "0"), ("--max-seqs", "5"), ("-c", "0.8"), ("--cov-mode", "5")],
    [("--split", "2"), ("--split-mode", "1")],
]
needs_ref = pytest.mark.skipif(not os.path.exists(REFSEEK), reason="oracle/_ref/refseek (the reference compiled from /root/reference) not present")


def _tag(kv):
    return " ".join(k + " " + v for k, v in kv)


@needs_ref
@pytest.mark.parametrize("kv", ALN_SWEEP, ids=_tag)
def test_structurealign_parameters_match_the_reference_module(work, kv):
    flags = _with(_params()["structurealign"], kv)
    _run(REFSEEK, work, "structurealign", "db", "db", "ref_pref_all", "out_ref", *flags)
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all", "out_got", *flags)
    _same(work, "out_got", "out_ref")


@needs_ref
@pytest.mark.parametrize("kv", PREF_SWEEP, ids=_tag)
def test_prefilter_parameters_match_the_reference_module(work, kv):
    flags = _with(_params()["prefilter"], kv)
    _run(REFSEEK, work, "prefilter", "db_ss", "db_ss", "out_ref", *flags)
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "out_got", *flags)
    _same(work, "out_got", "out_ref")


@needs_ref
def test_query_db_differs_from_target_db_with_self_matches(work):
    p = _params()
    for mod, pos in (("prefilter", ["q_ss", "db_ss"]), ("structurealign", ["q", "db", "ref_pref_q"])):
        flags = _with(p[mod], [("--add-self-matches", "1")])
        _run(REFSEEK, work, mod, *pos, "out_ref", *flags)
        _run(MOCK, work, mod, *pos, "out_got", *flags)
        _same(work, "out_got", "out_ref")


def _read_compressed_db(path):
    """entries of a --compressed 1 result DB (DBWriter.cpp:331-398): u32 stored size, zstd frame or raw bytes, terminator 0 / 0xFF"""
    import ctypes
    import struct
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = ctypes.c_size_t
    z.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    data = open(path, "rb").read()
    assert struct.unpack("<I", open(path + ".dbtype", "rb").read())[0] >> 31 == 1
    out = {}
    for line in open(path + ".index"):
        key, off, length = (int(x) for x in line.split())
        n = struct.unpack_from("<I", data, off)[0]
        body = data[off + 4:off + 4 + n]
        if data[off + 4 + n] == 0:
            buf = ctypes.create_string_buffer(length + 64)
            got = z.ZSTD_decompress(buf, len(buf), body, n)
            assert got == length - 1
            body = buf.raw[:got]
        out[key] = [tuple(r.split("\t")) for r in body.decode().split("\n") if r]
    return out


@needs_ref
def test_compressed_result_databases_both_ways(work):
    """--compressed 1: the reference's compressed prefilter DB is read by the host module, the host module's compressed prefilter DB is
    read by the reference's structurealign, and its compressed alignment DB holds the reference's rows"""
    p = _params()
    pc = _with(p["prefilter"], [("--compressed", "1")])
    _run(REFSEEK, work, "prefilter", "db_ss", "db_ss", "pref_rc", *pc)
    _run(REFSEEK, work, "structurealign", "db", "db", "pref_rc", "out_ref", *p["structurealign"])
    _run(MOCK, work, "structurealign", "db", "db", "pref_rc", "out_got", *p["structurealign"])
    assert _same(work, "out_got", "out_ref") >= 100
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_gc", *pc)
    _run(REFSEEK, work, "structurealign", "db", "db", "pref_gc", "out_ref2", *p["structurealign"])
    assert _same(work, "out_ref2", "ref_aln_all") >= 100
    _run(MOCK, work, "structurealign", "db", "db", "pref_gc", "aln_gc", *_with(p["structurealign"], [("--compressed", "1")]))
    got = _read_compressed_db(str(work / "aln_gc"))
    ref = synth.read_result_db(str(work / "ref_aln_all"))
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert [tuple(r) for r in ref[k]] == got[k], k
