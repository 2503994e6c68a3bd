"""`foldseek-b200 search <queryDB> <targetDB> <alnDB> <tmpDir>`: both stages of the search workflow in one process, the prefilter's hit
lists handed to the alignment stage in memory and both stages on ONE staged copy of the target.  Without a GPU, on the
oracle-backed stand-in for the C ABI (oracle/_mock, test infrastructure): the alignment DB must be the one the reference's two modules
write when chained the way data/structuresearch.sh chains them - the committed reference outputs, and live refseek chains over
parameter settings that touch both stages."""
import os
import shutil
import signal
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


@pytest.fixture()
def work(tmp_path):
    d = tmp_path / "ex"
    shutil.copytree(EX, str(d))
    return d


def _run(binary, work, *args, env=ENV):
    r = subprocess.run([binary] + list(args), cwd=str(work), capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (args, r.stderr[-600:], r.stdout[-300:])
    return r


def _same(work, got, ref):
    g, e = synth.read_result_db(str(work / got)), synth.read_result_db(str(work / ref))
    assert sorted(g) == sorted(e)
    for k in e:
        assert g[k] == e[k], (got, k, [x for x in e[k] if x not in g[k]][:2], [x for x in g[k] if x not in e[k]][:2])
    assert open(work / (got + ".dbtype"), "rb").read() == open(work / (ref + ".dbtype"), "rb").read()
    return sum(len(v) for v in e.values())


def test_search_writes_the_chained_modules_alignment_db(work):
    _run(MOCK, work, "search", "q", "db", "res_q", "tmp")                                   # defaults of `foldseek search`
    assert _same(work, "res_q", "ref_aln_q") >= 8
    _run(MOCK, work, "search", "db", "db", "res_all", "tmp", "--lanes", "2", "--threads", "4", "-s", "9.5", "--max-seqs", "1000", "-e", "10")
    assert _same(work, "res_all", "ref_aln_all") >= 100                                      # identity hits through both stages
    _run(MOCK, work, "search", "q", "db.idx", "res_idx", "tmp")                             # "${TARGET}${INDEXEXT}" of a createindex'ed DB
    assert _same(work, "res_idx", "ref_aln_q") >= 8
    r = subprocess.run([MOCK, "search", "q", "db", "res_x", "tmp", "--num-iterations", "3"], cwd=str(work), capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "--num-iterations" in r.stderr


SWEEP = [
    [("-c", "0.7"), ("--cov-mode", "2")],                                    # the prefilter's canBeCovered AND the aligner's coverage checks
    [("--max-seqs", "5"), ("-e", "0.1")],
    [("--split", "3"), ("--split-mode", "0"), ("--max-seqs", "10")],         # target split in the prefilter stage
    [("--alt-ali", "2"), ("-a", "1"), ("--min-ungapped-score", "40")],
    [("-s", "6"), ("--sort-by-structure-bits", "0"), ("--add-self-matches", "1")],
]


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="oracle/_ref/refseek (the reference compiled from /root/reference) not present")
@pytest.mark.parametrize("kv", SWEEP, ids=lambda kv: " ".join(k + " " + v for k, v in kv))
def test_search_parameters_match_the_chained_reference_modules(work, kv):
    p = _params()
    flags = {m: list(p[m]) for m in p}
    mine = []
    for k, v in kv:
        for m in flags:
            if k in flags[m]:
                flags[m][flags[m].index(k) + 1] = v
        mine += [k, v]
    _run(REFSEEK, work, "prefilter", "db_ss", "db_ss", "pref_ref", *flags["prefilter"])
    _run(REFSEEK, work, "structurealign", "db", "db", "pref_ref", "aln_ref", *flags["structurealign"])
    _run(MOCK, work, "search", "db", "db", "aln_got", "tmp", *mine)
    _same(work, "aln_got", "aln_ref")


def test_search_through_the_gpuserver(work, sockdir):
    """one staged copy of the target serves search calls; a call with other prefilter parameters replaces it"""
    env = dict(ENV, FSB200_SOCKET_DIR=sockdir)
    srv = subprocess.Popen([MOCK, "gpuserver", "db", "--lanes", "2"], cwd=str(work), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        on = ["--gpu-server", "1", "--gpu-server-wait-timeout", "60"]
        r = _run(MOCK, work, "search", "db", "db", "res_all", "tmp", *on, env=env)
        assert "served 1 structurealign call" in r.stdout and _same(work, "res_all", "ref_aln_all") >= 100
        r = _run(MOCK, work, "search", "q", "db", "res_q", "tmp", *on, env=env)
        assert "served 2 structurealign call" in r.stdout and _same(work, "res_q", "ref_aln_q") >= 8
        r = _run(MOCK, work, "search", "q", "db", "res_q5", "tmp", "--max-seqs", "5", *on, env=env)
        assert "served 1 structurealign call" in r.stdout                    # contexts are created with --max-seqs: restaged
        got, ref = synth.read_result_db(str(work / "res_q5")), synth.read_result_db(str(work / "ref_aln_q"))
        assert all(r_ in ref[k] for k in got for r_ in got[k]) and 0 < sum(map(len, got.values())) <= 5
        assert srv.poll() is None
    finally:
        srv.send_signal(signal.SIGTERM)
        srv.wait(timeout=30)


def test_search_then_convertalis_is_the_easy_search_output(work):
    """BASELINE configs[0] to its last byte: `foldseek easy-search example/d1asha_ example/ aln.m8 tmp` = createdb, search, convertalis
    (data/easystructuresearch.sh).  On the DBs the reference's createdb wrote, `foldseek-b200 search` + `foldseek-b200 convertalis` write
    the .m8 file the reference's search + convertalis write (tests/golden/example/ref_aln_q.m8); with backtraces in the alignment DB the
    mismatch / gap-open columns come from the CIGARs (ref_aln_all_alt2.m8)."""
    _run(MOCK, work, "search", "q", "db", "res_q", "tmp")
    _run(MOCK, work, "convertalis", "q", "db", "res_q", "aln.m8")
    assert open(work / "aln.m8", "rb").read() == open(work / "ref_aln_q.m8", "rb").read()
    _run(MOCK, work, "convertalis", "db", "db", "ref_aln_all_alt2", "alt.m8", "--format-mode", "0", "--threads", "4")
    assert open(work / "alt.m8", "rb").read() == open(work / "ref_aln_all_alt2.m8", "rb").read()
    r = subprocess.run([MOCK, "convertalis", "q", "db", "res_q", "x.m8", "--format-output", "query,target,qaln"], cwd=str(work), capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "--format-output" in r.stderr


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="oracle/_ref/refseek (the reference compiled from /root/reference) not present")
def test_convertalis_matches_the_reference_module_live(work):
    for q, aln in (("db", "ref_aln_all"), ("q", "ref_aln_q_cov"), ("db", "ref_aln_all_cov1"), ("db", "ref_aln_all_tmmode1")):
        _run(REFSEEK, work, "convertalis", q, "db", aln, "ref.m8", "--threads", "1")
        _run(MOCK, work, "convertalis", q, "db", aln, "got.m8")
        assert open(work / "got.m8", "rb").read() == open(work / "ref.m8", "rb").read(), aln


def test_fasta_identifier_forms(tmp_path):
    """Util::parseFastaHeader: first word, or the accession of the database-prefixed header forms"""
    import struct
    heads = ["d1asha_ a.1.1.2 (A:) Hemoglobin", "sp|P69905|HBA_HUMAN Hemoglobin subunit alpha", "gi|12345|ref|NP_000549.1| hemoglobin", "consensus_tr|Q9XYZ1|NAME x",
             "gnl|db|ID77 rest", "plain"]
    want = ["d1asha_", "P69905", "NP_000549.1", "Q9XYZ1", "ID77", "plain"]
    n = len(heads)

    def db(name, entries, dbtype):
        off = 0
        with open(str(tmp_path / name), "wb") as d, open(str(tmp_path / (name + ".index")), "w") as i:
            for k, e in enumerate(entries):
                d.write(e + b"\0")
                i.write("%d\t%d\t%d\n" % (k, off, len(e) + 1))
                off += len(e) + 1
        open(str(tmp_path / (name + ".dbtype")), "wb").write(struct.pack("<i", dbtype))
    db("s_h", [(h + "\n").encode() for h in heads], 12)
    db("aln", [("%d\t100\t0.500\t1.000E-05\t0\t9\t10\t0\t9\t10\n" % ((k + 1) % n)).encode() for k in range(n)], 5)
    _run(MOCK, tmp_path, "convertalis", "s", "s", "aln", "out.m8")
    rows = [l.split("\t") for l in open(str(tmp_path / "out.m8")).read().splitlines()]
    assert [r[0] for r in rows] == want and [r[1] for r in rows] == want[1:] + want[:1]
    assert rows[0][2:] == ["0.500", "10", "5", "0", "1", "10", "1", "10", "1.000E-05", "100"]
    if os.path.exists(REFSEEK):      # the reference's own parser on the same headers
        db("s", [b"ACDEFGHIKL\n"] * n, 0)
        _run(REFSEEK, tmp_path, "convertalis", "s", "s", "aln", "ref.m8", "--threads", "1")
        assert open(str(tmp_path / "ref.m8"), "rb").read() == open(str(tmp_path / "out.m8"), "rb").read()


def test_easy_search_on_databases(work):
    """`foldseek-b200 easy-search <queryDB> <targetDB> <aln.m8> <tmpDir>` = search + convertalis for database inputs (easystructuresearch.sh
    skips createdb when "<input>.dbtype" exists); a structure file as input is refused naming createdb, nothing is converted here"""
    _run(MOCK, work, "easy-search", "q", "db", "easy.m8", "tmp_easy", "-e", "10", "--threads", "2")
    assert open(work / "easy.m8", "rb").read() == open(work / "ref_aln_q.m8", "rb").read()
    assert not os.path.exists(str(work / "tmp_easy" / "result.dbtype"))                  # temporary alignment DB removed (default of the easy workflows)
    _run(MOCK, work, "easy-search", "q", "db", "easy2.m8", "tmp_easy", "--remove-tmp-files", "0")
    assert _same(work, "tmp_easy/result", "ref_aln_q") >= 8
    (work / "some.pdb").write_text("ATOM\n")
    r = subprocess.run([MOCK, "easy-search", "some.pdb", "db", "x.m8", "tmp_easy"], cwd=str(work), capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "createdb" in r.stderr and not os.path.exists(str(work / "x.m8"))


def test_padded_target_layout_through_structurealign_and_search(work):
    """the GPU-padded target of `foldseek makepaddeddb` (what `--gpu 1` searches run against): <pad>_ss in the padded numeric layout
    (DBTYPE_EXTENDED_GPU), <pad> / <pad>_ca / <pad>_h re-keyed in descending length order.  structurealign on the reference's prefilter
    result and the search module (prefilter on the padded 3Di DB, aligner on the re-keyed AA DB, one staged copy) write the reference's rows."""
    synth.rekey_like_padded(str(work))
    p = _params()
    _run(MOCK, work, "structurealign", "q", "pad", "ref_pref_q_pad", "aln_pad", *p["structurealign"])
    assert _same(work, "aln_pad", "ref_aln_q_pad") >= 8
    _run(MOCK, work, "search", "q", "pad", "res_pad", "tmp")
    assert _same(work, "res_pad", "ref_aln_q_pad") >= 8


def test_masked_3di_target(work):
    """lower-case (masked) stretches in the target's 3Di DB, default --mask-lower-case 1: the prefilter sees X there (index and ungapped
    scores), the aligner the letters.  The modules and `search` - whose prefilter then stages a masked copy of its own instead of sharing
    the aligner's - write the reference's DBs; with --mask-lower-case 0 `search` shares the one copy again and follows the reference too."""
    synth.mask_lower_case(str(work / "db_ss"))
    p = _params()
    _run(MOCK, work, "prefilter", "db_ss", "db_ss", "pref_m", *p["prefilter"])
    assert _same(work, "pref_m", "ref_pref_all_masked") >= 600
    assert synth.read_result_db(str(work / "pref_m")) != synth.read_result_db(str(work / "ref_pref_all"))       # the masking matters
    _run(MOCK, work, "structurealign", "db", "db", "ref_pref_all_masked", "aln_m", *p["structurealign"])
    assert _same(work, "aln_m", "ref_aln_all_masked") >= 100
    _run(MOCK, work, "search", "db", "db", "res_m", "tmp")
    assert _same(work, "res_m", "ref_aln_all_masked") >= 100
    if os.path.exists(REFSEEK):
        off = list(p["prefilter"]); off[off.index("--mask-lower-case") + 1] = "0"
        _run(REFSEEK, work, "prefilter", "db_ss", "db_ss", "pref_0", *off)
        _run(REFSEEK, work, "structurealign", "db", "db", "pref_0", "aln_0", *p["structurealign"])
        _run(MOCK, work, "search", "db", "db", "res_0", "tmp", "--mask-lower-case", "0")
        _same(work, "res_0", "aln_0")


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="oracle/_ref/refseek (the reference compiled from /root/reference) not present")
def test_ragged_lengths_and_x_residues_against_the_reference_modules(tmp_path):
    """targets of 3..900 residues (below the k-mer span, above the packed SW kernel's 384) with X residues, long planted queries: prefilter,
    structurealign (-a 1) and search with the workflow's parameter strings against live runs of the reference's modules"""
    synth.gen_edge_case_dbs(str(tmp_path))
    p = _params()
    al = list(p["structurealign"]); al[al.index("-a") + 1] = "1"
    _run(REFSEEK, tmp_path, "prefilter", "q_ss", "t_ss", "pref_ref", *p["prefilter"])
    _run(MOCK, tmp_path, "prefilter", "q_ss", "t_ss", "pref_got", *p["prefilter"])
    assert _same(tmp_path, "pref_got", "pref_ref") >= 1000
    _run(REFSEEK, tmp_path, "structurealign", "q", "t", "pref_ref", "aln_ref", *al)
    _run(MOCK, tmp_path, "structurealign", "q", "t", "pref_ref", "aln_got", *al)
    assert _same(tmp_path, "aln_got", "aln_ref") >= 10
    _run(MOCK, tmp_path, "search", "q", "t", "res", "tmp", "-a", "1")
    _same(tmp_path, "res", "aln_ref")


@pytest.mark.skipif(not os.path.exists(REFSEEK), reason="oracle/_ref/refseek (the reference compiled from /root/reference) not present")
def test_exhaustive_search(work):
    """--exhaustive-search 1 (PREFMODE=EXHAUSTIVE): no prefilter; the workflow's fake_pref makes the target DB's index the candidate list of
    every query (data/structuresearch.sh:27-41).  The structurealign module on such a fake prefilter DB and `search --exhaustive-search 1`
    write what the reference's structurealign writes for it."""
    size = os.path.getsize(str(work / "db_ss.index"))
    os.symlink(str(work / "db_ss.index"), str(work / "fp"))
    with open(str(work / "fp.index"), "w") as f:
        for line in open(str(work / "q_ss.index")):
            f.write("%s\t0\t%d\n" % (line.split()[0], size))
    open(str(work / "fp.dbtype"), "wb").write(bytes([7, 0, 0, 0]))
    p = _params()
    _run(REFSEEK, work, "structurealign", "q", "db", "fp", "aln_ref", *p["structurealign"])
    _run(MOCK, work, "structurealign", "q", "db", "fp", "aln_got", *p["structurealign"])
    assert _same(work, "aln_got", "aln_ref") >= 29
    _run(MOCK, work, "search", "q", "db", "res", "tmp", "--exhaustive-search", "1")
    _same(work, "res", "aln_ref")
