"""`foldseek-b200 gpuserver <targetDB>` + `prefilter|structurealign ... --gpu-server 1` (the reference's analogue: `gpuserver` keeps a
padded DB on the GPU and `ungappedprefilter --gpu-server 1` talks to it through shared memory, lib/mmseqs/src/util/gpuserver.cpp,
prefiltering/ungappedprefilter.cpp:72-141).  Here without a GPU: the host module linked against the oracle-backed stand-in for the C
ABI (oracle/_mock, test infrastructure).  The module calls served by the resident process must write the result DBs the reference's
modules wrote (tests/golden/example/ref_*), report errors and exit codes to the caller, survive a failed call, and stage the target
once for a series of calls."""
import os
import shutil
import signal
import subprocess
import time

import pytest

from foldseek_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "oracle", "_mock", "foldseek-b200-mock")
EX = os.path.join(ROOT, "tests", "golden", "example")

pytestmark = pytest.mark.skipif(not os.path.exists(MOCK), reason="oracle/_mock/foldseek-b200-mock not built (sh oracle/build.sh)")


def _params():
    p = {}
    for line in open(os.path.join(EX, "params.txt")):
        mod, _, rest = line.strip().partition(" ")
        p[mod] = rest.split()
    return p


def _same(work, got, ref):
    g, e = synth.read_result_db(str(work / got)), synth.read_result_db(str(work / ref))
    assert g == e, got
    return sum(len(v) for v in e.values())


@pytest.fixture()
def served(tmp_path, sockdir):
    work = tmp_path / "ex"
    shutil.copytree(EX, str(work))
    env = dict(os.environ, FSB200_DATA=os.path.join(ROOT, "foldseek_b200", "data"), FSB200_SOCKET_DIR=sockdir)
    srv = subprocess.Popen([MOCK, "gpuserver", "db", "--lanes", "2"], cwd=str(work), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)

    def call(*args, cwd=None):
        return subprocess.run([MOCK] + list(args) + ["--gpu-server", "1", "--gpu-server-wait-timeout", "60"], cwd=str(cwd or work), env=env,
                              capture_output=True, text=True, timeout=300)
    yield work, call, srv, env
    if srv.poll() is None:
        srv.send_signal(signal.SIGTERM)
        srv.wait(timeout=30)


def test_module_calls_served_by_the_resident_process(served, tmp_path):
    work, call, srv, env = served
    p = _params()
    r = call("prefilter", "db_ss", "db_ss", "pref_all", *p["prefilter"])          # waits for the server, then the preloaded state serves it
    assert r.returncode == 0, r.stderr
    assert "served 1 prefilter call" in r.stdout and "k-mers per position" in r.stdout
    assert _same(work, "pref_all", "ref_pref_all") >= 600
    r = call("prefilter", "q_ss", "db_ss", "pref_q", *p["prefilter"])
    assert r.returncode == 0 and "served 2 prefilter call" in r.stdout, r.stdout + r.stderr
    assert _same(work, "pref_q", "ref_pref_q") >= 20
    r = call("structurealign", "db", "db", "ref_pref_all", "aln_all", *p["structurealign"])
    assert r.returncode == 0 and "served 1 structurealign call" in r.stdout, r.stdout + r.stderr
    assert _same(work, "aln_all", "ref_aln_all") >= 100
    # a caller in another directory with relative paths: the server works in the caller's directory
    other = tmp_path / "elsewhere"
    other.mkdir()
    rel = os.path.relpath(str(work), str(other))
    r = call("structurealign", rel + "/q", rel + "/db", rel + "/ref_pref_q", "aln_q", *p["structurealign"], cwd=other)
    assert r.returncode == 0 and "served 2 structurealign call" in r.stdout, r.stdout + r.stderr
    assert synth.read_result_db(str(other / "aln_q")) == synth.read_result_db(str(work / "ref_aln_q"))
    # a failing call: message and exit code reach the caller, the server goes on
    r = call("structurealign", "q", "db", "no_such_pref", "aln_x", *p["structurealign"])
    assert r.returncode != 0 and "no_such_pref" in r.stderr
    r = call("prefilter", "q_ss", "db_ss", "pref_x", "--exact-kmer-matching", "1")
    assert r.returncode != 0 and "--exact-kmer-matching" in r.stderr
    # parameters that matter for the staged state replace it (here: the E-value threshold the contexts were created with)
    al = list(p["structurealign"]); al[al.index("-e") + 1] = "0.001"
    r = call("structurealign", "db", "db", "ref_pref_all", "aln_e", *al)
    assert r.returncode == 0 and "served 1 structurealign call" in r.stdout, r.stdout + r.stderr
    ref = synth.read_result_db(str(work / "ref_aln_all"))
    got = synth.read_result_db(str(work / "aln_e"))
    assert sorted(got) == sorted(ref) and all(r in ref[k] for k in ref for r in got[k]) and sum(map(len, got.values())) < sum(map(len, ref.values()))
    # alternative alignments through the server (second context per lane on the server's GPU)
    alt = list(p["structurealign"])
    for flag, v in (("--alt-ali", "2"), ("-a", "1")):
        alt[alt.index(flag) + 1] = v
    r = call("structurealign", "db", "db", "ref_pref_all", "aln_alt", *alt)
    assert r.returncode == 0, r.stderr
    assert _same(work, "aln_alt", "ref_aln_all_alt2") >= 950
    assert srv.poll() is None
    srv.send_signal(signal.SIGTERM)
    assert srv.wait(timeout=30) == 0
    assert not [f for f in os.listdir(env["FSB200_SOCKET_DIR"]) if f.endswith(".sock")]     # socket removed at shutdown


def test_no_server_no_fallback(tmp_path, sockdir):
    """--gpu-server 1 without a server: the call fails naming the database (ungappedprefilter.cpp:85-93); it does not run locally"""
    work = tmp_path / "ex"
    shutil.copytree(EX, str(work))
    env = dict(os.environ, FSB200_DATA=os.path.join(ROOT, "foldseek_b200", "data"), FSB200_SOCKET_DIR=sockdir)
    t0 = time.time()
    r = subprocess.run([MOCK, "prefilter", "q_ss", "db_ss", "pref_q", "--gpu-server", "1", "--gpu-server-wait-timeout", "0"], cwd=str(work), env=env,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "gpuserver for database db_ss not found" in r.stderr
    assert not os.path.exists(str(work / "pref_q.dbtype")) and time.time() - t0 < 20
    r = subprocess.run([MOCK, "prefilter", "q_ss", "db_ss", "pref_q", "--gpu-server", "1", "--gpu-server-wait-timeout", "1"], cwd=str(work), env=env,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "not found after" in r.stderr and "Waiting for `gpuserver`" in r.stderr


def test_second_server_for_the_same_database_refuses(served):
    work, call, srv, env = served
    for _ in range(600):                                                        # the socket appears once the first server has staged the DB
        if [f for f in os.listdir(env["FSB200_SOCKET_DIR"]) if f.endswith(".sock")] or srv.poll() is not None:
            break
        time.sleep(0.1)
    assert srv.poll() is None
    r = subprocess.run([MOCK, "gpuserver", "db"], cwd=str(work), env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "already running" in r.stderr
