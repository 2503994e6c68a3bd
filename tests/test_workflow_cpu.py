"""The workflow-level drop-in check of tests/test_workflow_gpu.py without a GPU: the reference's own data/structuresearch.sh, unchanged,
with $MMSEQS dispatching `prefilter` and `structurealign` to the host module on the oracle-backed stand-in for the C ABI (oracle/_mock,
test infrastructure) - once as plain module processes, once with --gpu-server 1 against a running `gpuserver` (the way
`foldseek search --gpu-server 1` reaches the reference's resident GPU process)."""
import os
import shutil
import signal
import subprocess

import pytest

from foldseek_b200 import synth
from test_workflow_gpu import EX, HDR, ROOT, run_search_workflow

MOCK = os.path.join(ROOT, "oracle", "_mock", "foldseek-b200-mock")
pytestmark = pytest.mark.skipif(not (os.path.exists(MOCK) and os.path.exists(HDR)), reason="oracle/_mock or oracle/_ref not built (sh oracle/build.sh)")
DATA = os.path.join(ROOT, "foldseek_b200", "data")


def test_reference_search_workflow_script_on_the_stand_in(tmp_path):
    d, got = run_search_workflow(MOCK, tmp_path, extra_env={"FSB200_DATA": DATA})
    assert got == synth.read_result_db(os.path.join(EX, "ref_aln_q"))


def test_reference_search_workflow_script_through_the_gpuserver(tmp_path, sockdir):
    env = dict(os.environ, FSB200_DATA=DATA, FSB200_SOCKET_DIR=sockdir)
    srv = subprocess.Popen([MOCK, "gpuserver", os.path.join(EX, "db")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        # the workflow's copy of the DBs lives in another directory than the one the server was started on: another database, so this
        # call must not find a server ...
        with pytest.raises(AssertionError, match="gpuserver for database"):
            run_search_workflow(MOCK, tmp_path / "a", extra_par=" --gpu-server 1 --gpu-server-wait-timeout 0", extra_env=env)
    finally:
        srv.send_signal(signal.SIGTERM)
        srv.wait(timeout=30)
    # ... and finds the one started on its own copy
    (tmp_path / "b").mkdir()
    shutil.copytree(EX, str(tmp_path / "b" / "wf"))
    srv = subprocess.Popen([MOCK, "gpuserver", str(tmp_path / "b" / "wf" / "db")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        d, got = run_search_workflow(MOCK, tmp_path / "b", extra_par=" --gpu-server 1 --gpu-server-wait-timeout 60", extra_env=env, copy=False)
        assert got == synth.read_result_db(os.path.join(EX, "ref_aln_q"))
        assert srv.poll() is None
    finally:
        srv.send_signal(signal.SIGTERM)
        out, err = srv.communicate(timeout=30)
    assert srv.returncode == 0, err
