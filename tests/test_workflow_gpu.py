"""Drop-in check at the workflow level: the reference's own search workflow script (data/structuresearch.sh, taken at test time
from the copy the oracle build places next to refseek: oracle/_ref/structuresearch.sh, a build artefact that is not committed) is run
UNCHANGED with $MMSEQS pointing at a dispatcher that sends `prefilter` and `structurealign` to foldseek-b200 (mvdb / rmdb, the
two file-management calls of the script, are done by the dispatcher).  The result DB must equal what the reference modules wrote
for the same DBs and the same serialised parameter strings (tests/golden/example/ref_aln_q)."""
import os
import shutil
import stat
import subprocess

import pytest

from foldseek_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "foldseek_b200", "bin", "foldseek-b200")
EX = os.path.join(ROOT, "tests", "golden", "example")
HDR = os.path.join(ROOT, "oracle", "_ref", "structuresearch.sh")     # copied next to refseek by oracle/build.sh (build artefact, not committed)

DISPATCH = """#!/bin/sh -e
# $MMSEQS of the workflow: the two hot-path modules go to foldseek-b200, the DB file management stays trivial
mod="$1"; shift
case "$mod" in
  prefilter|structurealign) exec "%s" "$mod" "$@" ;;
  mvdb) for f in "$1" "$1".index "$1".dbtype; do [ -e "$f" ] && mv "$f" "$2${f#$1}"; done ;;
  rmdb) rm -f "$1" "$1".index "$1".dbtype "$1".[0-9]* ;;
  *) echo "dispatcher: module $mod is outside the hot path" >&2; exit 1 ;;
esac
"""


def run_search_workflow(binary, tmp_path, extra_par="", extra_env=None, copy=True):
    """runs the reference's structuresearch.sh on the example DBs with `binary` behind $MMSEQS; returns (directory, result DB rows)"""
    script = open(HDR, "rb").read()
    d = tmp_path / "wf"
    if copy:
        shutil.copytree(EX, str(d))
    (d / "structuresearch.sh").write_bytes(script)
    disp = d / "mmseqs-dispatch"
    disp.write_text(DISPATCH % binary)
    disp.chmod(disp.stat().st_mode | stat.S_IEXEC)
    p = {}
    for line in open(os.path.join(EX, "params.txt")):
        mod, _, rest = line.strip().partition(" ")
        p[mod] = rest + extra_par
    tmp = d / "tmp"
    tmp.mkdir()
    env = dict(os.environ, MMSEQS=str(disp), RUNNER="", TMP_PATH=str(tmp), RESULTS=str(d / "result"), INDEXEXT="", PREFMODE="KMER",
               QUERY_PREFILTER=str(d / "q_ss"), TARGET_PREFILTER=str(d / "db_ss"), PREFILTER_PAR=p["prefilter"],
               QUERY_ALIGNMENT=str(d / "q"), TARGET_ALIGNMENT=str(d / "db"), ALIGNMENT_ALGO="structurealign", ALIGNMENT_PAR=p["structurealign"],
               VERBOSITY="-v 3", REMOVE_TMP="1", EXPAND="")
    env.update(extra_env or {})
    r = subprocess.run(["sh", "-e", str(d / "structuresearch.sh")], cwd=str(d), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not os.path.exists(str(tmp / "pref.dbtype"))      # REMOVE_TMP honoured through the dispatcher
    return d, synth.read_result_db(str(d / "result"))


@pytest.mark.skipif(not os.path.exists(HDR), reason="oracle/_ref was not built (needs /root/reference at build time)")
def test_reference_search_workflow_script_with_b200_modules(tmp_path):
    d, got = run_search_workflow(CLI, tmp_path)
    ref = synth.read_result_db(os.path.join(EX, "ref_aln_q"))
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert got[k] == ref[k], k
