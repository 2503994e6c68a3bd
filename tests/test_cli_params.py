"""Host module argument handling, without a GPU: `foldseek-b200` must know every parameter the search workflow serialises for the two
modules (tests/golden/example/params.txt = createParameterString of par.prefilter / par.structurealign, Parameters.cpp:472-505,
LocalParameters.cpp:160-167), reject unknown flags and unimplemented values by name before it touches a database or a device, and never
fall back to a CPU path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "foldseek_b200", "bin", "foldseek-b200")
EX = os.path.join(ROOT, "tests", "golden", "example")

pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="host module not built (python -c 'import __graft_entry__ as g; g.build()')")


def _params():
    p = {}
    for line in open(os.path.join(EX, "params.txt")):
        mod, _, rest = line.strip().partition(" ")
        p[mod] = rest.split()
    return p


def _run(*args):
    return subprocess.run([CLI] + list(args), capture_output=True, text=True, cwd="/tmp")


def test_workflow_parameter_lists_are_known():
    """with the full serialised lists the modules get past the parser: what stops them is the missing input DB, not a flag"""
    p = _params()
    for mod, pos in (("prefilter", ["no_such_q", "no_such_t", "out"]), ("structurealign", ["no_such_q", "no_such_t", "no_such_pref", "out"])):
        r = _run(mod, *pos, *p[mod], "--lanes", "2", "--gpu-list", "0")
        assert r.returncode != 0
        assert "unrecognized parameter" not in r.stderr and "not implemented" not in r.stderr, r.stderr
        assert "no_such" in r.stderr, r.stderr


@pytest.mark.parametrize("mod,flags", [
    ("prefilter", ["--no-such-flag", "1"]), ("structurealign", ["--no-such-flag", "1"]),
    ("structurealign", ["--alt-ali", "-1"]), ("structurealign", ["--exact-tmscore", "1"]), ("structurealign", ["--alignment-type", "1"]),
    ("prefilter", ["--exact-kmer-matching", "1"]), ("prefilter", ["-k", "7"]), ("prefilter", ["--mask", "1"]), ("prefilter", ["--lanes", "0"]),
])
def test_unknown_flags_and_unimplemented_values_are_named(mod, flags):
    pos = ["q", "t", "o"] if mod == "prefilter" else ["q", "t", "p", "o"]
    r = _run(mod, *pos, *flags)
    assert r.returncode != 0 and flags[0] in r.stderr, r.stderr


def test_no_cpu_path_behind_the_host_module():
    """on a machine without a CUDA device the module stops at context creation and says why"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = _params()
    r = subprocess.run([CLI, "prefilter", "q_ss", "db_ss", "/tmp/fsb200_no_gpu_pref"] + p["prefilter"], capture_output=True, text=True, cwd=EX)
    assert r.returncode != 0 and "CUDA device is required" in r.stderr, r.stderr
    assert not os.path.exists("/tmp/fsb200_no_gpu_pref.dbtype")


def test_help_lists_the_modules():
    r = _run("-h")
    assert r.returncode == 0 and all(m in r.stdout for m in ("easy-search", "search", "prefilter", "structurealign", "convertalis", "gpuserver"))
    assert "no CPU path" in r.stdout
