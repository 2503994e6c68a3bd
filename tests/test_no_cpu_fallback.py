"""CPU tests: the product path has no CPU fallback. Without a CUDA device the C ABI refuses to create a context and the host
modules stop with an error instead of computing anything (errors map to Debug(ERROR) + EXIT_FAILURE)."""
import ctypes as C
import os
import subprocess

import pytest

from foldseek_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "foldseek_b200", "bin", "foldseek-b200")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="needs a machine without a CUDA device")
def test_create_fails_without_cuda_device():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        api.Context(0)


@pytest.mark.skipif(_has_gpu() or not os.path.exists(CLI), reason="needs a machine without a CUDA device and the built host module")
def test_host_modules_stop_without_cuda_device(tmp_path):
    db = synth.gen_db(20, 80, 1, min_len=40)
    synth.write_foldseek_db(str(tmp_path / "t"), db, "t")
    for mod, args in (("prefilter", ["t_ss", "t_ss", "pref"]), ("structurealign", ["t", "t", "pref", "aln", "--sort-by-structure-bits", "0"])):
        if mod == "structurealign":
            synth._write_db(str(tmp_path / "pref"), [b"0\t100\t0\n"], 7)      # one prefilter entry: query 0 -> target 0
        r = subprocess.run([CLI, mod] + args, cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode != 0 and "Error" in r.stderr, (mod, r.stderr)
        assert not os.path.exists(tmp_path / (args[-1] if mod == "prefilter" else "aln.dbtype"))


def test_oracle_is_not_linked_into_the_product_library():
    """libfsb200.so must not contain or depend on the oracle (test infrastructure)."""
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "fso_" not in out
    deps = subprocess.run(["ldd", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in deps
    if os.path.exists(CLI):     # the host module links the CUDA library, never the oracle-backed stand-in (oracle/mock_fsb200.cpp, tests only)
        deps = subprocess.run(["ldd", CLI], capture_output=True, text=True).stdout
        assert "liboracle" not in deps and "libfsb200" in deps
