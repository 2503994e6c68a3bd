"""The C-ABI library loads and exports every symbol include/fsb200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "foldseek_b200", "libfsb200.so")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    return sorted(set(re.findall(r"\b(fsb200_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_python_binding_lists_the_same_symbols():
    from foldseek_b200 import api
    assert set(api.EXPORTS) <= set(declared_symbols())


def test_default_config_matches_reference_defaults(lib):
    from foldseek_b200 import api
    cfg = api.Config()
    lib.fsb200_default_config(ctypes.byref(cfg))
    assert (cfg.gap_open, cfg.gap_extend, cfg.kmer_size, cfg.max_seqs, cfg.min_diag_score) == (10, 1, 6, 1000, 30)
    assert abs(cfg.sensitivity - 9.5) < 1e-6 and cfg.evalue_thr == 10.0 and abs(cfg.comp_bias_scale_prefilter - 0.15) < 1e-6


def test_create_fails_loudly_without_gpu(lib):
    """no CPU fallback: on a machine without a CUDA device the context cannot be created"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from foldseek_b200 import api
    with pytest.raises(RuntimeError):
        api.Context(0)


def test_library_was_built_from_the_current_sources(lib):
    """the binary reports the hash of the sources it was compiled from (no reliance on modification times): a stale libfsb200.so
    cannot pass for HEAD"""
    import __graft_entry__ as g
    lib.fsb200_build_id.restype = ctypes.c_char_p
    assert lib.fsb200_build_id().decode() == g.source_id()


def test_python_struct_layouts_equal_the_headers(tmp_path):
    """foldseek_b200/api.py mirrors the structs of include/fsb200.h by hand (ctypes / numpy): sizes and field offsets must be the C
    compiler's, or every call through the binding reads the wrong bytes.  Checked against a C program compiled from the header."""
    import subprocess
    import numpy as np
    from foldseek_b200 import api
    hdr = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    structs = {"fsb200_config": api.Config, "fsb200_aln": api.ALN_DTYPE, "fsb200_bt": api.BT_DTYPE, "fsb200_sscore": api.SSCORE_DTYPE,
               "fsb200_hit": api.HIT_DTYPE}
    lines = []
    for name in structs:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                for f in decl.split(","):
                    fields.append(re.findall(r"\**\s*([A-Za-z_0-9]+)\s*$", f.strip())[0])
        lines.append('printf("%s %%zu", sizeof(%s));' % (name, name))
        lines += ['printf(" %s:%%zu", offsetof(%s, %s));' % (f, name, f) for f in fields]
        lines.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fsb200.h"\nint main(void) {\n%s\nreturn 0; }\n' % "\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        tok = line.split()
        name, size, offs = tok[0], int(tok[1]), [t.split(":") for t in tok[2:]]
        py = structs[name]
        if isinstance(py, np.dtype):
            assert py.itemsize == size, name
            assert [(f, py.fields[f][1]) for f in py.names] == [(f, int(o)) for f, o in offs], name
        else:
            assert ctypes.sizeof(py) == size, name
            assert [(f, getattr(py, f).offset) for f, _ in py._fields_] == [(f, int(o)) for f, o in offs], name
