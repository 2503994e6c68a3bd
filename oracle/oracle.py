"""ctypes loader for the CPU oracle (oracle/fs_oracle.cpp) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(os.path.dirname(_HERE), "foldseek_b200", "data")
REFSEEK = os.path.join(_HERE, "_ref", "refseek")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _lib():
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fs_oracle.cpp")
    hdr = os.path.join(_HERE, "block_aligner_port.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-mavx2", "-mfma", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.fso_create.restype = C.c_void_p
    lib.fso_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
    lib.fso_destroy.argtypes = [C.c_void_p]
    lib.fso_get_matrix.argtypes = [C.c_void_p, C.c_int, _i16p, _f64p]
    lib.fso_query_bias.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_float, _i8p, _i8p]
    lib.fso_mu_lambda.argtypes = [C.c_void_p, C.c_uint64, _u8p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.fso_align_query.argtypes = [C.c_void_p, C.c_uint64, _u8p, _u8p, C.c_int, C.c_float, _u8p, _u8p, _u64p, _u32p,
                                    C.c_int, C.c_int, C.c_int, _i32p, _f64p]
    lib.fso_index_create.restype = C.c_void_p
    lib.fso_index_create.argtypes = [C.c_void_p, _u8p, _u64p, _u32p, C.c_uint32, C.c_float]
    lib.fso_index_destroy.argtypes = [C.c_void_p]
    lib.fso_index_kmer_thr.argtypes = [C.c_void_p]
    lib.fso_index_entries.restype = C.c_uint64
    lib.fso_index_entries.argtypes = [C.c_void_p]
    lib.fso_prefilter_query.restype = C.c_uint32
    lib.fso_prefilter_query.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, _u8p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_uint32, _i32p, _u64p]
    lib.fso_ba_align.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _i8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, _i32p, C.c_char_p, C.c_int]
    lib.fso_ba_prefix_scan.argtypes = [_i16p, C.c_int, _i16p]
    lib.fso_backtrace.restype = C.c_int
    lib.fso_backtrace.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_float, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int32, _i32p, C.c_char_p, C.c_int]
    lib.fso_lddt.restype = C.c_double
    lib.fso_lddt.argtypes = [np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int,
                             np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int, C.c_int, C.c_int, C.c_char_p,
                             C.c_void_p, C.c_void_p]
    lib.fso_tmscore.restype = C.c_double
    lib.fso_tmscore.argtypes = [np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int,
                                np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int, C.c_int, C.c_int, C.c_char_p,
                                C.c_int, C.c_void_p]
    return lib


def tmscore(q_ca, t_ca, q_start, t_start, backtrace, norm_len):
    """tmscore of TMaligner::computeTMscore (approximate mode); norm_len = TMaligner::normalization(...)."""
    lib = _lib()
    q = np.ascontiguousarray(q_ca, np.float32)
    t = np.ascontiguousarray(t_ca, np.float32)
    return float(lib.fso_tmscore(q.reshape(-1), q.shape[1], t.reshape(-1), t.shape[1], int(q_start), int(t_start), backtrace.encode(),
                                 int(norm_len), None))


def lddt(q_ca, t_ca, q_start, t_start, backtrace):
    """avgLddtScore of LDDTCalculator::computeLDDTScore; q_ca / t_ca: float32 [3, L] (x, y, z rows)."""
    lib = _lib()
    q = np.ascontiguousarray(q_ca, np.float32)
    t = np.ascontiguousarray(t_ca, np.float32)
    return float(lib.fso_lddt(q.reshape(-1), q.shape[1], t.reshape(-1), t.shape[1], int(q_start), int(t_start), backtrace.encode(),
                              None, None))


def ba_prefix_scan(values16, gap):
    out = np.zeros(16, np.int16)
    _lib().fso_ba_prefix_scan(np.ascontiguousarray(values16, np.int16), int(gap), out)
    return out.tolist()


def ba_align(q, r, table, stride, null_code, gap_open, gap_extend, min_size, max_size, x_drop=0, trace=False, xdrop=False,
             eq=False):
    """Block-aligner engine restatement on converted byte strings; returns (score, query_idx, reference_idx, cigar)."""
    lib = _lib()
    out = np.zeros(3, np.int32)
    buf = C.create_string_buffer(16 * (len(q) + len(r) + 8))
    qa = np.ascontiguousarray(np.frombuffer(bytes(q), np.uint8))
    ra = np.ascontiguousarray(np.frombuffer(bytes(r), np.uint8))
    lib.fso_ba_align(qa if len(qa) else np.zeros(1, np.uint8), len(q), ra if len(ra) else np.zeros(1, np.uint8), len(r),
                     np.ascontiguousarray(table, np.int8).reshape(-1), stride, null_code, gap_open, gap_extend, min_size,
                     max_size, x_drop, (1 if trace else 0) | (2 if xdrop else 0) | (4 if eq else 0), out, buf, len(buf))
    return int(out[0]), int(out[1]), int(out[2]), buf.value.decode()


def ba_align_3di(q_aa, q_3di, r_aa, r_3di, tab_aa, tab_3di, stride, null_code, gap_open, gap_extend, min_size, max_size, x_drop=0,
                 trace=False, xdrop=False, eq=False, q_bias=None, r_bias=None):
    """The 3Di+AA entry of the block-aligner restatement (place_block_3di) with caller-supplied tables and position biases."""
    lib = _lib()
    lib.fso_ba_align_3di.restype = None
    i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
    lib.fso_ba_align_3di.argtypes = [_u8p, _u8p, C.c_void_p, C.c_int, _u8p, _u8p, C.c_void_p, C.c_int, _i8p, _i8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), C.c_char_p, C.c_int]
    out = np.zeros(3, np.int32)
    buf = C.create_string_buffer(16 * (len(q_aa) + len(r_aa) + 8))
    arr = lambda b: np.ascontiguousarray(np.frombuffer(bytes(b), np.uint8)) if len(b) else np.zeros(1, np.uint8)
    qb = np.ascontiguousarray(q_bias, np.int16) if q_bias is not None else None
    rb = np.ascontiguousarray(r_bias, np.int16) if r_bias is not None else None
    lib.fso_ba_align_3di(arr(q_aa), arr(q_3di), qb.ctypes.data if qb is not None else None, len(q_aa), arr(r_aa), arr(r_3di),
                         rb.ctypes.data if rb is not None else None, len(r_aa),
                         np.ascontiguousarray(tab_aa, np.int8).reshape(-1), np.ascontiguousarray(tab_3di, np.int8).reshape(-1), stride, null_code,
                         gap_open, gap_extend, min_size, max_size, x_drop, (1 if trace else 0) | (2 if xdrop else 0) | (4 if eq else 0), out, buf, len(buf))
    return int(out[0]), int(out[1]), int(out[2]), buf.value.decode()


def data_file(name):
    return open(os.path.join(_DATA, name), "rb").read()


class Oracle:
    def __init__(self):
        self.lib = _lib()
        nn = data_file("evalue_nn.kerasify")
        self.h = self.lib.fso_create(data_file("mat3di.out"), data_file("blosum62.out"), nn, len(nn))
        assert self.h, "oracle: matrix parse failed"

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fso_destroy(self.h)
            self.h = None

    def matrix(self, which):
        m = np.zeros(441, np.int16)
        pb = np.zeros(21, np.float64)
        self.lib.fso_get_matrix(self.h, which, m, pb)
        return m.reshape(21, 21), pb

    def query_bias(self, q_aa, q_3di, scale=0.5):
        L = len(q_aa)
        a = np.zeros(L, np.int8)
        s = np.zeros(L, np.int8)
        self.lib.fso_query_bias(self.h, np.ascontiguousarray(q_aa), np.ascontiguousarray(q_3di), L, scale, a, s)
        return a, s

    def backtrace(self, q_aa, q_3di, t_aa, t_3di, q_end, db_end, target_score, scale=0.5, go=10, ge=1):
        """alignStartPosBacktraceBlock: returns (ok, q_start, db_start, identical_aa, block_score, backtrace)."""
        out = np.zeros(4, np.int32)
        buf = C.create_string_buffer(len(q_aa) + len(t_aa) + 16)
        ok = self.lib.fso_backtrace(self.h, np.ascontiguousarray(q_aa), np.ascontiguousarray(q_3di), len(q_aa), scale,
                                    np.ascontiguousarray(t_aa), np.ascontiguousarray(t_3di), len(t_aa), go, ge, int(q_end),
                                    int(db_end), int(target_score), out, buf, len(buf))
        return bool(ok), int(out[0]), int(out[1]), int(out[2]), int(out[3]), buf.value.decode()

    def mu_lambda(self, db_residues, q_3di):
        a, b = C.c_double(), C.c_double()
        self.lib.fso_mu_lambda(self.h, db_residues, np.ascontiguousarray(q_3di), len(q_3di), C.byref(a), C.byref(b))
        return a.value, b.value

    def align_query(self, db_residues, q_aa, q_3di, t_aa_list, t_3di_list, go=10, ge=1, bias_scale=0.5):
        """Returns (int32 [n,4] = fwdScore,qEnd,dbEnd,revScore ; float64 [n,2] = E(fwd), E(fwd-rev))."""
        n = len(t_aa_list)
        lens = np.array([len(t) for t in t_aa_list], np.uint32)
        offs = np.zeros(n, np.uint64)
        if n:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        taa = np.concatenate(t_aa_list).astype(np.uint8) if n else np.zeros(1, np.uint8)
        tss = np.concatenate(t_3di_list).astype(np.uint8) if n else np.zeros(1, np.uint8)
        out = np.zeros((max(n, 1), 4), np.int32)
        ev = np.zeros((max(n, 1), 2), np.float64)
        self.lib.fso_align_query(self.h, db_residues, np.ascontiguousarray(q_aa), np.ascontiguousarray(q_3di), len(q_aa),
                                 bias_scale, taa, tss, offs, lens, n, go, ge, out, ev)
        return out[:n], ev[:n]


class PrefilterIndex:
    """Oracle k-mer index + per-query prefilter (stage 1)."""

    def __init__(self, oracle, t_3di_list, sensitivity=9.5):
        self.o = oracle
        lens = np.array([len(t) for t in t_3di_list], np.uint32)
        offs = np.zeros(len(lens), np.uint64)
        if len(lens) > 1:
            offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        flat = np.ascontiguousarray(np.concatenate(t_3di_list).astype(np.uint8))
        self.n = len(lens)
        self.ix = oracle.lib.fso_index_create(oracle.h, flat, offs, lens, self.n, sensitivity)

    def __del__(self):
        if getattr(self, "ix", None):
            self.o.lib.fso_index_destroy(self.ix)
            self.ix = None

    def query(self, q_3di, max_seqs=1000, min_diag_score=30, bins=2, bias_scale=0.15, identity=0xFFFFFFFF):
        out = np.zeros((max(max_seqs, 1), 3), np.int32)
        st = np.zeros(4, np.uint64)
        n = self.o.lib.fso_prefilter_query(self.o.h, self.ix, self.n, np.ascontiguousarray(q_3di), len(q_3di), bias_scale,
                                           max_seqs, min_diag_score, bins, identity, out, st)
        return out[:n].copy(), st
