"""ctypes binding of libfsb200.so (include/fsb200.h) used by the tests, bench.py and the Python-side tools.

The product path is the CUDA library: importing this module without the built library, or creating a
context without a CUDA device, raises — there is no CPU fallback.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSB200_LIB", os.path.join(_HERE, "libfsb200.so"))
DATA_DIR = os.path.join(_HERE, "data")

EXPORTS = [
    "fsb200_default_config", "fsb200_create", "fsb200_destroy", "fsb200_last_error", "fsb200_get_matrix",
    "fsb200_load_target", "fsb200_set_queries", "fsb200_query_info", "fsb200_align", "fsb200_align_upload",
    "fsb200_align_run", "fsb200_align_download", "fsb200_align_stats", "fsb200_prefilter", "fsb200_prefilter_stats",
    "fsb200_search_run", "fsb200_search_download", "fsb200_search", "fsb200_timer_start", "fsb200_timer_stop",
    "fsb200_staged_target_size", "fsb200_export_staged_target", "fsb200_load_target_device",
    "fsb200_backtrace_run", "fsb200_backtrace_size", "fsb200_backtrace_download", "fsb200_backtrace_stats",
    "fsb200_load_target_ca", "fsb200_set_queries_ca", "fsb200_structure_scores_run", "fsb200_structure_scores_download",
    "fsb200_structure_scores_stats", "fsb200_prefilter_stage_ns", "fsb200_set_identity", "fsb200_test_block_align",
    "fsb200_build_index", "fsb200_create_lane", "fsb200_set_host_threads", "fsb200_build_id",
]


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("mat3di_text", C.c_char_p), ("mataa_text", C.c_char_p),
        ("evalue_nn", C.c_char_p), ("evalue_nn_len", C.c_uint64),
        ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("comp_bias_corr", C.c_int32),
        ("comp_bias_scale_align", C.c_float), ("comp_bias_scale_prefilter", C.c_float),
        ("sensitivity", C.c_float), ("kmer_size", C.c_int32), ("max_seqs", C.c_int32), ("min_diag_score", C.c_int32),
        ("evalue_thr", C.c_double), ("aa_factor", C.c_float), ("diag_bin_count", C.c_int32), ("tmscore_norm_mode", C.c_int32),
    ]


ALN_DTYPE = np.dtype([("fwd_score", np.uint32), ("rev_score", np.uint32), ("q_end", np.int32), ("db_end", np.int32),
                      ("status", np.int32), ("score", np.int32), ("evalue", np.float64)], align=True)
BT_DTYPE = np.dtype([("q_start", np.int32), ("db_start", np.int32), ("identical_aa", np.uint32), ("aln_len", np.uint32),
                     ("bt_off", np.uint64), ("status", np.int32), ("block_score", np.int32)], align=True)
SSCORE_DTYPE = np.dtype([("tm", np.float64), ("lddt", np.float32), ("status", np.int32), ("score_bits", np.int32)], align=True)
HIT_DTYPE = np.dtype([("target", np.uint32), ("score", np.int32), ("diagonal", np.int32)], align=True)


def load_library():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libfsb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    lib = C.CDLL(LIB_PATH)
    lib.fsb200_last_error.restype = C.c_char_p
    for name in EXPORTS:
        getattr(lib, name)
    return lib


def _data(name):
    with open(os.path.join(DATA_DIR, name), "rb") as f:
        return f.read()


def flatten(seqs):
    """list of uint8 arrays -> (concatenated uint8, uint64 offsets, uint32 lengths)"""
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offs = np.zeros(len(seqs), dtype=np.uint64)
    if len(seqs) > 1:
        offs[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    flat = np.concatenate(seqs).astype(np.uint8) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(flat), offs, lens


class Context:
    def __init__(self, device=0, **overrides):
        self.lib = load_library()
        cfg = Config()
        self.lib.fsb200_default_config(C.byref(cfg))
        self._keep = (_data("mat3di.out"), _data("blosum62.out"), _data("evalue_nn.kerasify"))
        cfg.device = device
        cfg.mat3di_text, cfg.mataa_text = self._keep[0], self._keep[1]
        cfg.evalue_nn, cfg.evalue_nn_len = self._keep[2], len(self._keep[2])
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.lib.fsb200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError("fsb200_create failed with %d (CUDA device required; no CPU fallback)" % rc)
        self.h = h
        self.nq = 0
        # reuse_buffers: the large result arrays (alignments, hits, backtrace meta) are allocated once per size and handed out again
        # on the next call, the way a C++ host keeps its result buffers; the caller must consume a result before the next call
        self.reuse_buffers = False
        self._bufs = {}
        self._lanes = []

    def _out(self, name, n, dtype):
        if not self.reuse_buffers:
            return np.zeros(n, dtype)
        key = (name, int(n), np.dtype(dtype).str if not isinstance(dtype, np.dtype) or dtype.names is None else name)
        b = self._bufs.get(key)
        if b is None or len(b) != n:
            b = np.zeros(n, dtype)
            self._bufs[key] = b
        return b

    def close(self):
        for lane in getattr(self, "_lanes", []):      # lanes borrow this context's device arrays: they go first
            lane.close()
        self._lanes = []
        if getattr(self, "h", None):
            self.lib.fsb200_destroy(self.h)
            self.h = None

    def build_index(self):
        self._check(self.lib.fsb200_build_index(self.h))

    def set_host_threads(self, n):
        self._check(self.lib.fsb200_set_host_threads(self.h, C.c_int32(int(n))))

    def create_lane(self):
        """Another context on the same GPU that shares this context's staged target, C-alpha and k-mer index (fsb200_create_lane)."""
        h = C.c_void_p()
        self._check(self.lib.fsb200_create_lane(self.h, C.byref(h)))
        lane = Context.__new__(Context)
        lane.lib, lane.cfg, lane._keep, lane.h, lane.nq = self.lib, self.cfg, self._keep, h, 0
        lane.nt = self.nt
        lane.reuse_buffers, lane._bufs, lane._lanes = self.reuse_buffers, {}, []
        self._lanes.append(lane)
        return lane

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("fsb200 error %d: %s" % (rc, self.lib.fsb200_last_error(self.h).decode()))

    def matrix(self, which):
        m = np.zeros(441, np.int16)
        self._check(self.lib.fsb200_get_matrix(self.h, which, m.ctypes.data_as(C.c_void_p)))
        return m.reshape(21, 21)

    def load_target(self, aa_list, tdi_list, aa_db_residues=None):
        aa, off, ln = flatten(aa_list)
        tdi, _, _ = flatten(tdi_list)
        if aa_db_residues is None:
            aa_db_residues = int(ln.sum())
        self.load_target_flat(aa, tdi, off, ln, aa_db_residues)

    def load_target_flat(self, aa, tdi, off, ln, aa_db_residues):
        self._check(self.lib.fsb200_load_target(self.h, aa.ctypes.data_as(C.c_void_p), tdi.ctypes.data_as(C.c_void_p),
                                                off.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                                                C.c_uint32(len(ln)), C.c_uint64(aa_db_residues)))
        self.nt = len(ln)

    def set_queries(self, aa_list, tdi_list):
        aa, off, ln = flatten(aa_list)
        tdi, _, _ = flatten(tdi_list)
        self.set_queries_flat(aa, tdi, off, ln)

    def set_queries_flat(self, aa, tdi, off, ln):
        self._check(self.lib.fsb200_set_queries(self.h, aa.ctypes.data_as(C.c_void_p), tdi.ctypes.data_as(C.c_void_p),
                                                off.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p), C.c_uint32(len(ln))))
        self.nq = len(ln)
        self.qlen = ln.copy()

    def query_info(self, q):
        ml = np.zeros(2, np.float64)
        ba = np.zeros(int(self.qlen[q]), np.int8)
        bs = np.zeros(int(self.qlen[q]), np.int8)
        self._check(self.lib.fsb200_query_info(self.h, C.c_uint32(q), ml.ctypes.data_as(C.c_void_p),
                                               ba.ctypes.data_as(C.c_void_p), bs.ctypes.data_as(C.c_void_p)))
        return ml, ba, bs

    @staticmethod
    def csr(cand_lists):
        off = np.zeros(len(cand_lists) + 1, np.uint64)
        off[1:] = np.cumsum([len(c) for c in cand_lists], dtype=np.uint64)
        flat = np.concatenate([np.asarray(c, np.uint32) for c in cand_lists]) if len(cand_lists) else np.zeros(0, np.uint32)
        return off, np.ascontiguousarray(flat, dtype=np.uint32)

    def align(self, cand_off, cand_target):
        out = np.zeros(int(cand_off[-1]), ALN_DTYPE)
        self._check(self.lib.fsb200_align(self.h, cand_off.ctypes.data_as(C.c_void_p), cand_target.ctypes.data_as(C.c_void_p),
                                          out.ctypes.data_as(C.c_void_p)))
        return out

    def align_upload(self, cand_off, cand_target):
        self._n_cand = int(cand_off[-1])
        self._check(self.lib.fsb200_align_upload(self.h, cand_off.ctypes.data_as(C.c_void_p), cand_target.ctypes.data_as(C.c_void_p)))

    def align_run(self):
        self._check(self.lib.fsb200_align_run(self.h))

    def align_download(self):
        out = np.zeros(self._n_cand, ALN_DTYPE)
        self._check(self.lib.fsb200_align_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def align_stats(self):
        s = np.zeros(8, np.uint64)
        self._check(self.lib.fsb200_align_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return s

    @staticmethod
    def _flatten_ca(ca_list):
        lens = np.array([c.shape[1] for c in ca_list], np.uint64)
        off = np.zeros(len(ca_list), np.uint64)
        if len(ca_list) > 1:
            off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        xyz = [np.ascontiguousarray(np.concatenate([c[d] for c in ca_list]).astype(np.float32)) for d in range(3)]
        return xyz, off

    def load_target_ca(self, ca_list):
        """ca_list: float32 [3, L] per target (x, y, z rows), same order and lengths as load_target."""
        (x, y, z), off = self._flatten_ca(ca_list)
        self._check(self.lib.fsb200_load_target_ca(self.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                                   z.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))

    def load_target_ca_flat(self, x, y, z, off):
        """x/y/z float32 arrays; target i reads its residues at off[i] (offsets may alias)"""
        self._keep_ca = (x, y, z, off)
        self._check(self.lib.fsb200_load_target_ca(self.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                                   z.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))

    def set_queries_ca_flat(self, x, y, z, off):
        self._check(self.lib.fsb200_set_queries_ca(self.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                                   z.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))

    def structure_scores_run(self):
        self._check(self.lib.fsb200_structure_scores_run(self.h))

    def set_queries_ca(self, ca_list):
        (x, y, z), off = self._flatten_ca(ca_list)
        self._check(self.lib.fsb200_set_queries_ca(self.h, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                                   z.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)))

    def structure_scores(self, n_pairs):
        """LDDT / TM-score / structure-bits score of the traced pairs of the last backtrace run."""
        self._check(self.lib.fsb200_structure_scores_run(self.h))
        out = self._out("sscore", n_pairs, SSCORE_DTYPE)
        self._check(self.lib.fsb200_structure_scores_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def structure_scores_stats(self):
        s = np.zeros(8, np.uint64)
        self._check(self.lib.fsb200_structure_scores_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return s

    def backtrace_run(self):
        """Traceback of the accepted pairs of the last align/search run (results stay in the context)."""
        self._check(self.lib.fsb200_backtrace_run(self.h))

    def backtrace_download(self, n_pairs):
        nbytes = C.c_uint64(0)
        self._check(self.lib.fsb200_backtrace_size(self.h, C.byref(nbytes)))
        meta = self._out("bt_meta", n_pairs, BT_DTYPE)
        bt = self._out("bt", max(1, nbytes.value), np.uint8)
        self._check(self.lib.fsb200_backtrace_download(self.h, meta.ctypes.data_as(C.c_void_p), bt.ctypes.data_as(C.c_void_p),
                                                       C.c_uint64(len(bt))))
        return meta, bt

    def backtrace(self, n_pairs):
        """(meta[n_pairs], bytes of all backtraces); meta is indexed like the output of align()."""
        self.backtrace_run()
        return self.backtrace_download(n_pairs)

    def backtrace_stats(self):
        s = np.zeros(8, np.uint64)
        self._check(self.lib.fsb200_backtrace_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return s

    def prefilter(self):
        ms = int(self.cfg.max_seqs)
        hits = np.zeros(self.nq * ms, HIT_DTYPE)
        counts = np.zeros(self.nq, np.uint32)
        self._check(self.lib.fsb200_prefilter(self.h, hits.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p)))
        return [hits[q * ms:q * ms + int(counts[q])] for q in range(self.nq)]

    def prefilter_stats(self):
        s = np.zeros(8, np.uint64)
        self._check(self.lib.fsb200_prefilter_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return s

    STAGES = ("count_scan", "kmer_lists", "expand", "diag", "sort", "score", "select")

    def prefilter_stage_ms(self):
        """device time of the last prefilter per pipeline stage (CUDA events), in ms"""
        s = np.zeros(8, np.uint64)
        self._check(self.lib.fsb200_prefilter_stage_ns(self.h, s.ctypes.data_as(C.c_void_p)))
        return {n: float(s[i]) / 1e6 for i, n in enumerate(self.STAGES)}

    def set_identity(self, ids):
        """ids[q] = target index of query q itself (query DB == target DB), 0xFFFFFFFF for none; None clears"""
        if ids is None:
            self._check(self.lib.fsb200_set_identity(self.h, None))
            return
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        assert len(a) == self.nq
        self._check(self.lib.fsb200_set_identity(self.h, a.ctypes.data_as(C.c_void_p)))

    def test_block_align(self, q_aa, q_3di, t_aa, t_3di, tab_aa, tab_3di, gap_open, gap_extend, min_size, max_size, x_drop,
                         q_bias_aa=None, q_bias_ss=None):
        """test hook: one run of the CUDA block-aligner engine; returns (score, query_idx, reference_idx, M/I/D path)"""
        qa, qs = np.ascontiguousarray(q_aa, np.uint8), np.ascontiguousarray(q_3di, np.uint8)
        ta, ts = np.ascontiguousarray(t_aa, np.uint8), np.ascontiguousarray(t_3di, np.uint8)
        ba = np.ascontiguousarray(q_bias_aa if q_bias_aa is not None else np.zeros(len(qa)), np.int8)
        bs = np.ascontiguousarray(q_bias_ss if q_bias_ss is not None else np.zeros(len(qa)), np.int8)
        out = np.zeros(3, np.int32)
        buf = C.create_string_buffer(len(qa) + len(ta) + 8)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self._check(self.lib.fsb200_test_block_align(self.h, p(qa), p(qs), p(ba), p(bs), C.c_uint32(len(qa)), p(ta), p(ts), C.c_uint32(len(ta)),
                                                     p(np.ascontiguousarray(tab_aa, np.int8)), p(np.ascontiguousarray(tab_3di, np.int8)),
                                                     C.c_int32(gap_open), C.c_int32(gap_extend), C.c_uint32(min_size), C.c_uint32(max_size),
                                                     C.c_int32(x_drop), p(out), buf, C.c_uint32(len(buf))))
        return int(out[0]), int(out[1]), int(out[2]), buf.value.decode()

    def search_run(self):
        self._check(self.lib.fsb200_search_run(self.h))

    def search_download(self):
        ms = int(self.cfg.max_seqs)
        hits = self._out("hits", self.nq * ms, HIT_DTYPE)
        counts = self._out("counts", self.nq, np.uint32)
        aln = self._out("aln", self.nq * ms, ALN_DTYPE)
        self._check(self.lib.fsb200_search_download(self.h, hits.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                                    aln.ctypes.data_as(C.c_void_p)))
        return hits.reshape(self.nq, ms), counts, aln.reshape(self.nq, ms)

    def search(self):
        self.search_run()
        return self.search_download()

    def timer_start(self):
        self._check(self.lib.fsb200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.fsb200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def staged_target(self):
        """(device ptr aa, device ptr tdi, bytes, off, len) of the staged target DB (for the NCCL broadcast)."""
        b = C.c_uint64()
        self._check(self.lib.fsb200_staged_target_size(self.h, C.byref(b)))
        pa, pt = C.c_void_p(), C.c_void_p()
        off = np.zeros(self.nt, np.uint64)
        ln = np.zeros(self.nt, np.uint32)
        self._check(self.lib.fsb200_export_staged_target(self.h, C.byref(pa), C.byref(pt), off.ctypes.data_as(C.c_void_p),
                                                         ln.ctypes.data_as(C.c_void_p)))
        return pa.value, pt.value, b.value, off, ln

    def load_target_device(self, d_aa, d_tdi, nbytes, off, ln, aa_db_residues):
        self._check(self.lib.fsb200_load_target_device(self.h, C.c_void_p(d_aa), C.c_void_p(d_tdi), C.c_uint64(nbytes),
                                                       off.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                                                       C.c_uint32(len(ln)), C.c_uint64(aa_db_residues)))
        self.nt = len(ln)
