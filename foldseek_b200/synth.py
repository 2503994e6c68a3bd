"""Synthetic Foldseek databases (3Di + AA + C-alpha) in the MMseqs2 on-disk DB format.

Test/bench infrastructure.  The DB layout follows the reference's writer:
  data file  : each entry = payload + '\\n' + '\\0' for sequence DBs  (lib/mmseqs/src/commons/DBWriter.cpp:418-428)
  .index     : "key\\toffset\\tlength\\n", length includes the trailing "\\n\\0"
  .dbtype    : little-endian int32 (0 = amino acids, 12 = generic)   (DBWriter.cpp:193-213)
  <db>_ca    : raw float32 x[L],y[L],z[L] (+'\\0'); Coordinate16::read returns it as-is when
               entryLength >= 3*L*4 (src/commons/Coordinate16.h:15-18)
Letters are drawn i.i.d. from the background frequencies in the reference matrices
(data/mat3di.out:2, lib/mmseqs/data/blosum62.out:2), X excluded.
"""
import os
import struct
import numpy as np

ALPHABET = "ACDEFGHIKLMNPQRSTVWYX"
BG_3DI = np.array([0.0489372, 0.0306991, 0.101049, 0.0329671, 0.0276149, 0.0416262, 0.0452521, 0.030876,
                   0.0297251, 0.0607036, 0.0150238, 0.0215826, 0.0783843, 0.0512926, 0.0264886, 0.0610702,
                   0.0201311, 0.215998, 0.0310265, 0.0295417])
BG_AA = np.array([0.07422, 0.02469, 0.05363, 0.05431, 0.04742, 0.07415, 0.02621, 0.06792, 0.05815, 0.09891,
                  0.02499, 0.04465, 0.03854, 0.03426, 0.05161, 0.05723, 0.05089, 0.07292, 0.01303, 0.03228])
DBTYPE_AMINO_ACIDS = 0
DBTYPE_GENERIC = 12


def random_codes(rng, n, length, bg):
    """n x length uint8 codes in 0..19 drawn from background bg."""
    p = bg / bg.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    u = rng.random((n, length))
    return np.searchsorted(cdf, u, side="right").astype(np.uint8)


def random_walk_ca(rng, n, length):
    """n x 3 x length float32 C-alpha traces: 3.8 A steps in random directions."""
    v = rng.normal(size=(n, length, 3))
    v /= np.linalg.norm(v, axis=2, keepdims=True)
    v *= 3.8
    xyz = np.cumsum(v, axis=1)
    return np.ascontiguousarray(np.transpose(xyz, (0, 2, 1))).astype(np.float32)


class SynthSet:
    """A set of structures: lists of per-entry uint8 code arrays (aa, tdi) and float32 (3,L) ca."""

    def __init__(self, aa, tdi, ca):
        self.aa, self.tdi, self.ca = aa, tdi, ca

    def __len__(self):
        return len(self.aa)


def gen_db(n, length, seed, min_len=None):
    """n entries. length fixed, or uniform in [min_len, length] if min_len is given."""
    rng = np.random.default_rng(seed)
    aa = random_codes(rng, n, length, BG_AA)
    tdi = random_codes(rng, n, length, BG_3DI)
    ca = random_walk_ca(rng, n, length)
    if min_len is None:
        lens = np.full(n, length)
    else:
        lens = rng.integers(min_len, length + 1, size=n)
    return SynthSet([aa[i, :lens[i]].copy() for i in range(n)],
                    [tdi[i, :lens[i]].copy() for i in range(n)],
                    [ca[i, :, :lens[i]].copy() for i in range(n)])


def gen_queries(db, n, length, seed, planted_frac=0.1, sub_prob=0.3, indel_prob=0.02):
    """n queries; a fraction are mutated copies of db entries (so hit lists are not empty)."""
    rng = np.random.default_rng(seed)
    q = gen_db(n, length, seed + 7919)
    n_plant = int(round(n * planted_frac))
    for qi in range(n_plant):
        ti = int(rng.integers(0, len(db)))
        aa, tdi, ca = db.aa[ti], db.tdi[ti], db.ca[ti]
        out_aa, out_tdi, out_ca = [], [], []
        for pos in range(len(aa)):
            r = rng.random()
            if r < indel_prob / 2:      # deletion
                continue
            a, t = aa[pos], tdi[pos]
            if rng.random() < sub_prob:
                a = random_codes(rng, 1, 1, BG_AA)[0, 0]
            if rng.random() < sub_prob:
                t = random_codes(rng, 1, 1, BG_3DI)[0, 0]
            out_aa.append(a); out_tdi.append(t); out_ca.append(ca[:, pos] + rng.normal(scale=0.5, size=3))
            if r > 1 - indel_prob / 2:  # insertion
                out_aa.append(random_codes(rng, 1, 1, BG_AA)[0, 0])
                out_tdi.append(random_codes(rng, 1, 1, BG_3DI)[0, 0])
                out_ca.append(ca[:, pos] + rng.normal(scale=1.5, size=3))
        q.aa[qi] = np.array(out_aa, dtype=np.uint8)
        q.tdi[qi] = np.array(out_tdi, dtype=np.uint8)
        q.ca[qi] = np.ascontiguousarray(np.array(out_ca, dtype=np.float32).T)
    return q


_LUT = np.frombuffer(ALPHABET.encode(), dtype=np.uint8)


def _write_db(path, payloads, dbtype, keys=None):
    keys = range(len(payloads)) if keys is None else keys
    off = 0
    with open(path, "wb") as f, open(path + ".index", "w") as idx:
        for k, p in zip(keys, payloads):
            f.write(p)
            idx.write("%d\t%d\t%d\n" % (k, off, len(p)))
            off += len(p)
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", dbtype))


def write_foldseek_db(prefix, s, name="syn"):
    """Writes <prefix>, <prefix>_ss, <prefix>_ca, <prefix>_h (+.index/.dbtype) and <prefix>.lookup."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    _write_db(prefix, [_LUT[a].tobytes() + b"\n\0" for a in s.aa], DBTYPE_AMINO_ACIDS)
    _write_db(prefix + "_ss", [_LUT[a].tobytes() + b"\n\0" for a in s.tdi], DBTYPE_AMINO_ACIDS)
    _write_db(prefix + "_ca", [np.ascontiguousarray(c, dtype=np.float32).tobytes() + b"\0" for c in s.ca], DBTYPE_GENERIC)
    _write_db(prefix + "_h", [("%s_%d\n\0" % (name, i)).encode() for i in range(len(s))], DBTYPE_GENERIC)
    with open(prefix + ".lookup", "w") as f:
        for i in range(len(s)):
            f.write("%d\t%s_%d\t%d\n" % (i, name, i, i))


def read_result_db(path):
    """Reads an MMseqs2 result DB (prefilter or alignment) -> {key: [line fields...]}."""
    if os.path.exists(path):
        data = open(path, "rb").read()
    else:  # per-thread shards <path>.0, <path>.1, ... addressed as one concatenated file (DBReader.cpp)
        data, i = b"", 0
        while os.path.exists("%s.%d" % (path, i)):
            data += open("%s.%d" % (path, i), "rb").read()
            i += 1
    out = {}
    for line in open(path + ".index"):
        k, off, ln = line.split()
        body = data[int(off):int(off) + int(ln)].rstrip(b"\0").decode()
        out[int(k)] = [l.split("\t") for l in body.split("\n") if l]
    return out


def write_seq_db_fixed(path, codes, dbtype=DBTYPE_AMINO_ACIDS):
    """Fast writer for n equal-length sequences given as an (n, L) uint8 code matrix."""
    n, L = codes.shape
    buf = np.empty((n, L + 2), np.uint8)
    buf[:, :L] = _LUT[codes]
    buf[:, L] = 10
    buf[:, L + 1] = 0
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    buf.tofile(path)
    with open(path + ".index", "w") as f:
        f.write("".join("%d\t%d\t%d\n" % (i, i * (L + 2), L + 2) for i in range(n)))
    with open(path + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", dbtype))


def rekey_like_padded(dirname, src="db", pad="pad"):
    """What `foldseek makepaddeddb` does to the databases next to the padded 3Di DB (data/makepaddeddb.sh: renamedbkeys --subdb-mode 1 with
    the key mapping of <pad>_ss.lookup): <pad>, <pad>_ca and <pad>_h are the source DBs under the padded DB's keys (ids in descending
    length order), data files unchanged.  Lets tests run structurealign / search against the GPU-padded target layout."""
    import os
    import shutil
    m = {}
    for line in open(os.path.join(dirname, pad + "_ss.lookup")):
        new, _, old = line.rstrip("\n").split("\t")
        m[int(old)] = int(new)
    for suf in ("", "_ca", "_h"):
        rows = []
        for line in open(os.path.join(dirname, src + suf + ".index")):
            k, o, ln = line.split("\t")
            rows.append((m[int(k)], o, ln.strip()))
        rows.sort()
        with open(os.path.join(dirname, pad + suf + ".index"), "w") as f:
            for k, o, ln in rows:
                f.write("%d\t%s\t%s\n" % (k, o, ln))
        shutil.copy(os.path.join(dirname, src + suf), os.path.join(dirname, pad + suf))
        shutil.copy(os.path.join(dirname, src + suf + ".dbtype"), os.path.join(dirname, pad + suf + ".dbtype"))


def mask_lower_case(path, seed=5, stretches=2):
    """lower-cases `stretches` random runs of 8-24 letters in every entry of the sequence DB at `path`, in place: what `createdb` does to
    the 3Di (and AA) letters of low-confidence residues (--mask-bfactor-threshold), which --mask-lower-case 1 then hides from the k-mer index
    and the ungapped scoring while the aligner still sees them"""
    import random
    rnd = random.Random(seed)
    data = bytearray(open(path, "rb").read())
    for line in open(path + ".index"):
        _, o, ln = (int(x) for x in line.split())
        L = ln - 2
        for _ in range(stretches):
            a = rnd.randrange(0, max(1, L - 25))
            b = min(L, a + rnd.randrange(8, 25))
            for i in range(o + a, o + b):
                data[i] = ord(chr(data[i]).lower())
    open(path, "wb").write(bytes(data))


def gen_edge_case_dbs(dirname, n_targets=600, n_queries=16, max_len=900, seed=11):
    """target DB `t` with ragged lengths 3..max_len (shorter than the k-mer span of 10, longer than the 384 residues of the packed SW
    kernel) and 2 % X residues in both alphabets, query DB `q` of planted homologs (substitutions and indels): the inputs i.i.d. fixed-length
    data never produces"""
    import os
    db = gen_db(n_targets, max_len, seed, min_len=3)
    rng = np.random.default_rng(seed + 1)
    for i in range(len(db.aa)):
        L = len(db.aa[i])
        db.aa[i][rng.random(L) < 0.02] = 20
        db.tdi[i][rng.random(L) < 0.02] = 20
    q = gen_queries(db, n_queries, max_len, seed + 2, planted_frac=0.9, sub_prob=0.25, indel_prob=0.03)
    write_foldseek_db(os.path.join(dirname, "t"), db, "t")
    write_foldseek_db(os.path.join(dirname, "q"), q, "q")
    return db, q
