// cli_db.h - errors and console of a module run, MMseqs2 on-disk databases (read, write, zstd frames), residue and C-alpha decoding
// Part of the host module (foldseek-b200), included by cli_main.cpp inside its anonymous namespace.
#ifndef CLI_DB_H
#define CLI_DB_H

// Debug(Debug::ERROR) << ...; EXIT(EXIT_FAILURE): a module run ends with this message and a non-zero exit code.  Thrown instead of
// exiting on the spot because the same module code runs inside the gpuserver, where the error goes back to the calling process.
struct Fatal : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] void die(const std::string &msg) { throw Fatal(msg); }

// stdout / stderr of a module run: the process's own streams, or (inside the gpuserver) the text returned to the client
struct Console {
    std::mutex m; bool capture = false; std::string out, err;
    void put(bool toErr, const std::string &t) {
        std::lock_guard<std::mutex> g(m);
        if (capture) (toErr ? err : out) += t;
        else { fputs(t.c_str(), toErr ? stderr : stdout); fflush(toErr ? stderr : stdout); }
    }
} g_console;
void say_err(const std::string &t) { g_console.put(true, t); }
void say_out(const std::string &t) { g_console.put(false, t); }

// one host thread per context; the first error of any of them ends the run like an error on the main thread
template <class F> void run_workers(size_t n, F fn) {
    std::vector<std::thread> th; std::mutex m; std::string first; bool failed = false;
    for (size_t g = 0; g < n; g++) th.emplace_back([&, g]() {
        try { fn(g); } catch (const std::exception &e) { std::lock_guard<std::mutex> l(m); if (!failed) { failed = true; first = e.what(); } }
    });
    for (std::thread &t : th) t.join();
    if (failed) throw Fatal(first);
}

std::string slurp(const std::string &path, bool required = true) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { if (required) die("cannot read " + path); return std::string(); }
    std::ostringstream ss; ss << in.rdbuf();
    return ss.str();
}

// zstd frames of a compressed DB are inflated with the system's libzstd, looked up at run time (the image has the library but no
// headers; the two entry points used are part of zstd's stable ABI)
size_t zstd_inflate(void *dst, size_t cap, const void *src, size_t n) {
    typedef size_t (*dec_t)(void *, size_t, const void *, size_t);
    typedef unsigned (*err_t)(size_t);
    static dec_t dec = NULL; static err_t isErr = NULL;
    if (dec == NULL) {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (h == NULL) h = dlopen("libzstd.so", RTLD_NOW);
        if (h == NULL) die("compressed database: libzstd.so.1 not found on this system");
        dec = (dec_t) dlsym(h, "ZSTD_decompress"); isErr = (err_t) dlsym(h, "ZSTD_isError");
        if (dec == NULL || isErr == NULL) die("compressed database: libzstd lacks ZSTD_decompress");
    }
    const size_t r = dec(dst, cap, src, n);
    if (isErr(r)) die("compressed database: zstd frame cannot be decompressed");
    return r;
}

struct Db {   // read-only MMseqs2 DB (per-thread shards <name>.0, .1, ... are concatenated; compressed DBs are inflated on open)
    std::string data;
    uint64_t fileBytes = 0;        // size of the data file(s) as stored: DBReader::getAminoAcidDBSize() counts these bytes
    std::vector<uint32_t> key; std::vector<uint64_t> off, len;
    int dbtype = 0;
    bool padded = false;           // DBTYPE_EXTENDED_GPU (lib/mmseqs/src/util/makepaddedseqdb.cpp:27): numeric codes, no "\n\0"
    // appends the bytes of a file to `to` (one read into the final buffer: a 1M-structure DB is 350 MB per file)
    static bool append_file(const std::string &path, std::string &to) {
        FILE *f = fopen(path.c_str(), "rb");
        if (f == NULL) return false;
        struct stat st;
        const size_t have = to.size();
        if (fstat(fileno(f), &st) == 0 && st.st_size > 0) {
            to.resize(have + (size_t) st.st_size);
            const size_t got = fread(&to[have], 1, (size_t) st.st_size, f);
            to.resize(have + got);
        }
        fclose(f);
        return true;
    }
    void open(const std::string &name) {
        if (!append_file(name, data)) {
            for (int i = 0;; i++)
                if (!append_file(name + "." + std::to_string(i), data)) { if (i == 0) die("cannot open database " + name); break; }
        }
        std::string idx;
        if (!append_file(name + ".index", idx)) die("cannot open " + name + ".index");
        struct Row { uint32_t key; uint64_t off, len; };
        std::vector<Row> rows;
        rows.reserve(idx.size() / 12 + 1);
        bool sorted = true;
        for (const char *p = idx.c_str(), *e = p + idx.size(); p < e;) {   // key \t offset \t length \n
            char *q;
            while (p < e && isspace((unsigned char) *p)) p++;
            if (p >= e) break;
            const unsigned long long k = strtoull(p, &q, 10); if (q == p) break; p = q;
            const unsigned long long o = strtoull(p, &q, 10); if (q == p) break; p = q;
            const unsigned long long l = strtoull(p, &q, 10); if (q == p) break; p = q;
            if (!rows.empty() && rows.back().key > (uint32_t) k) sorted = false;
            rows.push_back(Row{(uint32_t) k, o, l});
        }
        // DBReader ids follow the key order of the index (DBReader::open sorts an unsorted index by key)
        if (!sorted) std::stable_sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) { return x.key < y.key; });
        key.resize(rows.size()); off.resize(rows.size()); len.resize(rows.size());
        for (size_t i = 0; i < rows.size(); i++) { key[i] = rows[i].key; off[i] = rows[i].off; len[i] = rows[i].len; }
        std::string t = slurp(name + ".dbtype");
        if (t.size() >= 4) memcpy(&dbtype, t.data(), 4);
        fileBytes = data.size();
        if (dbtype & (1 << 31)) {
            // DBWriter::writeEnd (lib/mmseqs/src/commons/DBWriter.cpp:331-398): u32 stored size, the bytes, then '\0' behind a zstd frame or
            // 0xFF behind an entry stored as is; the index keeps the ORIGINAL length (incl. the terminator).  DBReader::getDataCompressed
            // (DBReader.cpp:374-401).
            std::string plain;
            std::vector<uint64_t> noff(key.size());
            for (size_t i = 0; i < key.size(); i++) {
                if (off[i] + 4 > data.size()) die(name + ": truncated compressed entry");
                uint32_t cSize; memcpy(&cSize, data.data() + off[i], 4);
                const char *body = data.data() + off[i] + 4;
                if (off[i] + 4 + (uint64_t) cSize + 1 > data.size()) die(name + ": truncated compressed entry");
                noff[i] = plain.size();
                const uint64_t orig = len[i] > 0 ? len[i] - 1 : 0;
                if (body[cSize] == 0) {
                    std::string buf(orig + 64, '\0');
                    const size_t got = zstd_inflate(&buf[0], buf.size(), body, cSize);
                    if (got != orig) die(name + ": compressed entry does not have the length of its index line");
                    plain.append(buf.data(), got);
                } else {
                    if (cSize != orig) die(name + ": stored entry does not have the length of its index line");
                    plain.append(body, cSize);
                }
                plain.push_back('\0');
            }
            data.swap(plain); off.swap(noff);
            dbtype &= ~(1 << 31);
        }
        // extended flags live in bits 16..30 (DBReader.h:408-415); DBTYPE_EXTENDED_GPU = 8 (Parameters.h:94)
        padded = (((uint32_t) dbtype >> 16) & 0x7FFEu & 8u) != 0;
    }
    size_t size() const { return key.size(); }
};

// one zstd frame per entry for --compressed 1 (DBWriter.cpp:289-398; any valid frame is read back by DBReader::getDataCompressed)
std::string zstd_deflate(const std::string &in) {
    typedef size_t (*cmp_t)(void *, size_t, const void *, size_t, int);
    typedef size_t (*bnd_t)(size_t);
    typedef unsigned (*err_t)(size_t);
    static cmp_t cmp = NULL; static bnd_t bound = NULL; static err_t isErr = NULL;
    if (cmp == NULL) {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (h == NULL) h = dlopen("libzstd.so", RTLD_NOW);
        if (h == NULL) die("--compressed 1: libzstd.so.1 not found on this system");
        cmp = (cmp_t) dlsym(h, "ZSTD_compress"); bound = (bnd_t) dlsym(h, "ZSTD_compressBound"); isErr = (err_t) dlsym(h, "ZSTD_isError");
        if (cmp == NULL || bound == NULL || isErr == NULL) die("--compressed 1: libzstd lacks ZSTD_compress");
    }
    std::string out(bound(in.size()), '\0');
    const size_t n = cmp(&out[0], out.size(), in.data(), in.size(), 3);
    if (isErr(n)) die("--compressed 1: zstd compression failed");
    out.resize(n);
    return out;
}

void write_db(const std::string &name, const std::vector<uint32_t> &keys, const std::vector<std::string> &entries, int dbtype, bool compressed = false) {
    std::ofstream d(name, std::ios::binary), i(name + ".index");
    if (!d || !i) die("cannot write " + name);
    uint64_t off = 0;
    for (size_t k = 0; k < keys.size(); k++) {
        if (compressed) {   // u32 frame size, frame, '\0'; the index keeps the original length (DBWriter.cpp:359-396)
            const std::string frame = zstd_deflate(entries[k]);
            const uint32_t n = (uint32_t) frame.size();
            d.write(reinterpret_cast<const char *>(&n), 4); d.write(frame.data(), frame.size()); d.put('\0');
            i << keys[k] << '\t' << off << '\t' << (entries[k].size() + 1) << '\n';
            off += 4 + frame.size() + 1;
            continue;
        }
        d.write(entries[k].data(), entries[k].size()); d.put('\0');
        i << keys[k] << '\t' << off << '\t' << (entries[k].size() + 1) << '\n';
        off += entries[k].size() + 1;
    }
    if (compressed) dbtype |= (int) (1u << 31);
    d.close(); i.close();
    std::ofstream t(name + ".dbtype", std::ios::binary);   // last: resume guard
    t.write(reinterpret_cast<const char *>(&dbtype), 4);
}

// letters -> codes (BaseMatrix alphabet A C D E F G H I K L M N P Q R S T V W Y X; unknown -> X; lower == upper).
// Padded DBs (makepaddedseqdb.cpp:66-98) already hold the codes, +32 for lower-case/masked residues (DBReader.cpp:349-372).
void encode(const Db &db, bool maskLower, std::vector<uint8_t> &codes, std::vector<uint64_t> &off, std::vector<uint32_t> &len) {
    static const char *alph = "ACDEFGHIKLMNPQRSTVWYX";
    uint8_t map[256];
    for (int c = 0; c < 256; c++) map[c] = 20;
    for (int a = 0; a < 21; a++) { map[(unsigned char) alph[a]] = (uint8_t) a; map[(unsigned char) tolower(alph[a])] = (uint8_t) a; }
    off.resize(db.size()); len.resize(db.size());
    uint64_t total = 0;
    for (size_t i = 0; i < db.size(); i++) { len[i] = (uint32_t) (db.len[i] >= 2 ? db.len[i] - 2 : 0); off[i] = total; total += len[i]; }   // length includes "\n\0"
    codes.resize(total + 1);
    // one table for the whole byte -> code step; --mask-lower-case 1 (Foldseek default, LocalParameters.cpp:382-385): lower-case
    // (padded DBs: +32) target residues are masked
    uint8_t lut[256];
    for (int c = 0; c < 256; c++) {
        if (db.padded) { const int base = (c >= 32) ? c - 32 : c; lut[c] = (maskLower && c >= 32) ? 20 : (uint8_t) (base > 20 ? 20 : base); }
        else lut[c] = (maskLower && islower(c)) ? 20 : map[c];
    }
    const int64_t n = (int64_t) db.size();
#pragma omp parallel for schedule(static, 4096)
    for (int64_t i = 0; i < n; i++) {
        if (db.off[i] + len[i] > db.data.size()) continue;   // reported below
        const unsigned char *s = reinterpret_cast<const unsigned char *>(db.data.data()) + db.off[i];
        uint8_t *d = codes.data() + off[i];
        for (uint32_t k = 0; k < len[i]; k++) d[k] = lut[s[k]];
    }
    for (size_t i = 0; i < db.size(); i++) if (db.off[i] + len[i] > db.data.size()) die("database entry " + std::to_string(db.key[i]) + " lies outside the data file");
}


// Coordinate16::read (src/commons/Coordinate16.h:15-57): raw float32 x[L] y[L] z[L], or int32 start + int16 differences per axis
void decode_ca_into(const Db &db, size_t entry, uint32_t L, float *x, float *y, float *z) {
    const char *mem = db.data.data() + db.off[entry];
    const size_t entryLen = db.len[entry];
    float *out[3] = {x, y, z};
    if (entryLen >= (size_t) L * 3 * sizeof(float)) {
        for (int d = 0; d < 3; d++) memcpy(out[d], mem + (size_t) d * L * sizeof(float), (size_t) L * sizeof(float));
        return;
    }
    const char *p = mem;
    for (int d = 0; d < 3; d++) {
        int32_t start, diffSum = 0;
        memcpy(&start, p, 4); p += 4;
        if (L > 0) out[d][0] = start / 1000.0f;
        for (uint32_t i = 1; i < L; i++) {
            int16_t diff;
            memcpy(&diff, p, 2); p += 2;
            diffSum += diff;
            out[d][i] = (start + diffSum) / 1000.0f;
        }
    }
}
void decode_ca(const Db &db, size_t entry, uint32_t L, std::vector<float> &x, std::vector<float> &y, std::vector<float> &z) {
    const size_t base = x.size();
    x.resize(base + L); y.resize(base + L); z.resize(base + L);
    decode_ca_into(db, entry, L, x.data() + base, y.data() + base, z.data() + base);
}


// DBReader::decomposeDomainByAminoAcid (lib/mmseqs/src/commons/DBReader.cpp:1108-1150): entry ranges of `parts` chunks of about equal bytes
std::vector<std::pair<size_t, size_t>> decompose_by_residues(const Db &db, size_t parts) {
    std::vector<std::pair<size_t, size_t>> r(parts, {0, 0});
    const size_t entries = db.size();
    if (parts <= 1) { r.assign(1, {0, entries}); return r; }
    if (entries <= parts) { for (size_t p = 0; p < parts; p++) r[p] = {p < entries ? p : 0, p < entries ? 1 : 0}; return r; }
    const size_t chunk = (size_t) ceil((double) db.fileBytes / (double) parts);
    std::vector<size_t> per(parts, 0);
    size_t rank = 0, sum = 0;
    for (size_t i = 0; i < entries; i++) {
        if (sum >= chunk) { sum = 0; rank++; }
        sum += db.len[i];
        per[rank]++;
    }
    size_t start = 0;
    for (size_t p = 0; p < parts; p++) { r[p] = {start, per[p]}; start += per[p]; }
    return r;
}

#endif
