// cli_convertalis.h - `foldseek-b200 convertalis <queryDB> <targetDB> <alnDB> <out.m8>`: the alignment DB as the BLAST-tab file that
// `foldseek easy-search` ends with (data/easystructuresearch.sh:76-80 -> src/strucclustutils/structureconvertalis.cpp:253), default
// --format-mode 0 with the default --format-output (12 columns: query, target, fident, alnlen, mismatch, gapopen, qstart, qend, tstart,
// tend, evalue, bits).  Text in, text out: no GPU work, no context.
// Part of the host module (foldseek-b200), included by cli_main.cpp inside its anonymous namespace.
#ifndef CLI_CONVERTALIS_H
#define CLI_CONVERTALIS_H

// Util::parseFastaHeader / Util::getFastaHeaderPosition (lib/mmseqs/src/commons/Util.cpp:147-229): the identifier of a header line -
// the first word, or the accession between the bars of the database-prefixed forms (sp|ACC|..., gi|n|db|ACC|...)
std::string fasta_identifier(const char *headerPtr) {
    size_t len = 0;
    while (headerPtr[len] != '\0' && !isspace((unsigned char) headerPtr[len])) len++;   // Util::skipNoneWhitespace
    const std::string header(headerPtr, len);
    if (header.empty()) return "";
    size_t offset = 0;
    if (header.compare(0, 10, "consensus_") == 0) offset = 10;
    static const struct { const char *prefix; unsigned length, bar; } dbs[] = {
        {"cl|", 3, 1}, {"sp|", 3, 1}, {"tr|", 3, 1}, {"gb|", 3, 1}, {"ref|", 4, 1}, {"pdb|", 4, 1}, {"bbs|", 4, 1}, {"lcl|", 4, 1},
        {"pir||", 5, 1}, {"prf||", 5, 1}, {"gnl|", 4, 2}, {"pat|", 4, 2}, {"gi|", 3, 3}};
    for (const auto &d : dbs) {
        if (header.compare(offset, strlen(d.prefix), d.prefix) != 0) continue;
        size_t start = offset + d.length;
        for (unsigned j = 0; j + 1 < d.bar; j++) {
            const size_t end = header.find_first_of('|', start);
            if (end == std::string::npos) return "";
            start = end + 1;
        }
        size_t end = header.find_first_of('|', start);
        if (end == std::string::npos) end = header.find_first_of(" \n", start);
        if (end == std::string::npos) end = header.length();
        return header.substr(start, end - start);
    }
    size_t end = header.find_first_of(" \n", offset);
    if (end == std::string::npos) end = header.length();
    return header.substr(offset, end - offset);
}

int convertalis_main(int argc, char **argv) {
    Args a = parse(argc, argv, 2);
    if (a.pos.size() != 4) die("usage: foldseek-b200 convertalis <i:queryDb> <i:targetDb> <i:alignmentDB> <o:alignmentFile> [options]");
    Db qh, th, aln;
    qh.open(a.pos[0] + "_h"); th.open(a.pos[1] + "_h"); aln.open(a.pos[2]);
    auto by_key = [](const Db &db) {
        std::vector<uint32_t> m(db.size() ? *std::max_element(db.key.begin(), db.key.end()) + 1 : 0, 0xFFFFFFFFu);
        for (size_t i = 0; i < db.size(); i++) m[db.key[i]] = (uint32_t) i;
        return m;
    };
    const std::vector<uint32_t> qOf = by_key(qh), tOf = by_key(th);
    std::ofstream out(a.pos[3], std::ios::binary);
    if (!out) die("cannot write " + a.pos[3]);
    char buffer[1024];
    for (size_t i = 0; i < aln.size(); i++) {
        const uint32_t qKey = aln.key[i];
        if (qKey >= qOf.size() || qOf[qKey] == 0xFFFFFFFFu) die("alignment entry for a query key without header: " + std::to_string(qKey));
        const std::string queryId = fasta_identifier(qh.data.data() + qh.off[qOf[qKey]]);
        const char *data = aln.data.data() + aln.off[i];
        while (*data != '\0') {
            // Matcher::parseAlignmentRecord (lib/mmseqs/src/alignment/Matcher.cpp:205-260): 10 columns, or 11 with the backtrace
            std::vector<std::string> col;
            const char *p = data;
            while (*p != '\n' && *p != '\0') {
                const char *e = p;
                while (*e != '\t' && *e != ' ' && *e != '\n' && *e != '\0') e++;
                if (e > p) col.push_back(std::string(p, e));
                p = e;
                while (*p == '\t' || *p == ' ') p++;
            }
            data = (*p == '\n') ? p + 1 : p;
            if (col.size() < 10) die("Invalid alignment result record.");
            const uint32_t dbKey = (uint32_t) strtoul(col[0].c_str(), NULL, 10);
            const int score = atoi(col[1].c_str());
            const float seqId = (float) strtod(col[2].c_str(), NULL);
            const double eval = strtod(col[3].c_str(), NULL);
            const int qStart = atoi(col[4].c_str()), qEnd = atoi(col[5].c_str()), dbStart = atoi(col[7].c_str()), dbEnd = atoi(col[8].c_str());
            const std::string bt = col.size() > 10 ? col[10] : std::string();
            if (dbKey >= tOf.size() || tOf[dbKey] == 0xFFFFFFFFu) die("alignment row for a target key without header: " + std::to_string(dbKey));
            const std::string targetId = fasta_identifier(th.data.data() + th.off[tOf[dbKey]]);
            const int adjQ = qStart == -1 ? 0 : qStart, adjT = dbStart == -1 ? 0 : dbStart;
            unsigned int gapOpen = 0, mismatch = 0;
            unsigned int alnLen = (unsigned int) (std::max(abs(qEnd - adjQ), abs(dbEnd - adjT)) + 1);   // Matcher::computeAlnLength
            if (!bt.empty()) {   // structureconvertalis.cpp:731-757: counts from the compressed backtrace ("12M1D30M")
                size_t matchCount = 0;
                alnLen = 0;
                for (size_t pos = 0; pos < bt.size(); pos++) {
                    int cnt = 0;
                    if (isdigit((unsigned char) bt[pos])) {
                        cnt = atoi(bt.c_str() + pos);
                        while (pos < bt.size() && isdigit((unsigned char) bt[pos])) pos++;
                    }
                    alnLen += (unsigned int) cnt;
                    if (pos < bt.size()) {
                        if (bt[pos] == 'M') matchCount += (size_t) cnt;
                        else if (bt[pos] == 'D' || bt[pos] == 'I') gapOpen += 1;
                    }
                }
                const unsigned int identical = (unsigned int) (seqId * (float) alnLen + 0.5);
                mismatch = (unsigned int) (matchCount - identical);
            } else {             // :758-763: an estimate from the end positions
                const float best = (float) std::min(abs(qEnd - adjQ), abs(dbEnd - adjT));
                mismatch = (unsigned int) (best * (1.0f - seqId) + 0.5);
            }
            // the default --format-output goes through the per-column writer (:846-1000): SSTR(float) = "{:.3f}", SSTR(double) = "{:.3E}"
            // (lib/mmseqs/src/commons/Util.cpp:722-735), integers as they are
            const int n = snprintf(buffer, sizeof(buffer), "%s\t%s\t%.3f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.3E\t%d\n", queryId.c_str(), targetId.c_str(), seqId,
                                   (int) alnLen, (int) mismatch, (int) gapOpen, qStart + 1, qEnd + 1, dbStart + 1, dbEnd + 1, eval, score);
            if (n < 0 || (size_t) n >= sizeof(buffer)) { say_err("Truncated line in entry" + std::to_string(i) + "!\n"); continue; }
            out.write(buffer, n);
        }
    }
    return EXIT_SUCCESS;
}

#endif
