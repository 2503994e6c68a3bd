// oracle/refseek_main.cpp — TEST INFRASTRUCTURE, not product code.
// A minimal driver around the UNMODIFIED reference implementations of the two hot-path
// modules, compiled from /root/reference by oracle/Makefile:
//   prefilter       lib/mmseqs/src/prefiltering/Main.cpp:13
//   structurealign  src/strucclustutils/structurealign.cpp:141
// plus the reference's createdb (src/strucclustutils/structcreatedb.cpp) to turn the reference's example/ structures into DBs and its
// convertalis (src/strucclustutils/structureconvertalis.cpp) to turn an alignment DB into the .m8 file `easy-search` ends with.
// It replaces only the reference's command table / main (src/foldseek.cpp,
// lib/mmseqs/src/commons/Application.cpp) so that nothing else of the tool has to be linked.
#include <cstring>
#include <cstdio>
#include <vector>
#include "Command.h"
#include "LocalParameters.h"
#include "Debug.h"
#include "Timer.h"
#include "Prefiltering.h"

const char* binary_name = "refseek";
const char* tool_name = "refseek";
const char* tool_introduction = "reference hot-path driver";
const char* main_author = "";
const char* show_extended_help = "1";
const char* show_bash_info = NULL;
const char* index_version_compatible = "fs1";
const char* version = "oracle-ref";
bool hide_base_commands = true;
bool hide_base_downloads = true;
void (*initCommands)(void) = NULL;
// same table as src/FoldseekBase.cpp:585 (k=7 threshold line used by getKmerThreshold)
std::vector<KmerThreshold> externalThreshold = { {Parameters::DBTYPE_AMINO_ACIDS, 7, 197.0, 11.22}};
void (*validatorUpdate)(void) = NULL;

void initParameterSingleton();
extern int prefilter(int argc, const char **argv, const Command &command);
extern int structurealign(int argc, const char **argv, const Command &command);
extern int structcreatedb(int argc, const char **argv, const Command &command);
extern int makepaddedseqdb(int argc, const char **argv, const Command &command);
extern int structureconvertalis(int argc, const char **argv, const Command &command);   // `foldseek convertalis` (src/FoldseekBase.cpp:322)


// ---- dump commands: call the reference classes directly and print their intermediate values ----
#include "DBReader.h"
#include "Sequence.h"
#include "SubstitutionMatrix.h"
#include "StructureSmithWaterman.h"
#include "EvalueNeuralNet.h"
#include "Util.h"
#include "blosum62.out.h"
#include "mat3di.out.h"

static SubstitutionMatrix *makeMat(const char *name, const unsigned char *data, unsigned int len, float bitFactor, float bias) {
    std::string n(name), d((const char *) data, len);
    char *ser = BaseMatrix::serialize(n, d);
    SubstitutionMatrix *m = new SubstitutionMatrix(ser, bitFactor, bias);
    free(ser);
    return m;
}

static void printMat(const char *tag, SubstitutionMatrix *m) {
    printf("MAT %s %d\n", tag, m->alphabetSize);
    for (int i = 0; i < m->alphabetSize; i++) {
        for (int j = 0; j < m->alphabetSize; j++) printf("%d%c", (int) m->subMatrix[i][j], j + 1 == m->alphabetSize ? '\n' : ' ');
    }
    printf("PBACK %s", tag);
    for (int i = 0; i < m->alphabetSize; i++) printf(" %.17g", m->pBack[i]);
    printf("\n");
}

// matdump: the four matrices of the hot path (Prefiltering.cpp:68-69, structurealign.cpp:252-265)
static int matdump() {
    printMat("kmer3di_8.0", makeMat("3di.out", mat3di_out, mat3di_out_len, 8.0, -0.2f));
    printMat("ungapped3di_2.0", makeMat("3di.out", mat3di_out, mat3di_out_len, 2.0, -0.2f));
    printMat("sw3di_2.1", makeMat("3di.out", mat3di_out, mat3di_out_len, 2.1, 0.0));
    printMat("swaa_1.4", makeMat("blosum62.out", blosum62_out, blosum62_out_len, 1.4, 0.0));
    return 0;
}

// swdump <queryDB> <targetDB> <prefDB> : for every (query, candidate) pair of prefDB print what
// alignStructure (structurealign.cpp:37-74) computes before the traceback.
static int swdump(int argc, const char **argv) {
    if (argc < 3) return 1;
    std::string q(argv[0]), t(argv[1]), pdb(argv[2]);
    LocalParameters &par = LocalParameters::getLocalInstance();
    par.compBiasCorrection = 1; par.compBiasCorrectionScale = 0.5;
    DBReader<unsigned int> qaa(q.c_str(), (q + ".index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> qss((q + "_ss").c_str(), (q + "_ss.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> taa(t.c_str(), (t + ".index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> tss((t + "_ss").c_str(), (t + "_ss.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> pref(pdb.c_str(), (pdb + ".index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    qaa.open(DBReader<unsigned int>::NOSORT); qss.open(DBReader<unsigned int>::NOSORT);
    taa.open(DBReader<unsigned int>::NOSORT); tss.open(DBReader<unsigned int>::NOSORT);
    pref.open(DBReader<unsigned int>::LINEAR_ACCCESS);
    SubstitutionMatrix *subMat3Di = makeMat("3di.out", mat3di_out, mat3di_out_len, 2.1, 0.0);
    SubstitutionMatrix *subMatAA = makeMat("blosum62.out", blosum62_out, blosum62_out_len, 1.4, 0.0);
    int A = subMat3Di->alphabetSize;
    int8_t *tinyAA = (int8_t *) mem_align(ALIGN_INT, A * 32), *tiny3Di = (int8_t *) mem_align(ALIGN_INT, A * 32);
    for (int i = 0; i < A; i++) for (int j = 0; j < A; j++) { tiny3Di[i * A + j] = subMat3Di->subMatrix[i][j]; tinyAA[i * A + j] = subMatAA->subMatrix[i][j]; }
    EvalueNeuralNet evaluer(taa.getAminoAcidDBSize(), subMat3Di);
    StructureSmithWaterman ssw(par.maxSeqLen, A, true, 0.5, subMatAA, subMat3Di);
    StructureSmithWaterman rssw(par.maxSeqLen, A, true, 0.5, subMatAA, subMat3Di);
    Sequence qSeqAA(par.maxSeqLen, qaa.getDbtype(), subMatAA, 0, false, true), qSeq3Di(par.maxSeqLen, qss.getDbtype(), subMat3Di, 0, false, true);
    Sequence tSeqAA(par.maxSeqLen, Parameters::DBTYPE_AMINO_ACIDS, subMatAA, 0, false, true), tSeq3Di(par.maxSeqLen, Parameters::DBTYPE_AMINO_ACIDS, subMat3Di, 0, false, true);
    printf("DBRES %zu\n", taa.getAminoAcidDBSize());
    for (size_t id = 0; id < pref.getSize(); id++) {
        char *data = pref.getData(id, 0);
        unsigned int queryKey = pref.getDbKey(id);
        unsigned int queryId = qss.getId(queryKey);
        unsigned int qLen = qss.getSeqLen(queryId);
        qSeq3Di.mapSequence(id, queryKey, qss.getData(queryId, 0), qLen);
        qSeqAA.mapSequence(id, queryKey, qaa.getData(queryId, 0), qLen);
        std::pair<double, double> muLambda = evaluer.predictMuLambda(qSeq3Di.numSequence, qSeq3Di.L);
        ssw.ssw_init(&qSeqAA, &qSeq3Di, tinyAA, tiny3Di, subMatAA);
        qSeq3Di.reverse(); qSeqAA.reverse();
        rssw.ssw_init(&qSeqAA, &qSeq3Di, tinyAA, tiny3Di, subMatAA);
        printf("Q %u %u %.17g %.17g\n", queryKey, qLen, muLambda.first, muLambda.second);
        while (*data != '\0') {
            char keyBuf[256];
            Util::parseKey(data, keyBuf);
            data = Util::skipLine(data);
            unsigned int dbKey = (unsigned int) strtoul(keyBuf, NULL, 10);
            unsigned int targetId = tss.getId(dbKey);
            int tLen = tss.getSeqLen(targetId);
            tSeq3Di.mapSequence(targetId, dbKey, tss.getData(targetId, 0), tLen);
            tSeqAA.mapSequence(targetId, dbKey, taa.getData(targetId, 0), tLen);
            StructureSmithWaterman::s_align a = ssw.alignScoreEndPos<StructureSmithWaterman::PROFILE>(tSeqAA.numSequence, tSeq3Di.numSequence, tLen, 10, 1, qLen / 2);
            StructureSmithWaterman::s_align r = rssw.alignScoreEndPos<StructureSmithWaterman::PROFILE>(tSeqAA.numSequence, tSeq3Di.numSequence, tLen, 10, 1, qLen / 2);
            double e1 = evaluer.computeEvalueCorr(a.score1, muLambda.first, muLambda.second);
            double e2 = evaluer.computeEvalueCorr((int32_t) a.score1 - (int32_t) r.score1, muLambda.first, muLambda.second);
            printf("P %u %u %u %d %d %u %.17g %.17g\n", queryKey, dbKey, a.score1, a.qEndPos1, a.dbEndPos1, r.score1, e1, e2);
        }
    }
    return 0;
}

// sdump <queryDB> <targetDB> <alnDB> : for every line of an alignment DB written with -a, print what the reference's
// LDDTCalculator / TMaligner compute from the backtrace (structurealign.cpp:376-412).
#include "Matcher.h"
#include "TMaligner.h"
#include "LDDT.h"
#include "Coordinate16.h"
static int sdump(int argc, const char **argv) {
    if (argc < 3) return 1;
    std::string q(argv[0]), t(argv[1]), adb(argv[2]);
    DBReader<unsigned int> qss((q + "_ss").c_str(), (q + "_ss.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> tss((t + "_ss").c_str(), (t + "_ss.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> qca((q + "_ca").c_str(), (q + "_ca.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> tca((t + "_ca").c_str(), (t + "_ca.index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    DBReader<unsigned int> aln(adb.c_str(), (adb + ".index").c_str(), 1, DBReader<unsigned int>::USE_INDEX | DBReader<unsigned int>::USE_DATA);
    qss.open(DBReader<unsigned int>::NOSORT); tss.open(DBReader<unsigned int>::NOSORT);
    qca.open(DBReader<unsigned int>::NOSORT); tca.open(DBReader<unsigned int>::NOSORT);
    aln.open(DBReader<unsigned int>::LINEAR_ACCCESS);
    unsigned int maxLen = std::max(qss.getMaxSeqLen(), tss.getMaxSeqLen()) + 1;
    TMaligner tmaligner(maxLen, false, true, false);
    LDDTCalculator lddt(qss.getMaxSeqLen() + 1, tss.getMaxSeqLen() + 1);
    Coordinate16 qcoords, tcoords;
    std::vector<Matcher::result_t> res;
    for (size_t id = 0; id < aln.getSize(); id++) {
        unsigned int queryKey = aln.getDbKey(id);
        unsigned int queryId = qss.getId(queryKey);
        unsigned int qLen = qss.getSeqLen(queryId);
        size_t qId = qca.getId(queryKey);
        float *qc = qcoords.read(qca.getData(qId, 0), qLen, qca.getEntryLen(qId));
        tmaligner.initQuery(qc, &qc[qLen], &qc[2 * qLen], NULL, qLen);
        lddt.initQuery(qLen, qc, &qc[qLen], &qc[2 * qLen]);
        res.clear();
        Matcher::readAlignmentResults(res, aln.getData(id, 0), false);
        for (size_t k = 0; k < res.size(); k++) {
            Matcher::result_t &r = res[k];
            if (r.backtrace.empty()) continue;
            size_t tId = tca.getId(r.dbKey);
            float *tc = tcoords.read(tca.getData(tId, 0), r.dbLen, tca.getEntryLen(tId));
            TMaligner::TMscoreResult tm = tmaligner.computeTMscore(tc, &tc[r.dbLen], &tc[2 * r.dbLen], r.dbLen, r.qStartPos, r.dbStartPos, r.backtrace,
                TMaligner::normalization(0, std::min(r.qEndPos - r.qStartPos, r.dbEndPos - r.dbStartPos), r.qLen, r.dbLen));
            LDDTCalculator::LDDTScoreResult l = lddt.computeLDDTScore(r.dbLen, r.qStartPos, r.dbStartPos, r.backtrace, tc, &tc[r.dbLen], &tc[2 * r.dbLen]);
            printf("S %u %u %.17g %.17g %d\n", queryKey, r.dbKey, l.avgLddtScore, tm.tmscore, (int) (r.score * sqrt(l.avgLddtScore * tm.tmscore)));
        }
    }
    return 0;
}

int main(int argc, const char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: refseek <prefilter|structurealign> <args>\n");
        return 1;
    }
    initParameterSingleton();
    LocalParameters &par = LocalParameters::getLocalInstance();
    if (strcmp(argv[1], "matdump") == 0) return matdump();
    if (strcmp(argv[1], "paramdump") == 0) {
        // the parameter strings the search workflow hands to the two modules (StructureSearch.cpp:101-107: createParameterString of
        // the module's full parameter list; every parameter is serialised, not only the ones the user set)
        LocalParameters &p = LocalParameters::getLocalInstance();
        p.compBiasCorrectionScale = 0.15; p.sensitivity = 9.5; p.maxResListLen = 1000;
        printf("prefilter %s\n", p.createParameterString(p.prefilter).c_str());
        p.compBiasCorrectionScale = 0.5;
        printf("structurealign %s\n", p.createParameterString(p.structurealign).c_str());
        return 0;
    }
    if (strcmp(argv[1], "swdump") == 0) return swdump(argc - 2, argv + 2);
    if (strcmp(argv[1], "sdump") == 0) return sdump(argc - 2, argv + 2);
    Command cmds[] = {
        {"makepaddedseqdb", makepaddedseqdb, &par.makepaddedseqdb, COMMAND_HIDDEN, "", NULL, "", "<i:sequenceDB> <o:sequenceDB>", 0,
         {{"sequenceDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA | DbType::NEED_HEADER, &DbValidator::sequenceDb},
          {"sequenceDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA | DbType::NEED_HEADER, &DbValidator::sequenceDb}}},
        {"createdb", structcreatedb, &par.structurecreatedb, COMMAND_MAIN, "", NULL, "", "<i:PDB|mmCIF[.gz]> ... <o:sequenceDB>", 0,
         {{"PDB|mmCIF[.gz]|stdin|tar[.gz]|DB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA | DbType::VARIADIC, &DbValidator::flatfileStdinAndGeneric},
          {"sequenceDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::flatfile}}},
        {"prefilter", prefilter, &par.prefilter, COMMAND_PREFILTER, "", NULL, "", "<i:queryDB> <i:targetDB> <o:prefilterDB>", 0,
         {{"queryDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
          {"targetDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
          {"prefilterDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::prefilterDb}}},
        {"structurealign", structurealign, &par.structurealign, COMMAND_ALIGNMENT, "", NULL, "", "<i:queryDB> <i:targetDB> <i:prefilterDB> <o:resultDB>", 0,
         {{"queryDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
          {"targetDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::sequenceDb},
          {"resultDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::resultDb},
          {"alnDB", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::alignmentDb}}},
        {"convertalis", structureconvertalis, &par.convertalignments, COMMAND_FORMAT_CONVERSION, "", NULL, "", "<i:queryDb> <i:targetDb> <i:alignmentDB> <o:alignmentFile>", 0,
         {{"queryDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA | DbType::NEED_HEADER, &DbValidator::sequenceDb},
          {"targetDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA | DbType::NEED_HEADER, &DbValidator::sequenceDb},
          {"alignmentDB", DbType::ACCESS_MODE_INPUT, DbType::NEED_DATA, &DbValidator::alignmentDb},
          {"alignmentFile", DbType::ACCESS_MODE_OUTPUT, DbType::NEED_DATA, &DbValidator::flatfile}}},
    };
    for (size_t i = 0; i < sizeof(cmds) / sizeof(cmds[0]); i++) {
        if (strcmp(argv[1], cmds[i].cmd) == 0) {
            Timer timer;
            int status = cmds[i].commandFunction(argc - 2, argv + 2, cmds[i]);
            Debug(Debug::INFO) << "Time for processing: " << timer.lap() << "\n";
            return status;
        }
    }
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 1;
}
void initParameterSingleton() { new LocalParameters; }
