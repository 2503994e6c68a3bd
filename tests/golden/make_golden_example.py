#!/usr/bin/env python3
"""Real-protein fixture (BASELINE.json configs[0]): the reference's own example/ structures through the REFERENCE's createdb,
prefilter and structurealign (oracle/_ref/refseek = reference sources compiled by oracle/Makefile), with the parameter strings
the search workflow serialises for the two modules (refseek paramdump = createParameterString, StructureSearch.cpp:101-107).

    python tests/golden/make_golden_example.py        (build container, needs /root/reference)

Writes tests/golden/example/: db* (all 30 chains of example/), q* (d1asha_ alone), and the reference outputs
    ref_pref_all / ref_aln_all      db vs db   (query DB == target DB: identity hits, QueryMatcher.cpp:409-421)
    ref_pref_q   / ref_aln_q        d1asha_ vs db
    ref_aln_q_cov                   d1asha_ vs db with -c 0.8 --cov-mode 0 --min-seq-id 0.2 -a 1
    pad_ss* / ref_pref_q_pad        the reference's padded (GPU) layout of db_ss and d1asha_ vs it
    ref_pref_all_cov{1,2} / ref_aln_all_cov{1,2}   db vs db with -c 0.7 --cov-mode 1 | 2
    ref_aln_all_alt2                db vs db with --alt-ali 2 -a 1 (alternative alignments, structurealign.cpp:115-138)
    ref_aln_q.m8 / ref_aln_all_alt2.m8   the reference's convertalis on ref_aln_q (no backtraces: what `foldseek easy-search example/d1asha_
                                    example/ aln.m8 tmp` ends with) and on ref_aln_all_alt2 (with backtraces)
    ref_aln_q_pad                   d1asha_ vs the padded target layout (pad_ss padded + pad / pad_ca / pad_h re-keyed like makepaddeddb does, derived at
                                    run time by foldseek_b200.synth.rekey_like_padded), from ref_pref_q_pad
    ref_pref_all_nobias / ref_aln_all_nobias   db vs db with --comp-bias-corr 0 (prefilter also -s 4 --max-seqs 20; aligner -e 0.001 on ref_pref_all):
                                    the cluster workflow's early steps switch the composition-bias correction off
    ref_aln_all_3di                 db vs db with --alignment-type 0 -a 1 (3Di only: AA matrix with bit factor 0, structurealign.cpp:264-265)
    ref_pref_all_s75 / ref_aln_all_gap82   db vs db with -s 7.5 --min-ungapped-score 50 --max-seqs 20 (prefilter) and --gap-open 8 --gap-extend 2 -e 0.001 -a 1
                                    (aligner, on ref_pref_all): parameters the library reads from its configuration, away from their defaults
    ref_pref_all_masked / ref_aln_all_masked   db vs db after foldseek_b200.synth.mask_lower_case(db_ss): lower-case (masked) 3Di stretches in the target,
                                    default --mask-lower-case 1 (the prefilter sees them as X, the aligner does not); the masked DB is derived at run time
    ref_pref_all_split3             db vs db with --split 3 --split-mode 0 --max-seqs 10 (target split: three indexes, lists of 10 merged)
    ref_aln_all_tmmode1             db vs db with --tmscore-threshold-mode 1 --tmscore-threshold 0.3 (TM-score normalised by the query length)
Real proteins bring what i.i.d. synthetic data does not: X residues, ragged lengths, saturated diagonal scores, homologous
families, Coordinate16-encoded C-alpha traces."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFSEEK = os.path.join(ROOT, "oracle", "_ref", "refseek")
OUT = os.path.join(ROOT, "tests", "golden", "example")
EXAMPLE = "/root/reference/example"


def run(*a):
    r = subprocess.run([REFSEEK] + list(a), cwd=OUT, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit("refseek %s failed:\n%s\n%s" % (a[0], r.stdout[-2000:], r.stderr[-2000:]))
    return r.stdout


def params():
    out = {}
    for line in subprocess.run([REFSEEK, "paramdump"], capture_output=True, text=True, check=True).stdout.splitlines():
        mod, _, rest = line.partition(" ")
        out[mod] = rest.split()
        out[mod][out[mod].index("--threads") + 1] = "1"      # single output shard
    return out


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    run("createdb", EXAMPLE + "/", "db", "--threads", "1")
    run("createdb", EXAMPLE + "/d1asha_", "q", "--threads", "1")
    p = params()
    with open(os.path.join(OUT, "params.txt"), "w") as f:
        f.write("prefilter %s\nstructurealign %s\n" % (" ".join(p["prefilter"]), " ".join(p["structurealign"])))
    run("prefilter", "db_ss", "db_ss", "ref_pref_all", *p["prefilter"])
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all", *p["structurealign"])
    run("prefilter", "q_ss", "db_ss", "ref_pref_q", *p["prefilter"])
    run("structurealign", "q", "db", "ref_pref_q", "ref_aln_q", *p["structurealign"])
    cov = list(p["structurealign"])
    for flag, v in (("-c", "0.8"), ("--cov-mode", "0"), ("--min-seq-id", "0.2"), ("-a", "1")):
        cov[cov.index(flag) + 1] = v
    run("structurealign", "q", "db", "ref_pref_q", "ref_aln_q_cov", *cov)
    # padded (GPU) target DB, the layout of `foldseek makepaddeddb` (data/makepaddeddb.sh:9-20, lib/mmseqs/src/util/makepaddedseqdb.cpp):
    # numeric codes, entries re-keyed in descending length order and padded to 4 bytes, dbtype flag DBTYPE_EXTENDED_GPU
    for ext in ("", ".index", ".dbtype"):
        shutil.copy(os.path.join(OUT, "db_h" + ext), os.path.join(OUT, "db_ss_h" + ext))
    run("makepaddedseqdb", "db_ss", "pad_ss")
    for ext in ("", ".index", ".dbtype"):
        os.remove(os.path.join(OUT, "db_ss_h" + ext))
    run("prefilter", "q_ss", "pad_ss", "ref_pref_q_pad", *p["prefilter"])
    cov_modes(p)
    alt_ali(p)
    print("wrote", sorted(os.listdir(OUT)))


def cov_modes(p=None):
    """-c 0.7 with --cov-mode 1 (target coverage) and 2 (query coverage), all-vs-all: the prefilter's canBeCovered filter
    (Prefiltering.cpp:886-894 applies it for modes 0, 2 and 5) and the aligner's canBeCovered / hasCoverage / checkCriteria"""
    p = p or params()
    for mode in ("1", "2"):
        pf = list(p["prefilter"]); al = list(p["structurealign"])
        for lst in (pf, al):
            lst[lst.index("-c") + 1] = "0.7"
            lst[lst.index("--cov-mode") + 1] = mode
        run("prefilter", "db_ss", "db_ss", "ref_pref_all_cov" + mode, *pf)
        run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_cov" + mode, *al)


def alt_ali(p=None):
    """--alt-ali 2 -a 1, all-vs-all: two alternative alignments per accepted pair (computeAlternativeAlignment); and the TM-score
    normalised by the query length (--tmscore-threshold-mode 1) with a threshold of 0.3"""
    p = p or params()
    al = list(p["structurealign"])
    for flag, v in (("--alt-ali", "2"), ("-a", "1")):
        al[al.index(flag) + 1] = v
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_alt2", *al)
    al = list(p["structurealign"])
    for flag, v in (("--tmscore-threshold-mode", "1"), ("--tmscore-threshold", "0.3")):
        al[al.index(flag) + 1] = v
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_tmmode1", *al)
    pf = list(p["prefilter"])
    for flag, v in (("--split", "3"), ("--split-mode", "0"), ("--max-seqs", "10")):
        pf[pf.index(flag) + 1] = v
    run("prefilter", "db_ss", "db_ss", "ref_pref_all_split3", *pf)
    pf = list(p["prefilter"]); al = list(p["structurealign"])
    for flag, v in (("--comp-bias-corr", "0"), ("-s", "4"), ("--max-seqs", "20")):
        pf[pf.index(flag) + 1] = v
    for flag, v in (("--comp-bias-corr", "0"), ("-e", "0.001")):
        al[al.index(flag) + 1] = v
    pf2 = list(p["prefilter"]); al2 = list(p["structurealign"])
    for flag, v in (("-s", "7.5"), ("--min-ungapped-score", "50"), ("--max-seqs", "20")):
        pf2[pf2.index(flag) + 1] = v
    for flag, v in (("--gap-open", "aa:8,nucl:8"), ("--gap-extend", "aa:2,nucl:2"), ("-e", "0.001"), ("-a", "1")):
        al2[al2.index(flag) + 1] = v
    run("prefilter", "db_ss", "db_ss", "ref_pref_all_s75", *pf2)
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_gap82", *al2)
    al0 = list(p["structurealign"])
    for flag, v in (("--alignment-type", "0"), ("-a", "1")):
        al0[al0.index(flag) + 1] = v
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_3di", *al0)
    run("prefilter", "db_ss", "db_ss", "ref_pref_all_nobias", *pf)
    run("structurealign", "db", "db", "ref_pref_all", "ref_aln_all_nobias", *al)
    sys.path.insert(0, ROOT)
    from foldseek_b200 import synth
    shutil.copy(os.path.join(OUT, "db_ss"), os.path.join(OUT, "db_ss.keep"))
    synth.mask_lower_case(os.path.join(OUT, "db_ss"))
    run("prefilter", "db_ss", "db_ss", "ref_pref_all_masked", *p["prefilter"])
    run("structurealign", "db", "db", "ref_pref_all_masked", "ref_aln_all_masked", *p["structurealign"])
    shutil.move(os.path.join(OUT, "db_ss.keep"), os.path.join(OUT, "db_ss"))
    synth.rekey_like_padded(OUT)
    run("structurealign", "q", "pad", "ref_pref_q_pad", "ref_aln_q_pad", *p["structurealign"])
    for suf in ("", "_ca", "_h"):
        for ext in ("", ".index", ".dbtype"):
            os.remove(os.path.join(OUT, "pad" + suf + ext))
    run("convertalis", "q", "db", "ref_aln_q", "ref_aln_q.m8", "--threads", "1")
    run("convertalis", "db", "db", "ref_aln_all_alt2", "ref_aln_all_alt2.m8", "--threads", "1")
    for f in os.listdir(OUT):
        if f.endswith(".m8.index") or f.endswith(".m8.dbtype"):
            os.remove(os.path.join(OUT, f))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cov":
        cov_modes()        # adds the coverage-mode outputs to an existing tests/golden/example/
    elif len(sys.argv) > 1 and sys.argv[1] == "alt":
        alt_ali()
    else:
        main()
