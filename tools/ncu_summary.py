"""Summaries of an ncu --set full report for profiles/:
   python tools/ncu_summary.py <report.ncu-rep> <queries in the capture> <out prefix, e.g. profiles/r02>
   writes <prefix>_match_kernels.txt (all kernels, key metrics + top stall reasons) and <prefix>_<kernel>.txt per kernel
   (dram bytes in the form bench.py's ncu_traffic() reads)."""
import csv, re, subprocess, sys

rep, nq, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
agg = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").strip()
    short = re.sub(r"<.*", "", name)
    agg.setdefault(short, []).append(r)
with open(prefix + "_match_kernels.txt", "w") as f:
    f.write("ncu --set full --clock-control none, %s, %d queries x 1M targets (tools/prof_1m.py); one launch per kernel unless noted\n" % (rep.split("/")[-1], nq))
    for short, rs in agg.items():
        f.write("\n== %s (%d launch%s in the capture; first shown) ==\n" % (short, len(rs), "" if len(rs) == 1 else "es"))
        r = rs[0]
        for w in WANT:
            if w in idx:
                f.write("%-62s %-16s %s\n" % (w, units[idx[w]], r[idx[w]]))
        st = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h), reverse=True)
        f.write("stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st[:7]) + "\n")
scale = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for short, rs in agg.items():
    rd = sum(float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]] for r in rs)
    wr = sum(float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]] for r in rs)
    ms = sum(float(r[idx["gpu__time_duration.sum"]].replace(",", "")) for r in rs)
    with open("%s_%s.txt" % (prefix, short), "w") as f:
        f.write("# %s: sums over the %d launch(es) of one prefilter pass over %d queries x 1M targets (ncu --set full, %s)\n" % (short, len(rs), nq, rep.split("/")[-1]))
        f.write("dram__bytes_read.sum byte %.0f\n" % rd)
        f.write("dram__bytes_write.sum byte %.0f\n" % wr)
        f.write("gpu__time_duration.sum %s %.4f\n" % (units[idx["gpu__time_duration.sum"]], ms))
        f.write("queries_in_capture %d\n" % nq)
print("wrote", prefix + "_match_kernels.txt", "and", ", ".join(sorted(agg)))
