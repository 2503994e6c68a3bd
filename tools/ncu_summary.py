"""Summarise an .ncu-rep (ncu --set full) into text: selected raw metrics per kernel + the top stall locations.
usage: python tools/ncu_summary.py report.ncu-rep "title" > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "sm__cycles_elapsed.avg.per_second",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(title)
print("source: %s (ncu --set full --clock-control none --import-source on)\n" % rep)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("== %s" % name)
    for i, h in enumerate(hdr):
        if h in WANT:
            print("%-88s %-14s %s" % (h, units[i], r[i]))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    kname = lines[0][1]
    h = lines[1]
    ci = {x: i for i, x in enumerate(h)}
    if "# Samples" not in ci:
        continue
    data = []
    for r in lines[2:]:
        try:
            data.append((float(r[ci["# Samples"]]), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1.0
    data.sort(key=lambda x: -x[0])
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    print("== top stall locations of %s (warp stall sampling, %d samples)" % (kname[:80], tot))
    for v, r in data[:10]:
        top = sorted([(float(r[ci[s]] or 0), s) for s in stalls], reverse=True)[0]
        print("%9d %5.1f%%  exec %11s  %-72s %s" % (v, 100 * v / tot, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:72], top[1]))
    print()
