"""Per-source-line instruction counts / stall samples of one kernel from an ncu report (SASS page) joined with nvdisasm -gi line info.
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled function name> [top N]"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s/foldseek_b200/libfsb200.so >/dev/null 2>&1 && nvdisasm -c -gi *.cubin > all.sass 2>/dev/null" % (d, root), shell=True)
lines = open(os.path.join(d, "all.sass")).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + fn + ":"))
cur, insts = None, []          # (line tag) per instruction in order
for l in lines[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append((cur, l.strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = {h: i for i, h in enumerate(rows[hdr])}
body, seen = [], set()
for r in rows[hdr + 1:]:
    if len(r) > 5 and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); body.append(r)
body = body[:len(insts)] if len(body) > len(insts) and all("NOP" in r[1] for r in body[len(insts):]) else body
assert len(body) == len(insts), (len(body), len(insts))
agg = {}
tot_i = tot_s = 0
for (tag, txt), r in zip(insts, body):
    ie = int(r[H["Instructions Executed"]] or 0); ss = int(r[H["Warp Stall Sampling (All Samples)"]] or 0)
    a = agg.setdefault(tag, [0, 0]); a[0] += ie; a[1] += ss
    tot_i += ie; tot_s += ss
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for tag, (ie, ss) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    src = ""
    if tag:
        for base in (os.path.join(root, "foldseek_b200", "csrc"),):
            p = os.path.join(base, tag[0])
            if os.path.exists(p):
                src = open(p).read().split("\n")[tag[1] - 1].strip()[:110]
    print("%5.1f%% stalls %5.1f%% inst  %s:%s  %s" % (100.0 * ss / max(1, tot_s), 100.0 * ie / max(1, tot_i), tag[0] if tag else "?", tag[1] if tag else "?", src))
