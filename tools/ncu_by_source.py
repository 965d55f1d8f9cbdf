"""Per-source-section view of one profiled kernel: joins `ncu --page source --csv` (SASS rows with
executed / thread-executed instruction counts and stall samples) with the line table of the same
cubin (`cuobjdump -xelf all lib.so`, `nvdisasm -g -c narrow.sm_100a.cubin`).

  python tools/ncu_by_source.py report.ncu-rep narrow.dis 'round_kernelILb0EdEE' out.txt "comment"
"""
import csv, re, subprocess, sys

rep, dis_path, func_pat, out_path = sys.argv[1:5]
comment = sys.argv[5] if len(sys.argv) > 5 else ""
dis = open(dis_path).read().split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and func_pat in l][0]
line, off2line = None, {}
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File ".*narrow.cu", line (\d+)', l)
    if m:
        line = int(m.group(1)); continue
    if "//## File" in l:
        line = -1; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = line
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, data = rows[1], rows[2:]
ai, ie, te, sa = (h.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
base = int(data[0][ai], 16)
agg, tot, totT, totS = {}, 0.0, 0.0, 0.0
for r in data:
    try:
        n, t, s = float(r[ie]), float(r[te]), float(r[sa])
    except ValueError:
        continue
    a = agg.setdefault(off2line.get(int(r[ai], 16) - base, -2), [0, 0, 0])
    a[0] += n; a[1] += t; a[2] += s
    tot += n; totT += t; totS += s
src = open(sys.argv[6] if len(sys.argv) > 6 else "scalable-ccd_b200/csrc/narrow.cu").read().split("\n")
with open(out_path, "w") as f:
    f.write(f"# {comment}\n# kernel {rows[0][1][:120]}\n")
    f.write(f"# warp instructions {tot:.0f}, threads per instruction {totT / tot:.2f}, stall samples {totS:.0f}\n")
    f.write("# per source line of narrow.cu (top 60 by stall samples): share of executed warp instructions, "
            "active threads per instruction, share of stall samples\n")
    for ln, (n, t, s) in sorted(agg.items(), key=lambda x: -x[1][2])[:60]:
        text = src[ln - 1].strip()[:100] if ln > 0 else "(no line info)"
        f.write(f"L{ln:5d} inst {n / tot * 100:5.1f}%  thr/inst {t / max(n, 1):5.1f}  samples {s / totS * 100:5.1f}%  | {text}\n")
print("wrote", out_path)
