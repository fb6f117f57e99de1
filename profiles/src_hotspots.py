"""Per-source-line hot spots of one ncu capture (taken with --import-source on, kernels compiled with -lineinfo):
    ncu -i <rep> --page source --csv --print-source cuda,sass > x.csv ; python profiles/src_hotspots.py x.csv [top]
Prints, per source line, the warp-instructions executed and the stall samples (share of the kernel), sorted by instructions."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path, newline="")))
cur = None
lines = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    i_inst = hdr.index("Instructions Executed")
    i_samp = hdr.index("# Samples")
    key = (cur, ln)
    inst = int(r[i_inst]) if r[i_inst].isdigit() else 0
    samp = int(r[i_samp]) if r[i_samp].isdigit() else 0
    e = lines.setdefault(key, [0, 0, r[1].strip()])
    e[0] += inst
    e[1] += samp
tot_i = sum(v[0] for v in lines.values())
tot_s = sum(v[1] for v in lines.values())
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
byfile = {}
for (f, ln), v in lines.items():
    b = byfile.setdefault(f, [0, 0])
    b[0] += v[0]
    b[1] += v[1]
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:18s} inst {100 * b[0] / tot_i:5.1f} %  samples {100 * b[1] / max(tot_s, 1):5.1f} %")
print(f"{'file:line':24s} {'inst %':>7s} {'samp %':>7s}  source")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(ln):24s} {100 * v[0] / tot_i:7.2f} {100 * v[1] / max(tot_s, 1):7.2f}  {v[2][:110]}")
