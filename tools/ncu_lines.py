"""per-source-line totals (stall samples, warp instructions) of every kernel in an
`ncu --page source --csv --print-source cuda,sass` export"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kern, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Function Name":
        kern = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and kern and len(r) > 8 and r[2] == "-" and r[0].isdigit():
        a = data.setdefault(kern, {}).setdefault(int(r[0]), [r[1], 0, 0])
        a[1] += int(r[6] or 0)
        a[2] += int(r[7] or 0)
for k, lines in data.items():
    ts = sum(v[1] for v in lines.values()) or 1
    ti = sum(v[2] for v in lines.values()) or 1
    print(f"=== {k}: samples {ts} warp-instr {ti}")
    for ln, (src, s, i) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{ln:5d} {100 * s / ts:5.1f}%s {100 * i / ti:5.1f}%i  {src[:100]}")
