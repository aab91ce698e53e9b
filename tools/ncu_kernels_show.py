"""Developer: digest of tools/ncu_kernels_dev.sh launch lists (per kernel: mean us, warp instructions, ...)."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            per.setdefault((r[idc], r[kn].split("(")[0].replace("void ", "")), {})[r[mn]] = float(r[mv].replace(",", ""))
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, collections.Counter())
        a["n"] += 1
        for kk, v in m.items():
            a[kk] += v
    print(f)
    tot = tt = 0
    for k, a in agg.items():
        n = a["n"]
        print("  %-28s n=%d  %6.2f us  %8.0f inst  thr %.1f  issue %.1f%%  occ %.1f%%" % (
            k, n, a["gpu__time_duration.sum"] / n / 1e3, a["smsp__inst_executed.sum"] / n,
            a["smsp__thread_inst_executed_per_inst_executed.ratio"] / n,
            a["smsp__issue_active.avg.pct_of_peak_sustained_active"] / n,
            a["sm__warps_active.avg.pct_of_peak_sustained_active"] / n))
        tot += a["smsp__inst_executed.sum"] / n
        tt += a["gpu__time_duration.sum"] / n / 1e3
    print("  total: %.0f warp-instructions, %.1f us serialised" % (tot, tt))
