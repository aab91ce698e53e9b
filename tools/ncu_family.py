"""Summarise ncu launch lists (--csv, one metric per row) of the C2 cycle per cloud distribution and of
one sweep chunk into profiles/r2_cycle_ncu_summary.json.

    python tools/ncu_family.py gpurun_out/r2_launches_ OUT.json name1 name2 ... [sweep:FILE:ROBOTS]
"""
import collections
import csv
import json
import sys

M = {"gpu__time_duration.sum": "t_ns", "smsp__inst_executed.sum": "inst",
     "smsp__thread_inst_executed_per_inst_executed.ratio": "thr",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue",
     "sm__warps_active.avg.pct_of_peak_sustained_active": "occ"}


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in data:
        if len(r) > mv and r[mn] in M:
            per.setdefault((r[idc], r[kn].split("(")[0].replace("void ", "")), {})[M[r[mn]]] = float(r[mv].replace(",", ""))
    return per


def per_kernel(per):
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, collections.Counter())
        a["n"] += 1
        for key in ("t_ns", "inst", "thr", "issue", "occ"):
            a[key] += m.get(key, 0.0)
    return agg


def main():
    prefix, out = sys.argv[1], sys.argv[2]
    res = {"note": "ncu --metrics launch lists (--clock-control none): per-launch times are cold-cache and serialised, "
                   "compare shares; warp instructions are exact counts", "distributions": {}}
    for name in sys.argv[3:]:
        if name.startswith("sweep:"):
            _, path, robots = name.split(":")
            agg = per_kernel(load(path))
            robots = int(robots)
            # the tool runs the chunk set 5 times (warm-up, run, 3 replays)
            inst = sum(a["inst"] for a in agg.values()) / 5.0 / robots
            res["sweep"] = {"robots": robots, "warp_instructions_per_robot": inst,
                            "kernels": {k: {"launches": a["n"], "warp_inst_per_robot": a["inst"] / 5.0 / robots,
                                            "active_threads_per_inst": a["thr"] / a["n"],
                                            "issue_slot_busy_pct": a["issue"] / a["n"],
                                            "achieved_occupancy_pct": a["occ"] / a["n"]} for k, a in agg.items()}}
            continue
        agg = per_kernel(load(prefix + name + ".csv"))
        cycles = max(a["n"] for a in agg.values())
        ks = {}
        for k, a in agg.items():
            n = a["n"]
            ks[k] = {"launches_per_cycle": n / cycles, "us": a["t_ns"] / n / 1e3, "warp_instructions": a["inst"] / n,
                     "active_threads_per_inst": a["thr"] / n, "issue_slot_busy_pct": a["issue"] / n,
                     "achieved_occupancy_pct": a["occ"] / n}
        res["distributions"][name] = {
            "cycles_captured": cycles,
            "warp_instructions_per_cycle": sum(a["inst"] for a in agg.values()) / cycles,
            "sum_kernel_us_serialised": sum(a["t_ns"] for a in agg.values()) / cycles / 1e3, "kernels": ks}
    json.dump(res, open(out, "w"), indent=1)
    for n, d in res["distributions"].items():
        print(n, "%.2f M warp-inst/cycle, %.1f us serialised" % (d["warp_instructions_per_cycle"] / 1e6, d["sum_kernel_us_serialised"]))
    if "sweep" in res:
        print("sweep: %.2f M warp-inst/robot" % (res["sweep"]["warp_instructions_per_robot"] / 1e6))


if __name__ == "__main__":
    main()
