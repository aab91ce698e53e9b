import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import orc, workloads as wl
from parity_util import make_planner
pkg = ge.load_package()
kw = wl.cfg_c2()
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
pl = make_planner(pkg, kw, path)
cloud = wl.cloud_bench(0)
r = pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
c, a = pl.fetch_costs(r.n_slots); p = pl.fetch_pruned(r.n_slots)
print("admissible", a.sum(), "pruned", p.sum(), "winner cost", r.cost)
pl.set_tuning(7, 0)
r0 = pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
c0, a0 = pl.fetch_costs(r0.n_slots)
ca = np.sort(c0[a0 == 1])
print("exact totals: min %.4f p1 %.4f p10 %.4f p50 %.4f max %.4f" % (ca[0], ca[len(ca)//100], ca[len(ca)//10], ca[len(ca)//2], ca[-1]))
lb = c[(a == 1) & (p == 1)]
print("slack of pruned bounds vs exact: median %.4f" % np.median(c0[(a == 1) & (p == 1)] - lb) if len(lb) else "none pruned")
pl.set_tuning(7, 1); pl.set_tuning(4, 1)
for i in range(5):
    pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
for n, s, e in pl.debug_timeline(): print("%-18s %7.1f %7.1f %6.1f" % (n, s, e, e - s))
for mask in (0, 1):
    pl2 = make_planner(pkg, kw, path)
    pl2.set_tuning(5, mask); pl2.set_tuning(4, 1)
    for i in range(5):
        pl2.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
    print("reach mask", mask, [(n, round(e - s, 1)) for n, s, e in pl2.debug_timeline() if n.startswith("k_cost")])
    pl2.close()
