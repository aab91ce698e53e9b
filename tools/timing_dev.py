import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import orc, workloads as wl
from parity_util import make_planner
pkg = ge.load_package()
kw = wl.cfg_c2()
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
pl = make_planner(pkg, kw, path)
cloud = wl.cloud_bench(0)
for i in range(3):
    r = pl.cycle_cloud((1.0, 0, 0.0), (0.0, 0.0, 0.0), cloud, seg[0], seg[1])
c, a = pl.fetch_costs(r.n_slots)
tot = np.floor(c); roll = (c - tot) * 1e5
print("slots", len(c), "adm", a.sum())
for name, m in (("adm", a == 1), ("blocked", a == 0)):
    t = tot[m]
    print(name, "n", m.sum(), "cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f ; rollout+collision mean %.0f max %.0f" % (t.mean(), np.percentile(t, 50), np.percentile(t, 90), np.percentile(t, 99), t.max(), roll[m].mean(), roll[m].max()))
print("sum cycles", tot.sum(), "-> per warp slot (4736 warps)", tot.sum() / 4736)
idx = np.argsort(-tot)[:10]
print("slowest slots", idx, tot[idx])
