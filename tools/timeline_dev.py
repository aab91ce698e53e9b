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
t = c.astype(np.float64)
t = (t - t.min()) / 1e3
print("slot end times (us after the first slot end): p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(t, [10, 50, 90, 99, 100])))
h, e = np.histogram(t, bins=12)
print("histogram", list(zip(np.round(e[:-1], 1), h)))
cta = t.reshape(-1)[: (len(t) // 8) * 8].reshape(-1, 8).max(axis=1)
print("CTA end times by index (every 100th)", np.round(cta[::100], 1))
