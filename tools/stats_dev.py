"""developer: candidate-list statistics of one C2 cycle"""
import sys, os
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
r = pl.cycle_cloud((1.0, 0, 0.0), (0.0, 0.0, 0.0), cloud, seg[0], seg[1])
print(r.slot, r.cost, pl.debug_stats())
