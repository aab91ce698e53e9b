"""Developer micro-bench (not the driver contract): C2 cycle timings through the C-ABI."""
import sys, os, time
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
vel, pose = (1.0, 0, 0.0), (0.0, 0.0, 0.0)
for i in range(5):
    r = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
print("slots", r.n_slots, "adm", r.n_admissible, "P", r.n_points, "found", r.is_found, "slot", r.slot, "cost", r.cost)
ts = []
for i in range(200):
    t0 = time.perf_counter(); r = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1]); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("e2e ms p50 %.3f p90 %.3f p99 %.3f min %.3f" % (np.percentile(ts, 50), np.percentile(ts, 90), np.percentile(ts, 99), ts.min()))
pl.bank_alloc(8, 100000)
for s in range(8): pl.bank_upload(s, wl.cloud_bench(s))
for rep in range(2):
    tot, ev, last = pl.replay(0, 100, vel, pose, seg[0], seg[1], time_eval=True)
    print("replay(with eval events) 100 cycles: total %.3f ms -> %.1f us/cycle ; eval kernel %.1f us/cycle" % (tot, tot * 10, ev * 10))
    tot, _, last = pl.replay(0, 100, vel, pose, seg[0], seg[1])
    print("replay 100 cycles: total %.3f ms -> %.1f us/cycle" % (tot, tot * 10))
print("launches", pl.launch_count)
