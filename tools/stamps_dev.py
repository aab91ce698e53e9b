"""Developer: device time stamps inside k_cost_eval (needs a library built with -DKC_DBG_STAMPS at
kompass-core_b200/lib/libkompass_b200_dbg.so)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge

pkg = ge.load_package()
pkg.LIB_PATH = os.environ.get("KOMPASS_B200_LIB", pkg.LIB_PATH)
import orc
import workloads as wl
from parity_util import make_planner

kw = wl.cfg_c2()
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
pl = make_planner(pkg, kw, path)
cloud = wl.cloud_bench(0)
for i in range(5):
    pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
L = pkg.lib()
acc = []
for i in range(10):
    L.kc_planner_debug_stamps(pl._h, 1, None)
    pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
    out = (C.c_int64 * 8)()
    L.kc_planner_debug_stamps(pl._h, 0, out)
    acc.append([out[i] for i in range(6)])
a = np.median(np.array(acc), axis=0) / 1e3
print("us after kernel start: items done %.1f | all CTAs past ticket %.1f | last CTA starts %.1f | totals formed %.1f | published %.1f"
      % (a[1], a[2], a[3], a[4], a[5]))
