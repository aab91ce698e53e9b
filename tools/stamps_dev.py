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
import workloads as wl
from bench import ProductPath, make_planner

path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
L = pkg.lib()
for name in (sys.argv[1:] or ["friendly_ring"]):
    gen, w = wl.CLOUD_FAMILY[name]
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    cloud = wl.family_cloud(name, int(os.environ.get("KC_CLOUD", "0")))[0]  # KC_CLOUD: which member of the bank
    pa = pkg.PinnedArray((max(len(cloud), 1), 3), np.float32)
    pa.array[:len(cloud)] = cloud
    cloud = pa.array[:len(cloud)]
    for i in range(5):
        pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
    acc = []
    for i in range(10):
        L.kc_planner_debug_stamps(pl._h, 1, None)
        pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
        out = (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, 0, out)
        g0, g3 = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, -12, g0)
        L.kc_planner_debug_stamps(pl._h, -3, g3)
        g11 = (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, -11, g11)
        acc.append([out[i] for i in range(8)] + [g3[i] - g0[0] for i in range(4)] + [g11[i] for i in range(8)])
    a = np.median(np.array(acc), axis=0)
    print("%s: work items (slots) %d by_point %d | us after kernel start: items done %.1f | all CTAs past ticket %.1f | "
          "last CTA starts %.1f | totals formed %.1f | published %.1f"
          % (name, a[6], a[7], a[1] / 1e3, a[2] / 1e3, a[3] / 1e3, a[4] / 1e3, a[5] / 1e3))
    k = a[12:20]
    print("   whole-warp searches of the exact stage: candidate list %d (%.2f us each) | list-less cell %d (%.2f us each) | outside the window %d (%.2f us each)"
          % (k[1], k[0] / max(k[1], 1) / 1965, k[3], k[2] / max(k[3], 1) / 1965, k[5], k[4] / max(k[5], 1) / 1965))
    print("   publish: winner known %.1f | slot decoded %.1f | stores issued %.1f | fenced %.1f" % tuple(a[8:12] / 1e3))
    pl.close()
