"""Developer: phase timing inside k_rollout_collide (needs a library built with -DKC_DBG_STAMPS:
tools/build_variant.sh dbg -DKC_DBG_STAMPS; run with KOMPASS_B200_LIB=.../libkompass_b200_dbg.so)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
for name in (sys.argv[1:] or ["friendly_ring", "dense_cluster_on_path"]):
    gen, w = wl.CLOUD_FAMILY[name]
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    cloud = wl.family_cloud(name, 0)[0]
    for i in range(5):
        pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
    L = pkg.lib()
    acc = []
    for i in range(10):
        L.kc_planner_debug_stamps(pl._h, 1, None)
        pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
        out = (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, -1, out)
        out2 = (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, -2, out2)
        acc.append([out[i] for i in range(8)] + [out2[i] for i in range(8)])
    a = np.median(np.array(acc, dtype=np.float64), axis=0)
    n_cta = 310
    warps = max(a[4], 1)
    print(f"{name}: per CTA (cycles): decode {a[0]/n_cta:.0f} | phase A done {a[1]/n_cta:.0f} | per warp: phase B starts {a[5]/warps:.0f} "
          f"| end avg {a[2]/warps:.0f} max {a[3]:.0f} | warps {warps:.0f} | kernel span {(a[7]-a[6])/1e3:.1f} us")
    print("   thread 0 of every CTA (cycles): first barrier passed %.0f | slot decoded %.0f | increments of window 0 stored %.0f | "
          "barrier %.0f | chain of window 0 done %.0f | barrier %.0f" % tuple(a[8 + i] / n_cta for i in (4, 5, 0, 1, 2, 3)))
    pl.close()
