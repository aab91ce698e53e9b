"""Developer: host-side phase times of the end-to-end C2 cycle (KOMPASS_B200_HOST_PROF=1 makes the
library print them when the process exits) beside the wall-clock p50 of the call."""
import os
import sys
import time

import numpy as np

os.environ["KOMPASS_B200_HOST_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "friendly_ring"
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
gen, w = wl.CLOUD_FAMILY[name]
pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
vel, pose = (1.0, 0, 0.0), (0.0, 0.0, 0.0)
clouds = []
for s in range(16):
    c = wl.family_cloud(name, s)[0]
    pa = pkg.PinnedArray((len(c), 3), np.float32)
    pa.array[...] = c
    clouds.append(pa)
for i in range(100):
    pl.cycle_cloud(vel, pose, clouds[i % 16].array, seg[0], seg[1])
ts = []
for i in range(2000):
    t0 = time.perf_counter()
    pl.cycle_cloud(vel, pose, clouds[i % 16].array, seg[0], seg[1])
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e6
print(name, "e2e wall p50 %.1f us p90 %.1f min %.1f" % (np.percentile(ts, 50), np.percentile(ts, 90), ts.min()))
pl.close()
