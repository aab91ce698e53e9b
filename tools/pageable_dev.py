"""Developer: latency of the entry points that take pageable sensor buffers (critical zone cloud,
mapper cloud, planner cloud cycle) - p50 of 300 calls each. Env: KOMPASS_B200_STAGE, KOMPASS_B200_COPY_THREADS."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

pkg = ge.load_package()


def p50(fn, n=300):
    for _ in range(30):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return 1e3 * float(np.percentile(t, 50))


pts = wl.cloud_lattice(0)
data = wl.cloud_bytes_xyz16(pts)
ang = np.array([2 * math.pi * i / 360 for i in range(360)], np.float64)
cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.POINTCLOUD, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, ang, 0.1, 2.0, 20.0)
a = p50(lambda: cz.check(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True))
pinned = pkg.PinnedArray(data.shape, np.int8)
pinned.array[...] = data
b = p50(lambda: cz.check(pinned.array, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True))
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
pl = make_planner(pkg, wl.cfg_c2(), path)
cloud = np.ascontiguousarray(wl.family_cloud("friendly_ring", 0)[0])
c = p50(lambda: pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1]))
pc = pkg.PinnedArray(cloud.shape, np.float32)
pc.array[...] = cloud
d = p50(lambda: pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), pc.array, seg[0], seg[1]))
print("STAGE=%s THREADS=%s: critical zone 100k pageable %.4f ms (page-locked %.4f) | cycle pageable %.4f ms (page-locked %.4f)"
      % (os.environ.get("KOMPASS_B200_STAGE", "inplace"), os.environ.get("KOMPASS_B200_COPY_THREADS", "auto"), a, b, c, d))
