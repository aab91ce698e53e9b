"""Developer: mapper / critical-zone end-to-end p50 from pageable vs page-locked raw clouds."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl

pkg = ge.load_package()
pts = wl.cloud_lattice(0, 100_000)
data = wl.cloud_bytes_xyz16(pts)
n = len(pts)
pinned = pkg.PinnedArray(data.shape, np.int8)
pinned.array[...] = data


def p50(fn, iters=300):
    for _ in range(20):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return np.percentile(ts, 50) * 1e3


m = pkg.LocalMapperGPU(400, 400, 0.05, (0, 0, 0), 0.0, True, 1080, 0.01, 2.0, 0.1, 20.0, 256)
angles = np.arange(0.0, 2 * np.pi, 2 * np.pi / 360)
z = pkg.CriticalZoneCheckerGPU(1, 0, (0.51, 2.0), (0.22, 0.0, 0.4), (0, 0, 0.99, 0.0), 160.0, 0.3, 0.6, angles, 0.1, 2.0, 20.0)
for name, buf in (("pageable", data), ("page-locked", pinned.array)):
    print("%-12s mapper cloud->grid p50 %.4f ms | critical zone cloud p50 %.4f ms" % (
        name, p50(lambda: m.scan_to_grid(buf, 16, n * 16, 1, n, 0, 4, 8)),
        p50(lambda: z.check(buf, 16, n * 16, 1, n, 0, 4, 8, True))))
