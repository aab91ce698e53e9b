"""Developer A/B of the end-to-end C2 cycle from page-locked input over the host-path switches:
zero-copy cloud (tuning key 2), mapped result record (3), host polling of that record (8)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import orc
import workloads as wl
from parity_util import make_planner

pkg = ge.load_package()
kw = wl.cfg_c2()
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
vel, pose = (1.0, 0, 0.0), (0.0, 0.0, 0.0)
clouds = []
for s in range(16):
    pa = pkg.PinnedArray((100_000, 3), np.float32)
    pa.array[...] = wl.cloud_bench(s)
    clouds.append(pa)
ref = None
for zc, mr, poll in [(1, 1, 0), (1, 1, 1), (0, 0, 0), (1, 1, 0), (1, 1, 1)]:
    pl = make_planner(pkg, kw, path)
    pl.set_tuning(2, zc)
    pl.set_tuning(3, mr)
    pl.set_tuning(8, poll)
    for i in range(40):
        r = pl.cycle_cloud(vel, pose, clouds[i % 16].array, seg[0], seg[1])
    ts, res = [], []
    for i in range(1000):
        t0 = time.perf_counter()
        r = pl.cycle_cloud(vel, pose, clouds[i % 16].array, seg[0], seg[1])
        ts.append(time.perf_counter() - t0)
        if i < 16:
            res.append((r.slot, r.cost, r.n_admissible, r.x.tobytes()))
    if ref is None:
        ref = res
    assert res == ref, "results changed"
    ts = np.array(ts) * 1e3
    print("zero_copy=%d mapped_result=%d poll=%d: p50 %.4f p90 %.4f min %.4f ms" %
          (zc, mr, poll, np.percentile(ts, 50), np.percentile(ts, 90), ts.min()))
    pl.close()
