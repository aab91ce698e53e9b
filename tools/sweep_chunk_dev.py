"""Developer: resident / end-to-end time of the 1024-robot sweep against the chunk size (tuning key 12)."""
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

name = sys.argv[1] if len(sys.argv) > 1 else "dense_cluster_on_path"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
gen, w = wl.CLOUD_FAMILY[name]
n_pts = 100_000
host = pkg.PinnedArray((R * n_pts, 3), np.float32)
vels, poses = [], []
for r in range(R):
    rng = np.random.default_rng(wl.SEED + 7 * r)
    vels.append((float(rng.uniform(0.0, 2.0)), 0.0, float(rng.uniform(-2.0, 2.0))))
    poses.append((0.0, 0.0, 0.0))
    host.array[r * n_pts:(r + 1) * n_pts] = gen(5000 + r, n=n_pts)
offsets = np.arange(R, dtype=np.int64) * n_pts
counts = np.full(R, n_pts, np.int32)
for chunk in (8, 16, 32, 64):
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    pl.set_tuning(12, chunk)
    pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
        ts.append(time.perf_counter() - t0)
    ms, _ = pl.batch_replay(3, R)
    print(f"{name} R={R} chunk {chunk:2d}: resident {ms / 3:7.2f} ms ({ms / 3 / R * 1e3:5.1f} us/robot), end to end {np.median(ts) * 1e3:7.2f} ms")
    pl.close()
