"""Developer: one 64-robot chunk of the batched sweep (north_star config 5) for ncu launch lists.

    python tools/sweep_prof_dev.py [distribution] [robots]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

name = sys.argv[1] if len(sys.argv) > 1 else "dense_cluster_on_path"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
gen, w = wl.CLOUD_FAMILY[name]
pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
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
pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)   # warm-up (+ heavy-cell feedback)
res = pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
ms, _ = pl.batch_replay(3, R)
print(name, R, "robots:", ms / 3, "ms per chunk replay =", ms / 3 / R * 1e3, "us/robot; found", sum(1 for x in res if x[0]))
