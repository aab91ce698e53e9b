"""Developer: would two sweeps in flight at once fill the GPU better? N planner handles replay their
own resident chunk set from N host threads at the same time; compare us/robot with one handle.

    python tools/sweep_overlap_dev.py [distribution] [robots per handle] [handles]
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

name = sys.argv[1] if len(sys.argv) > 1 else "dense_cluster_on_path"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
gen, w = wl.CLOUD_FAMILY[name]
n_pts = 100_000
pls = []
for hnd in range(H):
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    host = pkg.PinnedArray((R * n_pts, 3), np.float32)
    vels, poses = [], []
    for r in range(R):
        rng = np.random.default_rng(wl.SEED + 7 * (r + hnd * R))
        vels.append((float(rng.uniform(0.0, 2.0)), 0.0, float(rng.uniform(-2.0, 2.0))))
        poses.append((0.0, 0.0, 0.0))
        host.array[r * n_pts:(r + 1) * n_pts] = gen(5000 + r + hnd * R, n=n_pts)
    offsets = np.arange(R, dtype=np.int64) * n_pts
    counts = np.full(R, n_pts, np.int32)
    pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
    pl.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
    pls.append((pl, host))
IT = 6
ms, _ = pls[0][0].batch_replay(IT, R)
print(name, "one handle alone: %.1f us/robot" % (ms / IT / R * 1e3))


def run(pl):
    pl.batch_replay(IT, R)


ths = [threading.Thread(target=run, args=(pl,)) for pl, _ in pls]
t0 = time.perf_counter()
for t in ths:
    t.start()
for t in ths:
    t.join()
dt = time.perf_counter() - t0
print(name, "%d handles at once: %.1f us/robot (wall, %d robots x %d iterations)" % (H, dt / (IT * R * H) * 1e6, R * H, IT))
