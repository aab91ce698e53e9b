"""Developer timeline of one C2 cycle in the real three-stream pipeline (tuning key 4): start/end of
every kernel from CUDA events on the stream it runs on. Not the driver contract."""
import os
import sys

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
pl = make_planner(pkg, kw, path)
clouds = []
for s in range(8):
    pa = pkg.PinnedArray((100_000, 3), np.float32)
    pa.array[...] = wl.cloud_bench(s)
    clouds.append(pa)
pl.set_tuning(4, 1)
acc = {}
N = 40
for i in range(N + 10):
    r = pl.cycle_cloud((1.0, 0, 0.0), (0.0, 0.0, 0.0), clouds[i % 8].array, seg[0], seg[1])
    if i < 10:
        continue
    for name, a, b in pl.debug_timeline():
        acc.setdefault(name, []).append((a, b))
print("kernel                start   end    dur (us, median of %d cycles)" % N)
for name, v in acc.items():
    v = np.array(v)
    print("%-20s %6.1f %6.1f %6.1f" % (name, np.median(v[:, 0]), np.median(v[:, 1]), np.median(v[:, 1] - v[:, 0])))
