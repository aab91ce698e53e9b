"""Developer: resident cycle time per bank slot for several values of one tuning key (default 13: the
survivor count up to which k_cost_eval works by (slot, point) pair).
    python tools/bypoint_dev.py DIST V1 V2 ...        python tools/bypoint_dev.py DIST key=10 V1 V2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
VEL, POSE = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)
name = sys.argv[1]
NS = 8
gen, w = wl.CLOUD_FAMILY[name]
pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
pl.bank_alloc(NS, 100_000)
for s in range(NS):
    pl.bank_upload(s, wl.family_cloud(name, s)[0])
KEY = 13
vals = sys.argv[2:]
if vals and vals[0].startswith("key="):
    KEY = int(vals[0][4:])
    vals = vals[1:]
for v in [int(x) for x in vals]:
    pl.set_tuning(KEY, v)
    pl.replay(0, 3, VEL, POSE, seg[0], seg[1])
    row = []
    for s in range(NS):
        pl.bank_alloc  # noqa
        # one slot at a time: replay n cycles starting at slot s advances through the bank, so time single cycles
        ts = []
        for _ in range(20):
            t, _, last = pl.replay(s, 1, VEL, POSE, seg[0], seg[1])
            ts.append(t * 1000)
        ts.sort()
        row.append("%.0f(%d)" % (ts[len(ts) // 2], last.n_admissible))
    tot, _, _ = pl.replay(0, 400, VEL, POSE, seg[0], seg[1])
    print(name, "key%d =" % KEY, v, "| per slot us (admissible):", " ".join(row), "| replay avg %.1f us" % (tot / 400 * 1000))
