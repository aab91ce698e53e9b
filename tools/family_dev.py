"""Developer: per-distribution timeline (tuning key 4), candidate-list statistics and resident time of
the C2 cycle over the cloud family of tests/workloads.py. Not the driver contract.

    python tools/family_dev.py [name ...]        (default: every member)
    python tools/family_dev.py --replay NAME N   (short resident replay of one member, for ncu)
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

pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
VEL, POSE = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)


def planner_for(name):
    gen, w = wl.CLOUD_FAMILY[name]
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    return pl, gen


def replay(name, n):
    pl, gen = planner_for(name)
    pl.bank_alloc(4, 100_000)
    for s in range(4):
        c = wl.family_cloud(name, s)[0]
        pl.bank_upload(s, c if len(c) else np.zeros((1, 3), np.float32)[:0])
    pl.replay(0, 2, VEL, POSE, seg[0], seg[1])  # warm-up; its result tells the next call whether heavy cells exist
    tot, _, last = pl.replay(0, n, VEL, POSE, seg[0], seg[1])
    print(name, tot / n * 1000, "us/cycle", last.slot, last.cost, last.n_admissible)


def timeline(name):
    pl, gen = planner_for(name)
    clouds = []
    for s in range(4):
        c = wl.family_cloud(name, s)[0]
        pa = pkg.PinnedArray((max(len(c), 1), 3), np.float32)
        pa.array[:len(c)] = c
        clouds.append((pa, len(c)))
    r = pl.cycle_cloud(VEL, POSE, clouds[0][0].array[:clouds[0][1]], seg[0], seg[1])
    print("==", name, "slot", r.slot, "cost", r.cost, "admissible", r.n_admissible, pl.debug_stats())
    pl.set_tuning(4, 1)
    acc = {}
    N = 24
    for i in range(N + 6):
        pa, n = clouds[i % 4]
        pl.cycle_cloud(VEL, POSE, pa.array[:n], seg[0], seg[1])
        if i < 6:
            continue
        for nm, a, b in pl.debug_timeline():
            acc.setdefault(nm, []).append((a, b))
    print("kernel                start   end    dur (us, median of %d cycles)" % N)
    for nm, v in acc.items():
        v = np.array(v)
        print("%-20s %6.1f %6.1f %6.1f" % (nm, np.median(v[:, 0]), np.median(v[:, 1]), np.median(v[:, 1] - v[:, 0])))
    pl.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--replay":
        replay(sys.argv[2], int(sys.argv[3]))
    else:
        for nm in (sys.argv[1:] or list(wl.CLOUD_FAMILY)):
            timeline(nm)
