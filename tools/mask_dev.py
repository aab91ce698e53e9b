"""Developer A/B of the analytic reach mask (tuning key 5) on the resident C2 / C3 cycles."""
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
for name in ("c2", "c3"):
    if name == "c2":
        kw = wl.cfg_c2()
        path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
        seg = wl.tracked_segment(path, 0, 2.0)
        pose = (0.0, 0.0, 0.0)
    else:
        kw = wl.cfg_c3(control_type=0)
        path = orc.Path(wl.circle34_points(), 0.01, 1.0)
        seg = wl.tracked_segment(path, 0, 4.0)
        pose = (float(path.X[0]), float(path.Y[0]), 0.0)
    vel = (1.0, 0, 0.0)
    out = []
    for mask in (0, 1, 0, 1):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(5, mask)
        pl.bank_alloc(8, 100000)
        for s in range(8):
            pl.bank_upload(s, wl.cloud_bench(s, center=pose[:2]))
        pl.replay(0, 16, vel, pose, seg[0], seg[1])
        tot, _, last = pl.replay(0, 200, vel, pose, seg[0], seg[1])
        r = pl.cycle_cloud(vel, pose, wl.cloud_bench(0, center=pose[:2]), seg[0], seg[1])
        st = pl.debug_stats()
        c, a = pl.fetch_costs(r.n_slots)
        out.append((r.slot, r.cost, c.tobytes()))
        print(name, "mask", mask, "%.1f us/cycle" % (tot * 5), "slot", last.slot, "stats", dict(st) if isinstance(st, dict) else st)
        pl.close()
    assert all(o == out[0] for o in out), "mask changed the results"
