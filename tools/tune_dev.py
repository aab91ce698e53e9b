"""Developer sweep of one tuning key on the resident C2 / C3 cycles: python tools/tune_dev.py KEY V1 V2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import orc
import workloads as wl
from parity_util import make_planner

key, values = int(sys.argv[1]), [int(v) for v in sys.argv[2:]]
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
    for v in values:
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(key, v)
        pl.bank_alloc(8, 100000)
        for s in range(8):
            pl.bank_upload(s, wl.cloud_bench(s, center=pose[:2]))
        pl.replay(0, 16, vel, pose, seg[0], seg[1])
        tot, _, last = pl.replay(0, 200, vel, pose, seg[0], seg[1])
        print(name, "key", key, "=", v, "%.1f us/cycle" % (tot * 5), "slot", last.slot)
        pl.close()
