import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import orc, workloads as wl
from parity_util import make_planner
pkg = ge.load_package()
kw = wl.cfg_c2()
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
for graphs in (1, 0):
    pl = make_planner(pkg, kw, path)
    pl.set_tuning(1, graphs)
    pl.bank_alloc(8, 100000)
    for s in range(8):
        pl.bank_upload(s, wl.cloud_bench(s))
    pl.replay(0, 50, (1.0, 0, 0.0), (0.0, 0.0, 0.0), seg[0], seg[1])
    for rep in range(2):
        tot, _, last = pl.replay(0, 200, (1.0, 0, 0.0), (0.0, 0.0, 0.0), seg[0], seg[1])
        print("graphs", graphs, "replay 200 cycles -> %.1f us/cycle" % (tot * 5))
    pl.close()
