"""Developer: a few calls of the mapper / binning / critical-zone entry points on the reference's published
shapes, for ncu captures (profiles/r2_aux_ncu_summary.json)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl

pkg = ge.load_package()
N = 3
angles, ranges = wl.mapping_scan(3600)
mp = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, False, 3600, 2 * math.pi / 3600, 2.0, 0.0, 20.0)
for _ in range(N):
    mp.scan_to_grid(angles, ranges, copy=False)
for _ in range(N):
    mp.scan_to_grid_baysian(angles, ranges, copy=False)
mp.close()
pts = wl.cloud_lattice(0)
data = wl.cloud_bytes_xyz16(pts)
mpc = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, True, 3600, 2 * math.pi / 3600, 2.0, 0.1, 20.0)
for _ in range(N):
    mpc.scan_to_grid(data, 16, 16 * len(pts), 1, len(pts), 0.0, 4.0, 8.0, copy=False)
mpc.close()
ang = np.array([2 * math.pi * i / 360 for i in range(360)], np.float64)
cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.POINTCLOUD, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, ang, 0.1, 2.0, 20.0)
for _ in range(N):
    cz.check(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True)
cz.close()
a36, r36 = wl.dense_slowdown_scan(3600)
cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.LASERSCAN, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, a36, 0.1, 2.0, 20.0)
for _ in range(N):
    cz.check(r36, True)
cz.close()
print("ok")
