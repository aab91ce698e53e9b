"""Latency of every BASELINE.json config through the public C-ABI calls (host buffers in, result in
host memory) plus the device-resident replay where one exists. Writes one JSON object.

    python tools/measure_configs.py > gpurun_out/configs.json
"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import orc  # noqa: E402  (path prep of the inputs only)
import workloads as wl  # noqa: E402
from parity_util import make_planner  # noqa: E402

pkg = ge.load_package()


def lat(fn, n=300, warm=30):
    for _ in range(warm):
        r = fn()
    t = np.zeros(n)
    for i in range(n):
        t0 = time.perf_counter()
        r = fn()
        t[i] = time.perf_counter() - t0
    return dict(p50_ms=float(np.percentile(t, 50) * 1e3), p90_ms=float(np.percentile(t, 90) * 1e3),
                p99_ms=float(np.percentile(t, 99) * 1e3)), r


def planner_case(name, kw, path, seg, vel, pose, scan=None, cloud=None, replay=True):
    pl = make_planner(pkg, kw, path)
    if scan is not None:
        fn = lambda: pl.cycle_scan(vel, pose, scan[0], scan[1], seg[0], seg[1])
    else:
        fn = lambda: pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
    stats, r = lat(fn)
    out = dict(stats, slots=r.n_slots, points=r.n_points, admissible=r.n_admissible, found=r.is_found,
               slot=r.slot, tracked_segment_points=seg[1],
               sensor_points=len(scan[0]) if scan is not None else len(cloud))
    if replay and cloud is not None:
        pl.bank_alloc(4, len(cloud))
        for s in range(4):
            pl.bank_upload(s, cloud)
        pl.replay(0, 20, vel, pose, seg[0], seg[1])
        tot, ev, _ = pl.replay(0, 100, vel, pose, seg[0], seg[1], time_eval=True)
        out["resident_us_per_cycle"] = tot * 10.0
        out["eval_kernel_us"] = ev * 10.0
    out["traj_steps_per_s_e2e"] = r.n_slots * r.n_points / (stats["p50_ms"] * 1e-3)
    pl.close()
    return name, out


res = {}
# C1: 441 slots, P=10, 360-beam scan (tests/test_controllers.py shape)
path1 = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
k, v = planner_case("c1_diff_441x10_scan360", wl.cfg_c1(weights=(1, 1, 1, 1, 1)), path1,
                    wl.tracked_segment(path1, 0, 1.0), (0.0, 0.0, 0.0), (-0.51731912, 0.0, 0.0),
                    scan=wl.scan_360())
res[k] = v
# C2
path2 = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
k, v = planner_case("c2_diff_10kx50_cloud100k", wl.cfg_c2(), path2, wl.tracked_segment(path2, 0, 2.0),
                    (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), cloud=wl.cloud_bench(0))
res[k] = v
# C3: Ackermann and Omni, P = 100, ~50k slots, 3/4 circle R = 10, C2 cloud around the start pose
path3 = orc.Path(wl.circle34_points(), 0.01, 1.0)
seg3 = wl.tracked_segment(path3, 0, 2.0)
pose3 = (float(path3.X[0]), float(path3.Y[0]), math.pi / 2)
cloud3 = wl.cloud_bench(1, center=(pose3[0], pose3[1]))
for nm, ct in (("ackermann", 0), ("omni", 2)):
    for drop in (True, False):
        k, v = planner_case("c3_%s_P100_%s" % (nm, "drop" if drop else "keep_padded"),
                            wl.cfg_c3(control_type=ct, drop_samples=drop), path3, seg3, (1.0, 0.0, 0.0), pose3,
                            cloud=cloud3)
        res[k] = v
# C4: mapper + critical zone
angles, ranges = wl.mapping_scan(1080)
mp = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, False, 1080, 2 * math.pi / 1080, 2.0, 0.1, 20.0)
res["c4_mapper_scan1080_400x400"], _ = lat(lambda: mp.scan_to_grid(angles, ranges))
res["c4_mapper_bayesian_scan1080_400x400"], _ = lat(lambda: mp.scan_to_grid_baysian(angles, ranges))
mp.close()
pts = wl.cloud_lattice(0)
data = wl.cloud_bytes_xyz16(pts)
mpc = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, True, 1080, 0.01, 2.0, 0.1, 20.0)
res["c4_mapper_cloud100k_400x400"], _ = lat(lambda: mpc.scan_to_grid(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8))
mpc.close()
ang = np.array([2 * math.pi * i / 360 for i in range(360)], np.float64)
cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.POINTCLOUD, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, ang, 0.1, 2.0, 20.0)
res["c4_critical_zone_cloud100k"], f = lat(lambda: cz.check(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True))
cz.close()
a2, r2 = wl.dense_slowdown_scan()
cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.LASERSCAN, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, a2, 0.1, 2.0, 20.0)
res["c4_critical_zone_scan3600"], f = lat(lambda: cz.check(r2, True))
cz.close()
print(json.dumps(res, indent=1))
