"""Developer: device-side timeline of one resident cycle from %globaltimer stamps inside every kernel
(first CTA in / last CTA out), taken in the real graph-launched three-stream pipeline without events.
Needs a library built with -DKC_DBG_STAMPS (tools/build_variant.sh dbg -DKC_DBG_STAMPS; run with
KOMPASS_B200_LIB=.../libkompass_b200_dbg.so).

    python tools/cycle_stamps_dev.py [distribution ...]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import workloads as wl
from bench import ProductPath, make_planner

NAMES = ["k_prep_points", "k_scan_dist", "k_scatter", "k_cell_cand", "k_cell_cand_heavy", "k_path_cand",
         "k_dilate", "k_rollout_collide", "k_cost_bounds", "k_cost_split", "k_cost_eval"]
pkg = ge.load_package()
path = ProductPath(pkg, wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
VEL, POSE = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)
ONLY = [int(a[7:]) for a in sys.argv if a.startswith("--slot=")]  # one cloud of the bank only
E2E = "--e2e" in sys.argv  # the call a user makes: page-locked host cloud read in place, mapped result record
import time
for name in ([a for a in sys.argv[1:] if not a.startswith("--")] or ["friendly_ring", "dense_cluster_on_path"]):
    gen, w = wl.CLOUD_FAMILY[name]
    pl = make_planner(pkg, wl.cfg_c2() if w is None else wl.cfg_c2(weights=w), path)
    pl.bank_alloc(8, 100_000)
    for s in range(8):
        pl.bank_upload(s, wl.family_cloud(name, s)[0])
    pl.replay(0, 8, VEL, POSE, seg[0], seg[1])
    L = pkg.lib()
    acc = []
    phases = []
    cats = []
    pins = []
    for s in range(4):
        c = wl.family_cloud(name, s)[0]
        pa = pkg.PinnedArray((max(len(c), 1), 3), np.float32)
        pa.array[:len(c)] = c
        pins.append(pa.array[:len(c)])
    wall = []
    for i in range(12):
        L.kc_planner_debug_stamps(pl._h, 1, None)
        if E2E:
            t0 = time.perf_counter()
            pl.cycle_cloud(VEL, POSE, pins[i % 4], seg[0], seg[1])
            wall.append((time.perf_counter() - t0) * 1e6)
        else:
            pl.replay(ONLY[0] if ONLY else i % 4, 1, VEL, POSE, seg[0], seg[1])
        v = []
        for g in (4, 5, 6, 7):
            out = (C.c_int64 * 8)()
            L.kc_planner_debug_stamps(pl._h, -g, out)
            v += [np.uint64(out[j] & 0xFFFFFFFFFFFFFFFF) for j in range(8)]
        cat = []
        for g in (8, 9):
            out = (C.c_int64 * 8)()
            L.kc_planner_debug_stamps(pl._h, -g, out)
            cat += [int(out[j]) for j in range(8)]
        cats.append(cat)
        out = (C.c_int64 * 8)()
        L.kc_planner_debug_stamps(pl._h, -10, out)
        phases.append([int(out[j]) for j in range(8)])
        acc.append(v)
    a = np.array(acc, dtype=np.uint64)
    print("==", name, ("e2e wall median %.1f us" % np.median(wall)) if E2E else "resident")
    t0s = a[:, 0:22:2].astype(np.float64)
    t0s[t0s > 1e19] = np.nan
    base = np.nanmin(t0s, axis=1)
    for k, nm in enumerate(NAMES):
        b = a[:, 2 * k].astype(np.float64)
        e = a[:, 2 * k + 1].astype(np.float64)
        if np.all(b > 1e19):
            continue
        print("  %-20s in %6.1f  out %6.1f  span %6.1f us" % (
            nm, np.median(b - base) / 1e3, np.median(e - base) / 1e3, np.median(e - b) / 1e3))
    ph = np.median(np.array(phases, dtype=np.float64), axis=0)
    print("  k_scatter: slowest classifying CTA (us): staged %.2f | columns walked %.2f | done %.2f (cap %d rows); slowest scatter CTA %.2f" % (
        ph[0] / 1965, ph[1] / 1965, ph[2] / 1965, ph[5], ph[4] / 1965))
    cc = np.array(cats, dtype=np.float64)
    cand_in = a[:, 6].astype(np.float64)
    for c, nm in enumerate(["no walk", "staged <= 64", "staged <= 384", "second walk"]):
        n = np.median(cc[:, 4 * c + 2])
        if n > 0:
            print("  k_cell_cand warps '%s': %d, mean %.2f us, max %.2f us, last one starts %.1f us after the kernel" % (
                nm, n, np.median(cc[:, 4 * c] / np.maximum(cc[:, 4 * c + 2], 1)) / 1965.0,
                np.median(cc[:, 4 * c + 1]) / 1965.0, np.median(cc[:, 4 * c + 3] - cand_in) / 1e3))
    pl.close()
