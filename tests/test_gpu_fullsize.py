"""Full-size parity of the BASELINE configurations against the CPU oracle.

configs[1] (10 201-slot grid x 50 points vs 100 000 points) over the whole cloud family of
tests/workloads.py and configs[2] (Ackermann + omni, ~50 k slots x 100 points, drop and keep) with
ALL FIVE weights, through the C-ABI, in both evaluation modes:
  exact mode   (tuning 7 = 0): n_admissible, every admissible row, winner slot/cost/rows and EVERY
               per-slot cost's bits;
  default mode (branch and bound): the same winner record; evaluated slots carry the oracle's bits,
               pruned slots report a lower bound below their true total and really lose.
The oracle runs on all host cores (orc.cost_evaluate(n_threads=os.cpu_count()); a config-2 cycle
takes 1.5-6 s on 8-32 cores, a config-3 cycle about a minute)."""
import os

import numpy as np
import pytest

import orc
import workloads as wl
from parity_util import assert_cycle_parity, make_planner, run_oracle_cycle
from test_gpu_planner import check_pruned_cycle

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _full_parity(pkg, kw, path, seg, vel, pose, cloud, min_admissible):
    ref = run_oracle_cycle(kw, path, seg, vel, pose, cloud=cloud, n_threads=THREADS)
    assert ref["n_admissible"] >= min_admissible, ref["n_admissible"]
    slots = ref["samples"]["slots"]
    stats = {}
    for prune in (0, 1):  # 1 = the default policy (on at these sizes)
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(7, prune)
        try:
            got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
            costs, adm = pl.fetch_costs(got.n_slots)
            prn = pl.fetch_pruned(got.n_slots)
            if prune == 0:
                assert not prn.any()
                assert_cycle_parity(got, ref, costs, adm, 1e-4)
                g = costs[slots]
                same = g.view(np.uint32) == ref["costs"].view(np.uint32)
                assert same.all(), f"{(~same).sum()} of {len(g)} per-slot costs differ in bits"
                assert np.all(costs[adm == 0] == np.finfo(np.float32).max)
                # every admissible row, bit for bit, in enumeration order
                rows = pl.generate_trajectories(vel, pose, cloud=cloud)
                assert np.array_equal(rows["slots"], slots)
                for k in ("vx", "vy", "omega", "x", "y"):
                    assert np.array_equal(rows[k].view(np.uint32), ref["samples"][k].view(np.uint32)), k
            else:
                check_pruned_cycle(got, ref, costs, adm, prn)
                stats["pruned"] = int(prn.sum())
        finally:
            pl.close()
    print(f"admissible {ref['n_admissible']} winner {ref['slot']} cost {ref['cost']:.6f} pruned {stats.get('pruned')}")
    return ref


@pytest.mark.parametrize("family", list(wl.CLOUD_FAMILY))
def test_c2_full_size_all_weights(pkg, family):
    """BASELINE configs[1] at full size: 9 900 slots x 50 points vs 100 000 points, every member of
    the cloud family (SURVEY 8d's own C2 cloud first)."""
    cloud, w = wl.family_cloud(family, 3)
    kw = wl.cfg_c2() if w is None else wl.cfg_c2(weights=w)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    _full_parity(pkg, kw, path, seg, (1.0, 0.0, 0.0), (0.0, 0.0, 0.0), cloud, 100)


@pytest.mark.parametrize("ctrl,drop", [(0, True), (0, False), (2, True), (2, False)])
def test_c3_full_size_all_weights(pkg, ctrl, drop):
    """BASELINE configs[2] at full size: Ackermann (50 175 slots) and omni (47 433 slots) x 100
    points vs 100 000 points, all five weights, dropping and keeping (zero-padded) samples."""
    kw = wl.cfg_c3(control_type=ctrl, drop_samples=drop)
    path = orc.Path(wl.circle34_points(), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 4.0)
    pose = (float(path.X[0]), float(path.Y[0]), 1.45)
    gen = wl.cloud_bench if drop else wl.cloud_pillars
    cloud = gen(9, n=100_000, center=pose[:2])
    _full_parity(pkg, kw, path, seg, (1.0, 0.0, 0.2), pose, cloud, 5000)
