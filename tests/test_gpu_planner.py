"""GPU parity tests of the DWA hot path: CUDA (through the C-ABI) vs the CPU oracle on the same
seeded inputs. Bars (north_star): selected trajectory index bit-exact, stored rollouts bit-exact,
per-trajectory costs within 1e-4 relative (we in fact expect identical bits and report the max)."""
import math

import numpy as np
import pytest

import orc
import workloads as wl
from parity_util import assert_cycle_parity, make_planner, run_oracle_cycle

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # tolerance stated by the reference's own CPU<->GPU parity harness (test_cost_parity.py:32)


def check_pruned_cycle(got, ref, costs, adm, prn):
    """Branch and bound on (the default): the winner is the oracle's, every slot that was evaluated
    has the oracle's bits, every pruned slot reports a valid lower bound and really loses."""
    assert_cycle_parity(got, ref)
    if not ref["found"]:
        return
    slots = ref["samples"]["slots"]
    assert np.array_equal(np.flatnonzero(adm), np.sort(slots))
    g, p, r = costs[slots], prn[slots].astype(bool), ref["costs"]
    assert np.array_equal(g[~p].view(np.uint32), r[~p].view(np.uint32)), \
        f"{(g[~p] != r[~p]).sum()} of {(~p).sum()} evaluated costs differ in bits"
    assert not p[slots == ref["slot"]].any()
    if p.any():
        assert np.all(g[p] <= r[p]), "a pruned slot's bound exceeds its true total"
        assert np.all(r[p] > np.float32(ref["cost"])), "a pruned slot would have won or tied"
    assert not prn[adm == 0].any()


def check_cycle(pkg, kw, path, seg, vel, pose, scan=None, cloud=None, exact_costs=True, tuning=None):
    """The cycle through the C-ABI against the oracle, twice: every slot evaluated exactly (tuning key
    7 = 0: all per-slot costs must carry the oracle's bits) and with the default branch and bound."""
    ref = run_oracle_cycle(kw, path, seg, vel, pose, scan=scan, cloud=cloud)
    got = None
    for prune in (0, 2):  # 2 = branch and bound forced on (the default skips it below 2048 slots)
        pl = make_planner(pkg, kw, path)
        if tuning is not None:
            pl.set_tuning(*tuning)
        pl.set_tuning(7, prune)
        try:
            if scan is not None:
                got = pl.cycle_scan(vel, pose, scan[0], scan[1], seg[0], seg[1])
            else:
                got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
            costs, adm = pl.fetch_costs(got.n_slots)
            prn = pl.fetch_pruned(got.n_slots)
            if prune:
                check_pruned_cycle(got, ref, costs, adm, prn)
            else:
                assert not prn.any()
                assert_cycle_parity(got, ref, costs, adm, RTOL)
                if exact_costs and ref["found"]:
                    g = costs[ref["samples"]["slots"]]
                    assert np.array_equal(g.view(np.uint32), ref["costs"].view(np.uint32)), \
                        f"{(g != ref['costs']).sum()} of {len(g)} costs differ in bits"
        finally:
            pl.close()
    return got, ref


def test_c1_default_weights(pkg):
    kw = wl.cfg_c1()
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    got, ref = check_cycle(pkg, kw, path, seg, (0, 0, 0), (-0.51731912, 0.0, 0.0), scan=wl.scan_360())
    assert got.n_slots == 20 * 21 or got.n_slots > 0


def test_c1_all_weights(pkg):
    kw = wl.cfg_c1(weights=(1.0, 1.0, 1.0, 1.0, 1.0))
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    check_cycle(pkg, kw, path, seg, (0.3, 0, 0.2), (-0.51731912, 0.0, 0.4), scan=wl.scan_360(3))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c2_reduced_cloud(pkg, seed):
    """config 2 shape at a size the oracle finishes in seconds: 41x41 slots x 50 pts vs 20k points"""
    kw = wl.cfg_c2(n_lin=40, n_ang=40)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = wl.cloud_c2(seed, n=20_000)
    got, ref = check_cycle(pkg, kw, path, seg, (1.0, 0, 0.0), (0.0, 0.0, 0.0), cloud=cloud)
    assert 0 < got.n_admissible < got.n_slots  # intruder wedge removes some samples


@pytest.mark.parametrize("cap", [0, 300])
def test_generic_obstacle_search_path(pkg, cap):
    """Candidate pool forced empty (every cell falls back to the generic exact search) or tiny (the
    pool overflows part-way: both strategies inside one cycle). Same bits either way."""
    kw = wl.cfg_c2(n_lin=30, n_ang=30)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = wl.cloud_c2(5, n=15_000)
    ref = run_oracle_cycle(kw, path, seg, (1.0, 0, 0.3), (0.0, 0.0, 0.0), cloud=cloud)
    results = []
    for tuning in (None, (0, cap)):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(5, 0)  # lists for the whole query window: "generic" then only means pool overflow
        pl.set_tuning(7, 0)  # every slot evaluated exactly: all per-slot costs are compared
        if tuning:
            pl.set_tuning(*tuning)
        got = pl.cycle_cloud((1.0, 0, 0.3), (0.0, 0.0, 0.0), cloud, seg[0], seg[1])
        costs, adm = pl.fetch_costs(got.n_slots)
        assert_cycle_parity(got, ref, costs, adm, RTOL)
        st = pl.debug_stats()
        results.append((costs.copy(), st))
        pl.close()
    assert np.array_equal(results[0][0].view(np.uint32), results[1][0].view(np.uint32))
    assert results[0][1]["generic_cells"] == 0 and results[0][1]["listed_cells"] > 0
    assert results[1][1]["generic_cells"] > 0
    if cap == 0:
        assert results[1][1]["listed_cells"] == 0


@pytest.mark.parametrize("case", ["forward", "reverse_window", "full_turn", "yawed_pose_ackermann", "box_keep"])
def test_reach_mask_never_changes_results(pkg, case):
    """Tuning key 5: candidate lists only inside the analytic reach set of the velocity window. The
    mask is a guess (stray queries take the generic exact search), so every cost must keep its bits
    and most query cells must really be skipped."""
    kw = wl.cfg_c2(n_lin=40, n_ang=40)
    vel, pose = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    if case == "reverse_window":  # window straddles zero: forward and reverse lobes
        vel = (0.2, 0.0, -1.0)
    elif case == "full_turn":  # omega * T beyond pi: the polygon wraps around
        kw.update(prediction_horizon=2.0, omega=(6.0, 300.0, 300.0))
    elif case == "yawed_pose_ackermann":
        kw.update(control_type=0)
        vel, pose = (1.5, 0.0, 0.5), (0.7, -0.4, 2.3)
    elif case == "box_keep":
        kw.update(shape=1, dims=(0.5, 0.3, 0.4), drop_samples=False)
        vel = (-0.5, 0.0, 0.8)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0 * kw["prediction_horizon"])
    cloud = wl.cloud_bench(21, n=30_000, center=pose[:2])
    out = []
    for mask in (0, 1):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(5, mask)
        pl.set_tuning(7, 0)  # every slot evaluated exactly: all per-slot costs are compared
        got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
        costs, adm = pl.fetch_costs(got.n_slots)
        out.append((got.slot, np.float32(got.cost), got.n_admissible, costs.copy(), adm.copy(), pl.debug_stats()))
        pl.close()
    assert out[0][:3] == out[1][:3] and out[0][2] > 100
    assert np.array_equal(out[0][3].view(np.uint32), out[1][3].view(np.uint32))
    assert np.array_equal(out[0][4], out[1][4])
    full, masked = out[0][5], out[1][5]
    assert full["generic_cells"] == 0 and full["listed_cells"] == full["query_cells"]
    if case != "full_turn":
        assert masked["listed_cells"] < 0.7 * full["listed_cells"]
    ref = run_oracle_cycle(kw, path, seg, vel, pose, cloud=cloud, max_traj=40)
    assert np.array_equal(out[1][3][ref["samples"]["slots"][:40]].view(np.uint32), ref["costs"].view(np.uint32))


@pytest.mark.parametrize("family", ["dense_cluster_on_path", "pillars_in_reach"])
def test_heavy_cell_kernel_never_changes_results(pkg, family):
    """Tuning key 10: query cells whose search disc holds more than N points get only their exact centre
    distance (k_cell_cand_heavy, one CTA per cell) and no candidate list; exact queries that land in
    them run the warp-cooperative search. Never (0), nearly every listed cell (16), the adaptive default
    (-1: first cycle builds heavy cells in place and reports them, the second cycle launches the
    kernel): every cost keeps its bits and the oracle agrees."""
    kw = wl.cfg_c2(n_lin=40, n_ang=40)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud, _ = wl.family_cloud(family, 11, n=60_000)
    vel, pose = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    outs = []
    for thr, prune in ((0, 0), (16, 0), (-1, 0), (16, 2), (-1, 2)):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(10, thr)
        pl.set_tuning(7, prune)  # 0: every slot evaluated exactly, all per-slot costs are compared
        for rep in range(2):  # adaptive policy: the second cycle runs with the first one's feedback
            got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
            costs, adm = pl.fetch_costs(got.n_slots)
            st = pl.debug_stats()
            outs.append((got.slot, np.float32(got.cost), got.n_admissible, costs.copy(), adm.copy(), st, thr, prune, rep))
        pl.close()
    exact = [o for o in outs if o[7] == 0]
    for o in outs[1:]:
        assert o[:3] == outs[0][:3], (o[6:], o[:3], outs[0][:3])
        assert np.array_equal(o[4], outs[0][4])
        assert o[5]["listed_cells"] + o[5]["generic_cells"] == outs[0][5]["listed_cells"] + outs[0][5]["generic_cells"]
    for o in exact[1:]:
        assert np.array_equal(o[3].view(np.uint32), exact[0][3].view(np.uint32)), o[6:]
    assert outs[0][2] > 100
    # the heavy path really ran: fewer listed cells once heavy cells carry no list
    always = [o for o in exact if o[6] == 16][0]
    adaptive2 = [o for o in exact if o[6] == -1 and o[8] == 1][0]
    assert always[5]["listed_cells"] < exact[0][5]["listed_cells"]
    if family == "dense_cluster_on_path":
        assert adaptive2[5]["listed_cells"] < exact[0][5]["listed_cells"]
    ref = run_oracle_cycle(kw, path, seg, vel, pose, cloud=cloud, max_traj=60)
    assert np.array_equal(adaptive2[3][ref["samples"]["slots"][:60]].view(np.uint32), ref["costs"].view(np.uint32))


@pytest.mark.parametrize("family", ["friendly_ring", "pillars_in_reach", "clutter_in_reach"])
def test_candidate_list_modes_never_change_results(pkg, family):
    """Tuning key 11: per-cell candidate lists always (1, the default), never (0: every exact query
    searches its own disc) or only when every slot is evaluated exactly (-1), with and without the branch
    and bound: same winner record; identical per-slot bits whenever every slot is evaluated."""
    kw = wl.cfg_c2(n_lin=50, n_ang=50)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud, _ = wl.family_cloud(family, 17, n=50_000)
    vel, pose = (1.0, 0.0, 0.1), (0.0, 0.0, 0.0)
    outs = []
    for lists in (1, 0, -1):
        for prune in (0, 2):
            pl = make_planner(pkg, kw, path)
            pl.set_tuning(11, lists)
            pl.set_tuning(7, prune)
            got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
            costs, adm = pl.fetch_costs(got.n_slots)
            outs.append((got.slot, np.float32(got.cost), got.n_admissible, costs.copy(), adm.copy(), pl.debug_stats(), lists, prune))
            pl.close()
    for o in outs[1:]:
        assert o[:3] == outs[0][:3], o[6:]
        assert np.array_equal(o[4], outs[0][4])
    exact = [o for o in outs if o[7] == 0]
    for o in exact[1:]:
        assert np.array_equal(o[3].view(np.uint32), exact[0][3].view(np.uint32)), o[6:]
    assert [o for o in outs if o[6] == 0][0][5]["listed_cells"] == 0
    assert exact[0][5]["listed_cells"] > 0 and outs[0][2] > 100
    # mode -1: lists in exact mode, none under the branch and bound
    d = {(o[6], o[7]): o[5]["listed_cells"] for o in outs}
    assert d[(-1, 0)] == d[(1, 0)] and d[(-1, 2)] == 0
    ref = run_oracle_cycle(kw, path, seg, vel, pose, cloud=cloud, max_traj=50)
    assert np.array_equal(exact[1][3][ref["samples"]["slots"][:50]].view(np.uint32), ref["costs"].view(np.uint32))


def test_long_horizon_long_segment(pkg):
    """P = 150 points per trajectory (five 32-point batches) and a 701-point tracked segment (more
    than one 32-window chunk of the two-level path search)."""
    kw = wl.cfg_c2(n_lin=16, n_ang=16)
    kw.update(prediction_horizon=3.0, time_step=0.02)
    path = orc.Path(wl.circle34_points(R=6.0, n=120), 0.01, 1.0)
    seg = wl.tracked_segment(path, 40, 7.0)
    assert seg[1] > 512
    cloud = wl.cloud_c2(7, n=6_000, center=(float(path.X[40]), float(path.Y[40])))
    pose = (float(path.X[40]) + 0.05, float(path.Y[40]) - 0.03, 1.2)
    got, ref = check_cycle(pkg, kw, path, seg, (1.2, 0, 0.4), pose, cloud=cloud)
    assert got.n_points == 150


def test_pinned_input_and_plain_launches_give_the_same_cycle(pkg):
    """The ways a cycle can be fed/launched are interchangeable: pageable input staged through the
    handle, page-locked caller buffers (read in place by the kernel, or DMA-ed first), winner record
    written to mapped host memory or copied back, CUDA graph replay on or off."""
    kw = wl.cfg_c2(n_lin=30, n_ang=30)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = wl.cloud_c2(9, n=12_000)
    ranges, angles = wl.scan_360(4)
    pinned_cloud = pkg.PinnedArray(cloud.shape, np.float32)
    pinned_cloud.array[...] = cloud
    pr, pa = pkg.PinnedArray(ranges.shape, np.float64), pkg.PinnedArray(angles.shape, np.float64)
    pr.array[...], pa.array[...] = ranges, angles
    outs = []
    for graphs, in_place, mapped in ((1, 1, 1), (0, 1, 1), (1, 0, 0), (1, 1, 0), (0, 0, 1)):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(1, graphs)
        pl.set_tuning(2, in_place)
        pl.set_tuning(3, mapped)
        for _ in range(2):  # second pass replays the cached graph
            a = pl.cycle_cloud((1.0, 0, 0.2), (0.0, 0.0, 0.0), cloud, seg[0], seg[1])
            ca, _ = pl.fetch_costs(a.n_slots)
            b = pl.cycle_cloud((1.0, 0, 0.2), (0.0, 0.0, 0.0), pinned_cloud.array, seg[0], seg[1])
            cb, _ = pl.fetch_costs(b.n_slots)
            c = pl.cycle_scan((0.5, 0, 0.0), (0.0, 0.0, 0.1), ranges, angles, seg[0], seg[1])
            cc, _ = pl.fetch_costs(c.n_slots)
            d = pl.cycle_scan((0.5, 0, 0.0), (0.0, 0.0, 0.1), pr.array, pa.array, seg[0], seg[1])
            cd, _ = pl.fetch_costs(d.n_slots)
            assert (a.slot, a.n_admissible) == (b.slot, b.n_admissible) and np.array_equal(ca.view(np.uint32), cb.view(np.uint32))
            assert (c.slot, c.n_admissible) == (d.slot, d.n_admissible) and np.array_equal(cc.view(np.uint32), cd.view(np.uint32))
            assert np.array_equal(a.x, b.x) and np.array_equal(a.vx, b.vx) and np.float32(a.cost) == np.float32(b.cost)
            outs.append((a.slot, np.float32(a.cost), c.slot, np.float32(c.cost), ca.copy(), cc.copy()))
        pl.close()
    for o in outs[1:]:
        assert o[:4] == outs[0][:4]
        assert np.array_equal(o[4].view(np.uint32), outs[0][4].view(np.uint32))
        assert np.array_equal(o[5].view(np.uint32), outs[0][5].view(np.uint32))
    ref = run_oracle_cycle(kw, path, seg, (1.0, 0, 0.2), (0.0, 0.0, 0.0), cloud=cloud)
    assert outs[0][0] == ref["slot"] and outs[0][1] == np.float32(ref["cost"])
    for x in (pinned_cloud, pr, pa):
        x.free()


def test_programmatic_dependent_launches_do_not_change_the_cycle(pkg):
    """Tuning key 9: the kernels behind the bounds stage launched as programmatic dependents (their
    prologue overlaps the producer's tail) or as plain stream-ordered kernels, under graph replay and
    plain launches, branch and bound forced on: same winner, same cost bits, same per-slot record."""
    kw = wl.cfg_c2(n_lin=40, n_ang=40)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = np.ascontiguousarray(wl.cloud_bench(0)[::8])
    outs = []
    for pdl, graphs in ((1, 1), (0, 1), (1, 0), (0, 0)):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(7, 2)
        pl.set_tuning(9, pdl)
        pl.set_tuning(1, graphs)
        for _ in range(3):
            r = pl.cycle_cloud((1.0, 0, 0.2), (0.0, 0.0, 0.0), cloud, seg[0], seg[1])
        costs, adm = pl.fetch_costs(r.n_slots)
        outs.append((r.is_found, r.slot, np.float32(r.cost), r.n_admissible, costs.copy(), adm.copy(),
                     pl.fetch_pruned(r.n_slots).copy()))
        pl.close()
    for o in outs[1:]:
        assert o[:4] == outs[0][:4]
        assert np.array_equal(o[4].view(np.uint32), outs[0][4].view(np.uint32))
        assert np.array_equal(o[5], outs[0][5]) and np.array_equal(o[6], outs[0][6])
    assert outs[0][6].sum() > 0  # the bound stage did prune
    ref = run_oracle_cycle(kw, path, seg, (1.0, 0, 0.2), (0.0, 0.0, 0.0), cloud=cloud)
    assert outs[0][1] == ref["slot"] and outs[0][2] == np.float32(ref["cost"])


@pytest.mark.parametrize("mount", ["flip_x", "flip_y", "flip_x_yawed"])
@pytest.mark.parametrize("shape,dims", [(0, (0.25, 1.2, 0.0)), (1, (0.5, 0.3, 1.2)), (2, (0.45, 0.0, 0.0))])
def test_upside_down_sensor_mounts(pkg, mount, shape, dims):
    """A laser mounted upside down (180 deg about x or y, e.g. the Python-side quaternion [1, 0, 0, 0])
    mirrors the scan about the robot's axis: the octree's voxel cubes stay axis-aligned with the
    upright robot solid (reflection of the xy block, z flipped). Full cycle against the oracle."""
    s, c = math.sin(0.45), math.cos(0.45)
    rot = {"flip_x": (1.0, 0.0, 0.0, 0.0), "flip_y": (0.0, 1.0, 0.0, 0.0),
           # yaw 0.9 about z composed with the flip about x: q = q_z * q_x = (c, s, 0, 0)
           "flip_x_yawed": (c, s, 0.0, 0.0)}[mount]
    kw = wl.cfg_c1(weights=(1.0, 1.0, 1.0, 1.0, 1.0))
    kw.update(shape=shape, dims=dims, sensor_position=(0.15, -0.1, 0.25), sensor_rotation=rot,
              octree_resolution=0.07)
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    ranges, angles = wl.scan_360(13, lo=0.75, hi=3.0)
    got, ref = check_cycle(pkg, kw, path, seg, (0.3, 0, 0.2), (-0.4, 0.1, 0.7), scan=(ranges, angles))
    assert 0 < got.n_admissible < got.n_slots  # the mirrored obstacles block part of the fan


@pytest.mark.parametrize("tilt", ["pitch", "roll_yaw", "random"])
@pytest.mark.parametrize("shape,dims", [(0, (0.25, 0.8, 0.0)), (1, (0.5, 0.3, 0.7)), (2, (0.35, 0.0, 0.0))])
def test_tilted_sensor_mounts(pkg, tilt, shape, dims):
    """A pitched / rolled laser: the scan's points land in a tilted octree whose cubes are oriented boxes
    for the upright robot solid. Full cycle (rows, admissible set, every cost, winner) against the oracle."""
    q = {"pitch": (0.0, math.sin(0.15), 0.0, math.cos(0.15)),
         "roll_yaw": (math.sin(0.2) * math.cos(0.3), math.sin(0.2) * math.sin(0.3), math.cos(0.2) * math.sin(0.3),
                      math.cos(0.2) * math.cos(0.3)),
         "random": tuple(float(np.float32(v)) for v in (lambda a: a / np.linalg.norm(a))(np.random.default_rng(77).normal(size=4)))}[tilt]
    kw = wl.cfg_c1(weights=(1.0, 1.0, 1.0, 1.0, 1.0))
    kw.update(shape=shape, dims=dims, sensor_position=(0.1, -0.05, 0.3), sensor_rotation=q, octree_resolution=0.08)
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    ranges, angles = wl.scan_360(23, lo=0.6, hi=2.5)
    got, ref = check_cycle(pkg, kw, path, seg, (0.3, 0, 0.2), (-0.4, 0.1, 0.7), scan=(ranges, angles))
    assert 0 < got.n_admissible < got.n_slots


def test_moving_pose_and_sensor_offset(pkg):
    kw = wl.cfg_c2(n_lin=20, n_ang=20)
    kw.update(sensor_position=(0.2, 0.05, 0.3), sensor_rotation=(0.0, 0.0, math.sin(0.25), math.cos(0.25)))
    path = orc.Path(wl.circle34_points(10.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 40, 2.0)
    pose = (float(path.X[40]) + 0.1, float(path.Y[40]) - 0.05, 1.7)
    ranges, angles = wl.scan_360(5, n=720, lo=0.4, hi=6.0)
    check_cycle(pkg, kw, path, seg, (0.8, 0, -0.5), pose, scan=(ranges, angles))


@pytest.mark.parametrize("ctrl", [0, 2])
@pytest.mark.parametrize("shape,dims", [(1, (0.5, 0.3, 0.4)), (2, (0.25, 0.0, 0.0)), (0, (0.2, 0.5, 0.0))])
def test_c3_kinematics_and_shapes(pkg, ctrl, shape, dims):
    """Ackermann + omni kinematics, box / sphere / cylinder solids, P = 100, reduced sample grid"""
    kw = wl.cfg_c3(control_type=ctrl, n=24, shape=shape, dims=dims)
    path = orc.Path(wl.circle34_points(10.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    pose = (10.0, 0.0, math.pi / 2)
    cloud = wl.cloud_c2(7, n=8_000, center=(10.0, 0.0))
    got, ref = check_cycle(pkg, kw, path, seg, (0.5, 0.1 if ctrl == 2 else 0.0, 0.1), pose, cloud=cloud)
    assert got.n_points == 100


def test_keep_samples_padding(pkg):
    """drop_samples = false: colliding samples are truncated and zero-padded -> smoothness / jerk
    costs become non-zero (trajectory_sampler.cpp:157-168)"""
    kw = wl.cfg_c3(control_type=0, n=20, drop_samples=False, shape=0, dims=(0.1, 0.5, 0.0))
    kw["num_ctrl_points"] = 5
    path = orc.Path(wl.straight_points(30.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 4.0)
    cloud = wl.cloud_c2(11, n=6_000)
    got, ref = check_cycle(pkg, kw, path, seg, (1.5, 0, 0), (0, 0, 0), cloud=cloud)
    kw2 = dict(kw, drop_samples=True)
    ref2 = run_oracle_cycle(kw2, path, seg, (1.5, 0, 0), (0, 0, 0), cloud=cloud)
    assert ref["n_admissible"] > ref2["n_admissible"]  # padding really kicked in


def test_sampler_rows_bit_exact(pkg):
    """TrajectorySampler::generateTrajectories: every admissible row, order preserved"""
    kw = wl.cfg_c2(n_lin=30, n_ang=30)
    cloud = wl.cloud_c2(2, n=10_000)
    common = {k: v for k, v in kw.items() if k not in ("weights", "max_local_range")}
    scfg = orc.sampler_cfg(**common)
    ref = orc.sampler_generate(scfg, (1.2, 0, 0.3), (0.0, 0.0, 0.2), cloud=cloud)
    pl = make_planner(pkg, kw)
    got = pl.generate_trajectories((1.2, 0, 0.3), (0.0, 0.0, 0.2), cloud=cloud)
    pl.close()
    assert np.array_equal(got["slots"], ref["slots"])
    for k in ("vx", "vy", "omega", "x", "y"):
        assert np.array_equal(got[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_no_obstacles_and_all_blocked(pkg):
    kw = wl.cfg_c1(weights=(1.0, 1.0, 1.0, 1.0, 1.0))
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    pose = (-0.51731912, 0.0, 0.0)
    # empty scan: collision and obstacle terms vanish
    check_cycle(pkg, kw, path, seg, (0, 0, 0), pose, scan=(np.zeros(0), np.zeros(0)))
    # a wall of returns right on top of the robot: nothing admissible -> found = False, cost 0
    ang = np.linspace(0, 2 * math.pi, 720, endpoint=False)
    got, ref = check_cycle(pkg, kw, path, seg, (0, 0, 0), pose, scan=(np.full(720, 0.12), ang))
    assert not got.is_found and got.n_admissible == 0 and got.cost == 0.0


def test_non_finite_scan_ranges(pkg):
    kw = wl.cfg_c1(weights=(1.0, 1.0, 1.0, 1.0, 1.0))
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 1.0)
    ranges, angles = wl.scan_360(9)
    ranges[::7] = np.inf
    ranges[3::31] = np.nan
    check_cycle(pkg, kw, path, seg, (0.2, 0, 0), (-0.51731912, 0.0, 0.0), scan=(ranges, angles))


def test_adaptive_horizon_changes_points(pkg):
    kw = wl.cfg_c2(n_lin=20, n_ang=20)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    pl = make_planner(pkg, kw, path)
    assert pl.num_points == 50
    assert pl.set_prediction_horizon(0.5) == 25
    assert pl.set_prediction_horizon(0.0) == 2      # clamp to 2 time steps
    assert pl.set_prediction_horizon(9.0) == 50     # clamp to the base horizon
    n = pl.set_prediction_horizon(0.6)
    kw2 = dict(kw, prediction_horizon=0.6)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = wl.cloud_c2(4, n=5_000)
    ref = run_oracle_cycle(kw2, path, seg, (1.0, 0, 0), (0, 0, 0), cloud=cloud)
    got = pl.cycle_cloud((1.0, 0, 0), (0, 0, 0), cloud, seg[0], seg[1])
    assert got.n_points == n == ref["P"]
    assert_cycle_parity(got, ref)
    pl.close()


# ---- CostEvaluator API on caller-provided samples: the reference's own known-answer cases --------
def _solo(name):
    w = dict(path=0.0, goal=0.0, obstacles=0.0, smooth=0.0, jerk=0.0)
    w[name] = 1.0
    return (w["path"], w["goal"], w["obstacles"], w["smooth"], w["jerk"])


def _kat_planner(pkg, weights, path):
    kw = wl.cfg_c1(weights=weights)
    kw.update(vx=(1.0, 1.0, 1.0), vy=(1.0, 1.0, 1.0), omega=(1.0, 1.0, 1.0))  # dummyControlLimits
    return make_planner(pkg, kw, path)


def _sample(pts, vels=None):
    pts = np.asarray(pts, np.float32)
    n = len(pts)
    s = dict(x=pts[None, :, 0].copy(), y=pts[None, :, 1].copy(), vx=np.zeros((1, n - 1), np.float32),
             vy=np.zeros((1, n - 1), np.float32), omega=np.zeros((1, n - 1), np.float32))
    if vels is not None:
        v = np.asarray(vels, np.float64)
        s["vx"][0], s["vy"][0], s["omega"][0] = v[:, 0], v[:, 1], v[:, 2]
    return s


@pytest.mark.parametrize("name,sample,obst,expected,tol", [
    ("goal", _sample([(4.0, 0.0)] * 5), None, 0.6, 1e-4),
    ("goal", _sample([(4.0, 0.1)] * 5), None, 0.61, 1e-4),
    ("goal", _sample([(4.0, 0.5)] * 5), None, 0.65, 1e-4),
    ("path", _sample([(float(i), 0.0) for i in range(5)]), None, 0.0, 1e-4),
    ("path", _sample([(float(i), 0.5) for i in range(5)]), None, (0.5 + 0.5 / 4.0) / 2.0, 1e-4),
    ("smooth", _sample([(0, 0)] * 5, [(1.0, 0, 0)] * 4), None, 0.0, 1e-4),
    ("smooth", _sample([(0, 0)] * 5, [(0.0, 0, 0), (1.0, 0, 0), (1.0, 0, 0), (1.0, 0, 0)]), None, 1 / 12, 1e-4),
    ("jerk", _sample([(0, 0)] * 5, [(0.1, 0, 0), (0.2, 0, 0), (0.3, 0, 0), (0.4, 0, 0)]), None, 0.0, 1e-4),
    ("jerk", _sample([(0, 0)] * 5, [(0.0, 0, 0), (1.0, 0, 0), (3.0, 0, 0), (6.0, 0, 0)]), None, 2 / 12, 1e-4),
    ("obstacles", _sample([(0, 0)] * 5), (20.0, 0.0, 0.0), 0.0, 1e-4),
    ("obstacles", _sample([(0, 0)] * 5), (0.0, 0.0, 0.0), 1.0, 1e-4),
    ("obstacles", _sample([(0, 0)] * 5), (5.0, 0.0, 0.0), 0.5, 1e-4),
])
def test_cost_evaluator_known_answers(pkg, name, sample, obst, expected, tol):
    """ref: src/kompass_cpp/tests/cost_evaluator_test.cpp:217-461 through kc_cost_evaluate"""
    path = orc.Path([(0.0, 0.0), (10.0, 0.0)], 1.0, 5.0)
    pl = _kat_planner(pkg, _solo(name), path)
    if obst is not None:
        pl.set_point_scan((0, 0, 0), cloud=[obst], max_sensor_range=30.0, multiple=3.0)
    seg = path.segment(0)
    res, costs = pl.get_min_trajectory_cost(sample, seg[0], seg[1])
    pl.close()
    assert res.is_found
    assert abs(res.cost - expected) <= tol * max(abs(expected), 1.0)
    assert costs[0] == np.float32(res.cost)


def test_cost_evaluator_batch_matches_oracle(pkg):
    """fluctuating-velocity batch (benchmark generator shape) + custom cost addend + obstacles"""
    rng = np.random.default_rng(wl.SEED + 77)
    n, P = 300, 120
    t = np.arange(P) * 0.05
    x = (t[None, :] * rng.uniform(0.2, 1.5, (n, 1))).astype(np.float32)
    y = (np.sin(t[None, :] * rng.uniform(0.1, 2.0, (n, 1))) * rng.uniform(0, 1.0, (n, 1))).astype(np.float32)
    samples = dict(x=x, y=y, vx=rng.uniform(-1, 1, (n, P - 1)).astype(np.float32),
                   vy=rng.uniform(-0.2, 0.2, (n, P - 1)).astype(np.float32),
                   omega=rng.uniform(-2, 2, (n, P - 1)).astype(np.float32))
    custom = rng.uniform(0, 0.3, n).astype(np.float32)
    path = orc.Path([(0.0, 0.0), (5.0, 0.0), (10.0, 0.0)], 0.01, 1000.0, 1000)
    seg = path.segment(0)
    kw = wl.cfg_c1(weights=(1.0, 2.0, 0.5, 1.5, 0.7))
    kw.update(vx=(1.0, 3.0, 5.0), vy=(1.0, 3.0, 5.0), omega=(3.14, 3.0, 5.0))
    cloud = wl.cloud_c2(21, n=3_000)
    ccfg = orc.cost_cfg(w_path=1.0, w_goal=2.0, w_obstacles=0.5, w_smooth=1.5, w_jerk=0.7,
                        acc_limits=(3.0, 3.0, 3.0))
    obs = orc.cost_points(ccfg, (0.5, 0.2, 0.1), cloud=cloud)
    D = float(np.float32(12.0) / np.float32(3.0))
    found, idx, cost, costs = orc.cost_evaluate(ccfg, samples, path, seg, obs, D, custom=custom)
    pl = make_planner(pkg, kw, path)
    pl.set_point_scan((0.5, 0.2, 0.1), cloud=cloud, max_sensor_range=12.0, multiple=3.0)
    res, gcosts = pl.get_min_trajectory_cost(samples, seg[0], seg[1], custom=custom)
    pl.close()
    assert res.is_found == found and res.slot == idx
    assert np.array_equal(gcosts.view(np.uint32), costs.view(np.uint32))
    assert np.array_equal(res.x, x[idx])


def test_batch_sweep_matches_single_cycles(pkg):
    """kc_planner_batch_cloud == R independent kc_planner_cycle_cloud calls (config 5 shape)"""
    kw = wl.cfg_c2(n_lin=24, n_ang=24)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    R = 6
    rng = np.random.default_rng(wl.SEED + 5)
    clouds = [wl.cloud_c2(100 + r, n=4_000 + 500 * r) for r in range(R)]
    vels = np.stack([rng.uniform(0.0, 2.0, R), np.zeros(R), rng.uniform(-1, 1, R)], axis=1)
    poses = np.zeros((R, 3))
    poses[:, 2] = rng.uniform(-0.5, 0.5, R)
    pl = make_planner(pkg, kw, path)
    batch = pl.batch_cloud(vels, poses, clouds, seg[0], seg[1])
    for r in range(R):
        one = pl.cycle_cloud(vels[r], poses[r], clouds[r], seg[0], seg[1])
        assert batch[r][0] == one.is_found and batch[r][2] == one.slot and batch[r][3] == one.n_admissible
        assert np.float32(batch[r][1]) == np.float32(one.cost)
        ref = run_oracle_cycle(kw, path, seg, vels[r], poses[r], cloud=clouds[r])
        assert ref["slot"] == one.slot
    pl.close()


def test_batch_sweep_chunks_and_shared_clouds(pkg):
    """A sweep larger than one chunk (64 robots per launch set): workspace reuse across chunks, the
    flat-array form with robots SHARING cloud slices (out-of-order offsets), a pinned input buffer, and
    the resident replay all give what single cycles give."""
    kw = wl.cfg_c2(n_lin=12, n_ang=12)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    R, n_distinct = 150, 7
    rng = np.random.default_rng(wl.SEED + 15)
    base = [wl.cloud_c2(300 + k, n=1_500 + 211 * k) for k in range(n_distinct)]
    starts = np.concatenate([[0], np.cumsum([len(b) for b in base])])
    pick = rng.integers(0, n_distinct, R)
    offsets, counts = starts[pick].astype(np.int64), np.array([len(base[k]) for k in pick], np.int32)
    counts[17] = 0  # a robot without any obstacle point
    vels = np.stack([rng.uniform(0.0, 2.0, R), np.zeros(R), rng.uniform(-1, 1, R)], axis=1)
    poses = np.zeros((R, 3))
    poses[:, 2] = rng.uniform(-0.5, 0.5, R)
    flat = np.concatenate(base).astype(np.float32)
    pinned = pkg.PinnedArray(flat.shape, np.float32)
    pinned.array[...] = flat
    pl = make_planner(pkg, kw, path)
    batch = pl.batch_cloud(vels, poses, pinned.array, seg[0], seg[1], offsets=offsets, counts=counts)
    _, replayed = pl.batch_replay(2, R)
    assert replayed == batch
    pageable = pl.batch_cloud(vels, poses, flat, seg[0], seg[1], offsets=offsets, counts=counts)
    assert pageable == batch
    for r in list(range(0, R, 11)) + [17, 63, 64, 127, 128, R - 1]:
        cloud = flat[offsets[r]:offsets[r] + counts[r]]
        one = pl.cycle_cloud(vels[r], poses[r], cloud, seg[0], seg[1])
        assert batch[r][0] == one.is_found and batch[r][2] == one.slot and batch[r][3] == one.n_admissible, r
        assert np.float32(batch[r][1]) == np.float32(one.cost), r
    pl.close()
    pinned.free()


def test_error_codes(pkg):
    kw = wl.cfg_c1()
    pl = make_planner(pkg, kw)
    with pytest.raises(ValueError, match="global path"):
        pl.cycle_scan((0, 0, 0), (0, 0, 0), [1.0], [0.0], 0, 1)
    path = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    pl.set_path(path.X, path.Y, path.acc, path.total_length)
    with pytest.raises(IndexError, match="Invalid range for path part"):
        pl.cycle_scan((0, 0, 0), (0, 0, 0), [1.0], [0.0], path.n - 2, 10)
    pl.close()
    kw = dict(wl.cfg_c1(), sensor_rotation=(0.3, 0.0, 0.0, 0.95))  # not a rotation (|q| = 0.996): loud, not silent
    pl = make_planner(pkg, kw, path)
    with pytest.raises(pkg.KompassB200Error, match="unit quaternion"):
        pl.cycle_scan((0, 0, 0), (0, 0, 0), [1.0], [0.0], 0, 10)
    pl.close()


# ---------------------------------------------------------------------------------------------
# Full-size property check (BASELINE configs 2 and 3): the pruned nearest-obstacle search of the
# cycle against the on-device brute force over all N*P*M pairs (kc_planner_bruteforce_obstacle_costs:
# the reference's minDist2D loop as written). With only the obstacle weight set, a slot's total IS
# its obstacle cost (float(0 + 1.0 * c)), so the two must agree bit for bit on every admissible slot.
# The brute force itself is pinned to the CPU oracle at a size the oracle finishes in seconds.
# ---------------------------------------------------------------------------------------------
def _obstacle_only_cycle(pkg, kw, path, seg, vel, pose, scan=None, cloud=None):
    kw = dict(kw)
    kw["weights"] = (0.0, 0.0, 1.0, 0.0, 0.0)
    # the branch and bound must pick the same winner; then every slot evaluated exactly
    pl = make_planner(pkg, kw, path)
    pl.set_tuning(7, 2)
    pruned = (pl.cycle_scan(vel, pose, scan[0], scan[1], seg[0], seg[1]) if scan is not None
              else pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1]))
    pcosts, padm = pl.fetch_costs(pruned.n_slots)
    pprn = pl.fetch_pruned(pruned.n_slots)
    pl.close()
    pl = make_planner(pkg, kw, path)
    pl.set_tuning(7, 0)
    if scan is not None:
        got = pl.cycle_scan(vel, pose, scan[0], scan[1], seg[0], seg[1])
    else:
        got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
    costs, adm = pl.fetch_costs(got.n_slots)
    brute, ms, pairs = pl.bruteforce_obstacle_costs(got.n_slots)
    pl.close()
    assert (pruned.slot, np.float32(pruned.cost), pruned.n_admissible) == (got.slot, np.float32(got.cost), got.n_admissible)
    assert np.array_equal(pruned.x, got.x) and np.array_equal(padm, adm)
    keep = (adm == 1) & (pprn == 0)
    assert np.array_equal(pcosts[keep].view(np.uint32), costs[keep].view(np.uint32))
    lost = (adm == 1) & (pprn == 1)
    assert np.all(pcosts[lost] <= costs[lost]) and np.all(costs[lost] > np.float32(got.cost))
    print(f"branch and bound: {int(lost.sum())} of {int((adm == 1).sum())} admissible slots pruned")
    return kw, got, costs, adm, brute, ms, pairs


def test_bruteforce_hook_matches_oracle(pkg):
    kw = wl.cfg_c2(n_lin=30, n_ang=30)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud = wl.cloud_c2(4, n=4_099)  # two full shared-memory tiles and a ragged one
    kw, got, costs, adm, brute, ms, pairs = _obstacle_only_cycle(pkg, kw, path, seg, (1.0, 0, 0.0), (0, 0, 0), cloud=cloud)
    ref = run_oracle_cycle(kw, path, seg, (1.0, 0, 0.0), (0, 0, 0), cloud=cloud)
    slots = ref["samples"]["slots"]
    assert pairs == len(slots) * 50 * 4_099
    assert np.array_equal(brute[slots].view(np.uint32), ref["costs"].view(np.uint32))
    assert np.all(brute[adm == 0] == np.finfo(np.float32).max)
    assert np.array_equal(brute.view(np.uint32), costs.view(np.uint32))
    # laser scan with inf / NaN ranges (kept for the cost, q8) and a pose + sensor offset
    kw1 = wl.cfg_c1()
    kw1.update(sensor_position=(0.2, -0.1, 0.3), sensor_rotation=(0, 0, math.sin(0.35), math.cos(0.35)))
    p1 = orc.Path(wl.GLOBAL_PATH_XY, 0.01, 1.0)
    s1 = wl.tracked_segment(p1, 0, 1.0)
    ranges, angles = wl.scan_360(11)
    ranges[::9] = np.inf
    ranges[4::23] = np.nan
    pose = (-0.4, 0.15, 0.6)
    kw1, got, costs, adm, brute, ms, pairs = _obstacle_only_cycle(pkg, kw1, p1, s1, (0.2, 0, 0.1), pose, scan=(ranges, angles))
    ref = run_oracle_cycle(kw1, p1, s1, (0.2, 0, 0.1), pose, scan=(ranges, angles))
    assert np.array_equal(brute[ref["samples"]["slots"]].view(np.uint32), ref["costs"].view(np.uint32))
    assert np.array_equal(brute.view(np.uint32), costs.view(np.uint32))


def test_bruteforce_packed_and_scalar_forms_agree(pkg, monkeypatch):
    """The packed FP32 form (FADD2 / FMUL2 / FFMA2, points paired in memory, odd count padded) and the
    scalar form (KC_BF_SCALAR=1, read at every call) of the brute-force hook return the same bits."""
    kw = wl.cfg_c2(n_lin=24, n_ang=24)
    kw["weights"] = (0.0, 0.0, 1.0, 0.0, 0.0)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    for n in (1, 2, 2_047, 2_048, 6_145):  # single point, one pair, ragged / full / odd multi-tile
        cloud = wl.cloud_c2(9, n=n)
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(7, 0)
        got = pl.cycle_cloud((1.0, 0, 0.0), (0, 0, 0), cloud, seg[0], seg[1])
        costs, adm = pl.fetch_costs(got.n_slots)
        monkeypatch.delenv("KC_BF_SCALAR", raising=False)
        packed = pl.bruteforce_obstacle_costs(got.n_slots)[0].copy()
        monkeypatch.setenv("KC_BF_SCALAR", "1")
        scalar = pl.bruteforce_obstacle_costs(got.n_slots)[0].copy()
        monkeypatch.delenv("KC_BF_SCALAR", raising=False)
        pl.close()
        assert np.array_equal(packed.view(np.uint32), scalar.view(np.uint32)), n
        assert np.array_equal(packed.view(np.uint32), costs.view(np.uint32)), n


@pytest.mark.parametrize("config", ["c2", "c3_ackermann_box", "c3_omni_keep"])
def test_full_size_pruned_search_equals_bruteforce(pkg, config):
    if config == "c2":  # 10 201 slots x 50 points vs 100 000 points
        kw, vel = wl.cfg_c2(), (1.0, 0.0, 0.0)
        path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
        seg = wl.tracked_segment(path, 0, 2.0)
    else:  # ~50 k slots x 100 points vs 100 000 points
        kw = wl.cfg_c3(control_type=0 if "ackermann" in config else 2, drop_samples="keep" not in config)
        vel = (1.0, 0.0, 0.0)
        path = orc.Path(wl.circle34_points(), 0.01, 1.0)
        seg = wl.tracked_segment(path, 0, 4.0)
    pose = (float(path.X[0]), float(path.Y[0]), 0.0) if config != "c2" else (0.0, 0.0, 0.0)
    cloud = wl.cloud_bench(5, n=100_000, center=pose[:2])
    kw, got, costs, adm, brute, ms, pairs = _obstacle_only_cycle(pkg, kw, path, seg, vel, pose, cloud=cloud)
    n_adm = int(adm.sum())
    assert got.is_found and n_adm == got.n_admissible and n_adm > 5000
    assert pairs == float(n_adm) * kw["prediction_horizon"] / kw["time_step"] * 100_000
    same = brute.view(np.uint32) == costs.view(np.uint32)
    assert same.all(), f"{(~same).sum()} of {len(same)} slots differ"
    c = costs[adm == 1]
    assert (c > 0).sum() > 100 and len(np.unique(c)) > 1000
    # the winner is the lowest-index minimum of those costs (cost_evaluator.cpp:102)
    assert got.slot == int(np.flatnonzero(adm == 1)[np.argmin(c)]) and np.float32(got.cost) == c.min()
    print(f"{config}: {pairs:.3g} pairs in {ms:.2f} ms (FP32 pass) = {pairs * 6 / ms / 1e9:.1f} TFLOP/s algorithmic")


_THREAD_SCRIPT = '''
import sys, threading
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import numpy as np
import __graft_entry__ as ge, orc, workloads as wl
from parity_util import make_planner, run_oracle_cycle
pkg = ge.load_package()
kw = wl.cfg_c2(n_lin=24, n_ang=24)
path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
seg = wl.tracked_segment(path, 0, 2.0)
cloud = wl.cloud_c2(3, n=6000)
pl = make_planner(pkg, kw, path)          # handle created on the main thread
mp = pkg.LocalMapperGPU(60, 60, 0.1, (0.0, 0.0, 0.0), 0.0, False, 90, 0.07, 2.0, 0.0, 20.0)
out = {}
def work():                                # ... and driven from another one
    out["a"] = pl.cycle_cloud((1.0, 0, 0.1), (0, 0, 0), cloud, seg[0], seg[1])
    ang = np.linspace(-3, 3, 90); out["g"] = mp.scan_to_grid(ang, np.full(90, 2.0))
t = threading.Thread(target=work); t.start(); t.join()
ref = run_oracle_cycle(kw, path, seg, (1.0, 0, 0.1), (0, 0, 0), cloud=cloud)
assert out["a"].slot == ref["slot"] and np.float32(out["a"].cost) == np.float32(ref["cost"]), (out["a"].slot, ref["slot"])
assert (out["g"] == 100).sum() > 0
b = pl.cycle_cloud((1.0, 0, 0.1), (0, 0, 0), cloud, seg[0], seg[1])   # back on the main thread
assert b.slot == ref["slot"]
print("ok", pkg.get_available_accelerators())
'''


def test_handles_can_be_driven_from_another_thread(pkg, tmp_path):
    """include/kompass_b200.h: distinct handles may be driven from distinct threads. The CUDA current
    device is per-thread state, so every entry point re-asserts the process's device: a
    handle created on the main thread runs its cycle on a worker thread - on the LAST device of the
    box (KOMPASS_B200_DEVICE), which is not the default device 0 wherever there are two or more."""
    import os
    import subprocess
    import sys
    import torch
    n_dev = torch.cuda.device_count()
    script = tmp_path / "threads.py"
    script.write_text(_THREAD_SCRIPT)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, KOMPASS_B200_DEVICE=str(max(0, int(n_dev) - 1)))
    env.pop("LOCAL_RANK", None)
    r = subprocess.run([sys.executable, str(script), root], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("family", ["clutter_in_reach", "pillars_in_reach", "dense_cluster_on_path", "all_ties_far_obstacles"])
def test_exact_stage_by_point_or_by_slot(pkg, family):
    """Tuning key 13: number of survivors of the bound stage up to which k_cost_eval spreads (slot, point)
    pairs over the grid in batches (per-lane short lists, whole-warp searches for the rest) instead of
    handing whole slots to warps. Never (0), small batches only (32), the default (2048), always: the same
    winner record, the same admissible set, and every surviving slot's cost keeps its bits; the oracle
    agrees with all of them."""
    gen, w = wl.CLOUD_FAMILY[family]
    kw = wl.cfg_c2(n_lin=60, n_ang=60) if w is None else wl.cfg_c2(n_lin=60, n_ang=60, weights=w)
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    cloud, _ = wl.family_cloud(family, 23, n=60_000)
    vel, pose = (1.0, 0.0, -0.2), (0.0, 0.0, 0.0)
    ref = run_oracle_cycle(kw, path, seg, vel, pose, cloud=cloud)
    outs = []
    for limit in (0, 32, 2048, 1 << 30):
        pl = make_planner(pkg, kw, path)
        pl.set_tuning(13, limit)
        for rep in range(2):  # (the second cycle runs with the heavy-cell feedback of the first)
            got = pl.cycle_cloud(vel, pose, cloud, seg[0], seg[1])
        costs, adm = pl.fetch_costs(got.n_slots)
        prn = pl.fetch_pruned(got.n_slots)
        outs.append((got.slot, np.float32(got.cost), got.n_admissible, costs.copy(), adm.copy(), prn.copy(), limit))
        assert got.is_found == ref["found"] and got.slot == ref["slot"], (limit, got.slot, ref["slot"])
        assert np.float32(got.cost) == np.float32(ref["cost"]), (limit, got.cost, ref["cost"])
        assert got.n_admissible == ref["n_admissible"]
        pl.close()
    for o in outs[1:]:
        assert o[:3] == outs[0][:3], o[6]
        assert np.array_equal(o[4], outs[0][4]) and np.array_equal(o[5], outs[0][5]), o[6]
        live = (o[4] == 1) & (o[5] == 0)  # admissible and not pruned: exact totals
        assert np.array_equal(o[3][live].view(np.uint32), outs[0][3][live].view(np.uint32)), o[6]
