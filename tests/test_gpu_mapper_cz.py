"""GPU parity tests: local mapper (grid cells bit-exact), point-cloud binning (ranges bit-exact) and
critical zone (factor bit-exact) vs the CPU oracle, plus the invariants the reference tests assert."""
import math
import struct

import numpy as np
import pytest

import orc
import workloads as wl

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ mapper
def _mapper(pkg, H=400, W=400, res=0.05, pos=(0.0, 0.0, 0.0), orient=0.0, cloud=False, scan_size=1080,
            max_h=2.0, min_h=0.1, range_max=20.0):
    return pkg.LocalMapperGPU(H, W, res, pos, orient, cloud, scan_size, 0.01, max_h, min_h, range_max, 256)


@pytest.mark.parametrize("variant", ["sine", "random", "offset"])
def test_scan_to_grid_cells_bit_exact(pkg, variant):
    H = W = 400
    res, pos, orient = 0.05, (0.0, 0.0, 0.0), 0.0
    if variant == "sine":
        angles, ranges = wl.mapping_scan(1080)
    elif variant == "random":
        angles, _ = wl.mapping_scan(1080)
        ranges = np.random.default_rng(wl.SEED).uniform(0.1, 14.1, 1080)
    else:
        H, W, res, pos, orient = 300, 500, 0.04, (1.3, -2.1, 0.0), 0.7
        angles, ranges = wl.mapping_scan(2000)
    m = _mapper(pkg, H, W, res, pos, orient)
    got = m.scan_to_grid(angles, ranges)
    m.close()
    ref = orc.mapper_scan_to_grid(H, W, res, pos, orient, angles, ranges)
    assert got.shape == (H, W) and got.dtype == np.int32
    assert np.array_equal(got, ref), f"{(got != ref).sum()} cells differ"
    # invariants the reference tests assert (tests/test_local_mapper_bindings.py:78-297)
    vals = set(np.unique(got).tolist())
    assert vals <= {-1, 0, 100}
    assert (got == 100).sum() > 0 and (got == 0).sum() > 0
    assert (got == -1).sum() + (got == 0).sum() + (got == 100).sum() == H * W


def test_scan_to_grid_edge_cases(pkg):
    m = _mapper(pkg, 100, 100, 0.1)
    empty = m.scan_to_grid(np.zeros(0), np.zeros(0))
    assert (empty == -1).all()
    angles = np.linspace(-math.pi, math.pi, 90, endpoint=False)
    far = m.scan_to_grid(angles, np.full(90, 500.0))  # every hit outside the grid
    ref = orc.mapper_scan_to_grid(100, 100, 0.1, (0, 0, 0), 0.0, angles, np.full(90, 500.0))
    assert np.array_equal(far, ref)
    assert (far == 100).sum() == 0 and (far == 0).sum() > 0
    zero = m.scan_to_grid(angles, np.zeros(90))  # zero range: start cell is the hit cell
    refz = orc.mapper_scan_to_grid(100, 100, 0.1, (0, 0, 0), 0.0, angles, np.zeros(90))
    assert np.array_equal(zero, refz)
    m.close()


def test_cloud_binning_bit_exact(pkg):
    pts = wl.cloud_lattice(0, 100_000)
    data = wl.cloud_bytes_xyz16(pts)
    n = len(pts)
    for bins, min_z, max_z in [(1080, 0.1, 2.0), (360, 0.0, -1.0), (3600, 0.5, 1.0)]:
        got = pkg.pointcloud_to_laserscan(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, min_z, max_z, bins)
        ref = orc.pointcloud_to_laserscan(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, min_z, max_z, bins)
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (bins, (got != ref).sum())


def test_cloud_binning_unaligned_layout_and_ring(pkg):
    # 3 rows x 500 points, 20-byte points with fields at odd offsets, trailing row padding
    rng = np.random.default_rng(wl.SEED + 3)
    width, height, ps, xo, yo, zo = 500, 3, 21, 1, 9, 13
    rs = width * ps + 11
    buf = np.zeros(height * rs, np.uint8)
    th = rng.uniform(0, 2 * math.pi, (height, width))
    for r in range(height):
        for c in range(width):
            o = r * rs + c * ps
            buf[o + xo:o + xo + 4] = np.frombuffer(struct.pack("<f", math.cos(th[r, c])), np.uint8)
            buf[o + yo:o + yo + 4] = np.frombuffer(struct.pack("<f", math.sin(th[r, c])), np.uint8)
            buf[o + zo:o + zo + 4] = np.frombuffer(struct.pack("<f", 0.5), np.uint8)
    data = buf.view(np.int8)
    got = pkg.pointcloud_to_laserscan(data, ps, rs, height, width, xo, yo, zo, 10.0, 0.0, 2.0, 360)
    ref = orc.pointcloud_to_laserscan(data, ps, rs, height, width, xo, yo, zo, 10.0, 0.0, 2.0, 360)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    hit = got < 10.0
    assert hit.mean() > 0.4 and np.all(np.abs(got[hit] - 1.0) < 1e-3)  # test_pointcloud_data.py:154-256
    # origin / z filters: everything stays at max_range
    got2 = pkg.pointcloud_to_laserscan(data, ps, rs, height, width, xo, yo, zo, 10.0, 0.6, 2.0, 360)
    assert np.all(got2 == 10.0)


def test_cloud_to_grid_bit_exact(pkg):
    pts = wl.cloud_lattice(1, 100_000)
    data = wl.cloud_bytes_xyz16(pts)
    n = len(pts)
    m = _mapper(pkg, cloud=True, scan_size=1080)
    got = m.scan_to_grid(data, 16, n * 16, 1, n, 0, 4, 8)
    ranges = orc.pointcloud_to_laserscan(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, 0.1, 2.0, 1080)
    angles = np.array([i * (2.0 * math.pi) / 1080 for i in range(1080)])
    ref = orc.mapper_scan_to_grid(400, 400, 0.05, (0, 0, 0), 0.0, angles, ranges)
    assert np.array_equal(got, ref), f"{(got != ref).sum()} cells differ"
    empty = m.scan_to_grid(np.zeros(0, np.int8), 16, 0, 1, 0, 0, 4, 8)  # empty cloud: rays at range_max
    refe = orc.mapper_scan_to_grid(400, 400, 0.05, (0, 0, 0), 0.0, angles, np.full(1080, 20.0))
    assert np.array_equal(empty, refe)
    m.close()


# ------------------------------------------------------------------ critical zone
def _init_scan(n, r):
    return np.full(n, r, np.float64), np.array([2.0 * math.pi * i / n for i in range(n)], np.float64)


def _set(angle, value, ranges, angles):
    a = math.fmod(angle, 2 * math.pi)
    if a < 0:
        a += 2 * math.pi
    ranges[int(np.argmin(np.abs(angles - a)))] = value


def _cz(pkg, angles, input_type=0, pos=(0.22, 0.0, 0.4), rot=(0, 0, 0.99, 0.0)):
    return pkg.CriticalZoneCheckerGPU(input_type, 0, (0.51, 2.0), pos, rot, 160.0, 0.3, 0.6, angles, 0.1,
                                      2.0, 20.0)


def test_critical_zone_laserscan_reference_cases(pkg):
    """ref: src/kompass_cpp/tests/critical_zone_test.cpp:39-190 (tests 1-8)"""
    ranges, angles = _init_scan(360, 10.0)
    z = _cz(pkg, angles)
    cfg = orc.cz_cfg()

    def chk(fwd):
        g = z.check(ranges, fwd)
        assert g == orc.cz_check_scan(cfg, angles, ranges, fwd)
        return g

    for a in (0.0, 0.1, -0.1):
        _set(a, 0.2, ranges, angles)
    assert chk(True) == 1.0
    ranges, _ = _init_scan(360, 10.0)
    assert chk(True) == 1.0
    for a in (math.pi, math.pi + 0.1, math.pi - 0.1):
        _set(a, 0.2, ranges, angles)
    assert chk(True) == 0.0
    assert chk(False) == 1.0
    for a in (0.0, 0.1, -0.1):
        _set(a, 0.2, ranges, angles)
    assert chk(False) == 0.0
    ranges, _ = _init_scan(360, 10.0)
    _set(0.0, 1.3, ranges, angles)
    assert 0.0 < chk(False) < 1.0
    assert chk(True) == 1.0
    _set(math.pi, 0.7, ranges, angles)
    assert 0.0 < chk(True) < 1.0
    z.close()


def _bytes(pts):
    b = b"".join(struct.pack("<ffff", x, y, zz, 0.0) for (x, y, zz) in pts)
    return np.frombuffer(b, dtype=np.int8) if b else np.zeros(0, np.int8)


def test_critical_zone_pointcloud_reference_cases(pkg):
    """ref: critical_zone_test.cpp:196-333 (tests 9-14)"""
    _, angles = _init_scan(360, 10.0)
    z = _cz(pkg, angles, 1, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0))
    cfg = orc.cz_cfg(sensor_position=(0, 0, 0), sensor_rotation=(0, 0, 0, 1))

    def run(pts, fwd):
        d = _bytes(pts)
        n = len(pts)
        g = z.check(d, 16, n * 16, 1, n, 0, 4, 8, fwd)
        assert g == orc.cz_check_cloud(cfg, angles, d, 16, n * 16, 1, n, 0, 4, 8, fwd)
        return g

    assert run([], True) == 1.0
    assert run([(0.7, 0.0, 0.5)], True) == 0.0
    assert run([(0.7, 0.0, 3.0)], True) == 1.0
    assert 0.4 < run([(0.95, 0.0, 0.5)], True) < 0.6
    assert run([(0.95, 0, 0.5), (1, 1, 0.5), (-1, -1, 0.5), (-0.1, -0.1, 3.0), (-0.1, -0.1, -3.0),
                (0.1, 0.2, 4.0), (0.1, 0.2, -4.0), (0.75, 0.0, 0.5)], True) == 0.0
    assert 0.4 < run([(0.95, 0, 0.5), (-0.95, 0, 0.5), (1, 1, 0.5), (-1, -1, 0.5), (-0.1, -0.1, 3.0),
                      (-0.1, -0.1, -3.0), (0.1, 0.2, 4.0), (0.1, 0.2, -4.0)], False) < 0.6
    z.close()


def test_critical_zone_benchmark_shapes(pkg):
    """config 4: 100k-point cloud and the 3600-ray all-slowdown scan, factor bit-exact"""
    angles360 = np.array([i * 2.0 * math.pi / 360.0 for i in range(360)])
    z = _cz(pkg, angles360, 1)
    cfg = orc.cz_cfg()
    for seed in range(3):
        pts = wl.cloud_lattice(seed, 100_000)
        d = wl.cloud_bytes_xyz16(pts)
        n = len(pts)
        for fwd in (True, False):
            assert z.check(d, 16, n * 16, 1, n, 0, 4, 8, fwd) == \
                orc.cz_check_cloud(cfg, angles360, d, 16, n * 16, 1, n, 0, 4, 8, fwd)
    # a sparse cloud that leaves a slowdown-only answer
    rng = np.random.default_rng(wl.SEED + 9)
    sp = np.zeros((2000, 4), np.float32)
    rr = rng.uniform(0.95, 6.0, 2000)
    aa = rng.uniform(0, 2 * math.pi, 2000)
    sp[:, 0], sp[:, 1], sp[:, 2] = rr * np.cos(aa), rr * np.sin(aa), 0.5
    d = wl.cloud_bytes_xyz16(sp)
    for fwd in (True, False):
        g = z.check(d, 16, 2000 * 16, 1, 2000, 0, 4, 8, fwd)
        assert g == orc.cz_check_cloud(cfg, angles360, d, 16, 2000 * 16, 1, 2000, 0, 4, 8, fwd)
    z.close()
    angles, ranges = wl.dense_slowdown_scan(3600)
    z = _cz(pkg, angles, 0)
    for fwd in (True, False):
        g = z.check(ranges, fwd)
        assert g == orc.cz_check_scan(cfg, angles, ranges, fwd)
    z.close()


# ------------------------------------------------------------------------------------------------
# Bayesian mapper + previous-grid warp
# ------------------------------------------------------------------------------------------------
def _circle_scan(radius, inc=0.01):  # ref: tests/mapper_test.cpp generateLaserScan("circle")
    angles = np.arange(0.0, 2 * math.pi, inc)
    return angles, np.full_like(angles, radius)


@pytest.mark.parametrize("variant", ["circles", "sine", "random", "offset_custom_params"])
def test_bayesian_grid_and_probabilities_bit_exact(pkg, variant):
    H, W, res, pos, orient = 60, 80, 0.05, (0.0, 0.0, 0.0), 0.0
    kwargs = dict(range_max=20.0)
    if variant == "circles":
        scans = [_circle_scan(r) for r in (0.3, 0.5, 2.0)]  # mapper_test.cpp:137-215
    elif variant == "sine":
        H, W = 400, 400
        scans = [wl.mapping_scan(1080)]
    elif variant == "random":
        H, W = 200, 160
        rng = np.random.default_rng(wl.SEED + 77)
        scans = [(np.array([-math.pi + 2 * math.pi * i / 720 for i in range(720)]), rng.uniform(0.1, 9.0, 720))]
    else:
        H, W, pos, orient = 120, 100, (0.4, -0.25, 0.1), 0.7
        kwargs = dict(p_prior=0.6, p_occupied=0.9, p_empty=0.1, range_sure=0.1, range_max=20.0, wall_size=0.2)
        scans = [wl.mapping_scan(900)]  # benchmark_runner.cpp:208-213 parameters
    mp = _mapper(pkg, H=H, W=W, res=res, pos=pos, orient=orient, range_max=kwargs["range_max"])
    if variant == "offset_custom_params":
        mp.set_bayesian_params(0.6, 0.9, 0.1, 0.1, 0.2)
    for angles, ranges in scans:
        g_ref, p_ref = orc.mapper_scan_to_grid_bayes(H, W, res, pos, orient, angles, ranges, **kwargs)
        g, p = mp.scan_to_grid_baysian(angles, ranges)
        assert np.array_equal(g, g_ref)
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32)), np.abs(p - p_ref).max()
        # the invariants the reference tests assert / log
        assert set(np.unique(g)) <= {-1, 0, 100}
        assert (g == 100).sum() > 0 and (g == 0).sum() > 0
        prior = np.float32(kwargs.get("p_prior", 0.5))
        assert np.all(p[g == -1] == prior)  # untouched cells keep the prior
    mp.close()


def test_previous_grid_warp_and_feedback_bit_exact(pkg):
    H, W, res = 90, 70, 0.1
    mp = _mapper(pkg, H=H, W=W, res=res)
    angles, ranges = wl.mapping_scan(720)
    prior = np.full((H, W), 0.5, np.float32)
    assert np.array_equal(mp.get_previous_grid(), prior)
    g, p = mp.scan_to_grid_baysian(angles, ranges * 0.6)
    # feed the posterior back, move the robot, warp, update again: every stage against the oracle
    mp.set_previous_grid(p)
    prev_ref = p.copy()
    for pos, yaw in [((0.35, -0.2), 0.3), ((-0.15, 0.4), -1.1), ((0.0, 0.0), 0.0), ((2.0, 1.0), 3.0)]:
        mp.get_previous_grid_in_current_pose(pos, yaw)
        prev_ref = orc.mapper_warp_previous(H, W, res, 0.5, pos, yaw, prev_ref)
        got = mp.get_previous_grid()
        assert np.array_equal(got.view(np.uint32), prev_ref.view(np.uint32)), np.abs(got - prev_ref).max()
        g_ref, p_ref = orc.mapper_scan_to_grid_bayes(H, W, res, (0, 0, 0), 0.0, angles, ranges * 0.6, prev=prev_ref)
        g, p = mp.scan_to_grid_baysian(angles, ranges * 0.6)
        assert np.array_equal(g, g_ref)
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
    mp.close()


def test_bayesian_edge_cases(pkg):
    mp = _mapper(pkg, H=50, W=50, res=0.1)
    g, p = mp.scan_to_grid_baysian(np.zeros(0), np.zeros(0))
    assert np.all(g == -1) and np.all(p == np.float32(0.5))
    # rays leaving the grid: no endpoint inside, probabilities only along the crossed cells
    angles, ranges = _circle_scan(30.0, 0.05)
    g_ref, p_ref = orc.mapper_scan_to_grid_bayes(50, 50, 0.1, (0, 0, 0), 0.0, angles, ranges)
    g, p = mp.scan_to_grid_baysian(angles, ranges)
    assert np.array_equal(g, g_ref) and np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
    assert (g == 100).sum() == 0
    with pytest.raises((IndexError, ValueError)):
        mp.set_bayesian_params(1.5, 0.6, 0.4, 1.0, 0.2)
    mp.close()


# ------------------------------------------------------------------------------------------------
# angle-step binning (pointcloud.h:116-177) and the Bayesian mapper's raw-cloud overload
# (local_mapper.cpp:253-264)
# ------------------------------------------------------------------------------------------------
def test_cloud_binning_angle_step_bit_exact(pkg):
    pts = wl.cloud_lattice(2, 100_000)
    data = wl.cloud_bytes_xyz16(pts)
    n = len(pts)
    for step, min_z, max_z in [(0.01, 0.1, 2.0), (float(np.float32(0.01)), 0.0, -1.0), (0.3, 0.5, 1.0),
                               (2 * math.pi / 360, 0.1, 2.0)]:
        got_r, got_a = pkg.pointcloud_to_laserscan_step(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, min_z, max_z, step)
        ref_r, ref_a = orc.pointcloud_to_laserscan_step(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, min_z, max_z, step)
        assert len(got_r) == len(ref_r) == math.ceil(2 * math.pi / step)
        assert np.array_equal(got_a.view(np.uint64), ref_a.view(np.uint64))
        assert np.array_equal(got_r.view(np.uint64), ref_r.view(np.uint64)), (step, (got_r != ref_r).sum())
    with pytest.raises((IndexError, ValueError)):
        pkg.pointcloud_to_laserscan_step(data, 16, n * 16, 1, n, 0, 4, 8, 20.0, 0.0, 2.0, 0.0)


@pytest.mark.parametrize("layout", ["xyz16", "unaligned"])
def test_bayesian_cloud_overload_bit_exact(pkg, layout):
    H, W, res = 400, 400, 0.05
    pts = wl.cloud_lattice(3, 60_000)
    if layout == "xyz16":
        data, ps, rs, h, w, xo, yo, zo = wl.cloud_bytes_xyz16(pts), 16, len(pts) * 16, 1, len(pts), 0, 4, 8
    else:  # 2 rows, 19-byte points, fields at odd offsets, padded rows
        ps, xo, yo, zo, h = 19, 3, 7, 11, 2
        w = len(pts) // h
        rs = w * ps + 5
        buf = np.zeros(h * rs, np.uint8)
        raw = np.ascontiguousarray(pts[:, :3]).view(np.uint8).reshape(len(pts), 12)
        for r in range(h):
            rows = raw[r * w:(r + 1) * w]
            base = r * rs + np.arange(w) * ps
            for k, off in enumerate((xo, yo, zo)):
                for b in range(4):
                    buf[base + off + b] = rows[:, 4 * k + b]
        data = buf.view(np.int8)
    mp = _mapper(pkg, H=H, W=W, res=res, cloud=True, scan_size=1080)  # ctor angle_step = 0.01f
    step = float(np.float32(0.01))
    ranges, angles = orc.pointcloud_to_laserscan_step(data, ps, rs, h, w, xo, yo, zo, 20.0, 0.1, 2.0, step)
    assert len(ranges) == 629
    prev = None
    for _ in range(2):  # second pass feeds the posterior back as the previous grid
        g_ref, p_ref = orc.mapper_scan_to_grid_bayes(H, W, res, (0, 0, 0), 0.0, angles, ranges, prev=prev)
        g, p = mp.scan_to_grid_baysian(data, ps, rs, h, w, xo, yo, zo)
        assert np.array_equal(g, g_ref), f"{(g != g_ref).sum()} cells differ"
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32)), np.abs(p - p_ref).max()
        assert (g == 100).sum() > 0 and (g == 0).sum() > 0
        mp.set_previous_grid(p)
        prev = p_ref
    # empty cloud: every ray runs to range_max
    g, p = mp.scan_to_grid_baysian(np.zeros(0, np.int8), 16, 0, 1, 0, 0, 4, 8)
    g_ref, p_ref = orc.mapper_scan_to_grid_bayes(H, W, res, (0, 0, 0), 0.0, angles, np.full(629, 20.0), prev=prev)
    assert np.array_equal(g, g_ref) and np.array_equal(p.view(np.uint32), p_ref.view(np.uint32))
    mp.close()


def test_page_locked_clouds_are_read_in_place(pkg):
    """A raw cloud in page-locked memory (kc_pinned_alloc) is consumed in place by the binning kernel;
    pageable memory is staged in chunks. Same grid / factor either way; a replay needs a resident copy."""
    pts = wl.cloud_lattice(6, 70_001)  # odd count: ragged last staging chunk
    data = wl.cloud_bytes_xyz16(pts)
    n = len(pts)
    pinned = pkg.PinnedArray(data.shape, np.int8)
    pinned.array[...] = data
    m = _mapper(pkg, cloud=True, scan_size=1080)
    g_page = m.scan_to_grid(data, 16, n * 16, 1, n, 0, 4, 8)
    assert m.replay(3) > 0.0
    g_pin = m.scan_to_grid(pinned.array, 16, n * 16, 1, n, 0, 4, 8)
    assert np.array_equal(g_page, g_pin) and (g_pin == 100).sum() > 0
    with pytest.raises(ValueError, match="nothing resident"):
        m.replay(1)
    gb_page, pb_page = m.scan_to_grid_baysian(data, 16, n * 16, 1, n, 0, 4, 8)
    m2 = _mapper(pkg, cloud=True, scan_size=1080)
    gb_pin, pb_pin = m2.scan_to_grid_baysian(pinned.array, 16, n * 16, 1, n, 0, 4, 8)
    assert np.array_equal(gb_page, gb_pin) and np.array_equal(pb_page.view(np.uint32), pb_pin.view(np.uint32))
    m.close()
    m2.close()
    angles = np.arange(0.0, 2 * math.pi, 2 * math.pi / 360)
    z = pkg.CriticalZoneCheckerGPU(1, 0, (0.51, 2.0), (0.22, 0.0, 0.4), (0, 0, 0.99, 0.0), 160.0, 0.3, 0.6,
                                   angles, 0.1, 2.0, 20.0)
    for fwd in (True, False):
        a = z.check(data, 16, n * 16, 1, n, 0, 4, 8, fwd)
        b = z.check(pinned.array, 16, n * 16, 1, n, 0, 4, 8, fwd)
        assert a == b
    z.close()
    pinned.free()
