"""Host-side logic of the Python front-end that needs no GPU: the result object cuts the five rows of a
kc_cycle_result out of ONE copy of the record (vx, vy, omega [P-1] each, then x, y [P] each, back to back:
include/kompass_b200.h), and the cycle call's argument checks."""
import ctypes as C

import numpy as np
import pytest


def _record(pkg, P, found=1):
    rows = np.arange(5 * P - 3, dtype=np.float32) + 0.5
    base = rows.ctypes.data
    fp = C.POINTER(C.c_float)
    r = pkg.CycleResult()
    r.found, r.cost, r.slot, r.n_points, r.n_slots, r.n_admissible = found, 1.25, 7, P, 99, 42
    r.vx = C.cast(base, fp)
    r.vy = C.cast(base + 4 * (P - 1), fp)
    r.omega = C.cast(base + 8 * (P - 1), fp)
    r.x = C.cast(base + 12 * (P - 1), fp)
    r.y = C.cast(base + 12 * (P - 1) + 4 * P, fp)
    return r, rows


@pytest.mark.parametrize("P", [2, 3, 50, 151])
def test_result_rows_are_cut_from_one_copy(pkg, P):
    r, rows = _record(pkg, P)
    t = pkg.TrajSearchResult(r)
    assert t.is_found and t.slot == 7 and t.n_points == P and t.n_admissible == 42 and t.cost == 1.25
    assert np.array_equal(t.vx, rows[:P - 1]) and np.array_equal(t.vy, rows[P - 1:2 * (P - 1)])
    assert np.array_equal(t.omega, rows[2 * (P - 1):3 * (P - 1)])
    assert np.array_equal(t.x, rows[3 * (P - 1):3 * (P - 1) + P]) and np.array_equal(t.y, rows[3 * (P - 1) + P:])
    assert len(t.x) == P and len(t.y) == P and len(t.vx) == P - 1
    rows[:] = -1.0  # the result owns a copy: the library reuses its record on the next cycle
    assert t.x[0] == 3 * (P - 1) + 0.5


def test_result_without_a_trajectory(pkg):
    r, _ = _record(pkg, 50, found=0)
    t = pkg.TrajSearchResult(r)
    assert not t.is_found and len(t.x) == 0 and len(t.vx) == 0


def test_cycle_cloud_rejects_a_ragged_cloud(pkg):
    """The fast path hands the array's address straight to the C entry point: the shape check that
    reshape(-1, 3) used to do is explicit."""
    p = pkg.Planner.__new__(pkg.Planner)  # no handle needed: the check comes before the call
    p._h, p._owned = C.c_void_p(), False
    with pytest.raises(ValueError):
        p.cycle_cloud((0, 0, 0), (0, 0, 0), np.zeros(10, np.float32), 0, 1)
