"""Helpers that run the same DWA cycle through the CPU oracle and through the C-ABI."""
import numpy as np

import orc


def _split(kw):
    w = kw.get("weights", (1.0,) * 5)
    common = dict(control_type=kw["control_type"], time_step=kw["time_step"],
                  prediction_horizon=kw["prediction_horizon"], control_horizon=kw["control_horizon"],
                  max_linear_samples=kw["max_linear_samples"], max_angular_samples=kw["max_angular_samples"],
                  vx=kw["vx"], vy=kw["vy"], omega=kw["omega"], shape=kw["shape"], dims=kw["dims"],
                  sensor_position=kw["sensor_position"], sensor_rotation=kw["sensor_rotation"],
                  octree_resolution=kw["octree_resolution"], drop_samples=kw["drop_samples"])
    if "num_ctrl_points" in kw:
        common["num_ctrl_points"] = kw["num_ctrl_points"]
    vy_acc = kw["vy"][1]
    ccfg = orc.cost_cfg(w_path=w[0], w_goal=w[1], w_obstacles=w[2], w_smooth=w[3], w_jerk=w[4],
                        acc_limits=(kw["vx"][1], vy_acc, kw["omega"][1]),
                        sensor_position=kw["sensor_position"], sensor_rotation=kw["sensor_rotation"])
    return common, ccfg


def custom_terms(samples, path, customs):
    """[n, n_custom] doubles weight_k * float(custom_k(trajectory, path)) (cost_evaluator.cpp:96-100;
    CustomCostFunction returns float, cost_evaluator.h:104-105)."""
    n = len(samples["slots"])
    out = np.zeros((n, len(customs)), np.float64)
    pd = dict(X=path.X, Y=path.Y, acc=path.acc, total_length=path.total_length)
    for i in range(n):
        t = {k: samples[k][i] for k in ("vx", "vy", "omega", "x", "y")}
        for k, (w, fn) in enumerate(customs):
            out[i, k] = float(w) * float(np.float32(fn(t, pd)))
    return out


def run_oracle_cycle(kw, path, seg, vel, pose, scan=None, cloud=None, n_threads=1, max_traj=None,
                     customs=None):
    """DWA::findBestPath through the oracle. Returns dict with winner + per-slot costs."""
    common, ccfg = _split(kw)
    scfg = orc.sampler_cfg(max_num_threads=n_threads, **common)
    samples = orc.sampler_generate(scfg, vel, pose, scan=scan, cloud=cloud)
    n = len(samples["slots"])
    out = dict(n_admissible=n, P=samples["P"], samples=samples)
    if n == 0:
        out.update(found=False, cost=0.0, slot=-1, costs=np.zeros(0, np.float32))
        return out
    D = float(np.float32(kw.get("max_local_range", 10.0)) / np.float32(3.0))
    obs = orc.cost_points(ccfg, pose, scan=scan, cloud=cloud)
    if len(obs[0]) == 0:
        obs = None
    ev = samples
    if max_traj is not None and n > max_traj:
        ev = {k: (v[:max_traj] if isinstance(v, np.ndarray) else v) for k, v in samples.items()}
    cu = custom_terms(ev, path, customs) if customs else None
    found, idx, cost, costs = orc.cost_evaluate(ccfg, ev, path, seg, obs, D, custom=cu, n_threads=n_threads)
    out.update(found=found, cost=cost, row=idx, slot=int(samples["slots"][idx]) if found else -1,
               costs=costs)
    if found:
        out.update(x=ev["x"][idx], y=ev["y"][idx], vx=ev["vx"][idx], vy=ev["vy"][idx],
                   omega=ev["omega"][idx])
    return out


def make_planner(pkg, kw, path=None):
    cfg = pkg.planner_config(**{k: v for k, v in kw.items()})
    p = pkg.Planner(cfg)
    if path is not None:
        p.set_path(path.X, path.Y, path.acc, path.total_length)
    return p


def assert_cycle_parity(got, ref, costs_gpu=None, adm_gpu=None, rtol=1e-4):
    """winner index bit-exact, winner cost within rtol (we additionally expect identical bits),
    per-trajectory costs within rtol relative (north_star tolerance)."""
    assert got.is_found == ref["found"]
    assert got.n_admissible == ref["n_admissible"], (got.n_admissible, ref["n_admissible"])
    if not ref["found"]:
        return
    assert got.slot == ref["slot"], (got.slot, ref["slot"], got.cost, ref["cost"])
    assert abs(got.cost - ref["cost"]) <= rtol * abs(ref["cost"]) + 1e-12
    assert np.array_equal(got.x, ref["x"]), np.abs(got.x - ref["x"]).max()
    assert np.array_equal(got.y, ref["y"]), np.abs(got.y - ref["y"]).max()
    assert np.array_equal(got.vx, ref["vx"]) and np.array_equal(got.omega, ref["omega"])
    if costs_gpu is not None:
        slots = ref["samples"]["slots"][: len(ref["costs"])]
        assert adm_gpu.sum() == ref["n_admissible"]
        assert np.all(adm_gpu[ref["samples"]["slots"]] == 1)
        g = costs_gpu[slots]
        r = ref["costs"]
        err = np.abs(g - r) / np.maximum(np.abs(r), 1e-12)
        assert err.max() <= rtol, (err.max(), int(err.argmax()))
