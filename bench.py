#!/usr/bin/env python
"""bench.py — driver contract for the B200-native DWA hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one DWA control cycle of one robot on BASELINE.json configs[1]: differential drive,
~10k velocity slots x 50 points (dt 0.02 s, 1 s horizon) against a 100 000-point synthetic cloud,
all five cost weights = 1. Unit of work: trajectory-steps = slots x points per cycle.

  value  device-resident throughput: K cycles enqueued back to back on the planner's stream, cycle i
         reading cloud (i mod B) of a bank of B distinct clouds (B x 1.2 MB > the 126 MB L2, so every
         cycle reads its cloud cold from HBM); timed with CUDA events on that stream (inside the
         C-ABI, because torch.cuda.Event only sees torch's streams); max over ranks.
  e2e    the same metric through the public call kc_planner_cycle_cloud with HOST buffers: per step
         one H2D of the cloud (staged through pinned memory) and one D2H of the winner; wall clock.
  N > 1  weak scaling: each rank (one process per GPU, torchrun) plans for its own robot / clouds;
         no data-path collective (robots are independent, north_star); barrier + max over ranks.
  --impl reference   the CPU restatement of the reference path (oracle/, the reference itself cannot
         be compiled in this image: no Eigen/FCL/octomap) with all host threads, bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "dwa_trajectory_steps_per_s"
UNIT = "trajectory-steps/s"
N_POINTS_CLOUD = 100_000
BANK = 128  # 128 x 1.2 MB = 154 MB > 126 MB L2


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL prints its version banner to stdout)
    must not add lines: keep a private handle of the real stdout and point fd 1 at stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload(rank):
    import orc  # path prep only (interpolation of the reference path happens above the hot path)
    import workloads as wl

    kw = wl.cfg_c2()
    path = orc.Path(wl.straight_points(20.0), 0.01, 1.0)
    seg = wl.tracked_segment(path, 0, 2.0)
    vel, pose = (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    return wl, kw, path, seg, vel, pose


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return world, rank, local, dist


def _is_nccl(dist):
    return dist is not None and dist.get_backend() == "nccl"


def barrier(dist, local):
    if dist is not None:
        if _is_nccl(dist):
            import torch

            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
        else:
            dist.barrier()


def max_over_ranks(dist, value, local):
    """MAX over ranks of a per-rank scalar (device time). Works on nccl (GPU box) and gloo (CPU tests)."""
    if dist is None:
        return value
    import torch

    dev = f"cuda:{local}" if _is_nccl(dist) else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_robots(n_robots, world, rank):
    """Contiguous shard of independent robots for this rank: [lo, hi). north_star: the sweep is
    partitioned by robot, no data-path collective."""
    lo = (n_robots * rank) // world
    hi = (n_robots * (rank + 1)) // world
    return lo, hi


def gather_results(dist, local_results, world, rank):
    """Final result gather: rank 0 receives every rank's [(found, cost, slot, n_admissible), ...]
    in robot order (8-16 bytes per robot; the only exchange of the sweep)."""
    if dist is None:
        return list(local_results)
    out = [None] * world if rank == 0 else None
    dist.gather_object(list(local_results), out, dst=0)
    if rank != 0:
        return None
    flat = []
    for part in out:
        flat.extend(part)
    return flat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def algorithmic_flops(n_adm, n_slots, P, M, S):
    """brute-force-equivalent FLOPs of one cycle (FMA = 2): what the reference's
    loops execute. Obstacle term N*P*M*6, path N*P*S*8, goal N*S*6, smooth+jerk N*(P-1)*3*7,
    rollout slots*(P-1)*40."""
    return (n_adm * P * M * 6.0 + n_adm * P * S * 8.0 + n_adm * S * 6.0 + n_adm * (P - 1) * 21.0 +
            n_slots * (P - 1) * 40.0)


def cpu_sample(wl, kw, path, seg, vel, pose, cloud, n_threads, budget_s):
    """Time the oracle on a bounded sample of one cycle: the sampler over ALL slots plus the five
    cost terms over the first `m` admissible trajectories against the full cloud; extrapolate the
    cost part to all admissible trajectories. Returns (traj_steps_per_s, description, seconds)."""
    import orc
    from parity_util import _split

    common, ccfg = _split(kw)
    scfg = orc.sampler_cfg(max_num_threads=n_threads, **common)
    t0 = time.perf_counter()
    samples = orc.sampler_generate(scfg, vel, pose, cloud=cloud)
    t_sampler = time.perf_counter() - t0
    n_adm = len(samples["slots"])
    n_slots = len(orc.velocity_samples(scfg, vel)[0])
    P = samples["P"]
    D = float(np.float32(kw["max_local_range"]) / np.float32(3.0))
    obs = orc.cost_points(ccfg, pose, cloud=cloud)
    # calibrate on a few trajectories, then size the sample to the budget
    m0 = max(1, min(n_adm, 4 * n_threads))
    sub = {k: (v[:m0] if isinstance(v, np.ndarray) else v) for k, v in samples.items()}
    t0 = time.perf_counter()
    orc.cost_evaluate(ccfg, sub, path, seg, obs, D, n_threads=n_threads)
    per_traj = (time.perf_counter() - t0) / m0
    m = int(max(m0, min(n_adm, budget_s / max(per_traj, 1e-9))))
    sub = {k: (v[:m] if isinstance(v, np.ndarray) else v) for k, v in samples.items()}
    t0 = time.perf_counter()
    orc.cost_evaluate(ccfg, sub, path, seg, obs, D, n_threads=n_threads)
    t_cost = time.perf_counter() - t0
    t_cycle = t_sampler + t_cost * (n_adm / max(m, 1))
    desc = (f"oracle sampler over all {n_slots} slots ({t_sampler:.2f} s) + 5 cost terms over the first "
            f"{m} of {n_adm} admissible trajectories vs the full {len(cloud)}-point cloud ({t_cost:.2f} s), "
            f"cost part extrapolated to all admissible; {n_threads} thread(s)")
    return n_slots * P / t_cycle, desc, t_sampler + t_cost, t_cycle


def run_sweep(pkg, wl, kw, path, seg, world, rank, local, dist, robots_total, iters):
    """north_star config 5: `robots_total` independent robots (own cloud, own current velocity),
    sharded contiguously by robot id across the ranks; no data-path collective, the final gather
    carries 20 bytes per robot. Returns a dict for the JSON line (rank 0) or None."""
    from parity_util import make_planner

    lo, hi = shard_robots(robots_total, world, rank)
    R = hi - lo
    n_distinct = 16
    base = [wl.cloud_bench(5000 + s) for s in range(n_distinct)]
    vels, poses, clouds = [], [], []
    for r in range(lo, hi):
        rng = np.random.default_rng(wl.SEED + 7 * r)
        vels.append((float(rng.uniform(0.0, 2.0)), 0.0, float(rng.uniform(-2.0, 2.0))))
        poses.append((0.0, 0.0, 0.0))
        clouds.append(base[r % n_distinct])
    planner = make_planner(pkg, kw, path)
    # end to end: every robot's own cloud sits in one page-locked host array (R x 100k x 12 B); the
    # call uploads it chunk by chunk beside the computation and returns the per-robot winners. One
    # untimed warm-up sweep first (device buffers are grown on first use, like the W warm-up steps).
    n_pts = len(base[0])
    host = pkg.PinnedArray((R * n_pts, 3), np.float32)
    for i, c in enumerate(clouds):
        host.array[i * n_pts:(i + 1) * n_pts] = c
    offsets = np.arange(R, dtype=np.int64) * n_pts
    counts = np.full(R, n_pts, np.int32)
    planner.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
    barrier(dist, local)
    t0 = time.perf_counter()
    res = planner.batch_cloud(vels, poses, host.array, seg[0], seg[1], offsets=offsets, counts=counts)
    e2e_s = time.perf_counter() - t0
    host.free()
    slots = list(planner.batch_slots)
    # device resident: the same launch set replayed on the resident batch
    planner.batch_replay(1, R)
    barrier(dist, local)
    ms, res2 = planner.batch_replay(iters, R)
    barrier(dist, local)
    ms = max_over_ranks(dist, ms, local)
    e2e_s = max_over_ranks(dist, e2e_s, local)
    assert [x[2] for x in res] == [x[2] for x in res2], "replayed sweep changed its winners"
    P = planner.num_points
    local_units = float(sum(slots)) * P
    units = local_units
    if dist is not None:
        import torch
        dev = f"cuda:{local}" if _is_nccl(dist) else "cpu"
        t = torch.tensor([local_units], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        units = float(t.item())
    gathered = gather_results(dist, res, world, rank)
    planner.close()
    if rank != 0:
        return None
    found = sum(1 for g in gathered if g[0])
    return {
        "workload": "configs[4]: %d independent robots x ~10k slots x %d points vs their own 100k-point cloud, "
                    "sharded by robot over %d GPU(s)" % (robots_total, P, world),
        "robots": robots_total, "robots_per_gpu": R, "iters": iters,
        "ms_per_sweep": ms / iters, "value": units / (ms / iters * 1e-3), "unit": UNIT,
        "robots_per_s": robots_total / (ms / iters * 1e-3),
        "e2e_ms_per_sweep": e2e_s * 1e3, "e2e_value": units / e2e_s,
        "h2d_bytes_per_sweep_per_gpu": R * N_POINTS_CLOUD * 12, "d2h_bytes_per_sweep_per_gpu": R * 20,
        "robots_with_a_trajectory": found, "scaling": "strong (fixed %d-robot job)" % robots_total,
    }


def run_aux(pkg, wl):
    """p50 latency of the other two entry points of the path through their public calls (host buffers
    in, result in host memory): LocalMapperGPU.scan_to_grid (config 4: 400x400 @ 0.05 m, 1080 beams)
    and CriticalZoneCheckerGPU.check on a 100k-point cloud."""
    out = {}
    angles, ranges = wl.mapping_scan(1080)
    mp = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, False, 1080, 2 * math.pi / 1080, 2.0, 0.1, 20.0)
    lat = []
    for i in range(220):
        t0 = time.perf_counter()
        g = mp.scan_to_grid(angles, ranges, copy=False)  # view of the mapper's buffer, as the reference binding returns
        lat.append(time.perf_counter() - t0)
    out["mapper_scan_to_grid_p50_ms"] = float(np.percentile(lat[20:], 50) * 1e3)
    out["mapper_grid"] = "400x400 @ 0.05 m, 1080 beams; occupied %d empty %d" % (int((g == 100).sum()), int((g == 0).sum()))
    mp.close()
    pts = wl.cloud_lattice(0)
    data = wl.cloud_bytes_xyz16(pts)
    ang = np.array([2 * math.pi * i / 360 for i in range(360)], np.float64)
    cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.POINTCLOUD, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                    (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, ang, 0.1, 2.0, 20.0)
    lat = []
    for i in range(220):
        t0 = time.perf_counter()
        f = cz.check(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True)
        lat.append(time.perf_counter() - t0)
    out["critical_zone_cloud_p50_ms"] = float(np.percentile(lat[20:], 50) * 1e3)
    out["critical_zone_cloud"] = "100000 points x 16 B (pageable numpy buffer), 360 bins; factor %.4f" % f
    # the same cloud in page-locked memory: read in place by the binning kernel
    pinned = pkg.PinnedArray(data.shape, np.int8)
    pinned.array[...] = data
    lat = []
    for i in range(220):
        t0 = time.perf_counter()
        f2 = cz.check(pinned.array, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True)
        lat.append(time.perf_counter() - t0)
    out["critical_zone_cloud_pinned_p50_ms"] = float(np.percentile(lat[20:], 50) * 1e3)
    assert f2 == f
    cz.close()
    pinned.free()
    return out


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import orc

    orc.build()
    wl, kw, path, seg, vel, pose = workload(0)
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    # every step is a bounded sample sized so that the whole run stays near 100 s of CPU work: the
    # parts that do not change between steps on the same cloud (sampler over all slots, timed once
    # per distinct cloud; cost-frame obstacle points; per-trajectory calibration) are hoisted, a step
    # times the five cost terms over the first m admissible trajectories against the full cloud
    from parity_util import _split
    total_s = float(os.environ.get("KC_BENCH_REF_SECONDS", "100"))  # CPU work of the whole run, roughly
    budget = max(0.02, min(3.0, total_s / max(steps + warmup, 1)))
    common, ccfg = _split(kw)
    scfg = orc.sampler_cfg(max_num_threads=threads, **common)
    D = float(np.float32(kw["max_local_range"]) / np.float32(3.0))
    n_slots = len(orc.velocity_samples(scfg, vel)[0])
    prepared = {}

    def prepare(ci):
        cloud = wl.cloud_bench(ci)
        t0 = time.perf_counter()
        samples = orc.sampler_generate(scfg, vel, pose, cloud=cloud)
        t_sampler = time.perf_counter() - t0
        obs = orc.cost_points(ccfg, pose, cloud=cloud)
        n_adm = len(samples["slots"])
        # (the port hands its threads blocks of trajectories: samples below ~16 per thread leave
        # threads idle and would make the CPU look slower than it is, so that is the floor)
        m0 = max(1, min(n_adm, 16 * threads))
        sub = {k: (v[:m0] if isinstance(v, np.ndarray) else v) for k, v in samples.items()}
        t0 = time.perf_counter()
        orc.cost_evaluate(ccfg, sub, path, seg, obs, D, n_threads=threads)
        per_traj = (time.perf_counter() - t0) / m0
        m = int(max(m0, min(n_adm, budget / max(per_traj, 1e-9))))
        m = min(n_adm, max(m0, m - m % (8 * threads)))
        sub = {k: (v[:m] if isinstance(v, np.ndarray) else v) for k, v in samples.items()}
        return dict(samples=sub, obs=obs, n_adm=n_adm, m=m, t_sampler=t_sampler, P=samples["P"], n_cloud=len(cloud))

    vals, secs, desc, tcyc = [], [], "", []
    for i in range(warmup + steps):
        c = prepared.get(i % 4)
        if c is None:
            c = prepared[i % 4] = prepare(i % 4)
        t0 = time.perf_counter()
        orc.cost_evaluate(ccfg, c["samples"], path, seg, c["obs"], D, n_threads=threads)
        t_cost = time.perf_counter() - t0
        t_cycle = c["t_sampler"] + t_cost * (c["n_adm"] / max(c["m"], 1))
        if i >= warmup:
            vals.append(n_slots * c["P"] / t_cycle)
            secs.append(t_cost)
            tcyc.append(t_cycle)
            desc = (f"per step: 5 cost terms over the first {c['m']} of {c['n_adm']} admissible trajectories vs "
                    f"the full {c['n_cloud']}-point cloud ({t_cost:.3f} s), extrapolated to all admissible, + the "
                    f"oracle sampler over all {n_slots} slots ({c['t_sampler']:.2f} s, timed once per distinct "
                    f"cloud); {threads} thread(s)")
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": float(np.mean(tcyc) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: DWA differential-drive, ~10k slots x 50 points vs 100k-point cloud, "
                               "all cost weights = 1", "cloud_points": N_POINTS_CLOUD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": desc + f"; ms_per_step is the extrapolated full-cycle time"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference cannot be compiled here (Eigen/FCL/octomap absent): oracle port with analytic "
                "voxel collision, which is cheaper than FCL -> this baseline flatters the CPU",
    }
    emit(line)
    return 0


def run_ours(args):
    world, rank, local, dist = dist_setup(args.gpus)
    import __graft_entry__ as ge

    pkg = ge.load_package()
    wl, kw, path, seg, vel, pose = workload(rank)
    from parity_util import make_planner

    planner = make_planner(pkg, kw, path)
    bank = BANK if not args.small_bank else 8
    planner.bank_alloc(bank, N_POINTS_CLOUD)
    clouds = []
    for s in range(bank):
        c = wl.cloud_bench(1000 * rank + s)
        planner.bank_upload(s, c)
        if s < 16:
            clouds.append(c)
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- warm-up + device-resident timed region -------------------------------------------------
    # untimed: W warm-up cycles, and at least one pass over the whole bank so that every resident
    # cloud's launch graph is captured before the clock starts
    planner.replay(0, max(warmup, bank), vel, pose, seg[0], seg[1])
    barrier(dist, local)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = planner.launch_count
    total_ms, _, last = planner.replay(warmup, steps, vel, pose, seg[0], seg[1])
    launches = planner.launch_count - l0
    barrier(dist, local)
    total_ms = max_over_ranks(dist, total_ms, local)
    n_slots, P = last.n_slots, last.n_points
    units_per_step = n_slots * P
    value = world * steps * units_per_step / (total_ms * 1e-3)

    # ---- dominant kernel (rollout+collision+cost) timed live with CUDA events --------------------
    ksteps = min(steps, 500)
    _, eval_ms, _ = planner.replay(warmup, ksteps, vel, pose, seg[0], seg[1], time_eval=True)
    eval_us = eval_ms * 1e3 / ksteps

    # ---- end-to-end through the public host-buffer call -----------------------------------------
    # inputs sit in page-locked host memory (the contract's "from pinned host memory"): every step
    # DMAs its own 1.2 MB cloud host->device and reads the winner back. A second, shorter loop
    # repeats the measurement with ordinary pageable numpy arrays (staged through the handle's
    # pinned buffer) and is reported beside it.
    esteps = min(steps, 1000)
    pinned = []
    for c in clouds:
        pa = pkg.PinnedArray(c.shape, np.float32)
        pa.array[...] = c
        pinned.append(pa)
    for i in range(min(warmup, 20)):
        planner.cycle_cloud(vel, pose, pinned[i % len(pinned)].array, seg[0], seg[1])
    barrier(dist, local)
    lat = np.zeros(esteps)
    t_begin = time.perf_counter()
    for i in range(esteps):
        t0 = time.perf_counter()
        r = planner.cycle_cloud(vel, pose, pinned[i % len(pinned)].array, seg[0], seg[1])
        lat[i] = time.perf_counter() - t0
    e2e_s = time.perf_counter() - t_begin
    e2e_s = max_over_ranks(dist, e2e_s, local)
    psteps = min(esteps, 300)
    lat_pageable = np.zeros(psteps)
    for i in range(psteps):
        t0 = time.perf_counter()
        planner.cycle_cloud(vel, pose, clouds[i % len(clouds)], seg[0], seg[1])
        lat_pageable[i] = time.perf_counter() - t0
    for pa in pinned:
        pa.free()
    clocks = sampler.stop() if sampler else None
    e2e_value = world * esteps * units_per_step / e2e_s
    h2d = N_POINTS_CLOUD * 12 + 4096
    d2h = 32 + 4 * (5 * P)

    # ---- brute-force reference kernel (verification hook, outside every timed region above) --------
    # the obstacle term as the reference's loops execute it (N*P*M pairs, no culling), on the GPU:
    # the FP32-roofline formulation the pruned cycle is compared with
    brute = None
    if rank == 0:
        try:
            r = planner.cycle_cloud(vel, pose, clouds[0], seg[0], seg[1])
            samples = [planner.bruteforce_obstacle_costs(r.n_slots)[1:] for _ in range(3)]
            bf_ms = float(np.median([s[0] for s in samples]))
            brute = {"pairs": samples[0][1], "ms_fp32_pass": bf_ms}
        except Exception as e:
            brute = {"error": repr(e)}
    planner.close()
    sweep = None
    if args.sweep_robots > 0:
        sweep = run_sweep(pkg, wl, kw, path, seg, world, rank, local, dist, args.sweep_robots, args.sweep_iters)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    traffic, executed = None, None
    aux = None
    try:
        aux = run_aux(pkg, wl)
    except Exception as e:  # the headline must not depend on the auxiliary entry points
        aux = {"error": repr(e)}
    ncu = None
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "eval_kernel_ncu_summary.json")))
        pair = [k for k in ncu.get("kernels", [])
                if any(n in k.get("kernel", "") for n in ("k_rollout_collide", "k_cost_bounds", "k_cost_split",
                                                          "k_cost_eval"))]
        if pair:  # the trajectory kernels the live CUDA events bracket
            traffic = sum(k.get("dram_traffic_bytes", 0.0) for k in pair)
            executed = {"executed_fp32_flop": sum(k.get("executed_fp32_flop", 0.0) for k in pair),
                        "executed_fp64_flop": sum(k.get("executed_fp64_flop", 0.0) for k in pair),
                        "issue_slot_busy_pct_when_active": max(k.get("issue_slot_busy_pct_when_active", 0.0) for k in pair)}
    except Exception:
        pass

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    fp32_peak = pkg.measure_fp32_peak_tflops() if hasattr(pkg, "measure_fp32_peak_tflops") else None
    S = seg[1]
    flops = algorithmic_flops(last.n_admissible, n_slots, P, N_POINTS_CLOUD, S)
    achieved = flops / (eval_us * 1e-6) / 1e12
    roofline = {
        "kernel": "k_rollout_collide<false> + k_cost_bounds + k_cost_split + k_cost_eval (the trajectory kernels, timed back to back on one stream)", "bound": "fp32", "unit": "TFLOP/s",
        "achieved": achieved, "peak": fp32_peak, "frac": (achieved / fp32_peak) if fp32_peak else None,
        "peak_source": "measured live: FP32 FMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP32 figure)",
        "kernel_us": eval_us, "share_of_step": eval_us / (total_ms * 1e3 / steps),
        "algorithmic_flop_per_launch": flops,
        "note": "achieved = brute-force-equivalent FLOPs of the reference loops (SURVEY 8d) / measured kernel "
                "time; > 1.0 is expected and is evidence of exact culling (grid-pruned nearest-obstacle search), "
                "not of a measurement error. Executed-instruction fractions from ncu are in profiles/.",
        "traffic": None,
        "hbm": {"bound": "hbm", "unit": "GB/s", "peak": peaks.get("hbm_gbs"),
                "achieved": (N_POINTS_CLOUD * 12) / (total_ms * 1e-3 / steps) / 1e9,
                "note": "algorithmic bytes of a whole cycle = the 1.2 MB cloud read once; the path is not HBM-bound"},
    }

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        import orc

        orc.build()
        v, desc, secs, tcyc = cpu_sample(wl, kw, path, seg, vel, pose, clouds[0], 1, args.cpu_budget)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc,
               "full_cycle_ms_extrapolated": tcyc * 1e3}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": "configs[1]: DWA differential-drive, %d velocity slots x %d points vs %d-point "
                               "cloud, all five cost weights = 1, cylinder r=0.2 robot, octree 0.1 m" %
                               (n_slots, P, N_POINTS_CLOUD),
                   "slots": n_slots, "points_per_trajectory": P, "admissible": last.n_admissible,
                   "cloud_points": N_POINTS_CLOUD, "tracked_segment_points": S,
                   "l2_policy": "inputs larger than L2: bank of %d distinct clouds (%.0f MB) cycled" %
                                (bank, bank * N_POINTS_CLOUD * 12 / 1e6),
                   "robots_per_gpu": 1},
        "p50_latency_ms": float(np.percentile(lat, 50) * 1e3),
        "p90_latency_ms": float(np.percentile(lat, 90) * 1e3),
        "p99_latency_ms": float(np.percentile(lat, 99) * 1e3),
        "p50_latency_ms_pageable_input": float(np.percentile(lat_pageable, 50) * 1e3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": esteps, "ms_per_step": e2e_s / esteps * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "winner": {"slot": last.slot, "cost": last.cost},
        "sweep": sweep, "aux": aux, "ncu_executed": ncu,
    }
    line["roofline"]["traffic"] = traffic
    if brute and "ms_fp32_pass" in brute and brute["ms_fp32_pass"] > 0:
        tf = brute["pairs"] * 6.0 / (brute["ms_fp32_pass"] * 1e-3) / 1e12
        brute.update({
            "kernel": "k_obstacle_bruteforce<false, packed>: minDist2D as written, every admissible trajectory "
                      "point x every cloud point (2 FADD + FMUL + FFMA + FMNMX per pair = 6 FLOP, SURVEY 8d; two "
                      "pairs per sm_100 packed FP32 instruction FADD2/FMUL2/FFMA2, same bits as the scalar form), "
                      "obstacle points staged through shared memory with cp.async, 8 register-resident entries "
                      "per lane",
            "achieved_tflops": tf, "frac_of_fp32_peak": (tf / fp32_peak) if fp32_peak else None,
            "vs_pruned_cycle": brute["ms_fp32_pass"] / (total_ms / steps),
            "note": "verification hook (tests/test_gpu_planner.py: the pruned search equals it bit for bit on "
                    "every slot at configs 2 and 3); the control path never runs it"})
    line["bruteforce_reference_kernel"] = brute
    if executed and fp32_peak and executed.get("executed_fp32_flop"):  # (absent when the capture had no op counters)
        # executed (not algorithmic) arithmetic of the same kernel from the committed ncu capture,
        # against the live-measured FP32 peak and the live kernel time
        ex = executed.get("executed_fp32_flop", 0.0) + executed.get("executed_fp64_flop", 0.0)
        line["roofline"]["executed_flop_per_launch_ncu"] = ex
        line["roofline"]["executed_frac_of_fp32_peak"] = ex / (eval_us * 1e-6) / 1e12 / fp32_peak
        line["roofline"]["issue_slot_busy_pct_ncu"] = executed.get("issue_slot_busy_pct_when_active")
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of oracle cost work")
    ap.add_argument("--small-bank", action="store_true", help="8-cloud bank (profiling runs)")
    ap.add_argument("--sweep-robots", type=int, default=1024,
                    help="robots of the batched multi-robot sweep (config 5), 0 = skip")
    ap.add_argument("--sweep-iters", type=int, default=3)
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
