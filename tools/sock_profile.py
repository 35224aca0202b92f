#!/usr/bin/env python3
"""Per-kernel device time of one socket / MD step (eager, event-bracketed) next to the wall-clock latency of the
graph-replayed step.   python tools/sock_profile.py [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fortnet_b200 as fb
import bench


def run(name, steps):
    ds, funcs, dims, wb, label = bench.workload(name, 1)
    ctx = fb.Context(device=0, precision=64)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    net = fb.Bpnn(ctx, dims, len(ds.atomic_numbers), "tanh")
    net.set_params(wb)
    rng = np.random.default_rng(1)
    traj = [ds.coords + rng.normal(scale=0.01, size=ds.coords.shape) for _ in range(8)]
    for k in range(20):
        ctx.socket_step(0, traj[k % 8])
    t0 = time.perf_counter()
    for k in range(steps):
        ctx.socket_step(0, traj[k % 8])
    t_graph = (time.perf_counter() - t0) / steps
    ctx.profile(True)
    for k in range(50):
        ctx.socket_step(0, traj[k % 8])
    rep = ctx.profile_report()
    ctx.profile(False)
    ctx.close()
    return {"workload": "%d atoms" % ds.n_atoms, "socket_step_us": t_graph * 1e6,
            "kernel_us": {k: round(v["ms_total"] / v["launches"] * 1e3, 2) for k, v in rep.items()},
            "launches_per_step": {k: v["launches"] / 50 for k, v in rep.items()}}


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    for name in ("c2", "c3"):
        print(json.dumps(run(name, steps)))
