#!/usr/bin/env python3
"""Latency of one MD / i-PI step for a single resident structure (the socket path,
prg_fnet/fortnet.F90:430-609): fnetgpu_socket_step vs the blocking call sequence
coords_update -> acsf_calculate -> predict -> forces.   python tools/md_latency.py [steps]"""
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
    zp = np.stack(acsf.zprec)
    for k in range(20):
        ctx.socket_step(0, traj[k % 8])
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for k in range(steps):
        g, raw, frc = ctx.socket_step(0, traj[k % 8])
    t_fused = (time.perf_counter() - t0) / steps
    launches = (ctx.launch_count() - l0) / steps
    for k in range(20):
        ctx.update_coords(0, traj[k % 8]); acsf.calculate(0, zprec=zp); net.predict_batch(0); net.forces(0)
    t0 = time.perf_counter()
    for k in range(steps):
        ctx.update_coords(0, traj[k % 8]); acsf.calculate(0, zprec=zp); r2 = net.predict_batch(0); f2 = net.forces(0)
    t_seq = (time.perf_counter() - t0) / steps
    ok = bool(np.allclose(r2, raw, rtol=1e-12, atol=1e-12) and np.allclose(f2, frc, rtol=1e-9, atol=1e-12))
    ctx.close()
    return {"workload": label.split(",")[0] + ", ONE structure of %d atoms" % ds.n_atoms, "steps": steps,
            "socket_step_us": t_fused * 1e6, "blocking_sequence_us": t_seq * 1e6, "kernel_launches_per_step": launches,
            "same_result": ok, "atoms_per_s": ds.n_atoms / t_fused}


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    for name in ("c2", "c3"):
        print(json.dumps(run(name, steps)))
