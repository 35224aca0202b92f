#!/usr/bin/env python3
"""Wall-clock + per-kernel breakdown of one end-to-end step (host coords in -> gradient out).
  python tools/e2e_breakdown.py [c2|c3] [n_struct]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fortnet_b200 as fb
import bench


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_struct = int(sys.argv[2]) if len(sys.argv) > 2 else (10000 if name == "c2" else 20000)
    ds, funcs, dims, wb, label = bench.workload(name, n_struct)
    ctx = fb.Context(device=0, precision=64)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    net = fb.Bpnn(ctx, dims, len(ds.atomic_numbers), "tanh")
    net.set_params(wb)
    cp = torch.from_numpy(ds.coords.copy()).pin_memory().numpy()
    lp = torch.from_numpy(ds.latvecs.copy()).pin_memory().numpy()
    for rep in range(4):
        ctx.profile(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.update_coords(0, cp, lp)
        t1 = time.perf_counter()
        acsf.calculate(0)
        t2 = time.perf_counter()
        dd, loss = net.update_gradients(0, "mse", fetch=True)
        t3 = time.perf_counter()
        prof = ctx.profile_report()
        ctx.profile(False)
        print(json.dumps({"rep": rep, "coords_update_ms": (t1 - t0) * 1e3, "acsf_calculate_ms": (t2 - t1) * 1e3,
                          "grad_ms": (t3 - t2) * 1e3, "total_ms": (t3 - t0) * 1e3,
                          "kernels_ms": {k: round(v["ms_total"], 4) for k, v in prof.items()}}))
    ctx.close()


if __name__ == "__main__":
    main()
