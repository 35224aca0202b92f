#!/usr/bin/env python3
"""C4 (BASELINE.json configs[3]): force-resolved prediction on a 1M-atom synthetic batch -- analytic
ACSF Cartesian derivatives + backprop forces, inference only.  Not the headline metric; the line
goes to profiles/.   python tools/bench_forces.py [n_struct] [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fortnet_b200 as fb
import bench


def main():
    n_struct = int(sys.argv[1]) if len(sys.argv) > 1 else 5209      # 5209 x 192 = 1 000 128 atoms
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ds, funcs, dims, wb, label = bench.workload("c3", n_struct)
    ctx = fb.Context(device=0, precision=64)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    zp = np.stack(acsf.zprec)
    net = fb.Bpnn(ctx, dims, len(ds.atomic_numbers), "tanh")
    net.set_params(wb)
    f = net.forces(0)                                   # warm-up (allocations, capacities)
    ctx.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        acsf.calculate(0, zprec=zp)
        raw = net.predict_batch(0)
        f = net.forces(0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    prof = ctx.profile_report()
    ctx.profile(False)
    # translation invariance: the forces of every structure sum to zero
    fsum = np.add.reduceat(f, ds.offsets[:-1].astype(int), axis=0)
    out = {"metric": "force_prediction_atoms_per_s", "value": ds.n_atoms / dt, "unit": "atoms/s", "ms_per_step": dt * 1e3,
           "config": {"workload": "C4 force-resolved prediction: %d TiO2-like structures x 192 atoms = %d atoms, 64 ACSF, "
                                  "64-32-32-32-1, energies + analytic forces, results copied to the host" % (n_struct, ds.n_atoms)},
           "kernel_ms_per_step": {k: round(v["ms_total"] / steps, 3) for k, v in prof.items()},
           "d2h_bytes_per_step": int(raw.nbytes + f.nbytes),
           "max_abs_force_sum_per_structure": float(np.abs(fsum).max()), "max_abs_force": float(np.abs(f).max())}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
