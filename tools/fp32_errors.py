#!/usr/bin/env python3
"""precision-32 error report (feeds the documented bound): raw and z-scored ACSF, per-atom
outputs and training gradient of the FP32 mode against the FP64 oracle, C2 / C3 / C5-like shapes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fortnet_b200 as fb
from fortnet_b200 import synthetic
from oracle import oracle as orc


def ntot(dims):
    return sum(a * b for a, b in zip(dims[:-1], dims[1:])) + dims[-1] + sum(dims)


cases = {
    "c2": (synthetic.si_bulk(n_struct=16, seed=20260001), fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16), [32, 20, 20, 1]),
    "c3": (synthetic.tio2(n_struct=4, seed=20260002), fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8]), [64, 32, 32, 32, 1]),
    "c5_like": (synthetic.dense_liquid(n_atoms=256, density_aa3=0.070, seed=99, n_struct=1), fb.GFunctions.from_auto_scheme(6.0 * fb.BOHR_PER_AA, 2, 64), [66, 16, 1]),
}
nt = os.cpu_count() or 1
out = {}
for name, (ds, funcs, dims) in cases.items():
    nsp = len(ds.atomic_numbers)
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(nsp, ntot(dims)))
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=nt)
    mu, sg = orc.zscore_stats(ds.offsets, ref, ds.weights)
    zref = orc.zscore_apply(ref, mu, sg)
    dd_o, raw_o = orc.grad(ds.offsets, zref, ds.globalsp, dims, "tanh", wb, "mse", ds.weights, ds.atomic_weights, ds.gtargets, ds.atargets, nthreads=nt)
    ctx = fb.Context(precision=32)
    ctx.upload(0, ds)
    a0 = fb.Acsf(ctx, funcs, standardize=False); a0.calculate(0); v = a0.features(0)
    acsf = fb.Acsf(ctx, funcs, standardize=True); acsf.calculate(0); z = acsf.features(0)
    net = fb.Bpnn(ctx, dims, nsp, "tanh"); net.set_params(wb)
    raw = net.predict_batch(0)
    dd, lossv = net.update_gradients(0, "mse")
    out[name] = {"raw_acsf_max_rel": float((np.abs(v - ref) / np.maximum(np.abs(ref), 1e-30)).max()),
                 "raw_acsf_max_abs_over_colmax": float((np.abs(v - ref).max(0) / np.maximum(np.abs(ref).max(0), 1e-300)).max()),
                 "zscored_max_abs": float(np.abs(z - zref).max()), "zscored_scale": float(np.abs(zref).max()),
                 "outputs_max_abs_over_scale": float(np.abs(raw - raw_o).max() / max(1.0, np.abs(raw_o).max())),
                 "gradient_max_abs_over_max": float(np.abs(dd - dd_o).max() / np.abs(dd_o).max()), "launch": ctx.acsf_launch_info(0)}
    ctx.close()
print(json.dumps(out, indent=1))
