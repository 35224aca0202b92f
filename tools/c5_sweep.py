#!/usr/bin/env python3
"""C5 (BASELINE.json configs[4]): ACSF-only stress on a dense liquid -- the HBM-roofline sweep.
4 096-atom periodic structures (rho = 0.070 / A^3) through the cell list; rc chosen for n = 25 / 50 / 100 / 150
neighbours, F_a = 16 / 64 / 128 angular functions (auto ladder) + a radial-only point per rc; both precisions.
Per point: atoms/s, achieved HBM GB/s of the algorithmic bytes (28 + s F per atom), fraction of the measured HBM
peak, algorithmic FP64/FP32 flop-equivalents/s (SURVEY.md 8d formula).   python tools/c5_sweep.py [n_struct] > out.json
"""
import json, os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fortnet_b200 as fb
from fortnet_b200 import synthetic


def main():
    n_struct = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rho = 0.070
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ds = synthetic.dense_liquid(n_atoms=4096, density_aa3=rho, seed=99, n_struct=n_struct)
    N = ds.n_atoms
    out = {"workload": "C5 dense liquid: %d structures x 4096 atoms, rho = %.3f / A^3, cell-list path" % (n_struct, rho),
           "hbm_peak_gbs": hbm_peak, "points": []}
    for prec in (64, 32):
        ctx = fb.Context(precision=prec)
        ctx.upload(0, ds)
        for n_target in (25, 50, 100, 150):
            rc_aa = (n_target / (rho * 4.0 / 3.0 * math.pi)) ** (1.0 / 3.0)
            for n_ang in (0, 16, 64, 128):
                n_rad = 16 if n_ang == 0 else 2
                funcs = fb.GFunctions.from_auto_scheme(rc_aa * fb.BOHR_PER_AA, n_rad, n_ang) if n_ang else \
                    fb.GFunctions(fb.GFunctions.from_auto_scheme(rc_aa * fb.BOHR_PER_AA, n_rad, 2).func[:n_rad])
                F = len(funcs)
                acsf = fb.Acsf(ctx, funcs, standardize=False)
                acsf.calculate(0)                         # capacities, cell list
                mx, mean = ctx.max_neighbors(0)
                reps = 3 if n_ang * n_target > 2000 else 10
                ctx.profile(True)
                for _ in range(reps):
                    acsf.calculate(0)
                ctx.synchronize()
                prof = ctx.profile_report()
                ctx.profile(False)
                ms = prof["acsf"]["ms_total"] / prof["acsf"]["launches"]
                li = ctx.acsf_launch_info(0)
                s = 8 if prec == 64 else 4
                bytes_atom = 28 + s * F
                gbs = bytes_atom * N / (ms * 1e-3) / 1e9
                flop_atom = mean * (10 + 4 * n_rad) + 0.5 * mean * (mean + 1) * (25 + 4 * n_ang) if n_ang else mean * (10 + 4 * n_rad)
                out["points"].append({
                    "precision": prec, "rc_A": round(rc_aa, 3), "mean_neighbours": round(mean, 1), "max_neighbours": mx,
                    "n_radial": n_rad, "n_angular": n_ang, "features": F, "kernel": "k_acsf_lean" if li["lean"] else "k_acsf", "launch": li, "acsf_ms": round(ms, 4), "atoms_per_s": N / (ms * 1e-3), "bytes_per_atom": bytes_atom,
                    "hbm_gbs": round(gbs, 2), "hbm_frac": gbs / hbm_peak, "tflop_equiv_per_s": flop_atom * N / (ms * 1e-3) / 1e12})
                print(json.dumps(out["points"][-1]), file=sys.stderr, flush=True)
        ctx.close()
    best = max(out["points"], key=lambda p: p["hbm_frac"])
    out["statement"] = ("largest HBM fraction of the sweep: %.1f %% (%s, precision %d, n = %.0f, %d radial + %d angular functions). "
                        "The north-star 60 %% of HBM peak is not reachable for ACSF: even the radial-only, n = 25 point needs ~%d flop-"
                        "equivalents per atom against %d bytes, and every angular point is bound by FP issue (see tflop_equiv_per_s)."
                        % (100 * best["hbm_frac"], best["kernel"], best["precision"], best["mean_neighbours"], best["n_radial"],
                           best["n_angular"], int(best["mean_neighbours"] * (10 + 4 * best["n_radial"])), best["bytes_per_atom"]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
