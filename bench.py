#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Fortnet hot path on B200.

metric : train atoms/s (ACSF featurisation + BPNN forward/backward + gradient reduction)
step   : one pass of the hot path over one batch of synthetic structures:
         fnetgpu_acsf_calculate (cell list resident, z-score fused) + fnetgpu_grad
workload (N = 1): BASELINE.json configs[1] -- single-species Si bulk, 10k structures x 64 atoms,
         32 ACSF features (Auto{RCut 4.0, NRadial 16, NAngular 16}), subnets [32,20,20,1] tanh,
         energy (one global target) training with MSE.  N > 1: the same shard per GPU (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, all host
threads; the Fortran binary cannot be built in this image) on a bounded sample of the same
workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_atoms_per_s"
UNIT = "atoms/s"


def workload(name, n_struct, seed_shift=0):
    import fortnet_b200 as fb
    from fortnet_b200 import synthetic
    if name == "c2":
        ds = synthetic.si_bulk(n_struct=n_struct, seed=20260001 + seed_shift)
        funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
        dims = [32, 20, 20, 1]
        label = "C2 single-species Si bulk, 10k structures x 64 atoms, 32 ACSF (G2x16+G5x16, rc 4 A), subnets 32-20-20-1 tanh, MSE energy training"
    elif name == "c3":
        ds = synthetic.tio2(n_struct=n_struct, seed=20260002 + seed_shift)
        funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8])
        dims = [64, 32, 32, 32, 1]
        label = "C3 two-species TiO2-like, 20k structures x 192 atoms, 64 species-resolved G2/G5 ACSF, subnets 64-32-32-32-1 tanh"
    else:
        raise SystemExit("unknown workload " + name)
    assert len(funcs) == dims[0]
    rng = np.random.default_rng(7)
    n_species = len(ds.atomic_numbers)
    # serialised length per species (network.F90:397-427): all weights incl. the dummy ww(d_L,1), all biases
    ntot = sum(a * b for a, b in zip(dims[:-1], dims[1:])) + dims[-1] + sum(dims)
    wb = rng.uniform(-0.5, 0.5, size=(n_species, ntot))
    return ds, funcs, dims, wb, label


class NvmlSampler:
    """SM clock / throttle reasons DURING the timed region through NVML in this process (a sample every
    ~4 ms; `nvidia-smi -lms` forks a heavier query that can land on a 2 ms step and lengthen it).
    Same quantities as the recipe's clocks line in B200_PROFILING.md; falls back to ClockSampler."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.ok = False
        self.samples = []
        self.run = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.h = h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        self.run = False
        self.t.join(timeout=1.0)
        sm = [a for a, _ in self.samples]
        reasons = set()
        for _, rs in self.samples:
            for bit, nm in self.REASONS.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(reasons), "how": "NVML in-process, one sample per ~4 ms inside the timed region"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(name, target_seconds=12.0, threads=None):
    """Times the oracle (C restatement of the reference algorithm, `kind: port`) on a bounded
    sample of the workload: ACSF once + one training-gradient pass per step."""
    from oracle import oracle as orc
    threads = threads or (os.cpu_count() or 1)
    per_thread = 1
    ds, funcs, dims, wb, _ = workload(name, max(threads * per_thread, 1))
    fd = funcs.asdicts()

    def one(dsx):
        t0 = time.perf_counter()
        vals = orc.acsf(dsx.offsets, dsx.coords, dsx.periodic, dsx.latvecs, dsx.atnum, fd, nthreads=threads)
        t1 = time.perf_counter()
        orc.grad(dsx.offsets, vals, dsx.globalsp, dims, "tanh", wb, "mse", dsx.weights, dsx.atomic_weights,
                 dsx.gtargets, dsx.atargets, nthreads=threads)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    ta, tg = one(ds)                                  # calibration: one structure per thread
    scale = max(1, int(target_seconds / max(ta + tg, 1e-3)))
    scale = min(scale, 64)
    ds2, _, _, _, _ = workload(name, threads * per_thread * scale)
    ta, tg = one(ds2)
    atoms = ds2.n_atoms
    # the reference's serial path: one thread on a sample sized for a few seconds
    n_ser = max(1, min(ds2.n_struct, int(round(ds2.n_struct * 3.0 / max((ta + tg) * threads, 1e-3)))))
    ds1, _, _, _, _ = workload(name, n_ser)
    t0 = time.perf_counter()
    v1 = orc.acsf(ds1.offsets, ds1.coords, ds1.periodic, ds1.latvecs, ds1.atnum, fd, nthreads=1)
    orc.grad(ds1.offsets, v1, ds1.globalsp, dims, "tanh", wb, "mse", ds1.weights, ds1.atomic_weights,
             ds1.gtargets, ds1.atargets, nthreads=1)
    t_ser = time.perf_counter() - t0
    return {"value": atoms / (ta + tg), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d structures (%d atoms): oracle ACSF %.2f s + fwd/bwd %.3f s, %d OpenMP threads, "
                      "reference block partition" % (ds2.n_struct, atoms, ta, tg, threads),
            "acsf_atoms_per_s": atoms / ta, "train_iter_atoms_per_s": atoms / tg,
            "serial_value": ds1.n_atoms / t_ser,
            "serial_sample": "%d structures (%d atoms) on 1 thread, %.2f s" % (ds1.n_struct, ds1.n_atoms, t_ser)}


def c4_forces(fb, n_struct=5209, steps=3):
    """C4: 5209 TiO2-like structures x 192 atoms = 1 000 128 atoms; ACSF (statistics given) + energies + analytic
    forces per step, results copied to the host; wall clock around the blocking API calls."""
    import torch
    ds, funcs, dims, wb, _ = workload("c3", n_struct, seed_shift=424242)
    ctx = fb.Context(device=torch.cuda.current_device(), precision=64)
    try:
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
        fsum = np.add.reduceat(f, ds.offsets[:-1].astype(int), axis=0)
        kms = {k: round(v["ms_total"] / steps, 3) for k, v in prof.items()}
        return {"metric": "force_prediction_atoms_per_s", "value": ds.n_atoms / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                "atoms": ds.n_atoms, "device_atoms_per_s": ds.n_atoms / (sum(kms.values()) * 1e-3),
                "kernel_ms_per_step": kms, "d2h_bytes_per_step": int(raw.nbytes + f.nbytes),
                "max_abs_force_sum_per_structure": float(np.abs(fsum).max()), "deterministic": True,
                "workload": "C4: %d TiO2-like structures x 192 atoms, 64 ACSF, 64-32-32-32-1, energies + analytic forces to the host" % n_struct}
    finally:
        ctx.close()


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads."""
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    name = args.workload
    # size one step to ~3 s
    ds, funcs, dims, wb, label = workload(name, threads)
    fd = funcs.asdicts()

    def step(dsx):
        vals = orc.acsf(dsx.offsets, dsx.coords, dsx.periodic, dsx.latvecs, dsx.atnum, fd, nthreads=threads)
        orc.grad(dsx.offsets, vals, dsx.globalsp, dims, "tanh", wb, "mse", dsx.weights, dsx.atomic_weights,
                 dsx.gtargets, dsx.atargets, nthreads=threads)

    t0 = time.perf_counter(); step(ds); t1 = time.perf_counter()
    scale = min(64, max(1, int(3.0 / max(t1 - t0, 1e-3))))
    ds, _, _, _, _ = workload(name, threads * scale)
    for _ in range(args.warmup):
        step(ds)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(ds)
    dt = (time.perf_counter() - t0) / args.steps
    val = ds.n_atoms / dt
    sample = "%d structures (%d atoms) per step, %d OpenMP threads" % (ds.n_struct, ds.n_atoms, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference algorithm (oracle/refcpu.c), not the Fortran binary: "
                "no Fortran compiler / HDF5 / MPI in this image",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--structures", type=int, default=0, help="structures per GPU (default: 10000 c2 / 20000 c3)")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's structure count per GPU; strong: that count split over the GPUs")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 force-prediction line (extra key, N = 1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE line, the JSON: everything libraries print while the bench runs (NCCL's version
    # banner, torchrun warnings of child imports) is sent to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import fortnet_b200 as fb
    from fortnet_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL logs to stdout by default: keep it to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_struct = args.structures or (10000 if args.workload == "c2" else 20000)
    if args.scaling == "strong":        # fixed total: contiguous blocks of structures (parallel.F90:43-54), remainder to the first ranks
        n_struct = n_struct // world + (1 if rank < n_struct % world else 0)
    ds, funcs, dims, wb, label = workload(args.workload, n_struct, seed_shift=1000 * rank)
    N = ds.n_atoms
    F = len(funcs)

    stream = torch.cuda.Stream()
    ctx = fb.Context(device=local, precision=args.precision)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        sharding.init_comm(ctx, dist)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)                      # first call: statistics of the training set (all ranks)
    net = fb.Bpnn(ctx, dims, len(ds.atomic_numbers), "tanh")
    net.set_params(wb)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step_resident():
        acsf.calculate(0)
        net.update_gradients(0, "mse", fetch=False)

    coords_pinned = torch.from_numpy(ds.coords.copy()).pin_memory()
    lat_pinned = torch.from_numpy(ds.latvecs.copy()).pin_memory()

    def step_e2e():
        acsf.calculate(0, coords=coords_pinned.numpy(), latvecs=lat_pinned.numpy())
        return net.update_gradients(0, "mse", fetch=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_resident()
        # ---------------- timed region: device-resident inputs ----------------
        sampler = NvmlSampler(local) if rank == 0 else None
        if rank == 0 and not sampler.ok:
            sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        ctx.profile(True)
        l0 = ctx.launch_count()
        for k in range(args.steps):
            flush.zero_()                                   # L2 flush, outside the event bracket
            ev[k][0].record(stream)
            step_resident()
            ev[k][1].record(stream)
        barrier()
        launches = ctx.launch_count() - l0
        prof = ctx.profile_report()
        ctx.profile(False)
        clocks = sampler.stop() if rank == 0 else None
        ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
        # ---------------- end to end: host buffers in, gradient + loss out ----------------
        # untimed warm-up of THIS path: the copy stream, its events and the pinned staging are created on its first calls,
        # and on a box whose PCIe link has been idle the host-to-device copies speed up over the first ~20 steps (measured:
        # 1.90 -> 1.33 ms per step, monotonically, in the first process on a fresh box; flat in the next process)
        n_e2e_warm = max(25, args.warmup)
        for _ in range(n_e2e_warm):
            step_e2e()
        barrier()
        t_e2e = 0.0
        e2e_steps_ms = []
        for k in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dd, loss = step_e2e()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t_e2e += dt
            e2e_steps_ms.append(round(dt * 1e3, 4))
        barrier()
        ms_e2e = t_e2e / args.steps * 1e3

    # C4 (BASELINE.json configs[3]): force-resolved prediction on 1 M atoms -- an extra key of the N = 1 line
    c4 = None
    if world == 1 and not args.no_c4:
        try:
            c4 = c4_forces(fb)
        except Exception as e:              # never lose the headline line to the extra one
            c4 = {"error": str(e)[:200]}
    mean_neigh = None
    try:
        mean_neigh = ctx.max_neighbors(0)[1]          # builds the cell list once: outside every timed region
    except Exception:
        pass
    grad_launch = None
    try:
        grad_launch = ctx.grad_launch_info(0)         # 0 two passes / 1 fused sums / 2 cluster-fused sums, cluster size, grid
    except Exception:
        pass
    live_peaks = None
    try:
        live_peaks = ctx.measure_peaks()              # this box, now (outside every timed region)
    except Exception:
        pass

    tms = torch.tensor([ms, ms_e2e, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tms.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tms.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        natoms = torch.tensor([float(N)], dtype=torch.float64, device="cuda"); dist.all_reduce(natoms)
        ms, ms_e2e, launches, total_atoms = tmax[0].item(), tmax[1].item(), int(tsum[2].item()), natoms.item()
    else:
        total_atoms = float(N)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        a = prof.get("acsf", {"ms_total": float("nan"), "launches": 1})
        acsf_ms = a["ms_total"] / max(a["launches"], 1)
        acsf_kernel_name = "k_acsf_lean" if ctx.acsf_kernel() == 1 else "k_acsf"
        sfeat = 8 if args.precision == 64 else 4
        bytes_per_atom = 3 * 8 + 4 + sfeat * F                   # coords + species + features
        alg_bytes = bytes_per_atom * N + 72 * ds.n_struct
        achieved = alg_bytes / (acsf_ms * 1e-3) / 1e9
        kshare = {k: v["ms_total"] / args.steps for k, v in prof.items()}
        # DRAM traffic of the same kernel: STATIC, from the committed ncu --set full capture of this build's
        # kernel (per launch, scaled by atoms when the launch size differs) -- a run under ncu is never a bench
        # run, so it cannot be measured here; None when no capture exists for this workload
        traffic, ncu_extra = None, {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json"))).get(args.workload)
            if tr and args.precision == 64:
                traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * (N / tr["atoms_per_launch"])
                ncu_extra = {"traffic_source": "static: " + tr["source"], "ncu_issue_active_pct": tr["issue_active_pct"],
                             "ncu_fp64_pipe_active_pct": tr["fp64_pipe_active_pct"],
                             "ncu_warp_inst_per_atom": tr["warp_inst_per_atom"], "ncu_kernel": tr["kernel"]}
        except Exception:
            pass
        # FP64 roofline of the same kernel: SURVEY.md 8(d) algorithmic flop-equivalents per atom,
        #   n (10 + 4 F_r) + n (n + 1) / 2 (25 + 4 F_a),  n = mean neighbours within rc,
        # against the measured DFMA peak of this pool's B200 (tools/peaks.cu -> profiles/r01_measured_peaks.json)
        roof64 = None
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "r01_measured_peaks.json")))
            f_r = sum(1 for f in funcs.func if f.radial)
            f_a = len(funcs.func) - f_r
            if len(ds.atomic_numbers) > 1:      # species-resolved copies share the pair geometry: per atom every pair feeds ONE copy
                f_r //= len(ds.atomic_numbers); f_a //= len(ds.atomic_numbers) * (len(ds.atomic_numbers) + 1) // 2
            nbar = mean_neigh if mean_neigh else 0.0
            flop_atom = nbar * (10 + 4 * f_r) + 0.5 * nbar * (nbar + 1) * (25 + 4 * f_a)
            ach = flop_atom * N / (acsf_ms * 1e-3) / 1e12
            peak64 = 2.0 * (live_peaks["dfma_tfma_s"] if live_peaks else pk["dfma_tfma_s"])
            roof64 = {"kernel": acsf_kernel_name, "bound": "fp64", "achieved": ach, "peak": peak64, "unit": "TFLOP/s", "frac": ach / peak64,
                      "algorithmic_flop_per_atom": flop_atom, "mean_neighbours": nbar,
                      "peak_source": ("DFMA issue peak measured on this box in this run (fnetgpu_measure_peaks: %.2f TFMA/s; "
                                      "round-1 box, tools/peaks.cu: %.1f)" % (live_peaks["dfma_tfma_s"], pk["dfma_tfma_s"])) if live_peaks else
                                     "measured DFMA issue peak, tools/peaks.cu (profiles/r01_measured_peaks.json: %.1f TFMA/s)" % pk["dfma_tfma_s"],
                      "note": "flop-equivalents of SURVEY.md 8(d) (c_a = 4 per angular function and pair), not executed instructions: "
                              "the kernel evaluates a whole xi-ladder from one table-driven power"}
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": total_atoms / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32",
            "data": "synthetic",
            "config": {"workload": label, "atoms_per_gpu": N, "structures_per_gpu": ds.n_struct,
                       "parallelism": "structure-sharded dp%d, one NCCL all-reduce of [ddSerial|loss] per step" % world,
                       "l2": "flushed between steps (256 MiB memset outside the per-step event brackets); "
                             "feature matrix %.0f MB > L2" % (N * F * sfeat / 1e6),
                       "step": "fnetgpu_acsf_calculate (z-score fused, stats from warm-up) + fnetgpu_grad"},
            "e2e": {"value": total_atoms / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "ms_steps_rank0": e2e_steps_ms, "warmup_steps": n_e2e_warm,
                    "h2d_bytes_per_step": int(ds.coords.nbytes + ds.latvecs.nbytes + 2 * F * 8),
                    "d2h_bytes_per_step": int((wb.size + 2) * 8 + 2 * 32),
                    "path": "fnetgpu_acsf_update_calculate(coords + lattices from pinned host memory, copy chunks overlapped with the ACSF kernel) -> fnetgpu_grad -> ddSerial + loss on the host (wall clock)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": acsf_kernel_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "algorithmic_bytes_per_atom": bytes_per_atom, "algorithmic_bytes": alg_bytes,
                         "avg_launch_ms": acsf_ms, "peak_source": peak_src,
                         "note": "FP64 angular ACSF is bound by FP64 issue, not by HBM (SURVEY.md 8d: ~13k flop-equivalents "
                                 "per atom against 284 B); the HBM fraction is reported as the contract asks, the binding "
                                 "figure is roofline_fp64 (and the ncu issue / FP64-pipe utilisation)", **ncu_extra},
            "roofline_fp64": roof64,
            "box_peaks_now": live_peaks,
            "kernel_ms_per_step": kshare,
            "grad_launch": grad_launch,
            "loss": loss,
        }
        if c4 is not None:
            out["c4_force_prediction"] = c4
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args.workload)
            out["cpu_baseline"] = cb
            # the blended headline ratio hides two very different ones: quote them separately
            grad_ms = sum(v for k, v in kshare.items() if k != "acsf")
            out["speedup_vs_cpu_baseline"] = {
                "cores": cb["cores"],
                "step_device": out["value"] / cb["value"],
                "step_e2e": out["e2e"]["value"] / cb["value"],
                "acsf_only": (N / (acsf_ms * 1e-3)) / cb["acsf_atoms_per_s"],
                "train_iteration_only": (N / (grad_ms * 1e-3)) / cb["train_iter_atoms_per_s"] if grad_ms > 0 else None,
                "note": "acsf_only is mostly algorithmic (the CPU port keeps the reference's per-(atom, function) neighbour-list rebuild); "
                        "train_iteration_only compares like with like (forward / backward / gradient reduction)"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
