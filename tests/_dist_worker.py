"""Worker for the multi-process tests (launched by torch.distributed.run).

mode cpu : gloo, every rank computes the ORACLE gradient of its structure shard and the
           host-side reduction of fortnet_b200.sharding sums them  (tests the host logic:
           atom-balanced sharding + [ddSerial | loss terms] reduction).
mode gpu : nccl, every rank runs libfnetgpu on its shard; the library's own NCCL all-reduce
           (z-score statistics, gradient + loss) must reproduce the single-process oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode = sys.argv[1]
    import torch
    import torch.distributed as dist
    import fortnet_b200 as fb
    from fortnet_b200 import sharding, synthetic
    from oracle import oracle as orc

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", 0))
    if mode == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")

    # ragged two-species dataset: TiO2 cells plus a few Si cells of different size
    ds = synthetic.tio2(n_struct=5, seed=11)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 4, 4).resolve_species([22, 8])
    dims = [len(funcs), 6, 5, 1]
    rng = np.random.default_rng(3)
    wb = rng.uniform(-0.5, 0.5, size=(2, orc.ntot(dims)))
    fd = funcs.asdicts()

    # single-process truth (oracle on the whole dataset)
    vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd)
    mu, sg = orc.zscore_stats(ds.offsets, vals, ds.weights)
    feats = orc.zscore_apply(vals, mu, sg)
    dd_ref, raw_ref = orc.grad(ds.offsets, feats, ds.globalsp, dims, "tanh", wb, "mse", ds.weights,
                               ds.atomic_weights, ds.gtargets, ds.atargets)
    loss_ref = orc.loss(ds.offsets, raw_ref, "mse", 1, 0, ds.gtargets, ds.atargets, ds.atomic_weights, ds.weights)

    sh = sharding.shard_dataset(ds, world, rank)
    if mode == "cpu":
        natoms = ds.offsets[1:] - ds.offsets[:-1]
        b, e = sharding.shard_bounds(natoms, world)[rank]
        f_sh = feats[ds.offsets[b]:ds.offsets[e]]
        dd, raw = orc.grad(sh.offsets, f_sh, sh.globalsp, dims, "tanh", wb, "mse", sh.weights,
                           sh.atomic_weights, sh.gtargets, sh.atargets)
        l = orc.loss(sh.offsets, raw, "mse", 1, 0, sh.gtargets, sh.atargets, sh.atomic_weights, sh.weights)
        den = float(np.sum(sh.weights * np.add.reduceat(sh.atomic_weights, sh.offsets[:-1].astype(int))))
        dd_all, num, den_all = sharding.allreduce_host(dd, l * den, den, dist)
        loss = num / den_all
    else:
        ctx = fb.Context(device=local)
        sharding.init_comm(ctx, dist)
        ctx.upload(0, sh)
        acsf = fb.Acsf(ctx, funcs, standardize=True)
        acsf.calculate(0)                       # statistics all-reduced inside the library
        assert np.allclose(acsf.zprec[0], mu, rtol=1e-9, atol=1e-10), np.abs(acsf.zprec[0] - mu).max()
        assert np.allclose(acsf.zprec[1], sg, rtol=1e-9, atol=1e-10), np.abs(acsf.zprec[1] - sg).max()
        net = fb.Bpnn(ctx, dims, 2, "tanh")
        net.set_params(wb)
        dd_all, loss = net.update_gradients(0, "mse")
        ctx.close()
    assert np.allclose(dd_all, dd_ref, rtol=1e-9, atol=1e-10), np.abs(dd_all - dd_ref).max()
    assert abs(loss - loss_ref) <= 1e-10 + 1e-9 * abs(loss_ref), (loss, loss_ref)
    dist.barrier()
    if rank == 0:
        print("DIST_OK mode=%s world=%d maxdiff=%.3e loss=%.12g" % (mode, world, np.abs(dd_all - dd_ref).max(), loss))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
