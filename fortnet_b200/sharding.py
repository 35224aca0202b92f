"""Structure-level data parallelism over GPUs (one process per GPU).

The reference splits the structure index range into contiguous blocks per MPI rank
(lib_common/parallel.F90:23-56, used at acsf.F90:599, bpnn.F90:257) and sum-allreduces padded
arrays.  Here every rank keeps only its shard resident on its GPU; shards are contiguous and
balanced by cumulative ATOM count (structures differ in size); the only exchanges are the
z-score statistics (once) and one all-reduce of [ddSerial | loss terms] per iteration, both
issued by libfnetgpu on its own stream (NCCL over NVLink).  On machines without NCCL/GPUs
(CPU tests) ``allreduce_host`` provides the same reduction through torch.distributed (gloo).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(natoms, n_ranks):
    """Contiguous structure ranges [b_r, e_r) with near-equal atom counts.  Every rank gets at
    least one structure when nStruct >= n_ranks."""
    natoms = np.asarray(natoms, np.int64)
    nS = len(natoms)
    cum = np.concatenate([[0], np.cumsum(natoms)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        b = int(np.searchsorted(cum, target, side="left"))
        # nearest boundary, keeping ranges non-empty when possible
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, nS)] - target):
            b -= 1
        b = max(b, bounds[-1] + (1 if nS >= n_ranks else 0))
        b = min(b, nS - (n_ranks - r) if nS >= n_ranks else nS)
        bounds.append(b)
    bounds.append(nS)
    return [(bounds[r], bounds[r + 1]) for r in range(n_ranks)]


def shard_dataset(ds, n_ranks, rank):
    natoms = ds.offsets[1:] - ds.offsets[:-1]
    b, e = shard_bounds(natoms, n_ranks)[rank]
    return ds.select(np.arange(b, e))


def init_comm(ctx, dist=None):
    """Creates the library's NCCL communicator for the current torch.distributed world."""
    import torch
    import torch.distributed as td
    dist = dist or td
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    obj = [ctx.comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx.comm_init(dist.get_world_size(), dist.get_rank(), obj[0])


def allreduce_host(dd, loss_num, loss_den, dist=None):
    """Host-side sum of per-shard (ddSerial, loss numerator, loss denominator) -- the gloo path
    used by the CPU multi-process tests; on GPUs libfnetgpu does this with NCCL itself."""
    import torch
    import torch.distributed as td
    dist = dist or td
    buf = torch.from_numpy(np.concatenate([np.asarray(dd, np.float64).reshape(-1), [loss_num, loss_den]]))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.numpy()
    return out[:-2].reshape(np.shape(dd)), float(out[-2]), float(out[-1])
