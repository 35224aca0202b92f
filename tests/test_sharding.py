"""Multi-process path: atom-balanced structure sharding + gradient/loss reduction.
CPU: world_size 2 and 3 over gloo (oracle as the per-shard compute).
GPU: world_size 2 over NCCL through libfnetgpu (skipped on single-GPU boxes)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from fortnet_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_dist_worker.py")


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    for n_ranks in (1, 2, 3, 8):
        natoms = rng.integers(2, 200, size=57)
        b = sharding.shard_bounds(natoms, n_ranks)
        assert b[0][0] == 0 and b[-1][1] == len(natoms)
        assert all(b[r][1] == b[r + 1][0] for r in range(n_ranks - 1))
        assert all(e > s for s, e in b)
        loads = [natoms[s:e].sum() for s, e in b]
        assert max(loads) - min(loads) <= 2 * natoms.max()
    assert sharding.shard_bounds([5, 5, 5], 3) == [(0, 1), (1, 2), (2, 3)]


def _run(mode, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nproc", [2, 3])
def test_gloo_sharded_gradient(nproc):
    _run("cpu", nproc, 29511 + nproc)


@pytest.mark.gpu
def test_nccl_sharded_gradient():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("gpu", 2, 29521)
