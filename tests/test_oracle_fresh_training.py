"""Replays the reference's FRESH-training goldens (test/prog/fortnet/training/**, input/** without a
stored netstat) on the CPU oracle: RANLUX seed -> truncated-normal Xavier weights
(tests/ranlux.py) -> ACSF (+ z-score statistics of the training set) -> gradient -> one steepest-descent
step -> the golden _fortnet.hdf5, at the reference comparator's tolerance (ATOL 1e-10 / RTOL 1e-9).
This pins the driver-side initialisation restated in the harness and, once more, the whole hot path.
CPU only."""
import json
import os

import numpy as np
import pytest

import golden_io as gio
import ranlux
from oracle import oracle as orc

with open(os.path.join(gio.GOLD, "index_fresh.json")) as _fh:
    FRESH = json.load(_fh)
SD1 = [e for e in FRESH if e["training"] in ("sd", "fire", "cg", "lbfgs") and e["niterations"] == 1]


class FreshCase(gio.Case):
    """a golden case without input netstat: topology, functions and statistics come from the golden
    OUTPUT netstat (they do not depend on the weights), the initial weights from the RANLUX seed"""

    def __init__(self, entry):
        self.entry = entry
        self.name = entry["case"]
        z = np.load(os.path.join(gio.GOLD, "fresh", entry["file"]))
        self.arr = {k: z[k] for k in z.files}
        self.meta = json.loads(bytes(self.arr.pop("meta")).decode())
        ns = self.meta["netstat"]
        self.mode = "train"
        self.dims = self.arr["ref_dims"].astype(np.int32)
        self.activation = ns["activation"]
        self.nG, self.nA = ns["nglobaltargets"], ns["natomictargets"]
        self.funcs = ns.get("functions", [])
        self.atomic_numbers = self.arr["ref_atomicnumbers"]
        self.zmeans = self.arr.get("ref_zmeans")
        self.zsigmas = self.arr.get("ref_zsigmas")
        self.ext_indices = self.arr.get("ref_extindices")
        self.dataset = gio.load_dataset(self.meta["dataset"])
        self.training = self.meta.get("training") or {}
        self.forces = None
        self.seed = int(self.meta["seed"])


def test_ranlux_stream_properties():
    """the generator restated in tests/ranlux.py: 24-bit mantissas in (0, 1), reproducible, and the
    luxury level changes the stream only after the first 24 numbers (ranlux.F90:355-368)"""
    a = ranlux.Ranlux(3, 123456).random(100)
    b = ranlux.Ranlux(3, 123456).random(100)
    c = ranlux.Ranlux(0, 123456).random(100)
    assert np.array_equal(a, b) and np.all((a > 0.0) & (a < 1.0))
    assert np.array_equal(a[:24], c[:24]) and not np.array_equal(a[24:], c[24:])
    assert np.all(a * 2 ** 24 == np.round(a * 2 ** 24)) or np.any(a < 2.0 ** -12)


@pytest.mark.parametrize("entry", SD1, ids=[e["case"] for e in SD1])
def test_fresh_sd_training_step(entry):
    case = FreshCase(entry)
    ds = case.dataset
    vals = None
    if case.funcs:
        vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, case.funcs, ext=ds.ext)
        if case.zmeans is not None:
            mu, sg = orc.zscore_stats(ds.offsets, vals, ds.weights)        # computed from the training set (acsf.F90:576-590)
            assert gio.allclose(mu, case.zmeans) and gio.allclose(sg, case.zsigmas)
            vals = orc.zscore_apply(vals, mu, sg)
    feats = case.assemble_features(vals)
    nsp = len(case.atomic_numbers)
    wb0 = ranlux.initial_parameters(case.seed, case.dims, nsp)
    dd, _raw = orc.grad(ds.offsets, feats, ds.globalsp, case.dims, case.activation, wb0, case.loss_name(),
                        ds.weights, ds.atomic_weights, ds.gtargets, ds.atargets)
    wb1 = {"fire": case.fire_update, "cg": case.cg_update, "lbfgs": case.lbfgs_update}.get(
        entry["training"], lambda w, d: case.sd_update(w, d)[0])(wb0, dd)
    ref = case.wb("ref_")
    # the unused last-layer array ww(d_L, 1) is drawn (and keeps its random values) but is not part of
    # the netstat file: compare everything else
    nW = case.n_weights()
    keep = np.ones(wb1.shape[1], bool)
    keep[nW - int(case.dims[-1]):nW] = False
    assert gio.allclose(wb1[:, keep], ref[:, keep]), gio.maxdiff(wb1[:, keep], ref[:, keep])
