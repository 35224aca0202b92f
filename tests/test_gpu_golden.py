"""GPU parity (through the C ABI) against the reference's own golden files and the CPU oracle.
Tolerance: the reference comparator's ATOL 1e-10 / RTOL 1e-9 (FP64 mode)."""
import numpy as np
import pytest

import golden_io as gio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import fortnet_b200 as fb
    return fb


PATHS = ["auto", "cells"]   # whole-structure neighbour search (all fixtures are small structures) and the cell list


def _setup(fb, case, precision=64, acsf_path="auto", mlp="auto"):
    ctx = fb.Context(precision=precision, acsf_path=acsf_path, mlp=mlp)
    ds = case.dataset
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, fb.GFunctions(case.funcs), standardize=case.zmeans is not None,
                   ext_indices=case.ext_indices)
    zp = np.stack([case.zmeans, case.zsigmas]) if case.zmeans is not None else None
    acsf.calculate(0, zprec=zp)
    net = fb.Bpnn(ctx, case.dims, len(case.atomic_numbers), case.activation)
    net.set_params(case.wb())
    return ctx, acsf, net


PRED = gio.cases(mode=("predict", "validate"), forces=(None, "analytical"))


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("entry", PRED, ids=[e["case"] for e in PRED])
def test_predictions_and_forces(fb, entry, path):
    case = gio.Case(entry)
    ctx, acsf, net = _setup(fb, case, acsf_path=path)
    if case.funcs:
        assert ctx.acsf_path(0) in ((2,) if path == "auto" else (0, 1)), ctx.acsf_path(0)
    ds = case.dataset
    raw = net.predict_batch(0)
    assert gio.allclose(raw, case.arr["out_rawpredictions"]), gio.maxdiff(raw, case.arr["out_rawpredictions"])
    if case.nG and "out_globalpredictions" in case.arr:
        glob = np.add.reduceat(raw[:, :case.nG], ds.offsets[:-1].astype(int), axis=0)
        assert gio.allclose(glob, case.arr["out_globalpredictions"])
    if case.forces == "analytical":
        f = net.forces(0)
        assert gio.allclose(f, case.arr["out_forces"]), gio.maxdiff(f, case.arr["out_forces"])
    ctx.close()


SD = gio.cases(mode=("train",), training=("sd", "cg", "lbfgs", "fire"))   # the first iteration of each optimiser (golden_io)


def _update(case, kind, wb0, dd):
    return {"fire": case.fire_update, "cg": case.cg_update, "lbfgs": case.lbfgs_update}.get(
        kind, lambda w, d: case.sd_update(w, d)[0])(wb0, dd)


@pytest.mark.parametrize("mlp", ["auto", "nofuse", "legacy"])   # DMMA kernels (+ fused per-structure sums where they apply) / register-tiled kernels
@pytest.mark.parametrize("entry", SD, ids=[e["case"] for e in SD])
def test_sd_training_step(fb, entry, mlp):
    from oracle import oracle as orc
    case = gio.Case(entry)
    ctx, acsf, net = _setup(fb, case, mlp=mlp)
    assert ctx.mlp_path() == (0 if mlp == "legacy" else 1)
    ds = case.dataset
    wb0 = case.wb()
    dd, loss = net.update_gradients(0, loss=case.loss_name())
    wb1 = _update(case, entry["training"], wb0, dd)
    ref = case.wb("ref_")
    assert gio.allclose(wb1, ref), gio.maxdiff(wb1, ref)
    # loss value and raw gradient against the oracle on the same features
    feats = acsf.features(0)
    dd_o, raw_o = orc.grad(ds.offsets, feats, ds.globalsp, case.dims, case.activation, wb0,
                           case.loss_name(), ds.weights, ds.atomic_weights, ds.gtargets, ds.atargets)
    assert np.allclose(dd, dd_o, rtol=1e-9, atol=1e-10), gio.maxdiff(dd, dd_o)
    loss_o = orc.loss(ds.offsets, raw_o, case.loss_name(), case.nG, case.nA, ds.gtargets, ds.atargets,
                      ds.atomic_weights, ds.weights)
    assert abs(loss - loss_o) <= 1e-10 + 1e-9 * abs(loss_o), (loss, loss_o)
    ctx.close()


ZS = [e for e in gio.cases(mode=("train",))]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("entry", ZS, ids=[e["case"] for e in ZS])
def test_features_and_zscore_statistics(fb, entry, path):
    """raw ACSF vs oracle, and statistics computed on the GPU vs the stored netstat values."""
    from oracle import oracle as orc
    case = gio.Case(entry)
    if not case.funcs:
        pytest.skip("no ACSF in this case")
    ds = case.dataset
    ctx = fb.Context(acsf_path=path)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, fb.GFunctions(case.funcs), standardize=False)
    acsf.calculate(0)
    vals = acsf.features(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, case.funcs, ext=ds.ext)
    assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), gio.maxdiff(vals, ref)
    if case.zmeans is not None and case.name != "input/weighting/datapoints/restart/multispecies/globalTargets/acsfPrec":
        acsf2 = fb.Acsf(ctx, fb.GFunctions(case.funcs), standardize=True)
        acsf2.calculate(0)
        assert gio.allclose(acsf2.zprec[0], case.zmeans), gio.maxdiff(acsf2.zprec[0], case.zmeans)
        assert gio.allclose(acsf2.zprec[1], case.zsigmas), gio.maxdiff(acsf2.zprec[1], case.zsigmas)
        z = acsf2.features(0)
        zo = orc.zscore_apply(ref, *orc.zscore_stats(ds.offsets, ref, ds.weights))
        assert np.allclose(z, zo, rtol=1e-9, atol=1e-10), gio.maxdiff(z, zo)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# fresh-training goldens (RANLUX initialisation restated in tests/ranlux.py)
# ---------------------------------------------------------------------------------------------
import json as _json
import os as _os

with open(_os.path.join(gio.GOLD, "index_fresh.json")) as _fh:
    _FRESH = [e for e in _json.load(_fh) if e["niterations"] == 1]


@pytest.mark.parametrize("entry", _FRESH, ids=[e["case"] for e in _FRESH])
def test_fresh_sd_training_step(fb, entry):
    """seed -> theta0 (harness) -> ACSF + statistics + gradient on the GPU -> SD step -> golden _fortnet.hdf5"""
    import ranlux
    from test_oracle_fresh_training import FreshCase
    case = FreshCase(entry)
    ds = case.dataset
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, fb.GFunctions(case.funcs), standardize=case.zmeans is not None, ext_indices=case.ext_indices)
    acsf.calculate(0)                                   # statistics from the training set, as the reference does
    if case.zmeans is not None:
        assert gio.allclose(acsf.zprec[0], case.zmeans) and gio.allclose(acsf.zprec[1], case.zsigmas)
    nsp = len(case.atomic_numbers)
    net = fb.Bpnn(ctx, case.dims, nsp, case.activation)
    wb0 = ranlux.initial_parameters(case.seed, case.dims, nsp)
    net.set_params(wb0)
    dd, _loss = net.update_gradients(0, loss=case.loss_name())
    wb1 = _update(case, entry["training"], wb0, dd)
    ref = case.wb("ref_")
    nW = case.n_weights()
    keep = np.ones(wb1.shape[1], bool)
    keep[nW - int(case.dims[-1]):nW] = False           # the unused ww(d_L, 1) is not in the netstat file
    assert gio.allclose(wb1[:, keep], ref[:, keep]), gio.maxdiff(wb1[:, keep], ref[:, keep])
    ctx.close()


_REGU = [e for e in _FRESH if "lossregu" in e["case"]]


@pytest.mark.parametrize("entry", _REGU, ids=[e["case"] for e in _REGU])
def test_fresh_training_step_regularized_on_device(fb, entry):
    """input/lossregu/{ridge,lasso,elasticnet} with TBpnn_update's gradient post-processing on the GPU
    (fnetgpu_regularization_set): elastic-net term + division by sum(weights) applied to the reduced gradient on the
    device, optimiser step on the host -> golden theta_1; reguLoss against loss.F90:119-196 evaluated in numpy"""
    import ranlux
    from test_oracle_fresh_training import FreshCase
    case = FreshCase(entry)
    ds = case.dataset
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, fb.GFunctions(case.funcs), standardize=case.zmeans is not None, ext_indices=case.ext_indices)
    acsf.calculate(0)
    nsp = len(case.atomic_numbers)
    net = fb.Bpnn(ctx, case.dims, nsp, case.activation)
    wb0 = ranlux.initial_parameters(case.seed, case.dims, nsp)
    net.set_params(wb0)
    lam, alpha = case.regularization()
    assert lam > 0.0 and entry["training"] == "sd"
    net.set_regularization(lam, alpha, float(np.sum(ds.weights)))
    dd, _loss = net.update_gradients(0, loss=case.loss_name())
    wb1 = case.sd_update(wb0, dd, pre_regularized=True)[0]
    ref = case.wb("ref_")
    nW = case.n_weights()
    keep = np.ones(wb1.shape[1], bool)
    keep[nW - int(case.dims[-1]):nW] = False
    assert gio.allclose(wb1[:, keep], ref[:, keep]), gio.maxdiff(wb1[:, keep], ref[:, keep])
    w = wb0[:, :nW]
    want = lam / nW * ((1.0 - alpha) / 2.0 * (w ** 2).sum(1) + alpha * np.abs(w).sum(1))
    assert np.allclose(net.regularization_loss(), want, rtol=1e-12, atol=0.0)
    net.set_regularization(0.0, 0.0, 0.0)                  # back to the plain summed gradient
    dd_plain, _ = net.update_gradients(0, loss=case.loss_name())
    wb1b = case.sd_update(wb0, dd_plain)[0]
    assert gio.allclose(wb1b[:, keep], ref[:, keep])
    ctx.close()
