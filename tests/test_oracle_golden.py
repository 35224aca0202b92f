"""Pins the CPU oracle (oracle/refcpu.c) against the reference's OWN regression goldens
(test/prog/fortnet/**, comparator tolerance ATOL 1e-10 / RTOL 1e-9): ACSF values (+z-score),
subnet forward, analytic forces, and one steepest-descent training step (gradient scaling,
serialisation order, loss gradient).  CPU only."""
import numpy as np
import pytest

import golden_io as gio
from oracle import oracle as orc


def _features(case):
    ds = case.dataset
    vals = None
    if case.funcs:
        vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, case.funcs, ext=ds.ext)
        if case.zmeans is not None:
            vals = orc.zscore_apply(vals, case.zmeans, case.zsigmas)
    return case.assemble_features(vals)


PRED = gio.cases(mode=("predict", "validate"), forces=(None, "analytical"))


@pytest.mark.parametrize("entry", PRED, ids=[e["case"] for e in PRED])
def test_predictions_and_forces(entry):
    case = gio.Case(entry)
    ds = case.dataset
    feats = _features(case)
    raw = orc.predict(feats, ds.globalsp, case.dims, case.activation, case.wb())
    assert gio.allclose(raw, case.arr["out_rawpredictions"]), gio.maxdiff(raw, case.arr["out_rawpredictions"])
    if case.nG and "out_globalpredictions" in case.arr:
        glob = np.add.reduceat(raw[:, :case.nG], ds.offsets[:-1].astype(int), axis=0)
        assert gio.allclose(glob, case.arr["out_globalpredictions"])
    if case.forces == "analytical":
        f = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, case.funcs, feats,
                       ds.globalsp, case.dims, case.activation, case.wb(), ext=ds.ext,
                       sigmas=case.zsigmas)
        assert gio.allclose(f, case.arr["out_forces"]), gio.maxdiff(f, case.arr["out_forces"])


SD = gio.cases(mode=("train",), training=("sd",))


@pytest.mark.parametrize("entry", SD, ids=[e["case"] for e in SD])
def test_sd_training_step(entry):
    """theta1 = theta0 - lr * g(theta0): fortnet.hdf5 -> _fortnet.hdf5 (bpnn.F90:277,298)."""
    case = gio.Case(entry)
    ds = case.dataset
    feats = _features(case)
    wb0 = case.wb()
    dd, _raw = orc.grad(ds.offsets, feats, ds.globalsp, case.dims, case.activation, wb0,
                        case.loss_name(), ds.weights, ds.atomic_weights, ds.gtargets, ds.atargets)
    wb1, _ = case.sd_update(wb0, dd)
    ref = case.wb("ref_")
    assert gio.allclose(wb1, ref), gio.maxdiff(wb1, ref)
    if "ref_zmeans" in case.arr and case.zmeans is not None:
        assert gio.allclose(case.zmeans, case.arr["ref_zmeans"])


FIRE = [e for e in gio.cases(mode=("train",), training=("fire", "cg", "lbfgs"))]


@pytest.mark.parametrize("entry", FIRE, ids=[e["case"] for e in FIRE])
def test_fire_cg_lbfgs_training_step(entry):
    """first FIRE iteration from rest: theta1 = theta0 - dt^2 g, dt = 0.1 MaxDisplacement (fire.F90:137-185);
    first conjugate-gradient iteration: trial step 5 |g| along -g, capped by MaxDisplacement (conjgrad.F90, linemin.F90);
    first L-BFGS iteration: unit trial step along -g/|g| (lbfgs.F90:283-359)"""
    case = gio.Case(entry)
    if int(case.training.get("niterations", 1)) != 1:
        pytest.skip("only the first iteration is restated")
    ds = case.dataset
    feats = _features(case)
    wb0 = case.wb()
    dd, _raw = orc.grad(ds.offsets, feats, ds.globalsp, case.dims, case.activation, wb0,
                        case.loss_name(), ds.weights, ds.atomic_weights, ds.gtargets, ds.atargets)
    wb1 = {"fire": case.fire_update, "cg": case.cg_update, "lbfgs": case.lbfgs_update}[entry["training"]](wb0, dd)
    ref = case.wb("ref_")
    assert gio.allclose(wb1, ref), gio.maxdiff(wb1, ref)


ZS = gio.cases(mode=("train",))


@pytest.mark.parametrize("entry", ZS, ids=[e["case"] for e in ZS])
def test_zscore_statistics(entry):
    """means / 'variances' (population sigma) stored in the netstat were computed by the
    reference from the same training set + ACSF config (acsf.F90:445-486)."""
    case = gio.Case(entry)
    if case.zmeans is None or not case.funcs:
        pytest.skip("no standardisation in this case")
    if case.name == "input/weighting/datapoints/restart/multispecies/globalTargets/acsfPrec":
        # the INPUT netstat of this one case was not generated from its own dataset (differs
        # by 3e-5 for weights [1..5] and for unit weights alike); its stats are an input of
        # the case, not an output, so there is nothing to pin here.  The weighted formula is
        # pinned by .../restart/singlespecies/globalTargets/acsfPrec.
        pytest.skip("input netstat statistics not derived from this dataset")
    ds = case.dataset
    vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, case.funcs, ext=ds.ext)
    mu, sg = orc.zscore_stats(ds.offsets, vals, ds.weights)
    assert gio.allclose(mu, case.zmeans), gio.maxdiff(mu, case.zmeans)
    assert gio.allclose(sg, case.zsigmas), gio.maxdiff(sg, case.zsigmas)
