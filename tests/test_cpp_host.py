"""The C++ host mirror (fortnet_b200/host/fnetgpu.hpp) driven in the Fortran driver's call order,
checked against the CPU oracle.  The build of the driver is also checked on CPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_driver.cpp")
SRC_MG = os.path.join(ROOT, "tests", "cpp", "host_driver_mg.cpp")


def _build(tmp_path, src=SRC, name="host_driver"):
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "fortnet_b200")
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", src, "-o", exe, "-L" + libdir, "-lfnetgpu",
                           "-Wl,-rpath," + libdir], env=env)
    return exe


def test_cpp_host_builds(tmp_path):
    assert os.path.exists(_build(tmp_path))
    assert os.path.exists(_build(tmp_path, SRC_MG, "host_driver_mg"))


def _write_case(d, fb, orc, ds, funcs, dims, wb):
    t = funcs.tables()
    F = len(funcs)
    np.array([ds.n_global_targets, ds.n_atomic_targets, 0, F, len(ds.atomic_numbers), fb.ACTIVATIONS.index("tanh"), 0, 1,
              len(dims)] + dims, np.int32).tofile(d + "meta.i32")
    ds.offsets.astype(np.int32).tofile(d + "offsets.i32"); ds.coords.tofile(d + "coords.f64")
    ds.periodic.astype(np.int32).tofile(d + "periodic.i32"); ds.latvecs.tofile(d + "lat.f64")
    ds.atnum.astype(np.int32).tofile(d + "atnum.i32"); ds.globalsp.astype(np.int32).tofile(d + "gsp.i32")
    ds.weights.astype(np.int32).tofile(d + "w.i32"); ds.atomic_weights.tofile(d + "aw.f64")
    ds.gtargets.tofile(d + "gt.f64"); ds.atargets.tofile(d + "at.f64"); ds.ext.tofile(d + "ext.f64")
    np.stack([t["ftype"], t["atomid"], t["atomicnumbers"][0::2], t["atomicnumbers"][1::2]], 1).astype(np.int32).tofile(d + "fint.i32")
    np.stack([t["rcut"], t["kappa"], t["rs"], t["eta"], t["lam"], t["xi"]], 1).tofile(d + "fpar.f64")
    wb.tofile(d + "wb.f64")


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [1, 2])
def test_cpp_host_single_process_multi_gpu(tmp_path, ndev):
    """ONE process, ndev GPUs through fnetgpu_mg_* (shards by atom count, ncclCommInitAll, one all-reduce per
    gradient): features, statistics, gradient, loss, predictions and forces of a ragged two-species batch
    against the oracle on the whole dataset"""
    import torch
    import fortnet_b200 as fb
    from fortnet_b200 import synthetic
    from oracle import oracle as orc
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    exe = _build(tmp_path, SRC_MG, "host_driver_mg")
    a = synthetic.tio2(n_struct=3, seed=15)
    rng = np.random.default_rng(19)
    # ragged: three 192-atom cells and four small clusters, so that the atom-balanced split is not the structure-balanced one
    natoms = [192, 192, 192, 5, 9, 2, 17]
    cl = [rng.uniform(0.0, 9.0, size=(n, 3)) for n in natoms[3:]]
    coords = np.concatenate([a.coords] + cl)
    N = sum(natoms)
    atnum = np.concatenate([a.atnum] + [rng.choice([22, 8], size=n) for n in natoms[3:]]).astype(np.int32)
    lat = np.concatenate([a.latvecs.reshape(3, 3, 3), np.zeros((4, 3, 3))])
    ds = fb.Dataset.build(natoms, coords, np.array([1, 1, 1, 0, 0, 0, 0], np.int32), lat, atnum,
                          gtargets=rng.uniform(-2.0, 2.0, size=(7, 1)), weights=rng.integers(1, 3, size=7),
                          atomic_numbers=[22, 8])
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 4, 6).resolve_species([22, 8])
    dims = [len(funcs), 7, 5, 1]
    wb = rng.uniform(-0.5, 0.5, size=(2, orc.ntot(dims)))
    d = str(tmp_path) + "/"
    _write_case(d, fb, orc, ds, funcs, dims, wb)
    r = subprocess.run([exe, d, str(ndev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HOST_MG_OK devices=%d" % ndev in r.stdout, r.stdout + r.stderr
    fd = funcs.asdicts()
    nt = os.cpu_count() or 1
    vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, nthreads=nt)
    mu, sg = orc.zscore_stats(ds.offsets, vals, ds.weights)
    feats = orc.zscore_apply(vals, mu, sg)
    dd, raw = orc.grad(ds.offsets, feats, ds.globalsp, dims, "tanh", wb, "mse", ds.weights, ds.atomic_weights,
                       ds.gtargets, ds.atargets, nthreads=nt)
    loss = orc.loss(ds.offsets, raw, "mse", 1, 0, ds.gtargets, ds.atargets, ds.atomic_weights, ds.weights)
    frc = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, feats, ds.globalsp, dims,
                     "tanh", wb, sigmas=sg, nthreads=nt)
    tol = dict(rtol=1e-9, atol=1e-10)
    assert np.allclose(np.fromfile(d + "out_feats.f64").reshape(feats.shape), feats, **tol)
    assert np.allclose(np.fromfile(d + "out_zprec.f64"), np.concatenate([mu, sg]), **tol)
    assert np.allclose(np.fromfile(d + "out_dd.f64").reshape(dd.shape), dd, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(dd).max()))
    assert np.allclose(np.fromfile(d + "out_loss.f64"), loss, **tol)
    assert np.allclose(np.fromfile(d + "out_gpred.f64"), np.add.reduceat(raw[:, 0], ds.offsets[:-1].astype(int)), **tol)
    assert np.allclose(np.fromfile(d + "out_raw.f64").reshape(raw.shape), raw, **tol)
    assert np.allclose(np.fromfile(d + "out_forces.f64").reshape(frc.shape), frc, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(frc).max()))


@pytest.mark.gpu
def test_cpp_host_matches_oracle(tmp_path):
    import fortnet_b200 as fb
    from fortnet_b200 import synthetic
    from oracle import oracle as orc
    exe = _build(tmp_path)
    ds = synthetic.tio2(n_struct=3, seed=5)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 4, 6).resolve_species([22, 8])
    dims = [len(funcs), 7, 5, 1]
    rng = np.random.default_rng(9)
    wb = rng.uniform(-0.5, 0.5, size=(2, orc.ntot(dims)))
    d = str(tmp_path) + "/"
    t = funcs.tables()
    F = len(funcs)
    np.array([1, 0, 0, F, 2, fb.ACTIVATIONS.index("tanh"), 0, 1, len(dims)] + dims, np.int32).tofile(d + "meta.i32")
    ds.offsets.astype(np.int32).tofile(d + "offsets.i32"); ds.coords.tofile(d + "coords.f64")
    ds.periodic.astype(np.int32).tofile(d + "periodic.i32"); ds.latvecs.tofile(d + "lat.f64")
    ds.atnum.astype(np.int32).tofile(d + "atnum.i32"); ds.globalsp.astype(np.int32).tofile(d + "gsp.i32")
    ds.weights.astype(np.int32).tofile(d + "w.i32"); ds.atomic_weights.tofile(d + "aw.f64")
    ds.gtargets.tofile(d + "gt.f64"); ds.atargets.tofile(d + "at.f64"); ds.ext.tofile(d + "ext.f64")
    np.stack([t["ftype"], t["atomid"], t["atomicnumbers"][0::2], t["atomicnumbers"][1::2]], 1).astype(np.int32).tofile(d + "fint.i32")
    np.stack([t["rcut"], t["kappa"], t["rs"], t["eta"], t["lam"], t["xi"]], 1).tofile(d + "fpar.f64")
    wb.tofile(d + "wb.f64")
    r = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_OK" in r.stdout, r.stdout + r.stderr
    fd = funcs.asdicts()
    vals = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd)
    mu, sg = orc.zscore_stats(ds.offsets, vals, ds.weights)
    feats = orc.zscore_apply(vals, mu, sg)
    dd, raw = orc.grad(ds.offsets, feats, ds.globalsp, dims, "tanh", wb, "mse", ds.weights, ds.atomic_weights,
                       ds.gtargets, ds.atargets)
    frc = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, feats, ds.globalsp, dims,
                     "tanh", wb, sigmas=sg)
    tol = dict(rtol=1e-9, atol=1e-10)
    assert np.allclose(np.fromfile(d + "out_feats.f64").reshape(feats.shape), feats, **tol)
    assert np.allclose(np.fromfile(d + "out_zprec.f64"), np.concatenate([mu, sg]), **tol)
    assert np.allclose(np.fromfile(d + "out_dd.f64").reshape(dd.shape), dd, **tol)
    assert np.allclose(np.fromfile(d + "out_raw.f64").reshape(raw.shape), raw, **tol)
    assert np.allclose(np.fromfile(d + "out_forces.f64").reshape(frc.shape), frc, **tol)
