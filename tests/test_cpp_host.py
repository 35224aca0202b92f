"""The C++ host mirror (fortnet_b200/host/fnetgpu.hpp) driven in the Fortran driver's call order,
checked against the CPU oracle.  The build of the driver is also checked on CPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_driver.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "host_driver")
    libdir = os.path.join(ROOT, "fortnet_b200")
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", SRC, "-o", exe, "-L" + libdir, "-lfnetgpu",
                           "-Wl,-rpath," + libdir], env=env)
    return exe


def test_cpp_host_builds(tmp_path):
    assert os.path.exists(_build(tmp_path))


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
