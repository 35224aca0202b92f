"""GPU parity (through the C ABI) against the CPU oracle on synthetic inputs of the BASELINE.json
shapes, the edge cases the reference handles (clusters, single atoms, ragged batches, thin and
triclinic cells, unfolded coordinates), and size-independent properties at full configuration size.

FP64 mode: the reference comparator's RTOL 1e-9 / ATOL 1e-10 (bin/testwithworkdir.py:24-25), raw
ACSF values at RTOL 1e-10.  FP32 mode: the bound documented in DESIGN.md section 4.5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-10


@pytest.fixture(scope="module")
def fb():
    import fortnet_b200 as fb
    return fb


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _nthreads():
    import os
    return os.cpu_count() or 1


def _ntot(dims):
    return sum(a * b for a, b in zip(dims[:-1], dims[1:])) + dims[-1] + sum(dims)


def _md(a, b):
    return "max abs diff %.3e (scale %.3e)" % (np.abs(a - b).max(), np.abs(b).max())


PATHS = ["auto", "cells"]      # neighbour search: whole-structure / minimum image where valid, or always the cell list
STRUCT, CELLS = (2,), (0, 1)   # values of Context.acsf_path()


def _full_check(fb, orc, ds, funcs, dims, act="tanh", loss="mse", forces=True, seed=3, acsf_path="auto",
                expect_path=None, mlp="auto", expect_mlp=None, acsf_kernel="auto", expect_lean=None, expect_fusion=None):
    """features (raw + z-scored), statistics, predictions, loss, gradient and forces vs the oracle"""
    fd = funcs.asdicts()
    nt = _nthreads()
    ctx = fb.Context(acsf_path=acsf_path, mlp=mlp, acsf_kernel=acsf_kernel)
    ctx.upload(0, ds)
    a_raw = fb.Acsf(ctx, funcs, standardize=False)
    a_raw.calculate(0)
    if acsf_path == "cells":
        expect_path = CELLS
    if expect_path is not None:
        assert ctx.acsf_path(0) in expect_path, ctx.acsf_path(0)
    if acsf_kernel == "generic":
        expect_lean = 0
    if expect_lean is not None:
        assert ctx.acsf_kernel() == expect_lean, ctx.acsf_kernel()
    vals = a_raw.features(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, ext=ds.ext, nthreads=nt)
    assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), _md(vals, ref)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    mu, sg = orc.zscore_stats(ds.offsets, ref, ds.weights)
    assert np.allclose(acsf.zprec[0], mu, rtol=RTOL, atol=ATOL), _md(acsf.zprec[0], mu)
    assert np.allclose(acsf.zprec[1], sg, rtol=RTOL, atol=ATOL), _md(acsf.zprec[1], sg)
    zref = orc.zscore_apply(ref, mu, sg)
    z = acsf.features(0)
    assert np.allclose(z, zref, rtol=RTOL, atol=ATOL), _md(z, zref)
    nsp = len(ds.atomic_numbers)
    net = fb.Bpnn(ctx, dims, nsp, act)
    wb = np.random.default_rng(seed).uniform(-0.5, 0.5, size=(nsp, _ntot(dims)))
    net.set_params(wb)
    if mlp == "legacy":
        expect_mlp = 0
    elif mlp == "nofuse" and expect_mlp is not None:
        expect_mlp = min(expect_mlp, 1)
    if expect_mlp is not None:
        assert ctx.mlp_path() == expect_mlp, ctx.mlp_path()
    raw = net.predict_batch(0)
    raw_o = orc.predict(zref, ds.globalsp, dims, act, wb, nthreads=nt)
    assert np.allclose(raw, raw_o, rtol=RTOL, atol=ATOL), _md(raw, raw_o)
    dd, lossv, gpred = net.update_gradients(0, loss, want_global=True)
    if expect_fusion is not None:     # 0 two passes, 1 sums fused into the gradient kernel, 2 fused across a thread-block cluster
        gi = ctx.grad_launch_info(0)
        assert gi["fusion"] == (expect_fusion if mlp == "auto" else 0), gi
    dd_o, raw_o2 = orc.grad(ds.offsets, zref, ds.globalsp, dims, act, wb, loss, ds.weights, ds.atomic_weights,
                            ds.gtargets, ds.atargets, nthreads=nt)
    assert np.allclose(dd, dd_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(dd_o).max())), _md(dd, dd_o)
    loss_o = orc.loss(ds.offsets, raw_o2, loss, ds.n_global_targets, ds.n_atomic_targets, ds.gtargets, ds.atargets,
                      ds.atomic_weights, ds.weights)
    assert abs(lossv - loss_o) <= ATOL + RTOL * abs(loss_o), (lossv, loss_o)
    assert abs(net.loss(0, loss) - loss_o) <= ATOL + RTOL * abs(loss_o)
    nG = ds.n_global_targets
    if nG:   # globalPredictions of updateGradients (fnetout.F90:116-117): per-structure sums of the atomic outputs
        g_o = np.add.reduceat(raw_o2[:, :nG], ds.offsets[:-1].astype(int), axis=0)
        assert np.allclose(gpred, g_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(g_o).max())), _md(gpred, g_o)
    if forces:
        f = net.forces(0)
        f_o = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, zref, ds.globalsp, dims, act,
                         wb, ext=ds.ext, sigmas=sg, nthreads=nt)
        assert np.allclose(f, f_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(f_o).max())), _md(f, f_o)
        if expect_path is not None:
            assert ctx.acsf_path(0) in expect_path, ctx.acsf_path(0)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs at oracle-sized samples
# ---------------------------------------------------------------------------------------------
MLPS = ["auto", "nofuse", "legacy"]   # precision 64 subnetworks: DMMA kernels (mlp_mma.cuh; "auto" fuses the per-structure
                                      # sums into the gradient kernel for single-species data), or the register-tiled ones


@pytest.mark.parametrize("mlp", MLPS)
@pytest.mark.parametrize("path", PATHS)
def test_c2_si_bulk(fb, orc, path, mlp):
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=12, seed=20260001)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
    _full_check(fb, orc, ds, funcs, [32, 20, 20, 1], acsf_path=path, expect_path=STRUCT, mlp=mlp, expect_mlp=1,
                expect_lean=1, expect_fusion=1)


KERNELS = ["auto", "generic"]   # ACSF value kernel: k_acsf_lean for automatic-scheme configurations, or always k_acsf


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("wl", ["c2", "c3", "c3_unresolved", "cluster"])
def test_value_kernels(fb, orc, wl, path, kernel):
    """both ACSF value kernels on the BASELINE shapes (and a ragged cluster batch) through both neighbour paths"""
    from fortnet_b200 import synthetic
    if wl == "c2":
        ds = synthetic.si_bulk(n_struct=5, seed=77)
        funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
    elif wl == "c3":
        ds = synthetic.tio2(n_struct=2, seed=78)
        funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8])
    elif wl == "c3_unresolved":
        ds = synthetic.tio2(n_struct=2, seed=79)
        funcs = fb.GFunctions.from_auto_scheme(5.0 * fb.BOHR_PER_AA, 12, 24)
    else:
        rng = np.random.default_rng(80)
        natoms = [1, 2, 40, 9, 33]
        coords = np.concatenate([rng.uniform(0.0, 6.0 + 2.0 * n ** (1 / 3), size=(n, 3)) for n in natoms])
        N = sum(natoms)
        ds = fb.Dataset.build(natoms, coords, np.zeros(len(natoms), np.int32), np.zeros((len(natoms), 3, 3)),
                              rng.choice([1, 8], size=N).astype(np.int32), gtargets=np.zeros((len(natoms), 1)),
                              atomic_numbers=[1, 8])
        funcs = fb.GFunctions.from_auto_scheme(6.0, 7, 10).resolve_species([1, 8])
    ctx = fb.Context(acsf_path=path, acsf_kernel=kernel)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False)
    assert ctx.acsf_kernel() == (1 if kernel == "auto" else 0)
    acsf.calculate(0)
    vals = acsf.features(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=_nthreads())
    assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), _md(vals, ref)
    acsf.calculate(0)                                    # second launch: capacities (and atoms per warp) from the first one
    assert np.allclose(acsf.features(0), vals, rtol=1e-13, atol=1e-300)
    ctx.close()


@pytest.mark.parametrize("mlp", MLPS)
@pytest.mark.parametrize("path", PATHS)
def test_c3_tio2(fb, orc, path, mlp):
    from fortnet_b200 import synthetic
    ds = synthetic.tio2(n_struct=3, seed=20260002)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8])
    assert len(funcs) == 64
    _full_check(fb, orc, ds, funcs, [64, 32, 32, 32, 1], acsf_path=path, expect_path=STRUCT, mlp=mlp, expect_mlp=1,
                expect_fusion=2)


def _ragged_two_species(fb, seed, natoms, n_targets):
    rng = np.random.default_rng(seed)
    rc = 3.5 * fb.BOHR_PER_AA
    coords = np.concatenate([rng.uniform(0.0, 2.0 * rc + 0.4 * rc * n ** (1 / 3), size=(n, 3)) for n in natoms])
    N = sum(natoms)
    atnum = rng.choice([22, 8, 8], size=N).astype(np.int32)
    atnum[0], atnum[1] = 22, 8
    ds = fb.Dataset.build(natoms, coords, np.zeros(len(natoms), np.int32), np.zeros((len(natoms), 3, 3)), atnum,
                          gtargets=rng.uniform(1.0, 2.0, size=(len(natoms), n_targets)),
                          weights=rng.integers(1, 4, size=len(natoms)), atomic_weights=rng.uniform(0.5, 1.5, size=N),
                          atomic_numbers=[22, 8])
    return ds, rc


@pytest.mark.parametrize("loss", ["mse", "rms", "mae", "mape"])
@pytest.mark.parametrize("mlp", ["auto", "nofuse"])
def test_cluster_fused_structure_sums_ragged_two_species(fb, orc, mlp, loss, monkeypatch):
    """two-species structures of 1..150 atoms (several small ones per super-round, structures whose species spans
    two rounds, single atoms of one species only), two global targets, dataset and atomic weights: the per-structure
    sums exchanged between the CTAs of a thread-block cluster (k_bpnn_mma<0,..,2>) against the oracle and against the
    two-pass kernels"""
    natoms = [1, 7, 64, 3, 150, 30, 5, 97, 63, 2, 17, 40, 24, 1, 1, 9, 130, 12]
    ds, rc = _ragged_two_species(fb, 52, natoms, 2)
    monkeypatch.setenv("FNETGPU_MLP_CLUSTER", "4")      # tiny batch: do not leave the choice to the cost estimate
    funcs = _mixed_functions(fb, rc)
    _full_check(fb, orc, ds, funcs, [len(funcs), 6, 5, 2], loss=loss, forces=False, mlp=mlp, expect_mlp=1, expect_fusion=2)


@pytest.mark.parametrize("cs", [2, 3, 4, 5, 8])
def test_cluster_fused_every_cluster_size(fb, cs, monkeypatch):
    """the cluster-fused gradient with the cluster size pinned to 2..8 (FNETGPU_MLP_CLUSTER) is the two-pass gradient:
    same loss, same per-structure energies, gradient to summation-order accuracy; repeat launches are bit-identical"""
    natoms = [40, 3, 64, 17, 64, 1, 60, 25, 25, 25, 9] if cs == 2 else [40, 3, 96, 17, 64, 1, 80, 25, 25, 25, 128, 9]
    ds, rc = _ragged_two_species(fb, 53, natoms, 1)
    funcs = fb.GFunctions.from_auto_scheme(rc, 5, 4).resolve_species([22, 8])
    dims = [len(funcs), 24, 16, 1]
    wb = np.random.default_rng(9).uniform(-0.5, 0.5, size=(2, _ntot(dims)))
    res = {}
    for mode in ("nofuse", "auto"):
        if mode == "auto":
            monkeypatch.setenv("FNETGPU_MLP_CLUSTER", str(cs))
        ctx = fb.Context(mlp=mode)
        ctx.upload(0, ds)
        fb.Acsf(ctx, funcs, standardize=True).calculate(0)
        net = fb.Bpnn(ctx, dims, 2, "tanh")
        net.set_params(wb)
        out = net.update_gradients(0, "mse", want_global=True)
        gi = ctx.grad_launch_info(0)
        if mode == "auto":
            assert gi["fusion"] == 2 and gi["cluster_size"] == cs and gi["grid"] % cs == 0, gi
            again = net.update_gradients(0, "mse", want_global=True)
            assert np.array_equal(out[0], again[0]) and out[1] == again[1] and np.array_equal(out[2], again[2])
        else:
            assert gi["fusion"] == 0, gi
        res[mode] = out
        ctx.close()
    (dd0, l0, g0), (dd1, l1, g1) = res["nofuse"], res["auto"]
    assert np.allclose(dd1, dd0, rtol=1e-11, atol=1e-12 * np.abs(dd0).max()), _md(dd1, dd0)
    assert abs(l1 - l0) <= 1e-12 * abs(l0)
    assert np.allclose(g1, g0, rtol=1e-12, atol=1e-13)


def test_cluster_fused_falls_back_for_atomic_targets_and_huge_structures(fb):
    """atomic targets, or a structure that needs more than 8 rounds of 64 atoms, keep the two-pass path"""
    from fortnet_b200 import synthetic
    big = synthetic.dense_liquid(n_atoms=600, density_aa3=0.05, seed=5, n_struct=1)
    ds = fb.Dataset.build([600], big.coords, np.ones(1, np.int32), big.latvecs, np.full(600, 14, np.int32),
                          gtargets=np.array([[1.0]]), atomic_numbers=[14])
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 4, 4)
    ctx = fb.Context()
    ctx.upload(0, ds)
    fb.Acsf(ctx, funcs, standardize=True).calculate(0)
    net = fb.Bpnn(ctx, [8, 6, 1], 1, "tanh")
    net.set_params(np.random.default_rng(1).uniform(-0.5, 0.5, size=(1, _ntot([8, 6, 1]))))
    net.update_gradients(0, "mse")
    assert ctx.grad_launch_info(0)["fusion"] == 0
    ctx.close()


@pytest.mark.parametrize("loss", ["mse", "rms", "mae", "mape"])
@pytest.mark.parametrize("mlp", ["auto", "nofuse"])
def test_fused_structure_sums_ragged_single_species(fb, orc, mlp, loss):
    """single-species structures of 1..64 atoms (several per 64-atom round, ragged last tiles), two global
    targets, dataset and atomic weights: the gradient kernel's fused per-structure sums / loss terms
    against the oracle and against the unfused kernels"""
    rng = np.random.default_rng(51)
    natoms = [1, 7, 64, 3, 30, 30, 5, 64, 63, 2, 17, 40, 24, 1, 1, 9]
    rc = 3.5 * fb.BOHR_PER_AA
    coords = np.concatenate([rng.uniform(0.0, 2.0 * rc + 0.4 * rc * n ** (1 / 3), size=(n, 3)) for n in natoms])
    N = sum(natoms)
    ds = fb.Dataset.build(natoms, coords, np.zeros(len(natoms), np.int32), np.zeros((len(natoms), 3, 3)),
                          np.full(N, 14, np.int32), gtargets=rng.uniform(1.0, 2.0, size=(len(natoms), 2)),
                          weights=rng.integers(1, 4, size=len(natoms)), atomic_weights=rng.uniform(0.5, 1.5, size=N),
                          atomic_numbers=[14])
    funcs = _mixed_functions(fb, rc)
    _full_check(fb, orc, ds, funcs, [len(funcs), 6, 5, 2], loss=loss, forces=False, mlp=mlp, expect_mlp=1)


def test_large_and_small_structures_in_one_batch(fb, orc):
    """a 300-atom cell next to small ones: the batch exceeds the whole-structure path's 256-atom limit, so
    the automatic choice must be the cell list for the whole slot; values, gradient and forces vs oracle"""
    from fortnet_b200 import synthetic
    big = synthetic.dense_liquid(n_atoms=300, density_aa3=0.05, seed=5, n_struct=1)      # L = 18.2 A >= 2 rc
    small = synthetic.si_bulk(n_struct=2, seed=6)
    ds = fb.Dataset.build([300, 64, 64], np.concatenate([big.coords, small.coords]), np.ones(3, np.int32),
                          np.concatenate([big.latvecs, small.latvecs]), np.full(428, 14, np.int32),
                          gtargets=np.array([[1.0], [2.0], [3.0]]), atomic_numbers=[14])
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 6, 6)
    _full_check(fb, orc, ds, funcs, [12, 6, 1], expect_path=CELLS)


def _force_workload(fb, wl):
    from fortnet_b200 import synthetic
    if wl == "c2":
        return synthetic.si_bulk(n_struct=4, seed=91), fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16), [32, 10, 2]
    if wl == "c3":
        return (synthetic.tio2(n_struct=2, seed=92),
                fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8]), [64, 12, 1])
    if wl == "long_ladder":     # 36 xi per lambda: two chained blocks of 32 / 4 functions (start power q^32), odd neighbour counts
        return synthetic.si_bulk(n_struct=2, seed=93), fb.GFunctions.from_auto_scheme(4.3 * fb.BOHR_PER_AA, 2, 72), [74, 6, 1]
    if wl == "mid_ladder":      # 9 xi per lambda: 2 x 2 chained slots
        return synthetic.tio2(n_struct=1, seed=94), fb.GFunctions.from_auto_scheme(3.6 * fb.BOHR_PER_AA, 17, 18), [35, 6, 1]
    rng = np.random.default_rng(95)
    natoms = [1, 2, 40, 9, 33, 3]
    coords = np.concatenate([rng.uniform(0.0, 6.0 + 2.0 * n ** (1 / 3), size=(n, 3)) for n in natoms])
    N = sum(natoms)
    ds = fb.Dataset.build(natoms, coords, np.zeros(len(natoms), np.int32), np.zeros((len(natoms), 3, 3)),
                          rng.choice([1, 8], size=N).astype(np.int32), gtargets=np.zeros((len(natoms), 1)),
                          atomic_numbers=[1, 8])
    funcs = fb.GFunctions.from_auto_scheme(6.0, 7, 10).resolve_species([1, 8])
    return ds, funcs, [len(funcs), 8, 1]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("wl", ["c2", "c3", "long_ladder", "mid_ladder", "cluster"])
def test_force_kernels(fb, orc, wl, path, kernel):
    """both force kernels (k_acsf_force_lean for automatic-scheme configurations, k_acsf_force) through both
    neighbour paths against the oracle's dense-tensor forces, z-scored features, one or two outputs"""
    ds, funcs, dims = _force_workload(fb, wl)
    nt = _nthreads()
    ctx = fb.Context(acsf_path=path, acsf_kernel=kernel)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    nsp = len(ds.atomic_numbers)
    net = fb.Bpnn(ctx, dims, nsp, "tanh")
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(nsp, _ntot(dims)))
    net.set_params(wb)
    f = net.forces(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=nt)
    mu, sg = orc.zscore_stats(ds.offsets, ref, ds.weights)
    zref = orc.zscore_apply(ref, mu, sg)
    f_o = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), zref, ds.globalsp, dims, "tanh",
                     wb, sigmas=sg, nthreads=nt)
    assert np.allclose(f, f_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(f_o).max())), _md(f, f_o)
    f2 = net.forces(0)
    if kernel == "auto" and path == "auto":
        assert np.array_equal(f, f2), "forces of the whole-structure path must be bit-reproducible"
    ctx.close()


def test_forces_bit_reproducible_full_size(fb):
    """10^5 atoms of C2 shape, two contexts: the deterministic force path gives identical bits"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=1500, seed=96)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
    dims = [32, 20, 20, 1]
    wb = np.random.default_rng(8).uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    out = []
    for rep in range(2):
        ctx = fb.Context()
        ctx.upload(0, ds)
        acsf = fb.Acsf(ctx, funcs, standardize=True)
        acsf.calculate(0)
        net = fb.Bpnn(ctx, dims, 1, "tanh")
        net.set_params(wb)
        out.append(net.forces(0))
        out.append(net.forces(0))
        ctx.close()
    for f in out[1:]:
        assert np.array_equal(out[0], f)
    assert np.abs(out[0].reshape(1500, 64, 3).sum(1)).max() < 1e-9 * np.abs(out[0]).max() * 64   # sum rule per structure


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("nrad,nang", [(24, 4), (3, 40), (17, 18), (32, 34), (9, 2), (2, 72), (5, 8)])
def test_auto_scheme_sizes(fb, orc, nrad, nang, path, kernel):
    """auto-scheme sizes that exercise every shape of the kernels' function grouping: 1 / 2 / 4 radial
    chunks of 8 (shared-memory reduction with 8 / 16 / 32 rows), 1 / 2 / 4 ladder slots per angular
    pass, continuing ladders (more than 8 functions per lambda) and partially filled ladders"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=2, seed=13)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, nrad, nang)
    ctx = fb.Context(acsf_path=path, acsf_kernel=kernel)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False)
    # nang = 4: xi step 15 -> the degree-5 power series is not accurate enough, k_acsf takes it
    assert ctx.acsf_kernel() == (1 if kernel == "auto" and nang != 4 else 0)
    acsf.calculate(0)
    vals = acsf.features(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=_nthreads())
    assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), _md(vals, ref)
    ctx.close()


@pytest.mark.parametrize("dims,expect", [([9, 1], 1), ([9, 3, 1], 1), ([9, 17, 9, 33, 2, 1], 1), ([9, 40, 40, 40, 1], 1),
                                         ([9, 100, 100, 1], 0), ([9, 70, 60, 1], 0)])
def test_subnetwork_shapes(fb, orc, dims, expect):
    """layer widths that are not multiples of the 8x8x4 DMMA tile, a network without hidden layer,
    deep / narrow ones, and widths beyond the DMMA path's limits (-> register-tiled kernels);
    ragged tiles (structures of 64 atoms, 3 structures = 192 atoms = 12 tiles of 16)"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=3, seed=12)
    ds.gtargets[:] = np.random.default_rng(6).uniform(1.0, 2.0, size=ds.gtargets.shape)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 5, 4)
    _full_check(fb, orc, ds, funcs, dims, forces=True, expect_mlp=expect)


@pytest.mark.parametrize("kernel", KERNELS)
def test_c5_dense_liquid_values(fb, orc, kernel):
    """C5 shape at oracle size: rc = 8 A, ~150 neighbours, 128 G5 on the auto ladder; box edge
    (12.2 A) < 2 rc, so atoms see several periodic images of the same neighbour."""
    from fortnet_b200 import synthetic
    ds = synthetic.dense_liquid(n_atoms=128, density_aa3=0.070, seed=99, n_struct=1)
    funcs = fb.GFunctions.from_auto_scheme(8.0 * fb.BOHR_PER_AA, 2, 128)
    ctx = fb.Context(acsf_kernel=kernel)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False)
    acsf.calculate(0)
    vals = acsf.features(0)
    assert ctx.acsf_path(0) in CELLS       # edge < 2 rc: the whole-structure path must have handed over to the cell list
    mx, mean = ctx.max_neighbors(0)
    assert 120 < mean < 180, mean
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=_nthreads())
    assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), _md(vals, ref)
    ctx.close()


def test_c5_full_structure_cell_list(fb, orc):
    """one full 4 096-atom C5 structure (cubic L = 38.8 A) through the cell list, rc = 5 A (~37 neighbours), 2 + 6
    functions, both value kernels: every atom against the oracle (whose brute-force neighbour search per function is
    what bounds the function count here -- it treats a structure on one thread)"""
    from fortnet_b200 import synthetic
    ds = synthetic.dense_liquid(n_atoms=4096, density_aa3=0.070, seed=99, n_struct=1)
    funcs = fb.GFunctions.from_auto_scheme(5.0 * fb.BOHR_PER_AA, 2, 6)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=1)
    for kernel in KERNELS:
        ctx = fb.Context(acsf_kernel=kernel)
        ctx.upload(0, ds)
        acsf = fb.Acsf(ctx, funcs, standardize=False)
        acsf.calculate(0)
        assert ctx.acsf_path(0) in CELLS
        assert ctx.acsf_kernel() == (1 if kernel == "auto" else 0)
        mx, mean = ctx.max_neighbors(0)
        assert 30 < mean < 45, mean
        vals = acsf.features(0)
        assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), (kernel, _md(vals, ref))
        ctx.close()


@pytest.mark.parametrize("act", ["gaussian", "relu", "lrelu", "softplus", "bent", "atan", "sigmoid", "heaviside", "tanh", "linear"])
@pytest.mark.parametrize("loss", ["mse", "rms", "mae", "mape"])
def test_activations_and_losses(fb, orc, act, loss):
    """every transfer function (transfer.F90) x every loss (loss.F90:217-281) on one small batch"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=3, seed=11)
    ds.gtargets[:] = np.random.default_rng(5).uniform(1.0, 2.0, size=ds.gtargets.shape)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 5, 4)
    _full_check(fb, orc, ds, funcs, [9, 7, 5, 1], act=act, loss=loss, forces=(loss == "mse"))


def test_reconfigure_with_larger_features_and_outputs(fb, orc):
    """a live slot is re-configured with MORE ACSF functions and a wider output layer: the per-slot dE/dG and
    force buffers of the first configuration must be re-sized (they were allocated once), forces vs oracle"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=3, seed=21)
    ctx = fb.Context()
    ctx.upload(0, ds)
    rng = np.random.default_rng(22)
    for nrad, nang, nout in [(3, 2, 1), (9, 10, 3)]:
        funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, nrad, nang)
        dims = [len(funcs), 6, nout]
        acsf = fb.Acsf(ctx, funcs, standardize=False)
        acsf.calculate(0)
        net = fb.Bpnn(ctx, dims, 1, "tanh")
        wb = rng.uniform(-0.5, 0.5, size=(1, _ntot(dims)))
        net.set_params(wb)
        f = net.forces(0)
        glob, raw, frc = ctx.socket_step(0, ds.coords, ds.latvecs)
        feats = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=_nthreads())
        f_o = orc.forces(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), feats, ds.globalsp, dims,
                         "tanh", wb, nthreads=_nthreads())
        assert f.shape == (ds.n_atoms, 3 * nout) and frc.shape == f.shape and raw.shape == (ds.n_atoms, nout)
        assert np.allclose(f, f_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(f_o).max())), _md(f, f_o)
        assert np.allclose(frc, f_o, rtol=RTOL, atol=ATOL * max(1.0, np.abs(f_o).max())), _md(frc, f_o)
    with pytest.raises(fb.FnetGpuError):
        ctx.socket_step(0, ds.coords, ds.latvecs, n_out=1)       # the network has three outputs
    ctx.close()


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def _mixed_functions(fb, rc, zs=None):
    G = fb.GFunction
    fs = [G("g1", rc), G("g2", rc, eta=0.3, rs=1.1), G("g2", 0.8 * rc, eta=1.7, rs=0.0), G("g3", rc, kappa=0.9),
          G("g4", rc, xi=1.0, eta=0.05, lam=1.0), G("g4", rc, xi=2.5, eta=0.05, lam=-1.0),
          G("g5", rc, xi=4.0, eta=0.02, lam=1.0), G("g5", 0.7 * rc, xi=1.5, eta=0.1, lam=-1.0),
          G("g5", rc, xi=0.0, eta=0.02, lam=1.0)]
    out = fb.GFunctions(fs)
    if zs is not None:
        out = out.resolve_species(zs)
    return out


@pytest.mark.parametrize("path", PATHS)
def test_ragged_clusters_and_single_atoms(fb, orc, path, mlp="auto"):
    """non-periodic structures of 1..40 atoms in one batch (prediction/** goldens are clusters);
    a single atom has no neighbours -> all ACSF are 0 (before the z-score)"""
    rng = np.random.default_rng(21)
    natoms = [1, 2, 3, 17, 1, 40, 5, 33]
    coords = np.concatenate([rng.uniform(0.0, 3.0 + 1.2 * n ** (1 / 3), size=(n, 3)) * fb.BOHR_PER_AA for n in natoms])
    N = sum(natoms)
    atnum = rng.choice([1, 8], size=N).astype(np.int32)
    atnum[0] = 1
    atnum[1] = 8
    ds = fb.Dataset.build(natoms, coords, np.zeros(len(natoms), np.int32), np.zeros((len(natoms), 3, 3)), atnum,
                          gtargets=rng.normal(size=(len(natoms), 1)), weights=rng.integers(1, 4, size=len(natoms)),
                          atomic_weights=rng.uniform(0.5, 1.5, size=N), atomic_numbers=[1, 8])
    funcs = _mixed_functions(fb, 3.0 * fb.BOHR_PER_AA, [1, 8])
    _full_check(fb, orc, ds, funcs, [len(funcs), 6, 4, 1], act="sigmoid", acsf_path=path, expect_path=STRUCT)


@pytest.mark.parametrize("path", PATHS)
def test_atomic_and_multiple_targets(fb, orc, path):
    """two global + two atomic targets (bothTargets goldens): nOut = 4, forces have 3*nOut columns"""
    rng = np.random.default_rng(22)
    natoms = [9, 14, 6]
    N = sum(natoms)
    L = 9.0 * fb.BOHR_PER_AA
    lat = np.stack([np.eye(3) * L] * 3)
    coords = np.concatenate([rng.uniform(0, L, size=(n, 3)) for n in natoms])
    atnum = rng.choice([14, 6], size=N).astype(np.int32)
    ds = fb.Dataset.build(natoms, coords, np.ones(3, np.int32), lat, atnum, gtargets=rng.normal(size=(3, 2)),
                          atargets=rng.normal(size=(N, 2)), atomic_weights=rng.uniform(0.5, 1.5, size=N),
                          atomic_numbers=[14, 6])
    funcs = _mixed_functions(fb, 4.0 * fb.BOHR_PER_AA)
    _full_check(fb, orc, ds, funcs, [len(funcs), 5, 4], act="tanh", acsf_path=path, expect_path=STRUCT)


def test_thin_triclinic_and_unfolded_cells(fb, orc):
    """cell edges < rc (an atom sees several images of a neighbour and of itself), triclinic
    lattices, and coordinates outside the unit cell (the reference does not fold them; we do).  Values
    only: the reference's dense derivative overwrites image contributions when an edge < 2 rc
    (acsf.F90:918), so forces parity is undefined here."""
    rng = np.random.default_rng(23)
    rc = 4.0 * fb.BOHR_PER_AA
    lats = np.array([
        [[0.6 * rc, 0, 0], [0, 2.3 * rc, 0], [0, 0, 1.1 * rc]],
        [[2.1 * rc, 0, 0], [0.7 * rc, 1.9 * rc, 0], [-0.4 * rc, 0.5 * rc, 2.2 * rc]],
        [[0.9 * rc, 0.2 * rc, 0], [0, 0.8 * rc, 0.1 * rc], [0.3 * rc, 0, 0.7 * rc]],
        [[3.0 * rc, 0, 0], [0, 3.0 * rc, 0], [0, 0, 3.0 * rc]],
    ])
    natoms = [3, 20, 1, 30]
    coords = []
    for s, n in enumerate(natoms):
        # The reference never folds coordinates and searches images in +-(floor(rc |b_k|) + 1)
        # (latpointiter.F90:204-253), which finds every image only while the fractional coordinates of
        # two atoms differ by less than 1 per axis (any reader output with 0 <= frac <= 1 does).  That is
        # the parity domain: a window of width 1 anywhere in space, here shifted by whole lattice vectors.
        if s != 3:
            frac = rng.uniform(0.3, 1.3, size=(n, 3)) + np.array([3.0, -2.0, 5.0])
        else:
            frac = rng.uniform(0.0, 1.0, size=(n, 3))
            frac[0] = [1.0, 0.0, 1.0]          # fractional coordinate exactly 1.0 stays unfolded in the reference
        coords.append(frac @ lats[s])
    coords = np.concatenate(coords)
    N = sum(natoms)
    atnum = rng.choice([22, 8], size=N).astype(np.int32)
    ds = fb.Dataset.build(natoms, coords, np.ones(4, np.int32), lats, atnum, gtargets=rng.normal(size=(4, 1)),
                          atomic_numbers=[22, 8])
    funcs = _mixed_functions(fb, rc, [22, 8])
    # auto mode: the in-kernel lattice check rejects the thin cells and the call falls back to the cell list
    _full_check(fb, orc, ds, funcs, [len(funcs), 4, 1], forces=False, expect_path=CELLS)


def test_whole_structure_path_triclinic_unfolded_mixed(fb, orc):
    """the whole-structure / minimum-image path on its own ground: skewed triclinic cells whose plane
    spacings are just above 2 rc, coordinates far outside the unit cell, fractional coordinate exactly
    1.0, clusters and periodic structures in one batch, 1..200 atoms; values, gradient AND forces
    (every edge >= 2 rc, so the reference's derivative is well defined)"""
    rng = np.random.default_rng(31)
    rc = 4.0 * fb.BOHR_PER_AA
    lats = np.array([
        [[2.05 * rc, 0, 0], [0.9 * rc, 2.02 * rc, 0], [-0.7 * rc, 0.8 * rc, 2.01 * rc]],     # rescaled below: min spacing 2.01 rc
        [[3.0 * rc, 0, 0], [0, 2.4 * rc, 0], [0, 0, 2.000001 * rc]],
        [[0, 0, 0], [0, 0, 0], [0, 0, 0]],                                                  # cluster
        [[2.6 * rc, 0.3 * rc, 0.2 * rc], [-0.2 * rc, 2.7 * rc, 0.4 * rc], [0.1 * rc, -0.3 * rc, 2.5 * rc]],
        [[0, 0, 0], [0, 0, 0], [0, 0, 0]],                                                  # single atom
    ])
    per = np.array([1, 1, 0, 1, 0], np.int32)
    for k in (0, 3):     # skewed cells: scale so that the smallest lattice-plane spacing is 2.01 rc
        h = 1.0 / np.linalg.norm(np.linalg.inv(lats[k]), axis=0)
        lats[k] *= 2.01 * rc / h.min()
    natoms = [40, 33, 25, 200, 1]
    coords = []
    for s, n in enumerate(natoms):
        if per[s]:
            frac = rng.uniform(0.2, 1.2, size=(n, 3)) + np.array([-4.0, 7.0, 2.0])
            if s == 1:
                frac = rng.uniform(0.0, 1.0, size=(n, 3))
                frac[0] = [1.0, 0.0, 1.0]
            coords.append(frac @ lats[s])
        else:
            coords.append(rng.uniform(0.0, 2.2 * rc, size=(n, 3)) + 50.0)
    coords = np.concatenate(coords)
    N = sum(natoms)
    atnum = rng.choice([22, 8], size=N).astype(np.int32)
    ds = fb.Dataset.build(natoms, coords, per, lats, atnum, gtargets=rng.normal(size=(5, 1)),
                          weights=rng.integers(1, 3, size=5), atomic_numbers=[22, 8])
    funcs = _mixed_functions(fb, rc, [22, 8])
    _full_check(fb, orc, ds, funcs, [len(funcs), 5, 1], expect_path=STRUCT)
    _full_check(fb, orc, ds, funcs, [len(funcs), 5, 1], acsf_path="cells")


def test_whole_structure_path_hands_over_after_lattice_update(fb, orc):
    """coords_update with a lattice that shrinks below 2 rc: the slot must switch to the cell list on
    its own (and back to the whole-structure path when the lattice grows again)"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=3, seed=8)
    rc = 4.0 * fb.BOHR_PER_AA
    funcs = fb.GFunctions.from_auto_scheme(rc, 6, 6)
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False)
    acsf.calculate(0)
    assert ctx.acsf_path(0) in STRUCT
    for scale, expect in ((0.7, CELLS), (1.0, STRUCT)):
        c2, l2 = ds.coords * scale, ds.latvecs * scale
        ctx.update_coords(0, c2, l2)
        acsf.calculate(0)
        assert ctx.acsf_path(0) in expect, (scale, ctx.acsf_path(0))
        ref = orc.acsf(ds.offsets, c2, ds.periodic, l2, ds.atnum, funcs.asdicts())
        vals = acsf.features(0)
        assert np.allclose(vals, ref, rtol=1e-10, atol=1e-12), _md(vals, ref)
    ctx.close()


@pytest.mark.parametrize("path", PATHS)
def test_update_calculate_pipelined_upload(fb, path):
    """fnetgpu_acsf_update_calculate (chunked upload overlapped with the kernel, 3 chunks here) gives
    bit-identical features to fnetgpu_coords_update + fnetgpu_acsf_calculate, also when the lattice changes"""
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=1600, seed=77)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 6, 6)
    ctx = fb.Context(acsf_path=path)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    rng = np.random.default_rng(3)
    for scale in (1.0, 1.02, 0.7):
        c = (ds.coords + rng.normal(scale=0.03, size=ds.coords.shape)) * scale
        l = ds.latvecs * scale
        acsf.calculate(0, coords=c, latvecs=l)
        a = acsf.features(0)
        if path == "auto":
            assert ctx.acsf_path(0) in (STRUCT if scale > 0.9 else CELLS)
        ctx.update_coords(0, c, l)
        acsf.calculate(0)
        b = acsf.features(0)
        assert np.array_equal(a, b), _md(a, b)
    ctx.close()


@pytest.mark.parametrize("path,precision", [("auto", 64), ("cells", 64), ("auto", 32)])
def test_socket_step(fb, orc, path, precision):
    """fnetgpu_socket_step (predictForSocketComm, fortnet.F90:503-609): an MD trajectory of a resident
    structure -- every step must equal the blocking call sequence and the oracle; the cell shrinking
    below 2 rc in mid-trajectory sends the step through the general path and back"""
    from fortnet_b200 import synthetic
    rng = np.random.default_rng(41)
    ds = synthetic.si_bulk(n_struct=2, seed=9)
    rc = 4.0 * fb.BOHR_PER_AA
    funcs = fb.GFunctions.from_auto_scheme(rc, 6, 6)
    dims = [12, 7, 5, 1]
    wb = rng.uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    ctx = fb.Context(acsf_path=path, precision=precision)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)                                   # fixes the z-score statistics (training-set role)
    mu, sg = acsf.zprec
    net = fb.Bpnn(ctx, dims, 1, "tanh")
    net.set_params(wb)
    rt, at = (RTOL, ATOL) if precision == 64 else (2e-4, 2e-4)
    # second context: the blocking call sequence on the cell list (forces for edges < 2 rc are the true
    # gradient, which the reference's dense derivative does not give -- SURVEY.md section 7)
    ctx2 = fb.Context(acsf_path="cells", precision=precision)
    ctx2.upload(0, ds)
    acsf2 = fb.Acsf(ctx2, funcs, standardize=True)
    acsf2.calculate(0, zprec=np.stack([mu, sg]))
    net2 = fb.Bpnn(ctx2, dims, 1, "tanh")
    net2.set_params(wb)
    coords = ds.coords.copy()
    for step, scale in enumerate([1.0, 1.0, 0.7, 0.7, 1.0, 1.0]):
        coords = coords + rng.normal(scale=0.02, size=coords.shape)
        c, l = coords * scale, ds.latvecs * scale
        glob, raw, frc = ctx.socket_step(0, c, l if step != 1 else None)
        if precision == 64:
            want = (2,) if (path == "auto" and scale == 1.0) else (0, 1)
            assert ctx.acsf_path(0) in want, (step, ctx.acsf_path(0))
        ref = orc.zscore_apply(orc.acsf(ds.offsets, c, ds.periodic, l, ds.atnum, funcs.asdicts()), mu, sg)
        raw_o = orc.predict(ref, ds.globalsp, dims, "tanh", wb)
        assert np.allclose(raw, raw_o, rtol=rt, atol=at), (step, _md(raw, raw_o))
        assert np.allclose(glob[:, 0], np.add.reduceat(raw_o[:, 0], ds.offsets[:-1].astype(int)), rtol=rt, atol=at * 64)
        if scale == 1.0:
            f_o = orc.forces(ds.offsets, c, ds.periodic, l, ds.atnum, funcs.asdicts(), ref, ds.globalsp, dims, "tanh", wb, sigmas=sg)
        else:
            ctx2.update_coords(0, c, l)
            acsf2.calculate(0, zprec=np.stack([mu, sg]))
            f_o = net2.forces(0)
        assert np.allclose(frc, f_o, rtol=rt, atol=at * max(1.0, np.abs(f_o).max())), (step, _md(frc, f_o))
    # and the blocking sequence on the last geometry gives the same numbers
    ctx.update_coords(0, c, l)
    acsf.calculate(0, zprec=np.stack([mu, sg]))
    assert np.allclose(net.predict_batch(0), raw, rtol=rt, atol=at)
    assert np.allclose(net.forces(0), frc, rtol=rt, atol=at * max(1.0, np.abs(frc).max()))
    ctx.close()
    ctx2.close()


def test_socket_step_survives_growth_of_the_result_buffer(fb):
    """the pinned result buffer (gradient fetch, socket outputs) and the socket step's staged geometry are separate
    allocations: growing the first after a socket step must not release the second (it once did, without resetting
    the pointer: the next MD step read freed host memory and fnetgpu_finalize freed it twice -- found by
    compute-sanitizer)"""
    from fortnet_b200 import synthetic
    rng = np.random.default_rng(5)
    ds = synthetic.si_bulk(n_struct=1, seed=3)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 6, 6)
    dims = [12, 24, 24, 1]                              # 937 parameters > 64 + 192 + 16 doubles of socket results
    wb = rng.uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    net = fb.Bpnn(ctx, dims, 1, "tanh")
    net.set_params(wb)
    c = ds.coords + rng.normal(scale=0.02, size=ds.coords.shape)
    first = [np.array(a) for a in ctx.socket_step(0, c)]
    first2 = [np.array(a) for a in ctx.socket_step(0, c)]       # (graph captured on the second step)
    net.update_gradients(0, "mse", fetch=True)                  # grows the pinned result buffer
    again = [np.array(a) for a in ctx.socket_step(0, c)]
    for a, b, d in zip(first, first2, again):
        assert np.array_equal(a, b) and np.array_equal(a, d)
    ctx.close()


@pytest.mark.parametrize("path", PATHS)
def test_atom_id_scaling_and_external_features(fb, orc, path):
    """q_i q_j prefactors from an external-feature row (acsf.F90:836-840,1003-1052) and external
    columns appended to the ACSF block (features.F90:227-241)"""
    from fortnet_b200 import synthetic
    rng = np.random.default_rng(24)
    base = synthetic.si_bulk(n_struct=2, seed=5)
    ext = rng.uniform(0.5, 1.5, size=(base.n_atoms, 3))
    ds = fb.Dataset.build(np.diff(base.offsets), base.coords, base.periodic, base.latvecs, base.atnum,
                          gtargets=base.gtargets, ext=ext, atomic_numbers=[14])
    rc = 4.0 * fb.BOHR_PER_AA
    G = fb.GFunction
    funcs = fb.GFunctions.from_auto_scheme(rc, 4, 4, atomid=2)
    funcs.append(fb.GFunctions([G("g4", rc, xi=2.0, eta=0.03, lam=1.0, atomid=1), G("g1", rc, atomid=3), G("g3", rc, kappa=0.5)]))
    fd = funcs.asdicts()
    ctx = fb.Context(acsf_path=path)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False, ext_indices=[3, 1])      # 1-based rows of extFeatures (features.F90:227-241)
    acsf.calculate(0)
    vals = acsf.features(0)
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, fd, ext=ds.ext)
    assert vals.shape == (ds.n_atoms, len(funcs) + 2)
    assert np.allclose(vals[:, :len(funcs)], ref, rtol=1e-10, atol=1e-12), _md(vals[:, :len(funcs)], ref)
    assert np.array_equal(vals[:, len(funcs):], ext[:, [2, 0]])
    ctx.close()


def test_empty_and_unconfigured_calls_fail_loudly(fb):
    """error convention of the boundary: non-zero return + message, never a silent fallback"""
    from fortnet_b200 import synthetic
    ctx = fb.Context()
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 4, 4)
    acsf = fb.Acsf(ctx, funcs)
    with pytest.raises(fb.FnetGpuError):
        acsf.calculate(3)                      # empty slot
    with pytest.raises(fb.FnetGpuError):
        fb.Bpnn(ctx, [8, 4, 1], 1, "nope")
    ds = synthetic.si_bulk(n_struct=1, seed=2)
    ctx.upload(0, ds)
    acsf.calculate(0)
    net = fb.Bpnn(ctx, [9, 4, 1], 1, "tanh")     # input width != number of features
    net.set_params(np.zeros((1, _ntot([9, 4, 1]))))
    with pytest.raises(fb.FnetGpuError):
        net.predict_batch(0)
    ctx.close()


# ---------------------------------------------------------------------------------------------
# size-independent properties at the full C2 size (640k atoms) -- no oracle involved
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_full(fb):
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=10000, seed=20260001)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
    return ds, funcs


def test_full_size_invariances(fb, c2_full):
    """ACSF are invariant under rigid translation (incl. across the cell boundary) and atom
    permutation within a structure; the gradient is invariant under both as well."""
    ds, funcs = c2_full
    dims = [32, 20, 20, 1]
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=False)
    acsf.calculate(0)
    f0 = acsf.features(0)
    assert np.isfinite(f0).all() and f0.min() >= 0.0            # G2/G5 sums of non-negative terms
    z = fb.Acsf(ctx, funcs, standardize=True)
    z.calculate(0)
    zf = z.features(0)
    on = z.zprec[1] >= 1e-8                                      # sigma < 1e-8: feature left as it is (acsf.F90:505)
    assert on.sum() >= 24
    assert np.allclose(zf[:, on].mean(0), 0.0, atol=1e-9) and np.allclose(zf[:, on].std(0), 1.0, atol=1e-9)
    assert np.array_equal(zf[:, ~on], f0[:, ~on])
    net = fb.Bpnn(ctx, dims, 1, "tanh")
    net.set_params(wb)
    dd0, loss0 = net.update_gradients(0, "mse")
    # translate every structure by its own vector, permute atoms within each structure
    rng = np.random.default_rng(31)
    n = 64
    shift = rng.uniform(-20.0, 20.0, size=(ds.n_struct, 1, 3))
    perm = np.argsort(rng.random((ds.n_struct, n)), axis=1)
    c = ds.coords.reshape(ds.n_struct, n, 3) + shift
    c = np.take_along_axis(c, perm[:, :, None], axis=1)
    ctx.update_coords(0, c.reshape(-1, 3), ds.latvecs)
    acsf.calculate(0)
    f1 = acsf.features(0).reshape(ds.n_struct, n, -1)
    f0p = np.take_along_axis(f0.reshape(ds.n_struct, n, -1), perm[:, :, None], axis=1)
    assert np.allclose(f1, f0p, rtol=1e-9, atol=1e-10), _md(f1, f0p)
    z.calculate(0)
    dd1, loss1 = net.update_gradients(0, "mse")
    assert np.allclose(dd1, dd0, rtol=1e-8, atol=1e-9 * np.abs(dd0).max()), _md(dd1, dd0)
    assert abs(loss1 - loss0) <= 1e-9 * abs(loss0)
    ctx.close()


def test_full_size_gradient_linearity_and_determinism(fb, c2_full):
    """dd is linear in the datapoint weights (bpnn.F90:443-450): dd(w) + dd(w') = dd(w + w');
    two runs on the same input are bit-identical (fixed-order reductions, no float atomics)."""
    ds, funcs = c2_full
    dims = [32, 20, 20, 1]
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    rng = np.random.default_rng(32)
    w1 = rng.integers(1, 5, size=ds.n_struct).astype(np.int32)
    w2 = rng.integers(1, 5, size=ds.n_struct).astype(np.int32)
    zp = None
    out = []
    for w in (w1, w2, w1 + w2, w1 + w2):
        ctx = fb.Context()
        d = fb.Dataset.build(np.diff(ds.offsets), ds.coords, ds.periodic, ds.latvecs, ds.atnum, weights=w,
                             gtargets=ds.gtargets, atomic_numbers=[14])
        ctx.upload(0, d)
        acsf = fb.Acsf(ctx, funcs, standardize=True)
        acsf.calculate(0, zprec=zp)
        if zp is None:
            zp = acsf.zprec.copy()
        net = fb.Bpnn(ctx, dims, 1, "tanh")
        net.set_params(wb)
        out.append(net.update_gradients(0, "mse")[0])
        ctx.close()
    assert np.array_equal(out[2], out[3])
    assert np.allclose(out[0] + out[1], out[2], rtol=1e-10, atol=1e-12 * np.abs(out[2]).max()), _md(out[0] + out[1], out[2])


def test_full_size_forces_sum_rule_and_finite_difference(fb):
    """C4 shape (TiO2-like 192-atom cells, forces on): the forces of every structure sum to zero
    (translation invariance), and F = -dE/dR against a central finite difference of the predicted
    energy for a few displaced atoms."""
    from fortnet_b200 import synthetic
    ds = synthetic.tio2(n_struct=600, seed=20260004)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 8, 16).resolve_species([22, 8])
    dims = [64, 32, 32, 32, 1]
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(2, _ntot(dims)))
    ctx = fb.Context()
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    net = fb.Bpnn(ctx, dims, 2, "tanh")
    net.set_params(wb)
    f = net.forces(0)
    fs = f.reshape(ds.n_struct, 192, 3).sum(1)
    assert np.abs(fs).max() <= 1e-9 * np.abs(f).max(), np.abs(fs).max()
    h = 1e-4
    rng = np.random.default_rng(33)
    for _ in range(3):
        s = int(rng.integers(ds.n_struct))
        i = int(ds.offsets[s] + rng.integers(192))
        c = int(rng.integers(3))
        e = []
        for sgn in (+1, -1):
            cc = ds.coords.copy()
            cc[i, c] += sgn * h
            ctx.update_coords(0, cc, ds.latvecs)
            acsf.calculate(0)
            raw = net.predict_batch(0)
            e.append(raw[ds.offsets[s]:ds.offsets[s + 1], 0].sum())
        fd = -(e[0] - e[1]) / (2 * h)
        assert abs(fd - f[i, c]) <= 1e-6 * max(1.0, abs(f[i, c])), (fd, f[i, c])
    ctx.close()


# ---------------------------------------------------------------------------------------------
# FP32 mode: documented bound
# ---------------------------------------------------------------------------------------------
def test_fp32_mode_bound(fb, orc):
    from fortnet_b200 import synthetic
    ds = synthetic.si_bulk(n_struct=16, seed=20260001)
    funcs = fb.GFunctions.from_auto_scheme(4.0 * fb.BOHR_PER_AA, 16, 16)
    dims = [32, 20, 20, 1]
    wb = np.random.default_rng(7).uniform(-0.5, 0.5, size=(1, _ntot(dims)))
    ref = orc.acsf(ds.offsets, ds.coords, ds.periodic, ds.latvecs, ds.atnum, funcs.asdicts(), nthreads=_nthreads())
    mu, sg = orc.zscore_stats(ds.offsets, ref, ds.weights)
    zref = orc.zscore_apply(ref, mu, sg)
    dd_o, raw_o = orc.grad(ds.offsets, zref, ds.globalsp, dims, "tanh", wb, "mse", ds.weights, ds.atomic_weights,
                           ds.gtargets, ds.atargets)
    ctx = fb.Context(precision=32)
    ctx.upload(0, ds)
    acsf = fb.Acsf(ctx, funcs, standardize=True)
    acsf.calculate(0)
    z = acsf.features(0)
    # precision 32: per-neighbour factors and radial sums in FP64, the angular pair sums in FP32 (FFMA ladders, ex2 / lg2
    # on the MUFU pipe), features stored as FP32.  Measured (tools/fp32_errors.py, C2 / C3 / C5-like): raw ACSF 2e-6
    # relative, z-scored 6e-6 absolute; the documented bound is 1e-5
    assert np.allclose(z, zref, rtol=1e-5, atol=1e-5), _md(z, zref)
    assert ctx.acsf_launch_info(0)["lean"] == 1
    net = fb.Bpnn(ctx, dims, 1, "tanh")
    net.set_params(wb)
    raw = net.predict_batch(0)
    assert np.abs(raw - raw_o).max() <= 2e-5 * max(1.0, np.abs(raw_o).max()), _md(raw, raw_o)
    dd, lossv = net.update_gradients(0, "mse")
    assert np.abs(dd - dd_o).max() <= 1e-4 * np.abs(dd_o).max(), _md(dd, dd_o)
    ctx.close()
