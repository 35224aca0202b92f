"""Synthetic workloads of the shapes BASELINE.json names.  Seeds are fixed;
all lengths are converted to Bohr (1 Angstrom = 1/0.529177249 Bohr, constants.F90:18-21).
Targets are sums of a smooth pair potential so that the loss is finite and non-trivial."""
from __future__ import annotations

import numpy as np

from .dataset import Dataset
from .gfunctions import BOHR_PER_AA


def _pair_energy(coords, lat, rc):
    """smooth cutoff pair potential summed over the minimum-image neighbours (targets only)."""
    inv = np.linalg.inv(lat)
    d = coords[:, None, :] - coords[None, :, :]
    s = d @ inv
    s -= np.round(s)
    d = s @ lat
    r = np.sqrt((d ** 2).sum(-1))
    np.fill_diagonal(r, np.inf)
    m = r < rc
    e = np.where(m, 0.5 * (np.cos(np.pi * np.minimum(r, rc) / rc) + 1.0) * np.exp(-0.3 * r), 0.0)
    return -0.5 * e.sum()


def _perturbed_cells(base_frac, lat0, n_struct, seed, sigma_aa=0.1, strain=0.03):
    rng = np.random.default_rng(seed)
    n = len(base_frac)
    coords = np.zeros((n_struct, n, 3))
    lats = np.zeros((n_struct, 3, 3))
    for s in range(n_struct):
        sc = 1.0 + rng.uniform(-strain, strain)
        lat = lat0 * sc
        cart = base_frac @ lat + rng.normal(0.0, sigma_aa * BOHR_PER_AA, size=(n, 3))
        coords[s] = cart
        lats[s] = lat
    return coords, lats


def si_bulk(n_struct=10000, seed=20260001, reps=2):
    """C2: diamond Si, reps^3 conventional cells (64 atoms for reps=2), a = 5.431 A."""
    a = 5.431 * BOHR_PER_AA
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.array([[i, j, k] for i in range(reps) for j in range(reps) for k in range(reps)])
    frac = ((basis[None, :, :] + cells[:, None, :]) / reps).reshape(-1, 3)
    lat0 = np.eye(3) * a * reps
    coords, lats = _perturbed_cells(frac, lat0, n_struct, seed)
    n = frac.shape[0]
    rc = 4.0 * BOHR_PER_AA
    assert lat0[0, 0] * 0.97 >= 2 * rc, "cell edge must be >= 2 rc (SURVEY.md section 7)"
    nE = min(n_struct, 64)
    e = np.array([_pair_energy(coords[s], lats[s], rc) for s in range(nE)])
    gt = np.resize(e, n_struct).reshape(n_struct, 1)
    return Dataset.build(np.full(n_struct, n), coords.reshape(-1, 3), np.ones(n_struct, np.int32), lats,
                         np.full(n_struct * n, 14, np.int32), gtargets=gt, atomic_numbers=[14])


def tio2(n_struct=20000, seed=20260002, reps=(2, 2, 8)):
    """C3: rutile-like TiO2 (a = 4.594, c = 2.959 A, 6 atoms/cell), 2x2x8 supercell = 192 atoms."""
    a, c, u = 4.594 * BOHR_PER_AA, 2.959 * BOHR_PER_AA, 0.305
    basis = np.array([[0, 0, 0], [.5, .5, .5], [u, u, 0], [1 - u, 1 - u, 0], [.5 + u, .5 - u, .5], [.5 - u, .5 + u, .5]])
    z = np.array([22, 22, 8, 8, 8, 8], np.int32)
    rx, ry, rz = reps
    cells = np.array([[i, j, k] for i in range(rx) for j in range(ry) for k in range(rz)])
    frac = ((basis[None, :, :] + cells[:, None, :]) / np.array(reps)).reshape(-1, 3)
    zz = np.tile(z, len(cells))
    lat0 = np.diag([a * rx, a * ry, c * rz])
    rc = 4.0 * BOHR_PER_AA
    assert min(np.diag(lat0)) * 0.97 >= 2 * rc, "cell edge must be >= 2 rc"
    coords, lats = _perturbed_cells(frac, lat0, n_struct, seed)
    n = frac.shape[0]
    nE = min(n_struct, 32)
    e = np.array([_pair_energy(coords[s], lats[s], rc) for s in range(nE)])
    gt = np.resize(e, n_struct).reshape(n_struct, 1)
    return Dataset.build(np.full(n_struct, n), coords.reshape(-1, 3), np.ones(n_struct, np.int32), lats,
                         np.tile(zz, n_struct), gtargets=gt, atomic_numbers=[22, 8])


def dense_liquid(n_atoms=4096, density_aa3=0.070, seed=99, min_dist_aa=1.6, n_struct=1):
    """C5: random dense liquid in a cubic box (jittered lattice start keeps a hard core)."""
    rng = np.random.default_rng(seed)
    L = (n_atoms / density_aa3) ** (1.0 / 3.0) * BOHR_PER_AA
    m = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    grid = np.array([[i, j, k] for i in range(m) for j in range(m) for k in range(m)], float)
    all_coords, lats = [], []
    for s in range(n_struct):
        sel = rng.permutation(len(grid))[:n_atoms]
        spacing = L / m
        jitter = max(0.0, 0.5 * (spacing - min_dist_aa * BOHR_PER_AA))
        pts = (grid[sel] + 0.5) * spacing + rng.uniform(-jitter, jitter, size=(n_atoms, 3))
        all_coords.append(pts)
        lats.append(np.eye(3) * L)
    coords = np.concatenate(all_coords)
    gt = np.zeros((n_struct, 1))
    return Dataset.build(np.full(n_struct, n_atoms), coords, np.ones(n_struct, np.int32), np.array(lats),
                         np.full(n_struct * n_atoms, 14, np.int32), gtargets=gt, atomic_numbers=[14])
