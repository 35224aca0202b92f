"""Flat, device-uploadable view of a Fortnet dataset.

Mirrors what the reference keeps in ``TDataset`` (lib_io/fnetdata.F90:36-120) for the hot
path only: geometries, species maps, targets, weights and external features, concatenated
over structures so that one ``fnetgpu_dataset_upload`` call ships them to HBM.

All lengths are Bohr, coordinates Cartesian.  ``latvecs[s, k, :]`` is lattice vector k of
structure s (= ``latVecs(:,k)``, the ``basis`` rows of fnetdata.hdf5).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def frac_to_cart(frac, latvecs):
    """coords = matmul(latVecs, frac) -- lib_io/fnetdata.F90:1148-1151 (no folding)."""
    return np.asarray(frac, np.float64) @ np.asarray(latvecs, np.float64)


@dataclass
class Dataset:
    offsets: np.ndarray          # (nS+1,) int32, atom offsets of the structures
    coords: np.ndarray           # (N,3) f64 Cartesian Bohr
    periodic: np.ndarray         # (nS,) int32
    latvecs: np.ndarray          # (nS,3,3) f64
    atnum: np.ndarray            # (N,) int32   localAtToAtNum
    globalsp: np.ndarray         # (N,) int32   localAtToGlobalSp (1-based)
    weights: np.ndarray          # (nS,) int32  datapoint weights
    atomic_weights: np.ndarray   # (N,) f64
    gtargets: np.ndarray         # (nS,nG) f64
    atargets: np.ndarray         # (N,nA) f64
    ext: np.ndarray              # (N,nExt) f64
    atomic_numbers: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    @property
    def n_struct(self):
        return len(self.offsets) - 1

    @property
    def n_atoms(self):
        return int(self.offsets[-1])

    @property
    def n_global_targets(self):
        return self.gtargets.shape[1]

    @property
    def n_atomic_targets(self):
        return self.atargets.shape[1]

    @property
    def n_ext(self):
        return self.ext.shape[1]

    @classmethod
    def build(cls, natoms, coords, periodic, latvecs, atnum, globalsp=None, weights=None,
              atomic_weights=None, gtargets=None, atargets=None, ext=None, atomic_numbers=None):
        natoms = np.asarray(natoms, np.int64)
        nS = len(natoms)
        offsets = np.zeros(nS + 1, np.int32)
        offsets[1:] = np.cumsum(natoms)
        N = int(offsets[-1])
        atnum = np.ascontiguousarray(atnum, np.int32)
        if atomic_numbers is None:
            _, first = np.unique(atnum, return_index=True)
            atomic_numbers = atnum[np.sort(first)]
        atomic_numbers = np.asarray(atomic_numbers, np.int32)
        if globalsp is None:
            lut = {int(z): i + 1 for i, z in enumerate(atomic_numbers)}
            globalsp = np.array([lut[int(z)] for z in atnum], np.int32)
        return cls(
            offsets=offsets,
            coords=np.ascontiguousarray(coords, np.float64).reshape(N, 3),
            periodic=np.ascontiguousarray(periodic, np.int32).reshape(nS),
            latvecs=np.ascontiguousarray(latvecs, np.float64).reshape(nS, 3, 3),
            atnum=atnum,
            globalsp=np.ascontiguousarray(globalsp, np.int32),
            weights=np.ones(nS, np.int32) if weights is None else np.ascontiguousarray(weights, np.int32),
            atomic_weights=np.ones(N) if atomic_weights is None else np.ascontiguousarray(atomic_weights, np.float64),
            gtargets=np.zeros((nS, 0)) if gtargets is None else np.ascontiguousarray(gtargets, np.float64).reshape(nS, -1),
            atargets=np.zeros((N, 0)) if atargets is None else np.ascontiguousarray(atargets, np.float64).reshape(N, -1),
            ext=np.zeros((N, 0)) if ext is None else np.ascontiguousarray(ext, np.float64).reshape(N, -1),
            atomic_numbers=atomic_numbers,
        )

    @classmethod
    def from_golden(cls, npz):
        """From a tests/golden/datasets/*.npz fixture (fractional coordinates are converted the
        way the reference's reader does, fnetdata.F90:1141-1152)."""
        natoms = npz["natoms"]
        coords = npz["coords"].copy()
        o = 0
        for s, n in enumerate(natoms):
            if npz["periodic"][s] and npz["fractional"][s]:
                coords[o:o + n] = frac_to_cart(coords[o:o + n], npz["latvecs"][s])
            o += n
        return cls.build(natoms, coords, npz["periodic"], npz["latvecs"], npz["atnum"],
                         globalsp=npz["globalsp"], weights=npz["weights"],
                         atomic_weights=npz["atomicweights"], gtargets=npz["globaltargets"],
                         atargets=npz["atomictargets"], ext=npz["extfeatures"],
                         atomic_numbers=npz["atomicnumbers"])

    def select(self, struct_ids):
        """Sub-dataset with the given structures (used for sharding over GPUs)."""
        struct_ids = np.asarray(struct_ids, np.int64)
        idx = np.concatenate([np.arange(self.offsets[s], self.offsets[s + 1]) for s in struct_ids]) \
            if len(struct_ids) else np.zeros(0, np.int64)
        natoms = (self.offsets[1:] - self.offsets[:-1])[struct_ids]
        return Dataset.build(natoms, self.coords[idx], self.periodic[struct_ids], self.latvecs[struct_ids],
                             self.atnum[idx], globalsp=self.globalsp[idx], weights=self.weights[struct_ids],
                             atomic_weights=self.atomic_weights[idx], gtargets=self.gtargets[struct_ids],
                             atargets=self.atargets[idx], ext=self.ext[idx],
                             atomic_numbers=self.atomic_numbers)
