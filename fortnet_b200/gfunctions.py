"""ACSF G-function tables: the host-side mirror of ``TGFunction`` / ``TGFunctions``
(lib_descriptors/acsf.F90:40-88) and of the two config expansions that decide the feature
ORDER the kernels must reproduce: the automatic parameter scheme
(``TGFunctions_fromAutoScheme``, acsf.F90:276-363) and the species-resolved expansion
(``processAcsfFunctions``, lib_fortnet/initprogram.F90:1453-1527).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, asdict

import numpy as np

BOHR_PER_AA = 1.0 / 0.529177249   # lib_dftbp/constants.F90:18-21
GTYPES = {"g1": 1, "g2": 2, "g3": 3, "g4": 4, "g5": 5}


@dataclass
class GFunction:
    type: str
    rcut: float
    kappa: float = 0.0
    rs: float = 0.0
    eta: float = 0.0
    lam: float = 0.0
    xi: float = 0.0
    atomid: int = 0
    atomicnumbers: tuple = (0, 0)

    @property
    def radial(self):
        return self.type.lower() in ("g1", "g2", "g3")

    def asdict(self):
        d = asdict(self)
        d["atomicnumbers"] = list(self.atomicnumbers)
        return d


class GFunctions:
    def __init__(self, funcs=()):
        self.func = [f if isinstance(f, GFunction) else GFunction(**{**f, "atomicnumbers": tuple(f.get("atomicnumbers", (0, 0)))})
                     for f in funcs]

    def __len__(self):
        return len(self.func)

    def append(self, other):
        self.func.extend(other.func)
        return self

    @classmethod
    def from_auto_scheme(cls, rcut, n_radial, n_angular, atomid=0):
        """acsf.F90:276-363.  rcut in Bohr."""
        rs_step = rcut / (n_radial - 1)
        g2eta = 5.0 * math.log(10.0) / (2.0 * rs_step) ** 2
        g2rs = [0.0] + [ii * rs_step for ii in range(1, n_radial)]
        g5eta = 2.0 * math.log(10.0) / rcut ** 2
        lam, xis = [], []
        for ii in range(0, int(math.ceil(n_angular / 2.0 - 1.0)) + 1):
            xi = 1.0 if n_angular <= 2 else 1.0 + ii * 30.0 / (n_angular - 2.0)
            for jj in (1, -1):
                if len(lam) >= n_angular:
                    break
                lam.append(float(jj))
                xis.append(xi)
        funcs = [GFunction("g2", rcut, eta=g2eta, rs=g2rs[i], atomid=atomid) for i in range(n_radial)]
        funcs += [GFunction("g5", rcut, xi=xis[i], eta=g5eta, lam=lam[i], atomid=atomid)
                  for i in range(n_angular)]
        return cls(funcs)

    def resolve_species(self, atomic_numbers):
        """initprogram.F90:1453-1527: per function, radial -> one copy per species (Z,0);
        angular -> one copy per unordered species pair (combinations with replacement)."""
        zs = [int(z) for z in atomic_numbers]
        if len(zs) <= 1:
            return self
        comb = list(itertools.combinations_with_replacement(zs, 2))
        out = []
        for f in self.func:
            if f.radial:
                for z in zs:
                    out.append(GFunction(**{**asdict(f), "atomicnumbers": (z, 0)}))
            else:
                for c in comb:
                    out.append(GFunction(**{**asdict(f), "atomicnumbers": tuple(c)}))
        return GFunctions(out)

    def tables(self):
        """Flat arrays in the order of ``fnetgpu_acsf_set`` (include/fnetgpu.h)."""
        f = self.func
        return dict(
            ftype=np.array([GTYPES[x.type.lower()] for x in f], np.int32),
            rcut=np.array([x.rcut for x in f], np.float64),
            kappa=np.array([x.kappa for x in f], np.float64),
            rs=np.array([x.rs for x in f], np.float64),
            eta=np.array([x.eta for x in f], np.float64),
            lam=np.array([x.lam for x in f], np.float64),
            xi=np.array([x.xi for x in f], np.float64),
            atomid=np.array([x.atomid for x in f], np.int32),
            atomicnumbers=np.array([list(x.atomicnumbers) for x in f], np.int32).reshape(-1),
        )

    def asdicts(self):
        return [x.asdict() for x in self.func]
