"""ctypes front-end of the CPU oracle (oracle/refcpu.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (fortnet_b200/) never does.

All functions take plain numpy arrays in the flat layouts documented at the top of refcpu.c
(coords (N,3) Cartesian Bohr, latvecs (nS,3,3) with row k = lattice vector k, features (N,F),
ext (N,nExt), wb (nSpecies,nTot) = the reference's serialised parameters per species).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ACTIVATIONS = ["gaussian", "relu", "lrelu", "softplus", "bent", "atan", "sigmoid", "heaviside",
               "tanh", "linear"]
LOSSES = ["mse", "rms", "mae", "mape"]
GTYPES = {"g1": 1, "g2": 2, "g3": 3, "g4": 4, "g5": 5}


def build(force=False):
    so = os.path.join(_HERE, "libfnetoracle.so")
    src = os.path.join(_HERE, "refcpu.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfnetoracle.so"], env=env,
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fnet_oracle_loss.restype = C.c_double
        _LIB.fnet_oracle_ntot.restype = C.c_int
        _LIB.fnet_oracle_max_threads.restype = C.c_int
    return _LIB


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads():
    return int(lib().fnet_oracle_max_threads())


def ntot(dims):
    dims = _i(dims)
    return int(lib().fnet_oracle_ntot(len(dims), _p(dims)))


class FuncTable:
    """Flat arrays for a list of G-functions (dicts with type,rcut,kappa,rs,eta,lam,xi,atomid,
    atomicnumbers) -- the shape fnetgpu_acsf_set() takes as well."""

    def __init__(self, funcs):
        self.F = len(funcs)
        self.ftype = _i([GTYPES[str(f["type"]).lower()] for f in funcs])
        self.rcut = _d([f["rcut"] for f in funcs])
        self.kappa = _d([f.get("kappa", 0.0) for f in funcs])
        self.rs = _d([f.get("rs", 0.0) for f in funcs])
        self.eta = _d([f.get("eta", 0.0) for f in funcs])
        self.lam = _d([f.get("lam", 0.0) for f in funcs])
        self.xi = _d([f.get("xi", 0.0) for f in funcs])
        self.atomid = _i([f.get("atomid", 0) for f in funcs])
        self.atomicnumbers = _i([f.get("atomicnumbers", [0, 0]) for f in funcs]).reshape(-1)

    def args(self):
        return (C.c_int(self.F), _p(self.ftype), _p(self.rcut), _p(self.kappa), _p(self.rs),
                _p(self.eta), _p(self.lam), _p(self.xi), _p(self.atomid), _p(self.atomicnumbers))


def acsf(offsets, coords, periodic, latvecs, atnum, funcs, ext=None, nthreads=1):
    """Raw ACSF values (N,F) -- acsf.F90:797-866."""
    ft = funcs if isinstance(funcs, FuncTable) else FuncTable(funcs)
    offsets, coords, periodic, latvecs, atnum = _i(offsets), _d(coords), _i(periodic), _d(latvecs), _i(atnum)
    N = coords.shape[0]
    ext = _d(ext) if ext is not None and ext.size else None
    nExt = ext.shape[1] if ext is not None else 0
    out = np.zeros((N, ft.F))
    lib().fnet_oracle_acsf(C.c_int(len(offsets) - 1), _p(offsets), _p(coords), _p(periodic),
                           _p(latvecs), _p(atnum), C.c_int(nExt), _p(ext), *ft.args(), _p(out),
                           C.c_int(nthreads))
    return out


def zscore_stats(offsets, vals, weights):
    offsets, vals, weights = _i(offsets), _d(vals), _i(weights)
    F = vals.shape[1]
    means, sig = np.zeros(F), np.zeros(F)
    lib().fnet_oracle_zscore_stats(C.c_int(len(offsets) - 1), _p(offsets), C.c_int(F), _p(vals),
                                   _p(weights), _p(means), _p(sig))
    return means, sig


def zscore_apply(vals, means, sigmas):
    vals = _d(vals).copy()
    lib().fnet_oracle_zscore_apply(C.c_int(vals.shape[0]), C.c_int(vals.shape[1]), _p(vals),
                                   _p(_d(means)), _p(_d(sigmas)))
    return vals


def acsf_prime_struct(coords, periodic, latvec, atnum, funcs, ext=None, sigmas=None):
    """Dense derivative tensor of ONE structure, returned as prime[f, i, a, c]
    (= array(c,a,i,f) of acsf.F90:437)."""
    ft = funcs if isinstance(funcs, FuncTable) else FuncTable(funcs)
    coords, latvec, atnum = _d(coords), _d(latvec), _i(atnum)
    n = coords.shape[0]
    ext = _d(ext) if ext is not None and ext.size else None
    nExt = ext.shape[1] if ext is not None else 0
    out = np.zeros((n, n, ft.F, 3))
    lib().fnet_oracle_acsf_prime_struct(C.c_int(n), _p(coords), C.c_int(int(periodic)), _p(latvec),
                                        _p(atnum), C.c_int(nExt), _p(ext), *ft.args(),
                                        _p(_d(sigmas)), _p(out))
    return out


def predict(feats, globalsp, dims, act, wb, nthreads=1):
    """Per-atom subnet outputs (N,nOut) -- bpnn.F90:867-900."""
    feats, globalsp, dims, wb = _d(feats), _i(globalsp), _i(dims), _d(wb)
    N = feats.shape[0]
    raw = np.zeros((N, int(dims[-1])))
    lib().fnet_oracle_predict(C.c_int(N), C.c_int(feats.shape[1]), _p(feats), _p(globalsp),
                              C.c_int(wb.shape[0]), C.c_int(len(dims)), _p(dims),
                              C.c_int(ACTIVATIONS.index(act)), _p(wb), _p(raw), C.c_int(nthreads))
    return raw


def grad(offsets, feats, globalsp, dims, act, wb, loss, ds_weights, atomic_weights, gtargets,
         atargets, shuffle=None, nthreads=1):
    """(ddSerial (nSpecies,nTot), raw (N,nOut)) -- bpnn.F90:394-481, 610-704."""
    offsets, feats, globalsp, dims, wb = _i(offsets), _d(feats), _i(globalsp), _i(dims), _d(wb)
    N = feats.shape[0]
    nS = len(offsets) - 1
    gt = _d(gtargets) if gtargets is not None and gtargets.size else None
    at = _d(atargets) if atargets is not None and atargets.size else None
    nG = gt.shape[1] if gt is not None else 0
    nA = at.shape[1] if at is not None else 0
    dd = np.zeros_like(wb)
    raw = np.zeros((N, int(dims[-1])))
    lib().fnet_oracle_grad(C.c_int(nS), _p(offsets), C.c_int(feats.shape[1]), _p(feats),
                           _p(globalsp), C.c_int(wb.shape[0]), C.c_int(len(dims)), _p(dims),
                           C.c_int(ACTIVATIONS.index(act)), _p(wb), C.c_int(LOSSES.index(loss)),
                           _p(_i(ds_weights)), _p(_d(atomic_weights)), C.c_int(nG), _p(gt),
                           C.c_int(nA), _p(at), _p(_i(shuffle)), _p(dd), _p(raw), C.c_int(nthreads))
    return dd, raw


def loss(offsets, raw, loss_name, nG, nA, gtargets, atargets, atomic_weights, ds_weights):
    offsets, raw = _i(offsets), _d(raw)
    gt = _d(gtargets) if nG else None
    at = _d(atargets) if nA else None
    return float(lib().fnet_oracle_loss(C.c_int(len(offsets) - 1), _p(offsets),
                                        C.c_int(LOSSES.index(loss_name)), C.c_int(nG), C.c_int(nA),
                                        _p(raw), _p(gt), _p(at), _p(_d(atomic_weights)),
                                        _p(_i(ds_weights))))


def jacobian(feats, globalsp, dims, act, wb):
    """(N, F, nOut): jac[i, a, t] = d out_t / d x_a -- network.F90:183-244."""
    feats, globalsp, dims, wb = _d(feats), _i(globalsp), _i(dims), _d(wb)
    N, F = feats.shape
    jac = np.zeros((N, F, int(dims[-1])))
    lib().fnet_oracle_jacobian(C.c_int(N), C.c_int(F), _p(feats), _p(globalsp), C.c_int(len(dims)),
                               _p(dims), C.c_int(ACTIVATIONS.index(act)), _p(wb), _p(jac))
    return jac


def forces(offsets, coords, periodic, latvecs, atnum, funcs, feats, globalsp, dims, act, wb,
           ext=None, sigmas=None, nthreads=1):
    """Analytic forces (N, 3*nOut), layout [atom][3*t+c] -- forces.F90:317-425."""
    ft = funcs if isinstance(funcs, FuncTable) else FuncTable(funcs)
    offsets, coords, periodic, latvecs, atnum = _i(offsets), _d(coords), _i(periodic), _d(latvecs), _i(atnum)
    feats, globalsp, dims, wb = _d(feats), _i(globalsp), _i(dims), _d(wb)
    N = coords.shape[0]
    ext = _d(ext) if ext is not None and ext.size else None
    nExt = ext.shape[1] if ext is not None else 0
    out = np.zeros((N, 3 * int(dims[-1])))
    lib().fnet_oracle_forces(C.c_int(len(offsets) - 1), _p(offsets), _p(coords), _p(periodic),
                             _p(latvecs), _p(atnum), C.c_int(nExt), _p(ext), *ft.args(),
                             _p(_d(sigmas)), _p(feats), _p(globalsp), C.c_int(len(dims)), _p(dims),
                             C.c_int(ACTIVATIONS.index(act)), _p(wb), _p(out), C.c_int(nthreads))
    return out


def sd_step(x, g, lr, max_disp):
    x = _d(x).copy().reshape(-1)
    g = _d(g).reshape(-1)
    lib().fnet_oracle_sd_step(C.c_int(x.size), _p(x), _p(g), C.c_double(lr), C.c_double(max_disp))
    return x


def start_end(n_systems, n_procs, i_proc):
    a, b = C.c_int(), C.c_int()
    lib().fnet_oracle_start_end(C.c_int(n_systems), C.c_int(n_procs), C.c_int(i_proc), C.byref(a), C.byref(b))
    return a.value, b.value
