"""Host-side mirror of the reference's hot-path interface on top of the C ABI.

``Context`` wraps one ``fnetgpu_ctx`` (one GPU).  ``Acsf`` and ``Bpnn`` mirror the type-bound
procedures the Fortran driver calls (lib_descriptors/acsf.F90: ``TAcsf%calculate`` /
``%calculatePrime``; lib_nn/bpnn.F90: ``TBpnn%updateGradients``, ``%predictBatch``,
``%nJacobian`` + lib_analysis/forces.F90: ``forceAnalysis_analytical``) with the same argument
meaning; data crosses the boundary exactly as it would from Fortran (host numpy arrays,
column-major ``array(F, nAtom)`` == C ``[nAtom][F]``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib, FnetGpuError
from .dataset import Dataset
from .gfunctions import GFunctions

ACTIVATIONS = ["gaussian", "relu", "lrelu", "softplus", "bent", "atan", "sigmoid", "heaviside",
               "tanh", "linear"]
LOSSES = ["mse", "rms", "mae", "mape"]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One GPU.  precision 64 (parity mode, default) or 32.  acsf_path "auto" (small structures:
    whole-structure / minimum-image neighbour search, else the cell list) or "cells"; mlp "auto"
    (precision 64: DMMA kernels when the network fits) or "legacy" (register-tiled FMA kernels)."""

    def __init__(self, device=-1, precision=64, deterministic=True, acsf_path="auto", mlp="auto", acsf_kernel="auto"):
        self._lib = lib()
        h = C.c_void_p()
        rc = self._lib.fnetgpu_init(C.byref(h), C.c_int(device), C.c_int(precision), C.c_int(int(deterministic)))
        if rc != 0:
            raise FnetGpuError(self._lib.fnetgpu_last_error(None).decode())
        self._h = h
        self.precision = precision
        if acsf_path not in ("auto", "cells"):
            raise ValueError("acsf_path must be 'auto' or 'cells'")
        if acsf_path == "cells":
            self._check(self._lib.fnetgpu_acsf_path_set(self._h, C.c_int(1)))
        if acsf_kernel not in ("auto", "generic"):
            raise ValueError("acsf_kernel must be 'auto' or 'generic'")
        if acsf_kernel == "generic":
            self._check(self._lib.fnetgpu_acsf_kernel_set(self._h, C.c_int(1)))
        if mlp not in ("auto", "legacy", "nofuse"):
            raise ValueError("mlp must be 'auto', 'legacy' or 'nofuse'")
        if mlp != "auto":
            self._check(self._lib.fnetgpu_mlp_path_set(self._h, C.c_int(1 if mlp == "legacy" else 2)))
        self.n_feat = {}
        self.n_atoms = {}
        self.n_struct = {}
        self.n_global = {}
        self.n_out = None            # output width of the configured network (set by Bpnn)

    def _check(self, rc):
        if rc != 0:
            raise FnetGpuError(self._lib.fnetgpu_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fnetgpu_finalize(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- dataset --------------------------------------------------------------------
    def upload(self, slot, ds: Dataset):
        nG, nA, nExt = ds.n_global_targets, ds.n_atomic_targets, ds.n_ext
        self._check(self._lib.fnetgpu_dataset_upload(
            self._h, C.c_int(slot), C.c_int(ds.n_struct), _p(_i(ds.offsets)), _p(_d(ds.coords)),
            _p(_i(ds.periodic)), _p(_d(ds.latvecs)), _p(_i(ds.atnum)), _p(_i(ds.globalsp)),
            _p(_i(ds.weights)), _p(_d(ds.atomic_weights)), C.c_int(nG), _p(_d(ds.gtargets)) if nG else None,
            C.c_int(nA), _p(_d(ds.atargets)) if nA else None, C.c_int(nExt), _p(_d(ds.ext)) if nExt else None))
        self.n_atoms[slot] = ds.n_atoms
        self.n_struct[slot] = ds.n_struct
        self.n_global[slot] = nG

    def update_coords(self, slot, coords, latvecs=None):
        self._check(self._lib.fnetgpu_coords_update(self._h, C.c_int(slot), _p(_d(coords)),
                                                    _p(_d(latvecs)) if latvecs is not None else None))

    def socket_step(self, slot, coords, latvecs=None, n_out=None, forces=True):
        """one MD / i-PI step (predictForSocketComm, fortnet.F90:503-609): new geometry in ->
        (global predictions [nStruct, nOut], atomic predictions [N, nOut], forces [N, 3*nOut]).
        The buffers are sized from the network's output width (recorded by ``Bpnn``); a caller-supplied
        n_out must agree with it -- the library writes nOut values per atom whatever the caller assumes."""
        if self.n_out is None:
            raise FnetGpuError("socket_step: no network configured (create a Bpnn on this context first)")
        if n_out is not None and int(n_out) != self.n_out:
            raise FnetGpuError("socket_step: n_out=%d does not match the network's output width %d" % (n_out, self.n_out))
        n_out = self.n_out
        N, nS = self.n_atoms[slot], self.n_struct[slot]
        glob = np.zeros((nS, n_out)); raw = np.zeros((N, n_out))
        frc = np.zeros((N, 3 * n_out)) if forces else None
        self._check(self._lib.fnetgpu_socket_step(self._h, C.c_int(slot), _p(_d(coords)),
                                                  _p(_d(latvecs)) if latvecs is not None else None,
                                                  _p(glob), _p(raw), _p(frc)))
        return glob, raw, frc

    # ---- plumbing -------------------------------------------------------------------
    def synchronize(self):
        self._check(self._lib.fnetgpu_synchronize(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.fnetgpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def launch_count(self):
        return int(self._lib.fnetgpu_launch_count(self._h))

    def profile(self, enable=True):
        self._check(self._lib.fnetgpu_profile(self._h, C.c_int(int(enable))))

    def profile_report(self):
        out = {}
        k = 0
        while True:
            name = self._lib.fnetgpu_kernel_name(C.c_int(k))
            if name is None:
                break
            ms, n = C.c_double(), C.c_longlong()
            self._check(self._lib.fnetgpu_profile_get(self._h, C.c_int(k), C.byref(ms), C.byref(n)))
            if n.value:
                out[name.decode()] = dict(ms_total=ms.value, launches=int(n.value))
            k += 1
        return out

    def mlp_path(self):
        """1: FP64 tensor-core (DMMA) subnetwork kernels, 0: register-tiled FMA kernels, -1: no network set"""
        return int(self._lib.fnetgpu_mlp_path_get(self._h))

    def acsf_path(self, slot):
        """path of the last ACSF launch: 0/1 cell list (direct/staged), 2 whole structure, -1 none"""
        return int(self._lib.fnetgpu_acsf_path_get(self._h, C.c_int(slot)))

    def acsf_kernel(self):
        """1: the configured ACSF run through k_acsf_lean (automatic-scheme configurations), 0: k_acsf, -1: none"""
        return int(self._lib.fnetgpu_acsf_kernel_get(self._h))

    def acsf_launch_info(self, slot):
        """what the last ACSF value launch used: dict(lean, atoms_per_warp, cap, cap_candidates, path, smem_bytes)"""
        info = (C.c_int * 6)()
        if self._lib.fnetgpu_acsf_launch_info(self._h, C.c_int(slot), info) != 0:
            raise FnetGpuError("acsf_launch_info: empty slot")
        return dict(zip(["lean", "atoms_per_warp", "cap", "cap_candidates", "path", "smem_bytes"], [int(v) for v in info]))

    def grad_launch_info(self, slot):
        """what the last update_gradients of the slot launched (fnetgpu_grad_launch_info)"""
        info = (C.c_int * 4)()
        if self._lib.fnetgpu_grad_launch_info(self._h, C.c_int(slot), info) != 0:
            raise FnetGpuError("grad_launch_info: empty slot")
        return dict(zip(["fusion", "cluster_size", "grid", "rounds"], [int(v) for v in info]))

    def measure_peaks(self):
        """live FP64 FMA / FP64 tensor (DMMA) / FP32 FMA issue rates of this device in T FMA/s"""
        out = (C.c_double * 3)()
        self._check(self._lib.fnetgpu_measure_peaks(self._h, out))
        return dict(dfma_tfma_s=out[0], dmma_tfma_s=out[1], ffma_tfma_s=out[2])

    def max_neighbors(self, slot):
        m, mean = C.c_int(), C.c_double()
        self._check(self._lib.fnetgpu_max_neighbors(self._h, C.c_int(slot), C.byref(m), C.byref(mean)))
        return m.value, mean.value

    # ---- multi-GPU ------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        if lib().fnetgpu_comm_unique_id(buf) != 0:
            raise FnetGpuError(lib().fnetgpu_last_error(None).decode())
        return buf.raw

    def comm_init(self, n_ranks, rank, uid: bytes):
        self._check(self._lib.fnetgpu_comm_init(self._h, C.c_int(n_ranks), C.c_int(rank), C.c_char_p(uid)))


class Acsf:
    """Mirror of ``TAcsf`` (lib_descriptors/acsf.F90:107-138)."""

    def __init__(self, ctx: Context, functions: GFunctions, standardize=False, ext_indices=None):
        self.ctx = ctx
        self.functions = functions
        self.t_zscore = bool(standardize)
        self.zprec = None            # (2,F): means, sigmas -- ``this%zPrec``
        t = functions.tables()
        F = len(functions)
        ctx._check(ctx._lib.fnetgpu_acsf_set(ctx._h, C.c_int(F), _p(t["ftype"]), _p(t["rcut"]), _p(t["kappa"]),
                                             _p(t["rs"]), _p(t["eta"]), _p(t["lam"]), _p(t["xi"]),
                                             _p(t["atomid"]), _p(t["atomicnumbers"])))
        idx = _i(ext_indices if ext_indices is not None else [])
        ctx._check(ctx._lib.fnetgpu_features_config(ctx._h, C.c_int(len(idx)), _p(idx)))
        self.n_acsf = F
        self.n_feat = F + len(idx)

    def calculate(self, slot, zprec=None, coords=None, latvecs=None):
        """``TAcsf%calculate`` (acsf.F90:540-639): features of the slot's dataset stay on the GPU.
        zprec given (or already stored) -> use it; else computed from this dataset (training set).
        coords (and optionally latvecs) given: new geometry for the slot, uploaded by the same call
        (copy overlapped with the kernel, ``fnetgpu_acsf_update_calculate``)."""
        if zprec is not None:
            self.zprec = _d(zprec).reshape(2, -1).copy()
        have = self.zprec is not None
        buf = self.zprec.reshape(-1).copy() if have else np.zeros(2 * max(self.n_acsf, 1))
        if coords is not None:
            self.ctx._check(self.ctx._lib.fnetgpu_acsf_update_calculate(
                self.ctx._h, C.c_int(slot), _p(_d(coords)), _p(_d(latvecs)) if latvecs is not None else None,
                C.c_int(int(self.t_zscore)), _p(buf), C.c_int(int(have))))
        else:
            self.ctx._check(self.ctx._lib.fnetgpu_acsf_calculate(self.ctx._h, C.c_int(slot), C.c_int(int(self.t_zscore)),
                                                                 _p(buf), C.c_int(int(have))))
        if self.t_zscore and not have and self.n_acsf:
            self.zprec = buf.reshape(2, -1)[:, :self.n_acsf].copy()
        self.ctx.n_feat[slot] = self.n_feat

    def features(self, slot):
        out = np.zeros((self.ctx.n_atoms[slot], self.ctx.n_feat[slot]))
        self.ctx._check(self.ctx._lib.fnetgpu_features_get(self.ctx._h, C.c_int(slot), _p(out)))
        return out


class Bpnn:
    """Mirror of ``TBpnn`` (lib_nn/bpnn.F90:46-90) for the hot path."""

    def __init__(self, ctx: Context, dims, n_species, activation="tanh"):
        self.ctx = ctx
        self.dims = _i(dims)
        self.n_species = int(n_species)
        self.activation = activation
        if activation not in ACTIVATIONS:
            raise FnetGpuError("unknown activation %r (no fallback)" % activation)
        ctx._check(ctx._lib.fnetgpu_net_set(ctx._h, C.c_int(self.n_species), C.c_int(len(self.dims)), _p(self.dims),
                                            C.c_int(ACTIVATIONS.index(activation))))
        self.n_tot = int(ctx._lib.fnetgpu_ntot(ctx._h))
        self.n_out = int(self.dims[-1])
        ctx.n_out = self.n_out

    def serial_weights_and_biases_fillup(self, wb):
        """``serialWeightsAndBiasesFillup`` (bpnn.F90:782-797): wb is (nSpecies, nTot)."""
        wb = _d(wb)
        assert wb.shape == (self.n_species, self.n_tot), wb.shape
        self.ctx._check(self.ctx._lib.fnetgpu_params_set(self.ctx._h, _p(wb)))

    set_params = serial_weights_and_biases_fillup

    def set_regularization(self, strength=0.0, alpha=0.0, n_datapoints=0.0):
        """``TBpnn_update``'s gradient post-processing on the device (bpnn.F90:750-767): elastic-net term
        (nestedtypes.F90:336-370) and division by sum(weights); (0, *, 0) switches both off."""
        self.ctx._check(self.ctx._lib.fnetgpu_regularization_set(self.ctx._h, C.c_double(strength), C.c_double(alpha),
                                                                 C.c_double(n_datapoints)))

    def regularization_loss(self):
        """``reguLoss`` of every species for the current parameters (loss.F90:119-196)"""
        out = np.zeros(self.n_species)
        self.ctx._check(self.ctx._lib.fnetgpu_regularization_loss(self.ctx._h, _p(out)))
        return out

    def update_gradients(self, slot, loss="mse", shuffle=None, want_global=False, fetch=True):
        """``updateGradients`` + ``loss`` (bpnn.F90:277-283, 394-481).
        Returns (ddSerial (nSpecies,nTot), loss[, globalPredictions (nStruct,nG)])."""
        dd = np.zeros((self.n_species, self.n_tot)) if fetch else None
        lossv = C.c_double(0.0)
        gp = np.zeros((self.ctx.n_struct[slot], self.ctx.n_global[slot])) if want_global else None
        self.ctx._check(self.ctx._lib.fnetgpu_grad(self.ctx._h, C.c_int(slot), C.c_int(LOSSES.index(loss)),
                                                   _p(_i(shuffle)) if shuffle is not None else None,
                                                   _p(dd), C.byref(lossv) if fetch else None, _p(gp)))
        if want_global:
            return dd, lossv.value, gp
        return dd, lossv.value

    def predict_batch(self, slot):
        """``predictBatch`` (bpnn.F90:1001-1058): (N, nOut) in dataset atom order."""
        raw = np.zeros((self.ctx.n_atoms[slot], self.n_out))
        self.ctx._check(self.ctx._lib.fnetgpu_predict(self.ctx._h, C.c_int(slot), _p(raw)))
        return raw

    def loss(self, slot, loss="mse"):
        v = C.c_double(0.0)
        self.ctx._check(self.ctx._lib.fnetgpu_loss(self.ctx._h, C.c_int(slot), C.c_int(LOSSES.index(loss)), C.byref(v)))
        return v.value

    def forces(self, slot):
        """``calculatePrime`` + ``nJacobian`` + ``forceAnalysis_analytical`` fused: (N, 3*nOut)."""
        f = np.zeros((self.ctx.n_atoms[slot], 3 * self.n_out))
        self.ctx._check(self.ctx._lib.fnetgpu_forces(self.ctx._h, C.c_int(slot), _p(f)))
        return f
