"""Loading of the committed golden fixtures (tests/golden/, made by make_golden.py from the
reference's regression suite) and the small amount of DRIVER-SIDE logic needed to replay a
case without Fortran: netstat <-> serialised parameter vector (network.F90:397-459), feature
assembly order (features.F90:200-265), elastic-net term (nestedtypes.F90:336-370), the
normalisation + steepest-descent step of TBpnn_update (bpnn.F90:751-776, steepdesc.F90:186-202).
Test infrastructure only.
"""
import json
import os

import numpy as np

from fortnet_b200.dataset import Dataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATOL, RTOL = 1e-10, 1e-9   # test/prog/fortnet/bin/testwithworkdir.py:24-25


def index():
    with open(os.path.join(GOLD, "index.json")) as fh:
        return json.load(fh)


def load_dataset(name):
    return Dataset.from_golden(np.load(os.path.join(GOLD, "datasets", name + ".npz")))


class Case:
    def __init__(self, entry):
        self.entry = entry
        self.name = entry["case"]
        z = np.load(os.path.join(GOLD, "cases", entry["file"]))
        self.arr = {k: z[k] for k in z.files}
        self.meta = json.loads(bytes(self.arr.pop("meta")).decode())
        ns = self.meta["netstat"]
        self.mode = self.meta["mode"]
        self.dims = self.arr["in_dims"].astype(np.int32)
        self.activation = ns["activation"]
        self.nG, self.nA = ns["nglobaltargets"], ns["natomictargets"]
        self.funcs = ns.get("functions", [])
        self.atomic_numbers = self.arr["in_atomicnumbers"]
        self.zmeans = self.arr.get("in_zmeans")
        self.zsigmas = self.arr.get("in_zsigmas")
        self.ext_indices = self.arr.get("in_extindices")  # 1-based rows of extfeatures
        self.dataset = load_dataset(self.meta["dataset"])
        self.training = self.meta.get("training") or {}
        self.forces = (self.meta.get("forces") or {}).get("_type")

    # --- netstat <-> serialised parameters (network.F90:397-459) -----------------------
    def wb(self, prefix="in_"):
        L = len(self.dims)
        out = []
        for sp in range(len(self.atomic_numbers)):
            ws, bs = [], [np.zeros(self.dims[0])]
            for l in range(1, L):
                ws.append(self.arr["%sw_%d_%d" % (prefix, sp, l)].reshape(-1))  # ww(d_l,d_{l+1}) col-major
                bs.append(self.arr["%sb_%d_%d" % (prefix, sp, l)].reshape(-1))
            ws.append(np.zeros(self.dims[-1]))                                    # dummy ww(d_L,1)
            out.append(np.concatenate(ws + bs))
        return np.asarray(out)

    def n_weights(self):
        d = self.dims
        return int(sum(d[i] * d[i + 1] for i in range(len(d) - 1)) + d[-1])

    def loss_name(self):
        return str(self.training.get("loss", "mse")).strip("'\"").lower()

    # --- TBpnn_update restated for SD (bpnn.F90:708-778) -----------------------------
    def regularization(self):
        """(strength, alpha) of the case's Regularization block, (0, 0) without one"""
        regu = self.training.get("regularization")
        if not regu:
            return 0.0, 0.0
        kind = regu.get("_type")
        return float(regu.get("strength", 0.0)), {"ridge": 0.0, "lasso": 1.0}.get(kind, float(regu.get("alpha", 0.0)))

    def sd_update(self, wb, dd, pre_regularized=False):
        """pre_regularized: dd already carries the elastic-net term and the division by sum(weights)
        (fnetgpu_regularization_set: done on the device)"""
        tr = self.training
        nW = self.n_weights()
        dd = dd.copy()
        regu = tr.get("regularization")
        if regu and not pre_regularized:
            kind = regu.get("_type")
            lam = float(regu.get("strength", 0.0))
            alpha = {"ridge": 0.0, "lasso": 1.0}.get(kind, float(regu.get("alpha", 0.0)))
            w = wb[:, :nW]
            dd[:, :nW] += lam / nW * ((1.0 - alpha) * w + alpha * np.sign(w))
        if not pre_regularized:
            dd /= float(np.sum(self.dataset.weights))
        lr = float(tr["learningrate"])
        maxd = float(tr["maxdisplacement"])
        thr = float(tr.get("threshold", 0.0))
        new = wb.copy()
        for sp in range(wb.shape[0]):
            if np.max(np.abs(dd[sp])) < thr:
                continue
            step = -lr * dd[sp]
            mx = np.max(np.abs(step))
            new[sp] = wb[sp] + (step if mx <= maxd else (maxd / mx) * step)
        return new, float(np.sqrt(np.sum(dd ** 2)))

    # --- first FIRE step (lib_dftbp/fire.F90:137-185; initprogram.F90:1175-1182) -----------------
    def fire_update(self, wb, dd):
        """One iteration from rest: v = 0 -> the velocity mixing is a no-op, v = -dt g, x = x0 + v dt with
        dt = dt_init = 0.1 * MaxDisplacement; one optimiser per species (bpnn.F90:761-774)."""
        tr = self.training
        nW = self.n_weights()
        dd = dd.copy()
        regu = tr.get("regularization")
        if regu:
            kind = regu.get("_type")
            lam = float(regu.get("strength", 0.0))
            alpha = {"ridge": 0.0, "lasso": 1.0}.get(kind, float(regu.get("alpha", 0.0)))
            w = wb[:, :nW]
            dd[:, :nW] += lam / nW * ((1.0 - alpha) * w + alpha * np.sign(w))
        dd /= float(np.sum(self.dataset.weights))
        dt = 0.1 * float(tr["maxdisplacement"])
        thr = float(tr.get("threshold", 0.0))
        new = wb.copy()
        for sp in range(wb.shape[0]):
            if np.max(np.abs(dd[sp])) < thr:
                continue
            new[sp] = wb[sp] + (-dt * dd[sp]) * dt
        return new

    # --- first conjugate-gradient step (lib_dftbp/conjgrad.F90, linemin.F90 reset + state 1) ----------
    def cg_update(self, wb, dd):
        """First iteration: search direction d0 = -g/|g|, trial step 5 |g| limited to
        MaxDisplacement / max|d0| (linemin.F90 reset / next_local state st_1), x1 = x0 + step d0."""
        tr = self.training
        nW = self.n_weights()
        dd = dd.copy()
        regu = tr.get("regularization")
        if regu:
            kind = regu.get("_type")
            lam = float(regu.get("strength", 0.0))
            alpha = {"ridge": 0.0, "lasso": 1.0}.get(kind, float(regu.get("alpha", 0.0)))
            w = wb[:, :nW]
            dd[:, :nW] += lam / nW * ((1.0 - alpha) * w + alpha * np.sign(w))
        dd /= float(np.sum(self.dataset.weights))
        maxd = float(tr["maxdisplacement"])
        thr = float(tr.get("threshold", 0.0))
        new = wb.copy()
        for sp in range(wb.shape[0]):
            g = dd[sp]
            if np.max(np.abs(g)) < thr:
                continue
            gn = np.sqrt(np.sum(g ** 2))
            d0 = -g / gn
            first = 5.0 * gn
            max_x = maxd / np.max(np.abs(d0))
            x = first if abs(first) <= max_x else max_x
            new[sp] = wb[sp] + x * d0
        return new

    # --- first L-BFGS step (lib_dftbp/lbfgs.F90:283-359, 387-403; line minimiser linemin.F90) ----------
    def lbfgs_update(self, wb, dd):
        """First iteration: no history -> direction -g; with Linemin = Yes (the default) a trial step of
        length 1 along -g/|g| capped by MaxDisplacement / max|d0|, else the direction itself rescaled so
        that its largest component is at most MaxDisplacement."""
        tr = self.training
        nW = self.n_weights()
        dd = dd.copy()
        regu = tr.get("regularization")
        if regu:
            kind = regu.get("_type")
            lam = float(regu.get("strength", 0.0))
            alpha = {"ridge": 0.0, "lasso": 1.0}.get(kind, float(regu.get("alpha", 0.0)))
            w = wb[:, :nW]
            dd[:, :nW] += lam / nW * ((1.0 - alpha) * w + alpha * np.sign(w))
        dd /= float(np.sum(self.dataset.weights))
        maxd = float(tr["maxdisplacement"])
        thr = float(tr.get("threshold", 0.0))
        linemin = str(tr.get("linemin", "yes")).strip("'\"").lower() in ("yes", "true", ".true.")
        new = wb.copy()
        for sp in range(wb.shape[0]):
            g = dd[sp]
            if np.max(np.abs(g)) < thr:
                continue
            if linemin:
                d0 = -g / np.sqrt(np.sum(g ** 2))
                max_x = maxd / np.max(np.abs(d0))
                new[sp] = wb[sp] + (1.0 if 1.0 <= max_x else max_x) * d0
            else:
                d = -g
                mx = np.max(np.abs(d))
                new[sp] = wb[sp] + (d * maxd / mx if (maxd > 0.0 and mx > maxd) else d)
        return new

    def assemble_features(self, acsf_vals):
        """features.F90:200-265: [ACSF ; ext(indices)]"""
        parts = []
        if acsf_vals is not None and acsf_vals.shape[1]:
            parts.append(acsf_vals)
        if self.ext_indices is not None and len(self.ext_indices):
            parts.append(self.dataset.ext[:, self.ext_indices.astype(int) - 1])
        return np.ascontiguousarray(np.concatenate(parts, axis=1))


def cases(mode=None, training=None, forces=None, limit=None):
    out = []
    for e in index():
        if mode is not None and e["mode"] not in mode:
            continue
        if training is not None and e["training"] not in training:
            continue
        if forces is not None and e["forces"] not in forces:
            continue
        out.append(e)
    return out[:limit] if limit else out


def allclose(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=ATOL)


def maxdiff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b))) if a.size else 0.0
