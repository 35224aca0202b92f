"""Driver-side initialisation restated for the test harness: the RANLUX generator
(lib_dftbp/ranlux.F90:127-188 init, :291-369 workhorse), the truncated-normal "Xavier" weights
(lib_common/random.F90:122-181) and the order in which a fresh BPNN draws them
(lib_nn/bpnn.F90:130-137 -> lib_nn/network.F90:95-118 -> lib_nn/layer.F90:55-90).  With these the
reference's fresh-training goldens (no stored netstat) replay without Fortran.
Test infrastructure only -- initialisation is driver-side and not part of the GPU hot path.
"""
import math

import numpy as np

_NDSKIP = (0, 24, 73, 199, 365)
_ITWO24 = 2 ** 24
_ICONS = 2147483563
_MASKLO = _ITWO24 - 1


def _trunc_div(a, b):
    """Fortran integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


class Ranlux:
    """ranlux.F90: subtract-with-borrow generator, luxury level 3 by default (initprogram.F90:392)."""

    def __init__(self, luxlev=3, seed=314159265):
        jseed = seed if seed > 0 else 314159265
        self.nskip = _NDSKIP[luxlev]
        self.in24 = 0
        self.twom24 = 1.0
        self.iseeds = [0] * 25            # 1-based like the reference
        self.next = [0] * 25
        for ii in range(1, 25):
            self.twom24 *= 0.5
            kk = _trunc_div(jseed, 53668)
            jseed = 40014 * (jseed - kk * 53668) - kk * 12211
            if jseed < 0:
                jseed += _ICONS
            self.iseeds[ii] = jseed % _ITWO24
            self.next[ii] = ii - 1
        self.twom12 = self.twom24 * 4096.0
        self.next[1] = 24
        self.i24, self.j24 = 24, 10
        self.icarry = 1 if (self.iseeds[24] & ~_MASKLO) != 0 else 0

    def _step(self):
        iuni = self.iseeds[self.j24] - self.iseeds[self.i24] - self.icarry
        if iuni & ~_MASKLO:               # negative (two's complement high bits) -> borrow
            iuni &= _MASKLO
            self.icarry = 1
        else:
            self.icarry = 0
        self.iseeds[self.i24] = iuni
        self.i24 = self.next[self.i24]
        self.j24 = self.next[self.j24]
        return iuni

    def random(self, n):
        out = np.empty(n)
        for k in range(n):
            iuni = self._step()
            uni = float(iuni) * self.twom24
            if uni < self.twom12:
                uni += (float(self.iseeds[self.j24]) * self.twom24) * self.twom24
                if uni <= 0.0:
                    uni = self.twom24 * self.twom24
            out[k] = uni
            self.in24 += 1
            if self.in24 == 24:
                self.in24 = 0
                for _ in range(self.nskip):
                    self._step()
        return out


def normal_xavier(rng, n, n_last, n_current, gain=1.0, minval=0.1):
    """random.F90:122-181 -- n values of the layer's weight array in generation (column-major) order"""
    rnd = rng.random(n)
    var = gain ** 2 * 2.0 / float(n_last + n_current)
    bound = math.sqrt(-2.0 * var * math.log(minval * math.sqrt(2.0 * math.pi * var)))
    rnd = (rnd - 0.5) * 2.0 * bound
    return np.exp(-0.5 * rnd ** 2 / var) / math.sqrt(2.0 * math.pi * var)


def initial_parameters(seed, dims, n_species):
    """serialised parameters (nSpecies, nTot) of a freshly initialised BPNN: per species, per layer
    ww(d_l, d_{l+1}) column-major incl. the unused last-layer ww(d_L, 1) -- it IS drawn and stored --
    then all biases = 0 (layer.F90:78-87)"""
    rng = Ranlux(3, int(seed))
    dims = [int(d) for d in dims]
    out = []
    for _ in range(n_species):
        ws = []
        for l in range(len(dims)):
            nxt = dims[l + 1] if l + 1 < len(dims) else 1
            ws.append(normal_xavier(rng, dims[l] * nxt, nxt, dims[l]))
        out.append(np.concatenate(ws + [np.zeros(sum(dims))]))
    return np.asarray(out)
