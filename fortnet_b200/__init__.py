"""fortnet_b200 -- B200-native hot path of Fortnet (ACSF + BPNN) behind a C ABI.

Host-side mirror of the reference interface for the hot path only; the compute lives in
libfnetgpu.so (fortnet_b200/csrc, include/fnetgpu.h).
"""
from .dataset import Dataset, frac_to_cart          # noqa: F401
from .gfunctions import GFunction, GFunctions, BOHR_PER_AA   # noqa: F401
from ._lib import FnetGpuError, LIB_PATH             # noqa: F401
from .context import Context, Acsf, Bpnn, ACTIVATIONS, LOSSES  # noqa: F401
