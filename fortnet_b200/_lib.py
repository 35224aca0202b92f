"""ctypes loader of libfnetgpu.so (C ABI: include/fnetgpu.h).  Fails loudly when the CUDA
library is missing -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNETGPU_LIB") or os.path.join(_HERE, "libfnetgpu.so")   # override: A/B builds
_lib = None

SYMBOLS = [
    "fnetgpu_init", "fnetgpu_finalize", "fnetgpu_last_error", "fnetgpu_synchronize",
    "fnetgpu_dataset_upload", "fnetgpu_coords_update", "fnetgpu_acsf_set", "fnetgpu_features_config",
    "fnetgpu_acsf_calculate", "fnetgpu_features_get", "fnetgpu_features_set", "fnetgpu_net_set",
    "fnetgpu_ntot", "fnetgpu_params_set", "fnetgpu_grad", "fnetgpu_predict", "fnetgpu_loss",
    "fnetgpu_forces", "fnetgpu_comm_unique_id", "fnetgpu_comm_init", "fnetgpu_set_stream",
    "fnetgpu_launch_count", "fnetgpu_profile", "fnetgpu_profile_get", "fnetgpu_kernel_name",
    "fnetgpu_max_neighbors", "fnetgpu_acsf_path_set", "fnetgpu_acsf_path_get",
    "fnetgpu_mlp_path_set", "fnetgpu_mlp_path_get", "fnetgpu_socket_step", "fnetgpu_acsf_update_calculate",
    "fnetgpu_acsf_kernel_set", "fnetgpu_acsf_kernel_get", "fnetgpu_acsf_launch_info", "fnetgpu_grad_launch_info",
    "fnetgpu_regularization_set", "fnetgpu_regularization_loss", "fnetgpu_measure_peaks",
    "fnetgpu_mg_init", "fnetgpu_mg_finalize", "fnetgpu_mg_last_error", "fnetgpu_mg_device_count", "fnetgpu_mg_context",
    "fnetgpu_mg_shard", "fnetgpu_mg_dataset_upload", "fnetgpu_mg_coords_update", "fnetgpu_mg_acsf_set",
    "fnetgpu_mg_features_config", "fnetgpu_mg_acsf_calculate", "fnetgpu_mg_features_get", "fnetgpu_mg_net_set",
    "fnetgpu_mg_params_set", "fnetgpu_mg_grad", "fnetgpu_mg_loss", "fnetgpu_mg_predict", "fnetgpu_mg_forces",
]


class FnetGpuError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FnetGpuError(
                "libfnetgpu.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C fortnet_b200/csrc`).  There is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.fnetgpu_last_error.restype = C.c_char_p
        _lib.fnetgpu_last_error.argtypes = [C.c_void_p]
        _lib.fnetgpu_kernel_name.restype = C.c_char_p
        _lib.fnetgpu_launch_count.restype = C.c_longlong
        _lib.fnetgpu_launch_count.argtypes = [C.c_void_p]
        _lib.fnetgpu_ntot.argtypes = [C.c_void_p]
    return _lib
