// kernels_lean.cu -- the instantiations of k_acsf_lean (acsf_lean.cuh), in a translation unit of their own: the code
// generated for this kernel then depends on its own source only (see the note above the kernel).
#define FNET_KERNEL_TU
#define FNET_DEFINE_LEAN_KERNEL
#include "internal.h"
#include "cells.cuh"
#include "acsf.cuh"
#include "acsf_lean.cuh"

template <int NL, int NC, int PATH, bool F32A>
static LeanKernelT lean_pick(bool sorted, int G) {
  if (PATH == FNET_PATH_DIRECT) return sorted ? (LeanKernelT)k_acsf_lean<NL, NC, PATH, true, 1, F32A> : (LeanKernelT)k_acsf_lean<NL, NC, PATH, false, 1, F32A>;
  if (sorted) {
    if (G == 2) return k_acsf_lean<NL, NC, PATH, true, 2, F32A>;
    if (G == 1) return k_acsf_lean<NL, NC, PATH, true, 1, F32A>;
    return nullptr;
  }
  if (G == 4) return k_acsf_lean<NL, NC, PATH, false, 4, F32A>;
  if (G == 2) return k_acsf_lean<NL, NC, PATH, false, 2, F32A>;
  if (G == 1) return k_acsf_lean<NL, NC, PATH, false, 1, F32A>;
  return nullptr;
}
template <int NL, int NC>
static LeanKernelT lean_pick_path(int path, bool sorted, int G, bool f32a) {
  if (G != 1 && path == FNET_PATH_DIRECT) return nullptr;
  switch (path) {
    case FNET_PATH_STRUCT: return f32a ? lean_pick<NL, NC, FNET_PATH_STRUCT, true>(sorted, G) : lean_pick<NL, NC, FNET_PATH_STRUCT, false>(sorted, G);
    case FNET_PATH_STAGED: return f32a ? lean_pick<NL, NC, FNET_PATH_STAGED, true>(sorted, G) : lean_pick<NL, NC, FNET_PATH_STAGED, false>(sorted, G);
    case FNET_PATH_DIRECT: return f32a ? lean_pick<NL, NC, FNET_PATH_DIRECT, true>(sorted, G) : lean_pick<NL, NC, FNET_PATH_DIRECT, false>(sorted, G);
  }
  return nullptr;
}
// the variants the launcher asks for (fnetgpu.cu, launch_acsf_values): (NL, NC) = (2, 1), (2, 2), (1, 4)
LeanKernelT fnet_lean_kernel(int NL, int NC, int PATH, bool SORTED, int G, bool F32A) {
  if (NL == 2 && NC == 1) return lean_pick_path<2, 1>(PATH, SORTED, G, F32A);
  if (NL == 2 && NC == 2) return lean_pick_path<2, 2>(PATH, SORTED, G, F32A);
  if (NL == 1 && NC == 4) return lean_pick_path<1, 4>(PATH, SORTED, G, F32A);
  return nullptr;
}
