// kernels_mma.cu -- the instantiations of k_bpnn_mma (mlp_mma.cuh), in a translation unit of their own (the code
// generated for the kernel then depends on its own source only; see acsf_lean.cuh).
#define FNET_KERNEL_TU
#define FNET_DEFINE_MMA_KERNEL
#include "internal.h"
#include "mlp.cuh"
#include "mlp_mma.cuh"

template <int MODE, int NSLOT, int FUSED>
static MmaKernelT mma_pick(int fch) {
  if (fch == 1) return k_bpnn_mma<MODE, NSLOT, 1, FUSED>;
  if (fch == 2) return k_bpnn_mma<MODE, NSLOT, 2, FUSED>;
  return nullptr;
}
// the variants the launchers ask for (fnetgpu.cu): forward / input gradients with one slot, the gradient kernel with
// 4 or FNET_MMA_MAXSLOTS gradient tiles per warp and the three ways of forming the per-structure sums
MmaKernelT fnet_mma_kernel(int MODE, int NSLOT, int FCH, int FUSED) {
  if (MODE == 2 && NSLOT == 1 && FUSED == 0) return mma_pick<2, 1, 0>(FCH);
  if (MODE == 1 && NSLOT == 1 && FUSED == 0) return mma_pick<1, 1, 0>(FCH);
  if (MODE == 0 && NSLOT == 4) return FUSED == 0 ? mma_pick<0, 4, 0>(FCH) : (FUSED == 1 ? mma_pick<0, 4, 1>(FCH) : mma_pick<0, 4, 2>(FCH));
  if (MODE == 0 && NSLOT == FNET_MMA_MAXSLOTS)
    return FUSED == 0 ? mma_pick<0, FNET_MMA_MAXSLOTS, 0>(FCH) : (FUSED == 1 ? mma_pick<0, FNET_MMA_MAXSLOTS, 1>(FCH) : mma_pick<0, FNET_MMA_MAXSLOTS, 2>(FCH));
  return nullptr;
}
