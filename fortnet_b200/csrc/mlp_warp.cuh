// mlp_warp.cuh -- latency path of the subnetworks: ONE WARP PER ATOM, forward sweep and the input gradients
// dE_k/dG in one kernel.
//
// Same mathematics and reference sites as mlp.cuh / mlp_mma.cuh (TNetwork_fprop, lib_nn/network.F90:146-180;
// the Jacobian of TBpnn_iJacobian / nJacobian, lib_nn/bpnn.F90:904-997, as one reverse sweep per output).  The
// throughput kernels process rounds of 64 atoms per CTA and need ~20 us for a round whatever its population: an MD /
// i-PI step of one 64..200-atom cell (fnetgpu_socket_step, prg_fnet/fortnet.F90:430-609) is one or three rounds on
// one to three SMs.  Here an atom is a matrix-vector chain of its own: lane = output neuron (forward) or input
// neuron (reverse), the species' weights in shared memory as [out][in] with an ODD row stride (conflict-free for both
// directions: lanes differ in `out` forward, in `in` backward), 4 atoms per CTA, so a 64-atom cell spreads over
// 16 SMs and the kernel takes a few microseconds.  FP64 only (the socket fast path is the parity mode).
#pragma once
#include "mlp.cuh"

#define FNET_WARP_WPB 4            // atoms (warps) per CTA

struct WarpMlpLayout {
  int wOff[FNET_MAX_LAYERS], ld[FNET_MAX_LAYERS], bOff[FNET_MAX_LAYERS];
  int wTotal, maxDim;
};
__host__ __device__ inline WarpMlpLayout warp_mlp_layout(const NetTables &net) {
  WarpMlpLayout m;
  int off = 0;
  m.maxDim = 0;
  for (int l = 0; l < net.L; l++) m.maxDim = net.dims[l] > m.maxDim ? net.dims[l] : m.maxDim;
  for (int l = 0; l + 1 < net.L; l++) {
    m.ld[l] = net.dims[l] | 1;
    m.wOff[l] = off;
    off += net.dims[l + 1] * m.ld[l];
  }
  m.bOff[0] = off;
  for (int l = 1; l < net.L; l++) { m.bOff[l] = off; off += net.dims[l]; }
  m.wTotal = (off + 1) & ~1;
  return m;
}
__host__ inline size_t bpnn_warp_smem_bytes(const NetTables &net) {
  const WarpMlpLayout m = warp_mlp_layout(net);
  return ((size_t)m.wTotal + FNET_EXP_TAB_N + (size_t)FNET_WARP_WPB * (2 * net.rowsA + 2 * m.maxDim)) * sizeof(double);
}

// tiles: (start, count <= FNET_WARP_WPB, species) triples of the species-sorted atom order `perm`
__global__ void __launch_bounds__(FNET_WARP_WPB * 32)
k_bpnn_warp(const int *__restrict__ tiles, const int *__restrict__ perm, const double *__restrict__ feat, int nFeat,
            const double *__restrict__ wb, NetTables net, double *__restrict__ raw, double *__restrict__ dEdG) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WarpMlpLayout m = warp_mlp_layout(net);
  const int L = net.L, d0 = net.dims[0], nOut = net.nOut;
  double *wsm = (double *)smem_raw;
  double *etab = wsm + m.wTotal;
  const int per = 2 * net.rowsA + 2 * m.maxDim;
  double *a = etab + FNET_EXP_TAB_N + (size_t)warp * per;      // activations a_0 .. a_{L-1}
  double *fp = a + net.rowsA;                                   // f'(z_l) at the same offsets
  double *dA = fp + net.rowsA, *dB = dA + m.maxDim;             // deltas of the reverse sweep (ping-pong)
  const int start = tiles[3 * blockIdx.x], count = tiles[3 * blockIdx.x + 1], sp = tiles[3 * blockIdx.x + 2];
  const int atom = warp < count ? perm[start + warp] : -1;
  FNET_PDL_TRIGGER();
  {   // weights, biases, exp table: constants of the step -- staged while the ACSF kernel in front may still be running
    const double *W = wb + (size_t)net.nTot * sp;
    for (int l = 0; l + 1 < L; l++) {
      const int din = net.dims[l], dout = net.dims[l + 1], ld = m.ld[l];
      const double *Wl = W + net.woff[l];                        // ww(i, o) at i + din * o (network.F90:413-419)
      double *dst = wsm + m.wOff[l];
      for (int e = threadIdx.x; e < din * dout; e += blockDim.x) {
        const int o = e / din, i = e - o * din;
        dst[o * ld + i] = Wl[e];
      }
    }
    for (int l = 1; l < L; l++)
      for (int e = threadIdx.x; e < net.dims[l]; e += blockDim.x) wsm[m.bOff[l] + e] = W[net.boff[l] + e];
    for (int e = threadIdx.x; e < FNET_EXP_TAB_N; e += blockDim.x) etab[e] = fnet_exp_tab_d[e];
  }
  FNET_PDL_WAIT();            // the features come from the kernel in front
  if (atom >= 0)
    for (int f = lane; f < d0; f += 32) a[f] = feat[(size_t)nFeat * atom + f];
  __syncthreads();
  if (atom < 0) return;
  // ---- forward: lane = output neuron ----
  for (int l = 1; l < L; l++) {
    const int din = net.dims[l - 1], dout = net.dims[l], ld = m.ld[l - 1];
    const bool last = (l == L - 1);
    const double *in = a + net.aoff[l - 1];
    double *out = a + net.aoff[l];
    for (int o = lane; o < dout; o += 32) {
      const double *w = wsm + m.wOff[l - 1] + o * ld;
      double z0 = wsm[m.bOff[l] + o], z1 = 0.0, z2 = 0.0, z3 = 0.0;
      int i = 0;
      for (; i + 3 < din; i += 4) {
        z0 = fma(w[i], in[i], z0); z1 = fma(w[i + 1], in[i + 1], z1);
        z2 = fma(w[i + 2], in[i + 2], z2); z3 = fma(w[i + 3], in[i + 3], z3);
      }
      for (; i < din; i++) z0 = fma(w[i], in[i], z0);
      const double z = (z0 + z1) + (z2 + z3);
      if (last) out[o] = z;                                      // network.F90:391: the output layer is linear
      else {
        const double v = net.act == FNETGPU_ACT_TANH ? fnet_tanh_em1(z, etab) : act_f<double>(net.act, z);
        out[o] = v;
        fp[net.aoff[l] + o] = act_d<double>(net.act, z, v);
      }
    }
    __syncwarp();
  }
  for (int k = lane; k < nOut; k += 32) raw[(size_t)nOut * atom + k] = a[net.aoff[L - 1] + k];
  if (!dEdG) return;
  // ---- input gradients: one reverse sweep per output, lane = input neuron of the layer ----
  for (int sweep = 0; sweep < nOut; sweep++) {
    const double *dn = nullptr;
    double *dcur = dA;
    for (int l = L - 2; l >= 0; l--) {
      const int din = net.dims[l], dout = net.dims[l + 1], ld = m.ld[l];
      const double *Wl = wsm + m.wOff[l];
      for (int i = lane; i < din; i += 32) {
        double s;
        if (l == L - 2) s = Wl[sweep * ld + i];                  // delta of the linear output layer = e_sweep
        else {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          int o = 0;
          for (; o + 3 < dout; o += 4) {
            s0 = fma(Wl[o * ld + i], dn[o], s0); s1 = fma(Wl[(o + 1) * ld + i], dn[o + 1], s1);
            s2 = fma(Wl[(o + 2) * ld + i], dn[o + 2], s2); s3 = fma(Wl[(o + 3) * ld + i], dn[o + 3], s3);
          }
          for (; o < dout; o++) s0 = fma(Wl[o * ld + i], dn[o], s0);
          s = (s0 + s1) + (s2 + s3);
        }
        if (l >= 1) dcur[i] = s * fp[net.aoff[l] + i];           // network.F90:282-288
        else dEdG[((size_t)nOut * atom + sweep) * d0 + i] = s;
      }
      __syncwarp();
      dn = dcur;
      dcur = (dcur == dA) ? dB : dA;
    }
  }
}
