// mlp.cuh -- grouped per-species Behler-Parrinello subnetworks: forward, loss gradient,
// backward and weight-gradient reduction.
//
// Replaces, per atom: TNetwork_fprop / iPredict (lib_nn/network.F90:146-180, 314-347),
// TNetwork_bprop (:248-296); per structure: TBpnn_sysTrain (lib_nn/bpnn.F90:610-704); per
// dataset: TBpnn_updateGradients (:394-481) and the loss (lib_common/loss.F90:370-721).
// Atoms are processed in species-sorted order (tiles never straddle a species), one thread
// per atom, activations of the tile in shared memory as [row][atom] (conflict-free, row
// stride T+1), the species' weights in shared memory transposed to [in][out].  Weight
// gradients are accumulated per tile as dW[i][o] = sum_t a[i][t] * delta[o][t] into a
// CTA-private partial buffer; a fixed-order second stage sums the CTA partials, so the result
// is bit-reproducible.  Activations are never written to HBM (the backward kernel recomputes
// the forward pass after the per-structure loss gradient is known).
#pragma once
#include "internal.h"

template <typename real> __device__ __forceinline__ real act_f(int id, real x);
template <typename real> __device__ __forceinline__ real act_d(int id, real x, real a);

// lib_nn/transfer.F90:54-342.  act_d receives both the argument x and the activation a = f(x).
template <> __device__ __forceinline__ double act_f<double>(int id, double x) {
  switch (id) {
    case FNETGPU_ACT_GAUSSIAN: return exp(-x * x);
    case FNETGPU_ACT_RELU: return fmax(0.0, x);
    case FNETGPU_ACT_LRELU: return fmax(0.01 * x, x);
    case FNETGPU_ACT_SOFTPLUS: return log(1.0 + exp(x));
    case FNETGPU_ACT_BENT: return (sqrt(x * x + 1.0) - 1.0) / 2.0 + x;
    case FNETGPU_ACT_ATAN: return atan(x);
    case FNETGPU_ACT_SIGMOID: return 1.0 / (1.0 + exp(-x));
    case FNETGPU_ACT_HEAVISIDE: return x > 0.0 ? 1.0 : 0.0;
    case FNETGPU_ACT_TANH: return tanh(x);
    default: return x;
  }
}
template <> __device__ __forceinline__ double act_d<double>(int id, double x, double a) {
  switch (id) {
    case FNETGPU_ACT_GAUSSIAN: return -2.0 * x * a;
    case FNETGPU_ACT_RELU: return x >= 0.0 ? 1.0 : 0.0;
    case FNETGPU_ACT_LRELU: return x >= 0.0 ? 1.0 : 0.01;
    case FNETGPU_ACT_SOFTPLUS: return 1.0 / (1.0 + exp(-x));
    case FNETGPU_ACT_BENT: return x / (2.0 * sqrt(x * x + 1.0)) + 1.0;
    case FNETGPU_ACT_ATAN: return 1.0 / (x * x + 1.0);
    case FNETGPU_ACT_SIGMOID: return a * (1.0 - a);
    case FNETGPU_ACT_HEAVISIDE: return 0.0;
    case FNETGPU_ACT_TANH: return 1.0 - a * a;
    default: return 1.0;
  }
}
template <> __device__ __forceinline__ float act_f<float>(int id, float x) {
  switch (id) {
    case FNETGPU_ACT_GAUSSIAN: return __expf(-x * x);
    case FNETGPU_ACT_RELU: return fmaxf(0.0f, x);
    case FNETGPU_ACT_LRELU: return fmaxf(0.01f * x, x);
    case FNETGPU_ACT_SOFTPLUS: return log1pf(__expf(x));
    case FNETGPU_ACT_BENT: return (sqrtf(x * x + 1.0f) - 1.0f) * 0.5f + x;
    case FNETGPU_ACT_ATAN: return atanf(x);
    case FNETGPU_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-x));
    case FNETGPU_ACT_HEAVISIDE: return x > 0.0f ? 1.0f : 0.0f;
    case FNETGPU_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
template <> __device__ __forceinline__ float act_d<float>(int id, float x, float a) {
  switch (id) {
    case FNETGPU_ACT_GAUSSIAN: return -2.0f * x * a;
    case FNETGPU_ACT_RELU: return x >= 0.0f ? 1.0f : 0.0f;
    case FNETGPU_ACT_LRELU: return x >= 0.0f ? 1.0f : 0.01f;
    case FNETGPU_ACT_SOFTPLUS: return 1.0f / (1.0f + __expf(-x));
    case FNETGPU_ACT_BENT: return x / (2.0f * sqrtf(x * x + 1.0f)) + 1.0f;
    case FNETGPU_ACT_ATAN: return 1.0f / (x * x + 1.0f);
    case FNETGPU_ACT_SIGMOID: return a * (1.0f - a);
    case FNETGPU_ACT_HEAVISIDE: return 0.0f;
    case FNETGPU_ACT_TANH: return 1.0f - a * a;
    default: return 1.0f;
  }
}

// lib_common/loss.F90:217-281
__device__ __forceinline__ double loss_grad_fn(int id, double p, double t) {
  switch (id) {
    case FNETGPU_LOSS_RMS: return (p - t) / sqrt((p - t) * (p - t));
    case FNETGPU_LOSS_MAE: return (p - t) / fabs(p - t);
    case FNETGPU_LOSS_MAPE: return 100.0 * (p - t) / (t * t * fabs(p / t - 1.0));
    default: return 2.0 * (p - t);
  }
}
// transposed weight offsets inside the shared-memory copy: wT[l] is [d_l][d_{l+1}] (out fastest),
// followed by the biases in serialised order.
struct SmemNet {
  int wT[FNET_MAX_LAYERS];
  int bOff[FNET_MAX_LAYERS];
  int total;
};
__host__ __device__ inline SmemNet smem_net_layout(const NetTables &net) {
  SmemNet s;
  int off = 0;
  for (int l = 0; l + 1 < net.L; l++) { s.wT[l] = off; off += net.dims[l] * net.dims[l + 1]; }
  for (int l = 0; l < net.L; l++) { s.bOff[l] = off; off += net.dims[l]; }
  s.total = off;
  return s;
}

template <typename real>
__device__ __forceinline__ void load_weights_T(const NetTables &net, const SmemNet &sn,
                                               const real *__restrict__ wb, real *__restrict__ wsm) {
  for (int l = 0; l + 1 < net.L; l++) {
    const int din = net.dims[l], dout = net.dims[l + 1];
    const real *W = wb + net.woff[l];
    for (int e = threadIdx.x; e < din * dout; e += blockDim.x) {
      int i = e % din, o = e / din;                 // serialised: ww(i,o) at i + din*o
      wsm[sn.wT[l] + i * dout + o] = W[e];
    }
  }
  for (int l = 0; l < net.L; l++)
    for (int e = threadIdx.x; e < net.dims[l]; e += blockDim.x) wsm[sn.bOff[l] + e] = wb[net.boff[l] + e];
}

// one dense layer for the thread's atom column: out[o][t] = f(sum_i wT[i][o] in[i][t] + b[o]);
// optionally stores f'(z) (DERIV).  4 outputs per sweep over the inputs.
template <typename real, bool DERIV>
__device__ __forceinline__ void dense_layer(int din, int dout, int actId, const real *__restrict__ wT,
                                            const real *__restrict__ bias, const real *__restrict__ in,
                                            real *__restrict__ out, real *__restrict__ dout_, int TS, int t) {
  for (int o = 0; o < dout; o += 4) {
    real acc[4];
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = (o + c < dout) ? bias[o + c] : (real)0;
    if (o + 4 <= dout) {
      for (int i = 0; i < din; i++) {
        const real a = in[i * TS + t];
        const real *wr = wT + i * dout + o;
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] += wr[c] * a;
      }
    } else {
      for (int i = 0; i < din; i++) {
        const real a = in[i * TS + t];
        const real *wr = wT + i * dout + o;
#pragma unroll
        for (int c = 0; c < 4; c++) if (o + c < dout) acc[c] += wr[c] * a;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (o + c < dout) {
        const real v = act_f<real>(actId, acc[c]);
        out[(o + c) * TS + t] = v;
        if (DERIV) dout_[(o + c) * TS + t] = act_d<real>(actId, acc[c], v);
      }
  }
}

// loads the feature rows of a tile into in[f][t] (coalesced over f within an atom row)
template <typename real>
__device__ __forceinline__ void load_tile_features(int start, int count, const int *__restrict__ perm,
                                                   const real *__restrict__ feat, int nFeat, int F,
                                                   real *__restrict__ in, int TS) {
  for (int e = threadIdx.x; e < count * F; e += blockDim.x) {
    int tt = e / F, f = e % F;
    in[f * TS + tt] = feat[(size_t)nFeat * perm[start + tt] + f];
  }
}

// ------------------------------------------------------------------------------------------
// forward only: raw[atom][k] -- TBpnn_iPredict (bpnn.F90:867-900)
// smem: weights | ping [dmax][TS] | pong [dmax][TS]
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void k_mlp_fwd(int nTiles, const int *__restrict__ tiles, const int *__restrict__ perm,
                          const real *__restrict__ feat, int nFeat, const real *__restrict__ wb,
                          NetTables net, int dmax, real *__restrict__ raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x, TS = T + 1, t = threadIdx.x;
  const SmemNet sn = smem_net_layout(net);
  real *wsm = (real *)smem_raw;
  real *buf0 = wsm + ((sn.total + 1) & ~1);
  real *buf1 = buf0 + (size_t)dmax * TS;
  int curSp = -1;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int start = tiles[3 * tile], count = tiles[3 * tile + 1], sp = tiles[3 * tile + 2];
    __syncthreads();
    if (sp != curSp) { load_weights_T<real>(net, sn, wb + (size_t)net.nTot * sp, wsm); curSp = sp; }
    load_tile_features<real>(start, count, perm, feat, nFeat, net.dims[0], buf0, TS);
    __syncthreads();
    if (t < count) {
      real *in = buf0, *out = buf1;
      for (int l = 1; l < net.L; l++) {
        const int actId = (l == net.L - 1) ? FNETGPU_ACT_LINEAR : net.act;   // network.F90:391
        dense_layer<real, false>(net.dims[l - 1], net.dims[l], actId, wsm + sn.wT[l - 1], wsm + sn.bOff[l],
                                 in, out, nullptr, TS, t);
        real *tmp = in; in = out; out = tmp;
      }
      const int atom = perm[start + t];
      for (int k = 0; k < net.nOut; k++) raw[(size_t)net.nOut * atom + k] = in[k * TS + t];
    }
  }
}

// ------------------------------------------------------------------------------------------
// per-structure energy sums, loss gradients of the global targets and loss terms
// (bpnn.F90:677-684, loss.F90:370-721).  One warp per structure, lane-strided + shuffle tree.
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void k_struct_loss(int nStruct, const int *__restrict__ offsets, int nOut, int nG, int nA,
                              const real *__restrict__ raw, const double *__restrict__ gt,
                              const double *__restrict__ at, const double *__restrict__ aw,
                              const double *__restrict__ dsw, int lossId, double *__restrict__ Es,
                              double *__restrict__ gS, double *__restrict__ lossPart) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nStruct) return;
  const int b = offsets[s], e = offsets[s + 1];
  double sw = 0.0, la = 0.0;
  for (int i = b + lane; i < e; i += 32) {
    sw += aw[i];
    if (nA > 0) {
      // simple*Loss over the nA atomic targets of this atom
      double ss = 0.0;
      for (int k = 0; k < nA; k++) {
        const double pv = (double)raw[(size_t)nOut * i + nG + k], tv = at[(size_t)nA * i + k];
        switch (lossId) {
          case FNETGPU_LOSS_MAE: ss += fabs(tv - pv); break;
          case FNETGPU_LOSS_MAPE: ss += fabs((tv - pv) / tv); break;
          default: ss += (tv - pv) * (tv - pv);
        }
      }
      double li;
      switch (lossId) {
        case FNETGPU_LOSS_RMS: li = sqrt(ss / nA); break;
        case FNETGPU_LOSS_MAPE: li = 100.0 * ss / nA; break;
        default: li = ss / nA;
      }
      la += aw[i] * li;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    la += __shfl_xor_sync(0xffffffffu, la, o);
  }
  double num = 0.0, den = 0.0;
  const double w = dsw[s];
  if (nG > 0) {
    double ss = 0.0;
    for (int k = 0; k < nG; k++) {
      double ek = 0.0;
      for (int i = b + lane; i < e; i += 32) ek += (double)raw[(size_t)nOut * i + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ek += __shfl_xor_sync(0xffffffffu, ek, o);
      const double tv = gt[(size_t)nG * s + k];
      if (lane == 0) { Es[(size_t)nG * s + k] = ek; gS[(size_t)nG * s + k] = loss_grad_fn(lossId, ek, tv); }
      switch (lossId) {
        case FNETGPU_LOSS_MAE: ss += fabs(tv - ek); break;
        case FNETGPU_LOSS_MAPE: ss += fabs((tv - ek) / tv); break;
        default: ss += (tv - ek) * (tv - ek);
      }
    }
    double lg;
    switch (lossId) {
      case FNETGPU_LOSS_RMS: lg = sqrt(ss / nG); break;
      case FNETGPU_LOSS_MAPE: lg = 100.0 * ss / nG; break;
      default: lg = ss / nG;
    }
    num += w * sw * lg;
    den += w * sw;
  }
  if (nA > 0) { num += w * la; den += w * sw; }
  if (lane == 0) { lossPart[2 * s] = num; lossPart[2 * s + 1] = den; }
}

// fixed-shape reduction of the per-structure loss terms -> out[0] = numerator, out[1] = denominator
__global__ void k_loss_final(int nStruct, const double *__restrict__ lossPart, double *__restrict__ out) {
  __shared__ double sm[2][1024];
  double a = 0.0, b = 0.0;
  for (int s = threadIdx.x; s < nStruct; s += blockDim.x) { a += lossPart[2 * s]; b += lossPart[2 * s + 1]; }
  sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sm[0][threadIdx.x] += sm[0][threadIdx.x + o]; sm[1][threadIdx.x] += sm[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sm[0][0]; out[1] = sm[1][0]; }
}

// ------------------------------------------------------------------------------------------
// forward (recomputed) + backward + weight-gradient accumulation.
// smem: weights | A[rowsA][TS] activations of all layers | D[rowsD][TS] f'(z) then delta
// MODE 0: training gradient -> partials[cta][nSpecies*nTot]
// MODE 1: input gradient for forces -> dEdG[atom][k][F] (one backward sweep per output k)
// ------------------------------------------------------------------------------------------
template <typename real, int MODE>
__global__ void k_mlp_bwd(int nTiles, const int *__restrict__ tiles, const int *__restrict__ perm,
                          const real *__restrict__ feat, int nFeat, const real *__restrict__ wb,
                          NetTables net, const int *__restrict__ structOf, const int *__restrict__ offsets,
                          const double *__restrict__ gS, const double *__restrict__ at,
                          const double *__restrict__ aw, const double *__restrict__ dsw, int nG, int nA,
                          int lossId, double *__restrict__ partials, real *__restrict__ dEdG) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x, TS = T + 1, t = threadIdx.x;
  const int L = net.L;
  const SmemNet sn = smem_net_layout(net);
  real *wsm = (real *)smem_raw;
  real *A = wsm + ((sn.total + 1) & ~1);
  real *D = A + (size_t)net.rowsA * TS;            // rows: layers 1..L-1, offset aoff[l]-dims[0]
  const int d0 = net.dims[0];
  double *part = (MODE == 0) ? partials + (size_t)blockIdx.x * net.nSpecies * net.nTot : nullptr;
  int curSp = -1;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int start = tiles[3 * tile], count = tiles[3 * tile + 1], sp = tiles[3 * tile + 2];
    __syncthreads();
    if (sp != curSp) { load_weights_T<real>(net, sn, wb + (size_t)net.nTot * sp, wsm); curSp = sp; }
    load_tile_features<real>(start, count, perm, feat, nFeat, d0, A, TS);
    __syncthreads();
    const int atom = (t < count) ? perm[start + t] : -1;
    if (t < count) {
      for (int l = 1; l < L; l++) {
        const int actId = (l == L - 1) ? FNETGPU_ACT_LINEAR : net.act;
        dense_layer<real, true>(net.dims[l - 1], net.dims[l], actId, wsm + sn.wT[l - 1], wsm + sn.bOff[l],
                                A + (size_t)net.aoff[l - 1] * TS, A + (size_t)net.aoff[l] * TS,
                                D + (size_t)(net.aoff[l] - d0) * TS, TS, t);
      }
    }
    const int nSweeps = (MODE == 0) ? 1 : net.nOut;
    for (int sweep = 0; sweep < nSweeps; sweep++) {
      if (t < count) {
        // output layer (linear): delta_L = lossgrad (x) f' (network.F90:276)
        real *dL = D + (size_t)(net.aoff[L - 1] - d0) * TS;
        if (MODE == 0) {
          const int s = structOf[atom];
          const double scale = dsw[s] * aw[atom] / (double)(offsets[s + 1] - offsets[s]);   // bpnn.F90:446,698
          for (int k = 0; k < net.nOut; k++) {
            double g;
            if (k < nG) g = gS[(size_t)nG * s + k];
            else g = loss_grad_fn(lossId, (double)A[(size_t)(net.aoff[L - 1] + k) * TS + t], at[(size_t)nA * atom + (k - nG)]);
            dL[k * TS + t] = (real)(g * scale);
          }
        } else {
          for (int k = 0; k < net.nOut; k++) dL[k * TS + t] = (k == sweep) ? (real)1 : (real)0;
        }
        // hidden layers: delta_l = (W_l delta_{l+1}) * f'(z_l)  (network.F90:282-288); in MODE 1 the
        // stored f' must survive for the next sweep, so deltas go to the A rows of the same layer
        // (activations of hidden layers are not needed for input gradients).
        for (int l = L - 1; l >= (MODE == 0 ? 2 : 1); l--) {
          const int din = net.dims[l - 1], dout = net.dims[l];
          const real *wT = wsm + sn.wT[l - 1];
          const real *dn = (MODE == 1 && l < L - 1) ? A + (size_t)net.aoff[l] * TS : D + (size_t)(net.aoff[l] - d0) * TS;
          for (int i = 0; i < din; i += 4) {
            real acc[4] = {0, 0, 0, 0};
            for (int o = 0; o < dout; o++) {
              const real dv = dn[o * TS + t];
#pragma unroll
              for (int c = 0; c < 4; c++) if (i + c < din) acc[c] += wT[(i + c) * dout + o] * dv;
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
              if (i + c < din) {
                if (l - 1 == 0) {
                  if (MODE == 1) dEdG[((size_t)net.nOut * atom + sweep) * d0 + (i + c)] = acc[c];
                } else if (MODE == 0) {
                  real *dp = D + (size_t)(net.aoff[l - 1] - d0) * TS + (size_t)(i + c) * TS + t;
                  *dp = acc[c] * (*dp);
                } else {
                  const real fp = D[(size_t)(net.aoff[l - 1] - d0) * TS + (size_t)(i + c) * TS + t];
                  A[(size_t)(net.aoff[l - 1] + i + c) * TS + t] = acc[c] * fp;
                }
              }
          }
        }
      }
    }
    if (MODE == 0) {
      __syncthreads();
      // dW_l[i][o] = sum_t a_{l}[i][t] delta_{l+1}[o][t]; db_{l+1}[o] = sum_t delta_{l+1}[o][t]
      double *ps = part + (size_t)net.nTot * sp;
      for (int l = 0; l + 1 < L; l++) {
        const int din = net.dims[l], dout = net.dims[l + 1];
        const real *Al = A + (size_t)net.aoff[l] * TS;
        const real *Dn = D + (size_t)(net.aoff[l + 1] - d0) * TS;
        for (int e = t; e < din * dout; e += T) {
          const int i = e % din, o = e / din;
          const real *ar = Al + (size_t)i * TS, *dr = Dn + (size_t)o * TS;
          double s = 0.0;
          for (int tt = 0; tt < count; tt++) s += (double)ar[tt] * (double)dr[tt];
          ps[net.woff[l] + e] += s;
        }
        for (int o = t; o < dout; o += T) {
          const real *dr = Dn + (size_t)o * TS;
          double s = 0.0;
          for (int tt = 0; tt < count; tt++) s += (double)dr[tt];
          ps[net.boff[l + 1] + o] += s;
        }
      }
    }
  }
}

// dd[p] = sum_cta partials[cta][p] in fixed order (deterministic)
__global__ void k_grad_reduce(int nCta, int nDD, const double *__restrict__ partials, double *__restrict__ dd) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nDD) return;
  double s = 0.0;
  for (int c = 0; c < nCta; c++) s += partials[(size_t)c * nDD + p];
  dd[p] = s;
}
