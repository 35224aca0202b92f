// mlp.cuh -- grouped per-species Behler-Parrinello subnetworks: forward, loss gradient,
// backward and weight-gradient reduction.
//
// Replaces, per atom: TNetwork_fprop / iPredict (lib_nn/network.F90:146-180, 314-347),
// TNetwork_bprop (:248-296); per structure: TBpnn_sysTrain (lib_nn/bpnn.F90:610-704); per
// dataset: TBpnn_updateGradients (:394-481) and the loss (lib_common/loss.F90:370-721).
// Atoms are processed in species-sorted order in tiles of T atoms that never straddle a species.
// Every layer of a tile is a small dense contraction Z[T x dout] = A[T x din] W[din x dout]
// done as a register-tiled FMA GEMM on the CUDA cores (layer widths of 2..100 are far too
// narrow for tcgen05 tiles and FP64/FP32 parity rules out TF32/BF16): each thread owns a
// 4-atom x 4-output tile, activations sit in shared memory as [row][atom] (row stride T+2
// doubles / T+4 floats: 16-byte aligned vector loads, conflict-free), the species' weights as
// [in][out] with a padded row stride.  Per 16 FMAs a thread issues two 16-byte activation loads
// and two 16-byte weight loads, so the kernel is bound by the FP64 (or FP32) FMA pipe, not by
// shared-memory bandwidth.  Backward: delta_l = (W_l delta_{l+1}) * f'(z_l) with the same
// tiling; weight gradients dW[i][o] = sum_t a[i][t] delta[o][t] as a third GEMM whose k-range
// (the atoms) is split over adjacent lanes and combined by shuffles, accumulated in a
// CTA-private shared-memory copy of the serialised gradient and written once per CTA; a
// fixed-order second stage sums the CTA partials, so the result is bit-reproducible (no float
// atomics).  Activations never go to HBM (the backward kernel recomputes the forward pass
// after the per-structure loss gradient is known).
#pragma once
#include "internal.h"
#include "fmath.cuh"

template <typename real> __device__ __forceinline__ real act_f(int id, real x);
template <typename real> __device__ __forceinline__ real act_d(int id, real x, real a);

// lib_nn/transfer.F90:54-342.  act_d receives both the argument x and the activation a = f(x).
template <> __device__ __forceinline__ double act_f<double>(int id, double x) {
  switch (id) {
    case FNETGPU_ACT_GAUSSIAN: return fnet_exp_lat(-x * x);
    case FNETGPU_ACT_RELU: return fmax(0.0, x);
    case FNETGPU_ACT_LRELU: return fmax(0.01 * x, x);
    case FNETGPU_ACT_SOFTPLUS: return log(1.0 + exp(x));
    case FNETGPU_ACT_BENT: return (sqrt(x * x + 1.0) - 1.0) / 2.0 + x;
    case FNETGPU_ACT_ATAN: return atan(x);
    case FNETGPU_ACT_SIGMOID: return fnet_rcp(1.0 + fnet_exp_lat(fmin(-x, 700.0)));
    case FNETGPU_ACT_HEAVISIDE: return x > 0.0 ? 1.0 : 0.0;
    case FNETGPU_ACT_TANH: return fnet_tanh(x);
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

// ------------------------------------------------------------------------------------------
// shared-memory layout of one species' parameters (padded):
//   W_l  [d_l][S_l], S_l = roundup4(d_{l+1}) + pad, element (i, o) at wOff[l] + i*S_l + o,
//   b_l  [roundup4(d_l)] at bOff[l]; padding is zero so padded outputs evaluate to f(0).
// ------------------------------------------------------------------------------------------
struct MlpLayout {
  int wOff[FNET_MAX_LAYERS], wS[FNET_MAX_LAYERS], bOff[FNET_MAX_LAYERS];
  int total;       // elements
};
template <typename real> struct VecPad;                       // row padding that keeps 16-byte alignment
template <> struct VecPad<double> { static const int value = 2; };
template <> struct VecPad<float> { static const int value = 4; };

template <typename real>
__host__ __device__ inline MlpLayout mlp_layout(const NetTables &net) {
  MlpLayout m;
  int off = 0;
  for (int l = 0; l + 1 < net.L; l++) {
    m.wS[l] = ((net.dims[l + 1] + 3) & ~3) + VecPad<real>::value;
    m.wOff[l] = off;
    off += net.dims[l] * m.wS[l];
  }
  for (int l = 0; l < net.L; l++) { m.bOff[l] = off; off += (net.dims[l] + 3) & ~3; }
  m.total = (off + 3) & ~3;
  return m;
}

template <typename real>
__device__ __forceinline__ void load_weights_padded(const NetTables &net, const MlpLayout &m,
                                                    const real *__restrict__ wb, real *__restrict__ wsm) {
  for (int e = threadIdx.x; e < m.total; e += blockDim.x) wsm[e] = (real)0;
  __syncthreads();
  for (int l = 0; l + 1 < net.L; l++) {
    const int din = net.dims[l], dout = net.dims[l + 1];
    const real *W = wb + net.woff[l];
    for (int e = threadIdx.x; e < din * dout; e += blockDim.x) {
      const int i = e % din, o = e / din;           // serialised: ww(i,o) at i + din*o (network.F90:413-419)
      wsm[m.wOff[l] + i * m.wS[l] + o] = W[e];
    }
  }
  for (int l = 0; l < net.L; l++)
    for (int e = threadIdx.x; e < net.dims[l]; e += blockDim.x) wsm[m.bOff[l] + e] = wb[net.boff[l] + e];
}

// 4-wide shared-memory vector access (16-byte aligned addresses)
__device__ __forceinline__ void ld4(const double *p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void ld4(const float *p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4 *>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void st4(double *p, const double (&v)[4]) {
  *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void ld2(const double *p, double (&v)[2]) {
  const double2 a = *reinterpret_cast<const double2 *>(p); v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void ld2(const float *p, float (&v)[2]) {
  const float2 a = *reinterpret_cast<const float2 *>(p); v[0] = a.x; v[1] = a.y;
}

// asynchronous global -> shared copies (LDGSTS): the features of the NEXT tile stream into the
// second input buffer while the current tile is being computed
__device__ __forceinline__ void cp_async_elem(double *smem, const double *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float *smem, const float *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// features of a tile -> A0[f][t]: one warp per atom row (coalesced over f), transposed on the way
// into shared memory; atoms beyond `count` are 0
template <typename real>
__device__ __forceinline__ void prefetch_tile_features(int start, int count, int T, const int *__restrict__ perm,
                                                       const real *__restrict__ feat, int nFeat, int F,
                                                       real *__restrict__ A0, int TS) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int tt = wib; tt < T; tt += nw) {
    if (tt < count) {
      const real *row = feat + (size_t)nFeat * perm[start + tt];
      for (int f = lane; f < F; f += 32) cp_async_elem(A0 + (size_t)f * TS + tt, row + f);
    } else {
      for (int f = lane; f < F; f += 32) A0[(size_t)f * TS + tt] = (real)0;
    }
  }
  cp_async_commit();
}

// Forward layer: out[o][t] = f(sum_i W[i][o] in[i][t] + b[o]); DERIV also stores f'(z).
// Work items e = cg*RG + rg: rg = 4-atom group, cg = 4-output group.
template <typename real, bool DERIV>
__device__ __forceinline__ void layer_forward(int din, int dout, int actId, const real *__restrict__ W, int S,
                                              const real *__restrict__ bias, const real *__restrict__ in,
                                              real *__restrict__ out, real *__restrict__ dact, int T, int TS) {
  const int RG = T >> 2, CG = (dout + 3) >> 2;
  for (int e = threadIdx.x; e < RG * CG; e += blockDim.x) {
    const int rg = e % RG, cg = e / RG;
    real acc[4][4];   // [out c][atom a]
    {
      real b4[4];
      ld4(bias + 4 * cg, b4);
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int a = 0; a < 4; a++) acc[c][a] = b4[c];
    }
    const real *ip = in + 4 * rg;
    const real *wp = W + 4 * cg;
#pragma unroll 4
    for (int i = 0; i < din; i++) {
      real a4[4], w4[4];
      ld4(ip + (size_t)i * TS, a4);
      ld4(wp + (size_t)i * S, w4);
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int a = 0; a < 4; a++) acc[c][a] += w4[c] * a4[a];
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int o = 4 * cg + c;
      if (o < dout) {
        real v4[4], d4[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          v4[a] = act_f<real>(actId, acc[c][a]);
          if (DERIV) d4[a] = act_d<real>(actId, acc[c][a], v4[a]);
        }
        st4(out + (size_t)o * TS + 4 * rg, v4);
        if (DERIV) st4(dact + (size_t)o * TS + 4 * rg, d4);
      }
    }
  }
}

// Backward layer: res[i][t] = sum_o W[i][o] dn[o][t]  (network.F90:282-288), 4 inputs x 4 atoms per item.
// store(i, rg, values[4 atoms]) consumes the result.
template <typename real, typename Store>
__device__ __forceinline__ void layer_backward(int din, int dout, const real *__restrict__ W, int S,
                                               const real *__restrict__ dn, int T, int TS, Store store) {
  const int RG = T >> 2, IG = (din + 3) >> 2;
  for (int e = threadIdx.x; e < RG * IG; e += blockDim.x) {
    const int rg = e % RG, ig = e / RG;
    real acc[4][4];   // [in c][atom a]
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int a = 0; a < 4; a++) acc[c][a] = (real)0;
    const real *wr[4];
#pragma unroll
    for (int c = 0; c < 4; c++) wr[c] = W + (size_t)min(4 * ig + c, din - 1) * S;
    const real *dp = dn + 4 * rg;
#pragma unroll 4
    for (int o = 0; o < dout; o++) {
      real d4[4];
      ld4(dp + (size_t)o * TS, d4);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const real wv = wr[c][o];
#pragma unroll
        for (int a = 0; a < 4; a++) acc[c][a] += wv * d4[a];
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (4 * ig + c < din) store(4 * ig + c, rg, acc[c]);
  }
}

// Weight gradient of one layer: G[woff + i + din*o] += sum_t a[i][t] dn[o][t] and
// G[boff + o] += sum_t dn[o][t].  Items = (4-input group, 4-output group); the atom range is
// split over KS adjacent lanes (KS a power of two, items*KS <= blockDim whenever possible) and
// combined by xor shuffles in a fixed order.
template <typename real>
__device__ __forceinline__ void layer_wgrad(int din, int dout, const real *__restrict__ Al,
                                            const real *__restrict__ Dn, int T, int TS,
                                            double *__restrict__ Gw, double *__restrict__ Gb) {
  const int IG = (din + 3) >> 2, OG = (dout + 3) >> 2, nIt = IG * OG;
  int KS = 1;
  while (KS < 8 && nIt * (KS << 1) <= (int)blockDim.x && (KS << 2) <= T) KS <<= 1;
  const int total = nIt * KS;
  for (int base = 0; base < total; base += blockDim.x) {   // uniform trip count: shuffles need whole warps
    const int e = base + threadIdx.x;
    const bool valid = e < total;
    const int kpart = e % KS, it = valid ? e / KS : 0;
    const int ig = it % IG, og = it / IG;
    double acc[4][4];   // [in][out]
#pragma unroll
    for (int ci = 0; ci < 4; ci++)
#pragma unroll
      for (int co = 0; co < 4; co++) acc[ci][co] = 0.0;
    const real *ar[4], *dr[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      ar[c] = Al + (size_t)min(4 * ig + c, din - 1) * TS;
      dr[c] = Dn + (size_t)min(4 * og + c, dout - 1) * TS;
    }
    for (int t = 2 * kpart; t < T; t += 2 * KS) {
      real a2[4][2], d2[4][2];
#pragma unroll
      for (int c = 0; c < 4; c++) { ld2(ar[c] + t, a2[c]); ld2(dr[c] + t, d2[c]); }
#pragma unroll
      for (int ci = 0; ci < 4; ci++)
#pragma unroll
        for (int co = 0; co < 4; co++)
          acc[ci][co] += (double)a2[ci][0] * (double)d2[co][0] + (double)a2[ci][1] * (double)d2[co][1];
    }
    for (int off = 1; off < KS; off <<= 1)
#pragma unroll
      for (int ci = 0; ci < 4; ci++)
#pragma unroll
        for (int co = 0; co < 4; co++) acc[ci][co] += __shfl_xor_sync(0xffffffffu, acc[ci][co], off);
    if (valid && kpart == 0) {
#pragma unroll
      for (int ci = 0; ci < 4; ci++)
#pragma unroll
        for (int co = 0; co < 4; co++) {
          const int i = 4 * ig + ci, o = 4 * og + co;
          if (i < din && o < dout) Gw[i + din * o] += acc[ci][co];
        }
    }
  }
  for (int o = threadIdx.x; o < dout; o += blockDim.x) {
    const real *dr = Dn + (size_t)o * TS;
    double s = 0.0;
    for (int t = 0; t < T; t++) s += (double)dr[t];
    Gb[o] += s;
  }
}

// ------------------------------------------------------------------------------------------
// The subnetwork kernel.  MODE 0: training gradient -> partials[cta][nSpecies*nTot]
// (TBpnn_sysTrain + updateGradients, bpnn.F90:394-481,610-704); MODE 1: input gradient
// dEdG[atom][k][F] for the forces (one reverse sweep per output k instead of the forward-mode
// Jacobian of network.F90:183-244); MODE 2: forward only, raw[atom][k] (iPredict, bpnn.F90:867-900).
// smem: weights | input buffer 0 | A[rowsA][TS] (input buffer 1 + activations) | D[rowsD][TS] f'(z) then delta | G[nTot]
// Each CTA walks a contiguous range of tiles, so it changes species at most nSpecies-1 times.
// ------------------------------------------------------------------------------------------
template <typename real, int MODE>
__global__ void __launch_bounds__(256)
k_bpnn(int nTiles, const int *__restrict__ tiles, const int *__restrict__ perm, const real *__restrict__ feat,
       int nFeat, const real *__restrict__ wb, NetTables net, int T, int gInSmem,
       const int *__restrict__ structOf, const int *__restrict__ offsets, const double *__restrict__ gS,
       const double *__restrict__ at, const double *__restrict__ aw, const double *__restrict__ dsw, int nG,
       int nA, int lossId, double *__restrict__ partials, real *__restrict__ dEdG, real *__restrict__ raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TS = T + VecPad<real>::value;
  const int L = net.L, d0 = net.dims[0];
  const MlpLayout m = mlp_layout<real>(net);
  real *wsm = (real *)smem_raw;
  real *A0b = wsm + m.total;                       // two input buffers [d0][TS] (double-buffered prefetch)
  real *A = A0b + (size_t)d0 * TS;                 // rows of layer l >= 1 at aoff[l] (row 0..d0-1 = buffer 1)
  real *D = A + (size_t)net.rowsA * TS;            // rows of layers 1..L-1 at (aoff[l] - d0)
  double *Gs = (double *)(D + (MODE == 2 ? 0 : (size_t)(net.rowsA - d0) * TS));
  const int nDD = net.nSpecies * net.nTot;
  double *G = nullptr;
  const int tile0 = (int)(((long long)nTiles * blockIdx.x) / gridDim.x);
  const int tile1 = (int)(((long long)nTiles * (blockIdx.x + 1)) / gridDim.x);
  int curSp = -1, buf = 0;
  if (tile0 < tile1)
    prefetch_tile_features<real>(tiles[3 * tile0], tiles[3 * tile0 + 1], T, perm, feat, nFeat, d0, A0b, TS);
  for (int tile = tile0; tile < tile1; tile++, buf ^= 1) {
    const int start = tiles[3 * tile], count = tiles[3 * tile + 1], sp = tiles[3 * tile + 2];
    real *A0 = A0b + (size_t)buf * d0 * TS;
    cp_async_wait_all();
    __syncthreads();                                 // previous tile done, this tile's features landed
    if (tile + 1 < tile1)
      prefetch_tile_features<real>(tiles[3 * tile + 3], tiles[3 * tile + 4], T, perm, feat, nFeat, d0,
                                   A0b + (size_t)(buf ^ 1) * d0 * TS, TS);
    if (sp != curSp) {
      if (MODE == 0) {
        double *gp = partials + (size_t)blockIdx.x * nDD;
        if (gInSmem) {
          if (curSp >= 0) for (int e = threadIdx.x; e < net.nTot; e += blockDim.x) gp[(size_t)net.nTot * curSp + e] = Gs[e];
          __syncthreads();
          for (int e = threadIdx.x; e < net.nTot; e += blockDim.x) Gs[e] = 0.0;
          G = Gs;
        } else {
          G = gp + (size_t)net.nTot * sp;           // accumulate in the (zeroed) global partial row
        }
      }
      load_weights_padded<real>(net, m, wb + (size_t)net.nTot * sp, wsm);
      curSp = sp;
      __syncthreads();
    }
    for (int l = 1; l < L; l++) {
      const int actId = (l == L - 1) ? FNETGPU_ACT_LINEAR : net.act;   // network.F90:391
      const real *lin = (l == 1) ? A0 : A + (size_t)net.aoff[l - 1] * TS;
      if (MODE == 2)
        layer_forward<real, false>(net.dims[l - 1], net.dims[l], actId, wsm + m.wOff[l - 1], m.wS[l - 1], wsm + m.bOff[l],
                                   lin, A + (size_t)net.aoff[l] * TS, nullptr, T, TS);
      else
        layer_forward<real, true>(net.dims[l - 1], net.dims[l], actId, wsm + m.wOff[l - 1], m.wS[l - 1], wsm + m.bOff[l],
                                  lin, A + (size_t)net.aoff[l] * TS,
                                  D + (size_t)(net.aoff[l] - d0) * TS, T, TS);
      __syncthreads();
    }
    if (MODE == 2) {
      for (int e = threadIdx.x; e < count * net.nOut; e += blockDim.x) {
        const int tt = e / net.nOut, k = e % net.nOut;
        raw[(size_t)net.nOut * perm[start + tt] + k] = A[(size_t)(net.aoff[L - 1] + k) * TS + tt];
      }
      continue;
    }
    const int nSweeps = (MODE == 0) ? 1 : net.nOut;
    real *dL = D + (size_t)(net.aoff[L - 1] - d0) * TS;
    for (int sweep = 0; sweep < nSweeps; sweep++) {
      // output layer (linear): delta_L = lossgrad (x) f' (network.F90:276); 0 for padding atoms
      for (int tt = threadIdx.x; tt < T; tt += blockDim.x) {
        if (MODE == 0) {
          const int atom = (tt < count) ? perm[start + tt] : -1;
          const int s = atom >= 0 ? structOf[atom] : 0;
          const double scale = atom >= 0 ? dsw[s] * aw[atom] / (double)(offsets[s + 1] - offsets[s]) : 0.0;   // bpnn.F90:446,698
          for (int k = 0; k < net.nOut; k++) {
            double g = 0.0;
            if (atom >= 0) {
              if (k < nG) g = gS[(size_t)nG * s + k];
              else g = loss_grad_fn(lossId, (double)A[(size_t)(net.aoff[L - 1] + k) * TS + tt], at[(size_t)nA * atom + (k - nG)]);
            }
            dL[(size_t)k * TS + tt] = (real)(g * scale);
          }
        } else {
          for (int k = 0; k < net.nOut; k++) dL[(size_t)k * TS + tt] = (k == sweep && tt < count) ? (real)1 : (real)0;
        }
      }
      __syncthreads();
      // hidden layers: delta_{l-1} = (W_{l-1} delta_l) * f'(z_{l-1}).  MODE 1 keeps f' for the next
      // sweep, so its deltas go to the A rows of the layer (hidden activations are not needed there).
      for (int l = L - 1; l >= (MODE == 0 ? 2 : 1); l--) {
        const int din = net.dims[l - 1], dout = net.dims[l];
        const real *dn = (MODE == 1 && l < L - 1) ? A + (size_t)net.aoff[l] * TS : D + (size_t)(net.aoff[l] - d0) * TS;
        if (l - 1 == 0) {
          layer_backward<real>(din, dout, wsm + m.wOff[0], m.wS[0], dn, T, TS, [&](int i, int rg, const real (&v)[4]) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
              const int tt = 4 * rg + a;
              if (tt < count) dEdG[((size_t)net.nOut * perm[start + tt] + sweep) * d0 + i] = v[a];
            }
          });
        } else {
          real *fp = D + (size_t)(net.aoff[l - 1] - d0) * TS;
          real *dst = (MODE == 0) ? fp : A + (size_t)net.aoff[l - 1] * TS;
          layer_backward<real>(din, dout, wsm + m.wOff[l - 1], m.wS[l - 1], dn, T, TS, [&](int i, int rg, const real (&v)[4]) {
            real f4[4], r4[4];
            ld4(fp + (size_t)i * TS + 4 * rg, f4);
#pragma unroll
            for (int a = 0; a < 4; a++) r4[a] = v[a] * f4[a];
            st4(dst + (size_t)i * TS + 4 * rg, r4);
          });
        }
        __syncthreads();
      }
    }
    if (MODE == 0) {
      for (int l = 0; l + 1 < L; l++)
        layer_wgrad<real>(net.dims[l], net.dims[l + 1], (l == 0) ? A0 : A + (size_t)net.aoff[l] * TS,
                          D + (size_t)(net.aoff[l + 1] - d0) * TS, T, TS, G + net.woff[l], G + net.boff[l + 1]);
    }
  }
  if (MODE == 0 && gInSmem && curSp >= 0) {
    __syncthreads();
    double *gp = partials + (size_t)blockIdx.x * nDD;
    for (int e = threadIdx.x; e < net.nTot; e += blockDim.x) gp[(size_t)net.nTot * curSp + e] = Gs[e];
  }
}

// shared-memory bytes of k_bpnn for tile size T
template <typename real>
__host__ inline size_t bpnn_smem_bytes(const NetTables &net, int T, int mode, bool gInSmem) {
  const MlpLayout m = mlp_layout<real>(net);
  const int TS = T + VecPad<real>::value;
  size_t b = ((size_t)m.total + (size_t)(net.rowsA + net.dims[0]) * TS) * sizeof(real);   // + second input buffer
  if (mode != 2) b += (size_t)(net.rowsA - net.dims[0]) * TS * sizeof(real);
  b = (b + 15) & ~(size_t)15;
  if (mode == 0 && gInSmem) b += (size_t)net.nTot * sizeof(double);
  return b;
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

#ifndef FNET_KERNEL_TU   // non-template kernels: defined once, in fnetgpu.cu (kernels_*.cu set FNET_KERNEL_TU)
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

// dd[p] = sum_cta partials[cta][p] in a fixed order (deterministic): a CTA of FNET_GRED_J x 32 threads takes 32 outputs,
// thread (j, pl) sums the CTAs c = j, j + J, ... (independent, coalesced loads), the J sub-sums are added in order of j.
// (One thread per output walked all CTAs alone: 296 dependent additions behind 4 loads in flight -- 35 us on C2.)
#define FNET_GRED_J 16
// With lossPart != nullptr the LAST CTA of the grid instead reduces the per-structure loss terms of the fused gradient
// kernels (what k_loss_final does as a launch of its own): lossOut[0] = numerator, lossOut[1] = denominator.
__global__ void __launch_bounds__(FNET_GRED_J * 32) k_grad_reduce(int nCta, int nDD, const double *__restrict__ partials, double *__restrict__ dd,
                                                                   int nStruct = 0, const double *__restrict__ lossPart = nullptr,
                                                                   double *__restrict__ lossOut = nullptr) {
  __shared__ double sh[FNET_GRED_J][32];
  if (lossPart && blockIdx.x == gridDim.x - 1) {
    __shared__ double sm[2][FNET_GRED_J * 32];
    double a = 0.0, b = 0.0;
    for (int s = threadIdx.x; s < nStruct; s += blockDim.x) { a += lossPart[2 * s]; b += lossPart[2 * s + 1]; }
    sm[0][threadIdx.x] = a; sm[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) { sm[0][threadIdx.x] += sm[0][threadIdx.x + o]; sm[1][threadIdx.x] += sm[1][threadIdx.x + o]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) { lossOut[0] = sm[0][0]; lossOut[1] = sm[1][0]; }
    return;
  }
  const int pl = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int p = blockIdx.x * 32 + pl;
  double s = 0.0;
  if (p < nDD) {
#pragma unroll 4
    for (int c = j; c < nCta; c += FNET_GRED_J) s += partials[(size_t)c * nDD + p];
  }
  sh[j][pl] = s;
  __syncthreads();
  if (j == 0 && p < nDD) {
    double t = sh[0][pl];
#pragma unroll
    for (int q = 1; q < FNET_GRED_J; q++) t += sh[q][pl];
    dd[p] = t;
  }
}
#endif
