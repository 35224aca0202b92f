// mlp_mma.cuh -- FP64 subnetwork kernels on the FP64 tensor-core path (DMMA m8n8k4).
//
// Same maths and the same reference sites as mlp.cuh (TNetwork_fprop / bprop, lib_nn/network.F90:
// 146-180, 248-296; TBpnn_sysTrain / updateGradients, lib_nn/bpnn.F90:394-481, 610-704); this file
// changes how the three small contractions of every layer are executed in precision 64:
//     forward    Z[t][o]  = sum_i A[t][i] W[i][o]          (M = atoms, N = outputs, K = inputs)
//     backward   D[t][i]  = sum_o Dn[t][o] W[i][o]         (M = atoms, N = inputs,  K = outputs)
//     gradient   dW[i][o] = sum_t A[t][i] Dn[t][o]         (M = inputs, N = outputs, K = atoms)
// Layer widths of 1..100 leave the register-tiled DFMA kernel of mlp.cuh latency- and
// barrier-bound (ncu: FP64 pipe 26 % busy, 7 CTA barriers per 64 atoms, idle warps in every
// phase).  mma.sync.m8n8k4.f64 issues 256 FMAs per warp instruction with one 8-byte operand per
// lane and matrix, so
//   * a WARP owns a tile of 16 atoms from the features to the input-layer deltas: no CTA
//     barrier inside the forward / backward sweeps, every lane busy, 2 + NT shared-memory loads
//     per 2*NT DMMAs;
//   * the weight gradients are accumulated in REGISTERS for the whole kernel: the 8x8 output
//     tiles of all dW_l are dealt out to the CTA's warps once (<= FNET_MMA_MAXSLOTS each); after
//     each round (one tile per warp) the warps sweep all tiles of the round, K = atoms.  Two CTA
//     barriers per round, no shared-memory gradient buffer, no atomics: the result is
//     bit-reproducible (fixed order), as before.
// Activations / deltas sit in shared memory as [row][atom] with a row stride of 20 doubles
// (= 4 mod 16) and the weights as [out][in] with a row stride = 4 mod 16, which makes every
// fragment load (8 rows x 4 columns or 4 rows x 8 columns of doubles) bank-conflict free.
// FP64 tensor cores run at the DFMA rate on B200 (tools/peaks.cu: 18.5 vs 16.9 TFMA/s) -- the
// gain is issue slots and latency, not peak.  Precision 32, the input-gradient sweep of the
// forces (MODE 1) and networks that do not fit the limits below stay on mlp.cuh.
#pragma once
#include <cooperative_groups.h>
#include "mlp.cuh"

// build-time knobs of the K loops (A/B, tools/gpu_ab2.sh, same box: unroll 4 instead of 2: C2 0.400 -> 0.398 ms, C3 7.13 -> 7.17;
// plain asm instead of asm volatile: 0.402 / 7.19 ms -- neither moves the kernel, the defaults stay)
#ifndef FNET_MMA_ASM
#define FNET_MMA_ASM asm volatile
#endif
#ifndef FNET_MMA_KUNROLL
#define FNET_MMA_KUNROLL 2
#endif
#define FNET_MMA_TA 8              // atoms per warp (one m8 tile)
#define FNET_MMA_TW 16             // atoms per shared-memory tile: two warps share a tile (columns 0-7 / 8-15)
#define FNET_MMA_TS 20             // row stride of the [row][atom] tiles, doubles
#define FNET_MMA_WARPS 8
#define FNET_MMA_TILES (FNET_MMA_WARPS / 2)
#define FNET_MMA_MAXSLOTS 9        // weight-gradient output tiles a warp can own
#define FNET_MMA_CSLOTS 32         // cluster-fused sums: structures per super-round (slots of the exchange buffer)

__host__ __device__ inline int fnet_ru4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int fnet_ru8(int x) { return (x + 7) & ~7; }

struct MmaLayout {
  int wS[FNET_MAX_LAYERS];     // row stride of W_l, stored [out][in] like the serialised ww(i,o): >= roundup4(din), = 4 mod 16
  int wOff[FNET_MAX_LAYERS];   // W_l: roundup8(dout) rows (zero padded)
  int bOff[FNET_MAX_LAYERS];   // bias of layer l >= 1: roundup8(dims[l]) entries (zero padded)
  int wTotal;                  // doubles (multiple of 2, + slack for the column over-read of the last row)
  int aOff[FNET_MAX_LAYERS];   // first row of a_l in a warp tile; block height roundup8(dims[l])
  int dOff[FNET_MAX_LAYERS];   // first row of f'(z_l), then delta_l (l >= 1); block height roundup8(dims[l])
  int aOffF[FNET_MAX_LAYERS];  // forward-only tile: layers ping-pong between two row blocks
  int sOff;                    // 8 scratch rows of the training tile (fused per-structure sums)
  int rowsA, rows;             // rows of the forward-only tile / of the training tile
  int nGradTiles;              // 8x8 output tiles of all weight gradients
  int nBias;                   // sum of dims[1..L-1]
};

__host__ __device__ inline MmaLayout mma_layout(const NetTables &net) {
  MmaLayout m;
  int off = 0;
  m.nGradTiles = 0; m.nBias = 0;
  for (int l = 0; l + 1 < net.L; l++) {
    int s = fnet_ru4(net.dims[l]);
    while ((s & 15) != 4) s += 4;
    m.wS[l] = s;
    m.wOff[l] = off;
    off += fnet_ru8(net.dims[l + 1]) * s;
    m.nGradTiles += (fnet_ru8(net.dims[l]) >> 3) * (fnet_ru8(net.dims[l + 1]) >> 3);
    m.nBias += net.dims[l + 1];
  }
  m.bOff[0] = off;
  for (int l = 1; l < net.L; l++) { m.bOff[l] = off; off += fnet_ru8(net.dims[l]); }
  m.wTotal = off + 8;
  int r = 0;
  for (int l = 0; l < net.L; l++) { m.aOff[l] = r; r += fnet_ru8(net.dims[l]); }
  {
    int x = 0, y = 0;
    for (int l = 0; l < net.L; l++) { if (l & 1) y = max(y, fnet_ru8(net.dims[l])); else x = max(x, fnet_ru8(net.dims[l])); }
    for (int l = 0; l < net.L; l++) m.aOffF[l] = (l & 1) ? x : 0;
    m.rowsA = x + y;
  }
  m.dOff[0] = r;
  for (int l = 1; l < net.L; l++) { m.dOff[l] = r; r += fnet_ru8(net.dims[l]); }
  m.sOff = r; r += 8;
  m.rows = r;
  return m;
}

// nGc >= 0: plus the double-buffered exchange area of the cluster-fused per-structure sums (nGc global targets)
__host__ inline size_t bpnn_mma_smem_bytes(const NetTables &net, int mode, int nGc = -1) {
  const MmaLayout m = mma_layout(net);
  const size_t ex = nGc >= 0 ? (size_t)2 * FNET_MMA_CSLOTS * (nGc + 2) : 0;
  return ((size_t)m.wTotal + FNET_EXP_TAB_N + (size_t)FNET_MMA_TILES * (mode == 2 ? m.rowsA : m.rows) * FNET_MMA_TS + ex) * sizeof(double);
}
// limits of this path (else mlp.cuh): bias accumulators are one per thread, gradient tiles
// <= MAXSLOTS per warp, everything in 220 KB of shared memory
__host__ inline bool bpnn_mma_fits(const NetTables &net) {
  const MmaLayout m = mma_layout(net);
  return m.nBias <= FNET_MMA_WARPS * 32 && m.nGradTiles <= FNET_MMA_WARPS * FNET_MMA_MAXSLOTS &&
         bpnn_mma_smem_bytes(net, 0) <= 220 * 1024;
}

// D += A(8x4, row) * B(4x8, col).  Lane (g = lane / 4, c = lane % 4) holds A[g][c], B[c][g] and
// D[g][2c], D[g][2c + 1].
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  FNET_MMA_ASM("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// weights of one species: serialised ww(i,o) at i + din*o (network.F90:413-419) -> [o][wS] padded
__device__ __forceinline__ void mma_load_weights(const NetTables &net, const MmaLayout &m,
                                                 const double *__restrict__ wb, double *__restrict__ wsm) {
  // (wTotal is even and wsm 16-byte aligned: 16-byte stores; rows by warps and columns by lanes: no division -- a CTA of
  // the C2 launch lives for 34 rounds only, and this prologue was 10 % of its instructions)
  for (int e = threadIdx.x; 2 * e < m.wTotal; e += blockDim.x) ((double2 *)wsm)[e] = make_double2(0.0, 0.0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int l = 0; l + 1 < net.L; l++) {
    const int din = net.dims[l], dout = net.dims[l + 1];
    const double *W = wb + net.woff[l];
    double *dst = wsm + m.wOff[l];
    const int wS = m.wS[l];
    for (int o = warp; o < dout; o += nwarp)
      for (int i = lane; i < din; i += 32) dst[o * wS + i] = W[o * din + i];
  }
  for (int l = 1; l < net.L; l++)
    for (int e = threadIdx.x; e < net.dims[l]; e += blockDim.x) wsm[m.bOff[l] + e] = wb[net.boff[l] + e];
}

// transfer functions other than tanh (mma_forward applies that one in registers) as a separate element-wise pass
// over a warp tile (z -> f(z), optionally f'(z)): an even split of the dout*8 elements over the lanes.  Inlined into its two call sites (the layer loop of the forward sweep, with
// and without derivative): as an out-of-line function its tile and table pointers were GENERIC (LD.E / ST.E instead of
// LDS / STS) and the two-element arrays lived in local memory (ncu: 38 % of the kernel's instructions in this pass).
template <bool DERIV>
__device__ __forceinline__ void mma_activate(int actId, int dout, double *__restrict__ out, double *__restrict__ dact, int lane,
                                          const double *__restrict__ etab) {
  // two elements per lane and iteration: independent dependency chains (the FP64 exp / reciprocal
  // sequences are latency-bound); the warp's 8 atoms are columns 0..7 of `out`
  const int n = dout * FNET_MMA_TA;
#pragma unroll 1
  for (int base = lane; base < n; base += 64) {
    double x[2], v[2];
    int off[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int idx = min(base + 32 * q, n - 1);
      off[q] = (idx >> 3) * FNET_MMA_TS + (idx & 7);
      x[q] = out[off[q]];
    }
    if (actId == FNETGPU_ACT_SIGMOID) {
#pragma unroll
      for (int q = 0; q < 2; q++) v[q] = act_f<double>(FNETGPU_ACT_SIGMOID, x[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 2; q++) v[q] = act_f<double>(actId, x[q]);
    }
    double d[2];
    if (DERIV) {
      if (actId == FNETGPU_ACT_SIGMOID) {
#pragma unroll
        for (int q = 0; q < 2; q++) d[q] = act_d<double>(FNETGPU_ACT_SIGMOID, x[q], v[q]);
      } else {
#pragma unroll
        for (int q = 0; q < 2; q++) d[q] = act_d<double>(actId, x[q], v[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 2; q++)
      if (base + 32 * q < n) {
        out[off[q]] = v[q];
        if (DERIV) dact[off[q]] = d[q];
      }
  }
}

// K loop of one group of NC <= 4 output tiles: NC is a template parameter (selected by a warp-uniform
// switch) so that the loop body carries no per-tile guards -- guarded mma.sync costs predicated-off
// issue slots and WARPSYNC / NOP padding (a 1-wide output layer would issue 4x its DMMAs)
template <int NC>
__device__ __forceinline__ void mma_k_loop(int KT, const double *__restrict__ ip, const double *__restrict__ wp,
                                           size_t aStep, size_t bStep, size_t bTile, double (&acc)[4][2]) {
  constexpr int KU = FNET_MMA_KUNROLL;
#pragma unroll KU
  for (int kt = 0; kt < KT; kt++) {
    const double a0 = ip[kt * aStep];
#pragma unroll
    for (int nc = 0; nc < NC; nc++) dmma(acc[nc], a0, wp[nc * bTile + kt * bStep]);
  }
}
__device__ __forceinline__ void mma_k_dispatch(int ncnt, int KT, const double *__restrict__ ip, const double *__restrict__ wp,
                                               size_t aStep, size_t bStep, size_t bTile, double (&acc)[4][2]) {
  switch (ncnt) {
    case 1: mma_k_loop<1>(KT, ip, wp, aStep, bStep, bTile, acc); break;
    case 2: mma_k_loop<2>(KT, ip, wp, aStep, bStep, bTile, acc); break;
    case 3: mma_k_loop<3>(KT, ip, wp, aStep, bStep, bTile, acc); break;
    default: mma_k_loop<4>(KT, ip, wp, aStep, bStep, bTile, acc); break;
  }
}

// forward layer of one warp tile: out[o][t] = f(sum_i W[o][i] in[i][t] + b[o]); DERIV also
// stores f'(z) (later overwritten by the delta).  tanh -- the reference's default transfer function
// (initprogram.F90) -- is applied to the DMMA accumulators IN REGISTERS (lane (g, c) holds z of atom g for the
// outputs 8 nt + 2 c, + 1: two independent chains per tile) and a, f'(z) are stored once; the separate
// element-wise pass (store z, read z, index arithmetic, store a / f') was 36 % of the kernel's instructions
// (ncu source page).  Other transfer functions keep that pass (mma_activate).
template <bool DERIV>
__device__ __forceinline__ void mma_forward(int din, int dout, int actId, const double *__restrict__ W, int wS,
                                            const double *__restrict__ bias, const double *__restrict__ in,
                                            double *__restrict__ out, double *__restrict__ dact, int lane,
                                            const double *__restrict__ etab) {
  const int g = lane >> 2, c = lane & 3;
  const int KT = fnet_ru4(din) >> 2, NT = fnet_ru8(dout) >> 3;
  const bool inreg = actId == FNETGPU_ACT_TANH;
  for (int nt0 = 0; nt0 < NT; nt0 += 4) {
    double acc[4][2];
#pragma unroll
    for (int nc = 0; nc < 4; nc++) {
      const int nb = 8 * min(nt0 + nc, NT - 1) + 2 * c;
      acc[nc][0] = bias[nb]; acc[nc][1] = bias[nb + 1];
    }
    const double *ip = in + c * FNET_MMA_TS + g;
    const double *wp = W + (size_t)(8 * nt0 + g) * wS + c;
    mma_k_dispatch(min(4, NT - nt0), KT, ip, wp, 4 * FNET_MMA_TS, 4, (size_t)8 * wS, acc);
    if (inreg) {
#pragma unroll
      for (int nc = 0; nc < 4; nc++)
        if (nt0 + nc < NT) {
          const int o = 8 * (nt0 + nc) + 2 * c;
          const double v0 = fnet_tanh_em1(acc[nc][0], etab), v1 = fnet_tanh_em1(acc[nc][1], etab);
          if (o < dout) {
            out[o * FNET_MMA_TS + g] = v0;
            if (DERIV) dact[o * FNET_MMA_TS + g] = fma(-v0, v0, 1.0);        // transfer.F90: 1 - tanh^2
          }
          if (o + 1 < dout) {
            out[(o + 1) * FNET_MMA_TS + g] = v1;
            if (DERIV) dact[(o + 1) * FNET_MMA_TS + g] = fma(-v1, v1, 1.0);
          }
        }
    } else {
#pragma unroll
      for (int nc = 0; nc < 4; nc++)
        if (nt0 + nc < NT) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int o = 8 * (nt0 + nc) + 2 * c + e;
            if (o < dout) out[o * FNET_MMA_TS + g] = acc[nc][e];
          }
        }
    }
  }
  if (!inreg && (actId != FNETGPU_ACT_LINEAR || DERIV)) {
    __syncwarp();
    mma_activate<DERIV>(actId, dout, out, dact, lane, etab);
  }
}

// backward layer of one warp tile: dst[i][t] = (sum_o W[o][i] dn[o][t]) * f'(z)[i][t]
// (network.F90:282-288); fp holds f'(z) (training: fp == dst, the delta replaces it; input-gradient
// sweeps keep f' for the next output and put the delta elsewhere); fp == nullptr: no factor (input layer)
__device__ __forceinline__ void mma_backward(int din, int dout, const double *__restrict__ W, int wS,
                                             const double *__restrict__ dn, const double *fp, double *dst, int lane) {
  const int g = lane >> 2, c = lane & 3;
  const int KT = fnet_ru4(dout) >> 2, NT = fnet_ru8(din) >> 3;
  for (int nt0 = 0; nt0 < NT; nt0 += 4) {
    double acc[4][2];
#pragma unroll
    for (int nc = 0; nc < 4; nc++) { acc[nc][0] = 0.0; acc[nc][1] = 0.0; }
    const double *ip = dn + c * FNET_MMA_TS + g;
    const double *wp = W + (size_t)c * wS + 8 * nt0 + g;
    mma_k_dispatch(min(4, NT - nt0), KT, ip, wp, 4 * FNET_MMA_TS, (size_t)4 * wS, 8, acc);
#pragma unroll
    for (int nc = 0; nc < 4; nc++)
      if (nt0 + nc < NT) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = 8 * (nt0 + nc) + 2 * c + e;
          if (i < din) dst[i * FNET_MMA_TS + g] = fp ? acc[nc][e] * fp[i * FNET_MMA_TS + g] : acc[nc][e];
        }
      }
  }
}

// gradient sweep over one tile for the first NM of the warp's output tiles (NM by a warp-uniform
// switch: no per-slot guards around mma.sync)
template <int NM, int NSLOT>
__device__ __forceinline__ void mma_wgrad_tile(const double *__restrict__ tb, const int (&aRow)[NSLOT], const int (&dRow)[NSLOT],
                                               const bool (&newA)[NSLOT], double (&acc)[NSLOT][2]) {
#pragma unroll
  for (int kt = 0; kt < FNET_MMA_TW / 4; kt++) {
    double a = 0.0;
#pragma unroll
    for (int s = 0; s < NM; s++) {
      if (newA[s]) a = tb[aRow[s] * FNET_MMA_TS + 4 * kt];
      dmma(acc[s], a, tb[dRow[s] * FNET_MMA_TS + 4 * kt]);
    }
  }
}
template <int NM, int NSLOT>
struct MmaWgradDispatch {
  static __device__ __forceinline__ void run(int nMine, const double *__restrict__ tb, const int (&aRow)[NSLOT],
                                             const int (&dRow)[NSLOT], const bool (&newA)[NSLOT], double (&acc)[NSLOT][2]) {
    if (nMine == NM) mma_wgrad_tile<NM, NSLOT>(tb, aRow, dRow, newA, acc);
    else MmaWgradDispatch<NM - 1, NSLOT>::run(nMine, tb, aRow, dRow, newA, acc);
  }
};
template <int NSLOT>
struct MmaWgradDispatch<0, NSLOT> {
  static __device__ __forceinline__ void run(int, const double *, const int (&)[NSLOT], const int (&)[NSLOT],
                                             const bool (&)[NSLOT], double (&)[NSLOT][2]) {}
};

// ------------------------------------------------------------------------------------------
// MODE 0: training gradient -> partials[cta][nSpecies*nTot]; MODE 2: forward only -> raw[atom][k];
// MODE 1: input gradients -> raw[atom][k][F] (the dE/dG of the analytic forces).
// `tiles` holds the ROUNDS: (start, count <= 64, species) triples of the species-sorted atom
// order; a CTA walks a contiguous range of rounds, warp w of the CTA takes atoms 8 w .. 8 w + 7.
// smem: weights | tile 0 | .. | tile 3   (tile: rows a_0 .. a_{L-1} | delta_1 .. delta_{L-1} | scratch,
// 16 atom columns: warps 2t and 2t+1 own columns 0-7 / 8-15 of tile t)
// ------------------------------------------------------------------------------------------
// FUSED (MODE 0, global targets only, every structure inside ONE round -- single-species data with
// <= 64 atoms per structure): the per-structure sums E_s, the loss gradients and the loss terms
// (TBpnn_sysTrain, bpnn.F90:677-684; loss.F90:370-721) are formed inside the round, between the
// forward and the backward sweep, so the separate forward kernel and k_struct_loss disappear and
// every atom is propagated forward exactly once per iteration.
// FUSED == 2 (cluster-fused): the same for multi-species data and structures of up to 64 x (cluster size) atoms.  The
// kernel is launched as thread-block CLUSTERS; `tiles` holds SUPER-ROUNDS of CS entries (start, count, species, first
// structure), one per CTA of the cluster: together the CS species-homogeneous rounds of a super-round hold all atoms of
// a group of <= FNET_MMA_CSLOTS whole structures (`perm` is ordered group / species / structure / atom, `segBE` the
// round-local range of every atom's structure).  After its forward sweep each CTA leaves the partial sums of its
// structures in shared memory, the cluster synchronises (barrier.cluster) and every CTA reads the other CTAs' partial
// sums through distributed shared memory in rank order: E_s, loss gradient and loss terms without a second forward pass.
// The kernel lives in its own translation unit (kernels_mma.cu defines FNET_DEFINE_MMA_KERNEL; see acsf_lean.cuh for the
// reason); the other units see a variable template of the same name that holds the kernel's address.
typedef void (*MmaKernelT)(int, const int *, const int *, const double *, int, const double *, NetTables, const int *,
                           const int *, const double *, const double *, const double *, const double *, int, int, int,
                           double *, double *, const double *, double *, double *, const int *);
MmaKernelT fnet_mma_kernel(int MODE, int NSLOT, int FCH, int FUSED);   // kernels_mma.cu (nullptr: not built)
#ifndef FNET_DEFINE_MMA_KERNEL
template <int MODE, int NSLOT, int FCH, int FUSED = 0>
static const MmaKernelT k_bpnn_mma = fnet_mma_kernel(MODE, NSLOT, FCH, FUSED);
#else
template <int MODE, int NSLOT, int FCH, int FUSED = 0>
__global__ void __launch_bounds__(FNET_MMA_WARPS * 32, (NSLOT <= 4 ? 2 : 1))
k_bpnn_mma(int nTiles, const int *__restrict__ tiles, const int *__restrict__ perm, const double *__restrict__ feat,
           int nFeat, const double *__restrict__ wb, NetTables net, const int *__restrict__ structOf,
           const int *__restrict__ offsets, const double *__restrict__ gS, const double *__restrict__ at,
           const double *__restrict__ aw, const double *__restrict__ dsw, int nG, int nA, int lossId,
           double *__restrict__ partials, double *__restrict__ raw, const double *__restrict__ gt = nullptr,
           double *__restrict__ Es = nullptr, double *__restrict__ lossPart = nullptr, const int *__restrict__ segBE = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TS = FNET_MMA_TS, TA = FNET_MMA_TA, TW = FNET_MMA_TW, NW = FNET_MMA_WARPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int L = net.L, d0 = net.dims[0];
  // the layer offsets are indexed with run-time layer numbers: as a local array they live in LOCAL memory, whose L1 lines
  // the streamed feature rows keep evicting (ncu: 5 % of the kernel's stall samples on the address arithmetic behind those
  // LDL at the head of every layer) -- one copy per CTA in shared memory instead
  __shared__ MmaLayout mShared;
  if (threadIdx.x == 0) mShared = mma_layout(net);
  __syncthreads();
  const MmaLayout &m = mShared;
  const int rows = (MODE == 2) ? m.rowsA : m.rows;
  double *wsm = (double *)smem_raw;
  double *etab = wsm + m.wTotal;                  // 2^(j/64) table of the transfer functions (fmath.cuh)
  double *tiles0 = etab + FNET_EXP_TAB_N;
  for (int e = threadIdx.x; e < FNET_EXP_TAB_N; e += blockDim.x) etab[e] = fnet_exp_tab_d[e];
  double *T = tiles0 + (size_t)(warp >> 1) * rows * TS + TA * (warp & 1);   // this warp's 8 columns of its tile
  for (int e = threadIdx.x; 2 * e < FNET_MMA_TILES * rows * TS; e += blockDim.x)               // padding rows stay zero from here on
    ((double2 *)tiles0)[e] = make_double2(0.0, 0.0);                                           // (TS is even, tiles0 16-byte aligned)
  __syncthreads();
  // cluster-fused sums: CS CTAs walk the super-rounds of their cluster in lock step
  int CS = 1, crank = 0;
  double *exch = tiles0 + (size_t)FNET_MMA_TILES * rows * TS;   // [2][FNET_MMA_CSLOTS][nG + 2]: E partials | sum of aw | atoms present
  if (FUSED == 2) {
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    CS = (int)cl.num_blocks(); crank = (int)cl.block_rank();
  }
  const int ES = FUSED == 2 ? 4 : 3;              // ints per entry

  // ---- weight-gradient tiles owned by this warp: q = warp*per + s <-> (layer, i-tile, o-tile) ----
  int aRow[NSLOT], dRow[NSLOT];
  bool newA[NSLOT];
  double acc[NSLOT][2];
  double bacc = 0.0;
  int nMine = 0, bRow = -1, bIdx = -1, q0 = 0;
  // gradient tile q -> (layer, i-tile, o-tile); recomputed in flush() rather than kept in registers
  auto tile_of = [&](int q, int &l, int &mt, int &nt) {
    l = 0;
    while (true) {
      const int cnt = (fnet_ru8(net.dims[l]) >> 3) * (fnet_ru8(net.dims[l + 1]) >> 3);
      if (q < cnt) break;
      q -= cnt; l++;
    }
    const int ntl = fnet_ru8(net.dims[l + 1]) >> 3;
    mt = q / ntl; nt = q % ntl;
  };
  if (MODE == 0) {
    const int per = (m.nGradTiles + NW - 1) / NW;
    q0 = warp * per;
    const int q1 = min(m.nGradTiles, q0 + per);
    nMine = max(q1 - q0, 0);
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
      acc[s][0] = 0.0; acc[s][1] = 0.0;
      aRow[s] = 0; dRow[s] = 0; newA[s] = false;
      if (s < nMine) {
        int l, mt, nt;
        tile_of(q0 + s, l, mt, nt);
        aRow[s] = m.aOff[l] + 8 * mt;
        dRow[s] = m.dOff[l + 1] + 8 * nt;
      }
    }
#pragma unroll
    for (int s = 0; s < NSLOT; s++) newA[s] = (s == 0) || (aRow[s] != aRow[s > 0 ? s - 1 : 0]);
    // bias gradients: thread tid <-> (layer l >= 1, output o)
    int t = threadIdx.x;
    if (t < m.nBias) {
      int l = 1;
      while (t >= net.dims[l]) { t -= net.dims[l]; l++; }
      bRow = m.dOff[l] + t;
      bIdx = net.boff[l] + t;
    }
  }
  auto flush = [&](int sp) {     // registers -> this CTA's partial row of species sp (visited once per CTA)
    double *gp = partials + ((size_t)blockIdx.x * net.nSpecies + sp) * net.nTot;
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
      if (s < nMine) {
        int l, mt, nt;
        tile_of(q0 + s, l, mt, nt);
        const int din = net.dims[l], dout = net.dims[l + 1];
        const int i = 8 * mt + g;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int o = 8 * nt + 2 * c + e;
          if (i < din && o < dout) gp[net.woff[l] + i + din * o] = acc[s][e];
        }
      }
      acc[s][0] = 0.0; acc[s][1] = 0.0;
    }
    if (bIdx >= 0) gp[bIdx] = bacc;
    bacc = 0.0;
  };

  // A CTA walks a contiguous range of ROUNDS: (start, count <= 64, species) entries of the
  // species-sorted atom order; warp w takes atoms [start + 16 w, start + 16 w + 16) of the round.
  // Software pipeline over the rounds, so that no tile waits for a chain of dependent HBM loads:
  //   entries two rounds ahead, atom indices two rounds ahead, feature rows of round r+2 pulled
  //   into L2, feature rows of round r+1 loaded into registers (pf) while round r's weight
  //   gradients are accumulated (MODE 0) / its hidden layers are evaluated (MODE 2).
  const int nWalk = gridDim.x / CS, walker = blockIdx.x / CS;   // CTAs (clusters) that share the rounds (super-rounds)
  const int round0 = (int)(((long long)nTiles * walker) / nWalk);
  const int round1 = (int)(((long long)nTiles * (walker + 1)) / nWalk);
  int curSp = -1;
  int e0[4] = {0, 0, 0, 0}, e1[4] = {0, 0, 0, 0};     // entries of round r and r + 1
  int atom0 = -1, atom1 = -1;                         // lane < 16: atom of this warp's tile in round r / r + 1
  double pf[FCH][TA];                                 // features of round r (lane <-> feature 32 ch + lane)
  auto load_entry = [&](int r, int (&e)[4]) {
    e[0] = e[1] = e[2] = e[3] = 0;
    if (r < round1) {
      const int *t = tiles + (size_t)ES * ((size_t)r * CS + crank);
      e[0] = t[0]; e[1] = t[1]; e[2] = t[2];
      if (FUSED == 2) e[3] = t[3];
    }
  };
  // the entry two rounds ahead is loaded lane-wise (lane & 3 = field) and broadcast by shuffles where it is first
  // needed, behind the sweeps: as warp-uniform loads the compiler moved the values to uniform registers at the load
  // (R2UR) and every warp waited for the global load there (ncu: 2.6 % of the stall samples)
  auto load_entry_lanes = [&](int r) -> int {
    const int f = lane & 3;
    return (r < round1 && f < ES) ? tiles[(size_t)ES * ((size_t)r * CS + crank) + f] : 0;
  };
  auto load_atom = [&](const int (&e)[4]) -> int {
    const int cnt = min(max(e[1] - TA * warp, 0), TA);
    return (lane < cnt) ? perm[e[0] + TA * warp + lane] : -1;
  };
  auto load_features = [&](int atomv) {
#pragma unroll
    for (int ch = 0; ch < FCH; ch++) {
      const int f = 32 * ch + lane;
#pragma unroll
      for (int u = 0; u < TA; u++) {
        const int atom = __shfl_sync(0xffffffffu, atomv, u);
        pf[ch][u] = (atom >= 0 && f < d0) ? feat[(size_t)nFeat * atom + f] : 0.0;
      }
    }
  };
  load_entry(round0, e0);
  load_entry(round0 + 1, e1);
  atom0 = load_atom(e0);
  atom1 = load_atom(e1);
  load_features(atom0);
  // structure of this lane's atom, one round ahead too: structOf -> offsets / weights / targets is a chain of dependent
  // global loads at the head of every round (ncu: 6 % of the kernel's stall samples on 2 % of its instructions)
  int strc0 = (MODE == 0 && atom0 >= 0) ? structOf[atom0] : 0;
  for (int r = round0; r < round1; r++) {
    const int sp = e0[2];
    const int nTl = (e0[1] + TW - 1) / TW;          // tiles with atoms in this round
    const int nIn = 2 * nTl;                        // warps at work: BOTH halves of every such tile (a half without atoms
                                                    // is swept with zero features so that no stale delta reaches the gradient sweep)
    const int count = min(max(e0[1] - TA * warp, 0), TA);
    const int myAtom = atom0;
    const int strc1 = (MODE == 0 && atom1 >= 0) ? structOf[atom1] : 0;
    const int ev2 = load_entry_lanes(r + 2);
    if (FUSED == 2) {      // this round's half of the exchange area: its last readers passed the previous cluster barrier
      double *ex = exch + (size_t)(r & 1) * FNET_MMA_CSLOTS * (nG + 2);
      for (int e = threadIdx.x; e < FNET_MMA_CSLOTS * (nG + 2); e += blockDim.x) ex[e] = 0.0;
    }
    if (sp != curSp && (FUSED != 2 || e0[1] > 0)) {
      __syncthreads();
      if (MODE == 0 && curSp >= 0) flush(curSp);
      mma_load_weights(net, m, wb + (size_t)net.nTot * sp, wsm);
      curSp = sp;
      __syncthreads();
    }
    // per-atom loss-gradient scale (MODE 0): issued now, consumed after the forward sweep
    double lgScale = 0.0, lgG0 = 0.0, lgG1 = 0.0, lgAw = 0.0, lgW = 0.0;
    int lgStruct = 0, lgB = 0, lgE = 0, lgSeg = 0;
    if (warp < nIn) {
      if (MODE == 0 && myAtom >= 0) {
        lgStruct = strc0;
        if (FUSED == 2) lgSeg = segBE[e0[0] + TA * warp + lane];
        lgB = offsets[lgStruct]; lgE = offsets[lgStruct + 1];
        lgAw = aw[myAtom]; lgW = dsw[lgStruct];       // (lgScale is formed after the forward sweep: no wait for these loads here)
        if (FUSED) {            // targets now, the sums after the forward sweep
          if (nG > 0) lgG0 = gt[(size_t)nG * lgStruct];
          if (nG > 1) lgG1 = gt[(size_t)nG * lgStruct + 1];
        } else {
          if (nG > 0) lgG0 = gS[(size_t)nG * lgStruct];
          if (nG > 1) lgG1 = gS[(size_t)nG * lgStruct + 1];
        }
      }
      // ---- features -> a_0[f][t] ----
#pragma unroll
      for (int ch = 0; ch < FCH; ch++) {
        const int f = 32 * ch + lane;
        if (f < d0) {
#pragma unroll
          for (int u = 0; u < TA; u++) T[f * TS + u] = pf[ch][u];
        }
      }
      for (int f0 = 32 * FCH; f0 < d0; f0 += 32) {     // rows beyond the register window: loaded in place
        const int f = f0 + lane;
        double v[TA];
#pragma unroll
        for (int u = 0; u < TA; u++) {
          const int atom = __shfl_sync(0xffffffffu, myAtom, u);
          v[u] = (atom >= 0 && f < d0) ? feat[(size_t)nFeat * atom + f] : 0.0;
        }
        if (f < d0) {
#pragma unroll
          for (int u = 0; u < TA; u++) T[f * TS + u] = v[u];
        }
      }
      if (MODE != 0) load_features(atom1);             // next round's rows: in flight during the whole sweep
      __syncwarp();
      // ---- forward ----
      for (int l = 1; l < L; l++) {
        const bool last = (l == L - 1);
        const int actId = last ? FNETGPU_ACT_LINEAR : net.act;   // network.F90:391
        const double *in = T + (MODE == 2 ? m.aOffF[l - 1] : m.aOff[l - 1]) * TS;
        double *out = T + (MODE == 2 ? m.aOffF[l] : m.aOff[l]) * TS;
        if (MODE == 2 || last)
          mma_forward<false>(net.dims[l - 1], net.dims[l], actId, wsm + m.wOff[l - 1], m.wS[l - 1], wsm + m.bOff[l], in, out, nullptr, lane, etab);
        else
          mma_forward<true>(net.dims[l - 1], net.dims[l], actId, wsm + m.wOff[l - 1], m.wS[l - 1], wsm + m.bOff[l], in, out,
                            T + m.dOff[l] * TS, lane, etab);
        __syncwarp();
      }
      if (MODE == 0) lgScale = myAtom >= 0 ? lgW * lgAw / (double)(lgE - lgB) : 0.0;   // bpnn.F90:446,698
      if (MODE == 2) {
        const double *o = T + m.aOffF[L - 1] * TS;
        if (lane < count)
          for (int k = 0; k < net.nOut; k++) raw[(size_t)net.nOut * myAtom + k] = o[k * TS + lane];
        __syncwarp();
      }
    }
    if (MODE == 0 && FUSED == 2) {
      // ---- cluster-fused: partial sums of this CTA's round -> exchange area; cluster barrier; totals in rank order ----
      if (warp < nIn && lane < TA) T[m.sOff * TS + lane] = lgAw;       // atomic weights of the tile (0 for padding)
      __syncthreads();
      const int NF = nG + 2;                                           // fields per slot: E_k partials | sum of aw | atoms present
      double *ex = exch + (size_t)(r & 1) * FNET_MMA_CSLOTS * NF;
      const bool mine = warp < nIn && lane < TA && myAtom >= 0;
      const int slot = lgStruct - e0[3];
      bool lead = false, uni = false;
      if (warp < nIn) {
        int b = lgSeg & 0xffff, e = lgSeg >> 16;                       // this structure's atoms in this round, round-local
        const int bU = __shfl_sync(0xffffffffu, b, 0), eU = __shfl_sync(0xffffffffu, e, 0);
        uni = count > 0 && __all_sync(0xffffffffu, !mine || (b == bU && e == eU));
        if (uni) { b = bU; e = eU; }
        lead = mine && (TA * warp + lane == b);
        for (int k = 0; k <= nG; k++) {                                // k == nG: the atomic weights (scratch row)
          const size_t rowk = (size_t)(k < nG ? m.aOff[L - 1] + k : m.sOff) * TS;
          double ek = 0.0;
          if (uni) {
            for (int i = b + lane; i < e; i += 32) ek += tiles0[(size_t)(i >> 4) * rows * TS + rowk + (i & 15)];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ek += __shfl_xor_sync(0xffffffffu, ek, o);
          } else if (mine) {
            for (int w2 = b >> 4; w2 <= (e - 1) >> 4; w2++) {
              const double *rp = tiles0 + (size_t)w2 * rows * TS + rowk;
              const int t1 = min(e - TW * w2, TW);
              for (int t = max(b - TW * w2, 0); t < t1; t++) ek += rp[t];
            }
          }
          if (lead) ex[slot * NF + k] = ek;
        }
        if (lead) ex[slot * NF + nG + 1] = 1.0;
      }
      cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
      cl.sync();
      if (warp < nIn) {
        double *dL = T + m.dOff[L - 1] * TS;
        double sw = 0.0, before = 0.0, E0 = 0.0, E1 = 0.0;
        const bool fast = uni && NF <= 4;                              // warp-uniform
        if (fast) {
          // one structure for the whole warp: lane = (field, rank) -- every remote value is ONE load in flight per lane,
          // summed over the ranks in rank order by shuffles
          const int slotU = __shfl_sync(0xffffffffu, slot, 0);
          const int f = lane >> 3, j = lane & 7;
          double v = 0.0;
          if (f < NF && j < CS) v = cl.map_shared_rank(ex, j)[slotU * NF + f];
          double tot = 0.0, bef = 0.0;
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const double x = __shfl_sync(0xffffffffu, v, (lane & ~7) + q);
            tot += x;
            if (q < crank) bef += x;
          }
          sw = __shfl_sync(0xffffffffu, tot, 8 * nG);
          before = __shfl_sync(0xffffffffu, bef, 8 * (nG + 1));
          E0 = __shfl_sync(0xffffffffu, tot, 0);
          E1 = __shfl_sync(0xffffffffu, tot, nG > 1 ? 8 : 0);
        } else if (mine) {
          for (int j = 0; j < CS; j++) {
            const double *rx = cl.map_shared_rank(ex, j) + slot * NF;
            sw += rx[nG];
            if (j < crank) before += rx[nG + 1];
          }
        }
        if (mine) {
          double ss = 0.0;
          const bool writer = lead && before == 0.0;                  // first round (in rank order) that holds atoms of the structure
          for (int k = 0; k < nG; k++) {
            double ek = (k == 0) ? E0 : E1;
            if (!fast) {
              ek = 0.0;
              for (int j = 0; j < CS; j++) ek += cl.map_shared_rank(ex, j)[slot * NF + k];
            }
            const double tv = (k == 0) ? lgG0 : (k == 1) ? lgG1 : gt[(size_t)nG * lgStruct + k];
            dL[k * TS + lane] = loss_grad_fn(lossId, ek, tv) * lgScale;
            if (writer) {
              Es[(size_t)nG * lgStruct + k] = ek;
              switch (lossId) {
                case FNETGPU_LOSS_MAE: ss += fabs(tv - ek); break;
                case FNETGPU_LOSS_MAPE: ss += fabs((tv - ek) / tv); break;
                default: ss += (tv - ek) * (tv - ek);
              }
            }
          }
          if (writer) {
            double lg;
            switch (lossId) {
              case FNETGPU_LOSS_RMS: lg = sqrt(ss / nG); break;
              case FNETGPU_LOSS_MAPE: lg = 100.0 * ss / nG; break;
              default: lg = ss / nG;
            }
            lossPart[2 * (size_t)lgStruct] = lgW * sw * lg;
            lossPart[2 * (size_t)lgStruct + 1] = lgW * sw;
          }
        } else if (lane < TA) {
          for (int k = 0; k < net.nOut; k++) dL[k * TS + lane] = 0.0;  // padding atoms
        }
      }
    }
    if (MODE == 0 && FUSED == 1) {
      // ---- per-structure sums across the warps of the round (fixed order), loss gradient, loss terms ----
      if (warp < nIn && lane < TA) T[m.sOff * TS + lane] = lgAw;       // atomic weights of the tile (0 for padding)
      __syncthreads();
      if (warp < nIn) {
        const bool mine = lane < TA && myAtom >= 0;
        // usual case: all atoms of this warp belong to one structure -> the 32 lanes share its sum
        const int bU = __shfl_sync(0xffffffffu, lgB, 0), eU = __shfl_sync(0xffffffffu, lgE, 0);
        const bool uni = count > 0 && __all_sync(0xffffffffu, !mine || (lgB == bU && lgE == eU));
        const int b = (uni ? bU : lgB) - e0[0], e = (uni ? eU : lgE) - e0[0];   // the structure's atoms, round-local
        const bool lead = mine && (TA * warp + lane == b);             // first atom of the structure: writes E_s and the loss terms
        double ss = 0.0;
        double *dL = T + m.dOff[L - 1] * TS;
        for (int k = 0; k < nG; k++) {
          const size_t rowk = (size_t)(m.aOff[L - 1] + k) * TS;
          double ek = 0.0;
          if (uni) {
            for (int i = b + lane; i < e; i += 32) ek += tiles0[(size_t)(i >> 4) * rows * TS + rowk + (i & 15)];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ek += __shfl_xor_sync(0xffffffffu, ek, o);
          } else if (mine) {
            for (int w2 = b >> 4; w2 <= (e - 1) >> 4; w2++) {          // tile by tile, in atom order
              const double *rp = tiles0 + (size_t)w2 * rows * TS + rowk;
              const int t1 = min(e - TW * w2, TW);
              for (int t = max(b - TW * w2, 0); t < t1; t++) ek += rp[t];
            }
          }
          if (mine) {
            const double tv = (k == 0) ? lgG0 : (k == 1) ? lgG1 : gt[(size_t)nG * lgStruct + k];
            dL[k * TS + lane] = loss_grad_fn(lossId, ek, tv) * lgScale;
            if (lead) {
              Es[(size_t)nG * lgStruct + k] = ek;
              switch (lossId) {
                case FNETGPU_LOSS_MAE: ss += fabs(tv - ek); break;
                case FNETGPU_LOSS_MAPE: ss += fabs((tv - ek) / tv); break;
                default: ss += (tv - ek) * (tv - ek);
              }
            }
          }
        }
        double sw = 0.0;                                             // sum of the structure's atomic weights
        if (uni) {
          for (int i = b + lane; i < e; i += 32) sw += tiles0[(size_t)(i >> 4) * rows * TS + (size_t)m.sOff * TS + (i & 15)];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
        } else if (lead) {
          for (int w2 = b >> 4; w2 <= (e - 1) >> 4; w2++) {
            const double *rp = tiles0 + (size_t)w2 * rows * TS + m.sOff * TS;
            const int t1 = min(e - TW * w2, TW);
            for (int t = max(b - TW * w2, 0); t < t1; t++) sw += rp[t];
          }
        }
        if (lead) {
          double lg;
          switch (lossId) {
            case FNETGPU_LOSS_RMS: lg = sqrt(ss / nG); break;
            case FNETGPU_LOSS_MAPE: lg = 100.0 * ss / nG; break;
            default: lg = ss / nG;
          }
          lossPart[2 * (size_t)lgStruct] = lgW * sw * lg;
          lossPart[2 * (size_t)lgStruct + 1] = lgW * sw;
        }
        if (!mine && lane < TA)
          for (int k = 0; k < net.nOut; k++) dL[k * TS + lane] = 0.0;  // padding atoms
      }
    }
    if (warp < nIn) {
      if (MODE == 0) {
        // ---- output layer (linear): delta = lossgrad * scale (network.F90:276, bpnn.F90:446,677-701) ----
        double *dL = T + m.dOff[L - 1] * TS;
        if (!FUSED && lane < TA) {
          const int s = lgStruct;
          const double scale = lgScale;
          for (int k = 0; k < net.nOut; k++) {
            double gr = 0.0;
            if (myAtom >= 0) {
              if (k < nG) gr = (k == 0) ? lgG0 : (k == 1) ? lgG1 : gS[(size_t)nG * s + k];
              else gr = loss_grad_fn(lossId, T[(m.aOff[L - 1] + k) * TS + lane], at[(size_t)nA * myAtom + (k - nG)]);
            }
            dL[k * TS + lane] = gr * scale;
          }
        }
        __syncwarp();
        // ---- hidden layers: delta_{l} = (W_l delta_{l+1}) * f'(z_l), l = L-2 .. 1 ----
        for (int l = L - 2; l >= 1; l--) {
          mma_backward(net.dims[l], net.dims[l + 1], wsm + m.wOff[l], m.wS[l], T + m.dOff[l + 1] * TS, T + m.dOff[l] * TS,
                       T + m.dOff[l] * TS, lane);
          __syncwarp();
        }
      }
      if (MODE == 1) {
        // ---- input gradients dE_k/dG (bpnn.F90:904-997 nJacobian, one reverse sweep per output k): f'(z_l) stays in
        // the delta rows, the deltas of a sweep go to the activation rows (not needed after the forward sweep) ----
        for (int sweep = 0; sweep < net.nOut; sweep++) {
          double *dL = T + m.dOff[L - 1] * TS;
          if (lane < TA)
            for (int k = 0; k < net.nOut; k++) dL[k * TS + lane] = (k == sweep && lane < count) ? 1.0 : 0.0;
          __syncwarp();
          for (int l = L - 2; l >= 1; l--) {
            const double *dn = (l == L - 2) ? dL : T + m.aOff[l + 1] * TS;
            mma_backward(net.dims[l], net.dims[l + 1], wsm + m.wOff[l], m.wS[l], dn, T + m.dOff[l] * TS, T + m.aOff[l] * TS, lane);
            __syncwarp();
          }
          mma_backward(d0, net.dims[1], wsm + m.wOff[0], m.wS[0], (L == 2) ? dL : T + m.aOff[1] * TS, nullptr, T + m.aOff[0] * TS, lane);
          __syncwarp();
          for (int f0 = 0; f0 < d0; f0 += 32) {
            const int f = f0 + lane;
#pragma unroll
            for (int u = 0; u < TA; u++) {
              const int atom = __shfl_sync(0xffffffffu, myAtom, u);
              if (atom >= 0 && f < d0) raw[((size_t)net.nOut * atom + sweep) * d0 + f] = T[f * TS + u];
            }
          }
          __syncwarp();
        }
      }
    }
    // atoms of round r + 2; their feature rows -> L2 (128-byte lines, four lanes per atom)
    int e2[4];
    e2[0] = __shfl_sync(0xffffffffu, ev2, 0); e2[1] = __shfl_sync(0xffffffffu, ev2, 1); e2[2] = __shfl_sync(0xffffffffu, ev2, 2);
    e2[3] = FUSED == 2 ? __shfl_sync(0xffffffffu, ev2, 3) : 0;
    const int atom2 = load_atom(e2);
    if (MODE == 0) {
      __syncthreads();
      load_features(atom1);                              // next round's rows: in flight during the gradient sweep
      // ---- weight gradients of the round: K = atoms of all tiles of the round ----
      for (int w2 = 0; w2 < nTl; w2++) {
        const double *tb = tiles0 + (size_t)w2 * rows * TS + g * TS + c;
        MmaWgradDispatch<NSLOT, NSLOT>::run(nMine, tb, aRow, dRow, newA, acc);
        if (bRow >= 0) {
          const double *dr = tiles0 + (size_t)w2 * rows * TS + bRow * TS;
          double sb = 0.0;
#pragma unroll
          for (int t = 0; t < TW; t++) sb += dr[t];
          bacc += sb;
        }
      }
      __syncthreads();
    }
    if (MODE != 0 && warp >= nIn) load_features(atom1);   // idle in this round (short last round of a species)
    {
      const int pa = __shfl_sync(0xffffffffu, atom2, lane & 7);
      if (pa >= 0) {
        const char *row = (const char *)(feat + (size_t)nFeat * pa);
        for (int b = (lane >> 3) * 128; b < d0 * 8; b += 512)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b));
      }
    }
    e0[0] = e1[0]; e0[1] = e1[1]; e0[2] = e1[2]; e0[3] = e1[3];
    e1[0] = e2[0]; e1[1] = e2[1]; e1[2] = e2[2]; e1[3] = e2[3];
    atom0 = atom1; atom1 = atom2;
    strc0 = strc1;
  }
  if (MODE == 0 && curSp >= 0) flush(curSp);
  if (FUSED == 2) cooperative_groups::this_cluster().sync();   // no CTA leaves while a peer may still read its exchange area
}
#endif   // FNET_DEFINE_MMA_KERNEL
