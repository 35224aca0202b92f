// cells.cuh -- GPU cell list replacing the reference's brute-force neighbour iterator
// (lib_dftbp/dynneighlist.F90:233-320 + lib_dftbp/latpointiter.F90:204-253, called once per
// (atom, G-function) from lib_descriptors/acsf.F90:829-834,1070-1138).
//
// Per structure: atoms are folded into the cell (the reference does NOT fold -- its image box
// +-(floor(rc*|b_k|)+1) makes folding unnecessary; the neighbour SET {(j,T): |r_j+T-r_i| <= rc}
// is the same), binned into nb[0] x nb[1] x nb[2] bins whose thickness is >= rc along each
// reciprocal direction, sorted by (bin, atom index) -> deterministic order.  Clusters use
// their bounding box without wrap-around.
#pragma once
#include "internal.h"

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// one thread per atom: fold, bin, histogram
__global__ void k_bin_count(int N, const double *__restrict__ coords, const int *__restrict__ structOf,
                            const StructInfo *__restrict__ sinfo, double *__restrict__ fpos,
                            int *__restrict__ atomCell, int *__restrict__ cellCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const StructInfo &S = sinfo[structOf[i]];
  double r[3] = {coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]};
  double s[3];
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; k++)
    s[k] = S.inv[3 * k] * (r[0] - S.lo[0]) + S.inv[3 * k + 1] * (r[1] - S.lo[1]) + S.inv[3 * k + 2] * (r[2] - S.lo[2]);
  if (S.periodic) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double fl = floor(s[k]);
      s[k] -= fl;
#pragma unroll
      for (int c = 0; c < 3; c++) r[c] -= fl * S.lat[3 * k + c];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int bb = (int)(s[k] * S.nb[k]);
    b[k] = min(max(bb, 0), S.nb[k] - 1);
  }
  fpos[3 * i] = r[0]; fpos[3 * i + 1] = r[1]; fpos[3 * i + 2] = r[2];
  int cell = S.binBase + (b[0] * S.nb[1] + b[1]) * S.nb[2] + b[2];
  atomCell[i] = cell;
  atomicAdd(&cellCount[cell], 1);
}

// single-block exclusive scan (runs once per geometry upload)
__global__ void k_bin_scan(int n, const int *__restrict__ count, int *__restrict__ start) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < n) ? count[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int s = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      wsum[lane] = s;
    }
    __syncthreads();
    int off = carry + (w > 0 ? wsum[w - 1] : 0);
    if (i < n) start[i] = off + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}

__global__ void k_bin_fill(int N, const int *__restrict__ atomCell, const int *__restrict__ cellStart,
                           int *__restrict__ cursor, int *__restrict__ cellAtoms) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c = atomCell[i];
  int slot = cellStart[c] + atomicAdd(&cursor[c], 1);
  cellAtoms[slot] = i;
}

// one thread per bin: order the bin by atom index (deterministic), then gather positions
__global__ void k_bin_sort(int nBins, const int *__restrict__ cellStart, int *__restrict__ cellAtoms,
                           const double *__restrict__ fpos, double *__restrict__ cpos) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nBins) return;
  int b = cellStart[c], e = cellStart[c + 1];
  for (int a = b + 1; a < e; a++) {
    int v = cellAtoms[a];
    int p = a - 1;
    while (p >= b && cellAtoms[p] > v) { cellAtoms[p + 1] = cellAtoms[p]; p--; }
    cellAtoms[p + 1] = v;
  }
  for (int a = b; a < e; a++) {
    int j = cellAtoms[a];
    cpos[3 * a] = fpos[3 * j]; cpos[3 * a + 1] = fpos[3 * j + 1]; cpos[3 * a + 2] = fpos[3 * j + 2];
  }
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative neighbour enumeration around atom i.  Calls visit(ok, dx, dy, dz, d2, j, selfImage)
// with all 32 lanes converged; `ok` marks lanes that hold a neighbour within rc.
// Lane l owns neighbour cell (cbase + l) and walks its atoms, so each step tests up to 32
// candidates from different cells.
// ------------------------------------------------------------------------------------------
template <typename Visit>
__device__ __forceinline__ void for_each_neighbor(int i, const StructInfo &S, const int *__restrict__ atomCell,
                                                  const int *__restrict__ cellStart,
                                                  const int *__restrict__ cellAtoms,
                                                  const double *__restrict__ fpos,
                                                  const double *__restrict__ cpos, double rc2, Visit visit) {
  const int lane = threadIdx.x & 31;
  const double rix = fpos[3 * i], riy = fpos[3 * i + 1], riz = fpos[3 * i + 2];
  int cell = atomCell[i] - S.binBase;
  const int b2 = cell % S.nb[2];
  const int b1 = (cell / S.nb[2]) % S.nb[1];
  const int b0 = cell / (S.nb[2] * S.nb[1]);
  const int w0 = 2 * S.D[0] + 1, w1 = 2 * S.D[1] + 1, w2 = 2 * S.D[2] + 1;
  const int ncells = w0 * w1 * w2;
  for (int cbase = 0; cbase < ncells; cbase += 32) {
    int c = cbase + lane;
    bool valid = c < ncells;
    int beg = 0, end = 0;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    bool zeroShift = true;
    if (valid) {
      int d2 = c % w2 - S.D[2];
      int d1 = (c / w2) % w1 - S.D[1];
      int d0 = c / (w2 * w1) - S.D[0];
      int g0 = b0 + d0, g1 = b1 + d1, g2 = b2 + d2;
      int s0 = 0, s1 = 0, s2 = 0;
      if (S.periodic) {
        s0 = floor_div(g0, S.nb[0]); g0 -= s0 * S.nb[0];
        s1 = floor_div(g1, S.nb[1]); g1 -= s1 * S.nb[1];
        s2 = floor_div(g2, S.nb[2]); g2 -= s2 * S.nb[2];
      } else {
        valid = g0 >= 0 && g0 < S.nb[0] && g1 >= 0 && g1 < S.nb[1] && g2 >= 0 && g2 < S.nb[2];
      }
      if (valid) {
        int gc = S.binBase + (g0 * S.nb[1] + g1) * S.nb[2] + g2;
        beg = cellStart[gc];
        end = cellStart[gc + 1];
        sx = s0 * S.lat[0] + s1 * S.lat[3] + s2 * S.lat[6];
        sy = s0 * S.lat[1] + s1 * S.lat[4] + s2 * S.lat[7];
        sz = s0 * S.lat[2] + s1 * S.lat[5] + s2 * S.lat[8];
        zeroShift = (s0 == 0 && s1 == 0 && s2 == 0);
      }
    }
    int len = end - beg;
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    for (int t = 0; t < maxlen; t++) {
      bool ok = false;
      double dx = 0, dy = 0, dz = 0, d2 = 0;
      int j = -1;
      if (t < len) {
        int slot = beg + t;
        j = cellAtoms[slot];
        dx = (cpos[3 * slot] + sx) - rix;
        dy = (cpos[3 * slot + 1] + sy) - riy;
        dz = (cpos[3 * slot + 2] + sz) - riz;
        d2 = dx * dx + dy * dy + dz * dz;
        ok = (d2 <= rc2) && !(j == i && zeroShift);   // dynneighlist.F90:271-276,294
      }
      visit(ok, dx, dy, dz, d2, j);
    }
  }
}

// neighbour count per atom (one warp per atom) -> sizes the shared-memory neighbour buffers
__global__ void k_neigh_count(int N, const int *__restrict__ structOf, const StructInfo *__restrict__ sinfo,
                              const int *__restrict__ atomCell, const int *__restrict__ cellStart,
                              const int *__restrict__ cellAtoms, const double *__restrict__ fpos,
                              const double *__restrict__ cpos, double rc2, int *__restrict__ neighCount,
                              int *__restrict__ flags /* [0]=max, [1..2]=sum lo/hi */) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const StructInfo &S = sinfo[structOf[warp]];
  int n = 0;
  for_each_neighbor(warp, S, atomCell, cellStart, cellAtoms, fpos, cpos, rc2,
                    [&](bool ok, double, double, double, double, int) {
                      n += __popc(__ballot_sync(0xffffffffu, ok));
                    });
  if (lane == 0) {
    neighCount[warp] = n;
    atomicMax(&flags[0], n);
    atomicAdd((unsigned long long *)(flags + 2), (unsigned long long)n);
  }
}
