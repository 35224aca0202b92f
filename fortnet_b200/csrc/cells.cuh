// cells.cuh -- GPU cell list replacing the reference's brute-force neighbour iterator
// (lib_dftbp/dynneighlist.F90:233-320 + lib_dftbp/latpointiter.F90:204-253, called once per
// (atom, G-function) from lib_descriptors/acsf.F90:829-834,1070-1138).
//
// Per structure: atoms are folded into the cell (the reference does NOT fold -- its image box
// +-(floor(rc*|b_k|)+1) makes folding unnecessary; the neighbour SET {(j,T): |r_j+T-r_i| <= rc}
// is the same), binned into nb[0] x nb[1] x nb[2] bins whose thickness is >= rc along each
// reciprocal direction, sorted by (bin, atom index) -> deterministic order.  Clusters use
// their bounding box without wrap-around.
//
// Layout: one 32-byte record per atom in CELL order (position, atom index, atomic number), so a
// candidate costs two 16-byte loads.  The ACSF kernels run one CTA per bin: all atoms of a bin
// share the same 27 (or (2D+1)^3) neighbour cells, so the candidate records are staged once per
// CTA in shared memory (shift applied) and every warp filters them for its own central atom.
#pragma once
#include "internal.h"

#define FNET_SHIFT_FLAG 0x40000000   // staged candidate is a periodic image (lattice shift != 0)
#define FNET_MAX_NCELLS 128          // neighbour cells per bin the staged path can hold

struct __align__(16) CRec {
  double x, y, z;
  int idx;          // atom index (dataset order)
  int zs;           // atomic number (| FNET_SHIFT_FLAG in staged copies)
};

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

#ifndef FNET_KERNEL_TU   // non-template kernels: defined once, in fnetgpu.cu (kernels_*.cu set FNET_KERNEL_TU)
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

// single-block exclusive scan (runs once per geometry upload); flags[4] = largest bin population
__global__ void k_bin_scan(int n, const int *__restrict__ count, int *__restrict__ start, int *__restrict__ flags) {
  __shared__ int wsum[32];
  __shared__ int carry;
  __shared__ int smax;
  if (threadIdx.x == 0) { carry = 0; smax = 0; }
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int mymax = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < n) ? count[i] : 0;
    mymax = max(mymax, v);
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
  atomicMax(&smax, mymax);
  __syncthreads();
  if (threadIdx.x == 0) { start[n] = carry; flags[4] = smax; }
}

__global__ void k_bin_fill(int N, const int *__restrict__ atomCell, const int *__restrict__ cellStart,
                           int *__restrict__ cursor, int *__restrict__ cellAtoms) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c = atomCell[i];
  int slot = cellStart[c] + atomicAdd(&cursor[c], 1);
  cellAtoms[slot] = i;
}

// one thread per bin: order the bin by atom index (deterministic), then write the records
__global__ void k_bin_sort(int nBins, const int *__restrict__ cellStart, int *__restrict__ cellAtoms,
                           const double *__restrict__ fpos, const int *__restrict__ atnum,
                           CRec *__restrict__ crec) {
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
    CRec r;
    r.x = fpos[3 * j]; r.y = fpos[3 * j + 1]; r.z = fpos[3 * j + 2];
    r.idx = j; r.zs = atnum[j];
    crec[a] = r;
  }
}
#endif

// ------------------------------------------------------------------------------------------
// Neighbour cells of a bin.  The bin's (2D0+1)(2D1+1)(2D2+1) neighbour cells are numbered
// c = 0..ncells-1; each maps to a global bin (wrapped for periodic cells) plus a lattice shift.
// ------------------------------------------------------------------------------------------
struct BinPos { int b0, b1, b2, w0, w1, w2, ncells; };

__device__ __forceinline__ BinPos bin_pos(const StructInfo &S, int bin) {
  BinPos p;
  const int cell = bin - S.binBase;
  p.b2 = cell % S.nb[2];
  p.b1 = (cell / S.nb[2]) % S.nb[1];
  p.b0 = cell / (S.nb[2] * S.nb[1]);
  p.w0 = 2 * S.D[0] + 1; p.w1 = 2 * S.D[1] + 1; p.w2 = 2 * S.D[2] + 1;
  p.ncells = p.w0 * p.w1 * p.w2;
  return p;
}

struct NCell { int beg, len, shifted; double sx, sy, sz; };

__device__ __forceinline__ NCell neighbor_cell(const StructInfo &S, const BinPos &p, int c,
                                               const int *__restrict__ cellStart) {
  NCell nc;
  nc.beg = 0; nc.len = 0; nc.shifted = 0; nc.sx = 0.0; nc.sy = 0.0; nc.sz = 0.0;
  if (c >= p.ncells) return nc;
  const int d2 = c % p.w2 - S.D[2];
  const int d1 = (c / p.w2) % p.w1 - S.D[1];
  const int d0 = c / (p.w2 * p.w1) - S.D[0];
  int g0 = p.b0 + d0, g1 = p.b1 + d1, g2 = p.b2 + d2;
  int s0 = 0, s1 = 0, s2 = 0;
  if (S.periodic) {
    s0 = floor_div(g0, S.nb[0]); g0 -= s0 * S.nb[0];
    s1 = floor_div(g1, S.nb[1]); g1 -= s1 * S.nb[1];
    s2 = floor_div(g2, S.nb[2]); g2 -= s2 * S.nb[2];
  } else if (g0 < 0 || g0 >= S.nb[0] || g1 < 0 || g1 >= S.nb[1] || g2 < 0 || g2 >= S.nb[2]) {
    return nc;
  }
  const int gc = S.binBase + (g0 * S.nb[1] + g1) * S.nb[2] + g2;
  nc.beg = cellStart[gc];
  nc.len = cellStart[gc + 1] - nc.beg;
  nc.sx = s0 * S.lat[0] + s1 * S.lat[3] + s2 * S.lat[6];
  nc.sy = s0 * S.lat[1] + s1 * S.lat[4] + s2 * S.lat[7];
  nc.sz = s0 * S.lat[2] + s1 * S.lat[5] + s2 * S.lat[8];
  nc.shifted = (s0 != 0 || s1 != 0 || s2 != 0) ? 1 : 0;
  return nc;
}

// CTA-cooperative staging of all candidate records of a bin's neighbour cells (shift applied).
// Each neighbour cell is resolved once (thread c < ncells) into the shared tables; warp w then
// copies cells w, w+nWarps, ...  Returns the candidate count (or -needed on overflow, or -1 when
// the bin has more than FNET_MAX_NCELLS neighbour cells).
struct StageTabs {                       // lives in shared memory behind the candidate records
  double sh[FNET_MAX_NCELLS][3];
  int beg[FNET_MAX_NCELLS];              // first record; bit 31 set when the lattice shift is non-zero
  int len[FNET_MAX_NCELLS];
  int pre[FNET_MAX_NCELLS + 4];
};

__device__ __forceinline__ int stage_candidates(const StructInfo &S, const BinPos &p,
                                                const int *__restrict__ cellStart,
                                                const CRec *__restrict__ crec, CRec *__restrict__ cand,
                                                int capC, StageTabs *__restrict__ tb) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (p.ncells > FNET_MAX_NCELLS) return -1;
  for (int c = threadIdx.x; c < FNET_MAX_NCELLS; c += blockDim.x) {
    int len = 0;
    if (c < p.ncells) {
      const NCell nc = neighbor_cell(S, p, c, cellStart);
      len = nc.len;
      tb->beg[c] = nc.beg | (nc.shifted ? 0x80000000 : 0);
      tb->sh[c][0] = nc.sx; tb->sh[c][1] = nc.sy; tb->sh[c][2] = nc.sz;
    }
    tb->len[c] = len;
  }
  __syncthreads();
  if (wib == 0) {   // exclusive scan over FNET_MAX_NCELLS = 4 x 32 entries
    int carry = 0;
#pragma unroll
    for (int q = 0; q < FNET_MAX_NCELLS / 32; q++) {
      const int v = tb->len[q * 32 + lane];
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      tb->pre[q * 32 + lane] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) tb->pre[FNET_MAX_NCELLS] = carry;
  }
  __syncthreads();
  const int total = tb->pre[FNET_MAX_NCELLS];
  if (total > capC) return -total;
  for (int c = wib; c < p.ncells; c += nw) {
    const int len = tb->len[c];
    if (len == 0) continue;
    const int begf = tb->beg[c], beg = begf & 0x7fffffff;
    const int flag = (begf < 0) ? FNET_SHIFT_FLAG : 0;
    const double sx = tb->sh[c][0], sy = tb->sh[c][1], sz = tb->sh[c][2];
    const int dst = tb->pre[c];
    for (int t = lane; t < len; t += 32) {
      CRec r = crec[beg + t];
      r.x += sx; r.y += sy; r.z += sz;
      r.zs |= flag;
      cand[dst + t] = r;
    }
  }
  __syncthreads();
  return total;
}

// Warp-cooperative enumeration of the candidates around the central atom `me`.  visit(valid, dx,
// dy, dz, idx, zs) is called with all 32 lanes converged; (dx, dy, dz) is the displacement from
// the central atom to the candidate image and zs carries FNET_SHIFT_FLAG for periodic images.
template <typename Visit>
__device__ __forceinline__ void for_each_candidate_staged(const CRec *__restrict__ cand, int total, const CRec &me,
                                                          Visit visit) {
  const int lane = threadIdx.x & 31;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    const bool valid = t < total;
    CRec r;
    r.x = 0.0; r.y = 0.0; r.z = 0.0; r.idx = -1; r.zs = 0;
    if (valid) r = cand[t];
    visit(valid, r.x - me.x, r.y - me.y, r.z - me.z, r.idx, r.zs);
  }
}

// direct variant (no staging): lane l owns neighbour cell (cbase + l) and walks its records
template <typename Visit>
__device__ __forceinline__ void for_each_candidate_direct(const StructInfo &S, const BinPos &p,
                                                          const int *__restrict__ cellStart,
                                                          const CRec *__restrict__ crec, const CRec &me,
                                                          Visit visit) {
  const int lane = threadIdx.x & 31;
  for (int cbase = 0; cbase < p.ncells; cbase += 32) {
    const NCell nc = neighbor_cell(S, p, cbase + lane, cellStart);
    int maxlen = nc.len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    for (int t = 0; t < maxlen; t++) {
      const bool valid = t < nc.len;
      CRec r;
      r.x = 0.0; r.y = 0.0; r.z = 0.0; r.idx = -1; r.zs = 0;
      if (valid) {
        r = crec[nc.beg + t];
        r.x += nc.sx; r.y += nc.sy; r.z += nc.sz;
        if (nc.shifted) r.zs |= FNET_SHIFT_FLAG;
      }
      visit(valid, r.x - me.x, r.y - me.y, r.z - me.z, r.idx, r.zs);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Whole-structure path for small structures (the usual training-set shape: tens to a few
// hundred atoms): no cell list at all.  One CTA stages every atom of its structure in shared
// memory; each central atom tests all of them and picks the periodic image by the minimum-image
// convention  d <- d - lat * rint(inv * d).  That is exact (the unique image within rc) when every
// lattice-plane spacing h_k >= 2 rc: an image with |d| <= rc has |inv_k . d| <= rc / h_k <= 1/2.
// The same bound excludes periodic images of the central atom itself.  stage_structure checks it
// and raises flags[4] otherwise (the host then falls back to the cell list).  Clusters: no wrap.
// ------------------------------------------------------------------------------------------
#define FNET_STRUCT_MAX_ATOMS 256    // candidates per central atom grow with the structure: beyond this the cell list wins

struct __align__(16) StructGeom {
  double lat[9];     // lat[3*k+c]: component c of lattice vector k
  double inv[9];     // fractional s_k = sum_c inv[3*k+c] * d_c
  int periodic;
  int diag;          // orthorhombic cell along the axes: the minimum image decouples per component
};

__device__ __forceinline__ int stage_structure(int st, int beg, int nAt, const double *__restrict__ coords,
                                               const int *__restrict__ atnum, const double *__restrict__ lat,
                                               const int *__restrict__ periodic, double rc, CRec *__restrict__ cand,
                                               int capC, StructGeom *__restrict__ g, int *__restrict__ flags) {
  if (nAt > capC) return -nAt;
#pragma unroll 1
  for (int t = threadIdx.x; t < nAt; t += blockDim.x) {
    const size_t j = (size_t)beg + t;
    CRec r;
    r.x = coords[3 * j]; r.y = coords[3 * j + 1]; r.z = coords[3 * j + 2];
    r.idx = (int)j; r.zs = atnum[j];
    cand[t] = r;
  }
  if (threadIdx.x == blockDim.x - 1) {
    const int per = periodic[st];
    g->periodic = per;
    if (per) {
      double m[3][3];                                  // m[c][k] = component c of lattice vector k
      for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) { m[c][k] = lat[(size_t)9 * st + 3 * k + c]; g->lat[3 * k + c] = m[c][k]; }
      const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2],
                   c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
      const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
      bool ok = fabs(det) >= 1e-12;
      const double id = ok ? 1.0 / det : 0.0;
      double r[3][3];                                  // r = m^-1: s_k = sum_c r[k][c] d_c
      r[0][0] = c00 * id; r[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; r[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
      r[1][0] = c01 * id; r[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; r[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
      r[2][0] = c02 * id; r[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; r[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
      for (int k = 0; k < 3; k++) {
        g->inv[3 * k] = r[k][0]; g->inv[3 * k + 1] = r[k][1]; g->inv[3 * k + 2] = r[k][2];
        // plane spacing h_k = 1 / |row k of m^-1|;  need h_k >= 2 rc (with a little slack)
        const double bn2 = r[k][0] * r[k][0] + r[k][1] * r[k][1] + r[k][2] * r[k][2];
        if (!(bn2 * (4.0 * rc * rc) * (1.0 + 1e-9) <= 1.0)) ok = false;
      }
      if (!ok) atomicOr(&flags[4], 1);
      g->diag = (m[0][1] == 0.0 && m[0][2] == 0.0 && m[1][0] == 0.0 && m[1][2] == 0.0 && m[2][0] == 0.0 && m[2][1] == 0.0) ? 1 : 0;
    } else {
      for (int e = 0; e < 9; e++) { g->lat[e] = 0.0; g->inv[e] = 0.0; }
      g->diag = 0;
    }
  }
  __syncthreads();
  return nAt;
}

template <typename Visit>
__device__ __forceinline__ void for_each_candidate_struct(const CRec *__restrict__ cand, int total,
                                                          const StructGeom *__restrict__ g, const CRec &me,
                                                          Visit visit) {
  const int lane = threadIdx.x & 31;
  const bool per = g->periodic != 0;
  const bool diag = g->diag != 0;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    const bool valid = t < total;
    CRec r;
    r.x = me.x; r.y = me.y; r.z = me.z; r.idx = -1; r.zs = 0;
    if (valid) r = cand[t];
    double dx = r.x - me.x, dy = r.y - me.y, dz = r.z - me.z;
    const double magic = 6755399441055744.0;           // 1.5 * 2^52: rint() through the adder
    if (diag) {                                        // same arithmetic as below with the zero terms dropped
      const double n0 = (g->inv[0] * dx + magic) - magic;
      const double n1 = (g->inv[4] * dy + magic) - magic;
      const double n2 = (g->inv[8] * dz + magic) - magic;
      dx -= n0 * g->lat[0];
      dy -= n1 * g->lat[4];
      dz -= n2 * g->lat[8];
    } else if (per) {
      const double n0 = (g->inv[0] * dx + g->inv[1] * dy + g->inv[2] * dz + magic) - magic;
      const double n1 = (g->inv[3] * dx + g->inv[4] * dy + g->inv[5] * dz + magic) - magic;
      const double n2 = (g->inv[6] * dx + g->inv[7] * dy + g->inv[8] * dz + magic) - magic;
      dx -= n0 * g->lat[0] + n1 * g->lat[3] + n2 * g->lat[6];
      dy -= n0 * g->lat[1] + n1 * g->lat[4] + n2 * g->lat[7];
      dz -= n0 * g->lat[2] + n1 * g->lat[5] + n2 * g->lat[8];
    }
    visit(valid, dx, dy, dz, r.idx, r.zs);             // no shift flag: j == i is the central atom itself
  }
}

// neighbour test of the reference: d2 <= rc2, the central atom itself excluded in the central
// cell only (dynneighlist.F90:271-276,294)
__device__ __forceinline__ bool is_neighbor(bool valid, double d2, double rc2, int j, int zs, int i) {
  return valid && (d2 <= rc2) && !(j == i && !(zs & FNET_SHIFT_FLAG));
}

#ifndef FNET_KERNEL_TU   // non-template kernels: defined once, in fnetgpu.cu (kernels_*.cu set FNET_KERNEL_TU)
// statistics that size the shared-memory buffers (one warp per atom, cell order):
// flags[0] = max neighbours per atom, flags[2..3] = sum of neighbours (64 bit),
// flags[5] = max candidates per bin, flags[6] = max neighbour cells per bin
__global__ void k_neigh_count(int N, const int *__restrict__ binOfSlot_unused, const int *__restrict__ binStruct,
                              const StructInfo *__restrict__ sinfo, const int *__restrict__ atomCell,
                              const int *__restrict__ cellStart, const CRec *__restrict__ crec,
                              double rc2, int *__restrict__ flags) {
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= N) return;
  const CRec me = crec[slot];
  const int bin = atomCell[me.idx];
  const StructInfo &S = sinfo[binStruct[bin]];
  const BinPos p = bin_pos(S, bin);
  int n = 0, ncand = 0;
  for_each_candidate_direct(S, p, cellStart, crec, me, [&](bool valid, double dx, double dy, double dz, int j, int zs) {
    const bool ok = is_neighbor(valid, dx * dx + dy * dy + dz * dz, rc2, j, zs, me.idx);
    n += __popc(__ballot_sync(0xffffffffu, ok));
    ncand += __popc(__ballot_sync(0xffffffffu, valid));
  });
  if (lane == 0) {
    atomicMax(&flags[0], n);
    atomicAdd((unsigned long long *)(flags + 2), (unsigned long long)n);
    atomicMax(&flags[5], ncand);
    atomicMax(&flags[6], p.ncells);
  }
}
#endif
