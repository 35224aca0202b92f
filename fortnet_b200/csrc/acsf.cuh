// acsf.cuh -- ACSF values: one warp per central atom.
//
// Replaces TAcsf_calculate -> iGeoAcsf -> buildGFunctionNeighborlists + g1..g5
// (lib_descriptors/acsf.F90:540-639, 797-866, 942-1066, 1287-1492).  The reference rebuilds
// the neighbour list for every (atom, function) and evaluates acos->cos, pow, exp per pair
// per function; here the list is gathered once per atom from the cell list into shared
// memory, sorted by species code, and each angular pass shares the pair geometry between all
// functions with equal (type, rc, eta, species pair): (1+lam*cos)^xi is evaluated for a whole
// xi-ladder as exp(xi0*L) * exp(dxi*L)^m with L = log(1+lam*cos) (one log + two exp per
// ladder per pair), unordered pairs are visited once (x2) where the reference walks ordered
// pairs.  Parity traps kept: neighbour test d2 <= rc^2 (dynneighlist.F90:294), cos(theta)
// denominator r_j r_k + 1e-13 (acsf.F90:1174), diagonal j==k pairs of identical lists
// (acsf.F90:1420-1431,1480-1487), periodic images of the central atom belong to EVERY
// species list (reduceGeometrySpecies keeps iAtom, acsf.F90:754-760), atom-id prefactors
// q_i q_j in the cutoffs and q_j q_k in G4's third cutoff (acsf.F90:1201,1428-1430).
#pragma once
#include "cells.cuh"

struct WarpSmem {
  double *dx, *dy, *dz, *r, *rinv, *qv, *fcE;
  int *idx, *seg;
  double *outv;
};

__host__ __device__ inline size_t acsf_warp_smem_bytes(int cap, int F) {
  size_t b = (size_t)cap * (7 * sizeof(double) + sizeof(int));
  b += (FNET_MAX_CODES + 4) * sizeof(int);
  b = (b + 7) & ~(size_t)7;
  b += (size_t)((F + 1) & ~1) * sizeof(double);
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ WarpSmem carve_warp_smem(unsigned char *base, int cap, int F) {
  WarpSmem w;
  double *d = (double *)base;
  w.dx = d; w.dy = d + cap; w.dz = d + 2 * cap; w.r = d + 3 * cap; w.rinv = d + 4 * cap;
  w.qv = d + 5 * cap; w.fcE = d + 6 * cap;
  w.idx = (int *)(d + 7 * cap);
  w.seg = w.idx + cap;
  size_t off = (size_t)cap * (7 * sizeof(double) + sizeof(int)) + (FNET_MAX_CODES + 4) * sizeof(int);
  off = (off + 7) & ~(size_t)7;
  w.outv = (double *)(base + off);
  return w;
}

__device__ __forceinline__ int species_code(const AcsfTables &tab, int z) {
  int code = tab.nCodes;  // "other"
#pragma unroll 1
  for (int c = 0; c < tab.nCodes; c++)
    if (tab.zcodes[c] == z) code = c;
  return code;
}

// Gathers the neighbours of atom i (within rcMax) into the warp's shared memory, sorted by
// species code [code 0 .. nCodes-1 | other | self-images]; returns n (or -needed on overflow).
__device__ __forceinline__ int gather_neighbors(int i, const StructInfo &S, const AcsfTables &tab,
                                                const int *__restrict__ atomCell,
                                                const int *__restrict__ cellStart,
                                                const int *__restrict__ cellAtoms,
                                                const double *__restrict__ fpos,
                                                const double *__restrict__ cpos,
                                                const int *__restrict__ atnum, int cap, WarpSmem &w) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const bool sorted = tab.nCodes > 0;
  // unsorted target = final arrays; sorted target = scratch aliased on (fcE, qv, r | rinv)
  double *gx = sorted ? w.fcE : w.dx, *gy = sorted ? w.qv : w.dy, *gz = sorted ? w.r : w.dz;
  int *gj = sorted ? (int *)w.rinv : w.idx;
  int n = 0;
  for_each_neighbor(i, S, atomCell, cellStart, cellAtoms, fpos, cpos, tab.rcMax * tab.rcMax,
                    [&](bool ok, double dx, double dy, double dz, double, int j) {
                      unsigned m = __ballot_sync(0xffffffffu, ok);
                      int pos = n + __popc(m & lt);
                      if (ok && pos < cap) { gx[pos] = dx; gy[pos] = dy; gz[pos] = dz; gj[pos] = j; }
                      n += __popc(m);
                    });
  if (n > cap) return -n;
  __syncwarp();
  if (sorted) {
    const int nc = tab.nCodes + 2;  // + other + self
    int mycount = 0;
    for (int base = 0; base < n; base += 32) {
      int t = base + lane;
      int code = -1;
      if (t < n) { int j = gj[t]; code = (j == i) ? tab.nCodes + 1 : species_code(tab, atnum[j]); }
      for (int c = 0; c < nc; c++) {
        unsigned m = __ballot_sync(0xffffffffu, code == c);
        if (lane == c) mycount += __popc(m);
      }
    }
    int incl = mycount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    int mybase = incl - mycount;
    if (lane <= nc) w.seg[lane] = (lane < nc) ? mybase : n;
    for (int base = 0; base < n; base += 32) {
      int t = base + lane;
      int code = -1, j = -1;
      double x = 0, y = 0, z = 0;
      if (t < n) { j = gj[t]; x = gx[t]; y = gy[t]; z = gz[t]; code = (j == i) ? tab.nCodes + 1 : species_code(tab, atnum[j]); }
      int pos = -1;
      for (int c = 0; c < nc; c++) {
        unsigned m = __ballot_sync(0xffffffffu, code == c);
        int b = __shfl_sync(0xffffffffu, mybase, c);
        if (code == c) pos = b + __popc(m & lt);
        if (lane == c) mybase += __popc(m);
      }
      if (pos >= 0) { w.dx[pos] = x; w.dy[pos] = y; w.dz[pos] = z; w.idx[pos] = j; }
    }
    __syncwarp();
  } else if (lane == 0) {
    w.seg[0] = 0; w.seg[1] = n; w.seg[2] = n;  // [other | self] unused
  }
  for (int t = lane; t < n; t += 32) {
    double d2 = w.dx[t] * w.dx[t] + w.dy[t] * w.dy[t] + w.dz[t] * w.dz[t];
    double rr = sqrt(d2);                      // dynneighlist.F90:311
    w.r[t] = rr;
    w.rinv[t] = 1.0 / rr;
  }
  __syncwarp();
  return n;
}

// A species list = segment(code) followed by the self-image segment; code < 0 = everything.
struct NbList { int s0, n0, s1, n1; };
__device__ __forceinline__ NbList make_list(const AcsfTables &tab, const WarpSmem &w, int code, int n) {
  NbList l;
  if (code < 0) { l.s0 = 0; l.n0 = n; l.s1 = 0; l.n1 = 0; }
  else {
    int self = tab.nCodes + 1;
    l.s0 = w.seg[code]; l.n0 = w.seg[code + 1] - l.s0;
    l.s1 = w.seg[self]; l.n1 = w.seg[self + 1] - l.s1;
  }
  return l;
}
__device__ __forceinline__ int list_at(const NbList &l, int t) { return t < l.n0 ? l.s0 + t : l.s1 + (t - l.n0); }

__device__ __forceinline__ double cutoff_fn(double rr, double qq, double invrc) {
  return 0.5 * qq * (cospi(rr * invrc) + 1.0);   // acsf.F90:1201 (pi*rr/rcut)
}

// (1 + lam*c)^xi ladder start and ratio from L = log(1 + lam*c), with the pow(0,0)=1 /
// pow(0,x>0)=0 conventions (log(0) = -inf, exp(-inf) = 0)
__device__ __forceinline__ void ladder_init(double L, double xi0, double dxi, double &p, double &q) {
  p = (xi0 == 0.0) ? 1.0 : exp(xi0 * L);
  q = (dxi == 0.0) ? 1.0 : exp(dxi * L);
}

template <typename real>
__global__ void __launch_bounds__(128)
k_acsf(int N, const int *__restrict__ structOf, const StructInfo *__restrict__ sinfo,
       const int *__restrict__ atomCell, const int *__restrict__ cellStart,
       const int *__restrict__ cellAtoms, const double *__restrict__ fpos,
       const double *__restrict__ cpos, const int *__restrict__ atnum, int nExt,
       const double *__restrict__ ext, AcsfTables tab, int cap, real *__restrict__ feat, int nFeat,
       const double *__restrict__ zprec, int nExtSel, const int *__restrict__ extIdx,
       int *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wib;
  if (i >= N) return;
  WarpSmem w = carve_warp_smem(smem_raw + (size_t)wib * acsf_warp_smem_bytes(cap, tab.F), cap, tab.F);
  const StructInfo &S = sinfo[structOf[i]];
  int n = gather_neighbors(i, S, tab, atomCell, cellStart, cellAtoms, fpos, cpos, atnum, cap, w);
  if (n < 0) { if (lane == 0) atomicMax(&flags[1], -n); return; }

  // ---------------- radial groups (acsf.F90:1287-1373) ----------------
  for (int g = 0; g < tab.nRadialGroups; g++) {
    const RadialGroup G = tab.rgroups[g];
    const NbList l = make_list(tab, w, G.code, n);
    const int nl = l.n0 + l.n1;
    const double qi = G.atomId > 0 ? ext[(size_t)nExt * i + G.atomId - 1] : 1.0;
    const double invrc = 1.0 / G.rc;
    const int nch = G.nChunksP2, per = 32 / nch;
    const int mychunk = lane % nch, sub = lane / nch;
    const int fbase = G.fBeg + mychunk * FNET_RCHUNK;
    const int fcnt = min(max(G.fCnt - mychunk * FNET_RCHUNK, 0), FNET_RCHUNK);
    double p1[FNET_RCHUNK], p2[FNET_RCHUNK], acc[FNET_RCHUNK];
#pragma unroll
    for (int f = 0; f < FNET_RCHUNK; f++) {
      acc[f] = 0.0;
      p1[f] = f < fcnt ? tab.rp1[fbase + f] : 0.0;
      p2[f] = f < fcnt ? tab.rp2[fbase + f] : 0.0;
    }
    if (fcnt > 0)
      for (int t = sub; t < nl; t += per) {
        const int a = list_at(l, t);
        const double rr = w.r[a];
        if (rr > G.rc) continue;                       // cutoff1d: rr > rcut -> 0
        const double qj = G.atomId > 0 ? ext[(size_t)nExt * w.idx[a] + G.atomId - 1] : 1.0;
        const double fc = cutoff_fn(rr, qi * qj, invrc);
        if (G.type == FNETGPU_G1) {
#pragma unroll
          for (int f = 0; f < FNET_RCHUNK; f++) acc[f] += fc;
        } else if (G.type == FNETGPU_G2) {
#pragma unroll
          for (int f = 0; f < FNET_RCHUNK; f++)
            if (f < fcnt) { double d = rr - p2[f]; acc[f] += exp(-p1[f] * d * d) * fc; }
        } else {
#pragma unroll
          for (int f = 0; f < FNET_RCHUNK; f++)
            if (f < fcnt) acc[f] += cos(p1[f] * rr) * fc;
        }
      }
#pragma unroll
    for (int f = 0; f < FNET_RCHUNK; f++)
      for (int o = 16; o >= nch; o >>= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], o);
    if (sub == 0) {
#pragma unroll
      for (int f = 0; f < FNET_RCHUNK; f++)
        if (f < fcnt) w.outv[tab.rfeat[fbase + f]] = acc[f];
    }
  }

  // ---------------- angular passes (acsf.F90:1377-1492) ----------------
  for (int pi_ = 0; pi_ < tab.nAngularPasses; pi_++) {
    const AngularPass *__restrict__ P = &tab.apasses[pi_];
    const int type = P->type, same = P->same, atomId = P->atomId, nSlots = P->nSlots;
    const double rc = P->rc, eta = P->eta, invrc = 1.0 / rc;
    const NbList l1 = make_list(tab, w, P->code1, n);
    const NbList l2 = same ? l1 : make_list(tab, w, P->code2, n);
    const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
    const double qi = atomId > 0 ? ext[(size_t)nExt * i + atomId - 1] : 1.0;
    __syncwarp();
    for (int t = lane; t < n; t += 32) {   // per-neighbour radial factor fc * exp(-eta r^2)
      const double rr = w.r[t];
      const double qj = atomId > 0 ? ext[(size_t)nExt * w.idx[t] + atomId - 1] : 1.0;
      w.qv[t] = qj;
      w.fcE[t] = (rr > rc) ? 0.0 : cutoff_fn(rr, qi * qj, invrc) * exp(-eta * rr * rr);
    }
    __syncwarp();
    double lam[FNET_SLOTS], xi0[FNET_SLOTS], dxi[FNET_SLOTS];
    int cnt[FNET_SLOTS];
    double acc[FNET_SLOTS][FNET_LADDER];
#pragma unroll
    for (int s = 0; s < FNET_SLOTS; s++) {
      lam[s] = s < nSlots ? P->slot[s].lam : 0.0;
      xi0[s] = s < nSlots ? P->slot[s].xi0 : 0.0;
      dxi[s] = s < nSlots ? P->slot[s].dxi : 0.0;
      cnt[s] = s < nSlots ? P->slot[s].count : 0;
#pragma unroll
      for (int f = 0; f < FNET_LADDER; f++) acc[s][f] = 0.0;
    }
    if (n1 > 0 && n2 > 0) {
      // flattened pair walk: row j holds k = k0(j) .. k0(j)+len(j)-1
      int j = 0, o = lane;
      while (j < n1) {
        int len = same ? n1 - j : n2;
        while (o >= len) { o -= len; j++; if (j >= n1) break; len = same ? n1 - j : n2; }
        if (j >= n1) break;
        const int k = (same ? j : 0) + o;
        const int a = list_at(l1, j), b = list_at(l2, k);
        double base = w.fcE[a] * w.fcE[b];
        if (same && a != b) base *= 2.0;
        if (type == FNETGPU_G4 && base != 0.0) {
          const double ex = w.dx[a] - w.dx[b], ey = w.dy[a] - w.dy[b], ez = w.dz[a] - w.dz[b];
          const double djk2 = ex * ex + ey * ey + ez * ez;
          const double djk = sqrt(djk2);
          base = (djk > rc) ? 0.0 : base * exp(-eta * djk2) * cutoff_fn(djk, w.qv[a] * w.qv[b], invrc);
        }
        if (base != 0.0) {
          const double dot = w.dx[a] * w.dx[b] + w.dy[a] * w.dy[b] + w.dz[a] * w.dz[b];
          const double pr = w.rinv[a] * w.rinv[b];
          const double c = dot * (pr * (1.0 - 1e-13 * pr));   // dot / (r_a r_b + 1e-13), acsf.F90:1173-1174
          double L = 0.0;
#pragma unroll
          for (int s = 0; s < FNET_SLOTS; s++) {
            if (cnt[s] > 0) {
              if (s == 0 || lam[s] != lam[s - 1]) L = log(fmax(1.0 + lam[s] * c, 0.0));
              double p, q;
              ladder_init(L, xi0[s], dxi[s], p, q);
              p *= base;
#pragma unroll
              for (int f = 0; f < FNET_LADDER; f++) {
                if (f < cnt[s]) acc[s][f] += p;
                p *= q;
              }
            }
          }
        }
        o += 32;
      }
    }
#pragma unroll
    for (int s = 0; s < FNET_SLOTS; s++) {
      if (cnt[s] > 0) {
#pragma unroll
        for (int f = 0; f < FNET_LADDER; f++) {
          double v = acc[s][f];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && f < cnt[s]) w.outv[P->slot[s].feat[f]] = v * P->slot[s].pref[f];
        }
      }
    }
  }
  __syncwarp();

  // ---------------- coalesced feature write (+ z-score, + external features) ----------------
  real *out = feat + (size_t)nFeat * i;
  for (int a = lane; a < tab.F; a += 32) {
    double v = w.outv[a];
    if (zprec) {
      const double sg = zprec[tab.F + a];
      if (!(sg < 1e-08)) v = (v - zprec[a]) / sg;   // acsf.F90:505-507
    }
    out[a] = (real)v;
  }
  for (int e = lane; e < nExtSel; e += 32) out[tab.F + e] = (real)ext[(size_t)nExt * i + extIdx[e]];
}

// external features only (no ACSF functions configured): features.F90:227-241
template <typename real>
__global__ void k_ext_concat(int N, int nExt, const double *__restrict__ ext, int F, int nExtSel,
                             const int *__restrict__ extIdx, real *__restrict__ feat, int nFeat) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * nExtSel) return;
  int i = t / nExtSel, e = t % nExtSel;
  feat[(size_t)nFeat * i + F + e] = (real)ext[(size_t)nExt * i + extIdx[e]];
}

// ------------------------------------------------------------------------------------------
// z-score statistics (acsf.F90:445-486): fixed-shape two-level sums -> deterministic.
// pass 0: part[b][a] = sum_{i in block b} w_s(i) * G[i][a]         (means == nullptr)
// pass 1: part[b][a] = sum w_s(i) * (G[i][a] - mean[a])^2
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void k_zstat(int N, int F, int nFeat, const real *__restrict__ feat,
                        const int *__restrict__ structOf, const double *__restrict__ dsw,
                        const double *__restrict__ means, int atomsPerBlock, double *__restrict__ part) {
  int a0 = blockIdx.x * atomsPerBlock;
  int a1 = min(N, a0 + atomsPerBlock);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double s = 0.0;
    const double mu = means ? means[f] : 0.0;
    for (int i = a0; i < a1; i++) {
      const double wgt = dsw[structOf[i]];
      const double v = (double)feat[(size_t)nFeat * i + f];
      s += means ? wgt * (v - mu) * (v - mu) : wgt * v;
    }
    part[(size_t)blockIdx.x * F + f] = s;
  }
}

// out[f] = sum_b part[b][f]  (sequential over blocks: fixed order)
__global__ void k_zstat_final(int nBlocks, int F, const double *__restrict__ part, double *__restrict__ out) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double s = 0.0;
  for (int b = 0; b < nBlocks; b++) s += part[(size_t)b * F + f];
  out[f] = s;
}

template <typename real>
__global__ void k_zapply(size_t N, int F, int nFeat, real *__restrict__ feat, const double *__restrict__ zprec) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * (size_t)F) return;
  size_t i = t / F;
  int a = (int)(t % F);
  const double sg = zprec[F + a];
  if (sg < 1e-08) return;
  real *p = feat + (size_t)nFeat * i + a;
  *p = (real)(((double)*p - zprec[a]) / sg);
}
