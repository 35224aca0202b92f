// acsf.cuh -- ACSF values: one CTA per cell-list bin, one warp per central atom.
//
// Replaces TAcsf_calculate -> iGeoAcsf -> buildGFunctionNeighborlists + g1..g5
// (lib_descriptors/acsf.F90:540-639, 797-866, 942-1066, 1287-1492).  The reference rebuilds
// the neighbour list for every (atom, function) and evaluates acos->cos, pow, exp per pair
// per function; here
//   * the candidate records of the bin's neighbour cells are staged once per CTA in shared
//     memory (cells.cuh), each warp filters them into its own neighbour list (sorted by
//     species code when species-resolved functions exist);
//   * radial G2 functions on an arithmetic rs-ladder with a common eta (the auto scheme,
//     acsf.F90:320-336) are evaluated by the Gaussian recurrence
//       g_{m+1}/g_m = exp(2 eta drs u) * exp(-eta drs^2 (2m+1)),  u = r - rs_first,
//     i.e. two exp per (neighbour, 8 functions) instead of eight;
//   * each angular pass shares the pair geometry between all functions with equal (type, rc,
//     eta, species pair): (1+lam*cos)^xi is evaluated for a whole arithmetic xi-ladder as
//     b^xi0 * (b^dxi)^m with b^dxi = exp(dxi*log b) (xi0 = 1 needs no exp; consecutive ladder
//     slots continue the running product), unordered pairs are visited once (x2) where the
//     reference walks ordered pairs;
//   * the per-lane partial sums are combined by a transposing butterfly (M values in M-1+log
//     shuffles instead of 5M).
// Parity traps kept: neighbour test d2 <= rc^2 (dynneighlist.F90:294), cos(theta) denominator
// r_j r_k + 1e-13 (acsf.F90:1174), diagonal j==k pairs of identical lists
// (acsf.F90:1420-1431,1480-1487), periodic images of the central atom belong to EVERY species
// list (reduceGeometrySpecies keeps iAtom, acsf.F90:754-760), atom-id prefactors q_i q_j in the
// cutoffs and q_j q_k in G4's third cutoff (acsf.F90:1201,1428-1430).
#pragma once
#include "cells.cuh"
#include "fmath.cuh"

#ifndef FNET_ACSF_MINB
#define FNET_ACSF_MINB 5          // resident CTAs per SM the value kernel is compiled for (NS <= 2)
#endif
static_assert(FNET_LADDER == 8, "ladder_accumulate is written for 8-function ladders");

struct WarpSmem {
  double *dx, *dy, *dz, *r, *rinv, *qv, *fcE;
  int *idx, *seg;
  double *outv;
  double *red;       // [redRows][FNET_RED_STRIDE] scratch of reduce_smem
};
#define FNET_RED_STRIDE 33          // 32 lanes + 1: row r, column c sits in bank (r + c) mod 16

__host__ __device__ inline size_t acsf_warp_smem_bytes(int cap, int F, int redRows) {
  size_t b = (size_t)cap * (7 * sizeof(double) + sizeof(int));
  b += (FNET_MAX_CODES + 4) * sizeof(int);
  b = (b + 7) & ~(size_t)7;
  b += (size_t)((F + 1) & ~1) * sizeof(double);
  b += (size_t)redRows * FNET_RED_STRIDE * sizeof(double);
  return (b + 15) & ~(size_t)15;
}
// CTA prefix: staged candidates + the neighbour-cell tables of stage_candidates (PATH 1) or the
// structure's lattice (PATH 2)
#define FNET_FTAB_BYTES (FNET_TAB_DOUBLES * sizeof(double))     // exp / log tables of fmath.cuh, first thing in the CTA's smem
__host__ __device__ inline size_t acsf_cta_prefix_bytes(int capC, int path = 1) {
  const size_t tail = path == 2 ? sizeof(StructGeom) : (path == 1 ? sizeof(StageTabs) : 0);
  return FNET_FTAB_BYTES + (size_t)capC * sizeof(CRec) + ((tail + 15) & ~(size_t)15);
}

// Where the central atoms and their candidates come from.  PATH 0: cell list, candidates read
// from global memory; PATH 1: cell list, candidates of the bin staged per CTA; PATH 2: whole
// structure staged per CTA, minimum image (cells.cuh).
#define FNET_PATH_DIRECT 0
#define FNET_PATH_STAGED 1
#define FNET_PATH_STRUCT 2
struct GeomArgs {                     // device pointers, passed by value to the kernels
  const int *binStruct; const StructInfo *sinfo; const int *cellStart; const CRec *crec;      // PATH 0 / 1
  const int *offsets; const double *coords; const double *lat; const int *periodic; const int *atnum;   // PATH 2
  int stBase;                         // PATH 2: first structure of this launch (chunked, copy-overlapped launches)
};
struct CtaGeom {                      // per-CTA view produced by acsf_cta_prologue
  const StructInfo *S; BinPos bp; const int *cellStart; const CRec *crec;
  const CRec *cand; int nCand; const StructGeom *sg;
  const double *ftab;                 // exp / log tables (shared memory)
  int a0, a1, first;                  // central atoms [a0, a1): slots of crec (PATH 0/1) or atoms (PATH 2, first = atomBeg)
};

// asynchronous global -> shared copies (LDGSTS): no register staging, nothing waits until fnet_cp_async_wait()
__device__ __forceinline__ void fnet_cp_async8(void *sdst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fnet_cp_async4(void *sdst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fnet_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Splits the CTA's bin (or structure) over blockIdx.y and stages the candidates.  Returns false
// when this CTA has nothing to do (or on overflow, flagged for the host).  ASYNCTAB: the table goes to shared memory
// by cp.async -- the caller waits (fnet_cp_async_wait) and synchronises the CTA before it reads c.ftab.
template <int PATH, bool EXPONLY = false, bool ASYNCTAB = false>
__device__ __forceinline__ bool acsf_cta_prologue(const GeomArgs &G, int nSplit, double rcMax, int capC,
                                                  unsigned char *smem_raw, int *__restrict__ flags, CtaGeom &c,
                                                  unsigned char *&wbase) {
  {   // tables first; the barriers of the staging below (or the explicit one of PATH 0) publish them
    double *ft = (double *)smem_raw;
#pragma unroll 1
    for (int e = threadIdx.x; e < (EXPONLY ? FNET_EXP_TAB_N : FNET_TAB_DOUBLES); e += blockDim.x) {
      if (ASYNCTAB) fnet_cp_async8(ft + e, e < FNET_EXP_TAB_N ? &fnet_exp_tab_d[e] : &fnet_log_tab_d[e - FNET_EXP_TAB_N]);
      else ft[e] = e < FNET_EXP_TAB_N ? fnet_exp_tab_d[e] : fnet_log_tab_d[e - FNET_EXP_TAB_N];
    }
    c.ftab = ft;
    smem_raw += EXPONLY ? FNET_EXP_TAB_N * sizeof(double) : FNET_FTAB_BYTES;
  }
  FNET_PDL_TRIGGER();     // socket step (internal.h): the next kernel of the chain may be scheduled now ...
  FNET_PDL_WAIT();        // ... and this one reads geometry / upstream results only after its predecessor has completed
  c.S = nullptr; c.cellStart = G.cellStart; c.crec = G.crec; c.cand = (const CRec *)smem_raw; c.nCand = 0; c.sg = nullptr;
  c.first = 0;
  wbase = smem_raw;
  if (PATH == FNET_PATH_STRUCT) {
    const int st = G.stBase + blockIdx.x;
    const int beg = G.offsets[st], end = G.offsets[st + 1];
    const int per = (end - beg + nSplit - 1) / nSplit;
    c.a0 = beg + blockIdx.y * per; c.a1 = min(end, c.a0 + per); c.first = beg;
    if (c.a0 >= c.a1) return false;
    StructGeom *sg = (StructGeom *)(smem_raw + (size_t)capC * sizeof(CRec));
    c.nCand = stage_structure(st, beg, end - beg, G.coords, G.atnum, G.lat, G.periodic, rcMax, (CRec *)smem_raw, capC, sg, flags);
    c.sg = sg;
    if (c.nCand < 0) { if (threadIdx.x == 0) atomicMax(&flags[7], -c.nCand); return false; }
    wbase += acsf_cta_prefix_bytes(capC, 2) - FNET_FTAB_BYTES;
    return true;
  }
  const int bin = blockIdx.x;
  const int beg = G.cellStart[bin], end = G.cellStart[bin + 1];
  const int per = (end - beg + nSplit - 1) / nSplit;
  c.a0 = beg + blockIdx.y * per; c.a1 = min(end, c.a0 + per);
  if (c.a0 >= c.a1) return false;
  c.S = &G.sinfo[G.binStruct[bin]];
  c.bp = bin_pos(*c.S, bin);
  if (PATH == FNET_PATH_STAGED) {
    StageTabs *tabs = (StageTabs *)(smem_raw + (size_t)capC * sizeof(CRec));
    c.nCand = stage_candidates(*c.S, c.bp, G.cellStart, G.crec, (CRec *)smem_raw, capC, tabs);
    if (c.nCand < 0) { if (threadIdx.x == 0) atomicMax(&flags[7], c.nCand == -1 ? 0x7fffffff : -c.nCand); return false; }
    wbase += acsf_cta_prefix_bytes(capC, 1) - FNET_FTAB_BYTES;
  } else {
    __syncthreads();
  }
  return true;
}
template <int PATH>
__device__ __forceinline__ CRec central_atom(const CtaGeom &c, int slot) {
  return PATH == FNET_PATH_STRUCT ? c.cand[slot - c.first] : c.crec[slot];
}

__device__ __forceinline__ WarpSmem carve_warp_smem(unsigned char *base, int cap, int F, int redRows = 0) {
  WarpSmem w;
  double *d = (double *)base;
  w.dx = d; w.dy = d + cap; w.dz = d + 2 * cap; w.r = d + 3 * cap; w.rinv = d + 4 * cap;
  w.qv = d + 5 * cap; w.fcE = d + 6 * cap;
  w.idx = (int *)(d + 7 * cap);
  w.seg = w.idx + cap;
  size_t off = (size_t)cap * (7 * sizeof(double) + sizeof(int)) + (FNET_MAX_CODES + 4) * sizeof(int);
  off = (off + 7) & ~(size_t)7;
  w.outv = (double *)(base + off);
  w.red = w.outv + ((F + 1) & ~1);
  return w;
}

__device__ __forceinline__ int species_code(const AcsfTables &tab, int z) {
  int code = tab.nCodes;  // "other"
#pragma unroll 1
  for (int c = 0; c < tab.nCodes; c++)
    if (tab.zcodes[c] == z) code = c;
  return code;
}

// Gathers the neighbours of the central atom `me` (within rcMax) into the warp's shared memory,
// sorted by species code [code 0 .. nCodes-1 | other | self-images]; returns n (or -needed on
// overflow).
template <int PATH>
__device__ __forceinline__ int gather_neighbors(const CRec &me, const CtaGeom &cg, const AcsfTables &tab, int cap,
                                                WarpSmem &w) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const bool sorted = tab.nCodes > 0;
  const double rc2 = tab.rcMax * tab.rcMax;
  // unsorted target = final arrays; sorted target = scratch aliased on (fcE, qv, r | rinv)
  double *gx = sorted ? w.fcE : w.dx, *gy = sorted ? w.qv : w.dy, *gz = sorted ? w.r : w.dz;
  int *gj = sorted ? (int *)w.rinv : w.idx;
  int *gc = (int *)w.rinv + cap;          // species codes (sorted path only)
  int n = 0;
  auto take = [&](bool valid, double dx, double dy, double dz, int j, int zs) {
    const bool ok = is_neighbor(valid, dx * dx + dy * dy + dz * dz, rc2, j, zs, me.idx);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int pos = n + __popc(m & lt);
    if (ok && pos < cap) {
      gx[pos] = dx; gy[pos] = dy; gz[pos] = dz; gj[pos] = j;
      if (sorted) gc[pos] = (j == me.idx) ? tab.nCodes + 1 : species_code(tab, zs & ~FNET_SHIFT_FLAG);
    }
    n += __popc(m);
  };
  if (PATH == FNET_PATH_STRUCT) for_each_candidate_struct(cg.cand, cg.nCand, cg.sg, me, take);
  else if (PATH == FNET_PATH_STAGED) for_each_candidate_staged(cg.cand, cg.nCand, me, take);
  else for_each_candidate_direct(*cg.S, cg.bp, cg.cellStart, cg.crec, me, take);
  if (n > cap) return -n;
  __syncwarp();
  if (sorted) {
    const int nc = tab.nCodes + 2;  // + other + self
    int mycount = 0;
    for (int base = 0; base < n; base += 32) {
      const int t = base + lane;
      const int code = (t < n) ? gc[t] : -1;
      for (int c = 0; c < nc; c++) {
        const unsigned m = __ballot_sync(0xffffffffu, code == c);
        if (lane == c) mycount += __popc(m);
      }
    }
    int incl = mycount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    int mybase = incl - mycount;
    if (lane <= nc) w.seg[lane] = (lane < nc) ? mybase : n;
    for (int base = 0; base < n; base += 32) {
      const int t = base + lane;
      int code = -1, j = -1;
      double x = 0, y = 0, z = 0;
      if (t < n) { j = gj[t]; x = gx[t]; y = gy[t]; z = gz[t]; code = gc[t]; }
      int pos = -1;
      for (int c = 0; c < nc; c++) {
        const unsigned m = __ballot_sync(0xffffffffu, code == c);
        const int b = __shfl_sync(0xffffffffu, mybase, c);
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
    const double d2 = w.dx[t] * w.dx[t] + w.dy[t] * w.dy[t] + w.dz[t] * w.dz[t];
    const double rr = sqrt(d2);                      // dynneighlist.F90:311
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

// fc = 0.5 q (cos(pi r / rc) + 1) (acsf.F90:1201) for 0 <= r <= rc, evaluated as q cos^2(pi r / (2 rc))
// with the Taylor polynomial of cos on [0, pi/2] (12 terms, truncation 2e-17; absolute error ~2e-16,
// what the reference's own cos + 1 cancellation has near rc) -- 14 FP64 operations instead of cospi's ~55
static __constant__ double fnet_cosh_cd[12] = {
  1.0, -1.2337005501361697, 0.25366950790104803, -0.02086348076335296, 0.0009192602748394266,
  -2.5202042373060607e-05, 4.710874778818172e-07, -6.386603083791852e-09, 6.565963114979473e-11,
  -5.294400200734623e-13, 3.437739179098607e-15, -1.8359916521552453e-17 };   // (-1)^k (pi/2)^(2k) / (2k)!
__device__ __forceinline__ double cutoff_fn(double rr, double qq, double invrc) {
  const double x = rr * invrc, u = x * x;
  double c = fnet_cosh_cd[11];
#pragma unroll
  for (int k = 10; k >= 0; k--) c = fma(c, u, fnet_cosh_cd[k]);
  return qq * (c * c);
}

// (1 + lam*c)^xi ladder start and ratio from b = max(1 + lam*c, 0) and L = log(b), with the
// pow(0,0)=1 / pow(0,x>0)=0 conventions (log(0) = -inf, exp(-inf) = 0)
static __device__ __noinline__ double fnet_exp_call(double x) { return fnet_exp(x); }   // keeps rare paths out of line
__device__ __forceinline__ void ladder_init(double b, double L, double xi0, double dxi, double &p, double &q,
                                            const double *__restrict__ ftab) {
  if (xi0 == 1.0) p = b;                       // auto scheme: every ladder starts at xi = 1 (acsf.F90:341)
  else if (xi0 == 0.0) p = 1.0;
  else p = fnet_exp_call(xi0 * L);
  q = (dxi == 0.0) ? 1.0 : fnet_exp_tab(dxi * L, ftab);
}
// acc[m] += pw q^m (m = 0..7): powers of q by doubling (dependency depth 3), then one FMA per
// accumulator; returns pw q^8 for a continuing ladder slot
__device__ __forceinline__ double ladder_accumulate(double *acc, double pw, double q) {
  const double q2 = q * q, q3 = q2 * q, q4 = q2 * q2;
  const double q5 = q4 * q, q6 = q4 * q2, q7 = q4 * q3;
  acc[0] += pw;
  acc[1] = fma(pw, q, acc[1]); acc[2] = fma(pw, q2, acc[2]); acc[3] = fma(pw, q3, acc[3]);
  acc[4] = fma(pw, q4, acc[4]); acc[5] = fma(pw, q5, acc[5]); acc[6] = fma(pw, q6, acc[6]);
  acc[7] = fma(pw, q7, acc[7]);
  return (pw * q4) * q4;
}

// pair index -> (row j, column k) of the flattened pair walk.  same: the upper triangle incl. the
// diagonal of an n1 x n1 matrix folded into a rectangle of width W = n1 | 1 -- row r of the
// triangle (n1 - r entries, k = r..n1-1) shares rectangle row r with triangle row n1 - r (n1 odd)
// or n1 - 1 - r (n1 even); else the full n1 x n2 rectangle (W = n2).  invW = 1 / W.
__device__ __forceinline__ void pair_decode(int p, int same, int n1, int W, float invW, int &j, int &k) {
  int r = (int)(((float)p + 0.5f) * invW);
  if (r * W > p) r--;
  else if ((r + 1) * W <= p) r++;
  const int c = p - r * W;
  if (same) {
    const int h = n1 - r;                       // entries of triangle row r
    const bool first = c < h;
    j = first ? r : h - ((n1 & 1) ^ 1);
    k = first ? r + c : j + (c - h);
  } else {
    j = r; k = c;
  }
}

// Transposing butterfly: every lane holds M partial values; on return the lane holds the sum
// over all lanes that differ from it in the lane bits {stride, 2*stride, ..., 16} of the value
// with index (lane / stride) % M.  M-1 + log2(32/(M*stride)) shuffles.
template <int M>
__device__ __forceinline__ double reduce_transpose(double (&v)[M], int lane, int stride) {
#pragma unroll
  for (int h = M / 2; h >= 1; h >>= 1) {
    const int off = h * stride;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < h; k++) {
      const double send = up ? v[k] : v[k + h];
      const double keep = up ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  double r = v[0];
  for (int off = M * stride; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// Same contract as reduce_transpose, through shared memory: every lane stores its M partial values
// as column (lane / stride) of rows (lane % stride) * M + f, then lane L sums a slice of row
// (L % stride) * M + (L / stride) % M and the slices are combined by xor shuffles.  M stores + 32 M /
// (stride M) ... = 256 / 32 loads and adds per lane for the radial shape, 16 for a 16-value angular
// pass -- about 40 % of the butterfly's instruction count (no 64-bit selects, two shuffles).
template <int M>
__device__ __forceinline__ double reduce_smem(const double (&v)[M], int lane, int stride, double *__restrict__ red) {
  const int lgs = __ffs(stride) - 1;            // stride is a power of two: shifts instead of integer divisions
  const int grp = lane & (stride - 1), sub = lane >> lgs;
  const int rows = stride * M;                  // <= 32 here (M = 8: stride <= 4; M = 8 NS: stride = 1)
  __syncwarp();
#pragma unroll
  for (int f = 0; f < M; f++) red[(grp * M + f) * FNET_RED_STRIDE + sub] = v[f];
  __syncwarp();
  // 32 / stride columns are in use; 32 / rows lanes share a row, each sums (32 / stride) / (32 / rows) = M columns
  const int row = grp * M + sub % M;
  const int part = sub / M;                     // = lane / rows
  const double *rp = red + row * FNET_RED_STRIDE + part * M;
  double r = 0.0;
#pragma unroll
  for (int c = 0; c < M; c++) r += rp[c];
  for (int off = rows; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// ------------------------------------------------------------------------------------------
// radial groups (acsf.F90:1287-1373)
//   ladder groups (G2, arithmetic rs, one eta): lanes = (neighbour sub-lane, 8-function chunk),
//     Gaussian recurrence, transposing butterfly over the sub-lanes;
//   generic groups (G1, G3, arbitrary G2): lanes = (neighbour sub-lane, function), one
//     accumulator per lane, xor-shuffle over the sub-lanes.  Kept small on purpose: this path
//     only has to be correct, the instruction cache belongs to the hot loops.
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ double radial_term_generic(int type, double p1, double p2, double rr) {
  if (type == FNETGPU_G1) return 1.0;
  if (type == FNETGPU_G2) { const double d = rr - p2; return exp(-p1 * d * d); }
  return cos(p1 * rr);
}

__device__ __forceinline__ void radial_groups(int i, int n, const AcsfTables &tab, const WarpSmem &w, int nExt,
                                              const double *__restrict__ ext, const double *__restrict__ ftab) {
  const int lane = threadIdx.x & 31;
  for (int g = 0; g < tab.nRadialGroups; g++) {
    const RadialGroup *__restrict__ G = &tab.rgroups[g];
    const int type = G->type, atomId = G->atomId, fBeg = G->fBeg, fCnt = G->fCnt, nch = G->nChunksP2;
    const double rc = G->rc;
    const NbList l = make_list(tab, w, G->code, n);
    const int nl = l.n0 + l.n1;
    const double qi = atomId > 0 ? ext[(size_t)nExt * i + atomId - 1] : 1.0;
    const double invrc = 1.0 / rc;
    if (G->ladder) {
      const int lgn = __ffs(nch) - 1;           // nch = 1, 2 or 4
      const int per = 32 >> lgn;
      const int mychunk = lane & (nch - 1), sub = lane >> lgn;
      const int fbase = fBeg + mychunk * FNET_RCHUNK;
      const int fcnt = min(max(fCnt - mychunk * FNET_RCHUNK, 0), FNET_RCHUNK);
      double acc[FNET_RCHUNK];
#pragma unroll
      for (int f = 0; f < FNET_RCHUNK; f++) acc[f] = 0.0;
      const double eta = G->eta, drs = G->drs;
      const double rsf = G->rs0 + (double)(mychunk * FNET_RCHUNK) * drs;
      double kk[FNET_RCHUNK - 1];
#pragma unroll
      for (int m = 0; m < FNET_RCHUNK - 1; m++) kk[m] = G->kk[m];
      if (fcnt > 0)
        for (int t = sub; t < nl; t += per) {
          const int a = list_at(l, t);
          const double rr = w.r[a];
          if (rr > rc) continue;                       // cutoff1d: rr > rcut -> 0
          const double qj = atomId > 0 ? ext[(size_t)nExt * w.idx[a] + atomId - 1] : 1.0;
          const double fc = cutoff_fn(rr, qi * qj, invrc);
          const double u = rr - rsf;
          const double e0 = eta * u * u, a1 = 2.0 * eta * drs * u;
          if (e0 < 690.0 && fabs(a1) < 690.0) {         // g_0 and the ratio stay normal numbers
            double gv = fnet_exp_tab(-e0, ftab) * fc;
            const double A = fnet_exp_tab(a1, ftab);
            acc[0] += gv;
#pragma unroll
            for (int m = 0; m < FNET_RCHUNK - 1; m++) { gv *= A * kk[m]; acc[m + 1] += gv; }
          } else {                                      // out of the recurrence's safe range (rare)
#pragma unroll 1
            for (int m = 0; m < FNET_RCHUNK; m++) {
              const double d = u - (double)m * drs;
              const double v = radial_term_generic(FNETGPU_G2, eta, 0.0, d) * fc;
#pragma unroll
              for (int f = 0; f < FNET_RCHUNK; f++) acc[f] += (f == m) ? v : 0.0;
            }
          }
        }
      // sum over the sub-lanes: lane ends with function (lane / nch) % 8 of chunk lane % nch
      const double v = reduce_smem<FNET_RCHUNK>(acc, lane, nch, w.red);
      const int f = (lane >> lgn) % FNET_RCHUNK;
      if (lane < nch * FNET_RCHUNK && f < fcnt) w.outv[tab.rfeat[fbase + f]] = v;
    } else {
      int nfP2 = 1;
      while (nfP2 < fCnt) nfP2 <<= 1;                  // fCnt <= 32
      const int per = 32 / nfP2;
      const int f = lane % nfP2, sub = lane / nfP2;
      double acc = 0.0;
      if (f < fCnt) {
        const double p1 = tab.rp1[fBeg + f], p2 = tab.rp2[fBeg + f];
#pragma unroll 1
        for (int t = sub; t < nl; t += per) {
          const int a = list_at(l, t);
          const double rr = w.r[a];
          if (rr > rc) continue;
          const double qj = atomId > 0 ? ext[(size_t)nExt * w.idx[a] + atomId - 1] : 1.0;
          acc += radial_term_generic(type, p1, p2, rr) * cutoff_fn(rr, qi * qj, invrc);
        }
      }
      for (int off = nfP2; off < 32; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane < nfP2 && f < fCnt) w.outv[tab.rfeat[fBeg + f]] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// one angular pass (acsf.F90:1377-1492) with up to NS ladder slots
// ------------------------------------------------------------------------------------------
template <int NS, bool FAST>
__device__ __forceinline__ void angular_pass(int i, int n, const AcsfTables &tab, const AngularPass *__restrict__ P,
                                             const WarpSmem &w, int nExt, const double *__restrict__ ext,
                                             const double *__restrict__ ftab) {
  const int lane = threadIdx.x & 31;
  const int type = P->type, same = P->same, atomId = P->atomId;
  const int nSlots = min(P->nSlots, NS);
  const double rc = P->rc, eta = P->eta, invrc = 1.0 / rc;
  const NbList l1 = make_list(tab, w, P->code1, n);
  const NbList l2 = same ? l1 : make_list(tab, w, P->code2, n);
  const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
  const double qi = atomId > 0 ? ext[(size_t)nExt * i + atomId - 1] : 1.0;
  __syncwarp();
  if (!P->keepFc)                        // species-resolved passes share (rc, eta): computed by the first of them
    for (int t = lane; t < n; t += 32) {   // per-neighbour radial factor fc * exp(-eta r^2)
      const double rr = w.r[t];
      const double qj = atomId > 0 ? ext[(size_t)nExt * w.idx[t] + atomId - 1] : 1.0;
      w.qv[t] = qj;
      w.fcE[t] = (rr > rc) ? 0.0 : cutoff_fn(rr, qi * qj, invrc) * fnet_exp_tab(-eta * rr * rr, ftab);
    }
  __syncwarp();
  double lam[NS], xi0[NS], dxi[NS];
  bool on[NS], cont[NS];
  double acc[NS * FNET_LADDER];
#pragma unroll
  for (int s = 0; s < NS; s++) {
    on[s] = s < nSlots;
    lam[s] = on[s] ? P->slot[s].lam : 0.0;
    xi0[s] = on[s] ? P->slot[s].xi0 : 0.0;
    dxi[s] = on[s] ? P->slot[s].dxi : 0.0;
    cont[s] = on[s] && s > 0 && P->slot[s].cont != 0;
#pragma unroll
    for (int f = 0; f < FNET_LADDER; f++) acc[s * FNET_LADDER + f] = 0.0;
  }
  const int nPairs = same ? (n1 * (n1 + 1)) >> 1 : n1 * n2;
  const int W = same ? (n1 | 1) : n2;
  const float invW = W > 0 ? 1.0f / (float)W : 0.0f;
  for (int p = lane; p < nPairs; p += 32) {
    int j, k;
    pair_decode(p, same, n1, W, invW, j, k);
    const int a = list_at(l1, j), b = list_at(l2, k);
    double base = w.fcE[a] * w.fcE[b];
    if (same && a != b) base *= 2.0;
    if (!FAST && type == FNETGPU_G4 && base != 0.0) {
      const double ex = w.dx[a] - w.dx[b], ey = w.dy[a] - w.dy[b], ez = w.dz[a] - w.dz[b];
      const double djk2 = ex * ex + ey * ey + ez * ez;
      const double djk = sqrt(djk2);
      base = (djk > rc) ? 0.0 : base * fnet_exp_tab(-eta * djk2, ftab) * cutoff_fn(djk, w.qv[a] * w.qv[b], invrc);
    }
    if (FAST) {
      // every slot: a fresh ladder from xi = 1 (p = b) with its own lam -- no branches, no joins
      const double dot = w.dx[a] * w.dx[b] + w.dy[a] * w.dy[b] + w.dz[a] * w.dz[b];
      const double pr = w.rinv[a] * w.rinv[b];
      const double c = dot * (pr * (1.0 - 1e-13 * pr));   // dot / (r_a r_b + 1e-13), acsf.F90:1173-1174
#pragma unroll
      for (int s = 0; s < NS; s++) {
        // |c| <= 1 up to rounding (the 1e-13 of the denominator pulls it inside for atomic distances);
        // a last-bit negative b gives log = -inf, q = 0 and a term of 1e-16 * base: no clamp needed
        const double bb = fma(lam[s], c, 1.0);
        const double q = fnet_exp_tab(dxi[s] * fnet_log_tab(bb, ftab), ftab);
        ladder_accumulate(&acc[s * FNET_LADDER], bb * base, q);
      }
    } else if (base != 0.0) {
      const double dot = w.dx[a] * w.dx[b] + w.dy[a] * w.dy[b] + w.dz[a] * w.dz[b];
      const double pr = w.rinv[a] * w.rinv[b];
      const double c = dot * (pr * (1.0 - 1e-13 * pr));   // dot / (r_a r_b + 1e-13), acsf.F90:1173-1174
      double L = 0.0, bb = 0.0, pw = 0.0, q = 1.0;
#pragma unroll
      for (int s = 0; s < NS; s++) {
        if (on[s]) {
          if (!cont[s]) {
            if (s == 0 || lam[s] != lam[s - 1]) { bb = fmax(1.0 + lam[s] * c, 0.0); L = fnet_log_tab(bb, ftab); }
            ladder_init(bb, L, xi0[s], dxi[s], pw, q, ftab);
            pw *= base;
          }
          pw = ladder_accumulate(&acc[s * FNET_LADDER], pw, q);
        }
      }
    }
  }
  const double v = (NS * FNET_LADDER < 32) ? reduce_smem<NS * FNET_LADDER>(acc, lane, 1, w.red)     // NS = 4: no scratch (redRows)
                                            : reduce_transpose<NS * FNET_LADDER>(acc, lane, 1);
  {
    const int e = lane % (NS * FNET_LADDER);
    const int s = e / FNET_LADDER, f = e % FNET_LADDER;
    if (lane < NS * FNET_LADDER && s < nSlots && f < P->slot[s].count)
      w.outv[P->slot[s].feat[f]] = v * P->slot[s].pref[f];
  }
}

template <typename real, int NS, int PATH>
__global__ void __launch_bounds__(128, (NS <= 2 ? FNET_ACSF_MINB : 3))
k_acsf(int nSplit, GeomArgs geo, int nExt, const double *__restrict__ ext, AcsfTables tab, int cap, int capC,
       real *__restrict__ feat, int nFeat, const double *__restrict__ zprec, int nExtSel,
       const int *__restrict__ extIdx, int *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  CtaGeom cg;
  unsigned char *wbase;
  if (!acsf_cta_prologue<PATH>(geo, nSplit, tab.rcMax, capC, smem_raw, flags, cg, wbase)) return;
  const int a0 = cg.a0, a1 = cg.a1;
  WarpSmem w = carve_warp_smem(wbase + (size_t)wib * acsf_warp_smem_bytes(cap, tab.F, tab.redRows), cap, tab.F, tab.redRows);
  for (int slot = a0 + wib; slot < a1; slot += nw) {
    const CRec me = central_atom<PATH>(cg, slot);
    const int i = me.idx;
    const int n = gather_neighbors<PATH>(me, cg, tab, cap, w);
    if (n < 0) { if (lane == 0) atomicMax(&flags[1], -n); continue; }
    radial_groups(i, n, tab, w, nExt, ext, cg.ftab);
    for (int pi_ = 0; pi_ < tab.nAngularPasses; pi_++) {
      if (tab.apasses[pi_].fast) angular_pass<NS, true>(i, n, tab, &tab.apasses[pi_], w, nExt, ext, cg.ftab);
      else angular_pass<NS, false>(i, n, tab, &tab.apasses[pi_], w, nExt, ext, cg.ftab);
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
    __syncwarp();
  }
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

#ifndef FNET_KERNEL_TU   // non-template kernels: defined once, in fnetgpu.cu (kernels_*.cu set FNET_KERNEL_TU)
// out[f] = sum_b part[b][f]: one CTA per feature, strided partial sums + a fixed-shape tree (deterministic)
__global__ void __launch_bounds__(256) k_zstat_final(int nBlocks, int F, const double *__restrict__ part, double *__restrict__ out) {
  __shared__ double sh[256];
  const int f = blockIdx.x;
  double s = 0.0;
  for (int b = threadIdx.x; b < nBlocks; b += 256) s += part[(size_t)b * F + f];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[f] = sh[0];
}
#endif

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
