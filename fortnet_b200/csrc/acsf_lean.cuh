// acsf_lean.cuh -- ACSF values for the automatic parameter scheme (TGFunctions_fromAutoScheme,
// lib_descriptors/acsf.F90:276-363, optionally species-resolved, initprogram.F90:1453-1527): the
// configuration Fortnet's documentation trains with, and the shape of BASELINE.json's C2 / C3 / C5.
//
// Same mathematics and the same parity traps as k_acsf (acsf.cuh), restructured so that the FP64 pipe,
// not the instruction issue, is the binding resource (B200: one FP64 warp instruction per 2 cycles
// and scheduler; k_acsf spent 70 % of its issue slots on integer / load / branch work):
//   * (1 + lam cos)^xi on the xi-ladder xi_m = 1 + m delta is b * (b^delta)^m.  b^delta is evaluated
//     DIRECTLY by a table-driven power instead of exp(delta * log b):  b = 2^k z, z in [1, 2),
//     z = c_i (1 + r) with c_i the midpoint of the i-th of 256 mantissa intervals (|r| <= 2^-9),
//       b^delta = 2^(k delta) * c_i^delta * (1 + r)^delta,
//     (1 + r)^delta by its binomial series to degree 5 (truncation binom(delta, 6) 2^-54),
//     2^(k delta) and (1/c_i, c_i^delta) from tables built on the host in long double for the
//     configuration's delta: 8 FP64 operations instead of 24 (fnet_log_tab + fnet_exp_tab);
//   * the diagonal (j == k) terms of identical lists (acsf.F90:1420-1431) are closed-form: there
//     cos = 1 - eps with eps = 1e-13 / r^2 (the 1e-13 of acsf.F90:1174), so
//       (1 + lam cos)^xi = (1 + lam)^xi (1 - xi lam eps / (1 + lam))   (+ O(eps^2) ~ 1e-27)
//     and the pair walk covers the strict triangle only -- 120 pairs = 3.75 sweeps for the 16
//     neighbours of bulk Si instead of 136 = 4.25 -> 5 sweeps; the factor 2 of the unordered pairs
//     moves into the per-function prefactor;
//   * the pair index -> (j, k) map of the strict triangle, p = k (k - 1) / 2 + j, does not depend on
//     the neighbour count: one 16-bit table for every atom; consecutive lanes read consecutive j
//     (conflict-free) and mostly the same k (broadcast);
//   * per neighbour: fc and fc exp(-eta r^2) are formed once (rsqrt + one third-order Newton step
//     instead of sqrt and a division) and shared by the radial groups and all angular passes with the
//     same (rc, eta), and the unit vector is stored as u' = u (1 - eps / 2), eps = 1e-13 / r^2: then
//     u'_j . u'_k = cos (1 - (eps_j + eps_k) / 2) reproduces d_j . d_k / (r_j r_k + 1e-13) =
//     cos (1 - sqrt(eps_j eps_k)) up to cos (sqrt(eps_j) - sqrt(eps_k))^2 / 2 < 1e-14 -- the regulariser
//     itself is 1e-13 -- and costs no arithmetic in the pair loop; one 48-byte record per neighbour
//     (u'_x, u'_y | u'_z, fc E | r, fc), two 16-byte loads per pair member; invalid lanes of the last
//     sweep are routed to a dummy neighbour with fc = 0, so the pair loop has no predicate at all;
//   * G = 1, 2 or 4 central atoms per warp (32 / G lanes each), chosen from the neighbour count: for the
//     ~16 neighbours of bulk Si every per-atom step (candidate sweep, per-neighbour factors, radial
//     sums, the cross-lane reductions, table loads) then fills the warp instead of half of it;
//   * the z-score is applied as (g - mu) * (1 / sigma) with the reciprocals formed once per CTA.
// Everything that is not a fresh xi-ladder from xi = 1 with one common delta (explicit function
// lists, G1 / G3 / G4, atom-id scaling, lam < -1) runs through k_acsf.
#pragma once
#include <type_traits>
#include "acsf.cuh"


__host__ __device__ inline size_t lean_group_bytes(int cap, int F, bool sorted, bool f32a = false) {   // cap: multiple of 8
  size_t b = (size_t)cap * 6 * sizeof(double);                         // neighbour records
  if (f32a) b += (size_t)cap * 4 * sizeof(float);                      // FP32 pair arithmetic: (u'_x, u'_y, u'_z, fc E) as float4
  (void)sorted;                                                        // (the sort scratch lives in the reduction area, below)
  b += (size_t)((FNET_MAX_CODES + 4 + 3) & ~3) * sizeof(int);          // list segments
  b += (size_t)((F + 1) & ~1) * sizeof(double);                        // feature row
  return (b + 15) & ~(size_t)15;
}
// species-resolved configurations: unsorted displacements + species codes of a group, alive only between the neighbour
// gather and the counting sort -- they share the warp's reduction scratch, which is first touched after the records
// are final (C3: 63 -> 52 KB per CTA, 3 -> 4 CTAs per SM)
__host__ __device__ inline size_t lean_scratch_bytes(int cap) {
  return ((size_t)cap * (3 * sizeof(double) + sizeof(int)) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t lean_warp_smem_bytes(int cap, int F, int redRows, bool sorted, int G, bool f32a = false) {
  size_t red = ((size_t)redRows * FNET_RED_STRIDE * sizeof(double) + 15) & ~(size_t)15;
  const size_t sc = sorted ? (size_t)G * lean_scratch_bytes(cap) : 0;
  if (sc > red) red = sc;
  return (size_t)G * lean_group_bytes(cap, F, sorted, f32a) + red;
}
__host__ __device__ inline int lean_pair_entries(int cap) {        // staged part of the strict-triangle pair table
  const int need = cap * (cap - 1) / 2 + 1;
  return need < FNET_PAIR_TAB_N ? need : FNET_PAIR_TAB_N;
}
// power tables + (mu, 1/sigma) + atomic number -> species code + pass / radial tables + pair table
__host__ __device__ inline size_t lean_cta_tables_bytes(int F, int cap, int stageBytes) {
  size_t b = (size_t)(FNET_POW_DOUBLES + 2 * ((F + 1) & ~1)) * sizeof(double) + 128;
  b += (size_t)((stageBytes + 15) & ~15);
  b += ((size_t)lean_pair_entries(cap) * sizeof(unsigned short) + 15) & ~(size_t)15;
  return (b + 15) & ~(size_t)15;
}
// what the CTA prefix grows by against k_acsf's: the tables above minus the log table, which the lean
// kernel does not stage (acsf_cta_prologue<PATH, true>); may be "negative" (unsigned wrap-around is intended)
__host__ __device__ inline size_t lean_cta_extra_bytes(int F, int cap, int stageBytes) {
  return lean_cta_tables_bytes(F, cap, stageBytes) - 2 * FNET_LOG_TAB_N * sizeof(double);
}

// b^delta for b in [0, 2] (header comment); b <= 2^-62 (and a last-bit negative b) returns a finite
// value of the size of 2^(-62 delta): the caller multiplies it by b
__device__ __forceinline__ double lean_pow(double b, const double *__restrict__ pt, const LeanTables &lt) {
  const int hi = __double2hiint(b), lo = __double2loint(b);
  const int i = (hi >> 12) & (FNET_POW_TAB_N - 1);
  const int k = max((hi >> 20) - 1023, FNET_POW_KMIN);
  const double z = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 e = *(const double2 *)(pt + 2 * i);
  const double r = fma(z, e.x, -1.0);
  double s = fma(lt.powC[4], r, lt.powC[3]);
  s = fma(s, r, lt.powC[2]);
  s = fma(s, r, lt.powC[1]);
  s = fma(s, r, lt.powC[0]);
  const double t = s * r;
  const double v = fma(t, e.y, e.y);
  return v * pt[2 * FNET_POW_TAB_N + (k - FNET_POW_KMIN)];
}

// 1 / sqrt(d2) and sqrt(d2) for d2 > 0: hardware seed (2^-22) + one third-order step, then one
// residual correction of the root -- 9 FP64 operations, within an ulp of sqrt() and 1.0 / sqrt()
__device__ __forceinline__ void lean_rsqrt(double d2, double &rinv, double &r) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d2));
  const double t = d2 * y;
  const double e = fma(-t, y, 1.0);
  const double h = fma(0.375, e, 0.5);
  y = fma(y * e, h, y);
  double rr = d2 * y;
  rr = fma(fma(-rr, rr, d2), 0.5 * y, rr);
  rinv = y; r = rr;
}

// species list = segment(code) followed by the self-image segment (the last one, ending at n)
__device__ __forceinline__ NbList lean_list(const AcsfTables &tab, const int *__restrict__ seg, int code, int n) {
  NbList l;
  if (code < 0) { l.s0 = 0; l.n0 = n; l.s1 = n; l.n1 = 0; }
  else {
    const int self = tab.nCodes + 1;
    l.s0 = seg[code]; l.n0 = seg[code + 1] - l.s0;
    l.s1 = seg[self]; l.n1 = seg[self + 1] - l.s1;
  }
  return l;
}

// per-neighbour factor fc(r) exp(-eta r^2) (0 beyond rc) for one (rc, eta)
__device__ __forceinline__ double lean_fce(double rr, double rc, double invrc, double eta, const double *__restrict__ ftab,
                                           double &fc) {
  fc = (rr > rc) ? 0.0 : cutoff_fn(rr, 1.0, invrc);
  return fc * fnet_exp_tab(-eta * rr * rr, ftab);
}

// acc[m] += pw q^m (m = 0..7); q2, q4 are shared with the chained slots of the same ladder
__device__ __forceinline__ void lean_ladder8(double *acc, double pw, double q, double q2, double q4) {
  const double q3 = q2 * q, q5 = q4 * q, q6 = q4 * q2, q7 = q4 * q3;
  acc[0] += pw;
  acc[1] = fma(pw, q, acc[1]); acc[2] = fma(pw, q2, acc[2]); acc[3] = fma(pw, q3, acc[3]);
  acc[4] = fma(pw, q4, acc[4]); acc[5] = fma(pw, q5, acc[5]); acc[6] = fma(pw, q6, acc[6]);
  acc[7] = fma(pw, q7, acc[7]);
}

// strict-triangle pair p = k (k - 1) / 2 + j  ->  (j, k), j < k
__device__ __forceinline__ void lean_tri_decode(int p, int &j, int &k) {
  int kk = (int)(sqrtf(2.0f * (float)p + 0.25f) + 0.5f);
  int tri = (kk * (kk - 1)) >> 1;
  if (tri > p) { kk--; tri -= kk; }
  else if (tri + kk <= p) { tri += kk; kk++; }
  j = p - tri; k = kk;
}

// Sums of the M per-lane partial values over the LPA lanes of each group through the warp-wide
// scratch red[M][33] (column = lane: conflict-free stores; lane (g, sl) then reads rows sl, sl + LPA, ...
// over its group's columns, bank (row + column) mod 16: conflict-free as well).
//   M >= LPA: lane sl returns the totals of rows sl + i LPA in out[i], i < M / LPA
//   M <  LPA: LPA / M lanes share row sl % M (M columns each, combined by xor shuffles), total in out[0]
template <int M, int LPA>
__device__ __forceinline__ void lean_group_reduce(const double (&v)[M], int lane, double *__restrict__ red,
                                                  double (&out)[(M >= LPA ? M / LPA : 1)]) {
  const int sl = lane & (LPA - 1), gcol = lane & ~(LPA - 1);
  __syncwarp();
#pragma unroll
  for (int f = 0; f < M; f++) red[f * FNET_RED_STRIDE + lane] = v[f];
  __syncwarp();
  if (M >= LPA) {
#pragma unroll
    for (int i = 0; i < (M >= LPA ? M / LPA : 1); i++) {
      const double *rp = red + (sl + i * LPA) * FNET_RED_STRIDE + gcol;
      double r = 0.0;
#pragma unroll
      for (int c = 0; c < LPA; c++) r += rp[c];
      out[i] = r;
    }
  } else {
    const int row = sl & (M - 1), part = sl / M;
    const double *rp = red + row * FNET_RED_STRIDE + gcol + part * M;
    double r = 0.0;
#pragma unroll
    for (int c = 0; c < M; c++) r += rp[c];
#pragma unroll
    for (int off = M; off < LPA; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    out[0] = r;
  }
}

// One pair of the angular pass: records a, b -> the NL x NC ladders
template <int NL, int NC>
__device__ __forceinline__ void lean_pair_eval(double (&acc)[NL * NC * FNET_LADDER], const double *__restrict__ rec,
                                               const float4 *__restrict__, int a, int b,
                                               int m0, const double (&lam)[NL], const double *__restrict__ pt,
                                               const LeanTables &lt) {
  const double2 a0 = *(const double2 *)(rec + 6 * a), a1 = *(const double2 *)(rec + 6 * a + 2);
  const double2 b0 = *(const double2 *)(rec + 6 * b), b1 = *(const double2 *)(rec + 6 * b + 2);
  const double base = a1.y * b1.y;
  double c = a0.x * b0.x;                      // d_a . d_b / (r_a r_b + 1e-13), acsf.F90:1173-1174 (header comment)
  c = fma(a0.y, b0.y, c);
  c = fma(a1.x, b1.x, c);
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const double bb = fma(lam[l], c, 1.0);
    const double q = lean_pow(bb, pt, lt);
    double pw = bb * base;
    const double q2 = q * q, q4 = q2 * q2;
    if (m0 > 0) {                              // later 8 NC-function blocks of a long ladder: b q^m0, m0 a multiple of 8
      double qm = 1.0, qb = q4 * q4;
      for (int t = m0 >> 3; t; t >>= 1) { if (t & 1) qm *= qb; qb *= qb; }
      pw *= qm;
    }
#pragma unroll
    for (int ch = 0; ch < NC; ch++) {
      lean_ladder8(&acc[(l * NC + ch) * FNET_LADDER], pw, q, q2, q4);
      if (ch + 1 < NC) pw = (pw * q4) * q4;
    }
  }
}
// precision 32: the same pair in FP32 -- one 16-byte record per member (u'_x, u'_y, u'_z, fc E), b^delta =
// ex2(delta lg2 b) on the MUFU pipe (ex2.approx / lg2.approx, ~2 ulp each), FFMA ladders, FP32 partial sums
// per lane (summed over the lanes in FP64)
template <int NL, int NC>
__device__ __forceinline__ void lean_pair_eval(float (&acc)[NL * NC * FNET_LADDER], const double *__restrict__,
                                               const float4 *__restrict__ recf, int a, int b,
                                               int m0, const double (&lam)[NL], const double *__restrict__,
                                               const LeanTables &lt) {
  const float4 A = recf[a], B = recf[b];
  const float base = A.w * B.w;
  float c = A.x * B.x;
  c = fmaf(A.y, B.y, c);
  c = fmaf(A.z, B.z, c);
  const float delta = (float)lt.powC[0];
#pragma unroll
  for (int l = 0; l < NL; l++) {
    const float bb = fmaxf(fmaf((float)lam[l], c, 1.0f), 1e-30f);
    const float q = exp2f(delta * __log2f(bb));
    float pw = bb * base;
    const float q2 = q * q, q4 = q2 * q2;
    if (m0 > 0) {
      float qm = 1.0f, qb = q4 * q4;
      for (int t = m0 >> 3; t; t >>= 1) { if (t & 1) qm *= qb; qb *= qb; }
      pw *= qm;
    }
#pragma unroll
    for (int ch = 0; ch < NC; ch++) {
      float *ac = &acc[(l * NC + ch) * FNET_LADDER];
      const float q3 = q2 * q, q5 = q4 * q, q6 = q4 * q2, q7 = q4 * q3;
      ac[0] += pw;
      ac[1] = fmaf(pw, q, ac[1]); ac[2] = fmaf(pw, q2, ac[2]); ac[3] = fmaf(pw, q3, ac[3]);
      ac[4] = fmaf(pw, q4, ac[4]); ac[5] = fmaf(pw, q5, ac[5]); ac[6] = fmaf(pw, q6, ac[6]);
      ac[7] = fmaf(pw, q7, ac[7]);
      if (ch + 1 < NC) pw = (pw * q4) * q4;
    }
  }
}

// pair index -> list positions.  KIND 0: identical lists, strict triangle p = k (k-1)/2 + j through the pair table;
// 1: identical lists longer than the table; 2: two lists (p = j n2 + k).  KIND 1 / 2 decode the first index of a lane
// once and then ADVANCE (j, k) by the loop stride -- no division, no square root in the loop.
__device__ __forceinline__ void lean_rect_decode(int p, int n2, int &j, int &k) {
  const float invW = n2 > 0 ? 1.0f / (float)n2 : 0.0f;
  j = (int)(((float)p + 0.5f) * invW);
  if (j * n2 > p) j--;
  else if ((j + 1) * n2 <= p) j++;
  k = p - j * n2;
}

#ifndef FNET_LEAN_UNROLL
#define FNET_LEAN_UNROLL 2
#endif
// species-resolved configurations: ONE pair per lane and iteration -- they execute two instantiations of the pair loop
// (identical lists, two lists) plus the counting sort, and with two pairs in flight their executed code no longer fits
// the instruction caches (L0 ~6 KB, L1.5 32 KB): C3 ACSF 17.3 -> 14.0 ms in a same-box A/B of deterministic builds
#ifndef FNET_LEAN_UNROLL_SORTED
#define FNET_LEAN_UNROLL_SORTED 1
#endif
// neighbours per lane in flight in the record and radial loops of single-list configurations.  Measured: 2 (two
// independent rsqrt / cutoff / exp chains per lane) LOSES -- C2 ACSF 0.722 -> 0.860 ms in a same-box A/B: at the
// 128-register cap the second chain is paid with rematerialisation, and the executed code grows past the caches.
#ifndef FNET_LEAN_NB_UNROLL
#define FNET_LEAN_NB_UNROLL 1
#endif
#ifndef FNET_LEAN_PAIR_PRAGMA
#define FNET_LEAN_PAIR_PRAGMA
#endif
// One angular pass over the pairs of (l1, l2); indices beyond the last pair map to (0, n1) resp. (n1, 0): the dummy
// neighbour.  Two pairs per lane and iteration for the 16-accumulator shapes of SINGLE-LIST configurations: two
// independent pairs hide the latency of the dependent table-lookup / DFMA chains at 16 resident warps per SM (C2,
// deterministic builds: 0.749 -> 0.725 ms; 5 CTAs at 96 registers spill and lose).  Species-resolved configurations
// walk one pair per lane (FNET_LEAN_UNROLL_SORTED above): their code has to fit the instruction caches first.
template <int NL, int NC, int LPA, bool SORTED, int KIND, typename AT>
__device__ __forceinline__ void lean_pair_loop(AT (&acc)[NL * NC * FNET_LADDER], const double *__restrict__ rec,
                                               const float4 *__restrict__ recf,
                                               const NbList &l1, const NbList &l2, int n1, int n2, int m0,
                                               const double (&lam)[NL], const double *__restrict__ pt,
                                               const unsigned short *__restrict__ ptab, const LeanTables &lt, int sl) {
  constexpr int U = NL * NC <= 2 ? (SORTED ? FNET_LEAN_UNROLL_SORTED : FNET_LEAN_UNROLL) : 1;
  constexpr int S = U * LPA;
  const int nP = KIND == 2 ? n1 * n2 : (n1 * (n1 - 1)) >> 1;
  int jj[U], kk[U];
  if (KIND != 0) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (KIND == 1) lean_tri_decode(u * LPA + sl, jj[u], kk[u]);
      else lean_rect_decode(u * LPA + sl, n2, jj[u], kk[u]);
    }
  }
  FNET_LEAN_PAIR_PRAGMA
  for (int p0 = 0; p0 < nP; p0 += S) {
    int a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int j, k;
      if (KIND == 0) { const unsigned jk = ptab[min(p0 + u * LPA + sl, nP)]; j = jk & 255; k = jk >> 8; }
      else if (KIND == 1) {
        j = jj[u]; k = kk[u];
        if (k >= n1) { j = 0; k = n1; }
        jj[u] += S;
        while (jj[u] >= kk[u]) { jj[u] -= kk[u]; kk[u]++; }
      } else {
        j = jj[u]; k = kk[u];
        if (j >= n1) { j = n1; k = 0; }
        kk[u] += S;
        while (kk[u] >= n2) { kk[u] -= n2; jj[u]++; }
      }
      a[u] = SORTED ? list_at(l1, j) : j; b[u] = SORTED ? list_at(l2, k) : k;
    }
#pragma unroll
    for (int u = 0; u < U; u++) lean_pair_eval<NL, NC>(acc, rec, recf, a[u], b[u], m0, lam, pt, lt);
  }
}

// Neighbour gather over a LINEAR candidate array (whole structure: minimum image; staged bin candidates) for the G
// central atoms of a warp at once: every lane loads ONE candidate per sweep and tests it against all G central atoms
// (their coordinates live in registers), so a sweep costs 2 shared-memory loads + G distance tests for 32 candidates --
// G times fewer sweeps than one candidate array walk per group.  Accepted candidates are compacted (ballot) in candidate
// order into group q's arrays: displacements to rec[RS * pos + 0..2] (or the unsorted scratch when SORTED, with the
// species code), atom indices when WITHIDX.  Returns this lane's group's neighbour count.
template <int PATH, bool SORTED, int G, int RS, bool WITHIDX>
__device__ __forceinline__ int lean_gather_linear(const CtaGeom &cg, const AcsfTables &tab, const CRec &me, bool act, int cap,
                                                  double *rec0, double *gx0, double *gy0, double *gz0, int *gc0, int *gi0,
                                                  size_t gbytes, size_t sbytes, const unsigned char *__restrict__ zcode, int lane) {
  constexpr int LPA = 32 / G;
  const int grp = lane / LPA;
  const unsigned ltm = (1u << lane) - 1u;
  const double rc2 = tab.rcMax * tab.rcMax;
  double mx[G], my[G], mz[G];
  int mi[G], nq[G];
  bool aq[G];
#pragma unroll
  for (int q = 0; q < G; q++) {
    mx[q] = __shfl_sync(0xffffffffu, me.x, q * LPA); my[q] = __shfl_sync(0xffffffffu, me.y, q * LPA);
    mz[q] = __shfl_sync(0xffffffffu, me.z, q * LPA); mi[q] = __shfl_sync(0xffffffffu, me.idx, q * LPA);
    aq[q] = __shfl_sync(0xffffffffu, act ? 1 : 0, q * LPA) != 0;
    nq[q] = 0;
  }
  const bool per = PATH == FNET_PATH_STRUCT && cg.sg->periodic != 0;
  const bool diag = PATH == FNET_PATH_STRUCT && cg.sg->diag != 0;
  const StructGeom *__restrict__ sg = cg.sg;
  for (int base = 0; base < cg.nCand; base += 32) {
    const int t = base + lane;
    const bool valid = t < cg.nCand;
    CRec r;
    r.x = 0.0; r.y = 0.0; r.z = 0.0; r.idx = -1; r.zs = 0;
    if (valid) r = cg.cand[t];
    int code = 0;
    if (SORTED) code = (int)zcode[(r.zs & ~FNET_SHIFT_FLAG) & 127];
#pragma unroll
    for (int q = 0; q < G; q++) {
      double dx = r.x - mx[q], dy = r.y - my[q], dz = r.z - mz[q];
      if (PATH == FNET_PATH_STRUCT) {                      // minimum image (cells.cuh for_each_candidate_struct)
        const double magic = 6755399441055744.0;
        if (diag) {
          const double n0 = (sg->inv[0] * dx + magic) - magic;
          const double n1 = (sg->inv[4] * dy + magic) - magic;
          const double n2 = (sg->inv[8] * dz + magic) - magic;
          dx -= n0 * sg->lat[0]; dy -= n1 * sg->lat[4]; dz -= n2 * sg->lat[8];
        } else if (per) {
          const double n0 = (sg->inv[0] * dx + sg->inv[1] * dy + sg->inv[2] * dz + magic) - magic;
          const double n1 = (sg->inv[3] * dx + sg->inv[4] * dy + sg->inv[5] * dz + magic) - magic;
          const double n2 = (sg->inv[6] * dx + sg->inv[7] * dy + sg->inv[8] * dz + magic) - magic;
          dx -= n0 * sg->lat[0] + n1 * sg->lat[3] + n2 * sg->lat[6];
          dy -= n0 * sg->lat[1] + n1 * sg->lat[4] + n2 * sg->lat[7];
          dz -= n0 * sg->lat[2] + n1 * sg->lat[5] + n2 * sg->lat[8];
        }
      }
      const bool ok = is_neighbor(valid && aq[q], dx * dx + dy * dy + dz * dz, rc2, r.idx, r.zs, mi[q]);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int pos = nq[q] + __popc(m & ltm);
      if (ok && pos < cap - 1) {
        const size_t go = (size_t)q * gbytes, so = (size_t)q * sbytes;
        if (SORTED) {
          ((double *)((unsigned char *)gx0 + so))[pos] = dx; ((double *)((unsigned char *)gy0 + so))[pos] = dy;
          ((double *)((unsigned char *)gz0 + so))[pos] = dz;
          ((int *)((unsigned char *)gc0 + so))[pos] = (r.idx == mi[q]) ? tab.nCodes + 1 : code;
          if (WITHIDX) ((int *)((unsigned char *)gi0 + so))[pos] = r.idx;
        } else {
          double *qq = (double *)((unsigned char *)rec0 + go) + (size_t)RS * pos;
          qq[0] = dx; qq[1] = dy; qq[2] = dz;
          if (WITHIDX) ((int *)(qq + RS - 1))[0] = r.idx;
        }
      }
      nq[q] += __popc(m);
    }
  }
  int n = nq[0];
#pragma unroll
  for (int q = 1; q < G; q++) n = (grp == q) ? nq[q] : n;
  return n;
}

#ifndef FNET_LEAN_MINB
#define FNET_LEAN_MINB 4
#endif
// The kernel lives in its OWN translation unit (kernels_lean.cu defines FNET_DEFINE_LEAN_KERNEL and instantiates every
// variant): in one unit with the rest of the library its code generation changed with edits to unrelated kernels -- the
// same source compiled to 5 176 or 5 568 instructions, and the species-resolved C3 launch took 15.3 or 19.3 ms.  The
// other units see a variable template of the same name that holds the kernel's address, so launch sites read alike.
typedef void (*LeanKernelT)(int, GeomArgs, int, const double *, AcsfTables, LeanTables, int, int, void *, int, int,
                            const double *, int, const int *, int *);
LeanKernelT fnet_lean_kernel(int NL, int NC, int PATH, bool SORTED, int G, bool F32A);   // kernels_lean.cu (nullptr: not built)
#ifndef FNET_DEFINE_LEAN_KERNEL
template <int NL, int NC, int PATH, bool SORTED, int G, bool F32A>
static const LeanKernelT k_acsf_lean = fnet_lean_kernel(NL, NC, PATH, SORTED, G, F32A);
#else
template <int NL, int NC, int PATH, bool SORTED, int G, bool F32A>
__global__ void __launch_bounds__(128, (NL * NC <= 2 ? FNET_LEAN_MINB : 3))
k_acsf_lean(int nSplit, GeomArgs geo, int nExt, const double *__restrict__ ext, AcsfTables tab, LeanTables lt, int cap,
            int capC, void *__restrict__ featv, int f32, int nFeat, const double *__restrict__ zprec, int nExtSel,
            const int *__restrict__ extIdx, int *__restrict__ flags) {
  constexpr int M = NL * NC * FNET_LADDER;          // accumulators per lane
  constexpr int LPA = 32 / G;                       // lanes per central atom
  constexpr int RA = M >= LPA ? M / LPA : 1;        // finished angular values per lane after a reduction
  constexpr int RR = FNET_RCHUNK >= LPA ? FNET_RCHUNK / LPA : 1;
  constexpr int NBU = SORTED ? 1 : FNET_LEAN_NB_UNROLL;   // neighbours per lane in flight (record / radial loops)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int grp = lane / LPA, sl = lane & (LPA - 1), gshift = grp * LPA;
  const unsigned lowmask = LPA == 32 ? 0xffffffffu : ((1u << LPA) - 1u);
  const unsigned ltmask = (1u << sl) - 1u;
  CtaGeom cg;
  unsigned char *wbase;
  typedef typename std::conditional<F32A, float, double>::type AT;   // arithmetic type of the pair sums
  // Tables that do not depend on the geometry (power tables, pass / radial tables, pair table) go to shared memory as
  // ASYNCHRONOUS copies (cp.async) issued before anything else: their global-memory latency then overlaps the staging
  // of the structure (offsets -> coordinates, lattice inverse, a barrier) instead of following it -- the CTA prologue was
  // 9 % of the kernel's stall samples even at 64 atoms per CTA, most of it rolled load -> store loops waiting one
  // global-memory round trip per iteration.
  const int F = tab.F, Fp = (F + 1) & ~1;
  unsigned char *wbase0 = smem_raw + FNET_EXP_TAB_N * sizeof(double) +
      (PATH == FNET_PATH_DIRECT ? 0 : acsf_cta_prefix_bytes(capC, PATH == FNET_PATH_STRUCT ? 2 : 1) - FNET_FTAB_BYTES);
  double *pt = (double *)wbase0;                    // power tables
  double *zmu = pt + FNET_POW_DOUBLES, *zis = zmu + Fp;
  unsigned char *zcode = (unsigned char *)(zis + Fp);   // atomic number -> species code (SORTED)
  unsigned char *stage = zcode + 128;               // pass / radial tables (when small): shared-memory latency instead of L1's
  unsigned short *ptab = (unsigned short *)(stage + ((lt.stageBytes + 15) & ~15));   // the used part of the pair table
  const LeanPass *passes = lt.pass;
  const LeanRadial *rads = lt.rad;
#pragma unroll 1
  for (int e = threadIdx.x; e < FNET_POW_DOUBLES; e += blockDim.x) fnet_cp_async8(pt + e, lt.powtab + e);
  if (lt.stageBytes > 0) {
    const int nw8 = lt.stageBytes >> 3, np8 = (int)((lt.nPasses * sizeof(LeanPass)) >> 3);
    double *dst = (double *)stage;
    const double *srcP = (const double *)lt.pass, *srcR = (const double *)lt.rad;
#pragma unroll 1
    for (int e = threadIdx.x; e < nw8; e += blockDim.x) fnet_cp_async8(dst + e, e < np8 ? srcP + e : srcR + (e - np8));
    passes = (const LeanPass *)stage;
    rads = (const LeanRadial *)(stage + (size_t)np8 * 8);
  }
  {
    const int ne2 = (lean_pair_entries(cap) + 1) >> 1;                 // 32-bit copies
    const unsigned *src = (const unsigned *)lt.pairtab;
#pragma unroll 1
    for (int e = threadIdx.x; e < ne2; e += blockDim.x) fnet_cp_async4((unsigned *)ptab + e, src + e);
  }
  double zMu = 0.0, zIs = 1.0;                      // z-score row of this thread (F <= blockDim.x in practice): loads in flight too
  if (zprec && (int)threadIdx.x < F) { zMu = zprec[threadIdx.x]; zIs = zprec[F + threadIdx.x]; }
  if (!acsf_cta_prologue<PATH, true, true>(geo, nSplit, tab.rcMax, capC, smem_raw, flags, cg, wbase)) { fnet_cp_async_wait(); return; }   // exp table only, asynchronously
#pragma unroll 1
  for (int a = threadIdx.x; a < F; a += blockDim.x) {
    double mu = 0.0, is = 1.0;
    if (zprec) {
      const double sg = a < (int)blockDim.x ? zIs : zprec[F + a];
      if (!(sg < 1e-08)) { mu = a < (int)blockDim.x ? zMu : zprec[a]; is = 1.0 / sg; }   // acsf.F90:505-507
    }
    zmu[a] = mu; zis[a] = is;
  }
  if (SORTED)
#pragma unroll 1
    for (int z = threadIdx.x; z < 128; z += blockDim.x) zcode[z] = (unsigned char)species_code(tab, z);
  fnet_cp_async_wait();
  __syncthreads();
  wbase += lean_cta_tables_bytes(F, cap, lt.stageBytes);
  const int a0 = cg.a0, a1 = cg.a1;
  const size_t gbytes = lean_group_bytes(cap, F, SORTED, F32A);
  unsigned char *wb = wbase + (size_t)wib * lean_warp_smem_bytes(cap, F, lt.redRows, SORTED, G, F32A);
  unsigned char *gb = wb + (size_t)grp * gbytes;
  double *rec = (double *)gb;                                        // [cap][6]: u'_x, u'_y | u'_z, fc E | r, fc
  float4 *recf = (float4 *)(rec + 6 * cap);                          // F32A: [cap] (u'_x, u'_y, u'_z, fc E)
  int *seg = (int *)(rec + 6 * cap + (F32A ? 2 * cap : 0));
  double *outv = (double *)(seg + ((FNET_MAX_CODES + 4 + 3) & ~3));
  double *red = (double *)(wb + (size_t)G * gbytes);
  const size_t sbytes = lean_scratch_bytes(cap);                     // SORTED: unsorted displacements and species codes of the
  double *gx = (double *)((unsigned char *)red + (size_t)grp * sbytes), *gy = gx + cap, *gz = gy + cap;   // group, inside the
  int *gc = (int *)(gz + cap);                                       //         reduction scratch (dead until the records are final)
  const double *ftab = cg.ftab;
  int nmaxW = 0;
  for (int s0 = a0 + wib * G; s0 < a1; s0 += nw * G) {
    const int slot = s0 + grp;
    bool act = slot < a1;
    const CRec me = central_atom<PATH>(cg, act ? slot : a0);
    const int i = me.idx;
    // ---------------- neighbours of each group's central atom: candidates -> compacted displacements ----------------
    int n = 0;
    const double rc2 = tab.rcMax * tab.rcMax;
    auto take = [&](bool ok, double dx, double dy, double dz, int j, int zs) {
      const unsigned mg = (__ballot_sync(0xffffffffu, ok) >> gshift) & lowmask;
      const int pos = n + __popc(mg & ltmask);
      if (ok && pos < cap - 1) {
        if (SORTED) {
          gx[pos] = dx; gy[pos] = dy; gz[pos] = dz;
          gc[pos] = (j == i) ? tab.nCodes + 1 : (int)zcode[(zs & ~FNET_SHIFT_FLAG) & 127];
        } else {
          rec[6 * pos] = dx; rec[6 * pos + 1] = dy; rec[6 * pos + 2] = dz;
        }
      }
      n += __popc(mg);
    };
    if (PATH == FNET_PATH_DIRECT) {            // candidates straight from the cell list (bins too crowded to stage): G = 1
      for_each_candidate_direct(*cg.S, cg.bp, cg.cellStart, cg.crec, me, [&](bool valid, double dx, double dy, double dz, int j, int zs) {
        take(is_neighbor(valid && act, dx * dx + dy * dy + dz * dz, rc2, j, zs, i), dx, dy, dz, j, zs);
      });
    } else {
      double *x0 = red;                                    // group 0's records at wb, its scratch at red; group q's are gbytes / sbytes * q further
      n = lean_gather_linear<PATH, SORTED, G, 6, false>(cg, tab, me, act, cap, (double *)wb, x0, x0 + cap, x0 + 2 * cap, (int *)(x0 + 3 * cap),
                                                        nullptr, gbytes, sbytes, zcode, lane);
    }
    if (n > cap - 1) {                                       // one slot is the dummy neighbour
      if (sl == 0) atomicMax(&flags[1], n + 1);
      act = false; n = 0;
    }
    nmaxW = max(nmaxW, n);
    __syncwarp();
    if (SORTED) {   // stable counting sort by species code: [code 0 .. | other | self-images] (acsf.F90:754-760)
      const int nc = tab.nCodes + 2;
      int nAll = n;
#pragma unroll
      for (int o = 16; o >= LPA; o >>= 1) nAll = max(nAll, __shfl_xor_sync(0xffffffffu, nAll, o));
      int mycount = 0;
      for (int base = 0; base < nAll; base += LPA) {
        const int t = base + sl;
        const int code = (t < n) ? gc[t] : -1;
        for (int c = 0; c < nc; c++) {
          const unsigned mg = (__ballot_sync(0xffffffffu, code == c) >> gshift) & lowmask;
          if (sl == c) mycount += __popc(mg);
        }
      }
      int incl = mycount;
#pragma unroll
      for (int o = 1; o < LPA; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o, LPA); if (sl >= o) incl += y; }
      int mybase = incl - mycount;
      if (sl < nc) seg[sl] = mybase;
      if (sl == 0) seg[nc] = n;
      for (int base = 0; base < nAll; base += LPA) {
        const int t = base + sl;
        int code = -1;
        double x = 0, y = 0, z = 0;
        if (t < n) { x = gx[t]; y = gy[t]; z = gz[t]; code = gc[t]; }
        int pos = -1;
        for (int c = 0; c < nc; c++) {
          const unsigned mg = (__ballot_sync(0xffffffffu, code == c) >> gshift) & lowmask;
          const int b = __shfl_sync(0xffffffffu, mybase, c, LPA);
          if (code == c) pos = b + __popc(mg & ltmask);
          if (sl == c) mybase += __popc(mg);
        }
        if (pos >= 0) { rec[6 * pos] = x; rec[6 * pos + 1] = y; rec[6 * pos + 2] = z; }
      }
      __syncwarp();
    }
    // ---------------- per-neighbour record; entry n is the dummy neighbour ----------------
    // single list (!SORTED): the diagonal sums of the passes with the shared (rc, eta) -- S0 = sum (fc E)^2,
    // S1 = sum (fc E)^2 eps -- are formed here, once per atom, instead of in a loop of their own per pass
    double dS0 = 0.0, dS1 = 0.0;
#pragma unroll NBU
    for (int t = sl; t <= n; t += LPA) {
      double ux = 0.0, uy = 0.0, uz = 0.0, fe = 0.0, rr = 2.0 * tab.rcMax, fc = 0.0;
      if (t < n) {
        const double dx = rec[6 * t], dy = rec[6 * t + 1], dz = rec[6 * t + 2];
        double ri;
        lean_rsqrt(dx * dx + dy * dy + dz * dz, ri, rr);   // dynneighlist.F90:311
        const double ri2 = ri * ri;
        const double sc = ri * fma(ri2, -0.5e-13, 1.0);    // u (1 - eps / 2), eps = 1e-13 / r^2
        ux = dx * sc; uy = dy * sc; uz = dz * sc;          // acsf.F90:1565
        fe = lean_fce(rr, lt.rcShared, lt.invrcShared, lt.etaShared, ftab, fc);
        if (!SORTED) { const double e2 = fe * fe; dS0 += e2; dS1 = fma(e2, ri2 * 1e-13, dS1); }
      }
      *(double2 *)(rec + 6 * t) = make_double2(ux, uy);
      *(double2 *)(rec + 6 * t + 2) = make_double2(uz, fe);
      *(double2 *)(rec + 6 * t + 4) = make_double2(rr, fc);
      if (F32A) recf[t] = make_float4((float)ux, (float)uy, (float)uz, (float)fe);
    }
    if (!SORTED) {
#pragma unroll
      for (int o = LPA / 2; o > 0; o >>= 1) {
        dS0 += __shfl_xor_sync(0xffffffffu, dS0, o);
        dS1 += __shfl_xor_sync(0xffffffffu, dS1, o);
      }
    }
    __syncwarp();
    // ---------------- radial ladder groups (acsf.F90:1287-1373): lanes = neighbours, 16 functions per sweep ----------------
    // g_m = fc exp(-eta (u - m drs)^2) = g_{m-1} A k_{m-1}, A = exp(2 eta drs u), k_m = exp(-eta drs^2 (2m+1)): two
    // exponentials per neighbour serve a PAIR of 8-function chunks (functions 8..15 continue the recurrence with
    // A c16 k_{m-8}, c16 = exp(-16 eta drs^2)) -- the 16 G2 of the automatic scheme cost 2 exponentials, not 4.
    for (int g = 0; g < lt.nRadial; g++) {
      const LeanRadial *__restrict__ R = &rads[g];
      const int fCnt = R->fCnt;
      const NbList l = lean_list(tab, seg, SORTED ? R->code : -1, n);
      const int nl = l.n0 + l.n1;
      const double rc = R->rc, invrc = R->invrc, eta = R->eta, drs = R->drs, rs0 = R->rs0;
      const bool shared = R->sharedFc != 0;
      double kk[FNET_RCHUNK - 1];
#pragma unroll
      for (int m = 0; m < FNET_RCHUNK - 1; m++) kk[m] = R->kk[m];
      const double kk7 = R->kk7, c16 = R->c16;
      for (int ch = 0; ch * FNET_RCHUNK < fCnt; ch += 2) {
        const bool two = (ch + 1) * FNET_RCHUNK < fCnt;      // warp-uniform: the second chunk of the pair exists
        const double rsf = rs0 + (double)(ch * FNET_RCHUNK) * drs;
        double acc[FNET_RCHUNK], acc2[FNET_RCHUNK];
#pragma unroll
        for (int f = 0; f < FNET_RCHUNK; f++) { acc[f] = 0.0; acc2[f] = 0.0; }
#pragma unroll NBU
        for (int t = sl; t < nl; t += LPA) {
          const int a = SORTED ? list_at(l, t) : t;
          const double2 rf = *(const double2 *)(rec + 6 * a + 4);
          const double rr = rf.x;
          const double fc = shared ? rf.y : ((rr > rc) ? 0.0 : cutoff_fn(rr, 1.0, invrc));
          const double u = rr - rsf;
          const double e0 = eta * u * u, a1 = 2.0 * eta * drs * u;
          if (e0 < 690.0 && fabs(a1) < 690.0) {         // g_0 and the ratio stay normal numbers
            double gv = fnet_exp_tab(-e0, ftab) * fc;
            const double A = fnet_exp_tab(a1, ftab);
            acc[0] += gv;
#pragma unroll
            for (int m = 0; m < FNET_RCHUNK - 1; m++) { gv *= A * kk[m]; acc[m + 1] += gv; }
            if (two) {
              gv *= A * kk7; acc2[0] += gv;
              const double A2 = A * c16;
#pragma unroll
              for (int m = 0; m < FNET_RCHUNK - 1; m++) { gv *= A2 * kk[m]; acc2[m + 1] += gv; }
            }
          } else {                                      // out of the recurrence's safe range (rare)
#pragma unroll 1
            for (int m = 0; m < (two ? 2 : 1) * FNET_RCHUNK; m++) {
              const double d = u - (double)m * drs;
              const double v = radial_term_generic(FNETGPU_G2, eta, 0.0, d) * fc;
#pragma unroll
              for (int f = 0; f < FNET_RCHUNK; f++) { acc[f] += (f == m) ? v : 0.0; acc2[f] += (f + FNET_RCHUNK == m) ? v : 0.0; }
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (h == 1 && !two) break;
          double v[RR];
          lean_group_reduce<FNET_RCHUNK, LPA>(h ? acc2 : acc, lane, red, v);
#pragma unroll
          for (int q = 0; q < RR; q++) {
            const int f = (ch + h) * FNET_RCHUNK + (FNET_RCHUNK >= LPA ? sl + q * LPA : sl);
            if ((FNET_RCHUNK >= LPA || sl < FNET_RCHUNK) && f < fCnt) outv[tab.rfeat[R->fBeg + f]] = v[q];
          }
        }
      }
    }
    // ---------------- angular passes (acsf.F90:1377-1492) ----------------
    bool sharedFe = true;                             // the records still hold fc E of the shared (rc, eta)
    for (int pi_ = 0; pi_ < lt.nPasses; pi_++) {
      const LeanPass *__restrict__ P = &passes[pi_];
      const int same = P->same, m0 = P->m0;
      const NbList l1 = lean_list(tab, seg, SORTED ? P->code1 : -1, n);
      const NbList l2 = (!SORTED || same) ? l1 : lean_list(tab, seg, P->code2, n);
      const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
      __syncwarp();
      if (P->recomp) {
        sharedFe = false;
        const double rc = P->rc, invrc = P->invrc, eta = P->eta;
#pragma unroll 1
        for (int t = sl; t < n; t += LPA) {
          double fc;
          const double fe = lean_fce(rec[6 * t + 4], rc, invrc, eta, ftab, fc);
          rec[6 * t + 3] = fe;
          if (F32A) recf[t].w = (float)fe;
        }
        __syncwarp();
      }
      double lam[NL];
#pragma unroll
      for (int l = 0; l < NL; l++) lam[l] = P->lam[l];
      AT acc[M];
#pragma unroll
      for (int f = 0; f < M; f++) acc[f] = (AT)0;
      if (same) {
        int nmaxG = n1;                              // the table covers lists of <= FNET_PAIR_TAB_MAXN neighbours: warp-uniform choice
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nmaxG = max(nmaxG, __shfl_xor_sync(0xffffffffu, nmaxG, o));
        if (nmaxG <= FNET_PAIR_TAB_MAXN) lean_pair_loop<NL, NC, LPA, SORTED, 0, AT>(acc, rec, recf, l1, l2, n1, n2, m0, lam, pt, ptab, lt, sl);
        else lean_pair_loop<NL, NC, LPA, SORTED, 1, AT>(acc, rec, recf, l1, l2, n1, n2, m0, lam, pt, ptab, lt, sl);
      } else {
        lean_pair_loop<NL, NC, LPA, SORTED, 2, AT>(acc, rec, recf, l1, l2, n1, n2, m0, lam, pt, ptab, lt, sl);
      }
      // diagonal of identical lists: S0 = sum fcE^2, S1 = sum fcE^2 eps, eps = 1e-13 / r^2 (a 1e-14 correction: FP32 reciprocal)
      double S0 = dS0, S1 = dS1;
      if (same && (SORTED || !sharedFe)) {
        S0 = 0.0; S1 = 0.0;
#pragma unroll 2
        for (int t = sl; t < n1; t += LPA) {
          const int a = SORTED ? list_at(l1, t) : t;
          const double fe = rec[6 * a + 3], rr = rec[6 * a + 4];
          const double e2 = fe * fe;
          const float rf = (float)rr;
          S0 += e2; S1 = fma(e2, (double)__fdividef(1e-13f, rf * rf), S1);
        }
#pragma unroll
        for (int o = LPA / 2; o > 0; o >>= 1) {
          S0 += __shfl_xor_sync(0xffffffffu, S0, o);
          S1 += __shfl_xor_sync(0xffffffffu, S1, o);
        }
      }
      double v[RA];
      if constexpr (F32A) {
        double accd[M];
#pragma unroll
        for (int f = 0; f < M; f++) accd[f] = (double)acc[f];
        lean_group_reduce<M, LPA>(accd, lane, red, v);
      } else {
        lean_group_reduce<M, LPA>(acc, lane, red, v);
      }
#pragma unroll
      for (int q = 0; q < RA; q++) {
        const int e = M >= LPA ? sl + q * LPA : (sl & (M - 1));
        const int ofeat = P->feat[e];
        if ((M >= LPA || sl < M) && ofeat >= 0) outv[ofeat] = fma(P->pref[e], v[q], fma(P->dA[e], S0, P->dB[e] * S1));
      }
    }
    __syncwarp();
    // ---------------- feature row write (+ z-score, + external features) ----------------
    if (act) {
      if (f32) {
        float *out = (float *)featv + (size_t)nFeat * i;
#pragma unroll 1
        for (int a = sl; a < F; a += LPA) out[a] = (float)((outv[a] - zmu[a]) * zis[a]);
#pragma unroll 1
        for (int e = sl; e < nExtSel; e += LPA) out[F + e] = (float)ext[(size_t)nExt * i + extIdx[e]];
      } else {
        double *out = (double *)featv + (size_t)nFeat * i;
#pragma unroll 2
        for (int a = sl; a < F; a += LPA) out[a] = (outv[a] - zmu[a]) * zis[a];
#pragma unroll 1
        for (int e = sl; e < nExtSel; e += LPA) out[F + e] = ext[(size_t)nExt * i + extIdx[e]];
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o >= LPA; o >>= 1) nmaxW = max(nmaxW, __shfl_xor_sync(0xffffffffu, nmaxW, o));
  if (lane == 0 && nmaxW > 0) atomicMax(&flags[0], nmaxW);   // exact capacity hint for the next launch
}
#endif   // FNET_DEFINE_LEAN_KERNEL
