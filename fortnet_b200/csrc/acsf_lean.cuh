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
//   * per neighbour: unit vectors, sqrt(1e-13) / r, fc and fc exp(-eta r^2) are formed once (rsqrt +
//     one third-order Newton step instead of sqrt and a division) and shared by the radial groups and
//     all angular passes with the same (rc, eta); invalid lanes of the last sweep are routed to a
//     dummy neighbour with fc = 0, so the pair loop has no predicate at all;
//   * the z-score is applied as (g - mu) * (1 / sigma) with the reciprocals formed once per CTA.
// Everything that is not a fresh xi-ladder from xi = 1 with one common delta (explicit function
// lists, G1 / G3 / G4, atom-id scaling, lam < -1) runs through k_acsf.
#pragma once
#include "acsf.cuh"

struct LeanWarp {
  double *ux, *uy, *uz, *w, *fcE, *r, *fc;
  int *seg;
  double *outv, *red;
};

__host__ __device__ inline size_t lean_warp_smem_bytes(int cap, int F, int redRows) {
  size_t b = (size_t)cap * 7 * sizeof(double);
  b += (FNET_MAX_CODES + 4) * sizeof(int);
  b = (b + 15) & ~(size_t)15;
  b += (size_t)((F + 1) & ~1) * sizeof(double);
  b += (size_t)redRows * FNET_RED_STRIDE * sizeof(double);
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t lean_cta_extra_bytes(int F) {   // power tables + (mu, 1/sigma)
  return ((size_t)(FNET_POW_DOUBLES + 2 * ((F + 1) & ~1)) * sizeof(double) + 15) & ~(size_t)15;
}
__device__ __forceinline__ LeanWarp lean_carve(unsigned char *base, int cap, int F) {
  LeanWarp w;
  double *d = (double *)base;
  w.ux = d; w.uy = d + cap; w.uz = d + 2 * cap; w.w = d + 3 * cap; w.fcE = d + 4 * cap; w.r = d + 5 * cap; w.fc = d + 6 * cap;
  w.seg = (int *)(d + 7 * cap);
  size_t off = (size_t)cap * 7 * sizeof(double) + (FNET_MAX_CODES + 4) * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  w.outv = (double *)(base + off);
  w.red = w.outv + ((F + 1) & ~1);
  return w;
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

// Gathers the neighbours of `me` into (ux, uy, uz) as raw displacements (sorted by species code
// when SORTED: [code 0 .. | other | self-images], acsf.F90:754-760); returns n or -needed.
template <int PATH, bool SORTED>
__device__ __forceinline__ int lean_gather(const CRec &me, const CtaGeom &cg, const AcsfTables &tab, int cap,
                                           LeanWarp &w) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double rc2 = tab.rcMax * tab.rcMax;
  double *gx = SORTED ? w.fcE : w.ux, *gy = SORTED ? w.r : w.uy, *gz = SORTED ? w.fc : w.uz;
  int *gc = (int *)w.w;
  int n = 0;
  auto take = [&](bool valid, double dx, double dy, double dz, int j, int zs) {
    const bool ok = is_neighbor(valid, dx * dx + dy * dy + dz * dz, rc2, j, zs, me.idx);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int pos = n + __popc(m & lt);
    if (ok && pos < cap - 1) {
      gx[pos] = dx; gy[pos] = dy; gz[pos] = dz;
      if (SORTED) gc[pos] = (j == me.idx) ? tab.nCodes + 1 : species_code(tab, zs & ~FNET_SHIFT_FLAG);
    }
    n += __popc(m);
  };
  if (PATH == FNET_PATH_STRUCT) for_each_candidate_struct(cg.cand, cg.nCand, cg.sg, me, take);
  else if (PATH == FNET_PATH_STAGED) for_each_candidate_staged(cg.cand, cg.nCand, me, take);
  else for_each_candidate_direct(*cg.S, cg.bp, cg.cellStart, cg.crec, me, take);
  if (n > cap - 1) return -(n + 1);          // one slot is the dummy neighbour
  __syncwarp();
  if (SORTED) {
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
      int code = -1;
      double x = 0, y = 0, z = 0;
      if (t < n) { x = gx[t]; y = gy[t]; z = gz[t]; code = gc[t]; }
      int pos = -1;
      for (int c = 0; c < nc; c++) {
        const unsigned m = __ballot_sync(0xffffffffu, code == c);
        const int b = __shfl_sync(0xffffffffu, mybase, c);
        if (code == c) pos = b + __popc(m & lt);
        if (lane == c) mybase += __popc(m);
      }
      if (pos >= 0) { w.ux[pos] = x; w.uy[pos] = y; w.uz[pos] = z; }
    }
    __syncwarp();
  }
  return n;
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

template <typename real, int NL, int NC, int PATH, bool SORTED>
__global__ void __launch_bounds__(128, (NL * NC <= 2 ? 5 : 3))
k_acsf_lean(int nSplit, GeomArgs geo, int nExt, const double *__restrict__ ext, AcsfTables tab, LeanTables lt, int cap,
            int capC, real *__restrict__ feat, int nFeat, const double *__restrict__ zprec, int nExtSel,
            const int *__restrict__ extIdx, int *__restrict__ flags) {
  constexpr int M = NL * NC * FNET_LADDER;          // accumulators per lane
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  CtaGeom cg;
  unsigned char *wbase;
  if (!acsf_cta_prologue<PATH>(geo, nSplit, tab.rcMax, capC, smem_raw, flags, cg, wbase)) return;
  const int F = tab.F, Fp = (F + 1) & ~1;
  double *pt = (double *)wbase;                     // power tables
  double *zmu = pt + FNET_POW_DOUBLES, *zis = zmu + Fp;
  for (int e = threadIdx.x; e < FNET_POW_DOUBLES; e += blockDim.x) pt[e] = lt.powtab[e];
  for (int a = threadIdx.x; a < F; a += blockDim.x) {
    double mu = 0.0, is = 1.0;
    if (zprec) { const double sg = zprec[F + a]; if (!(sg < 1e-08)) { mu = zprec[a]; is = 1.0 / sg; } }   // acsf.F90:505-507
    zmu[a] = mu; zis[a] = is;
  }
  __syncthreads();
  wbase += lean_cta_extra_bytes(F);
  const int a0 = cg.a0, a1 = cg.a1;
  LeanWarp w = lean_carve(wbase + (size_t)wib * lean_warp_smem_bytes(cap, F, lt.redRows), cap, F);
  const double *ftab = cg.ftab;
  const double sq13 = 3.1622776601683794e-07;       // sqrt(1e-13)
  int nmaxW = 0;
  for (int slot = a0 + wib; slot < a1; slot += nw) {
    const CRec me = central_atom<PATH>(cg, slot);
    const int i = me.idx;
    const int n = lean_gather<PATH, SORTED>(me, cg, tab, cap, w);
    if (n < 0) { if (lane == 0) atomicMax(&flags[1], -n); continue; }
    nmaxW = max(nmaxW, n);
    // ---------------- per-neighbour quantities; entry n is the dummy neighbour ----------------
    for (int t = lane; t <= n; t += 32) {
      double ux = 0.0, uy = 0.0, uz = 0.0, ww = 0.0, fe = 0.0, rr = 2.0 * tab.rcMax, fc = 0.0;
      if (t < n) {
        const double dx = w.ux[t], dy = w.uy[t], dz = w.uz[t];
        double ri;
        lean_rsqrt(dx * dx + dy * dy + dz * dz, ri, rr);   // dynneighlist.F90:311
        ux = dx * ri; uy = dy * ri; uz = dz * ri;          // acsf.F90:1565
        ww = ri * sq13;
        fe = lean_fce(rr, lt.rcShared, lt.invrcShared, lt.etaShared, ftab, fc);
      }
      w.ux[t] = ux; w.uy[t] = uy; w.uz[t] = uz; w.w[t] = ww; w.fcE[t] = fe; w.r[t] = rr; w.fc[t] = fc;
    }
    if (!SORTED && lane == 0) { w.seg[0] = 0; w.seg[1] = n; w.seg[2] = n; }
    __syncwarp();
    // ---------------- radial ladder groups (acsf.F90:1287-1373) ----------------
    for (int g = 0; g < lt.nRadial; g++) {
      const LeanRadial *__restrict__ G = &lt.rad[g];
      const int nch = G->nch, lgn = G->lgn, fCnt = G->fCnt;
      const NbList l = lean_list(tab, w.seg, SORTED ? G->code : -1, n);
      const int nl = l.n0 + l.n1;
      const int per = 32 >> lgn;
      const int mychunk = lane & (nch - 1), sub = lane >> lgn;
      const int fcnt = min(max(fCnt - mychunk * FNET_RCHUNK, 0), FNET_RCHUNK);
      const double rc = G->rc, invrc = G->invrc, eta = G->eta, drs = G->drs;
      const double rsf = G->rs0 + (double)(mychunk * FNET_RCHUNK) * drs;
      const bool shared = G->sharedFc != 0;
      double acc[FNET_RCHUNK], kk[FNET_RCHUNK - 1];
#pragma unroll
      for (int f = 0; f < FNET_RCHUNK; f++) acc[f] = 0.0;
#pragma unroll
      for (int m = 0; m < FNET_RCHUNK - 1; m++) kk[m] = G->kk[m];
      if (fcnt > 0)
        for (int t = sub; t < nl; t += per) {
          const int a = SORTED ? list_at(l, t) : t;
          const double rr = w.r[a];
          const double fc = shared ? w.fc[a] : ((rr > rc) ? 0.0 : cutoff_fn(rr, 1.0, invrc));
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
      const double v = reduce_smem<FNET_RCHUNK>(acc, lane, nch, w.red);
      const int f = (lane >> lgn) & (FNET_RCHUNK - 1);
      if (lane < nch * FNET_RCHUNK && f < fcnt) w.outv[tab.rfeat[G->fBeg + mychunk * FNET_RCHUNK + f]] = v;
    }
    // ---------------- angular passes (acsf.F90:1377-1492) ----------------
    for (int pi_ = 0; pi_ < lt.nPasses; pi_++) {
      const LeanPass *__restrict__ P = &lt.pass[pi_];
      const int same = P->same, m0 = P->m0;
      const NbList l1 = lean_list(tab, w.seg, SORTED ? P->code1 : -1, n);
      const NbList l2 = (!SORTED || same) ? l1 : lean_list(tab, w.seg, P->code2, n);
      const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
      __syncwarp();
      if (P->recomp) {
        const double rc = P->rc, invrc = P->invrc, eta = P->eta;
        for (int t = lane; t < n; t += 32) { double fc; w.fcE[t] = lean_fce(w.r[t], rc, invrc, eta, ftab, fc); }
        __syncwarp();
      }
      // prefetch the epilogue's per-function constants (lane e of the pass)
      const int e = lane & (M - 1);
      const int ofeat = __ldg(&P->feat[e]);
      const double opref = __ldg(&P->pref[e]), odA = __ldg(&P->dA[e]), odB = __ldg(&P->dB[e]);
      double lam[NL];
#pragma unroll
      for (int l = 0; l < NL; l++) lam[l] = P->lam[l];
      double acc[M];
#pragma unroll
      for (int f = 0; f < M; f++) acc[f] = 0.0;
      const int nP = same ? (n1 * (n1 - 1)) >> 1 : n1 * n2;
      const float invW = n2 > 0 ? 1.0f / (float)n2 : 0.0f;
      const bool useTab = same && n1 <= FNET_PAIR_TAB_MAXN;
      for (int p0 = 0; p0 < nP; p0 += 32) {
        const int p = min(p0 + lane, nP);            // p = nP: (0, n1) resp. (n1, 0) -> the dummy neighbour
        int j, k;
        if (same) {
          if (useTab) { const unsigned jk = __ldg(&lt.pairtab[p]); j = jk & 255; k = jk >> 8; }
          else lean_tri_decode(p, j, k);
        } else {
          j = (int)(((float)p + 0.5f) * invW);
          if (j * n2 > p) j--;
          else if ((j + 1) * n2 <= p) j++;
          k = p - j * n2;
        }
        const int a = SORTED ? list_at(l1, j) : j, b = SORTED ? list_at(l2, k) : k;
        const double base = w.fcE[a] * w.fcE[b];
        double dot = w.ux[a] * w.ux[b];
        dot = fma(w.uy[a], w.uy[b], dot);
        dot = fma(w.uz[a], w.uz[b], dot);
        const double c = fma(-dot, w.w[a] * w.w[b], dot);   // dot / (r_a r_b + 1e-13), acsf.F90:1173-1174
#pragma unroll
        for (int l = 0; l < NL; l++) {
          const double bb = fma(lam[l], c, 1.0);
          const double q = lean_pow(bb, pt, lt);
          double pw = bb * base;
          const double q2 = q * q, q4 = q2 * q2;
          if (m0 > 0) {                                      // later 8 NC-function blocks of a long ladder: b q^m0
            double qm = 1.0, qb = q4 * q4;                   // m0 is a multiple of 8
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
      // diagonal of identical lists: S0 = sum fcE^2, S1 = sum fcE^2 eps, eps = 1e-13 / r^2
      double S0 = 0.0, S1 = 0.0;
      if (same) {
        for (int t = lane; t < n1; t += 32) {
          const int a = SORTED ? list_at(l1, t) : t;
          const double fe = w.fcE[a], ww = w.w[a];
          const double e2 = fe * fe;
          S0 += e2; S1 = fma(e2, ww * ww, S1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          S0 += __shfl_xor_sync(0xffffffffu, S0, o);
          S1 += __shfl_xor_sync(0xffffffffu, S1, o);
        }
      }
      const double v = reduce_smem<M>(acc, lane, 1, w.red);
      if (lane < M && ofeat >= 0) w.outv[ofeat] = fma(opref, v, fma(odA, S0, odB * S1));
    }
    __syncwarp();
    // ---------------- coalesced feature write (+ z-score, + external features) ----------------
    real *out = feat + (size_t)nFeat * i;
    for (int a = lane; a < F; a += 32) out[a] = (real)((w.outv[a] - zmu[a]) * zis[a]);
    for (int e = lane; e < nExtSel; e += 32) out[F + e] = (real)ext[(size_t)nExt * i + extIdx[e]];
    __syncwarp();
  }
  if (lane == 0 && nmaxW > 0) atomicMax(&flags[0], nmaxW);   // exact capacity hint for the next launch
}
