// acsf_force_lean.cuh -- fused analytic forces for automatic-scheme ACSF configurations: the derivative
// counterpart of acsf_lean.cuh (same tables, same G central atoms per warp, same b^delta power tables).
//
// Replaces TAcsf_calculatePrime -> iGeoAcsfGrad -> gFuncGrad (lib_descriptors/acsf.F90:643-717, 870-939,
// 1496-1697), applyZscorePrime (:515-536) and the contraction of forceAnalysis_analytical
// (lib_analysis/forces.F90:400-413) like k_acsf_force (acsf_force.cuh), with three changes:
//   * CYCLIC pair walk: the lane that owns list position j visits the pairs (j, (j + o) mod n),
//     o = 1 .. n/2.  Every unordered pair is met once, the first member's record and its force
//     accumulator stay in REGISTERS for the whole walk, and in every step the second members of the
//     lanes are pairwise different -- their contributions are plain shared-memory read-modify-writes
//     in a fixed order: no atomics, no segmented reductions, no pair-index decode;
//   * per pair and lambda ONE Horner sweep gives both P(q) = sum_m c_m q^m and P'(q) (q = b^delta, c_m =
//     dE/dG_m 2^(1-xi_m)), from which S0 = sum c_m b^xi_m = b P and S1 = sum c_m xi_m lam b^(xi_m - 1) =
//     lam (P + delta q P'): one coefficient set in registers instead of two;
//   * DETERMINISTIC forces on the whole-structure path: one CTA owns a structure, every warp adds the
//     contributions of its central atoms to its own shared-memory copy of the structure's forces (groups
//     of a warp in turn), the copies are summed in a fixed order and STORED -- no floating-point atomics
//     anywhere, results are bit-identical from run to run (forces.F90:400-413 is order-fixed, too).
//     The cell-list path (structures too large to stage) scatters with red.global.add.f64.
// The diagonal j == k of identical lists is closed-form as in acsf_lean.cuh: d/dR_j of dA fcE_j^2.
#pragma once
#include "acsf_lean.cuh"

#define FNET_FREC 10   // doubles per neighbour record: u_x, u_y | u_z, E | E', 1/r | r, fc | fc', atom index (int bits)

__host__ __device__ inline size_t force_lean_group_bytes(int cap, int F, int M, bool sorted) {   // cap: multiple of 8
  size_t b = (size_t)cap * FNET_FREC * sizeof(double);                       // records
  b += (size_t)((FNET_MAX_CODES + 4 + 3) & ~3) * sizeof(int);               // list segments
  // force accumulators | dE/dG row | pass coefficients, diagonal sums -- and, species-resolved configurations, in the SAME
  // bytes the unsorted displacements, codes and atom indices of the group: they are dead once the counting sort has
  // filled the records, before the accumulators are cleared and the dE/dG row is loaded (C3: 82.5 -> 74.5 KB per CTA,
  // 3 CTAs / 12 warps per SM instead of 2 / 8)
  size_t u = (size_t)cap * 3 * sizeof(double) + (size_t)(((F + 1) & ~1) + M + 2) * sizeof(double);
  const size_t sc = sorted ? (size_t)cap * (3 * sizeof(double) + 2 * sizeof(int)) : 0;
  if (sc > u) u = sc;
  b += u;
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t force_lean_warp_bytes(int cap, int F, int M, bool sorted, int G, int localAtoms) {
  return (size_t)G * force_lean_group_bytes(cap, F, M, sorted) + (((size_t)3 * localAtoms * sizeof(double) + 15) & ~(size_t)15);
}
__host__ __device__ inline size_t force_lean_cta_tables_bytes(int stageBytes) {
  return ((size_t)FNET_POW_DOUBLES * sizeof(double) + 128 + (size_t)((stageBytes + 15) & ~15) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t force_lean_cta_extra_bytes(int stageBytes) {   // see lean_cta_extra_bytes
  return force_lean_cta_tables_bytes(stageBytes) - 2 * FNET_LOG_TAB_N * sizeof(double);
}

// d fc / dr of fc = cos^2(pi r / (2 rc)): -(pi / (2 rc)) sin(pi r / rc)   (dfCutoffWoCheck, acsf.F90:1253-1254)
__device__ __forceinline__ double lean_dcutoff(double rr, double invrc) {
  return -0.5 * (3.14159265358979323846 * invrc) * sinpi(rr * invrc);
}

template <int NL, int NC, int PATH, bool SORTED, int G, bool LOCAL>
__global__ void __launch_bounds__(128, (NL * NC <= 2 ? 4 : 3))
k_acsf_force_lean(int nSplit, GeomArgs geo, AcsfTables tab, LeanTables lt, int cap, int capC, int localAtoms,
                  const double *__restrict__ dEdG, int nOut, const double *__restrict__ zprec,
                  double *__restrict__ forces, double *__restrict__ fpart, int *__restrict__ flags) {
  constexpr int MH = NC * FNET_LADDER;              // coefficients per lambda-group
  constexpr int M = NL * MH;
  constexpr int LPA = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int grp = lane / LPA, sl = lane & (LPA - 1), gshift = grp * LPA;
  const unsigned lowmask = LPA == 32 ? 0xffffffffu : ((1u << LPA) - 1u);
  const unsigned gmask = lowmask << gshift;
  const unsigned ltmask = (1u << sl) - 1u;
  const int kt = blockIdx.z;                        // target index
  CtaGeom cg;
  unsigned char *wbase;
  if (!acsf_cta_prologue<PATH, true>(geo, nSplit, tab.rcMax, capC, smem_raw, flags, cg, wbase)) return;
  const int F = tab.F, Fp = (F + 1) & ~1;
  double *pt = (double *)wbase;
  unsigned char *zcode = (unsigned char *)(pt + FNET_POW_DOUBLES);
  for (int e = threadIdx.x; e < FNET_POW_DOUBLES; e += blockDim.x) pt[e] = lt.powtab[e];
  if (SORTED) for (int z = threadIdx.x; z < 128; z += blockDim.x) zcode[z] = (unsigned char)species_code(tab, z);
  unsigned char *stage = zcode + 128;
  const LeanPass *passes = lt.pass;
  const LeanRadial *rads = lt.rad;
  if (lt.stageBytes > 0) {
    const int nw8 = lt.stageBytes >> 3, np8 = (int)((lt.nPasses * sizeof(LeanPass)) >> 3);
    double *dst = (double *)stage;
    const double *srcP = (const double *)lt.pass, *srcR = (const double *)lt.rad;
    for (int e = threadIdx.x; e < nw8; e += blockDim.x) dst[e] = e < np8 ? srcP[e] : srcR[e - np8];
    passes = (const LeanPass *)stage;
    rads = (const LeanRadial *)(stage + (size_t)np8 * 8);
  }
  wbase += force_lean_cta_tables_bytes(lt.stageBytes);
  const int a0 = cg.a0, a1 = cg.a1;
  const size_t gbytes = force_lean_group_bytes(cap, F, M, SORTED);
  const size_t wbytes = force_lean_warp_bytes(cap, F, M, SORTED, G, localAtoms);
  unsigned char *wb = wbase + (size_t)wib * wbytes;
  unsigned char *gb = wb + (size_t)grp * gbytes;
  double *rec = (double *)gb;                                        // [cap][FNET_FREC]
  int *seg = (int *)(rec + (size_t)FNET_FREC * cap);
  double *fa = (double *)(seg + ((FNET_MAX_CODES + 4 + 3) & ~3));    // [cap][3] force accumulators of the neighbours
  double *gx = fa, *gy = gx + cap, *gz = gy + cap;                   // SORTED: unsorted displacements, species codes, atom indices --
  int *gc = (int *)(gz + cap), *gi = gc + cap;                       //   in the bytes of fa | Dv | cbuf (dead until the records are sorted)
  double *Dv = fa + 3 * cap;                                         // dE/dG row of the central atom (/ sigma)
  double *cbuf = Dv + Fp;                                            // [M] pass coefficients, [M], [M + 1]: diagonal sums
  double *loc = (double *)(wb + (size_t)G * gbytes);                 // LOCAL: this warp's copy of the structure's forces
  if (LOCAL) for (int e = lane; e < 3 * localAtoms; e += 32) loc[e] = 0.0;
  __syncthreads();
  const double *ftab = cg.ftab;
  const double delta = lt.powC[0];
  const int stride = 3 * nOut;
  for (int s0 = a0 + wib * G; s0 < a1; s0 += nw * G) {
    const int slot = s0 + grp;
    bool act = slot < a1;
    const CRec me = central_atom<PATH>(cg, act ? slot : a0);
    const int i = me.idx;
    // ---------------- neighbours (as in k_acsf_lean, plus the atom index of every neighbour) ----------------
    const int n0 = lean_gather_linear<PATH, SORTED, G, FNET_FREC, true>(cg, tab, me, act, cap, (double *)wb, (double *)wb + (gx - rec),
                                                                        (double *)wb + (gy - rec), (double *)wb + (gz - rec),
                                                                        (int *)wb + (gc - (int *)rec), (int *)wb + (gi - (int *)rec),
                                                                        gbytes, gbytes, zcode, lane);
    int n = n0;
    if (n > cap - 1) {
      if (sl == 0) atomicMax(&flags[1], n + 1);
      act = false; n = 0;
    }
    __syncwarp();
    if (SORTED) {
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
        int code = -1, j = -1;
        double x = 0, y = 0, z = 0;
        if (t < n) { x = gx[t]; y = gy[t]; z = gz[t]; code = gc[t]; j = gi[t]; }
        int pos = -1;
        for (int c = 0; c < nc; c++) {
          const unsigned mg = (__ballot_sync(0xffffffffu, code == c) >> gshift) & lowmask;
          const int b = __shfl_sync(0xffffffffu, mybase, c, LPA);
          if (code == c) pos = b + __popc(mg & ltmask);
          if (sl == c) mybase += __popc(mg);
        }
        if (pos >= 0) { double *q = rec + (size_t)FNET_FREC * pos; q[0] = x; q[1] = y; q[2] = z; ((int *)(q + 9))[0] = j; }
      }
      __syncwarp();
    }
    // ---------------- per-neighbour record (entry n: dummy), zeroed accumulators, dE/dG row ----------------
    for (int t = sl; t <= n; t += LPA) {
      double *q = rec + (size_t)FNET_FREC * t;
      double ux = 0.0, uy = 0.0, uz = 0.0, E = 0.0, Ep = 0.0, ri = 0.0, rr = 2.0 * tab.rcMax, fc = 0.0, fcp = 0.0;
      if (t < n) {
        const double dx = q[0], dy = q[1], dz = q[2];
        lean_rsqrt(dx * dx + dy * dy + dz * dz, ri, rr);
        ux = dx * ri; uy = dy * ri; uz = dz * ri;            // acsf.F90:1565
        if (!(rr > lt.rcShared)) {
          fc = cutoff_fn(rr, 1.0, lt.invrcShared);
          fcp = lean_dcutoff(rr, lt.invrcShared);
          const double ex = fnet_exp_tab(-lt.etaShared * rr * rr, ftab);
          E = fc * ex;
          Ep = (fcp - 2.0 * lt.etaShared * rr * fc) * ex;    // acsf.F90:1577-1578
        }
      }
      *(double2 *)(q) = make_double2(ux, uy);
      *(double2 *)(q + 2) = make_double2(uz, E);
      *(double2 *)(q + 4) = make_double2(Ep, ri);
      *(double2 *)(q + 6) = make_double2(rr, fc);
      q[8] = fcp;
      fa[3 * t] = 0.0; fa[3 * t + 1] = 0.0; fa[3 * t + 2] = 0.0;
    }
    for (int a = sl; a < F; a += LPA) {       // z-score: derivatives / sigma (acsf.F90:528-534)
      double d = act ? dEdG[((size_t)nOut * i + kt) * F + a] : 0.0;
      if (zprec) { const double sg = zprec[F + a]; if (!(sg < 1e-08)) d /= sg; }
      Dv[a] = d;
    }
    __syncwarp();
    // ---------------- radial (acsf.F90:1527-1557): lanes = neighbours ----------------
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
      for (int ch = 0; ch * FNET_RCHUNK < fCnt; ch++) {
        const double rsf = rs0 + (double)(ch * FNET_RCHUNK) * drs;
        double D[FNET_RCHUNK];
#pragma unroll
        for (int m = 0; m < FNET_RCHUNK; m++) {
          const int f = ch * FNET_RCHUNK + m;
          D[m] = f < fCnt ? Dv[tab.rfeat[R->fBeg + f]] : 0.0;
        }
        for (int t = sl; t < nl; t += LPA) {
          const int a = SORTED ? list_at(l, t) : t;
          double *q = rec + (size_t)FNET_FREC * a;
          const double rr = q[6];
          double fc = q[7], fcp = q[8];
          if (!shared) {
            if (rr > rc) { fc = 0.0; fcp = 0.0; }
            else { fc = cutoff_fn(rr, 1.0, invrc); fcp = lean_dcutoff(rr, invrc); }
          }
          const double u = rr - rsf;
          const double e0 = eta * u * u, a1 = 2.0 * eta * drs * u;
          double s = 0.0;
          if (e0 < 690.0 && fabs(a1) < 690.0) {
            double gv = fnet_exp_tab(-e0, ftab);
            const double A = fnet_exp_tab(a1, ftab);
            const double te = 2.0 * eta * fc;
            double w = fma(-te, u, fcp);                       // fc' - 2 eta (r - rs_m) fc, rs_m = rsf + m drs
            const double dw = te * drs;
            s = D[0] * w * gv;
#pragma unroll
            for (int m = 1; m < FNET_RCHUNK; m++) { gv *= A * kk[m - 1]; w += dw; s = fma(D[m] * w, gv, s); }
          } else {
#pragma unroll 1
            for (int m = 0; m < FNET_RCHUNK; m++) {
              const double d = u - (double)m * drs;
              double Dm = 0.0;
#pragma unroll
              for (int f = 0; f < FNET_RCHUNK; f++) Dm += (f == m) ? D[f] : 0.0;
              s += Dm * (fcp - 2.0 * eta * d * fc) * radial_term_generic(FNETGPU_G2, eta, 0.0, d);
            }
          }
          fa[3 * a] = fma(s, q[0], fa[3 * a]); fa[3 * a + 1] = fma(s, q[1], fa[3 * a + 1]); fa[3 * a + 2] = fma(s, q[2], fa[3 * a + 2]);
        }
        __syncwarp();
      }
    }
    // ---------------- angular passes (acsf.F90:1559-1668) ----------------
    for (int pi_ = 0; pi_ < lt.nPasses; pi_++) {
      const LeanPass *__restrict__ P = &passes[pi_];
      const int same = P->same, m0 = P->m0;
      const NbList l1 = lean_list(tab, seg, SORTED ? P->code1 : -1, n);
      const NbList l2 = (!SORTED || same) ? l1 : lean_list(tab, seg, P->code2, n);
      const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
      __syncwarp();
      if (P->recomp) {
        const double rc = P->rc, invrc = P->invrc, eta = P->eta;
        for (int t = sl; t < n; t += LPA) {
          double *q = rec + (size_t)FNET_FREC * t;
          const double rr = q[6];
          double E = 0.0, Ep = 0.0;
          if (!(rr > rc)) {
            const double fc = cutoff_fn(rr, 1.0, invrc), fcp = lean_dcutoff(rr, invrc);
            const double ex = fnet_exp_tab(-eta * rr * rr, ftab);
            E = fc * ex; Ep = (fcp - 2.0 * eta * rr * fc) * ex;
          }
          q[3] = E; q[4] = Ep;
        }
      }
      // coefficients c_e = dE/dG_e 2^(1 - xi_e) (x 2 for unordered pairs); diagonal sums  sum_e D_e dA_e, sum_e D_e dB_e
      {
        double sA = 0.0, sB = 0.0;
        for (int e = sl; e < M; e += LPA) {
          const int of = P->feat[e];
          const double d = of >= 0 ? Dv[of] : 0.0;
          cbuf[e] = d * P->pref[e];
          sA = fma(d, P->dA[e], sA); sB = fma(d, P->dB[e], sB);
        }
#pragma unroll
        for (int o = LPA / 2; o > 0; o >>= 1) {
          sA += __shfl_xor_sync(0xffffffffu, sA, o);
          sB += __shfl_xor_sync(0xffffffffu, sB, o);
        }
        if (sl == 0) { cbuf[M] = sA; cbuf[M + 1] = sB; }
      }
      __syncwarp();
      double lam[NL], c0[NL][MH];
#pragma unroll
      for (int l = 0; l < NL; l++) {
        lam[l] = P->lam[l];
#pragma unroll
        for (int m = 0; m < MH; m++) c0[l][m] = cbuf[l * MH + m];
      }
      double qm0s = 1.0;                               // S1 factor of a later ladder block: xi = 1 + (m0 + m) delta
      if (m0 > 0) qm0s = fma((double)m0, delta, 1.0);
      // cyclic walk: rows j = jb + sl of list 1; identical lists: partners (j + o) mod n1, o = 1 .. n1 / 2
      // (n1 even: the last step visits every pair twice -> lanes j < n1 / 2 only); two lists: all n2 partners,
      // start rotated by j (blocks of <= n2 rows keep the partners of a step pairwise different)
      const int blk = same ? LPA : min(LPA, max(n2, 1));
      const int nSteps = same ? (n1 >> 1) : n2;
      const int oBeg = same ? 1 : 0, oEnd = same ? nSteps + 1 : nSteps;
      if (n1 > 0 && nSteps > 0)
        for (int jb = 0; jb < n1; jb += blk) {
          const int j = jb + sl;
          const bool rowOn = sl < blk && j < n1;
          const int a = rowOn ? (SORTED ? list_at(l1, j) : j) : n;
          const double *qa = rec + (size_t)FNET_FREC * a;
          const double2 A0 = *(const double2 *)(qa), A1 = *(const double2 *)(qa + 2), A2 = *(const double2 *)(qa + 4);
          double gax = 0.0, gay = 0.0, gaz = 0.0;
          const int nmod = same ? n1 : n2;
          const int jm = same ? j : j % max(n2, 1);         // two lists: j can exceed n2
          for (int o = oBeg; o < oEnd; o++) {
            int kp = jm + o;
            if (kp >= nmod) kp -= nmod;
            const bool on = rowOn && !(same && 2 * o == n1 && j >= o);
            const int b = on ? (SORTED ? list_at(l2, kp) : kp) : n;
            double *qb = rec + (size_t)FNET_FREC * b;
            const double2 B0 = *(const double2 *)(qb), B1 = *(const double2 *)(qb + 2), B2 = *(const double2 *)(qb + 4);
            double c = A0.x * B0.x;                          // acsf.F90:1591: unit vectors, no regulariser in the derivative
            c = fma(A0.y, B0.y, c);
            c = fma(A1.x, B1.x, c);
            double S0 = 0.0, S1 = 0.0;
#pragma unroll
            for (int l = 0; l < NL; l++) {
              const double bb = fma(lam[l], c, 1.0);         // a last-bit negative b: lean_pow stays finite, the term is ~1e-16
              const double q = lean_pow(bb, pt, lt);
              double Pq = c0[l][MH - 1], dP = 0.0;
#pragma unroll
              for (int m = MH - 2; m >= 0; m--) { dP = fma(dP, q, Pq); Pq = fma(Pq, q, c0[l][m]); }
              double u1 = fma(delta * q, dP, Pq * qm0s);   // sum_m c_m xi_m q^m / q^m0-part
              if (m0 > 0) {
                const double q2 = q * q, q4 = q2 * q2;
                double qm = 1.0, qb2 = q4 * q4;
                for (int t = m0 >> 3; t; t >>= 1) { if (t & 1) qm *= qb2; qb2 *= qb2; }
                Pq *= qm; u1 *= qm;
              }
              S0 = fma(bb, Pq, S0);
              S1 = fma(lam[l], u1, S1);
            }
            // g_a = Eb [ S1 Ea (u_b - c u_a) / r_a + S0 Ea' u_a ],  g_b likewise
            const double X = S1 * (A1.y * B1.y);
            const double ta = X * A2.y, tb = X * B2.y;
            const double sa = fma(S0 * B1.y, A2.x, -ta * c), sb = fma(S0 * A1.y, B2.x, -tb * c);
            gax = fma(ta, B0.x, fma(sa, A0.x, gax));
            gay = fma(ta, B0.y, fma(sa, A0.y, gay));
            gaz = fma(ta, B1.x, fma(sa, A1.x, gaz));
            double *fb = fa + 3 * b;
            fb[0] += fma(tb, A0.x, sb * B0.x);
            fb[1] += fma(tb, A0.y, sb * B0.y);
            fb[2] += fma(tb, A1.x, sb * B1.x);
            __syncwarp(gmask);
          }
          double *fap = fa + 3 * a;
          fap[0] += gax; fap[1] += gay; fap[2] += gaz;
          __syncwarp(gmask);
        }
      if (same) {   // diagonal: d/dR_t of (dA + dB eps_t) fcE_t^2, eps_t = 1e-13 / r_t^2
        const double sA = cbuf[M], sB = cbuf[M + 1];
        for (int t = sl; t < n1; t += LPA) {
          const int a = SORTED ? list_at(l1, t) : t;
          double *q = rec + (size_t)FNET_FREC * a;
          const double ri = q[5];
          const double s = 2.0 * fma(sB, 1e-13 * ri * ri, sA) * q[3] * q[4];
          fa[3 * a] = fma(s, q[0], fa[3 * a]); fa[3 * a + 1] = fma(s, q[1], fa[3 * a + 1]); fa[3 * a + 2] = fma(s, q[2], fa[3 * a + 2]);
        }
      }
      __syncwarp();
    }
    // ---------------- scatter: neighbours get -g, the central atom +sum(g) (forces.F90:400-413) ----------------
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int t = sl; t < n; t += LPA) { sx += fa[3 * t]; sy += fa[3 * t + 1]; sz += fa[3 * t + 2]; }
#pragma unroll
    for (int o = LPA / 2; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (LOCAL) {
#pragma unroll 1
      for (int gq = 0; gq < G; gq++) {       // groups in turn: fixed order of the additions into the warp's copy
        if (grp == gq && act) {
          for (int t = sl; t < n; t += LPA) {
            const int jl = 3 * (((const int *)(rec + (size_t)FNET_FREC * t + 9))[0] - cg.first);
            loc[jl] -= fa[3 * t]; loc[jl + 1] -= fa[3 * t + 1]; loc[jl + 2] -= fa[3 * t + 2];
          }
        }
        __syncwarp();
        if (grp == gq && act && sl == 0) {
          const int il = 3 * (i - cg.first);
          loc[il] += sx; loc[il + 1] += sy; loc[il + 2] += sz;
        }
        __syncwarp();
      }
    } else if (act) {
      for (int t = sl; t < n; t += LPA) {
        const double gxv = fa[3 * t], gyv = fa[3 * t + 1], gzv = fa[3 * t + 2];
        if (gxv != 0.0 || gyv != 0.0 || gzv != 0.0) {
          double *ff = forces + (size_t)stride * ((const int *)(rec + (size_t)FNET_FREC * t + 9))[0] + 3 * kt;
          atomicAdd(ff, -gxv); atomicAdd(ff + 1, -gyv); atomicAdd(ff + 2, -gzv);
        }
      }
      if (sl == 0) {
        double *ff = forces + (size_t)stride * i + 3 * kt;
        atomicAdd(ff, sx); atomicAdd(ff + 1, sy); atomicAdd(ff + 2, sz);
      }
    }
    __syncwarp();
  }
  if (LOCAL) {   // fixed-order sum of the warps' copies, plain stores: to the forces when this CTA owns the whole structure,
    __syncthreads();   // else to this CTA's partial (k_force_reduce sums the CTAs of a structure in a fixed order)
    const int nAt = cg.nCand;                       // atoms of the structure (PATH_STRUCT)
    double *part = fpart + (((size_t)blockIdx.x * gridDim.y + blockIdx.y) * nOut + kt) * (size_t)(3 * localAtoms);
    for (int e = threadIdx.x; e < 3 * nAt; e += blockDim.x) {
      double s = 0.0;
      for (int w2 = 0; w2 < nw; w2++) s += *(const double *)(wbase + (size_t)w2 * wbytes + (size_t)G * gbytes + (size_t)e * sizeof(double));
      if (gridDim.y == 1) forces[(size_t)stride * (cg.first + e / 3) + 3 * kt + e % 3] = s;
      else part[e] = s;
    }
  }
}

// forces of structures that were split over several CTAs (few, large structures: an MD step of one cell):
// out = sum over the nSplit partials in a fixed order
// (and the partials are cleared behind the sum: the next launch finds the zeros its empty CTAs rely on)
__global__ void k_force_reduce(int nStruct, int nSplit, int nOut, int localAtoms, const int *__restrict__ offsets,
                               double *__restrict__ fpart, double *__restrict__ forces) {
  FNET_PDL_TRIGGER();
  FNET_PDL_WAIT();
  const int st = blockIdx.y;
  const int beg = offsets[st], nAt = offsets[st + 1] - beg;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;      // (k, 3 atom + c)
  if (e >= nOut * 3 * nAt) return;
  const int k = e / (3 * nAt), r = e % (3 * nAt);
  double s = 0.0;
  double *q0 = fpart + ((size_t)st * nSplit * nOut + k) * (size_t)(3 * localAtoms) + r;
  const size_t step = (size_t)nOut * (size_t)(3 * localAtoms);
  int y = 0;
  for (; y + 8 <= nSplit; y += 8) {          // eight loads in flight, summed in the fixed order
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldcg(q0 + (size_t)(y + u) * step);
#pragma unroll
    for (int u = 0; u < 8; u++) s += v[u];
  }
  for (; y < nSplit; y++) s += __ldcg(q0 + (size_t)y * step);
  for (y = 0; y < nSplit; y++) q0[(size_t)y * step] = 0.0;
  forces[(size_t)(3 * nOut) * (beg + r / 3) + 3 * k + r % 3] = s;
}
