// acsf_force.cuh -- fused analytic forces: F_f -= sum_i sum_a (dE/dG_ia) dG_ia/dR_f.
//
// Replaces TAcsf_calculatePrime -> iGeoAcsfGrad -> gFuncGrad (lib_descriptors/acsf.F90:643-717,
// 870-939, 1496-1697), applyZscorePrime (:515-536) and the contraction of
// forceAnalysis_analytical (lib_analysis/forces.F90:400-413).  The reference materialises a
// dense [3,F,N,N] tensor per structure and a full Jacobian per atom; here dE/dG comes from one
// reverse sweep of the subnetwork (k_mlp_bwd<MODE 1>) and the derivative of every pair term
// is contracted with it on the fly: per pair and ladder only S0 = sum_f D_f 2^(1-xi_f) b^xi_f
// and S1 = sum_f D_f 2^(1-xi_f) xi_f lam b^(xi_f - 1) are formed, then applied to the three
// geometric directions (Appendix C of SURVEY.md).  One warp per central atom; per-neighbour
// force accumulators live in shared memory, the scatter to atoms uses FP64 red.global.add.
// Note: for cells with an edge < 2*rc the reference overwrites instead
// of summing contributions of repeated periodic images (acsf.F90:918); we compute the true
// gradient, parity is defined for cells with every edge >= 2*rc (all reference goldens).
#pragma once
#include "acsf.cuh"

struct ForceSmem {
  WarpSmem w;          // w.dx..dz hold UNIT vectors here
  double *dE;          // d/dr of fc*exp(-eta r^2)
  double *fx, *fy, *fz;
};

__host__ __device__ inline size_t force_warp_smem_bytes(int cap, int F) {
  return acsf_warp_smem_bytes(cap, F, 0) + (size_t)cap * 4 * sizeof(double);
}

__device__ __forceinline__ double dcutoff_fn(double rr, double qq, double rc, double invrc) {
  // dfCutoffWoCheck, deriv = 1 (acsf.F90:1253-1254): 0.5 q q (pi/rc) cos(pi r/rc + pi/2)
  return -0.5 * qq * (3.14159265358979323846 * invrc) * sinpi(rr * invrc);
}

template <int PATH>
__global__ void __launch_bounds__(128)
k_acsf_force(int nSplit, GeomArgs geo, int nExt, const double *__restrict__ ext, AcsfTables tab, int cap, int capC,
             const double *__restrict__ dEdG, int nOut, const double *__restrict__ zprec,
             double *__restrict__ forces, int *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int kt = blockIdx.z;                        // target index
  CtaGeom cg;
  unsigned char *wbase;
  if (!acsf_cta_prologue<PATH>(geo, nSplit, tab.rcMax, capC, smem_raw, flags, cg, wbase)) return;
  const int a0 = cg.a0, a1 = cg.a1;
  unsigned char *base = wbase + (size_t)wib * force_warp_smem_bytes(cap, tab.F);
  WarpSmem w = carve_warp_smem(base, cap, tab.F);
  double *extra = (double *)(base + acsf_warp_smem_bytes(cap, tab.F, 0));
  double *dE = extra, *fx = extra + cap, *fy = extra + 2 * cap, *fz = extra + 3 * cap;
  for (int slot = a0 + wib; slot < a1; slot += nw) {
  const CRec me = central_atom<PATH>(cg, slot);
  const int i = me.idx;
  const int n = gather_neighbors<PATH>(me, cg, tab, cap, w);
  if (n < 0) { if (lane == 0) atomicMax(&flags[1], -n); continue; }
  for (int t = lane; t < n; t += 32) {
    const double ri = w.rinv[t];
    w.dx[t] *= ri; w.dy[t] *= ri; w.dz[t] *= ri;    // unit vectors (acsf.F90:1565)
    fx[t] = 0.0; fy[t] = 0.0; fz[t] = 0.0;
  }
  // dE/dG of this atom and target (z-score: derivatives / sigma, acsf.F90:528-534)
  for (int a = lane; a < tab.F; a += 32) {
    double d = dEdG[((size_t)nOut * i + kt) * tab.F + a];
    if (zprec) { const double sg = zprec[tab.F + a]; if (!(sg < 1e-08)) d /= sg; }
    w.outv[a] = d;
  }
  __syncwarp();

  // ---------------- radial (acsf.F90:1527-1557) ----------------
  for (int g = 0; g < tab.nRadialGroups; g++) {
    const RadialGroup G = tab.rgroups[g];
    const NbList l = make_list(tab, w, G.code, n);
    const int nl = l.n0 + l.n1;
    const double qi = G.atomId > 0 ? ext[(size_t)nExt * i + G.atomId - 1] : 1.0;
    const double invrc = 1.0 / G.rc;
    for (int t = lane; t < nl; t += 32) {
      const int a = list_at(l, t);
      const double rr = w.r[a];
      if (rr > G.rc) continue;
      const double qj = G.atomId > 0 ? ext[(size_t)nExt * w.idx[a] + G.atomId - 1] : 1.0;
      const double fc = cutoff_fn(rr, qi * qj, invrc), dfc = dcutoff_fn(rr, qi * qj, G.rc, invrc);
      double s = 0.0;
      for (int f = 0; f < G.fCnt; f++) {
        const double D = w.outv[tab.rfeat[G.fBeg + f]];
        const double p1 = tab.rp1[G.fBeg + f], p2 = tab.rp2[G.fBeg + f];
        if (G.type == FNETGPU_G1) s += D * dfc;
        else if (G.type == FNETGPU_G2) {
          const double d = rr - p2, e = p1 * d * d;
          s += D * (dfc - fc * 2.0 * p1 * d) * (e < 700.0 ? fnet_exp_tab(-e, cg.ftab) : 0.0);
        }
        else { double sn, cs; sincos(p1 * rr, &sn, &cs); s += D * (cs * dfc - sn * fc * p1); }
      }
      // each slot is owned by exactly one lane inside this loop, but lists of different groups
      // overlap across iterations only sequentially -> plain adds are safe here
      fx[a] += s * w.dx[a]; fy[a] += s * w.dy[a]; fz[a] += s * w.dz[a];
    }
    __syncwarp();
  }

  // ---------------- angular (acsf.F90:1559-1668) ----------------
  for (int pi_ = 0; pi_ < tab.nAngularPasses; pi_++) {
    const AngularPass *__restrict__ P = &tab.apasses[pi_];
    const int type = P->type, same = P->same, atomId = P->atomId, nSlots = P->nSlots;
    const double rc = P->rc, eta = P->eta, invrc = 1.0 / rc;
    const NbList l1 = make_list(tab, w, P->code1, n);
    const NbList l2 = same ? l1 : make_list(tab, w, P->code2, n);
    const int n1 = l1.n0 + l1.n1, n2 = l2.n0 + l2.n1;
    const double qi = atomId > 0 ? ext[(size_t)nExt * i + atomId - 1] : 1.0;
    for (int t = lane; t < n; t += 32) {
      const double rr = w.r[t];
      const double qj = atomId > 0 ? ext[(size_t)nExt * w.idx[t] + atomId - 1] : 1.0;
      w.qv[t] = qj;
      if (rr > rc) { w.fcE[t] = 0.0; dE[t] = 0.0; }
      else {
        const double fc = cutoff_fn(rr, qi * qj, invrc), dfc = dcutoff_fn(rr, qi * qj, rc, invrc);
        const double ex = fnet_exp_tab(-eta * rr * rr, cg.ftab);
        w.fcE[t] = fc * ex;
        dE[t] = (dfc - 2.0 * eta * rr * fc) * ex;   // acsf.F90:1577-1578
      }
    }
    __syncwarp();
    double lam[FNET_SLOTS], xi0[FNET_SLOTS], dxi[FNET_SLOTS];
    int cnt[FNET_SLOTS];
    double c0[FNET_SLOTS][FNET_LADDER], c1[FNET_SLOTS][FNET_LADDER];   // D*pref and D*pref*xi*lam
#pragma unroll
    for (int s = 0; s < FNET_SLOTS; s++) {
      lam[s] = s < nSlots ? P->slot[s].lam : 0.0;
      xi0[s] = s < nSlots ? P->slot[s].xi0 : 0.0;
      dxi[s] = s < nSlots ? P->slot[s].dxi : 0.0;
      cnt[s] = s < nSlots ? P->slot[s].count : 0;
#pragma unroll
      for (int f = 0; f < FNET_LADDER; f++) {
        const bool on = s < nSlots && f < cnt[s];
        const double D = on ? w.outv[P->slot[s].feat[f]] * P->slot[s].pref[f] : 0.0;
        c0[s][f] = D;
        c1[s][f] = on ? D * P->slot[s].xi[f] * lam[s] : 0.0;
      }
    }
    {
      const int nPairs = same ? (n1 * (n1 + 1)) >> 1 : n1 * n2;
      const int W = same ? (n1 | 1) : n2;
      const float invW = W > 0 ? 1.0f / (float)W : 0.0f;
      // Lanes of one sweep hold consecutive pairs, i.e. runs of pairs with the same first neighbour a
      // (row j of the pair walk): their contributions to a are combined by a segmented warp
      // reduction and added once per run -- a plain atomicAdd per pair would serialise up to 32
      // same-address shared-memory atomics.  The second neighbour b differs within a run.
      for (int pbase = 0; pbase < nPairs; pbase += 32) {
        const int p = pbase + lane;
        const bool valid = p < nPairs;
        int j = 0, k = 0;
        if (valid) pair_decode(p, same, n1, W, invW, j, k);
        const int a = list_at(l1, j), b = list_at(l2, k);
        const double Ea = w.fcE[a], Eb = w.fcE[b];
        double wgt = (same && a != b) ? 2.0 : 1.0;
        double H = 1.0, dH = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
        double gax = 0.0, gay = 0.0, gaz = 0.0;
        bool on = valid && (Ea != 0.0 || dE[a] != 0.0) && (Eb != 0.0 || dE[b] != 0.0);
        if (type == FNETGPU_G4 && on) {
          // third leg r_ab (acsf.F90:1603-1611): d_a - d_b in Cartesian = r_a u_a - r_b u_b
          vx = w.r[a] * w.dx[a] - w.r[b] * w.dx[b];
          vy = w.r[a] * w.dy[a] - w.r[b] * w.dy[b];
          vz = w.r[a] * w.dz[a] - w.r[b] * w.dz[b];
          const double d2 = vx * vx + vy * vy + vz * vz;
          const double dab = sqrt(d2);
          if (dab > rc) on = false;
          else {
            const double qq = w.qv[a] * w.qv[b];
            const double fc = cutoff_fn(dab, qq, invrc), ex = fnet_exp_tab(-eta * d2, cg.ftab);
            H = fc * ex;
            if (a != b) {
              dH = (dcutoff_fn(dab, qq, rc, invrc) - 2.0 * eta * dab * fc) * ex;
              const double inv = 1.0 / dab;
              vx *= inv; vy *= inv; vz *= inv;
            }
          }
        }
        if (on) {
          const double c = w.dx[a] * w.dx[b] + w.dy[a] * w.dy[b] + w.dz[a] * w.dz[b];   // acsf.F90:1591
          double S0 = 0.0, S1 = 0.0, L = 0.0, bb = 0.0;
#pragma unroll
          for (int s = 0; s < FNET_SLOTS; s++) {
            if (cnt[s] > 0) {
              if (s == 0 || lam[s] != lam[s - 1]) { bb = fmax(1.0 + lam[s] * c, 0.0); L = fnet_log_tab(bb, cg.ftab); }
              double pm, q;
              ladder_init(bb, L, xi0[s] - 1.0, dxi[s], pm, q, cg.ftab);     // b^(xi-1)
              // sum_f c[f] pm q^f = pm * P(q): two Horner chains instead of a serial pm *= q ladder
              double P1 = c1[s][FNET_LADDER - 1], P0 = c0[s][FNET_LADDER - 1];
#pragma unroll
              for (int f = FNET_LADDER - 2; f >= 0; f--) { P1 = fma(P1, q, c1[s][f]); P0 = fma(P0, q, c0[s][f]); }
              S1 = fma(P1, pm, S1);
              S0 = fma(P0, pm * bb, S0);
            }
          }
          // g_a = w [ Eb H ( S1 Ea (u_b - c u_a)/r_a + S0 dEa u_a ) + S0 Ea Eb dH u_ab ]
          const double ta = wgt * Eb * H * S1 * Ea * w.rinv[a];
          const double sa = wgt * Eb * H * S0 * dE[a] - ta * c;
          const double tb = wgt * Ea * H * S1 * Eb * w.rinv[b];
          const double sb = wgt * Ea * H * S0 * dE[b] - tb * c;
          const double th = wgt * S0 * Ea * Eb * dH;
          gax = ta * w.dx[b] + sa * w.dx[a] + th * vx;
          gay = ta * w.dy[b] + sa * w.dy[a] + th * vy;
          gaz = ta * w.dz[b] + sa * w.dz[a] + th * vz;
          atomicAdd(&fx[b], tb * w.dx[a] + sb * w.dx[b] - th * vx);
          atomicAdd(&fy[b], tb * w.dy[a] + sb * w.dy[b] - th * vy);
          atomicAdd(&fz[b], tb * w.dz[a] + sb * w.dz[b] - th * vz);
        }
        const int key = valid ? a : -1 - lane;           // runs of equal a are contiguous in the pair walk
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int ko = __shfl_down_sync(0xffffffffu, key, off);
          const double x2 = __shfl_down_sync(0xffffffffu, gax, off), y2 = __shfl_down_sync(0xffffffffu, gay, off),
                       z2 = __shfl_down_sync(0xffffffffu, gaz, off);
          if (lane + off < 32 && ko == key) { gax += x2; gay += y2; gaz += z2; }
        }
        const int kp = __shfl_up_sync(0xffffffffu, key, 1);
        if (valid && (lane == 0 || kp != key) && (gax != 0.0 || gay != 0.0 || gaz != 0.0)) {
          atomicAdd(&fx[a], gax); atomicAdd(&fy[a], gay); atomicAdd(&fz[a], gaz);
        }
      }
    }
    __syncwarp();
  }

  // ---------------- scatter: neighbours get -g, the central atom +sum(g) ----------------
  double sx = 0.0, sy = 0.0, sz = 0.0;
  const int stride = 3 * nOut;
  for (int t = lane; t < n; t += 32) {
    const double gx = fx[t], gy = fy[t], gz = fz[t];
    if (gx != 0.0 || gy != 0.0 || gz != 0.0) {
      double *ff = forces + (size_t)stride * w.idx[t] + 3 * kt;
      atomicAdd(ff, -gx); atomicAdd(ff + 1, -gy); atomicAdd(ff + 2, -gz);
      sx += gx; sy += gy; sz += gz;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) {
    double *ff = forces + (size_t)stride * i + 3 * kt;
    atomicAdd(ff, sx); atomicAdd(ff + 1, sy); atomicAdd(ff + 2, sz);
  }
  __syncwarp();
  }   // slot loop
}
