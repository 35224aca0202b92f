/*
 * refcpu.c -- CPU ORACLE for the Fortnet hot path (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C, FP64 restatement of the reference algorithm (vanderhe/fortnet v0.7.4),
 * keeping the reference's own algorithmic structure (neighbour list rebuilt by brute
 * force for every (atom, G-function); ordered pair loops with acos->cos and pow; per-atom
 * fprop/bprop; dense [3,F,N,N] derivative tensor and the 5-deep force loop).  It is
 *   (a) the checker the GPU parity tests compare against, and
 *   (b) the CPU baseline bench.py times ("kind": "port") -- the Fortran binary cannot be
 *       built in this image (no gfortran / HDF5 / MPI).
 * Parity is PINNED: tests/test_oracle_golden.py replays the reference's own regression
 * goldens (tests/golden/, made by tests/golden/make_golden.py from
 * /root/reference/test/prog/fortnet) through this file at rtol 1e-9 / atol 1e-10.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path (fortnet_b200/) never does.
 *
 * Each function cites the reference file:line (relative to
 * /root/reference/prog/fortnet/) it follows.
 *
 * Array conventions (same as the C ABI in include/fnetgpu.h):
 *   coords[3*i+c]            Cartesian, Bohr, atom i of the concatenated dataset
 *   latvecs[9*s+3*k+c]       component c of lattice vector k of structure s (= latVecs(c,k))
 *   feats[F*i+a]             feature a of atom i             (= array(a,i) per structure)
 *   ext[nExt*i+e]            external feature e of atom i    (= extFeatures(e,i))
 *   wb[nTot*sp + p]          serialised parameters of species sp (TDerivs_serialized order)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846264338327950288

enum { G1 = 1, G2 = 2, G3 = 3, G4 = 4, G5 = 5 };

typedef struct {
  int type;        /* 1..5 */
  double rcut, kappa, rs, eta, lambda, xi;
  int atomId;      /* 0 = none, else 1-based row of ext */
  int z1, z2;      /* atomicNumbers(2) */
} GFunc;

/* ------------------------------------------------------------------------------------------
 * lib_common/parallel.F90:23-56  getStartAndEndIndex (0-based iProc, 1-based inclusive range)
 * ---------------------------------------------------------------------------------------- */
void fnet_oracle_start_end(int nSystems, int nProcs, int iProc, int *iStart, int *iEnd) {
  int splitSize = nSystems / nProcs;
  *iStart = iProc * splitSize + 1;
  *iEnd = *iStart + splitSize - 1;
  int offset = nProcs - nSystems % nProcs;
  if (iProc + 1 > offset) {
    *iStart = *iStart + iProc - offset;
    *iEnd = *iEnd + iProc - offset + 1;
  }
}

/* ------------------------------------------------------------------------------------------
 * Geometry helper: lib_dftbp/simplealgebra.F90:90-125 invert33 + the transposition done in
 * lib_io/fnetdata.F90:1159-1161 -> recVecs2p(:,k) = row k of inverse(latVecs).
 * L[c][k] = latvecs[3*k+c].
 * ---------------------------------------------------------------------------------------- */
static void rec_vecs(const double *lv, double rec[3][3] /* rec[k][c] */) {
  double o[3][3]; /* o[r][c] = orig(r+1,c+1) = latVecs(r,c) = lv[3*c + r] */
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o[r][c] = lv[3 * c + r];
  double det = o[0][0] * (o[1][1] * o[2][2] - o[1][2] * o[2][1])
             - o[0][1] * (o[1][0] * o[2][2] - o[1][2] * o[2][0])
             + o[0][2] * (o[1][0] * o[2][1] - o[1][1] * o[2][0]);
  double inv[3][3];
  inv[0][0] = -o[1][2] * o[2][1] + o[1][1] * o[2][2];
  inv[1][0] =  o[1][2] * o[2][0] - o[1][0] * o[2][2];
  inv[2][0] = -o[1][1] * o[2][0] + o[1][0] * o[2][1];
  inv[0][1] =  o[0][2] * o[2][1] - o[0][1] * o[2][2];
  inv[1][1] = -o[0][2] * o[2][0] + o[0][0] * o[2][2];
  inv[2][1] =  o[0][1] * o[2][0] - o[0][0] * o[2][1];
  inv[0][2] = -o[0][2] * o[1][1] + o[0][1] * o[1][2];
  inv[1][2] =  o[0][2] * o[1][0] - o[0][0] * o[1][2];
  inv[2][2] = -o[0][1] * o[1][0] + o[0][0] * o[1][1];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) inv[r][c] /= det;
  /* recVecs2p = transpose(inv): recVecs2p(:,k) = inv(k,:) */
  for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) rec[k][c] = inv[k][c];
}

/* growable neighbour list */
typedef struct {
  int n, cap;
  double *dist;    /* n   */
  double *xyz;     /* 3n  absolute image coordinates */
  int *idx;        /* n   index into the (reduced) geometry, later mapped to the full one */
} NList;

static void nl_init(NList *l) { l->n = 0; l->cap = 0; l->dist = NULL; l->xyz = NULL; l->idx = NULL; }
static void nl_free(NList *l) { free(l->dist); free(l->xyz); free(l->idx); nl_init(l); }
static void nl_push(NList *l, double d, const double *x, int idx) {
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 64;
    l->dist = (double *)realloc(l->dist, sizeof(double) * l->cap);
    l->xyz = (double *)realloc(l->xyz, sizeof(double) * 3 * l->cap);
    l->idx = (int *)realloc(l->idx, sizeof(int) * l->cap);
  }
  l->dist[l->n] = d;
  l->xyz[3 * l->n] = x[0]; l->xyz[3 * l->n + 1] = x[1]; l->xyz[3 * l->n + 2] = x[2];
  l->idx[l->n] = idx;
  l->n++;
}
static void nl_copy(NList *dst, const NList *src) {
  dst->n = 0;
  for (int i = 0; i < src->n; i++) nl_push(dst, src->dist[i], src->xyz + 3 * i, src->idx[i]);
}

/* ------------------------------------------------------------------------------------------
 * lib_descriptors/acsf.F90:1070-1138 buildNeighborlist
 *   -> lib_dftbp/dynneighlist.F90:182-320 (iteration order, dist2 <= cutoff2 test :294,
 *      sqrt :311) and lib_dftbp/latpointiter.F90:62-117,204-253 (image box, x slowest /
 *      z fastest, origin skipped).
 * geometry = nAt atoms with coordinates xyz[3*a+c]; iAtom 0-based.
 * ---------------------------------------------------------------------------------------- */
static void build_neighborlist(int nAt, const double *xyz, int periodic, const double *lv,
                               double rcut, int iAtom, NList *out) {
  out->n = 0;
  const double cutoff2 = rcut * rcut;
  const double *ci = xyz + 3 * iAtom;
  /* central cell: atoms iAtom+1 ... wrapping, self skipped (dynneighlist.F90:216-218,270-276) */
  for (int it = 1; it < nAt; it++) {
    int a2 = (iAtom + it) % nAt;
    const double *cj = xyz + 3 * a2;
    double d2 = (ci[0] - cj[0]) * (ci[0] - cj[0]) + (ci[1] - cj[1]) * (ci[1] - cj[1])
              + (ci[2] - cj[2]) * (ci[2] - cj[2]);
    if (d2 <= cutoff2) nl_push(out, sqrt(d2), cj, a2);
  }
  if (!periodic) return;
  double rec[3][3];
  rec_vecs(lv, rec);
  int lo[3], hi[3];
  for (int k = 0; k < 3; k++) { /* latpointiter.F90:247-251 with posExt = negExt = 1 */
    int iTmp = (int)floor(rcut * sqrt(rec[k][0] * rec[k][0] + rec[k][1] * rec[k][1] + rec[k][2] * rec[k][2]));
    lo[k] = -(iTmp + 1);
    hi[k] = iTmp + 1;
  }
  for (int p0 = lo[0]; p0 <= hi[0]; p0++)
    for (int p1 = lo[1]; p1 <= hi[1]; p1++)
      for (int p2 = lo[2]; p2 <= hi[2]; p2++) {
        if (p0 == 0 && p1 == 0 && p2 == 0) continue;
        double cell[3];
        for (int c = 0; c < 3; c++) /* cellVec = matmul(latVecs, point), dynneighlist.F90:280 */
          cell[c] = lv[c] * (double)p0 + lv[3 + c] * (double)p1 + lv[6 + c] * (double)p2;
        for (int it = 0; it < nAt; it++) { /* atoms iAtom, iAtom+1, ... wrapping (:287-303) */
          int a2 = (iAtom + it) % nAt;
          double cj[3] = {xyz[3 * a2] + cell[0], xyz[3 * a2 + 1] + cell[1], xyz[3 * a2 + 2] + cell[2]};
          double d2 = (ci[0] - cj[0]) * (ci[0] - cj[0]) + (ci[1] - cj[1]) * (ci[1] - cj[1])
                    + (ci[2] - cj[2]) * (ci[2] - cj[2]);
          if (d2 <= cutoff2) nl_push(out, sqrt(d2), cj, a2);
        }
      }
}

/* ------------------------------------------------------------------------------------------
 * acsf.F90:720-793 reduceGeometrySpecies + :942-1066 buildGFunctionNeighborlists.
 * Produces list1/list2 (indices mapped back to the FULL geometry, as atomIndices?_ are,
 * :1021-1022,1061-1062) and the per-neighbour atom-id prefactors q1/q2.
 * ---------------------------------------------------------------------------------------- */
typedef struct { NList l1, l2; double *q1, *q2; int qcap1, qcap2; double *rx; int *keep; int rcap; } Work;

static void work_init(Work *w) { nl_init(&w->l1); nl_init(&w->l2); w->q1 = w->q2 = NULL; w->qcap1 = w->qcap2 = 0; w->rx = NULL; w->keep = NULL; w->rcap = 0; }
static void work_free(Work *w) { nl_free(&w->l1); nl_free(&w->l2); free(w->q1); free(w->q2); free(w->rx); free(w->keep); }

static void reduced_list(Work *w, int nAt, const double *xyz, const int *atnum, int periodic,
                         const double *lv, int iAtom, int Z, double rcut, NList *out) {
  if (w->rcap < nAt) { w->rcap = nAt; w->rx = (double *)realloc(w->rx, sizeof(double) * 3 * nAt); w->keep = (int *)realloc(w->keep, sizeof(int) * nAt); }
  /* geo1 = geo reduced to species Z plus iAtom itself (deep copy, :754-777) */
  int m = 0, iOut = -1;
  for (int a = 0; a < nAt; a++) {
    if (a == iAtom) { iOut = m; w->keep[m] = a; memcpy(w->rx + 3 * m, xyz + 3 * a, 3 * sizeof(double)); m++; }
    else if (atnum[a] == Z) { w->keep[m] = a; memcpy(w->rx + 3 * m, xyz + 3 * a, 3 * sizeof(double)); m++; }
  }
  build_neighborlist(m, w->rx, periodic, lv, rcut, iOut, out);
  for (int i = 0; i < out->n; i++) out->idx[i] = w->keep[out->idx[i]];
}

static void set_q(double **q, int *cap, const NList *l, const GFunc *g, int nExt, const double *ext) {
  if (*cap < l->n) { *cap = l->n + 64; *q = (double *)realloc(*q, sizeof(double) * (*cap)); }
  for (int i = 0; i < l->n; i++)
    (*q)[i] = (g->atomId > 0) ? ext[(size_t)nExt * l->idx[i] + (g->atomId - 1)] : 1.0;
}

static void build_gfunc_lists(Work *w, int nAt, const double *xyz, const int *atnum, int periodic,
                              const double *lv, int nExt, const double *ext, const GFunc *g, int iAtom) {
  int resolved = !(g->z1 == 0 && g->z2 == 0); /* :992-996 */
  if (!resolved) {
    if (w->rcap < nAt) { w->rcap = nAt; w->rx = (double *)realloc(w->rx, sizeof(double) * 3 * nAt); w->keep = (int *)realloc(w->keep, sizeof(int) * nAt); }
    memcpy(w->rx, xyz, sizeof(double) * 3 * nAt); /* geo1 = geo (deep copy, :999) */
    build_neighborlist(nAt, w->rx, periodic, lv, g->rcut, iAtom, &w->l1);
    nl_copy(&w->l2, &w->l1);
  } else {
    reduced_list(w, nAt, xyz, atnum, periodic, lv, iAtom, g->z1, g->rcut, &w->l1);
    if (g->type >= G4) reduced_list(w, nAt, xyz, atnum, periodic, lv, iAtom, g->z2, g->rcut, &w->l2);
    else nl_copy(&w->l2, &w->l1);
  }
  set_q(&w->q1, &w->qcap1, &w->l1, g, nExt, ext);
  set_q(&w->q2, &w->qcap2, &w->l2, g, nExt, ext);
}

/* acsf.F90:1181-1256 cutoff functions */
static double cutoff_chk(double rr, double a1, double a2, double rcut) {
  if (rr > rcut) return 0.0;
  return 0.5 * a1 * a2 * (cos(PI * rr / rcut) + 1.0);
}
static double cutoff_nochk(double rr, double a1, double a2, double rcut) {
  return 0.5 * a1 * a2 * (cos(PI * rr / rcut) + 1.0);
}
static double dcutoff_nochk(double rr, double a1, double a2, double rcut) { /* deriv = 1 */
  return 0.5 * a1 * a2 * (PI / rcut) * cos((PI * rr / rcut) + 0.5 * PI);
}
/* acsf.F90:1159-1176 theta */
static double theta(const double *ci, const double *c1, const double *c2, double d1, double d2) {
  double dot = (c1[0] - ci[0]) * (c2[0] - ci[0]) + (c1[1] - ci[1]) * (c2[1] - ci[1]) + (c1[2] - ci[2]) * (c2[2] - ci[2]);
  return acos(dot / (d1 * d2 + 1e-13));
}

/* acsf.F90:1287-1492  g1..g5 values */
static double gfunc_value(const GFunc *g, const double *ci, double qi, const Work *w) {
  const NList *l1 = &w->l1, *l2 = &w->l2;
  double s = 0.0;
  switch (g->type) {
  case G1:
    for (int j = 0; j < l1->n; j++) s += cutoff_chk(l1->dist[j], qi, w->q1[j], g->rcut);
    return s;
  case G2:
    for (int j = 0; j < l1->n; j++) {
      double r = l1->dist[j];
      s += exp(-g->eta * (r - g->rs) * (r - g->rs)) * cutoff_chk(r, qi, w->q1[j], g->rcut);
    }
    return s;
  case G3:
    for (int j = 0; j < l1->n; j++) {
      double r = l1->dist[j];
      s += cos(g->kappa * r) * cutoff_chk(r, qi, w->q1[j], g->rcut);
    }
    return s;
  case G4:
    if (l1->n == 0 && l2->n == 0) return 0.0;
    for (int j = 0; j < l1->n; j++)
      for (int k = 0; k < l2->n; k++) {
        const double *cj = l1->xyz + 3 * j, *ck = l2->xyz + 3 * k;
        double djk = sqrt((cj[0] - ck[0]) * (cj[0] - ck[0]) + (cj[1] - ck[1]) * (cj[1] - ck[1]) + (cj[2] - ck[2]) * (cj[2] - ck[2]));
        double rj = l1->dist[j], rk = l2->dist[k];
        s += pow(1.0 + g->lambda * cos(theta(ci, cj, ck, rj, rk)), g->xi)
           * exp(-g->eta * (rj * rj + rk * rk + djk * djk))
           * cutoff_nochk(rj, qi, w->q1[j], g->rcut) * cutoff_nochk(rk, qi, w->q2[k], g->rcut)
           * cutoff_chk(djk, w->q1[j], w->q2[k], g->rcut);
      }
    return s * pow(2.0, 1.0 - g->xi);
  case G5:
    if (l1->n == 0 && l2->n == 0) return 0.0;
    for (int j = 0; j < l1->n; j++)
      for (int k = 0; k < l2->n; k++) {
        const double *cj = l1->xyz + 3 * j, *ck = l2->xyz + 3 * k;
        double rj = l1->dist[j], rk = l2->dist[k];
        s += pow(1.0 + g->lambda * cos(theta(ci, cj, ck, rj, rk)), g->xi)
           * exp(-g->eta * (rj * rj + rk * rk))
           * cutoff_nochk(rj, qi, w->q1[j], g->rcut) * cutoff_nochk(rk, qi, w->q2[k], g->rcut);
      }
    return s * pow(2.0, 1.0 - g->xi);
  }
  return 0.0;
}

static void unpack_funcs(int F, const int *ftype, const double *rcut, const double *kappa,
                         const double *rs, const double *eta, const double *lambda, const double *xi,
                         const int *atomid, const int *atomicnumbers, GFunc *g) {
  for (int a = 0; a < F; a++) {
    g[a].type = ftype[a]; g[a].rcut = rcut[a]; g[a].kappa = kappa[a]; g[a].rs = rs[a];
    g[a].eta = eta[a]; g[a].lambda = lambda[a]; g[a].xi = xi[a]; g[a].atomId = atomid[a];
    g[a].z1 = atomicnumbers[2 * a]; g[a].z2 = atomicnumbers[2 * a + 1];
  }
}

/* ------------------------------------------------------------------------------------------
 * acsf.F90:797-866 iGeoAcsf for structures [sBeg,sEnd) ; raw values (no z-score).
 * nThreads > 1 = "MPI-style": the structure range is split with getStartAndEndIndex and
 * every worker handles its contiguous block (acsf.F90:597-614).
 * ---------------------------------------------------------------------------------------- */
void fnet_oracle_acsf(int nStruct, const int *offsets, const double *coords, const int *periodic,
                      const double *latvecs, const int *atnum, int nExt, const double *ext,
                      int F, const int *ftype, const double *rcut, const double *kappa,
                      const double *rs, const double *eta, const double *lambda, const double *xi,
                      const int *atomid, const int *atomicnumbers, double *out, int nThreads) {
  GFunc *g = (GFunc *)malloc(sizeof(GFunc) * (F > 0 ? F : 1));
  unpack_funcs(F, ftype, rcut, kappa, rs, eta, lambda, xi, atomid, atomicnumbers, g);
  if (nThreads < 1) nThreads = 1;
#pragma omp parallel for num_threads(nThreads) schedule(static, 1)
  for (int p = 0; p < nThreads; p++) {
    int s0, s1;
    fnet_oracle_start_end(nStruct, nThreads, p, &s0, &s1);
    Work w; work_init(&w);
    for (int s = s0 - 1; s < s1; s++) {
      int o = offsets[s], nAt = offsets[s + 1] - offsets[s];
      const double *xyz = coords + 3 * (size_t)o;
      const double *e = ext ? ext + (size_t)nExt * o : NULL;
      for (int i = 0; i < nAt; i++)
        for (int a = 0; a < F; a++) {
          build_gfunc_lists(&w, nAt, xyz, atnum + o, periodic[s], latvecs + 9 * (size_t)s, nExt, e, &g[a], i);
          double qi = (g[a].atomId > 0) ? e[(size_t)nExt * i + (g[a].atomId - 1)] : 1.0; /* :836-840 */
          out[(size_t)F * (o + i) + a] = gfunc_value(&g[a], xyz + 3 * i, qi, &w);
        }
    }
    work_free(&w);
  }
  free(g);
}

/* acsf.F90:445-486 getMeansAndVariances (two-pass, integer dataset weights, population sigma) */
void fnet_oracle_zscore_stats(int nStruct, const int *offsets, int F, const double *vals,
                              const int *weights, double *means, double *sigmas) {
  long nTot = 0;
  for (int a = 0; a < F; a++) { means[a] = 0.0; sigmas[a] = 0.0; }
  for (int s = 0; s < nStruct; s++)
    for (int i = offsets[s]; i < offsets[s + 1]; i++) {
      nTot += weights[s];
      for (int a = 0; a < F; a++) means[a] += (double)weights[s] * vals[(size_t)F * i + a];
    }
  for (int a = 0; a < F; a++) means[a] /= (double)nTot;
  for (int s = 0; s < nStruct; s++)
    for (int i = offsets[s]; i < offsets[s + 1]; i++)
      for (int a = 0; a < F; a++) {
        double d = vals[(size_t)F * i + a] - means[a];
        sigmas[a] += (double)weights[s] * d * d;
      }
  for (int a = 0; a < F; a++) sigmas[a] = sqrt(sigmas[a] / (double)nTot);
}

/* acsf.F90:490-511 applyZscore */
void fnet_oracle_zscore_apply(int N, int F, double *vals, const double *means, const double *sigmas) {
  for (int a = 0; a < F; a++) {
    if (sigmas[a] < 1e-08) continue;
    for (int i = 0; i < N; i++) vals[(size_t)F * i + a] = (vals[(size_t)F * i + a] - means[a]) / sigmas[a];
  }
}

/* ------------------------------------------------------------------------------------------
 * acsf.F90:1496-1670 gFuncGrad + :1674-1697 T_matr.   grad[3*n + c], n over list1.
 * (la,qa) play the role of neighCoords1/neighDists1/atomIds1, (lb,qb) of ...2.
 * ---------------------------------------------------------------------------------------- */
static void t_matr(const double *u, double nrm, double T[3][3]) {
  for (int i = 0; i < 3; i++) {
    T[i][i] = 1.0 - u[i] * u[i];
    for (int j = i + 1; j < 3; j++) { T[i][j] = -u[i] * u[j]; T[j][i] = T[i][j]; }
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] /= nrm;
}

static void gfunc_grad(const GFunc *g, const double *ci, double qi, const NList *la, const double *qa,
                       const NList *lb, const double *qb, double *grad) {
  int n1 = la->n, n2 = lb->n;
  for (int i = 0; i < 3 * n1; i++) grad[i] = 0.0;
  if (g->type == G1) {
    for (int j = 0; j < n1; j++) {
      double r = la->dist[j], f = dcutoff_nochk(r, qi, qa[j], g->rcut);
      for (int c = 0; c < 3; c++) grad[3 * j + c] = f * (la->xyz[3 * j + c] - ci[c]) / r;
    }
    return;
  }
  if (g->type == G2) {
    for (int j = 0; j < n1; j++) {
      double r = la->dist[j];
      double f = (dcutoff_nochk(r, qi, qa[j], g->rcut) - cutoff_nochk(r, qi, qa[j], g->rcut) * 2.0 * g->eta * (r - g->rs))
               * exp(-g->eta * (r - g->rs) * (r - g->rs));
      for (int c = 0; c < 3; c++) grad[3 * j + c] = f * (la->xyz[3 * j + c] - ci[c]) / r;
    }
    return;
  }
  if (g->type == G3) {
    for (int j = 0; j < n1; j++) {
      double r = la->dist[j];
      double f = cos(g->kappa * r) * dcutoff_nochk(r, qi, qa[j], g->rcut)
               - sin(g->kappa * r) * cutoff_nochk(r, qi, qa[j], g->rcut) * g->kappa;
      for (int c = 0; c < 3; c++) grad[3 * j + c] = f * (la->xyz[3 * j + c] - ci[c]) / r;
    }
    return;
  }
  double *uv2 = (double *)malloc(sizeof(double) * 3 * (n2 + 1));
  double *sq2 = (double *)malloc(sizeof(double) * (n2 + 1));
  double *fc2 = (double *)calloc(n2 + 1, sizeof(double));
  for (int j = 0; j < n2; j++) { /* :1564-1569 */
    for (int c = 0; c < 3; c++) uv2[3 * j + c] = (lb->xyz[3 * j + c] - ci[c]) / lb->dist[j];
    sq2[j] = lb->dist[j] * lb->dist[j];
    if (lb->dist[j] > g->rcut) continue;
    fc2[j] = cutoff_nochk(lb->dist[j], qi, qb[j], g->rcut);
  }
  double T[3][3], tmp[3];
  if (g->z1 == g->z2) { /* :1573-1625 same neighbourhood */
    for (int k = 0; k < n1; k++) {
      t_matr(uv2 + 3 * k, la->dist[k], T);
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) T[a][b] *= g->lambda * g->xi;
      double dfc_ik = -2.0 * g->eta * la->dist[k] * fc2[k] + dcutoff_nochk(la->dist[k], qi, qa[k], g->rcut);
      double pp = 1.0;
      if (g->type == G4) pp = qa[k] * qb[k];
      double diag = pow(1.0 + g->lambda, g->xi) * fc2[k] * dfc_ik * exp(-2.0 * g->eta * sq2[k]) * pp;
      for (int c = 0; c < 3; c++) grad[3 * k + c] += diag * uv2[3 * k + c];
      for (int j = 0; j < n2; j++) {
        if (j == k) continue;
        double a_ijk = 1.0 + g->lambda * (uv2[3 * j] * uv2[3 * k] + uv2[3 * j + 1] * uv2[3 * k + 1] + uv2[3 * j + 2] * uv2[3 * k + 2]);
        for (int a = 0; a < 3; a++) tmp[a] = T[a][0] * uv2[3 * j] + T[a][1] * uv2[3 * j + 1] + T[a][2] * uv2[3 * j + 2];
        if (g->type == G5) {
          double pre = fc2[j] * exp(-g->eta * (sq2[k] + sq2[j])) * pow(a_ijk, g->xi - 1.0);
          for (int c = 0; c < 3; c++) grad[3 * k + c] += pre * (fc2[k] * tmp[c] + a_ijk * dfc_ik * uv2[3 * k + c]);
          continue;
        }
        double ujk[3] = {la->xyz[3 * k] - lb->xyz[3 * j], la->xyz[3 * k + 1] - lb->xyz[3 * j + 1], la->xyz[3 * k + 2] - lb->xyz[3 * j + 2]};
        double djk = sqrt(ujk[0] * ujk[0] + ujk[1] * ujk[1] + ujk[2] * ujk[2]);
        if (djk > g->rcut) continue;
        for (int c = 0; c < 3; c++) ujk[c] /= djk;
        double fc_jk = cutoff_nochk(djk, qa[k], qb[j], g->rcut);
        double dfc_jk = -2.0 * g->eta * djk * fc_jk + dcutoff_nochk(djk, qa[k], qb[j], g->rcut);
        double pre = fc2[j] * exp(-g->eta * (sq2[k] + sq2[j] + djk * djk)) * pow(a_ijk, g->xi - 1.0);
        for (int c = 0; c < 3; c++)
          grad[3 * k + c] += pre * (fc2[k] * fc_jk * tmp[c] + a_ijk * (fc_jk * dfc_ik * uv2[3 * k + c] + fc2[k] * dfc_jk * ujk[c]));
      }
    }
    double sc = pow(2.0, 2.0 - g->xi);
    for (int i = 0; i < 3 * n1; i++) grad[i] *= sc;
  } else { /* :1627-1668  Z1 /= Z2 */
    for (int k = 0; k < n1; k++) {
      double fc_ik = cutoff_nochk(la->dist[k], qi, qa[k], g->rcut);
      double dfc_ik = -2.0 * g->eta * la->dist[k] * fc_ik + dcutoff_nochk(la->dist[k], qi, qa[k], g->rcut);
      double uik[3];
      for (int c = 0; c < 3; c++) uik[c] = (la->xyz[3 * k + c] - ci[c]) / la->dist[k];
      t_matr(uik, la->dist[k], T);
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) T[a][b] *= g->lambda * g->xi;
      for (int j = 0; j < n2; j++) {
        double a_ijk = 1.0 + g->lambda * (uv2[3 * j] * uik[0] + uv2[3 * j + 1] * uik[1] + uv2[3 * j + 2] * uik[2]);
        for (int a = 0; a < 3; a++) tmp[a] = T[a][0] * uv2[3 * j] + T[a][1] * uv2[3 * j + 1] + T[a][2] * uv2[3 * j + 2];
        if (g->type == G5) {
          double pre = fc2[j] * exp(-g->eta * (la->dist[k] * la->dist[k] + sq2[j])) * pow(a_ijk, g->xi - 1.0);
          for (int c = 0; c < 3; c++) grad[3 * k + c] += pre * (fc_ik * tmp[c] + a_ijk * dfc_ik * uik[c]);
          continue;
        }
        double ujk[3] = {la->xyz[3 * k] - lb->xyz[3 * j], la->xyz[3 * k + 1] - lb->xyz[3 * j + 1], la->xyz[3 * k + 2] - lb->xyz[3 * j + 2]};
        double djk = sqrt(ujk[0] * ujk[0] + ujk[1] * ujk[1] + ujk[2] * ujk[2]);
        if (djk > g->rcut) continue;
        for (int c = 0; c < 3; c++) ujk[c] /= djk;
        double fc_jk = cutoff_nochk(djk, qa[k], qb[j], g->rcut);
        double dfc_jk = -2.0 * g->eta * djk * fc_jk + dcutoff_nochk(djk, qa[k], qb[j], g->rcut);
        double pre = fc2[j] * exp(-g->eta * (la->dist[k] * la->dist[k] + sq2[j] + djk * djk)) * pow(a_ijk, g->xi - 1.0);
        for (int c = 0; c < 3; c++)
          grad[3 * k + c] += pre * (fc_ik * fc_jk * tmp[c] + a_ijk * (fc_jk * dfc_ik * uik[c] + fc_ik * dfc_jk * ujk[c]));
      }
    }
    double sc = pow(2.0, 1.0 - g->xi);
    for (int i = 0; i < 3 * n1; i++) grad[i] *= sc;
  }
  free(uv2); free(sq2); free(fc2);
}

/* ------------------------------------------------------------------------------------------
 * acsf.F90:870-939 iGeoAcsfGrad for ONE structure: dense prime[c + 3*(a + F*(i + nAt*f))]
 * (= array(c,a,i,f), Fortran order), f = atom the derivative is taken with respect to.
 * The vector-subscript assignment at :918 keeps the LAST value when an atom index repeats
 * (several periodic images), and :922 sums the stored columns once per list entry -- both
 * restated literally; they only matter for cells with an edge < 2*rc.
 * sigmas != NULL applies acsf.F90:515-536 applyZscorePrime.
 * ---------------------------------------------------------------------------------------- */
void fnet_oracle_acsf_prime_struct(int nAt, const double *xyz, int periodic, const double *lv,
                                   const int *atnum, int nExt, const double *ext, int F,
                                   const int *ftype, const double *rcut, const double *kappa,
                                   const double *rs, const double *eta, const double *lambda,
                                   const double *xi, const int *atomid, const int *atomicnumbers,
                                   const double *sigmas, double *prime) {
  GFunc *g = (GFunc *)malloc(sizeof(GFunc) * (F > 0 ? F : 1));
  unpack_funcs(F, ftype, rcut, kappa, rs, eta, lambda, xi, atomid, atomicnumbers, g);
  memset(prime, 0, sizeof(double) * 3 * (size_t)F * nAt * nAt);
  Work w; work_init(&w);
  double *grad = NULL; int gcap = 0;
#define PR(c, a, i, f) prime[(c) + 3 * ((size_t)(a) + (size_t)F * ((size_t)(i) + (size_t)nAt * (f)))]
  for (int i = 0; i < nAt; i++)
    for (int a = 0; a < F; a++) {
      build_gfunc_lists(&w, nAt, xyz, atnum, periodic, lv, nExt, ext, &g[a], i);
      double qi = (g[a].atomId > 0) ? ext[(size_t)nExt * i + (g[a].atomId - 1)] : 1.0;
      int need = 3 * (w.l1.n > w.l2.n ? w.l1.n : w.l2.n) + 3;
      if (gcap < need) { gcap = need; grad = (double *)realloc(grad, sizeof(double) * gcap); }
      gfunc_grad(&g[a], xyz + 3 * i, qi, &w.l1, w.q1, &w.l2, w.q2, grad);
      for (int n = 0; n < w.l1.n; n++) for (int c = 0; c < 3; c++) PR(c, a, i, w.l1.idx[n]) = grad[3 * n + c];
      double sum[3] = {0, 0, 0};
      for (int n = 0; n < w.l1.n; n++) for (int c = 0; c < 3; c++) sum[c] += PR(c, a, i, w.l1.idx[n]);
      for (int c = 0; c < 3; c++) PR(c, a, i, i) = -sum[c];
      if (g[a].type >= G4 && g[a].z1 != g[a].z2) { /* :924-933 */
        gfunc_grad(&g[a], xyz + 3 * i, qi, &w.l2, w.q2, &w.l1, w.q1, grad);
        for (int n = 0; n < w.l2.n; n++) for (int c = 0; c < 3; c++) PR(c, a, i, w.l2.idx[n]) = grad[3 * n + c];
        double s2[3] = {0, 0, 0};
        for (int n = 0; n < w.l2.n; n++) for (int c = 0; c < 3; c++) s2[c] += PR(c, a, i, w.l2.idx[n]);
        for (int c = 0; c < 3; c++) PR(c, a, i, i) -= s2[c];
      }
    }
  if (sigmas)
    for (int a = 0; a < F; a++) {
      if (sigmas[a] < 1e-08) continue;
      for (int f = 0; f < nAt; f++) for (int i = 0; i < nAt; i++) for (int c = 0; c < 3; c++) PR(c, a, i, f) /= sigmas[a];
    }
#undef PR
  free(grad); work_free(&w); free(g);
}

/* ------------------------------------------------------------------------------------------
 * lib_nn/transfer.F90:54-342 activation functions and derivatives.
 * ids: 0 gaussian, 1 relu, 2 lrelu, 3 softplus, 4 bent, 5 atan, 6 sigmoid, 7 heaviside,
 *      8 tanh, 9 linear
 * ---------------------------------------------------------------------------------------- */
static double act_f(int id, double x) {
  switch (id) {
  case 0: return exp(-x * x);
  case 1: return fmax(0.0, x);
  case 2: return fmax(0.01 * x, x);
  case 3: return log(1.0 + exp(x));
  case 4: return (sqrt(x * x + 1.0) - 1.0) / 2.0 + x;
  case 5: return atan(x);
  case 6: return 1.0 / (1.0 + exp(-x));
  case 7: return x > 0.0 ? 1.0 : 0.0;
  case 8: return tanh(x);
  default: return x;
  }
}
static double act_d(int id, double x) {
  switch (id) {
  case 0: return -2.0 * x * exp(-x * x);
  case 1: return x >= 0.0 ? 1.0 : 0.0;
  case 2: return x >= 0.0 ? 1.0 : 0.01;
  case 3: return 1.0 / (1.0 + exp(-x));
  case 4: return x / (2.0 * sqrt(x * x + 1.0)) + 1.0;
  case 5: return 1.0 / (x * x + 1.0);
  case 6: { double s = 1.0 / (1.0 + exp(-x)); return s * (1.0 - s); }
  case 7: return 0.0;
  case 8: { double t = tanh(x); return 1.0 - t * t; }
  default: return 1.0;
  }
}

/* Serialised parameter layout, lib_nn/network.F90:397-427 + 95-142 + nestedtypes.F90:428-470:
 * weights layer 1..L (ww(d_l, d_{l+1}) column-major; layer L is the dummy ww(d_L,1)), then
 * biases layer 1..L (bb(d_l); bb(d_1) unused).  nW = sum d_l d_{l+1} + d_L, nB = sum d_l. */
typedef struct { int L; const int *d; int woff[16], boff[16], nTot; } Topo;
static void topo_init(Topo *t, int L, const int *dims) {
  t->L = L; t->d = dims;
  int ind = 0;
  for (int l = 0; l < L; l++) { t->woff[l] = ind; ind += dims[l] * (l + 1 < L ? dims[l + 1] : 1); }
  for (int l = 0; l < L; l++) { t->boff[l] = ind; ind += dims[l]; }
  t->nTot = ind;
}
int fnet_oracle_ntot(int L, const int *dims) { Topo t; topo_init(&t, L, dims); return t.nTot; }

/* network.F90:146-180 fprop: aa[l], aarg[l] for all layers; hidden act `act`, output linear (:391) */
static void fprop(const Topo *t, const double *wb, int act, const double *x, double **aa, double **aarg) {
  memcpy(aa[0], x, sizeof(double) * t->d[0]);
  for (int l = 1; l < t->L; l++) {
    const double *W = wb + t->woff[l - 1]; /* ww(i,o) = W[i + d_{l-1}*o] */
    const double *b = wb + t->boff[l];
    int din = t->d[l - 1], dout = t->d[l];
    int id = (l == t->L - 1) ? 9 : act;
    for (int o = 0; o < dout; o++) {
      double s = 0.0;
      for (int i = 0; i < din; i++) s += W[i + din * o] * aa[l - 1][i];
      aarg[l][o] = s + b[o];
      aa[l][o] = act_f(id, aarg[l][o]);
    }
  }
}

/* network.F90:248-296 bprop: dd (serialised layout, nTot) += scale * gradient of one atom */
static void bprop_acc(const Topo *t, const double *wb, int act, double **aa, double **aarg,
                      const double *lossgrad, double scale, double *dd, double **delta) {
  int L = t->L;
  for (int o = 0; o < t->d[L - 1]; o++) delta[L - 1][o] = lossgrad[o] * act_d(9, aarg[L - 1][o]);
  for (int l = L - 1; l >= 1; l--) {
    int din = t->d[l - 1], dout = t->d[l];
    /* dw(l-1) = aa(l-1) (x) db(l) ; db(l) = delta(l) */
    double *dw = dd + t->woff[l - 1];
    for (int o = 0; o < dout; o++) {
      for (int i = 0; i < din; i++) dw[i + din * o] += aa[l - 1][i] * delta[l][o] * scale;
      dd[t->boff[l] + o] += delta[l][o] * scale;
    }
    if (l >= 2) {
      const double *W = wb + t->woff[l - 1];
      int id = act; /* layers 2..L-1 are hidden */
      for (int i = 0; i < din; i++) {
        double s = 0.0;
        for (int o = 0; o < dout; o++) s += W[i + din * o] * delta[l][o];
        delta[l - 1][i] = s * act_d(id, aarg[l - 1][i]);
      }
    }
  }
}

static double **alloc_layers(const Topo *t) {
  double **p = (double **)malloc(sizeof(double *) * t->L);
  for (int l = 0; l < t->L; l++) p[l] = (double *)calloc(t->d[l] > 0 ? t->d[l] : 1, sizeof(double));
  return p;
}
static void free_layers(const Topo *t, double **p) { for (int l = 0; l < t->L; l++) free(p[l]); free(p); }

/* lib_nn/bpnn.F90:867-900,1001-1058 iPredict / predictBatch: raw[nOut*i + t] */
void fnet_oracle_predict(int N, int nFeat, const double *feats, const int *globalsp, int nSpecies,
                         int L, const int *dims, int act, const double *wb, double *raw, int nThreads) {
  Topo t; topo_init(&t, L, dims);
  (void)nSpecies;
  if (nThreads < 1) nThreads = 1;
#pragma omp parallel num_threads(nThreads)
  {
    double **aa = alloc_layers(&t), **aarg = alloc_layers(&t);
#pragma omp for schedule(static)
    for (int i = 0; i < N; i++) {
      fprop(&t, wb + (size_t)t.nTot * (globalsp[i] - 1), act, feats + (size_t)nFeat * i, aa, aarg);
      memcpy(raw + (size_t)dims[L - 1] * i, aa[L - 1], sizeof(double) * dims[L - 1]);
    }
    free_layers(&t, aa); free_layers(&t, aarg);
  }
}

/* lib_common/loss.F90:217-281 loss gradients; ids: 0 mse, 1 rms, 2 mae, 3 mape */
static double loss_grad(int id, double p, double t) {
  switch (id) {
  case 1: return (p - t) / sqrt((p - t) * (p - t));
  case 2: return (p - t) / fabs(p - t);
  case 3: return 100.0 * (p - t) / (t * t * fabs(p / t - 1.0));
  default: return 2.0 * (p - t);
  }
}
/* loss.F90:284-366 simple*Loss over n values */
static double simple_loss(int id, int n, const double *p, const double *t) {
  double s = 0.0;
  switch (id) {
  case 1: for (int i = 0; i < n; i++) s += (t[i] - p[i]) * (t[i] - p[i]); return sqrt(s / n);
  case 2: for (int i = 0; i < n; i++) s += fabs(t[i] - p[i]); return s / n;
  case 3: for (int i = 0; i < n; i++) s += fabs((t[i] - p[i]) / t[i]); return 100.0 * s / n;
  default: for (int i = 0; i < n; i++) s += (t[i] - p[i]) * (t[i] - p[i]); return s / n;
  }
}

/* loss.F90:370-721 (maLoss/mapLoss/msLoss/rmsLoss share one shape) */
double fnet_oracle_loss(int nStruct, const int *offsets, int lossId, int nG, int nA, const double *raw,
                        const double *gTargets, const double *aTargets, const double *atomicWeights,
                        const int *dsWeights) {
  int nOut = nG + nA;
  double loss = 0.0, nValues = 0.0;
  double *gp = (double *)malloc(sizeof(double) * (nG > 0 ? nG : 1));
  if (nG > 0)
    for (int s = 0; s < nStruct; s++) {
      double sw = 0.0;
      for (int t = 0; t < nG; t++) gp[t] = 0.0;
      for (int i = offsets[s]; i < offsets[s + 1]; i++) {
        for (int t = 0; t < nG; t++) gp[t] += raw[(size_t)nOut * i + t];
        sw += atomicWeights[i];
      }
      double w = dsWeights ? (double)dsWeights[s] : 1.0;
      loss += w * sw * simple_loss(lossId, nG, gp, gTargets + (size_t)nG * s);
      nValues += w * sw;
    }
  if (nA > 0)
    for (int s = 0; s < nStruct; s++) {
      double w = dsWeights ? (double)dsWeights[s] : 1.0;
      for (int i = offsets[s]; i < offsets[s + 1]; i++) {
        loss += w * atomicWeights[i] * simple_loss(lossId, nA, raw + (size_t)nOut * i + nG, aTargets + (size_t)nA * i);
        nValues += w * atomicWeights[i];
      }
    }
  free(gp);
  return loss / nValues;
}

/* ------------------------------------------------------------------------------------------
 * bpnn.F90:394-481 updateGradients + :610-704 sysTrain.  ddSerial[nTot*sp + p] is
 * TDerivs_serialized(resDd) (un-normalised, before regularisation); raw = resPredicts.
 * nThreads > 1: block partition over structures + sum (the MPI path :436-467).
 * shuffle (1-based, or NULL) only permutes the visiting order (:436-437).
 * ---------------------------------------------------------------------------------------- */
void fnet_oracle_grad(int nStruct, const int *offsets, int nFeat, const double *feats,
                      const int *globalsp, int nSpecies, int L, const int *dims, int act,
                      const double *wb, int lossId, const int *dsWeights, const double *atomicWeights,
                      int nG, const double *gTargets, int nA, const double *aTargets,
                      const int *shuffle, double *ddSerial, double *raw, int nThreads) {
  Topo t; topo_init(&t, L, dims);
  int nOut = dims[L - 1];
  size_t nDD = (size_t)t.nTot * nSpecies;
  if (nThreads < 1) nThreads = 1;
  double *part = (double *)calloc(nDD * nThreads, sizeof(double));
#pragma omp parallel for num_threads(nThreads) schedule(static, 1)
  for (int p = 0; p < nThreads; p++) {
    int s0, s1;
    fnet_oracle_start_end(nStruct, nThreads, p, &s0, &s1);
    double *dd = part + nDD * p;
    double *ddTmp = (double *)malloc(sizeof(double) * nDD);
    double **delta = alloc_layers(&t);
    double *lossgrads = NULL; double ***AA = NULL, ***AARG = NULL; int cap = 0;
    double *gp = (double *)malloc(sizeof(double) * (nG > 0 ? nG : 1));
    for (int ii = s0 - 1; ii < s1; ii++) {
      int s = shuffle ? shuffle[ii] - 1 : ii;
      int o = offsets[s], nAt = offsets[s + 1] - offsets[s];
      if (cap < nAt) { /* TMultiLayerStruc: activations of every atom are kept (:665,672) */
        AA = (double ***)realloc(AA, sizeof(double **) * nAt);
        AARG = (double ***)realloc(AARG, sizeof(double **) * nAt);
        for (int i = cap; i < nAt; i++) { AA[i] = alloc_layers(&t); AARG[i] = alloc_layers(&t); }
        lossgrads = (double *)realloc(lossgrads, sizeof(double) * nOut * nAt);
        cap = nAt;
      }
      memset(ddTmp, 0, sizeof(double) * nDD);
      for (int i = 0; i < nAt; i++) {
        int sp = globalsp[o + i] - 1;
        fprop(&t, wb + (size_t)t.nTot * sp, act, feats + (size_t)nFeat * (o + i), AA[i], AARG[i]);
        memcpy(raw + (size_t)nOut * (o + i), AA[i][L - 1], sizeof(double) * nOut);
      }
      if (nG > 0) { /* :677-684 */
        for (int k = 0; k < nG; k++) gp[k] = 0.0;
        for (int i = 0; i < nAt; i++) for (int k = 0; k < nG; k++) gp[k] += raw[(size_t)nOut * (o + i) + k];
        for (int i = 0; i < nAt; i++)
          for (int k = 0; k < nG; k++) lossgrads[nOut * i + k] = loss_grad(lossId, gp[k], gTargets[(size_t)nG * s + k]);
      }
      if (nA > 0) /* :686-689 */
        for (int i = 0; i < nAt; i++)
          for (int k = 0; k < nA; k++)
            lossgrads[nOut * i + nG + k] = loss_grad(lossId, raw[(size_t)nOut * (o + i) + nG + k], aTargets[(size_t)nA * (o + i) + k]);
      for (int i = 0; i < nAt; i++) { /* :692-702 */
        int sp = globalsp[o + i] - 1;
        bprop_acc(&t, wb + (size_t)t.nTot * sp, act, AA[i], AARG[i], lossgrads + nOut * i,
                  atomicWeights[o + i] / (double)nAt, ddTmp + (size_t)t.nTot * sp, delta);
      }
      double w = dsWeights ? (double)dsWeights[s] : 1.0; /* :443-450 */
      for (size_t q = 0; q < nDD; q++) dd[q] += ddTmp[q] * w;
    }
    for (int i = 0; i < cap; i++) { free_layers(&t, AA[i]); free_layers(&t, AARG[i]); }
    free(AA); free(AARG); free(lossgrads); free(gp); free(ddTmp); free_layers(&t, delta);
  }
  for (size_t q = 0; q < nDD; q++) { /* mpifx_allreduce(SUM) :460-467 */
    double s = 0.0;
    for (int p = 0; p < nThreads; p++) s += part[nDD * p + q];
    ddSerial[q] = s;
  }
  free(part);
}

/* network.F90:183-244 fdevi: full Jacobian jac[t + nOut*a] = d out_t / d x_a by forward mode */
static void fdevi(const Topo *t, const double *wb, int act, const double *x, double *jac) {
  int L = t->L, F = t->d[0];
  int dmax = 1;
  for (int l = 0; l < L; l++) if (t->d[l] > dmax) dmax = t->d[l];
  double *aa = (double *)malloc(sizeof(double) * dmax), *an = (double *)malloc(sizeof(double) * dmax);
  double *J = (double *)calloc((size_t)dmax * F, sizeof(double)), *Jn = (double *)calloc((size_t)dmax * F, sizeof(double));
  memcpy(aa, x, sizeof(double) * F);
  for (int a = 0; a < F; a++) J[a + (size_t)dmax * a] = 1.0; /* J[row + dmax*col] */
  for (int l = 1; l < L; l++) {
    const double *W = wb + t->woff[l - 1];
    const double *b = wb + t->boff[l];
    int din = t->d[l - 1], dout = t->d[l];
    int id = (l == L - 1) ? 9 : act;
    for (int o = 0; o < dout; o++) {
      double s = 0.0;
      for (int i = 0; i < din; i++) s += W[i + din * o] * aa[i];
      double arg = s + b[o];
      an[o] = act_f(id, arg);
      double d = act_d(id, arg);
      for (int a = 0; a < F; a++) {
        double m = 0.0;
        for (int i = 0; i < din; i++) m += W[i + din * o] * J[i + (size_t)dmax * a];
        Jn[o + (size_t)dmax * a] = d * m;
      }
    }
    double *tp = aa; aa = an; an = tp;
    tp = J; J = Jn; Jn = tp;
  }
  int nOut = t->d[L - 1];
  for (int a = 0; a < F; a++) for (int o = 0; o < nOut; o++) jac[o + nOut * a] = J[o + (size_t)dmax * a];
  free(aa); free(an); free(J); free(Jn);
}

/* bpnn.F90:904-997 nJacobian: jac[(nOut*F)*i + t + nOut*a] */
void fnet_oracle_jacobian(int N, int nFeat, const double *feats, const int *globalsp, int L,
                          const int *dims, int act, const double *wb, double *jac) {
  Topo t; topo_init(&t, L, dims);
  int nOut = dims[L - 1];
  for (int i = 0; i < N; i++)
    fdevi(&t, wb + (size_t)t.nTot * (globalsp[i] - 1), act, feats + (size_t)nFeat * i, jac + (size_t)nOut * nFeat * i);
}

/* ------------------------------------------------------------------------------------------
 * lib_analysis/forces.F90:317-425 forceAnalysis_analytical (loop :400-413), built on
 * TAcsf_calculatePrime (dense tensor per structure) and nJacobian.
 * forces[(3*nOut)*f + 3*t + c]  (= forces%geos(iSys)%array(c + 3*(t-1), f)).
 * nThreads > 1: block partition over structures (forces.F90:355).
 * ---------------------------------------------------------------------------------------- */
void fnet_oracle_forces(int nStruct, const int *offsets, const double *coords, const int *periodic,
                        const double *latvecs, const int *atnum, int nExt, const double *ext,
                        int F, const int *ftype, const double *rcut, const double *kappa,
                        const double *rs, const double *eta, const double *lambda, const double *xi,
                        const int *atomid, const int *atomicnumbers, const double *sigmas,
                        const double *feats, const int *globalsp, int L, const int *dims, int act,
                        const double *wb, double *forces, int nThreads) {
  int nOut = dims[L - 1];
  Topo t; topo_init(&t, L, dims);
  if (nThreads < 1) nThreads = 1;
#pragma omp parallel for num_threads(nThreads) schedule(static, 1)
  for (int p = 0; p < nThreads; p++) {
    int s0, s1;
    fnet_oracle_start_end(nStruct, nThreads, p, &s0, &s1);
    for (int s = s0 - 1; s < s1; s++) {
      int o = offsets[s], nAt = offsets[s + 1] - offsets[s];
      double *prime = (double *)malloc(sizeof(double) * 3 * (size_t)F * nAt * nAt);
      double *jac = (double *)malloc(sizeof(double) * (size_t)nOut * F * nAt);
      fnet_oracle_acsf_prime_struct(nAt, coords + 3 * (size_t)o, periodic[s], latvecs + 9 * (size_t)s,
                                    atnum + o, nExt, ext ? ext + (size_t)nExt * o : NULL, F, ftype, rcut,
                                    kappa, rs, eta, lambda, xi, atomid, atomicnumbers, sigmas, prime);
      for (int i = 0; i < nAt; i++)
        fdevi(&t, wb + (size_t)t.nTot * (globalsp[o + i] - 1), act, feats + (size_t)F * (o + i), jac + (size_t)nOut * F * i);
      for (int f = 0; f < nAt; f++) {
        double *ff = forces + (size_t)3 * nOut * (o + f);
        for (int q = 0; q < 3 * nOut; q++) ff[q] = 0.0;
        for (int i = 0; i < nAt; i++)
          for (int a = 0; a < F; a++)
            for (int c = 0; c < 3; c++) {
              double d = prime[c + 3 * ((size_t)a + (size_t)F * ((size_t)i + (size_t)nAt * f))];
              for (int k = 0; k < nOut; k++) ff[3 * k + c] -= jac[(size_t)nOut * F * i + k + nOut * a] * d;
            }
      }
      free(prime); free(jac);
    }
  }
}

/* lib_dftbp/steepdesc.F90:186-202 next_local with weight(:) = learning rate */
void fnet_oracle_sd_step(int n, double *x, const double *grad, double lr, double maxDisp) {
  double maxX = 0.0;
  for (int i = 0; i < n; i++) { double v = fabs(lr * grad[i]); if (v > maxX) maxX = v; }
  double sc = (maxX <= maxDisp) ? 1.0 : maxDisp / maxX;
  for (int i = 0; i < n; i++) x[i] += sc * (-lr * grad[i]);
}

int fnet_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
