// multi.cuh -- single-process multi-GPU layer of the C ABI (fnetgpu_mg_*, include/fnetgpu.h).
//
// Replaces the reference's MPI structure-level parallelism for a driver that stays ONE process:
// getStartAndEndIndex (lib_common/parallel.F90:23-56) becomes a contiguous split of the structures
// balanced by atom count, the per-structure mpifx_allreduce of lib_nn/bpnn.F90:455-467 one
// ncclAllReduce of [ddSerial | loss numerator | denominator] per iteration (communicators from
// ncclCommInitAll), the z-score statistics two small all-reduces (acsf.F90:618-636).  One context per
// device; every call fans out to one host thread per device (the per-device entry points block), shards
// and gathers the caller's arrays, and reports the first error.  Features, predictions and forces stay sharded on
// the devices; gradient, loss and statistics come back replicated (device 0's copy is returned).
#pragma once
#include <thread>
#include <functional>

struct fnetgpu_mg {
  int nDev = 0;
  std::vector<fnetgpu_ctx *> ctx;
  std::string err;
  struct Shard { int st0 = 0, st1 = 0, a0 = 0, a1 = 0; };
  std::vector<Shard> shards[FNETGPU_MAX_SLOTS];      // per slot: structures / atoms of every device
  int nStruct[FNETGPU_MAX_SLOTS] = {0}, nAtoms[FNETGPU_MAX_SLOTS] = {0}, nG[FNETGPU_MAX_SLOTS] = {0};
};

static thread_local std::string g_mg_err;

// runs fn(d) on one thread per device; first non-zero status wins, its message is kept
static int mg_fanout(fnetgpu_mg *mg, const std::function<int(int)> &fn) {
  std::vector<int> rc(mg->nDev, 0);
  std::vector<std::thread> th;
  for (int d = 1; d < mg->nDev; d++) th.emplace_back([&, d] { rc[d] = fn(d); });
  rc[0] = fn(0);
  for (auto &t : th) t.join();
  for (int d = 0; d < mg->nDev; d++)
    if (rc[d] != 0) {
      mg->err = "device " + std::to_string(mg->ctx[d]->device) + ": " + mg->ctx[d]->err;
      return 1;
    }
  return 0;
}

extern "C" const char *fnetgpu_mg_last_error(const fnetgpu_mg *mg) { return mg ? mg->err.c_str() : g_mg_err.c_str(); }
extern "C" int fnetgpu_mg_device_count(const fnetgpu_mg *mg) { return mg ? mg->nDev : 0; }
extern "C" fnetgpu_ctx *fnetgpu_mg_context(fnetgpu_mg *mg, int d) { return (mg && d >= 0 && d < mg->nDev) ? mg->ctx[d] : nullptr; }

extern "C" int fnetgpu_mg_finalize(fnetgpu_mg *mg) {
  if (!mg) return 0;
  for (fnetgpu_ctx *c : mg->ctx) fnetgpu_finalize(c);
  delete mg;
  return 0;
}

// nDevicesRequested <= 0: every visible device.  Devices 0 .. n-1, one context each, one NCCL communicator
// over all of them (none for a single device).
extern "C" int fnetgpu_mg_init(fnetgpu_mg **out, int nDevicesRequested, int precision, int deterministic) {
  if (!out) { g_mg_err = "null out pointer"; return 1; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_mg_err = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return 1;
  }
  const int n = nDevicesRequested <= 0 ? ndev : nDevicesRequested;
  if (n > ndev) { g_mg_err = "requested " + std::to_string(n) + " devices, " + std::to_string(ndev) + " visible"; return 1; }
  fnetgpu_mg *mg = new fnetgpu_mg();
  mg->nDev = n;
  for (int d = 0; d < n; d++) {
    fnetgpu_ctx *c = nullptr;
    if (fnetgpu_init(&c, d, precision, deterministic)) {
      g_mg_err = g_err;
      fnetgpu_mg_finalize(mg);
      return 1;
    }
    mg->ctx.push_back(c);
  }
  if (n > 1) {
    void *h = nccl_handle();
    if (!h) { g_mg_err = std::string("cannot load libnccl: ") + dlerror(); fnetgpu_mg_finalize(mg); return 1; }
    typedef int (*initall_t)(void **, int, const int *);
    initall_t f = (initall_t)dlsym(h, "ncclCommInitAll");
    std::vector<void *> comms(n, nullptr);
    std::vector<int> devs(n);
    for (int d = 0; d < n; d++) devs[d] = d;
    const int rc = f ? f(comms.data(), n, devs.data()) : -1;
    if (rc != 0) { g_mg_err = "ncclCommInitAll failed (" + std::to_string(rc) + ")"; fnetgpu_mg_finalize(mg); return 1; }
    for (int d = 0; d < n; d++) { mg->ctx[d]->nccl = h; mg->ctx[d]->comm = comms[d]; mg->ctx[d]->nRanks = n; mg->ctx[d]->rank = d; }
  }
  *out = mg;
  return 0;
}

// contiguous blocks of structures, balanced by atom count (the reference splits by structure count,
// parallel.F90:43-54); every device gets at least one structure
static void mg_partition(int nStruct, const int *offsets, int nDev, std::vector<fnetgpu_mg::Shard> &sh) {
  sh.assign(nDev, fnetgpu_mg::Shard());
  const long long N = offsets[nStruct];
  int st = 0;
  for (int d = 0; d < nDev; d++) {
    const int left = nDev - d;                       // devices still to fill (this one included)
    const long long target = offsets[st] + (N - offsets[st] + left - 1) / left;
    int e = st + 1;
    while (e < nStruct - (left - 1) && offsets[e] < target) e++;
    if (d == nDev - 1) e = nStruct;
    sh[d].st0 = st; sh[d].st1 = e; sh[d].a0 = offsets[st]; sh[d].a1 = offsets[e];
    st = e;
  }
}

extern "C" int fnetgpu_mg_dataset_upload(fnetgpu_mg *mg, int slot, int nStruct, const int *offsets, const double *coords,
                                         const int *periodic, const double *latvecs, const int *atnum, const int *globalsp,
                                         const int *dsWeights, const double *atomicWeights, int nG, const double *gTargets,
                                         int nA, const double *aTargets, int nExt, const double *ext) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  mg->err.clear();
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS) { mg->err = "slot out of range"; return 1; }
  if (nStruct < mg->nDev) { mg->err = "fewer structures than devices"; return 1; }
  if (!offsets || offsets[0] != 0) { mg->err = "dataset_upload: offsets must start at 0"; return 1; }
  mg_partition(nStruct, offsets, mg->nDev, mg->shards[slot]);
  mg->nStruct[slot] = nStruct; mg->nAtoms[slot] = offsets[nStruct]; mg->nG[slot] = nG;
  return mg_fanout(mg, [&](int d) {
    const fnetgpu_mg::Shard &S = mg->shards[slot][d];
    std::vector<int> off(S.st1 - S.st0 + 1);
    for (int s = S.st0; s <= S.st1; s++) off[s - S.st0] = offsets[s] - S.a0;
    return fnetgpu_dataset_upload(mg->ctx[d], slot, S.st1 - S.st0, off.data(), coords ? coords + (size_t)3 * S.a0 : nullptr,
                                  periodic ? periodic + S.st0 : nullptr, latvecs ? latvecs + (size_t)9 * S.st0 : nullptr,
                                  atnum + S.a0, globalsp + S.a0, dsWeights ? dsWeights + S.st0 : nullptr,
                                  atomicWeights ? atomicWeights + S.a0 : nullptr, nG, gTargets ? gTargets + (size_t)nG * S.st0 : nullptr,
                                  nA, aTargets ? aTargets + (size_t)nA * S.a0 : nullptr, nExt, ext ? ext + (size_t)nExt * S.a0 : nullptr);
  });
}

extern "C" int fnetgpu_mg_coords_update(fnetgpu_mg *mg, int slot, const double *coords, const double *latvecs) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS || mg->shards[slot].empty()) { mg->err = "coords_update: empty slot"; return 1; }
  return mg_fanout(mg, [&](int d) {
    const fnetgpu_mg::Shard &S = mg->shards[slot][d];
    return fnetgpu_coords_update(mg->ctx[d], slot, coords + (size_t)3 * S.a0, latvecs ? latvecs + (size_t)9 * S.st0 : nullptr);
  });
}

extern "C" int fnetgpu_mg_acsf_set(fnetgpu_mg *mg, int F, const int *type, const double *rcut, const double *kappa, const double *rs,
                                   const double *eta, const double *lambda, const double *xi, const int *atomid,
                                   const int *atomicnumbers) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  return mg_fanout(mg, [&](int d) { return fnetgpu_acsf_set(mg->ctx[d], F, type, rcut, kappa, rs, eta, lambda, xi, atomid, atomicnumbers); });
}
extern "C" int fnetgpu_mg_features_config(fnetgpu_mg *mg, int nExtSel, const int *extIndices) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  return mg_fanout(mg, [&](int d) { return fnetgpu_features_config(mg->ctx[d], nExtSel, extIndices); });
}
extern "C" int fnetgpu_mg_net_set(fnetgpu_mg *mg, int nSpecies, int nLayers, const int *dims, int activationId) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  return mg_fanout(mg, [&](int d) { return fnetgpu_net_set(mg->ctx[d], nSpecies, nLayers, dims, activationId); });
}
extern "C" int fnetgpu_mg_params_set(fnetgpu_mg *mg, const double *wb) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  return mg_fanout(mg, [&](int d) { return fnetgpu_params_set(mg->ctx[d], wb); });
}

// TAcsf_calculate over all devices; the statistics (when computed here) are those of the WHOLE dataset
// (all-reduced inside fnetgpu_acsf_calculate) and identical on every device
extern "C" int fnetgpu_mg_acsf_calculate(fnetgpu_mg *mg, int slot, int standardize, double *zprec, int have_zprec) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  const int F = mg->ctx[0]->acsf.F;
  std::vector<std::vector<double>> zp(mg->nDev);
  for (int d = 0; d < mg->nDev; d++) {
    zp[d].assign((size_t)2 * std::max(F, 1), 0.0);
    if (zprec && have_zprec) zp[d].assign(zprec, zprec + (size_t)2 * F);
  }
  if (mg_fanout(mg, [&](int d) { return fnetgpu_acsf_calculate(mg->ctx[d], slot, standardize, zp[d].data(), have_zprec); })) return 1;
  if (zprec && standardize && !have_zprec) memcpy(zprec, zp[0].data(), (size_t)2 * F * sizeof(double));
  return 0;
}

// updateGradients over all devices: one all-reduce inside; every device returns the same gradient and loss
extern "C" int fnetgpu_mg_grad(fnetgpu_mg *mg, int slot, int lossId, const int *shuffle, double *ddSerial, double *loss,
                               double *globalPred) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS || mg->shards[slot].empty()) { mg->err = "grad: empty slot"; return 1; }
  std::vector<double> lossv(mg->nDev, 0.0);
  const int nG = mg->nG[slot];
  if (mg_fanout(mg, [&](int d) {
        const fnetgpu_mg::Shard &S = mg->shards[slot][d];
        return fnetgpu_grad(mg->ctx[d], slot, lossId, shuffle, d == 0 ? ddSerial : nullptr, &lossv[d],
                            (globalPred && nG > 0) ? globalPred + (size_t)nG * S.st0 : nullptr);
      })) return 1;
  if (loss) *loss = lossv[0];
  return 0;
}

extern "C" int fnetgpu_mg_loss(fnetgpu_mg *mg, int slot, int lossId, double *loss) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  std::vector<double> lossv(mg->nDev, 0.0);
  if (mg_fanout(mg, [&](int d) { return fnetgpu_loss(mg->ctx[d], slot, lossId, &lossv[d]); })) return 1;
  if (loss) *loss = lossv[0];
  return 0;
}

extern "C" int fnetgpu_mg_predict(fnetgpu_mg *mg, int slot, double *raw) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS || mg->shards[slot].empty()) { mg->err = "predict: empty slot"; return 1; }
  const int nOut = mg->ctx[0]->net.nOut;
  return mg_fanout(mg, [&](int d) { return fnetgpu_predict(mg->ctx[d], slot, raw ? raw + (size_t)nOut * mg->shards[slot][d].a0 : nullptr); });
}

extern "C" int fnetgpu_mg_forces(fnetgpu_mg *mg, int slot, double *forces) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS || mg->shards[slot].empty()) { mg->err = "forces: empty slot"; return 1; }
  const int nOut = mg->ctx[0]->net.nOut;
  return mg_fanout(mg, [&](int d) { return fnetgpu_forces(mg->ctx[d], slot, forces ? forces + (size_t)3 * nOut * mg->shards[slot][d].a0 : nullptr); });
}

extern "C" int fnetgpu_mg_features_get(fnetgpu_mg *mg, int slot, double *out) {
  if (!mg) { g_mg_err = "null multi-GPU context"; return 1; }
  if (slot < 0 || slot >= FNETGPU_MAX_SLOTS || mg->shards[slot].empty()) { mg->err = "features_get: empty slot"; return 1; }
  return mg_fanout(mg, [&](int d) {
    const int nFeat = mg->ctx[d]->slots[slot].nFeat;
    return fnetgpu_features_get(mg->ctx[d], slot, out + (size_t)nFeat * mg->shards[slot][d].a0);
  });
}

// shard of device d in slot: [st0, st1) structures, [a0, a1) atoms (what getStartAndEndIndex returned per rank)
extern "C" int fnetgpu_mg_shard(const fnetgpu_mg *mg, int slot, int d, int *st0, int *st1, int *a0, int *a1) {
  if (!mg || slot < 0 || slot >= FNETGPU_MAX_SLOTS || d < 0 || d >= mg->nDev || mg->shards[slot].empty()) return 1;
  const fnetgpu_mg::Shard &S = mg->shards[slot][d];
  if (st0) *st0 = S.st0; if (st1) *st1 = S.st1; if (a0) *a0 = S.a0; if (a1) *a1 = S.a1;
  return 0;
}
