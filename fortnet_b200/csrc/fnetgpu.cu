// fnetgpu.cu -- the C ABI of libfnetgpu.so (include/fnetgpu.h): host-side orchestration of the
// sm_100a kernels in cells.cuh / acsf.cuh / acsf_force.cuh / mlp.cuh.  No torch types, no CPU
// fallback: every entry point either runs the CUDA path or returns an error.
#include "internal.h"
#include "cells.cuh"
#include "acsf.cuh"
#include "acsf_lean.cuh"
#include "acsf_force.cuh"
#include "acsf_force_lean.cuh"
#include "mlp.cuh"
#include "mlp_mma.cuh"
#include "mlp_warp.cuh"

#include <dlfcn.h>
#include <map>
#include <tuple>
#include <type_traits>

static const char *kKernelNames[K_NUM_KERNELS] = {
    "bin_count", "bin_scan", "bin_fill", "bin_sort", "neigh_count", "acsf", "zstat", "zstat_final",
    "zapply", "ext_concat", "mlp_fwd", "struct_loss", "loss_final", "mlp_grad", "grad_reduce",
    "mlp_ingrad", "acsf_force", "misc", "mlp_warp"};

extern "C" const char *fnetgpu_kernel_name(int kernelId) {
  return (kernelId >= 0 && kernelId < K_NUM_KERNELS) ? kKernelNames[kernelId] : nullptr;
}

static thread_local std::string g_err;  // errors before a context exists

// dynamic shared memory opt-in of a kernel + the MAXIMUM shared-memory carve-out of the unified L1: without the explicit
// preference the driver picks the split per launch from recent history, and a context that alternates kernels with
// different footprints can settle on a smaller carve-out -- measured on this pool as one CTA per SM less for the ACSF
// kernels (C3: 21.6 instead of 16.7 ms at identical clocks).  Occupancy must not depend on launch history.
template <typename K>
static cudaError_t fnet_smem_attr(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

#define CHECK_CTX(ctx) do { if (!(ctx)) { g_err = "null context"; return 1; } (ctx)->err.clear(); } while (0)
#define CHECK_SLOT(ctx, slot)                                                        \
  do { if ((slot) < 0 || (slot) >= FNETGPU_MAX_SLOTS) FNET_FAIL(ctx, "slot out of range"); } while (0)

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" int fnetgpu_init(fnetgpu_ctx **out, int device, int precision, int deterministic) {
  if (!out) { g_err = "null out pointer"; return 1; }
  *out = nullptr;
  if (precision != 64 && precision != 32) { g_err = "precision must be 64 or 32"; return 1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_err = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return 1;
  }
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % ndev : 0;
  }
  if (device >= ndev) { g_err = "device index out of range"; return 1; }
  fnetgpu_ctx *ctx = new fnetgpu_ctx();
  ctx->device = device; ctx->precision = precision; ctx->deterministic = deterministic;
  memset(ctx->kms, 0, sizeof(ctx->kms)); memset(ctx->klaunch, 0, sizeof(ctx->klaunch)); memset(ctx->kprof, 0, sizeof(ctx->kprof));
  memset(&ctx->acsf, 0, sizeof(ctx->acsf)); memset(&ctx->net, 0, sizeof(ctx->net));
  if (cudaSetDevice(device) != cudaSuccess) { g_err = "cudaSetDevice failed"; delete ctx; return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->nSM = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { g_err = "stream creation failed"; delete ctx; return 1; }
  {   // every allocation of the context is checked: a half-built context would fail later, far from the cause
    cudaError_t ce = cudaEventCreate(&ctx->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&ctx->ev1);
    if (ce == cudaSuccess) ce = cudaMalloc((void **)&ctx->d_flags, 16 * sizeof(int));   // [0..7] per-launch flags, [8..15] cell-list statistics
    if (ce == cudaSuccess) ce = cudaMemset(ctx->d_flags, 0, 16 * sizeof(int));
    if (ce != cudaSuccess) {
      g_err = std::string("fnetgpu_init: ") + cudaGetErrorString(ce);
      if (ctx->ev0) cudaEventDestroy(ctx->ev0);
      if (ctx->ev1) cudaEventDestroy(ctx->ev1);
      cudaFree(ctx->d_flags); cudaStreamDestroy(ctx->stream);
      delete ctx;
      return 1;
    }
  }
  { const char *pm = getenv("FNETGPU_MLP"); ctx->mlpLegacy = (pm && strcmp(pm, "legacy") == 0) ? 1 : 0;
    ctx->mlpNoFuse = (pm && strcmp(pm, "nofuse") == 0) ? 1 : 0; }
  { const char *pm = getenv("FNETGPU_GRAPHS"); ctx->useGraphs = !(pm && strcmp(pm, "0") == 0); }
  { const char *pm = getenv("FNETGPU_PDL"); ctx->usePdl = !(pm && strcmp(pm, "0") == 0); }
  { const char *pm = getenv("FNETGPU_ACSF_KERNEL"); ctx->acsfGeneric = (pm && strcmp(pm, "generic") == 0) ? 1 : 0; }
  { const char *pm = getenv("FNETGPU_ACSF_PATH"); ctx->acsfPathCells = (pm && strcmp(pm, "cells") == 0) ? 1 : 0; }
  *out = ctx;
  return 0;
}

static void free_slot(Slot &s) {
  if (s.sockGraph) { cudaGraphExecDestroy(s.sockGraph); s.sockGraph = nullptr; }
  cudaFree(s.d_offsets); cudaFree(s.d_structOf); cudaFree(s.d_atnum); cudaFree(s.d_sp); cudaFree(s.d_periodic);
  cudaFree(s.d_coords); cudaFree(s.d_lat); cudaFree(s.d_fpos); cudaFree(s.d_crec); cudaFree(s.d_binStruct); cudaFree(s.d_sinfo); cudaFree(s.d_atomCell);
  cudaFree(s.d_cellStart); cudaFree(s.d_cellCount); cudaFree(s.d_cellAtoms); cudaFree(s.d_dsw); cudaFree(s.d_aw);
  cudaFree(s.d_gt); cudaFree(s.d_at); cudaFree(s.d_ext); cudaFree(s.d_feat);
  cudaFree(s.d_perm); cudaFree(s.d_tiles); cudaFree(s.d_tiles16); cudaFree(s.d_tilesS); cudaFree(s.d_tilesW); cudaFree(s.d_tilesC); cudaFree(s.d_permC); cudaFree(s.d_segBE); cudaFree(s.d_raw); cudaFree(s.d_gS); cudaFree(s.d_Es);
  cudaFree(s.d_lossPart); cudaFree(s.d_dEdG); cudaFree(s.d_forces);
  s = Slot();
}

extern "C" int fnetgpu_finalize(fnetgpu_ctx *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->arStream) { cudaStreamSynchronize(ctx->arStream); cudaStreamDestroy(ctx->arStream); cudaEventDestroy(ctx->evGrad); cudaEventDestroy(ctx->evAR); }
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < FNETGPU_MAX_SLOTS; i++) free_slot(ctx->slots[i]);
  cudaFree(ctx->d_rgroups); cudaFree(ctx->d_rfeat); cudaFree(ctx->d_rp1); cudaFree(ctx->d_rp2);
  cudaFree(ctx->d_apasses); cudaFree(ctx->d_lrad); cudaFree(ctx->d_lpass); cudaFree(ctx->d_powtab); cudaFree(ctx->d_pairtab); cudaFree(ctx->d_extIdx); cudaFree(ctx->d_zprec); cudaFree(ctx->d_wb);
  cudaFree(ctx->d_wb64); cudaFree(ctx->d_fpart); cudaFree(ctx->d_conv); cudaFree(ctx->d_partials); cudaFree(ctx->d_dd); cudaFree(ctx->d_flags);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_pinIn) cudaFreeHost(ctx->h_pinIn);
  if (ctx->comm && ctx->nccl) {
    typedef int (*destroy_t)(void *);
    destroy_t d = (destroy_t)dlsym(ctx->nccl, "ncclCommDestroy");
    if (d) d(ctx->comm);
  }
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  if (ctx->copyStream) { cudaStreamDestroy(ctx->copyStream); for (int i = 0; i < 8; i++) cudaEventDestroy(ctx->evChunk[i]); }
  if (ctx->prof) { for (int i = 0; i < FNET_PROF_RING; i++) { cudaEventDestroy(ctx->prof[i].a); cudaEventDestroy(ctx->prof[i].b); } delete[] ctx->prof; }
  if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" const char *fnetgpu_last_error(const fnetgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int fnetgpu_synchronize(fnetgpu_ctx *ctx) {
  CHECK_CTX(ctx);
  if (ctx->arPending) { CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evAR, 0)); ctx->arPending = false; }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fnetgpu_set_stream(fnetgpu_ctx *ctx, void *stream) {
  CHECK_CTX(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)stream;
  ctx->ownStream = false;
  return 0;
}

extern "C" long long fnetgpu_launch_count(const fnetgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

// resolve the recorded event pairs into per-kernel totals
static void profile_flush(fnetgpu_ctx *ctx) {
  if (!ctx->prof || ctx->profUsed == 0) return;
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < ctx->profUsed; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof[i].a, ctx->prof[i].b) == cudaSuccess) {
      ctx->kms[ctx->prof[i].kernel] += ms;
      ctx->kprof[ctx->prof[i].kernel]++;
    }
  }
  ctx->profUsed = 0;
}

extern "C" int fnetgpu_profile(fnetgpu_ctx *ctx, int enable) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  if (enable && !ctx->prof) {
    ctx->prof = new ProfEvent[FNET_PROF_RING];
    for (int i = 0; i < FNET_PROF_RING; i++) { cudaEventCreate(&ctx->prof[i].a); cudaEventCreate(&ctx->prof[i].b); ctx->prof[i].kernel = 0; }
  }
  if (enable) {
    ctx->profUsed = 0;
    memset(ctx->kms, 0, sizeof(ctx->kms)); memset(ctx->kprof, 0, sizeof(ctx->kprof));
  } else {
    profile_flush(ctx);
  }
  ctx->profiling = enable != 0;
  return 0;
}
// ms_total / launches: summed CUDA-event durations and number of timed launches of this kernel
// since profiling was enabled
extern "C" int fnetgpu_profile_get(fnetgpu_ctx *ctx, int kid, double *ms, long long *launches) {
  CHECK_CTX(ctx);
  if (kid < 0 || kid >= K_NUM_KERNELS) FNET_FAIL(ctx, "kernel id out of range");
  profile_flush(ctx);
  if (ms) *ms = ctx->kms[kid];
  if (launches) *launches = ctx->kprof[kid];
  return 0;
}

static int ensure_pinned(fnetgpu_ctx *ctx, size_t n) {
  if (ctx->pinnedN >= n) return 0;
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  // (h_pinIn, the socket step's staged geometry, has its own capacity and is NOT released here: it used to be freed without
  // being reset, so that growing this buffer after a socket step left a dangling pointer for the next step and a second
  // cudaFreeHost in fnetgpu_finalize -- found by compute-sanitizer on test_reconfigure_with_larger_features_and_outputs)
  ctx->h_pinned = nullptr; ctx->pinnedN = 0;
  CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_pinned, n * sizeof(double)));
  ctx->pinnedN = n;
  return 0;
}

// ------------------------------------------------------------------------------------------
// dataset upload
// ------------------------------------------------------------------------------------------
extern "C" int fnetgpu_dataset_upload(fnetgpu_ctx *ctx, int slot, int nStruct, const int *offsets,
                                      const double *coords, const int *periodic, const double *latvecs,
                                      const int *atnum, const int *globalsp, const int *dsWeights,
                                      const double *atomicWeights, int nG, const double *gTargets, int nA,
                                      const double *aTargets, int nExt, const double *ext) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  if (nStruct <= 0 || !offsets || !atnum || !globalsp) FNET_FAIL(ctx, "dataset_upload: missing arrays");
  cudaSetDevice(ctx->device);
  const int N = offsets[nStruct];
  if (N <= 0 || offsets[0] != 0) FNET_FAIL(ctx, "dataset_upload: offsets must start at 0 and end at N > 0");
  // validate everything before the slot is touched: a failed upload leaves the previous contents usable
  for (int st = 0; st < nStruct; st++)
    if (offsets[st + 1] <= offsets[st]) FNET_FAIL(ctx, "dataset_upload: empty structure");
  for (int i = 0; i < N; i++)
    if (globalsp[i] < 1) FNET_FAIL(ctx, "dataset_upload: globalsp must be 1-based");
  if (nG < 0 || nA < 0 || nExt < 0) FNET_FAIL(ctx, "dataset_upload: negative target / feature count");
  if (nG > 0 && !gTargets) FNET_FAIL(ctx, "gTargets missing");
  if (nA > 0 && !aTargets) FNET_FAIL(ctx, "aTargets missing");
  if (nExt > 0 && !ext) FNET_FAIL(ctx, "ext missing");
  Slot &s = ctx->slots[slot];
  free_slot(s);
  struct SlotGuard { Slot &s; bool ok; ~SlotGuard() { if (!ok) free_slot(s); } } guard{s, false};   // CUDA failures below: slot left empty
  s.nStruct = nStruct; s.N = N; s.nG = nG; s.nA = nA; s.nExt = nExt;
  s.h_offsets.assign(offsets, offsets + nStruct + 1);
  s.h_periodic.assign(nStruct, 0);
  if (periodic) s.h_periodic.assign(periodic, periodic + nStruct);
  s.h_lat.assign((size_t)9 * nStruct, 0.0);
  if (latvecs) s.h_lat.assign(latvecs, latvecs + (size_t)9 * nStruct);
  std::vector<int> structOf(N);
  for (int st = 0; st < nStruct; st++) {
    if (offsets[st + 1] <= offsets[st]) FNET_FAIL(ctx, "dataset_upload: empty structure");
    s.maxAtoms = std::max(s.maxAtoms, offsets[st + 1] - offsets[st]);
    for (int i = offsets[st]; i < offsets[st + 1]; i++) structOf[i] = st;
  }
  int maxSp = 0;
  s.h_globalsp.resize(N);
  for (int i = 0; i < N; i++) {
    if (globalsp[i] < 1) FNET_FAIL(ctx, "dataset_upload: globalsp must be 1-based");
    s.h_globalsp[i] = globalsp[i] - 1;
    maxSp = std::max(maxSp, globalsp[i]);
  }
  // species-sorted processing order (stable)
  std::vector<int> perm(N);
  s.spBeg.assign(maxSp + 1, 0);
  for (int i = 0; i < N; i++) s.spBeg[s.h_globalsp[i] + 1]++;
  for (int k = 0; k < maxSp; k++) s.spBeg[k + 1] += s.spBeg[k];
  {
    std::vector<int> cur(s.spBeg.begin(), s.spBeg.end() - 1);
    for (int i = 0; i < N; i++) perm[cur[s.h_globalsp[i]]++] = i;
  }
  std::vector<double> dsw(nStruct, 1.0), aw(N, 1.0);
  if (dsWeights) for (int st = 0; st < nStruct; st++) dsw[st] = (double)dsWeights[st];
  if (atomicWeights) aw.assign(atomicWeights, atomicWeights + N);
  if (dev_upload(ctx, &s.d_offsets, offsets, (size_t)nStruct + 1)) return 1;
  if (dev_upload(ctx, &s.d_structOf, structOf.data(), (size_t)N)) return 1;
  if (dev_upload(ctx, &s.d_atnum, atnum, (size_t)N)) return 1;
  if (dev_upload(ctx, &s.d_sp, s.h_globalsp.data(), (size_t)N)) return 1;
  if (dev_upload(ctx, &s.d_perm, perm.data(), (size_t)N)) return 1;
  if (dev_upload(ctx, &s.d_dsw, dsw.data(), (size_t)nStruct)) return 1;
  if (dev_upload(ctx, &s.d_aw, aw.data(), (size_t)N)) return 1;
  if (coords) { if (dev_upload(ctx, &s.d_coords, coords, (size_t)3 * N)) return 1; s.capCoords = (size_t)3 * N; }
  if (dev_upload(ctx, &s.d_lat, s.h_lat.data(), (size_t)9 * nStruct)) return 1;
  if (dev_upload(ctx, &s.d_periodic, s.h_periodic.data(), (size_t)nStruct)) return 1;
  if (nG > 0) { if (!gTargets) FNET_FAIL(ctx, "gTargets missing"); if (dev_upload(ctx, &s.d_gt, gTargets, (size_t)nG * nStruct)) return 1; }
  if (nA > 0) { if (!aTargets) FNET_FAIL(ctx, "aTargets missing"); if (dev_upload(ctx, &s.d_at, aTargets, (size_t)nA * N)) return 1; }
  if (nExt > 0) { if (!ext) FNET_FAIL(ctx, "ext missing"); if (dev_upload(ctx, &s.d_ext, ext, (size_t)nExt * N)) return 1; }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  s.used = true; guard.ok = true;
  return 0;
}

extern "C" int fnetgpu_coords_update(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used) FNET_FAIL(ctx, "coords_update: empty slot");
  cudaSetDevice(ctx->device);
  if (!coords) FNET_FAIL(ctx, "coords_update: coords missing");
  if (dev_reserve(ctx, &s.d_coords, &s.capCoords, (size_t)3 * s.N)) return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(s.d_coords, coords, (size_t)3 * s.N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (latvecs) {
    CUDA_TRY(ctx, cudaMemcpyAsync(s.d_lat, latvecs, (size_t)9 * s.nStruct * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // the whole-structure path is re-tried only when a lattice actually changed (an MD driver passes the same
    // cell every step: a slot that was sent to the cell list stays there)
    if (s.h_lat.size() != (size_t)9 * s.nStruct || memcmp(s.h_lat.data(), latvecs, (size_t)9 * s.nStruct * sizeof(double)) != 0) {
      s.h_lat.assign(latvecs, latvecs + (size_t)9 * s.nStruct);   // overlaps the copies above
      s.structPath = 1;
    }
  }
  s.cellRc = -1.0; s.neighStale = true; s.featValid = false; s.geomEpoch++;   // maxNeigh stays as a capacity hint
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // caller may reuse its buffer on return
  return 0;
}

// ------------------------------------------------------------------------------------------
// ACSF configuration: grouping of the function list into radial groups and angular passes
// ------------------------------------------------------------------------------------------
extern "C" int fnetgpu_acsf_set(fnetgpu_ctx *ctx, int F, const int *type, const double *rcut,
                                const double *kappa, const double *rs, const double *eta,
                                const double *lambda, const double *xi, const int *atomid,
                                const int *atomicnumbers) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  if (F < 0) FNET_FAIL(ctx, "acsf_set: negative function count");
  AcsfTables &T = ctx->acsf;
  memset(&T, 0, sizeof(T));
  T.F = F;
  ctx->h_rgroups.clear(); ctx->h_apasses.clear(); ctx->maxSlots = 1;
  std::vector<int> rfeat; std::vector<double> rp1, rp2;
  // species codes
  std::vector<int> codes;
  auto code_of = [&](int z) -> int {
    for (size_t c = 0; c < codes.size(); c++) if (codes[c] == z) return (int)c;
    codes.push_back(z);
    return (int)codes.size() - 1;
  };
  double rcMax = 0.0;
  int anyAtomId = 0;
  typedef std::tuple<int, double, int, int> RKey;                       // type, rc, atomId, code
  typedef std::tuple<int, double, double, int, int, int> AKey;          // type, rc, eta, atomId, code1, code2
  std::vector<RKey> rkeys; std::vector<std::vector<int>> rmembers;
  std::vector<AKey> akeys; std::vector<std::vector<int>> amembers;
  for (int a = 0; a < F; a++) {
    if (type[a] < 1 || type[a] > 5) FNET_FAIL(ctx, "acsf_set: invalid function type (expected 1..5 for G1..G5)");
    if (!(rcut[a] > 0.0)) FNET_FAIL(ctx, "acsf_set: invalid cutoff");
    if (atomid[a] < 0) FNET_FAIL(ctx, "acsf_set: negative atomId");
    rcMax = std::max(rcMax, rcut[a]);
    if (atomid[a] > 0) anyAtomId = 1;
    const int z1 = atomicnumbers[2 * a], z2 = atomicnumbers[2 * a + 1];
    const bool resolved = !(z1 == 0 && z2 == 0);                        // acsf.F90:992-996
    if (type[a] <= 3) {
      RKey k(type[a], rcut[a], atomid[a], resolved ? code_of(z1) : -1);
      size_t g = 0;
      for (; g < rkeys.size(); g++) if (rkeys[g] == k) break;
      if (g == rkeys.size()) { rkeys.push_back(k); rmembers.emplace_back(); }
      rmembers[g].push_back(a);
    } else {
      int c1 = -1, c2 = -1;
      if (resolved) { c1 = code_of(z1); c2 = code_of(z2); }
      AKey k(type[a], rcut[a], eta[a], atomid[a], c1, c2);
      size_t g = 0;
      for (; g < akeys.size(); g++) if (akeys[g] == k) break;
      if (g == akeys.size()) { akeys.push_back(k); amembers.emplace_back(); }
      amembers[g].push_back(a);
    }
  }
  if ((int)codes.size() > FNET_MAX_CODES) FNET_FAIL(ctx, "acsf_set: too many distinct atomic numbers in species-resolved functions");
  T.nCodes = (int)codes.size();
  for (size_t c = 0; c < codes.size(); c++) T.zcodes[c] = codes[c];
  T.rcMax = rcMax; T.anyAtomId = anyAtomId;
  // radial groups: <= 4 chunks of FNET_RCHUNK functions each.  G2 members that form an arithmetic
  // rs-ladder with one eta (auto scheme, acsf.F90:320-336) become ladder groups evaluated by the
  // Gaussian recurrence (acsf.cuh); everything else goes into generic groups.
  const size_t RG_MAX = 4 * FNET_RCHUNK;
  auto emit_group = [&](const RKey &key, const std::vector<int> &m, size_t beg, size_t cnt, bool lad) {
    RadialGroup G; memset(&G, 0, sizeof(G));
    G.type = std::get<0>(key); G.rc = std::get<1>(key); G.atomId = std::get<2>(key); G.code = std::get<3>(key);
    G.fBeg = (int)rfeat.size(); G.fCnt = (int)cnt;
    int nch = ((int)cnt + FNET_RCHUNK - 1) / FNET_RCHUNK, p2 = 1;
    while (p2 < nch) p2 <<= 1;
    G.nChunksP2 = p2;
    if (lad) {
      const double d = rs[m[beg + 1]] - rs[m[beg]];
      G.ladder = 1; G.eta = eta[m[beg]]; G.rs0 = rs[m[beg]]; G.drs = d;
      for (int q = 0; q < FNET_RCHUNK - 1; q++) G.kk[q] = exp(-G.eta * d * d * (2.0 * q + 1.0));
    }
    for (size_t q = 0; q < cnt; q++) {
      int a = m[beg + q];
      rfeat.push_back(a);
      if (G.type == FNETGPU_G2) { rp1.push_back(eta[a]); rp2.push_back(rs[a]); }
      else if (G.type == FNETGPU_G3) { rp1.push_back(kappa[a]); rp2.push_back(0.0); }
      else { rp1.push_back(0.0); rp2.push_back(0.0); }
    }
    ctx->h_rgroups.push_back(G);
  };
  for (size_t g = 0; g < rkeys.size(); g++) {
    std::vector<int> m = rmembers[g], rest;
    if (std::get<0>(rkeys[g]) == FNETGPU_G2) {
      std::stable_sort(m.begin(), m.end(), [&](int a, int b) { return eta[a] != eta[b] ? eta[a] < eta[b] : rs[a] < rs[b]; });
      size_t beg = 0;
      while (beg < m.size()) {
        size_t run = 1;
        while (beg + run < m.size() && eta[m[beg + run]] == eta[m[beg]]) run++;
        bool lad = run >= 2;
        const double r0 = rs[m[beg]], d = lad ? rs[m[beg + 1]] - r0 : 0.0;
        lad = lad && d > 0.0;
        for (size_t q = 0; q < run && lad; q++) {
          const double expect = r0 + (double)q * d, v = rs[m[beg + q]];
          if (fabs(v - expect) > 8.0 * 2.220446049250313e-16 * std::max(fabs(v), fabs(d))) lad = false;
        }
        if (lad) {
          // chunks restart the recurrence from a direct exp, so groups may be cut anywhere
          for (size_t b2 = 0; b2 < run; b2 += RG_MAX) {
            size_t cnt = std::min(run - b2, RG_MAX);
            if (cnt >= 2) emit_group(rkeys[g], m, beg + b2, cnt, true);
            else rest.push_back(m[beg + b2]);
          }
        } else {
          for (size_t q = 0; q < run; q++) rest.push_back(m[beg + q]);
        }
        beg += run;
      }
    } else {
      rest = m;
    }
    for (size_t beg = 0; beg < rest.size(); beg += RG_MAX)
      emit_group(rkeys[g], rest, beg, std::min(rest.size() - beg, RG_MAX), false);
  }
  // angular passes: per key, split by lambda, sort by xi, cut into arithmetic ladders
  for (size_t g = 0; g < akeys.size(); g++) {
    std::vector<LadderSlot> slots;
    std::vector<int> m = amembers[g];
    std::stable_sort(m.begin(), m.end(), [&](int a, int b) {
      if (lambda[a] != lambda[b]) return lambda[a] > lambda[b];
      return xi[a] < xi[b];
    });
    size_t p = 0;
    while (p < m.size()) {
      LadderSlot sl; memset(&sl, 0, sizeof(sl));
      sl.lam = lambda[m[p]]; sl.xi0 = xi[m[p]]; sl.dxi = 0.0; sl.count = 1;
      size_t q = p + 1;
      if (q < m.size() && lambda[m[q]] == sl.lam) {
        sl.dxi = xi[m[q]] - xi[m[p]];
        while (q < m.size() && sl.count < FNET_LADDER && lambda[m[q]] == sl.lam) {
          double expect = sl.xi0 + sl.count * sl.dxi;
          if (fabs(xi[m[q]] - expect) > 1e-12 * std::max(1.0, fabs(expect))) break;
          sl.count++; q++;
        }
      }
      if (sl.count == 1) sl.dxi = 0.0;
      for (int f = 0; f < sl.count; f++) {
        int a = m[p + f];
        sl.feat[f] = a; sl.xi[f] = xi[a]; sl.pref[f] = pow(2.0, 1.0 - xi[a]);   // acsf.F90:1434,1490
      }
      slots.push_back(sl);
      p += sl.count;
    }
    for (size_t b = 0; b < slots.size(); b += FNET_SLOTS) {
      AngularPass P; memset(&P, 0, sizeof(P));
      P.type = std::get<0>(akeys[g]); P.rc = std::get<1>(akeys[g]); P.eta = std::get<2>(akeys[g]);
      P.atomId = std::get<3>(akeys[g]); P.code1 = std::get<4>(akeys[g]); P.code2 = std::get<5>(akeys[g]);
      P.same = (P.code1 == P.code2) ? 1 : 0;   // unresolved (-1,-1) or Z1 == Z2 (acsf.F90:1573)
      P.nSlots = (int)std::min(slots.size() - b, (size_t)FNET_SLOTS);
      for (int q = 0; q < P.nSlots; q++) {
        P.slot[q] = slots[b + q];
        P.slot[q].cont = 0;
        if (q > 0) {   // continues the running product of the previous slot (acsf.cuh angular_pass)
          const LadderSlot &pv = P.slot[q - 1];
          const LadderSlot &cu = P.slot[q];
          const double expect = pv.xi0 + FNET_LADDER * pv.dxi;
          if (pv.count == FNET_LADDER && pv.dxi != 0.0 && cu.lam == pv.lam &&
              fabs(cu.xi0 - expect) <= 1e-12 * std::max(1.0, fabs(expect)) &&
              (cu.count == 1 || fabs(cu.dxi - pv.dxi) <= 1e-12 * std::max(1.0, fabs(pv.dxi))))
            P.slot[q].cont = 1;
        }
      }
      ctx->maxSlots = std::max(ctx->maxSlots, P.nSlots);
      ctx->h_apasses.push_back(P);
    }
  }
  {   // straight-line variant of the pair loop (acsf.cuh angular_pass<NS, true>)
    const int ns = ctx->maxSlots <= 1 ? 1 : (ctx->maxSlots <= 2 ? 2 : 4);
    for (size_t k = 0; k < ctx->h_apasses.size(); k++) {
      AngularPass &Q = ctx->h_apasses[k];
      Q.keepFc = (k > 0 && Q.rc == ctx->h_apasses[k - 1].rc && Q.eta == ctx->h_apasses[k - 1].eta &&
                  Q.atomId == ctx->h_apasses[k - 1].atomId) ? 1 : 0;
    }
    for (AngularPass &P : ctx->h_apasses) {
      bool fast = P.type == FNETGPU_G5 && P.nSlots == ns;
      for (int q = 0; q < P.nSlots && fast; q++) fast = P.slot[q].cont == 0 && P.slot[q].xi0 == 1.0;
      P.fast = fast ? 1 : 0;
    }
    int rows = 8 * std::min(ns, 4);
    if (ns == 4) rows = 0;                       // 32 angular values per lane use the shuffle butterfly
    for (const RadialGroup &G : ctx->h_rgroups) if (G.ladder) rows = std::max(rows, 8 * G.nChunksP2);
    T.redRows = rows;
  }
  // ---- lean kernel tables (acsf_lean.cuh): every radial group a G2 ladder, every angular key a set of
  // lambda-groups that are fresh xi-ladders from xi = 1 with ONE common step delta (the automatic scheme,
  // acsf.F90:276-363, also after the species-resolved expansion), no atom-id scaling ----
  ctx->leanOK = false;
  std::vector<LeanRadial> lrad; std::vector<LeanPass> lpass;
  double leanDelta = 0.0;
  {
    bool ok = F > 0 && !anyAtomId;
    int maxChunks = 1;
    for (const RadialGroup &G : ctx->h_rgroups) {
      if (!G.ladder) { ok = false; break; }
      LeanRadial R; memset(&R, 0, sizeof(R));
      R.code = G.code; R.fBeg = G.fBeg; R.fCnt = G.fCnt; R.nch = G.nChunksP2;
      R.lgn = G.nChunksP2 == 1 ? 0 : (G.nChunksP2 == 2 ? 1 : 2);
      R.rc = G.rc; R.invrc = 1.0 / G.rc; R.eta = G.eta; R.rs0 = G.rs0; R.drs = G.drs;
      for (int q = 0; q < FNET_RCHUNK - 1; q++) R.kk[q] = G.kk[q];
      R.kk7 = exp(-G.eta * G.drs * G.drs * 15.0); R.c16 = exp(-G.eta * G.drs * G.drs * 16.0);
      maxChunks = std::max(maxChunks, G.nChunksP2);
      lrad.push_back(R);
    }
    // lambda-groups per key
    struct LamGroup { double lam; std::vector<int> m; };
    std::vector<std::vector<LamGroup>> keyGroups(akeys.size());
    bool haveDelta = false;
    size_t maxM = 1; size_t maxGroups = 1;
    for (size_t g = 0; g < akeys.size() && ok; g++) {
      if (std::get<0>(akeys[g]) != FNETGPU_G5) { ok = false; break; }
      std::vector<int> m = amembers[g];
      std::stable_sort(m.begin(), m.end(), [&](int a, int b) {
        if (lambda[a] != lambda[b]) return lambda[a] > lambda[b];
        return xi[a] < xi[b];
      });
      for (size_t p = 0; p < m.size();) {
        LamGroup L; L.lam = lambda[m[p]];
        size_t q = p;
        while (q < m.size() && lambda[m[q]] == L.lam) L.m.push_back(m[q++]);
        if (!(L.lam >= -1.0) || xi[L.m[0]] != 1.0) { ok = false; break; }
        for (size_t f = 1; f < L.m.size() && ok; f++) {
          const double d = (xi[L.m[f]] - 1.0) / (double)f;
          if (!haveDelta) { leanDelta = d; haveDelta = true; }
          if (!(d > 0.0) || fabs(d - leanDelta) > 1e-12 * std::max(1.0, fabs(leanDelta))) ok = false;
        }
        maxM = std::max(maxM, L.m.size());
        keyGroups[g].push_back(L);
        p = q;
      }
      maxGroups = std::max(maxGroups, keyGroups[g].size());
    }
    if (ok && haveDelta) {   // degree-5 binomial series of (1 + r)^delta, |r| <= 2^-9: truncation binom(delta, 6) 2^-54
      long double b6 = 1.0L;
      for (int k = 0; k < 6; k++) b6 *= ((long double)leanDelta - k) / (long double)(k + 1);
      if (fabsl(b6) > 18.0L) ok = false;
    }
    if (ok) {
      const int NC = maxM <= 8 ? 1 : (maxM <= 16 ? 2 : 4);
      const int NL = NC == 4 ? 1 : 2;
      (void)maxGroups;
      const int blk = FNET_LADDER * NC;
      for (size_t g = 0; g < akeys.size(); g++) {
        const std::vector<LamGroup> &LG = keyGroups[g];
        for (size_t l0 = 0; l0 < LG.size(); l0 += NL) {
          size_t longest = 0;
          for (int l = 0; l < NL && l0 + l < LG.size(); l++) longest = std::max(longest, LG[l0 + l].m.size());
          for (size_t m0 = 0; m0 < longest; m0 += blk) {
            LeanPass P; memset(&P, 0, sizeof(P));
            P.code1 = std::get<4>(akeys[g]); P.code2 = std::get<5>(akeys[g]);
            P.same = (P.code1 == P.code2) ? 1 : 0;
            P.rc = std::get<1>(akeys[g]); P.invrc = 1.0 / P.rc; P.eta = std::get<2>(akeys[g]);
            P.m0 = (int)m0;
            for (int e = 0; e < FNET_LEAN_MAXACC; e++) P.feat[e] = -1;
            for (int l = 0; l < NL; l++) {
              if (l0 + l >= LG.size()) { P.lam[l] = 0.0; continue; }
              const LamGroup &L = LG[l0 + l];
              P.lam[l] = L.lam;
              for (int f = 0; f < blk; f++) {
                const size_t mi = m0 + f;
                if (mi >= L.m.size()) break;
                const int a = L.m[mi], e = l * blk + f;
                const double x = xi[a], pre = pow(2.0, 1.0 - x);        // acsf.F90:1434,1490
                P.feat[e] = a;
                P.pref[e] = P.same ? 2.0 * pre : pre;
                if (P.same) {   // diagonal j == k: cos = 1 - eps (acsf_lean.cuh)
                  if (L.lam > -1.0) { P.dA[e] = pre * pow(1.0 + L.lam, x); P.dB[e] = -P.dA[e] * x * L.lam / (1.0 + L.lam); }
                  else { P.dA[e] = 0.0; P.dB[e] = (x == 1.0) ? pre : 0.0; }
                }
              }
            }
            const bool first = lpass.empty();
            P.recomp = (!first && (P.rc != lpass.back().rc || P.eta != lpass.back().eta)) ? 1 : 0;
            lpass.push_back(P);
          }
        }
      }
      ctx->leanNL = NL; ctx->leanNC = NC;
      ctx->leanSorted = T.nCodes > 0;
      LeanTables &LT = ctx->lean;
      memset(&LT, 0, sizeof(LT));
      LT.nRadial = (int)lrad.size(); LT.nPasses = (int)lpass.size();
      LT.redRows = std::max(FNET_LADDER * NL * NC, FNET_RCHUNK);
      { const size_t tb = lpass.size() * sizeof(LeanPass) + lrad.size() * sizeof(LeanRadial);
        LT.stageBytes = tb <= 8192 ? (int)tb : 0; }
      (void)maxChunks;
      if (!lpass.empty()) { LT.rcShared = lpass[0].rc; LT.etaShared = lpass[0].eta; }
      else { LT.rcShared = lrad.empty() ? rcMax : lrad[0].rc; LT.etaShared = 0.0; }
      LT.invrcShared = 1.0 / LT.rcShared;
      for (LeanRadial &R : lrad) R.sharedFc = (R.rc == LT.rcShared) ? 1 : 0;
      {
        long double c = 1.0L;
        for (int k = 1; k <= 5; k++) { c *= ((long double)leanDelta - (k - 1)) / (long double)k; LT.powC[k - 1] = (double)c; }
      }
      std::vector<double> pt(FNET_POW_DOUBLES);
      for (int i = 0; i < FNET_POW_TAB_N; i++) {
        const long double mid = 1.0L + ((long double)i + 0.5L) / (long double)FNET_POW_TAB_N;
        const double invc = (double)(1.0L / mid);
        pt[2 * i] = invc;
        pt[2 * i + 1] = (double)powl(1.0L / (long double)invc, (long double)leanDelta);
      }
      for (int k = FNET_POW_KMIN; k <= 1; k++) pt[2 * FNET_POW_TAB_N + (k - FNET_POW_KMIN)] = (double)powl(2.0L, (long double)k * (long double)leanDelta);
      std::vector<unsigned short> pairs(FNET_PAIR_TAB_N + 1);
      for (int k = 1; k <= FNET_PAIR_TAB_MAXN; k++)
        for (int j = 0; j < k; j++) {
          const int p = k * (k - 1) / 2 + j;
          if (p < FNET_PAIR_TAB_N) pairs[p] = (unsigned short)(j | (k << 8));
        }
      if (dev_upload(ctx, &ctx->d_lrad, lrad.data(), lrad.size())) return 1;
      if (dev_upload(ctx, &ctx->d_lpass, lpass.data(), lpass.size())) return 1;
      if (dev_upload(ctx, &ctx->d_powtab, pt.data(), pt.size())) return 1;
      if (dev_upload(ctx, &ctx->d_pairtab, pairs.data(), pairs.size())) return 1;
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors above are stack-lifetime
      LT.rad = ctx->d_lrad; LT.pass = ctx->d_lpass; LT.powtab = ctx->d_powtab; LT.pairtab = ctx->d_pairtab;
      ctx->leanOK = true;
    }
  }
  T.nRadialGroups = (int)ctx->h_rgroups.size();
  T.nAngularPasses = (int)ctx->h_apasses.size();
  if (dev_upload(ctx, &ctx->d_rgroups, ctx->h_rgroups.data(), ctx->h_rgroups.size())) return 1;
  if (dev_upload(ctx, &ctx->d_rfeat, rfeat.data(), rfeat.size())) return 1;
  if (dev_upload(ctx, &ctx->d_rp1, rp1.data(), rp1.size())) return 1;
  if (dev_upload(ctx, &ctx->d_rp2, rp2.data(), rp2.size())) return 1;
  if (dev_upload(ctx, &ctx->d_apasses, ctx->h_apasses.data(), ctx->h_apasses.size())) return 1;
  T.rgroups = ctx->d_rgroups; T.rfeat = ctx->d_rfeat; T.rp1 = ctx->d_rp1; T.rp2 = ctx->d_rp2;
  T.apasses = ctx->d_apasses;
  if (dev_alloc(ctx, &ctx->d_zprec, (size_t)2 * std::max(F, 1))) return 1;
  ctx->haveZ = false;
  ctx->acsfSet = true; ctx->acsfEpoch++;
  for (int i = 0; i < FNETGPU_MAX_SLOTS; i++) { ctx->slots[i].featValid = false; ctx->slots[i].maxNeigh = -1; ctx->slots[i].maxCand = -1; ctx->slots[i].okEpoch = 0; }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fnetgpu_features_config(fnetgpu_ctx *ctx, int nExtSel, const int *extIndices) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  ctx->extIdx.clear();
  for (int e = 0; e < nExtSel; e++) {
    if (extIndices[e] < 1) FNET_FAIL(ctx, "features_config: external feature indices are 1-based");
    ctx->extIdx.push_back(extIndices[e] - 1);
  }
  if (dev_upload(ctx, &ctx->d_extIdx, ctx->extIdx.data(), ctx->extIdx.size())) return 1;
  for (int i = 0; i < FNETGPU_MAX_SLOTS; i++) ctx->slots[i].featValid = false;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// cell list construction for a slot (host: per-structure bin geometry; device: binning)
// ------------------------------------------------------------------------------------------
static bool invert3(const double *L /* L[3*k+c] = lat vec k comp c */, double *inv /* inv[3*k+c] */) {
  // M[c][k] = L[3k+c]; inv = M^-1, inv[3*k+c] = (M^-1)[k][c]
  double m[3][3];
  for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) m[c][k] = L[3 * k + c];
  double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
               m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  if (fabs(det) < 1e-12) return false;   // fnetdata.F90:1155-1158
  double id = 1.0 / det;
  double r[3][3];
  r[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id;
  r[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  r[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  r[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id;
  r[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  r[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  r[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id;
  r[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  r[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) inv[3 * k + c] = r[k][c];
  return true;
}

static int ensure_cells(fnetgpu_ctx *ctx, Slot &s, double rc, const double *h_coords_or_null) {
  if (s.cellRc == rc) return 0;
  if (!s.d_coords) FNET_FAIL(ctx, "slot has no geometry (coords were not uploaded)");
  // non-periodic structures need their bounding box: fetch coordinates once
  std::vector<double> hc;
  bool anyCluster = false;
  for (int st = 0; st < s.nStruct; st++) if (!s.h_periodic[st]) { anyCluster = true; break; }
  if (anyCluster) {
    hc.resize((size_t)3 * s.N);
    CUDA_TRY(ctx, cudaMemcpyAsync(hc.data(), s.d_coords, hc.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  (void)h_coords_or_null;
  std::vector<StructInfo> si(s.nStruct);
  std::vector<int> binStruct;
  long long totalBins = 0;
  s.maxCells = 1;
  for (int st = 0; st < s.nStruct; st++) {
    StructInfo &S = si[st];
    memset(&S, 0, sizeof(S));
    S.atomBeg = s.h_offsets[st]; S.atomEnd = s.h_offsets[st + 1];
    const int nAt = S.atomEnd - S.atomBeg;
    S.periodic = s.h_periodic[st];
    if (S.periodic) {
      memcpy(S.lat, &s.h_lat[(size_t)9 * st], 9 * sizeof(double));
      if (!invert3(S.lat, S.inv)) FNET_FAIL(ctx, "dependent lattice vectors");
    } else {
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int i = S.atomBeg; i < S.atomEnd; i++)
        for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], hc[3 * (size_t)i + c]); hi[c] = std::max(hi[c], hc[3 * (size_t)i + c]); }
      for (int c = 0; c < 3; c++) {
        double len = std::max(hi[c] - lo[c], 1e-3) * (1.0 + 1e-9) + 1e-9;
        S.lo[c] = lo[c];
        S.lat[3 * c + c] = len;
        S.inv[3 * c + c] = 1.0 / len;
      }
    }
    for (int k = 0; k < 3; k++) {
      double bn = sqrt(S.inv[3 * k] * S.inv[3 * k] + S.inv[3 * k + 1] * S.inv[3 * k + 1] + S.inv[3 * k + 2] * S.inv[3 * k + 2]);
      double h = 1.0 / bn;                       // spacing of the lattice planes normal to b_k
      int nb = (int)floor(h / rc);
      if (nb < 1) nb = 1;
      S.nb[k] = nb;
      S.D[k] = (h / nb >= rc) ? 1 : (int)ceil(rc / h);
      if (!S.periodic) S.D[k] = std::min(S.D[k], 1);
    }
    // keep the number of bins comparable to the number of atoms
    while ((long long)S.nb[0] * S.nb[1] * S.nb[2] > 2LL * nAt + 8) {
      int k = 0;
      if (S.nb[1] > S.nb[k]) k = 1;
      if (S.nb[2] > S.nb[k]) k = 2;
      if (S.nb[k] == 1) break;
      S.nb[k] = (S.nb[k] + 1) / 2;
    }
    S.binBase = (int)totalBins;
    totalBins += (long long)S.nb[0] * S.nb[1] * S.nb[2];
    if (totalBins > 2000000000LL) FNET_FAIL(ctx, "cell list too large");
    binStruct.resize((size_t)totalBins, st);
    s.maxCells = std::max(s.maxCells, (2 * S.D[0] + 1) * (2 * S.D[1] + 1) * (2 * S.D[2] + 1));
  }
  s.totalBins = (int)totalBins;
  if (dev_reserve(ctx, &s.d_sinfo, &s.capSinfo, si.size())) return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(s.d_sinfo, si.data(), si.size() * sizeof(StructInfo), cudaMemcpyHostToDevice, ctx->stream));
  if (dev_reserve(ctx, &s.d_binStruct, &s.capBinStruct, binStruct.size())) return 1;
  CUDA_TRY(ctx, cudaMemcpyAsync(s.d_binStruct, binStruct.data(), binStruct.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // si / binStruct are stack-lifetime host vectors
  if (dev_reserve(ctx, &s.d_fpos, &s.capFpos, (size_t)3 * s.N)) return 1;
  if (dev_reserve(ctx, &s.d_crec, &s.capCrec, (size_t)s.N)) return 1;
  if (dev_reserve(ctx, &s.d_atomCell, &s.capAtomCell, (size_t)s.N)) return 1;
  if (dev_reserve(ctx, &s.d_cellAtoms, &s.capCellAtoms, (size_t)s.N)) return 1;
  if (dev_reserve(ctx, &s.d_cellStart, &s.capCellStart, (size_t)s.totalBins + 1)) return 1;
  if (dev_reserve(ctx, &s.d_cellCount, &s.capCellCount, (size_t)s.totalBins)) return 1;
  CUDA_TRY(ctx, cudaMemsetAsync(s.d_cellCount, 0, (size_t)s.totalBins * sizeof(int), ctx->stream));
  const int B = 256;
  LAUNCH(ctx, K_BIN_COUNT, (k_bin_count<<<(s.N + B - 1) / B, B, 0, ctx->stream>>>(s.N, s.d_coords, s.d_structOf, s.d_sinfo, s.d_fpos, s.d_atomCell, s.d_cellCount)));
  LAUNCH(ctx, K_BIN_SCAN, (k_bin_scan<<<1, 1024, 0, ctx->stream>>>(s.totalBins, s.d_cellCount, s.d_cellStart, ctx->d_flags + 8)));
  CUDA_TRY(ctx, cudaMemsetAsync(s.d_cellCount, 0, (size_t)s.totalBins * sizeof(int), ctx->stream));
  LAUNCH(ctx, K_BIN_FILL, (k_bin_fill<<<(s.N + B - 1) / B, B, 0, ctx->stream>>>(s.N, s.d_atomCell, s.d_cellStart, s.d_cellCount, s.d_cellAtoms)));
  LAUNCH(ctx, K_BIN_SORT, (k_bin_sort<<<(s.totalBins + B - 1) / B, B, 0, ctx->stream>>>(s.totalBins, s.d_cellStart, s.d_cellAtoms, s.d_fpos, s.d_atnum, s.d_crec)));
  s.cellRc = rc;
  s.neighStale = true;
  return 0;
}

static int ensure_neigh_count(fnetgpu_ctx *ctx, Slot &s) {
  if (s.maxNeigh >= 0 && !s.neighStale) return 0;
  const double rc = ctx->acsf.rcMax;
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), ctx->stream));
  const int B = 128;
  const int grid = (int)(((long long)s.N * 32 + B - 1) / B);
  LAUNCH(ctx, K_NEIGH_COUNT, (k_neigh_count<<<grid, B, 0, ctx->stream>>>(s.N, nullptr, s.d_binStruct, s.d_sinfo, s.d_atomCell, s.d_cellStart, s.d_crec, rc * rc, ctx->d_flags)));
  int h[16];
  CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_flags, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  s.maxNeigh = h[0];
  unsigned long long tot;
  memcpy(&tot, &h[2], sizeof(tot));
  s.meanNeigh = (double)tot / (double)s.N;
  s.maxCand = h[5];
  s.maxBinPop = h[8 + 4];
  s.neighStale = false;
  return 0;
}

// launch geometry shared by the ACSF value and force kernels (one CTA per bin x split)
struct AcsfLaunch { int cap, capC, wpb, nSplit; bool staged; int path; size_t smem; dim3 grid; int stBase = 0; bool lean = false; int G = 1; bool local = false; };
// the whole-structure path (cells.cuh): small structures, lattice check passed so far
static bool use_struct_path(const fnetgpu_ctx *ctx, const Slot &s) {
  return !ctx->acsfPathCells && s.structPath && s.maxAtoms <= FNET_STRUCT_MAX_ATOMS && s.d_coords && s.d_lat;
}
static inline int struct_cap(const Slot &s) {   // neighbours per atom <= atoms - 1 under the minimum-image convention
  const int hint = s.maxNeigh > 0 ? s.maxNeigh : 32;
  return std::max(32, (std::min(hint, std::max(s.maxAtoms - 1, 1)) + 31) & ~31);
}
static GeomArgs geom_args(const Slot &s) {
  GeomArgs g;
  g.binStruct = s.d_binStruct; g.sinfo = s.d_sinfo; g.cellStart = s.d_cellStart; g.crec = s.d_crec;
  g.offsets = s.d_offsets; g.coords = s.d_coords; g.lat = s.d_lat; g.periodic = s.d_periodic; g.atnum = s.d_atnum;
  g.stBase = 0;
  return g;
}
static int plan_struct_launch(fnetgpu_ctx *ctx, const Slot &s, size_t warpBytes, AcsfLaunch &L, int cap = -1, size_t extraCta = 0,
                              bool wholeStructure = false, int atomsPerIter = 1) {
  L.path = FNET_PATH_STRUCT; L.staged = false;
  L.cap = cap > 0 ? cap : struct_cap(s);
  L.capC = (s.maxAtoms + 31) & ~31;
  L.wpb = 4;
  const size_t prefix = acsf_cta_prefix_bytes(L.capC, 2) + extraCta;
  L.smem = prefix + warpBytes * L.wpb;
  while (L.smem > 220 * 1024 && L.wpb > 1) { L.wpb >>= 1; L.smem = prefix + warpBytes * L.wpb; }
  if (L.smem > 220 * 1024) FNET_FAIL(ctx, "too many neighbours per atom for the shared-memory neighbour buffers");
  // <= 32 central atoms per warp: the CTA prologue (structure, lattice inverse, power / pair / pass tables -> shared
  // memory, two barriers) was 15 % of the lean kernel's stall samples at 8 atoms per warp (ncu source page); same-box
  // A/B 8 -> 16 -> 32: C2 ACSF 0.842 -> 0.799 -> 0.800 ms, C3 15.38 -> 14.90 -> 14.74 ms.  More splits when there are
  // too few structures to fill the GPU.  FNETGPU_ACSF_ATOMS_PER_WARP overrides (A/B).
  static const int apw = [] { const char *e = getenv("FNETGPU_ACSF_ATOMS_PER_WARP"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 32; }();
  int nSplit = (s.maxAtoms + apw * L.wpb - 1) / (apw * L.wpb);
  const long long want = 8LL * ctx->nSM;
  if ((long long)s.nStruct * nSplit < want)
    nSplit = (int)std::min<long long>((s.maxAtoms + L.wpb - 1) / L.wpb, (want + s.nStruct - 1) / s.nStruct);
  if (wholeStructure) {
    // deterministic force accumulation: one CTA per structure; only when there are too few structures to fill the
    // GPU (an MD step of one cell) the atoms are split over CTAs, whose partial forces k_force_reduce sums in order
    const long long wantF = 2LL * ctx->nSM;
    nSplit = 1;
    if ((long long)s.nStruct < wantF)
      nSplit = (int)std::min<long long>((s.maxAtoms + L.wpb * atomsPerIter - 1) / (L.wpb * atomsPerIter), (wantF + s.nStruct - 1) / s.nStruct);
  }
  L.nSplit = std::max(1, std::min(nSplit, 65535));
  L.grid = dim3(s.nStruct, L.nSplit);
  return 0;
}
static int plan_acsf_launch(fnetgpu_ctx *ctx, const Slot &s, size_t warpBytes, AcsfLaunch &L, int cap = -1, size_t extraCta = 0) {
  L.cap = cap > 0 ? cap : std::max(32, (s.maxNeigh + 31) & ~31);
  L.staged = s.maxCells <= FNET_MAX_NCELLS && s.maxCand >= 0 && s.maxCand <= 1536;
  L.capC = L.staged ? ((s.maxCand + s.maxCand / 8 + 31) & ~31) : 0;
  L.wpb = 4;
  const size_t prefix = (L.staged ? acsf_cta_prefix_bytes(L.capC) : acsf_cta_prefix_bytes(0, 0)) + extraCta;
  L.smem = prefix + warpBytes * L.wpb;
  while (L.smem > 220 * 1024 && L.wpb > 1) { L.wpb >>= 1; L.smem = prefix + warpBytes * L.wpb; }
  if (L.smem > 220 * 1024 && L.staged) { L.staged = false; L.capC = 0; L.wpb = 4; L.smem = acsf_cta_prefix_bytes(0, 0) + extraCta + warpBytes * L.wpb;
    while (L.smem > 220 * 1024 && L.wpb > 1) { L.wpb >>= 1; L.smem = acsf_cta_prefix_bytes(0, 0) + extraCta + warpBytes * L.wpb; } }
  if (L.smem > 220 * 1024) FNET_FAIL(ctx, "too many neighbours per atom for the shared-memory neighbour buffers");
  // atoms of a bin are split over blockIdx.y so that a CTA sees ~4 rounds of its warps and the
  // grid still fills the GPU when a structure has few, crowded bins
  const int pop = std::max(1, s.maxBinPop);
  int nSplit = (pop + 4 * L.wpb - 1) / (4 * L.wpb);
  const long long want = 8LL * ctx->nSM;
  if ((long long)s.totalBins * nSplit < want) nSplit = (int)std::min<long long>((pop + L.wpb - 1) / L.wpb, (want + s.totalBins - 1) / s.totalBins);
  L.nSplit = std::max(1, std::min(nSplit, 65535));
  L.path = L.staged ? FNET_PATH_STAGED : FNET_PATH_DIRECT;
  L.grid = dim3(s.totalBins, L.nSplit);
  return 0;
}

// launch plan of the VALUE kernel: the lean kernel (acsf_lean.cuh) when the configuration is an
// automatic-scheme one and the candidates can be staged, else k_acsf
static bool use_lean(const fnetgpu_ctx *ctx) { return ctx->leanOK && !ctx->acsfGeneric; }
static int plan_values(fnetgpu_ctx *ctx, const Slot &s, bool structPath, AcsfLaunch &L) {
  const AcsfTables &T = ctx->acsf;
  if (use_lean(ctx)) {
    // capacity: neighbours + the dummy neighbour of the pair walk, in steps of 8; central atoms per warp from
    // the neighbour count (FNETGPU_LEAN_G overrides: A/B)
    int hint = s.maxNeigh > 0 ? s.maxNeigh : 32;
    if (structPath) hint = std::min(hint, std::max(s.maxAtoms - 1, 1));
    const int cap = std::max(8, (hint + 1 + 7) & ~7);
    static const int gEnv = [] { const char *e = getenv("FNETGPU_LEAN_G"); const int v = e ? atoi(e) : 0; return (v == 1 || v == 2 || v == 4) ? v : 0; }();
    int G = ctx->leanSorted ? (hint <= 48 ? 2 : 1) : (hint <= 20 ? 4 : (hint <= 64 ? 2 : 1));
    if (s.N <= 2 * ctx->nSM) G = 1;          // a handful of atoms (an MD step of one cell): latency, not lane occupancy -- a warp per atom
    if (gEnv) G = (ctx->leanSorted && gEnv == 4) ? 2 : gEnv;
    const size_t extra = lean_cta_extra_bytes(T.F, cap, ctx->lean.stageBytes);
    const bool f32a = ctx->precision == 32;   // FP32 pair arithmetic (acsf_lean.cuh)
    for (; G >= 1; G >>= 1) {
      AcsfLaunch Q;
      const size_t wb = lean_warp_smem_bytes(cap, T.F, ctx->lean.redRows, ctx->leanSorted, G, f32a);
      const int rc = structPath ? plan_struct_launch(ctx, s, wb, Q, cap, extra) : plan_acsf_launch(ctx, s, wb, Q, cap, extra);
      // bins too crowded to stage their candidates (PATH_DIRECT): the lean kernel walks the cell list with a full warp per atom
      if (rc == 0 && Q.wpb == 4 && (Q.path != FNET_PATH_DIRECT || G == 1)) { L = Q; L.lean = true; L.G = G; return 0; }
      ctx->err.clear();
    }
  }
  L.lean = false;
  if (structPath) return plan_struct_launch(ctx, s, acsf_warp_smem_bytes(struct_cap(s), T.F, T.redRows), L);
  return plan_acsf_launch(ctx, s, acsf_warp_smem_bytes(std::max(32, (s.maxNeigh + 31) & ~31), T.F, T.redRows), L);
}

// launch plan of the FORCE kernel: k_acsf_force_lean for automatic-scheme configurations (whole-structure
// path: one CTA per structure, deterministic shared-memory accumulation), else k_acsf_force
static int lean_M(const fnetgpu_ctx *ctx) { return FNET_LADDER * ctx->leanNL * ctx->leanNC; }
static int plan_forces(fnetgpu_ctx *ctx, const Slot &s, bool structPath, AcsfLaunch &L) {
  const AcsfTables &T = ctx->acsf;
  if (use_lean(ctx)) {
    int hint = s.maxNeigh > 0 ? s.maxNeigh : 32;
    if (structPath) hint = std::min(hint, std::max(s.maxAtoms - 1, 1));
    const int cap = std::max(8, (hint + 1 + 7) & ~7);
    static const int gEnv = [] { const char *e = getenv("FNETGPU_LEAN_G"); const int v = e ? atoi(e) : 0; return (v == 1 || v == 2 || v == 4) ? v : 0; }();
    int G = ctx->leanSorted ? (hint <= 48 ? 2 : 1) : (hint <= 20 ? 4 : (hint <= 64 ? 2 : 1));
    if (s.N <= 2 * ctx->nSM) G = 1;          // a handful of atoms (an MD step of one cell): latency, not lane occupancy -- a warp per atom
    if (gEnv) G = (ctx->leanSorted && gEnv == 4) ? 2 : gEnv;
    const size_t extra = force_lean_cta_extra_bytes(ctx->lean.stageBytes);
    const int localAtoms = structPath ? ((s.maxAtoms + 1) & ~1) : 0;
    for (; G >= 1; G >>= 1) {
      AcsfLaunch Q;
      const size_t wb = force_lean_warp_bytes(cap, T.F, lean_M(ctx), ctx->leanSorted, G, localAtoms);
      const int rc = structPath ? plan_struct_launch(ctx, s, wb, Q, cap, extra, true, G) : plan_acsf_launch(ctx, s, wb, Q, cap, extra);
      if (rc == 0 && Q.path != FNET_PATH_DIRECT && Q.wpb == 4) { L = Q; L.lean = true; L.G = G; L.local = structPath; return 0; }
      ctx->err.clear();
    }
  }
  L.lean = false; L.local = false;
  if (structPath) return plan_struct_launch(ctx, s, force_warp_smem_bytes(struct_cap(s), T.F), L);
  return plan_acsf_launch(ctx, s, force_warp_smem_bytes(std::max(32, (s.maxNeigh + 31) & ~31), T.F), L);
}

extern "C" int fnetgpu_acsf_path_set(fnetgpu_ctx *ctx, int mode) {
  CHECK_CTX(ctx);
  if (mode != 0 && mode != 1) FNET_FAIL(ctx, "acsf_path_set: mode must be 0 (auto) or 1 (cell list)");
  ctx->acsfPathCells = mode;
  return 0;
}
extern "C" int fnetgpu_acsf_kernel_set(fnetgpu_ctx *ctx, int mode) {
  CHECK_CTX(ctx);
  if (mode != 0 && mode != 1) FNET_FAIL(ctx, "acsf_kernel_set: mode must be 0 (auto: lean kernel for automatic-scheme configurations) or 1 (always k_acsf)");
  ctx->acsfGeneric = mode;
  return 0;
}
// 1: the value kernel of the current configuration is k_acsf_lean, 0: k_acsf, -1: no configuration
extern "C" int fnetgpu_acsf_kernel_get(const fnetgpu_ctx *ctx) {
  if (!ctx || !ctx->acsfSet) return -1;
  return (ctx->leanOK && !ctx->acsfGeneric) ? 1 : 0;
}
extern "C" int fnetgpu_mlp_path_set(fnetgpu_ctx *ctx, int mode) {
  CHECK_CTX(ctx);
  if (mode < 0 || mode > 2) FNET_FAIL(ctx, "mlp_path_set: mode must be 0 (auto), 1 (register-tiled kernels) or 2 (DMMA without fused sums)");
  ctx->mlpLegacy = (mode == 1) ? 1 : 0;
  ctx->mlpNoFuse = (mode == 2) ? 1 : 0;
  return 0;
}
extern "C" int fnetgpu_mlp_path_get(const fnetgpu_ctx *ctx) {
  if (!ctx || !ctx->netSet) return -1;
  return (ctx->precision == 64 && !ctx->mlpLegacy && bpnn_mma_fits(ctx->net)) ? 1 : 0;
}
extern "C" int fnetgpu_grad_launch_info(const fnetgpu_ctx *ctx, int slot, int *info) {
  if (!ctx || slot < 0 || slot >= FNETGPU_MAX_SLOTS || !ctx->slots[slot].used || !info) return 1;
  for (int k = 0; k < 4; k++) info[k] = ctx->slots[slot].lastGrad[k];
  return 0;
}
extern "C" int fnetgpu_acsf_path_get(const fnetgpu_ctx *ctx, int slot) {
  if (!ctx || slot < 0 || slot >= FNETGPU_MAX_SLOTS || !ctx->slots[slot].used) return -1;
  return ctx->slots[slot].lastPath;
}

// what the last ACSF value launch of the slot used: info[0] = 1 k_acsf_lean / 0 k_acsf, [1] = central atoms per
// warp, [2] = neighbour capacity, [3] = staged candidates capacity, [4] = path, [5] = dynamic shared memory (bytes)
extern "C" int fnetgpu_acsf_launch_info(const fnetgpu_ctx *ctx, int slot, int *info) {
  if (!ctx || slot < 0 || slot >= FNETGPU_MAX_SLOTS || !ctx->slots[slot].used || !info) return 1;
  for (int k = 0; k < 6; k++) info[k] = ctx->slots[slot].lastLaunch[k];
  return 0;
}

extern "C" int fnetgpu_max_neighbors(fnetgpu_ctx *ctx, int slot, int *maxNeigh, double *meanNeigh) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used || !ctx->acsfSet || ctx->acsf.F == 0) FNET_FAIL(ctx, "max_neighbors: need a dataset and an ACSF configuration");
  cudaSetDevice(ctx->device);
  if (ensure_cells(ctx, s, ctx->acsf.rcMax, nullptr)) return 1;
  if (ensure_neigh_count(ctx, s)) return 1;
  if (maxNeigh) *maxNeigh = s.maxNeigh;
  if (meanNeigh) *meanNeigh = s.meanNeigh;
  return 0;
}

// ------------------------------------------------------------------------------------------
// multi-GPU plumbing: NCCL loaded lazily (dlopen) so the library has no link-time dependency
// ------------------------------------------------------------------------------------------
struct NcclUid { char internal[128]; };
typedef int (*nccl_getuid_t)(NcclUid *);
typedef int (*nccl_initrank_t)(void **, int, NcclUid, int);
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_t)(int);
static void *g_nccl = nullptr;
static void *nccl_handle() {
  if (!g_nccl) g_nccl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl) g_nccl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return g_nccl;
}

extern "C" int fnetgpu_comm_unique_id(char *id) {
  void *h = nccl_handle();
  if (!h) { g_err = std::string("cannot load libnccl: ") + dlerror(); return 1; }
  nccl_getuid_t f = (nccl_getuid_t)dlsym(h, "ncclGetUniqueId");
  NcclUid u;
  if (!f || f(&u) != 0) { g_err = "ncclGetUniqueId failed"; return 1; }
  memcpy(id, u.internal, FNETGPU_UNIQUE_ID_BYTES);
  return 0;
}

extern "C" int fnetgpu_comm_init(fnetgpu_ctx *ctx, int nRanks, int rank, const char *id) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  if (nRanks < 1 || rank < 0 || rank >= nRanks) FNET_FAIL(ctx, "comm_init: bad rank / size");
  void *h = nccl_handle();
  if (!h) FNET_FAIL(ctx, std::string("cannot load libnccl: ") + dlerror());
  nccl_initrank_t f = (nccl_initrank_t)dlsym(h, "ncclCommInitRank");
  if (!f) FNET_FAIL(ctx, "ncclCommInitRank not found");
  NcclUid u;
  memcpy(u.internal, id, FNETGPU_UNIQUE_ID_BYTES);
  void *comm = nullptr;
  int rc = f(&comm, nRanks, u, rank);
  if (rc != 0) FNET_FAIL(ctx, "ncclCommInitRank failed (" + std::to_string(rc) + ")");
  ctx->nccl = h; ctx->comm = comm; ctx->nRanks = nRanks; ctx->rank = rank;
  return 0;
}

// in-place sum all-reduce of n doubles on the library's stream (no-op for a single rank)
static int allreduce_sum(fnetgpu_ctx *ctx, double *d_buf, size_t n, cudaStream_t stream = nullptr) {
  if (ctx->nRanks <= 1 || !ctx->comm) return 0;
  static nccl_allreduce_t f = nullptr;
  if (!f) f = (nccl_allreduce_t)dlsym(ctx->nccl, "ncclAllReduce");
  if (!f) FNET_FAIL(ctx, "ncclAllReduce not found");
  int rc = f(d_buf, d_buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->comm, stream ? stream : ctx->stream);
  if (rc != 0) FNET_FAIL(ctx, "ncclAllReduce failed (" + std::to_string(rc) + ")");
  ctx->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------
// ACSF calculation
// ------------------------------------------------------------------------------------------
// one launch of the ACSF value kernel for the planned geometry path (no flag read-back)
template <typename real>
static int launch_acsf_values(fnetgpu_ctx *ctx, Slot &s, const AcsfLaunch &L, const double *zp) {
  const AcsfTables &T = ctx->acsf;
  const int nExtSel = (int)ctx->extIdx.size();
  const int nFeat = T.F + nExtSel;
  real *feat = (real *)s.d_feat;
  GeomArgs geo = geom_args(s);
  geo.stBase = L.stBase;
#define FNET_ACSF_LAUNCH(NS, PATH)                                                                             \
  do {                                                                                                         \
    CUDA_TRY(ctx, fnet_smem_attr(k_acsf<real, NS, PATH>, L.smem)); \
    LAUNCH(ctx, K_ACSF, (k_acsf<real, NS, PATH><<<L.grid, L.wpb * 32, L.smem, ctx->stream>>>(                   \
                            L.nSplit, geo, s.nExt, s.d_ext, T, L.cap, L.capC, feat, nFeat, zp, nExtSel,        \
                            ctx->d_extIdx, ctx->d_flags)));                                                    \
  } while (0)
#define FNET_ACSF_LAUNCH_NS(PATH)                                                                              \
  do { if (ns == 1) FNET_ACSF_LAUNCH(1, PATH); else if (ns == 2) FNET_ACSF_LAUNCH(2, PATH); else FNET_ACSF_LAUNCH(4, PATH); } while (0)
  s.lastLaunch[0] = L.lean ? 1 : 0; s.lastLaunch[1] = L.lean ? L.G : 1; s.lastLaunch[2] = L.cap; s.lastLaunch[3] = L.capC;
  s.lastLaunch[4] = L.path; s.lastLaunch[5] = (int)L.smem;
  if (L.lean) {
    const LeanTables &LT = ctx->lean;
    const int f32 = std::is_same<real, float>::value ? 1 : 0;
#define FNET_LEAN_LAUNCH2(NL, NC, PATH, SORTED, G, F32A)                                                        \
  do {                                                                                                         \
    CUDA_TRY(ctx, fnet_smem_attr(k_acsf_lean<NL, NC, PATH, SORTED, G, F32A>, L.smem)); \
    LAUNCH(ctx, K_ACSF, (fnet_launch_k(ctx->pdl, k_acsf_lean<NL, NC, PATH, SORTED, G, F32A>, L.grid, dim3(L.wpb * 32), L.smem, ctx->stream, \
                            L.nSplit, geo, s.nExt, s.d_ext, T, LT, L.cap, L.capC, (void *)feat, f32, nFeat, zp, \
                            nExtSel, ctx->d_extIdx, ctx->d_flags)));                                           \
  } while (0)
#define FNET_LEAN_LAUNCH(NL, NC, PATH, SORTED, G)                                                              \
  do { if (f32) FNET_LEAN_LAUNCH2(NL, NC, PATH, SORTED, G, true); else FNET_LEAN_LAUNCH2(NL, NC, PATH, SORTED, G, false); } while (0)
#define FNET_LEAN_LAUNCH_S(NL, NC, PATH)                                                                       \
  do {                                                                                                         \
    if (ctx->leanSorted) { if (L.G == 2) FNET_LEAN_LAUNCH(NL, NC, PATH, true, 2); else FNET_LEAN_LAUNCH(NL, NC, PATH, true, 1); } \
    else if (L.G == 4) FNET_LEAN_LAUNCH(NL, NC, PATH, false, 4);                                               \
    else if (L.G == 2) FNET_LEAN_LAUNCH(NL, NC, PATH, false, 2);                                               \
    else FNET_LEAN_LAUNCH(NL, NC, PATH, false, 1);                                                             \
  } while (0)
#define FNET_LEAN_LAUNCH_P(NL, NC)                                                                             \
  do {                                                                                                         \
    if (L.path == FNET_PATH_STRUCT) FNET_LEAN_LAUNCH_S(NL, NC, FNET_PATH_STRUCT);                              \
    else if (L.path == FNET_PATH_STAGED) FNET_LEAN_LAUNCH_S(NL, NC, FNET_PATH_STAGED);                         \
    else if (ctx->leanSorted) FNET_LEAN_LAUNCH(NL, NC, FNET_PATH_DIRECT, true, 1);                             \
    else FNET_LEAN_LAUNCH(NL, NC, FNET_PATH_DIRECT, false, 1);                                                 \
  } while (0)
    if (ctx->leanNC == 1) FNET_LEAN_LAUNCH_P(2, 1);
    else if (ctx->leanNC == 2) FNET_LEAN_LAUNCH_P(2, 2);
    else FNET_LEAN_LAUNCH_P(1, 4);
#undef FNET_LEAN_LAUNCH_P
#undef FNET_LEAN_LAUNCH_S
#undef FNET_LEAN_LAUNCH
#undef FNET_LEAN_LAUNCH2
    return 0;
  }
  const int ns = ctx->maxSlots <= 1 ? 1 : (ctx->maxSlots <= 2 ? 2 : 4);
  if (L.path == FNET_PATH_STRUCT) FNET_ACSF_LAUNCH_NS(FNET_PATH_STRUCT);
  else if (L.path == FNET_PATH_STAGED) FNET_ACSF_LAUNCH_NS(FNET_PATH_STAGED);
  else FNET_ACSF_LAUNCH_NS(FNET_PATH_DIRECT);
#undef FNET_ACSF_LAUNCH_NS
#undef FNET_ACSF_LAUNCH
  return 0;
}


template <typename real>
static int acsf_calculate_t(fnetgpu_ctx *ctx, Slot &s, int standardize, double *zprec, int have_zprec,
                            const double *h_coords = nullptr) {
  // h_coords: new coordinates still on the host (fnetgpu_acsf_update_calculate) -- uploaded in
  // chunks of structures on a second stream while the kernel already works on the earlier chunks
  const AcsfTables &T = ctx->acsf;
  const int F = T.F, nExtSel = (int)ctx->extIdx.size();
  const int nFeat = F + nExtSel;
  if (nFeat == 0) FNET_FAIL(ctx, "acsf_calculate: no features configured");
  if (nExtSel > 0) {
    for (int e : ctx->extIdx) if (e >= s.nExt) FNET_FAIL(ctx, "external feature index exceeds the dataset's extfeatures");
  }
  if (s.nFeat != nFeat || !s.d_feat) {
    real *p = nullptr;
    if (dev_alloc(ctx, &p, (size_t)s.N * nFeat)) return 1;
    cudaFree(s.d_feat);
    s.d_feat = p; s.nFeat = nFeat;
  }
  real *feat = (real *)s.d_feat;
  const bool useGiven = standardize && have_zprec;
  if (useGiven && have_zprec == 2) {            // internal: the statistics already on the device (fnetgpu_socket_step)
    if (!ctx->haveZ) FNET_FAIL(ctx, "acsf_calculate: no z-score statistics on the device");
  } else if (useGiven) {
    if (!zprec) FNET_FAIL(ctx, "acsf_calculate: zprec missing");
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_zprec, zprec, (size_t)2 * F * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->haveZ = true;
  }
  bool verifiedRepeat = false;
  if (F > 0) {
    // buffer capacities (neighbours per atom, candidates per bin): counted once per slot; after a
    // geometry update the previous maxima are reused as hints and the kernel's overflow flags
    // trigger a retry.  Small structures take the whole-structure path (no cell list); its lattice
    // check (flags[4]) sends the slot back to the cell list.
    for (int attempt = 0; attempt < 4; attempt++) {
      AcsfLaunch L;
      const bool sp = use_struct_path(ctx, s);
      if (sp && plan_values(ctx, s, true, L)) {
        // the whole-structure buffers do not fit (very many neighbours per atom): the cell list takes the slot
        ctx->err.clear();
        s.structPath = 0; s.maxNeigh = -1;
        continue;
      }
      if (!sp) {
        if (h_coords) {                         // the cell list is built from the device copy: plain upload first
          CUDA_TRY(ctx, cudaMemcpyAsync(s.d_coords, h_coords, (size_t)3 * s.N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
          h_coords = nullptr;
        }
        if (ensure_cells(ctx, s, T.rcMax, nullptr)) return 1;
        if (s.maxNeigh < 0 || s.maxCand < 0) { if (ensure_neigh_count(ctx, s)) return 1; }
        if (plan_values(ctx, s, false, L)) return 1;
      }
      if (!h_coords && s.okEpoch == s.geomEpoch && L.cap == s.okCap && L.capC == s.okCapC && L.path == s.okPath &&
          L.G == s.okG && (int)L.lean == s.okLean) {
        // same plan on the geometry of a launch whose flags were clean: nothing to read back, nothing to wait for
        if (launch_acsf_values<real>(ctx, s, L, useGiven ? ctx->d_zprec : nullptr)) return 1;
        s.lastPath = L.path;
        verifiedRepeat = true;
        break;
      }
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), ctx->stream));
      if (h_coords && sp) {
        // whole-structure path: structures are independent launches -> pipeline copy and kernel
        if (!ctx->copyStream) {
          CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
          for (int c = 0; c < 8; c++) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evChunk[c], cudaEventDisableTiming));
        }
        // (measured alternative: up to 16 equal wave-aligned chunks -- every launch boundary drains the SMs: C2 +0.06 ms,
        // C3 +0.7 ms end to end on a host whose copy is faster than the kernel; only a copy-bound host would gain)
        // chunks of 1, 2, 4 full WAVES of CTAs and the rest: the first kernel starts after one wave's worth of the copy,
        // the copy of chunk c + 1 (twice the bytes, ~2.5x the kernel's rate per byte) hides behind the kernel of chunk c,
        // and only the last launch has a partially filled wave (chunks of nStruct / 15 ... 8 nStruct / 15 structures cost
        // three extra waves out of 17 on C2 once a CTA takes a whole 64-atom structure)
        const int ctasPerSM = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (L.smem + 1024)));
        const long long wave = std::max<long long>(1, (long long)ctx->nSM * ctasPerSM / std::max(1, L.nSplit));
        int bound[9], nChunk = 0;
        bound[0] = 0;
        {
          long long done = 0, waves = 1;
          while (nChunk < 3 && (long long)s.nStruct - done >= 2 * wave * waves) { done += wave * waves; bound[++nChunk] = (int)done; waves *= 2; }
          bound[++nChunk] = s.nStruct;
        }
        for (int c = 0; c < nChunk; c++) {
          const int st0 = bound[c], st1 = bound[c + 1];
          const size_t a0 = s.h_offsets[st0], a1 = s.h_offsets[st1];
          CUDA_TRY(ctx, cudaMemcpyAsync(s.d_coords + 3 * a0, h_coords + 3 * a0, 3 * (a1 - a0) * sizeof(double), cudaMemcpyHostToDevice, ctx->copyStream));
          CUDA_TRY(ctx, cudaEventRecord(ctx->evChunk[c], ctx->copyStream));
        }
        for (int c = 0; c < nChunk; c++) {
          const int st0 = bound[c], st1 = bound[c + 1];
          CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evChunk[c], 0));
          AcsfLaunch Lc = L;
          Lc.stBase = st0; Lc.grid = dim3(st1 - st0, L.nSplit);
          if (launch_acsf_values<real>(ctx, s, Lc, useGiven ? ctx->d_zprec : nullptr)) return 1;
        }
        h_coords = nullptr;                     // on the device now: retries relaunch over the whole slot
      } else {
        if (launch_acsf_values<real>(ctx, s, L, useGiven ? ctx->d_zprec : nullptr)) return 1;
      }
      int h[16];
      CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_flags, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      s.lastPath = L.path;
      if (L.lean && sp && h[4] == 0 && h[1] == 0 && h[7] == 0 && h[0] > 0) s.maxNeigh = h[0];   // exact maximum of this geometry
      if (h[4] == 0 && h[1] == 0 && h[7] == 0) {
        s.okEpoch = s.geomEpoch; s.okCap = L.cap; s.okCapC = L.capC; s.okPath = L.path; s.okG = L.G; s.okLean = (int)L.lean;
      }
      if (sp) {
        if (h[4] != 0) { s.structPath = 0; s.maxNeigh = -1; continue; }   // a lattice is too small for the minimum image: cell list
        if (h[1] == 0 && h[7] == 0) break;
        if (h[1] != 0) s.maxNeigh = h[1];
        if (h[7] != 0) FNET_FAIL(ctx, "structure larger than the staged-structure buffer");
      } else {
        s.maxBinPop = h[8 + 4];
        if (h[1] == 0 && h[7] == 0) break;
        if (h[1] != 0) s.maxNeigh = h[1];
        if (h[7] != 0) s.maxCand = h[7];     // 0x7fffffff: too many neighbour cells -> direct path
      }
      if (attempt == 3) FNET_FAIL(ctx, "neighbour buffer overflow");
    }
  } else {
    const int B = 256;
    const long long tot = (long long)s.N * nExtSel;
    LAUNCH(ctx, K_EXT_CONCAT, (k_ext_concat<real><<<(int)((tot + B - 1) / B), B, 0, ctx->stream>>>(s.N, s.nExt, s.d_ext, 0, nExtSel, ctx->d_extIdx, feat, nFeat)));
  }
  if (standardize && !have_zprec && F > 0) {
    // acsf.F90:445-486 two-pass statistics with the dataset weights
    const int apb = 256;
    const int nb = (s.N + apb - 1) / apb;
    size_t need = (size_t)nb * F + 2 * F + 8;
    if (ctx->partialsN < need) { if (dev_alloc(ctx, &ctx->d_partials, need)) return 1; ctx->partialsN = need; }
    double *part = ctx->d_partials, *sums = ctx->d_partials + (size_t)nb * F;   // sums[F] (+1 count)
    if (ensure_pinned(ctx, (size_t)2 * F + 8)) return 1;
    double wN = 0.0;                             // sum_s w_s N_s (acsf.F90:462-470)
    {
      std::vector<double> dsw(s.nStruct);
      CUDA_TRY(ctx, cudaMemcpyAsync(dsw.data(), s.d_dsw, s.nStruct * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      wN = 0.0;
      for (int st = 0; st < s.nStruct; st++) wN += dsw[st] * (double)(s.h_offsets[st + 1] - s.h_offsets[st]);
    }
    double *h = ctx->h_pinned;
    for (int pass = 0; pass < 2; pass++) {
      LAUNCH(ctx, K_ZSTAT, (k_zstat<real><<<nb, 128, 0, ctx->stream>>>(s.N, F, nFeat, feat, s.d_structOf, s.d_dsw, pass ? ctx->d_zprec : nullptr, apb, part)));
      LAUNCH(ctx, K_ZSTAT_FINAL, (k_zstat_final<<<F, 256, 0, ctx->stream>>>(nb, F, part, sums)));
      if (pass == 0) {
        // append the weighted atom count so one all-reduce carries both
        CUDA_TRY(ctx, cudaMemcpyAsync(sums + F, &wN, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (allreduce_sum(ctx, sums, (size_t)F + 1)) return 1;
        CUDA_TRY(ctx, cudaMemcpyAsync(h, sums, ((size_t)F + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        wN = h[F];
        for (int a = 0; a < F; a++) h[a] = h[a] / wN;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_zprec, h, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (zprec) memcpy(zprec, h, (size_t)F * sizeof(double));
      } else {
        if (allreduce_sum(ctx, sums, (size_t)F)) return 1;
        CUDA_TRY(ctx, cudaMemcpyAsync(h, sums, (size_t)F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int a = 0; a < F; a++) h[a] = sqrt(h[a] / wN);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_zprec + F, h, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (zprec) memcpy(zprec + F, h, (size_t)F * sizeof(double));
      }
    }
    ctx->haveZ = true;
    const long long tot = (long long)s.N * F;
    LAUNCH(ctx, K_ZAPPLY, (k_zapply<real><<<(int)((tot + 255) / 256), 256, 0, ctx->stream>>>((size_t)s.N, F, nFeat, feat, ctx->d_zprec)));
  }
  if (!standardize) ctx->haveZ = false;
  if (!verifiedRepeat) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // (a verified repeat launch stays asynchronous: everything that reads the features is ordered on the same stream)
  s.featValid = true;
  s.zscored = standardize && F > 0;
  return 0;
}

extern "C" int fnetgpu_acsf_calculate(fnetgpu_ctx *ctx, int slot, int standardize, double *zprec, int have_zprec) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used) FNET_FAIL(ctx, "acsf_calculate: empty slot");
  if (!ctx->acsfSet && ctx->extIdx.empty()) FNET_FAIL(ctx, "acsf_calculate: call fnetgpu_acsf_set / fnetgpu_features_config first");
  cudaSetDevice(ctx->device);
  have_zprec = have_zprec ? 1 : 0;
  if (ctx->precision == 64) return acsf_calculate_t<double>(ctx, s, standardize, zprec, have_zprec);
  return acsf_calculate_t<float>(ctx, s, standardize, zprec, have_zprec);
}

// fnetgpu_coords_update + fnetgpu_acsf_calculate in one blocking call; for small structures the
// upload is pipelined with the kernel (chunks of structures, second stream)
extern "C" int fnetgpu_acsf_update_calculate(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs,
                                             int standardize, double *zprec, int have_zprec) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used) FNET_FAIL(ctx, "acsf_update_calculate: empty slot");
  if (!coords) FNET_FAIL(ctx, "acsf_update_calculate: coords missing");
  if (!ctx->acsfSet || ctx->acsf.F == 0) FNET_FAIL(ctx, "acsf_update_calculate: call fnetgpu_acsf_set first");
  cudaSetDevice(ctx->device);
  if (dev_reserve(ctx, &s.d_coords, &s.capCoords, (size_t)3 * s.N)) return 1;
  if (latvecs) {
    CUDA_TRY(ctx, cudaMemcpyAsync(s.d_lat, latvecs, (size_t)9 * s.nStruct * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (s.h_lat.size() != (size_t)9 * s.nStruct || memcmp(s.h_lat.data(), latvecs, (size_t)9 * s.nStruct * sizeof(double)) != 0) {
      s.h_lat.assign(latvecs, latvecs + (size_t)9 * s.nStruct);
      s.structPath = 1;
    }
  }
  s.cellRc = -1.0; s.neighStale = true; s.featValid = false; s.geomEpoch++;
  have_zprec = have_zprec ? 1 : 0;
  if (ctx->precision == 64) return acsf_calculate_t<double>(ctx, s, standardize, zprec, have_zprec, coords);
  return acsf_calculate_t<float>(ctx, s, standardize, zprec, have_zprec, coords);
}

template <typename real> __global__ void k_convert_out(size_t n, const real *__restrict__ in, double *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (double)in[t];
}
template <typename real> __global__ void k_convert_in(size_t n, const double *__restrict__ in, real *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (real)in[t];
}

// device real buffer -> host double buffer
static int download_real(fnetgpu_ctx *ctx, const void *d_src, size_t n, double *h_dst) {
  if (ctx->precision == 64) {
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    if (dev_reserve(ctx, &ctx->d_conv, &ctx->convN, n)) return 1;     // grow-only conversion buffer of the context
    double *tmp = ctx->d_conv;
    LAUNCH(ctx, K_MISC, (k_convert_out<float><<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(n, (const float *)d_src, tmp)));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int fnetgpu_features_get(fnetgpu_ctx *ctx, int slot, double *out) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used || !s.featValid) FNET_FAIL(ctx, "features_get: features not computed");
  cudaSetDevice(ctx->device);
  return download_real(ctx, s.d_feat, (size_t)s.N * s.nFeat, out);
}

extern "C" int fnetgpu_features_set(fnetgpu_ctx *ctx, int slot, int nFeat, const double *in) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used) FNET_FAIL(ctx, "features_set: empty slot");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)s.N * nFeat;
  cudaFree(s.d_feat); s.d_feat = nullptr;
  if (ctx->precision == 64) {
    double *p = nullptr;
    if (dev_upload(ctx, &p, in, n)) return 1;
    s.d_feat = p;
  } else {
    double *tmp = nullptr; float *p = nullptr;
    if (dev_upload(ctx, &tmp, in, n)) return 1;
    if (dev_alloc(ctx, &p, n)) return 1;
    LAUNCH(ctx, K_MISC, (k_convert_in<float><<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(n, tmp, p)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    s.d_feat = p;
  }
  s.nFeat = nFeat; s.featValid = true;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// network
// ------------------------------------------------------------------------------------------
extern "C" int fnetgpu_net_set(fnetgpu_ctx *ctx, int nSpecies, int nLayers, const int *dims, int activationId) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  if (nSpecies < 1) FNET_FAIL(ctx, "net_set: nSpecies < 1");
  if (nLayers < 2 || nLayers > FNET_MAX_LAYERS) FNET_FAIL(ctx, "net_set: unsupported number of layers");
  if (activationId < 0 || activationId > FNETGPU_ACT_LINEAR) FNET_FAIL(ctx, "net_set: unknown activation (no fallback)");
  NetTables &n = ctx->net;
  memset(&n, 0, sizeof(n));
  n.nSpecies = nSpecies; n.L = nLayers; n.act = activationId; n.nOut = dims[nLayers - 1];
  int ind = 0, rows = 0;
  for (int l = 0; l < nLayers; l++) {
    if (dims[l] < 1) FNET_FAIL(ctx, "net_set: layer width < 1");
    n.dims[l] = dims[l];
    n.aoff[l] = rows; rows += dims[l];
  }
  n.rowsA = rows;
  for (int l = 0; l < nLayers; l++) { n.woff[l] = ind; ind += dims[l] * (l + 1 < nLayers ? dims[l + 1] : 1); }   // network.F90:413-419
  for (int l = 0; l < nLayers; l++) { n.boff[l] = ind; ind += dims[l]; }                                        // network.F90:422-425
  n.nTot = ind;
  cudaFree(ctx->d_wb); ctx->d_wb = nullptr;
  size_t bytes = (size_t)n.nTot * nSpecies * (ctx->precision == 64 ? 8 : 4);
  CUDA_TRY(ctx, cudaMalloc(&ctx->d_wb, bytes));
  if (dev_alloc(ctx, &ctx->d_wb64, (size_t)n.nTot * nSpecies)) return 1;
  if (dev_alloc(ctx, &ctx->d_dd, (size_t)n.nTot * nSpecies + 8)) return 1;
  ctx->netSet = true; ctx->paramsSet = false; ctx->netEpoch++;
  for (int i = 0; i < FNETGPU_MAX_SLOTS; i++) { ctx->slots[i].nTiles = 0; ctx->slots[i].nTiles16 = 0; ctx->slots[i].nTilesS = 0; ctx->slots[i].nTilesC = -1; ctx->slots[i].nTilesW = 0; }
  return 0;
}

extern "C" int fnetgpu_ntot(const fnetgpu_ctx *ctx) { return (ctx && ctx->netSet) ? ctx->net.nTot : -1; }

extern "C" int fnetgpu_params_set(fnetgpu_ctx *ctx, const double *wb) {
  CHECK_CTX(ctx);
  if (!ctx->netSet) FNET_FAIL(ctx, "params_set: call fnetgpu_net_set first");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->net.nTot * ctx->net.nSpecies;
  if (ctx->precision == 64) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_wb, wb, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_wb64, wb, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, K_MISC, (k_convert_in<float><<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->d_wb64, (float *)ctx->d_wb)));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->paramsSet = true;
  return 0;
}

// tile table of the species-sorted atom order; T = atoms per tile (k_bpnn, mlp.cuh)
template <typename real>
static int ensure_tiles(fnetgpu_ctx *ctx, Slot &s) {
  if (s.nTiles > 0) return 0;
  const NetTables &n = ctx->net;
  if ((int)s.spBeg.size() - 1 > n.nSpecies) FNET_FAIL(ctx, "dataset references more species than the network has sub-networks");
  int T = 0;
  const int cand[3] = {64, 32, 16};
  for (int c = 0; c < 3 && !T; c++)
    if (bpnn_smem_bytes<real>(n, cand[c], 0, false) <= 220 * 1024) T = cand[c];
  if (!T) FNET_FAIL(ctx, "network too large for the shared-memory tile (layer widths sum too big)");
  std::vector<int> tiles;
  for (int sp = 0; sp + 1 < (int)s.spBeg.size(); sp++)
    for (int b = s.spBeg[sp]; b < s.spBeg[sp + 1]; b += T) {
      tiles.push_back(b); tiles.push_back(std::min(T, s.spBeg[sp + 1] - b)); tiles.push_back(sp);
    }
  s.nTiles = (int)tiles.size() / 3; s.tileT = T;
  if (dev_upload(ctx, &s.d_tiles, tiles.data(), tiles.size())) return 1;
  {   // rounds of the FP64 tensor-core path: 4 warps x 16 atoms of one species
    std::vector<int> t16;
    const int R = FNET_MMA_TA * FNET_MMA_WARPS;
    for (int sp = 0; sp + 1 < (int)s.spBeg.size(); sp++)
      for (int b = s.spBeg[sp]; b < s.spBeg[sp + 1]; b += R) {
        t16.push_back(b); t16.push_back(std::min(R, s.spBeg[sp + 1] - b)); t16.push_back(sp);
      }
    s.nTiles16 = (int)t16.size() / 3;
    if (dev_upload(ctx, &s.d_tiles16, t16.data(), t16.size())) return 1;
    // structure-aligned rounds: single-species data (the species-sorted order is the atom order) with
    // <= 64 atoms per structure -> whole structures packed greedily into rounds of <= 64 atoms
    s.nTilesS = 0;
    if ((int)s.spBeg.size() == 2 && s.maxAtoms <= R) {
      std::vector<int> ts;
      int st = 0;
      while (st < s.nStruct) {
        const int b = s.h_offsets[st];
        int e = st;
        while (e < s.nStruct && s.h_offsets[e + 1] - b <= R) e++;
        ts.push_back(b); ts.push_back(s.h_offsets[e] - b); ts.push_back(0);
        st = e;
      }
      s.nTilesS = (int)ts.size() / 3;
      if (dev_upload(ctx, &s.d_tilesS, ts.data(), ts.size())) return 1;
    }
  }
  real *raw = nullptr;
  if (dev_alloc(ctx, &raw, (size_t)s.N * n.nOut)) return 1;
  cudaFree(s.d_raw); s.d_raw = raw;
  if (dev_alloc(ctx, &s.d_gS, (size_t)s.nStruct * std::max(s.nG, 1))) return 1;
  if (dev_alloc(ctx, &s.d_Es, (size_t)s.nStruct * std::max(s.nG, 1))) return 1;
  if (dev_alloc(ctx, &s.d_lossPart, (size_t)2 * s.nStruct)) return 1;
  return 0;
}

// launch geometry of k_bpnn: threads = largest per-layer item count (4-atom x 4-output register
// tiles), persistent grid sized by the shared-memory footprint
struct BpnnLaunch { int threads, grid, gInSmem; size_t smem; };
template <typename real>
static BpnnLaunch plan_bpnn(fnetgpu_ctx *ctx, const Slot &s, int mode) {
  const NetTables &n = ctx->net;
  BpnnLaunch B;
  const int RG = s.tileT / 4;
  int items = 32;
  for (int l = 0; l < n.L; l++) items = std::max(items, RG * ((n.dims[l] + 3) / 4));
  B.threads = std::min(256, (items + 31) & ~31);
  B.gInSmem = (mode == 0 && bpnn_smem_bytes<real>(n, s.tileT, 0, true) <= 220 * 1024) ? 1 : 0;
  B.smem = bpnn_smem_bytes<real>(n, s.tileT, mode, B.gInSmem != 0);
  const int perSM = std::max(1, std::min((int)(225 * 1024 / (B.smem + 1024)), 2048 / B.threads));
  B.grid = std::max(1, std::min(s.nTiles, ctx->nSM * std::min(perSM, 8)));
  return B;
}

template <typename real>
static int check_ready(fnetgpu_ctx *ctx, Slot &s, bool needTargets) {
  if (!s.used) FNET_FAIL(ctx, "empty dataset slot");
  if (!ctx->netSet || !ctx->paramsSet) FNET_FAIL(ctx, "network / parameters not set");
  if (!s.featValid) FNET_FAIL(ctx, "features of this slot have not been computed");
  if (s.nFeat != ctx->net.dims[0]) FNET_FAIL(ctx, "feature count does not match the input layer width");
  if (needTargets && s.nG + s.nA != ctx->net.nOut) FNET_FAIL(ctx, "number of targets does not match the output layer width");
  return ensure_tiles<real>(ctx, s);
}

// precision 64: DMMA kernels (mlp_mma.cuh) whenever the network fits their limits
template <typename real>
static bool use_mma(const fnetgpu_ctx *ctx) {
  return std::is_same<real, double>::value && !ctx->mlpLegacy && bpnn_mma_fits(ctx->net);
}
struct MmaLaunch { int grid; size_t smem; };
static MmaLaunch plan_mma(const fnetgpu_ctx *ctx, const Slot &s, int mode) {
  MmaLaunch M;
  M.smem = bpnn_mma_smem_bytes(ctx->net, mode);
  const int perSM = std::max(1, std::min((int)(225 * 1024 / (M.smem + 1024)), 4));
  M.grid = std::max(1, std::min(s.nTiles16, ctx->nSM * perSM));
  return M;
}
// persistent grid = what is actually resident (registers can bind before shared memory does)
template <typename K>
static int mma_grid(const fnetgpu_ctx *ctx, const Slot &s, K kernel, size_t smem, int fallback) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, FNET_MMA_WARPS * 32, smem) != cudaSuccess || nb < 1) return fallback;
  return std::max(1, std::min(s.nTiles16, ctx->nSM * nb));
}

template <typename real>
static int run_forward(fnetgpu_ctx *ctx, Slot &s) {
  const NetTables &n = ctx->net;
  if constexpr (std::is_same<real, double>::value) {
    if (use_mma<real>(ctx)) {
      const MmaLaunch M = plan_mma(ctx, s, 2);
#define FNET_MMA_FWD(FCH)                                                                                         \
      do {                                                                                                        \
        CUDA_TRY(ctx, fnet_smem_attr(k_bpnn_mma<2, 1, FCH>, M.smem)); \
        const int fgrid = mma_grid(ctx, s, k_bpnn_mma<2, 1, FCH>, M.smem, M.grid);                                \
        LAUNCH(ctx, K_MLP_FWD, (k_bpnn_mma<2, 1, FCH><<<fgrid, FNET_MMA_WARPS * 32, M.smem, ctx->stream>>>(       \
                                   s.nTiles16, s.d_tiles16, s.d_perm, (const double *)s.d_feat, s.nFeat,          \
                                   (const double *)ctx->d_wb, n, nullptr, nullptr, nullptr, nullptr, nullptr,     \
                                   nullptr, s.nG, s.nA, 0, nullptr, (double *)s.d_raw, nullptr, nullptr, nullptr, nullptr))); \
      } while (0)
      if (n.dims[0] <= 32) FNET_MMA_FWD(1); else FNET_MMA_FWD(2);
#undef FNET_MMA_FWD
      return 0;
    }
  }
  const BpnnLaunch B = plan_bpnn<real>(ctx, s, 2);
  CUDA_TRY(ctx, fnet_smem_attr(k_bpnn<real, 2>, B.smem));
  LAUNCH(ctx, K_MLP_FWD, (k_bpnn<real, 2><<<B.grid, B.threads, B.smem, ctx->stream>>>(
                             s.nTiles, s.d_tiles, s.d_perm, (const real *)s.d_feat, s.nFeat, (const real *)ctx->d_wb, n,
                             s.tileT, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s.nG, s.nA, 0, nullptr,
                             (real *)nullptr, (real *)s.d_raw)));
  return 0;
}

// dE_k/dG for every atom and output (nJacobian, bpnn.F90:904-997): one reverse sweep per output; precision 64: DMMA kernel
template <typename real>
static int run_ingrad(fnetgpu_ctx *ctx, Slot &s) {
  const NetTables &n = ctx->net;
  if constexpr (std::is_same<real, double>::value) {
    if (use_mma<real>(ctx)) {
      const MmaLaunch M = plan_mma(ctx, s, 1);
#define FNET_MMA_ING(FCH)                                                                                         \
      do {                                                                                                        \
        CUDA_TRY(ctx, fnet_smem_attr(k_bpnn_mma<1, 1, FCH>, M.smem)); \
        const int fgrid = mma_grid(ctx, s, k_bpnn_mma<1, 1, FCH>, M.smem, M.grid);                                \
        LAUNCH(ctx, K_MLP_INGRAD, (k_bpnn_mma<1, 1, FCH><<<fgrid, FNET_MMA_WARPS * 32, M.smem, ctx->stream>>>(    \
                                      s.nTiles16, s.d_tiles16, s.d_perm, (const double *)s.d_feat, s.nFeat,       \
                                      (const double *)ctx->d_wb, n, nullptr, nullptr, nullptr, nullptr, nullptr,  \
                                      nullptr, s.nG, s.nA, 0, nullptr, (double *)s.d_dEdG, nullptr, nullptr, nullptr, nullptr))); \
      } while (0)
      if (n.dims[0] <= 32) FNET_MMA_ING(1); else FNET_MMA_ING(2);
#undef FNET_MMA_ING
      return 0;
    }
  }
  const BpnnLaunch B = plan_bpnn<real>(ctx, s, 1);
  CUDA_TRY(ctx, fnet_smem_attr(k_bpnn<real, 1>, B.smem));
  LAUNCH(ctx, K_MLP_INGRAD, (k_bpnn<real, 1><<<B.grid, B.threads, B.smem, ctx->stream>>>(
                                s.nTiles, s.d_tiles, s.d_perm, (const real *)s.d_feat, s.nFeat, (const real *)ctx->d_wb, n,
                                s.tileT, 0, s.d_structOf, s.d_offsets, nullptr, nullptr, nullptr, nullptr, s.nG, s.nA, 0,
                                nullptr, (real *)s.d_dEdG, (real *)nullptr)));
  return 0;
}

template <typename real>
static int run_struct_loss(fnetgpu_ctx *ctx, Slot &s, int lossId) {
  const NetTables &n = ctx->net;
  const int B = 128;
  const int grid = (int)(((long long)s.nStruct * 32 + B - 1) / B);
  LAUNCH(ctx, K_STRUCT_LOSS, (k_struct_loss<real><<<grid, B, 0, ctx->stream>>>(s.nStruct, s.d_offsets, n.nOut, s.nG, s.nA, (const real *)s.d_raw, s.d_gt, s.d_at, s.d_aw, s.d_dsw, lossId, s.d_Es, s.d_gS, s.d_lossPart)));
  const size_t nDD = (size_t)n.nTot * n.nSpecies;
  LAUNCH(ctx, K_LOSS_FINAL, (k_loss_final<<<1, 1024, 0, ctx->stream>>>(s.nStruct, s.d_lossPart, ctx->d_dd + nDD)));
  return 0;
}

// TWeightDerivs_elasticNetRegularization (lib_common/nestedtypes.F90:336-370) + the division by the number of
// datapoints of TBpnn_update (lib_nn/bpnn.F90:750-767) on the reduced gradient:
//   dd(w) += lambda / nWeights * ((1 - alpha) w + alpha sign(w))  for the nW weight entries of every species, then dd *= inv
template <typename real>
__global__ void k_regularize(int nSpecies, int nTot, int nW, const real *__restrict__ wb, double *__restrict__ dd, double c,
                             double alpha, double inv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nSpecies * nTot) return;
  double g = dd[e];
  if (e % nTot < nW) {
    const double w = (double)wb[e];
    const double sg = w > 0.0 ? 1.0 : (w < 0.0 ? -1.0 : 0.0);
    g += c * ((1.0 - alpha) * w + alpha * sg);
  }
  dd[e] = g * inv;
}

// strength = 0 and nDatapoints <= 0 switch both off (the default: fnetgpu_grad returns the plain summed gradient)
extern "C" int fnetgpu_regularization_set(fnetgpu_ctx *ctx, double strength, double alpha, double nDatapoints) {
  CHECK_CTX(ctx);
  if (strength < 0.0 || alpha < 0.0 || alpha > 1.0) FNET_FAIL(ctx, "regularization_set: need strength >= 0 and 0 <= alpha <= 1");
  ctx->reguLambda = strength; ctx->reguAlpha = alpha; ctx->reguDiv = nDatapoints > 0.0 ? nDatapoints : 0.0;
  return 0;
}
// reguLoss of every species (lib_common/loss.F90:119-196: lambda / nW ((1 - alpha) / 2 sum w^2 + alpha sum |w|)), from
// the parameters on the device; out[nSpecies]
extern "C" int fnetgpu_regularization_loss(fnetgpu_ctx *ctx, double *out) {
  CHECK_CTX(ctx);
  if (!ctx->netSet || !ctx->paramsSet) FNET_FAIL(ctx, "regularization_loss: network / parameters not set");
  cudaSetDevice(ctx->device);
  const NetTables &n = ctx->net;
  const int nW = n.boff[0];
  const size_t tot = (size_t)n.nTot * n.nSpecies;
  std::vector<double> w(tot);
  if (ctx->precision == 64) {
    CUDA_TRY(ctx, cudaMemcpyAsync(w.data(), ctx->d_wb, tot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    CUDA_TRY(ctx, cudaMemcpyAsync(w.data(), ctx->d_wb64, tot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int sp = 0; sp < n.nSpecies; sp++) {      // O(nW) numbers: summed on the host in the reference's order
    double s2 = 0.0, s1 = 0.0;
    for (int e = 0; e < nW; e++) { const double v = w[(size_t)n.nTot * sp + e]; s2 += v * v; s1 += fabs(v); }
    out[sp] = ctx->reguLambda / (double)nW * ((1.0 - ctx->reguAlpha) / 2.0 * s2 + ctx->reguAlpha * s1);
  }
  return 0;
}

// the main stream must not touch d_dd before a pending all-reduce (side stream) is done with it
static int wait_allreduce(fnetgpu_ctx *ctx) {
  if (!ctx->arPending) return 0;
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evAR, 0));
  ctx->arPending = false;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Cluster-fused per-structure sums (k_bpnn_mma<0, .., 2>, mlp_mma.cuh): multi-species data and structures of more than
// 64 atoms get E_s / the loss gradient inside the gradient kernel too.  Consecutive structures are packed into groups
// whose species-homogeneous rounds (<= 64 atoms) number <= CS; the CS CTAs of a thread-block cluster take one round
// each and exchange their partial sums through distributed shared memory.  CS is chosen per dataset from the
// co-resident clusters the device reports (cudaOccupancyMaxActiveClusters: GPC sizes strand SMs for some CS) and
// the padding of the groups; when the two-pass path (forward kernel + k_struct_loss + gradient kernel) is estimated
// cheaper -- or a structure needs more than 8 rounds -- it stays.
// ------------------------------------------------------------------------------------------
static MmaKernelT mma_cluster_kernel(const NetTables &n) {
  const MmaLayout ml = mma_layout(n);
  const int perWarp = (ml.nGradTiles + FNET_MMA_WARPS - 1) / FNET_MMA_WARPS;
  const bool f1 = n.dims[0] <= 32;
  if (perWarp <= 4) return f1 ? (MmaKernelT)k_bpnn_mma<0, 4, 1, 2> : (MmaKernelT)k_bpnn_mma<0, 4, 2, 2>;
  return f1 ? (MmaKernelT)k_bpnn_mma<0, FNET_MMA_MAXSLOTS, 1, 2> : (MmaKernelT)k_bpnn_mma<0, FNET_MMA_MAXSLOTS, 2, 2>;
}
struct ClusterPlan { int CS = 0, nSuper = 0; double cost = 0.0; std::vector<int> tiles, perm, seg; };
// spCount: [nStruct][nSp] atoms per structure and species; false: some structure needs more than CS rounds
static bool build_cluster_plan(const Slot &s, const std::vector<int> &spCount, int CS, bool fill, ClusterPlan &P) {
  const int R = FNET_MMA_TA * FNET_MMA_WARPS;
  const int nSp = (int)s.spBeg.size() - 1;
  P.CS = CS; P.nSuper = 0; P.cost = 0.0; P.tiles.clear();
  if (fill) { P.perm.assign(s.N, 0); P.seg.assign(s.N, 0); }
  std::vector<int> cnt(nSp), lastSp(CS, 0), run, runSt;
  int st = 0, pos = 0;
  while (st < s.nStruct) {
    std::fill(cnt.begin(), cnt.end(), 0);
    int e = st;
    while (e < s.nStruct && e - st < FNET_MMA_CSLOTS) {
      int rounds = 0;
      for (int sp = 0; sp < nSp; sp++) rounds += (cnt[sp] + spCount[(size_t)e * nSp + sp] + R - 1) / R;
      if (rounds > CS) break;
      for (int sp = 0; sp < nSp; sp++) cnt[sp] += spCount[(size_t)e * nSp + sp];
      e++;
    }
    if (e == st) return false;
    int rank = 0, maxTiles = 0;
    for (int sp = 0; sp < nSp; sp++) {
      if (cnt[sp] == 0) continue;
      const int nr = (cnt[sp] + R - 1) / R;
      const int per = std::min(R, (((cnt[sp] + nr - 1) / nr) + 7) & ~7);
      if (fill) {
        run.clear(); runSt.clear();
        for (int t = st; t < e; t++)
          for (int i = s.h_offsets[t]; i < s.h_offsets[t + 1]; i++)
            if (s.h_globalsp[i] == sp) { run.push_back(i); runSt.push_back(t); }
      }
      for (int q = 0; q < nr; q++) {
        const int b = q * per, c = std::min(per, cnt[sp] - b);
        maxTiles = std::max(maxTiles, (c + FNET_MMA_TW - 1) / FNET_MMA_TW);
        P.tiles.push_back(pos); P.tiles.push_back(c); P.tiles.push_back(sp); P.tiles.push_back(st);
        if (fill) {
          for (int t = 0; t < c;) {           // segments: the atoms of one structure in this round
            int segE = t + 1;
            while (segE < c && runSt[b + segE] == runSt[b + t]) segE++;
            for (int u = t; u < segE; u++) { P.perm[pos + u] = run[b + u]; P.seg[pos + u] = t | (segE << 16); }
            t = segE;
          }
        }
        pos += c;
        lastSp[rank] = sp;
        rank++;
      }
    }
    for (; rank < CS; rank++) { P.tiles.push_back(0); P.tiles.push_back(0); P.tiles.push_back(lastSp[rank]); P.tiles.push_back(st); }
    P.cost += (double)maxTiles + 1.0;
    P.nSuper++;
    st = e;
  }
  return true;
}
static int ensure_cluster_plan(fnetgpu_ctx *ctx, Slot &s) {
  if (s.nTilesC >= 0) return 0;
  s.nTilesC = 0;
  const NetTables &n = ctx->net;
  const int nSp = (int)s.spBeg.size() - 1;
  if (s.nA != 0 || s.nG < 1 || nSp < 1) return 0;
  const size_t smem = bpnn_mma_smem_bytes(n, 0, s.nG);
  if (smem > 227 * 1024) return 0;
  const char *env = getenv("FNETGPU_MLP_CLUSTER");          // 0: never; 2..8: pin the cluster size (tests, A/B)
  const int pin = env ? atoi(env) : -1;
  if (pin == 0) return 0;
  MmaKernelT kc = mma_cluster_kernel(n);
  if (fnet_smem_attr(kc, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  const int R = FNET_MMA_TA * FNET_MMA_WARPS;
  std::vector<int> spCount((size_t)s.nStruct * nSp, 0);
  int minCS = 1;
  for (int st = 0; st < s.nStruct; st++) {
    for (int i = s.h_offsets[st]; i < s.h_offsets[st + 1]; i++) spCount[(size_t)st * nSp + s.h_globalsp[i]]++;
    int rounds = 0;
    for (int sp = 0; sp < nSp; sp++) rounds += (spCount[(size_t)st * nSp + sp] + R - 1) / R;
    minCS = std::max(minCS, rounds);
  }
  if (minCS > 8) return 0;
  // two-pass estimate in the same unit (16-atom tiles of the slowest warp pair + 1 per round, per resident CTA):
  // the forward kernel costs about a third of the gradient kernel
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc, FNET_MMA_WARPS * 32, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return 0; }
  double cost2 = 0.0;
  int rounds2 = 0;
  for (int sp = 0; sp < nSp; sp++) {
    const int c = s.spBeg[sp + 1] - s.spBeg[sp];
    cost2 += (double)(c / R) * (R / FNET_MMA_TW + 1) + ((c % R) ? (double)((c % R + FNET_MMA_TW - 1) / FNET_MMA_TW + 1) : 0.0);
    rounds2 += (c + R - 1) / R;
  }
  double best = pin > 0 ? 1e300 : 1.35 * cost2 / (double)std::max(1, std::min(ctx->nSM * occ, rounds2));
  ClusterPlan P, bestP;
  int bestClusters = 0;
  for (int CS = std::max(minCS, pin > 0 ? pin : 1); CS <= (pin > 0 ? pin : 8); CS++) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * ctx->nSM * occ); cfg.blockDim = dim3(FNET_MMA_WARPS * 32); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kc, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); continue; }
    if (!build_cluster_plan(s, spCount, CS, false, P)) continue;
    const double est = P.cost / (double)std::min(nc, P.nSuper);
    if (est < best) { best = est; bestP.CS = CS; bestClusters = nc; }
  }
  if (bestP.CS == 0) return 0;
  if (!build_cluster_plan(s, spCount, bestP.CS, true, bestP)) return 0;
  if (dev_upload(ctx, &s.d_tilesC, bestP.tiles.data(), bestP.tiles.size())) return 1;
  if (dev_upload(ctx, &s.d_permC, bestP.perm.data(), bestP.perm.size())) return 1;
  if (dev_upload(ctx, &s.d_segBE, bestP.seg.data(), bestP.seg.size())) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // the vectors go out of scope
  s.clusterCS = bestP.CS;
  s.clusterGrid = bestP.CS * std::max(1, std::min(bestClusters, bestP.nSuper));
  s.nTilesC = bestP.nSuper;
  return 0;
}

template <typename real>
static int grad_t(fnetgpu_ctx *ctx, Slot &s, int lossId, double *ddSerial, double *loss, double *globalPred) {
  if (check_ready<real>(ctx, s, true)) return 1;
  if (wait_allreduce(ctx)) return 1;
  const NetTables &n = ctx->net;
  const size_t nDD = (size_t)n.nTot * n.nSpecies;
  const bool mma = use_mma<real>(ctx);
  // every structure inside one round of the DMMA kernel: E_s, loss gradient and loss terms are formed
  // in the gradient kernel itself -- no separate forward pass, no k_struct_loss
  bool fused = mma && s.nTilesS > 0 && s.nA == 0 && s.nG >= 1 && !ctx->mlpNoFuse;
  // ... or inside one super-round of a thread-block cluster
  bool cfused = false;
  if (mma && !fused && !ctx->mlpNoFuse) {
    if (ensure_cluster_plan(ctx, s)) return 1;
    cfused = s.nTilesC > 0;
  }
  if (!fused && !cfused) {
    if (run_forward<real>(ctx, s)) return 1;
    if (run_struct_loss<real>(ctx, s, lossId)) return 1;
  }
  const BpnnLaunch B = plan_bpnn<real>(ctx, s, 0);
  MmaLaunch M = plan_mma(ctx, s, 0);
  if (fused) M.grid = std::max(1, std::min(M.grid, s.nTilesS));
  if (cfused) { M.grid = s.clusterGrid; M.smem = bpnn_mma_smem_bytes(n, 0, s.nG); }
  const int grid = mma ? M.grid : B.grid;
  s.lastGrad[0] = cfused ? 2 : (fused ? 1 : 0); s.lastGrad[1] = cfused ? s.clusterCS : 1; s.lastGrad[2] = grid;
  s.lastGrad[3] = cfused ? s.nTilesC : (fused ? s.nTilesS : (mma ? s.nTiles16 : s.nTiles));
  size_t need = (size_t)grid * nDD;
  if (ctx->partialsN < need) { if (dev_alloc(ctx, &ctx->d_partials, need)) return 1; ctx->partialsN = need; }
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_partials, 0, need * sizeof(double), ctx->stream));
  bool launched = false;
  if constexpr (std::is_same<real, double>::value) {
    if (mma) {
      const MmaLayout ml = mma_layout(n);
      const int perWarp = (ml.nGradTiles + FNET_MMA_WARPS - 1) / FNET_MMA_WARPS;
#define FNET_MMA_GRAD(NSLOT)                                                                                      \
      do { if (n.dims[0] <= 32) FNET_MMA_GRAD2(NSLOT, 1); else FNET_MMA_GRAD2(NSLOT, 2); } while (0)
#define FNET_MMA_GRAD2(NSLOT, FCH)                                                                                \
      do {                                                                                                        \
        if (fused) {                                                                                              \
          CUDA_TRY(ctx, fnet_smem_attr(k_bpnn_mma<0, NSLOT, FCH, 1>, M.smem)); \
          LAUNCH(ctx, K_MLP_GRAD, (k_bpnn_mma<0, NSLOT, FCH, 1><<<grid, FNET_MMA_WARPS * 32, M.smem, ctx->stream>>>( \
                                      s.nTilesS, s.d_tilesS, s.d_perm, (const double *)s.d_feat, s.nFeat,         \
                                      (const double *)ctx->d_wb, n, s.d_structOf, s.d_offsets, nullptr, s.d_at,   \
                                      s.d_aw, s.d_dsw, s.nG, s.nA, lossId, ctx->d_partials, (double *)nullptr,    \
                                      s.d_gt, s.d_Es, s.d_lossPart, nullptr)));                                   \
          break;                                                                                                  \
        }                                                                                                         \
        CUDA_TRY(ctx, fnet_smem_attr(k_bpnn_mma<0, NSLOT, FCH>, M.smem)); \
        LAUNCH(ctx, K_MLP_GRAD, (k_bpnn_mma<0, NSLOT, FCH><<<grid, FNET_MMA_WARPS * 32, M.smem, ctx->stream>>>(   \
                                    s.nTiles16, s.d_tiles16, s.d_perm, (const double *)s.d_feat, s.nFeat,         \
                                    (const double *)ctx->d_wb, n, s.d_structOf, s.d_offsets, s.d_gS, s.d_at,      \
                                    s.d_aw, s.d_dsw, s.nG, s.nA, lossId, ctx->d_partials, (double *)nullptr,      \
                                    nullptr, nullptr, nullptr, nullptr)));                                        \
      } while (0)
      if (cfused) {
        MmaKernelT kc = mma_cluster_kernel(n);
        CUDA_TRY(ctx, fnet_smem_attr(kc, M.smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FNET_MMA_WARPS * 32); cfg.dynamicSmemBytes = M.smem; cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = s.clusterCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        LAUNCH(ctx, K_MLP_GRAD, (cudaLaunchKernelEx(&cfg, kc, s.nTilesC, (const int *)s.d_tilesC, (const int *)s.d_permC,
                                                    (const double *)s.d_feat, s.nFeat, (const double *)ctx->d_wb, n,
                                                    (const int *)s.d_structOf, (const int *)s.d_offsets, (const double *)nullptr,
                                                    (const double *)s.d_at, (const double *)s.d_aw, (const double *)s.d_dsw, s.nG, s.nA,
                                                    lossId, ctx->d_partials, (double *)nullptr, (const double *)s.d_gt, s.d_Es,
                                                    s.d_lossPart, (const int *)s.d_segBE)));
      } else if (perWarp <= 4) FNET_MMA_GRAD(4); else FNET_MMA_GRAD(FNET_MMA_MAXSLOTS);
#undef FNET_MMA_GRAD
#undef FNET_MMA_GRAD2
      launched = true;
    }
  }
  if (!launched) {
    CUDA_TRY(ctx, fnet_smem_attr(k_bpnn<real, 0>, B.smem));
    LAUNCH(ctx, K_MLP_GRAD, (k_bpnn<real, 0><<<grid, B.threads, B.smem, ctx->stream>>>(
                                s.nTiles, s.d_tiles, s.d_perm, (const real *)s.d_feat, s.nFeat, (const real *)ctx->d_wb, n,
                                s.tileT, B.gInSmem, s.d_structOf, s.d_offsets, s.d_gS, s.d_at, s.d_aw, s.d_dsw, s.nG, s.nA,
                                lossId, ctx->d_partials, (real *)nullptr, (real *)nullptr)));
  }
  // (fused sums: the loss terms of the structures are reduced by one more CTA of the same launch)
  if (fused || cfused)
    LAUNCH(ctx, K_GRAD_REDUCE, (k_grad_reduce<<<(int)((nDD + 31) / 32) + 1, FNET_GRED_J * 32, 0, ctx->stream>>>(
                                   grid, (int)nDD, ctx->d_partials, ctx->d_dd, s.nStruct, s.d_lossPart, ctx->d_dd + nDD)));
  else
    LAUNCH(ctx, K_GRAD_REDUCE, (k_grad_reduce<<<(int)((nDD + 31) / 32), FNET_GRED_J * 32, 0, ctx->stream>>>(grid, (int)nDD, ctx->d_partials, ctx->d_dd)));
  if (ctx->nRanks > 1 && ctx->comm) {
    // gradient | loss numerator | denominator: ONE all-reduce, on its own stream -- when the caller does not fetch the
    // result now, it overlaps whatever comes next on the main stream (the next step's ACSF kernel)
    if (!ctx->arStream) {
      CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->arStream, cudaStreamNonBlocking));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evGrad, cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evAR, cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->evGrad, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->arStream, ctx->evGrad, 0));
    if (allreduce_sum(ctx, ctx->d_dd, nDD + 2, ctx->arStream)) return 1;
    CUDA_TRY(ctx, cudaEventRecord(ctx->evAR, ctx->arStream));
    ctx->arPending = true;
  }
  if (ctx->reguLambda != 0.0 || ctx->reguDiv > 0.0) {   // regularised, normalised gradient (once, after the all-reduce)
    if (wait_allreduce(ctx)) return 1;
    const int nW = n.boff[0];
    const double c = ctx->reguLambda / (double)nW, inv = ctx->reguDiv > 0.0 ? 1.0 / ctx->reguDiv : 1.0;
    if (ctx->precision == 64)
      LAUNCH(ctx, K_MISC, (k_regularize<double><<<(int)((nDD + 127) / 128), 128, 0, ctx->stream>>>(n.nSpecies, n.nTot, nW, (const double *)ctx->d_wb, ctx->d_dd, c, ctx->reguAlpha, inv)));
    else
      LAUNCH(ctx, K_MISC, (k_regularize<double><<<(int)((nDD + 127) / 128), 128, 0, ctx->stream>>>(n.nSpecies, n.nTot, nW, ctx->d_wb64, ctx->d_dd, c, ctx->reguAlpha, inv)));
  }
  if (ddSerial || loss) {
    if (wait_allreduce(ctx)) return 1;
    if (ensure_pinned(ctx, nDD + 8)) return 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dd, (nDD + 2) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ddSerial) memcpy(ddSerial, ctx->h_pinned, nDD * sizeof(double));
    if (loss) *loss = ctx->h_pinned[nDD] / ctx->h_pinned[nDD + 1];
  }
  if (globalPred && s.nG > 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(globalPred, s.d_Es, (size_t)s.nG * s.nStruct * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int fnetgpu_grad(fnetgpu_ctx *ctx, int slot, int lossId, const int *shuffle, double *ddSerial,
                            double *loss, double *globalPred) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  (void)shuffle;   // only permutes the summation order in the reference (bpnn.F90:436-437)
  if (lossId < 0 || lossId > FNETGPU_LOSS_MAPE) FNET_FAIL(ctx, "grad: unknown loss id");
  cudaSetDevice(ctx->device);
  Slot &s = ctx->slots[slot];
  if (ctx->precision == 64) return grad_t<double>(ctx, s, lossId, ddSerial, loss, globalPred);
  return grad_t<float>(ctx, s, lossId, ddSerial, loss, globalPred);
}

template <typename real>
static int predict_t(fnetgpu_ctx *ctx, Slot &s, double *raw) {
  if (check_ready<real>(ctx, s, false)) return 1;
  if (run_forward<real>(ctx, s)) return 1;
  if (raw) return download_real(ctx, s.d_raw, (size_t)s.N * ctx->net.nOut, raw);
  return 0;
}

extern "C" int fnetgpu_predict(fnetgpu_ctx *ctx, int slot, double *raw) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  cudaSetDevice(ctx->device);
  Slot &s = ctx->slots[slot];
  if (ctx->precision == 64) return predict_t<double>(ctx, s, raw);
  return predict_t<float>(ctx, s, raw);
}

template <typename real>
static int loss_t(fnetgpu_ctx *ctx, Slot &s, int lossId, double *loss) {
  if (check_ready<real>(ctx, s, true)) return 1;
  if (wait_allreduce(ctx)) return 1;
  const size_t nDD = (size_t)ctx->net.nTot * ctx->net.nSpecies;
  if (run_forward<real>(ctx, s)) return 1;
  if (run_struct_loss<real>(ctx, s, lossId)) return 1;
  if (allreduce_sum(ctx, ctx->d_dd + nDD, 2)) return 1;
  double h[2];
  CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_dd + nDD, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (loss) *loss = h[0] / h[1];
  return 0;
}

extern "C" int fnetgpu_loss(fnetgpu_ctx *ctx, int slot, int lossId, double *loss) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  if (lossId < 0 || lossId > FNETGPU_LOSS_MAPE) FNET_FAIL(ctx, "loss: unknown loss id");
  cudaSetDevice(ctx->device);
  Slot &s = ctx->slots[slot];
  if (ctx->precision == 64) return loss_t<double>(ctx, s, lossId, loss);
  return loss_t<float>(ctx, s, lossId, loss);
}

// ------------------------------------------------------------------------------------------
// forces
// ------------------------------------------------------------------------------------------
// dE/dG [N][nOut][F] and forces [N][3 nOut] of a slot: grow-only, sized for the CURRENT ACSF / network
// configuration (a later fnetgpu_acsf_set / fnetgpu_net_set with a larger F or nOut re-allocates)
static int ensure_force_buffers(fnetgpu_ctx *ctx, Slot &s, size_t realBytes) {
  const size_t needD = (size_t)s.N * ctx->net.nOut * ctx->acsf.F * realBytes;
  if (!s.d_dEdG || s.capDEdGBytes < needD) {
    cudaFree(s.d_dEdG); s.d_dEdG = nullptr; s.capDEdGBytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&s.d_dEdG, std::max<size_t>(needD, 8)));
    s.capDEdGBytes = needD;
  }
  size_t capF = s.capForces;
  if (dev_reserve(ctx, &s.d_forces, &capF, (size_t)3 * ctx->net.nOut * s.N)) return 1;
  s.capForces = capF;
  return 0;
}
// one launch of the fused force kernel for the planned geometry path (no flag read-back)
// forcesOut: where the forces go (default: s.d_forces).  The deterministic whole-structure path only STORES them, so the
// socket step passes mapped pinned host memory and needs no device-to-host copy.
static int launch_acsf_forces(fnetgpu_ctx *ctx, Slot &s, const AcsfLaunch &L, const double *dEdG64, const double *zp,
                              double *forcesOut = nullptr) {
  const AcsfTables &T = ctx->acsf;
  const NetTables &n = ctx->net;
  const dim3 g(L.grid.x, L.grid.y, n.nOut);
  const GeomArgs geo = geom_args(s);
  double *fout = forcesOut ? forcesOut : s.d_forces;
  if (L.lean) {
    const LeanTables &LT = ctx->lean;
    const int localAtoms = L.local ? ((s.maxAtoms + 1) & ~1) : 0;
    double *fpart = nullptr;
    if (L.local && L.nSplit > 1) {       // partial forces of the CTAs of a structure
      const size_t need = (size_t)s.nStruct * L.nSplit * n.nOut * 3 * localAtoms;
      if (!ctx->d_fpart || ctx->fpartN < need) ctx->fpartZeroN = 0;
      if (dev_reserve(ctx, &ctx->d_fpart, &ctx->fpartN, need)) return 1;
      // CTAs without atoms write nothing: the partials start from zero.  k_force_reduce clears what it has summed, so
      // only the first use of the buffer (or a larger one) needs the memset
      if (ctx->fpartZeroN < need) CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_fpart, 0, ctx->fpartN * sizeof(double), ctx->stream));
      ctx->fpartZeroN = 0;                  // until the reduction below has been enqueued
      fpart = ctx->d_fpart;
    }
#define FNET_FLEAN(NL, NC, PATH, SORTED, G, LOCAL)                                                               \
  do {                                                                                                          \
    CUDA_TRY(ctx, fnet_smem_attr(k_acsf_force_lean<NL, NC, PATH, SORTED, G, LOCAL>, L.smem)); \
    LAUNCH(ctx, K_ACSF_FORCE, (fnet_launch_k(ctx->pdl, k_acsf_force_lean<NL, NC, PATH, SORTED, G, LOCAL>, g, dim3(L.wpb * 32), L.smem, ctx->stream, \
                                  L.nSplit, geo, T, LT, L.cap, L.capC, localAtoms, dEdG64, n.nOut, zp, fout, fpart, ctx->d_flags))); \
  } while (0)
#define FNET_FLEAN_S(NL, NC, PATH, LOCAL)                                                                        \
  do {                                                                                                          \
    if (ctx->leanSorted) { if (L.G == 2) FNET_FLEAN(NL, NC, PATH, true, 2, LOCAL); else FNET_FLEAN(NL, NC, PATH, true, 1, LOCAL); } \
    else if (L.G == 4) FNET_FLEAN(NL, NC, PATH, false, 4, LOCAL);                                                \
    else if (L.G == 2) FNET_FLEAN(NL, NC, PATH, false, 2, LOCAL);                                                \
    else FNET_FLEAN(NL, NC, PATH, false, 1, LOCAL);                                                              \
  } while (0)
#define FNET_FLEAN_P(NL, NC)                                                                                     \
  do { if (L.path == FNET_PATH_STRUCT) FNET_FLEAN_S(NL, NC, FNET_PATH_STRUCT, true); else FNET_FLEAN_S(NL, NC, FNET_PATH_STAGED, false); } while (0)
    if (ctx->leanNC == 1) FNET_FLEAN_P(2, 1);
    else if (ctx->leanNC == 2) FNET_FLEAN_P(2, 2);
    else FNET_FLEAN_P(1, 4);
#undef FNET_FLEAN_P
#undef FNET_FLEAN_S
#undef FNET_FLEAN
    if (fpart) {
      const dim3 rg((n.nOut * 3 * s.maxAtoms + 127) / 128, s.nStruct);
      LAUNCH(ctx, K_ACSF_FORCE, (fnet_launch_k(ctx->pdl, k_force_reduce, rg, dim3(128), 0, ctx->stream, s.nStruct, L.nSplit, n.nOut, localAtoms, (const int *)s.d_offsets, fpart, fout)));
      ctx->fpartZeroN = ctx->fpartN;
    }
    return 0;
  }
#define FNET_FORCE_LAUNCH(PATH)                                                                                 \
  do {                                                                                                          \
    CUDA_TRY(ctx, fnet_smem_attr(k_acsf_force<PATH>, L.smem)); \
    LAUNCH(ctx, K_ACSF_FORCE, (k_acsf_force<PATH><<<g, L.wpb * 32, L.smem, ctx->stream>>>(                      \
                                  L.nSplit, geo, s.nExt, s.d_ext, T, L.cap, L.capC, dEdG64, n.nOut, zp,         \
                                  fout, ctx->d_flags)));                                                        \
  } while (0)
  if (L.path == FNET_PATH_STRUCT) FNET_FORCE_LAUNCH(FNET_PATH_STRUCT);
  else if (L.path == FNET_PATH_STAGED) FNET_FORCE_LAUNCH(FNET_PATH_STAGED);
  else FNET_FORCE_LAUNCH(FNET_PATH_DIRECT);
#undef FNET_FORCE_LAUNCH
  return 0;
}


template <typename real>
static int forces_t(fnetgpu_ctx *ctx, Slot &s, double *forces) {
  if (check_ready<real>(ctx, s, false)) return 1;
  const AcsfTables &T = ctx->acsf;
  const NetTables &n = ctx->net;
  if (!ctx->acsfSet || T.F == 0) FNET_FAIL(ctx, "forces: need an ACSF configuration");
  if (!ctx->extIdx.empty()) FNET_FAIL(ctx, "forces: not defined with external features (initprogram.F90:1543-1548)");
  // (1) dE_k/dG for every atom and output: one reverse sweep per output
  if (ensure_force_buffers(ctx, s, sizeof(real))) return 1;
  if (run_ingrad<real>(ctx, s)) return 1;
  // the force kernel contracts in FP64
  const double *dEdG64 = nullptr;
  double *tmp64 = nullptr;
  const size_t nD = (size_t)s.N * n.nOut * T.F;
  if (ctx->precision == 64) dEdG64 = (const double *)s.d_dEdG;
  else {
    CUDA_TRY(ctx, cudaMalloc((void **)&tmp64, nD * sizeof(double)));
    LAUNCH(ctx, K_MISC, (k_convert_out<float><<<(int)((nD + 255) / 256), 256, 0, ctx->stream>>>(nD, (const float *)s.d_dEdG, tmp64)));
    dEdG64 = tmp64;
  }
  const double *zp = s.zscored ? ctx->d_zprec : nullptr;
  int h[16];
  for (int attempt = 0; attempt < 4; attempt++) {
    // same path selection as the value kernel; the forces are accumulated with atomics, so a retry
    // (neighbour-buffer overflow, lattice too small for the whole-structure path) starts from zero
    AcsfLaunch L;
    const bool sp = use_struct_path(ctx, s);
    if (sp) {
      if (plan_forces(ctx, s, true, L)) return 1;
    } else {
      if (ensure_cells(ctx, s, T.rcMax, nullptr)) return 1;
      if (ensure_neigh_count(ctx, s)) return 1;
      if (plan_forces(ctx, s, false, L)) return 1;
    }
    CUDA_TRY(ctx, cudaMemsetAsync(s.d_forces, 0, (size_t)3 * n.nOut * s.N * sizeof(double), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flags, 0, 8 * sizeof(int), ctx->stream));
    if (launch_acsf_forces(ctx, s, L, dEdG64, zp)) return 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_flags, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    s.lastPath = L.path;
    if (sp && h[4] != 0) { s.structPath = 0; s.maxNeigh = -1; continue; }
    if (h[1] == 0 && h[7] == 0) break;
    if (!sp || attempt == 3) break;            // cell-list path: capacities were counted exactly
    if (h[1] != 0) s.maxNeigh = h[1];
  }
  if (forces && h[1] == 0 && h[7] == 0) {
    CUDA_TRY(ctx, cudaMemcpyAsync(forces, s.d_forces, (size_t)3 * n.nOut * s.N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (tmp64) cudaFree(tmp64);
  if (h[1] != 0 || h[7] != 0) FNET_FAIL(ctx, "neighbour buffer overflow in the force kernel");
  return 0;
}

extern "C" int fnetgpu_forces(fnetgpu_ctx *ctx, int slot, double *forces) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  cudaSetDevice(ctx->device);
  Slot &s = ctx->slots[slot];
  if (ctx->precision == 64) return forces_t<double>(ctx, s, forces);
  return forces_t<float>(ctx, s, forces);
}

// ------------------------------------------------------------------------------------------
// one MD / i-PI step for a resident slot: new geometry in -> predictions + forces out
// (calculateMappingsForSocketComm + predictForSocketComm, prg_fnet/fortnet.F90:430-609).
// Fast path (precision 64, whole-structure neighbour search): everything is enqueued on the
// stream with the capacities of the previous step and the host synchronises ONCE; overflow /
// lattice flags are inspected afterwards and send the step through the general entry points.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// socket / MD step helpers: geometry in and flags out without copy nodes.  The step's device work is a chain of small
// dependent kernels; every cudaMemcpy / cudaMemset node in that chain costs as much as a kernel.  Inputs are staged in
// pinned host memory, which the device reads directly (k_sock_prologue also clears the flags); the subnetwork outputs
// and the forces are STORED straight into mapped pinned memory by the kernels that produce them; k_sock_epilogue
// publishes the flags.
// ------------------------------------------------------------------------------------------
__global__ void k_sock_prologue(int n3, const double *__restrict__ hCoords, double *__restrict__ coords, int n9,
                                const double *__restrict__ hLat, double *__restrict__ lat, int *__restrict__ flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  FNET_PDL_TRIGGER();
  if (t < 8) flags[t] = 0;
  for (int e = t; e < n3; e += nt) coords[e] = hCoords[e];
  for (int e = t; e < n9; e += nt) lat[e] = hLat[e];
}
__global__ void k_sock_epilogue(const int *__restrict__ flags, int *__restrict__ hFlags) {
  FNET_PDL_WAIT();
  if (threadIdx.x < 8) hFlags[threadIdx.x] = flags[threadIdx.x];
}
// (start, count <= FNET_WARP_WPB, species) entries of the species-sorted order for k_bpnn_warp
static int ensure_tiles_warp(fnetgpu_ctx *ctx, Slot &s) {
  if (s.nTilesW > 0) return 0;
  std::vector<int> t;
  for (int sp = 0; sp + 1 < (int)s.spBeg.size(); sp++)
    for (int b = s.spBeg[sp]; b < s.spBeg[sp + 1]; b += FNET_WARP_WPB) {
      t.push_back(b); t.push_back(std::min(FNET_WARP_WPB, s.spBeg[sp + 1] - b)); t.push_back(sp);
    }
  if (dev_upload(ctx, &s.d_tilesW, t.data(), t.size())) return 1;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  s.nTilesW = (int)t.size() / 3;
  return 0;
}
// the latency kernel serves small batches in precision 64 (FNETGPU_SOCK_MLP=mma keeps the throughput kernels: A/B)
static bool use_warp_mlp(const fnetgpu_ctx *ctx, const Slot &s) {
  static const bool off = [] { const char *e = getenv("FNETGPU_SOCK_MLP"); return e && !strcmp(e, "mma"); }();
  return !off && ctx->precision == 64 && s.N <= 16 * ctx->nSM && bpnn_warp_smem_bytes(ctx->net) <= 200 * 1024;
}
static int run_warp_mlp(fnetgpu_ctx *ctx, Slot &s, double *rawOut, double *dEdG) {
  if ((int)s.spBeg.size() - 1 > ctx->net.nSpecies) FNET_FAIL(ctx, "dataset references more species than the network has sub-networks");
  if (ensure_tiles_warp(ctx, s)) return 1;
  const size_t smem = bpnn_warp_smem_bytes(ctx->net);
  CUDA_TRY(ctx, fnet_smem_attr(k_bpnn_warp, smem));
  LAUNCH(ctx, K_MLP_WARP, (fnet_launch_k(ctx->pdl, k_bpnn_warp, dim3(s.nTilesW), dim3(FNET_WARP_WPB * 32), smem, ctx->stream,
                                         (const int *)s.d_tilesW, (const int *)s.d_perm, (const double *)s.d_feat, s.nFeat,
                                         (const double *)ctx->d_wb, ctx->net, rawOut, dEdG)));
  return 0;
}

extern "C" int fnetgpu_socket_step(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs,
                                   double *globalPred, double *atomicPred, double *forces) {
  CHECK_CTX(ctx); CHECK_SLOT(ctx, slot);
  Slot &s = ctx->slots[slot];
  if (!s.used) FNET_FAIL(ctx, "socket_step: empty slot");
  if (!coords) FNET_FAIL(ctx, "socket_step: coords missing");
  cudaSetDevice(ctx->device);
  const AcsfTables &T = ctx->acsf;
  const NetTables &n = ctx->net;
  if (!ctx->acsfSet || T.F == 0) FNET_FAIL(ctx, "socket_step: need an ACSF configuration");
  if (!ctx->extIdx.empty()) FNET_FAIL(ctx, "socket_step: not defined with external features (fortnet.F90:560-561)");
  if (!ctx->netSet || !ctx->paramsSet) FNET_FAIL(ctx, "socket_step: network / parameters not set");
  if (!s.d_feat || s.nFeat != T.F)
    FNET_FAIL(ctx, "socket_step: call fnetgpu_acsf_calculate on this slot once first (it fixes the standardisation)");
  if (s.zscored && !ctx->haveZ) FNET_FAIL(ctx, "socket_step: z-score statistics missing");
  const size_t nRaw = (size_t)s.N * n.nOut, nFrc = (size_t)3 * n.nOut * s.N;
  bool done = false;
  if (ctx->precision == 64) {
    for (int attempt = 0; attempt < 3 && !done && use_struct_path(ctx, s); attempt++) {
      if (latvecs) s.h_lat.assign(latvecs, latvecs + (size_t)9 * s.nStruct);
      s.cellRc = -1.0; s.neighStale = true; s.geomEpoch++;
      const double *zp = s.zscored ? ctx->d_zprec : nullptr;
      AcsfLaunch Lv, Lf;
      if (plan_values(ctx, s, true, Lv)) return 1;
      if (plan_forces(ctx, s, true, Lf)) return 1;
      if (ensure_pinned(ctx, nRaw + nFrc + 16)) return 1;
      const size_t nIn = (size_t)3 * s.N + (size_t)9 * s.nStruct;
      if (ctx->pinInN < nIn) {
        if (ctx->h_pinIn) cudaFreeHost(ctx->h_pinIn);
        ctx->h_pinIn = nullptr; ctx->pinInN = 0;
        CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_pinIn, nIn * sizeof(double)));
        ctx->pinInN = nIn;
      }
      double *hp = ctx->h_pinned, *hin = ctx->h_pinIn;
      const bool warpMlp = use_warp_mlp(ctx, s);
      // the step's device work: geometry in (read from pinned host memory), ACSF, subnetworks + input gradients,
      // forces, flags out; outputs and forces are stored straight into pinned host memory where the kernels allow it
      auto enqueue = [&](const double *srcCoords, const double *srcLat) -> int {
        LAUNCH(ctx, K_MISC, (k_sock_prologue<<<1, 256, 0, ctx->stream>>>(3 * s.N, srcCoords, s.d_coords, 9 * s.nStruct, srcLat, s.d_lat, ctx->d_flags)));
        if (launch_acsf_values<double>(ctx, s, Lv, zp)) return 1;
        s.featValid = true; s.lastPath = FNET_PATH_STRUCT;
        if (check_ready<double>(ctx, s, false)) return 1;
        if (ensure_force_buffers(ctx, s, sizeof(double))) return 1;
        if (warpMlp) {
          if (run_warp_mlp(ctx, s, hp, (double *)s.d_dEdG)) return 1;
        } else {
          if (run_forward<double>(ctx, s)) return 1;
          if (run_ingrad<double>(ctx, s)) return 1;
          CUDA_TRY(ctx, cudaMemcpyAsync(hp, s.d_raw, nRaw * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (Lf.local) {
          if (launch_acsf_forces(ctx, s, Lf, (const double *)s.d_dEdG, zp, hp + nRaw)) return 1;
        } else {                                            // atomics path accumulates on the device
          CUDA_TRY(ctx, cudaMemsetAsync(s.d_forces, 0, nFrc * sizeof(double), ctx->stream));
          if (launch_acsf_forces(ctx, s, Lf, (const double *)s.d_dEdG, zp)) return 1;
          CUDA_TRY(ctx, cudaMemcpyAsync(hp + nRaw, s.d_forces, nFrc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        LAUNCH(ctx, K_MISC, (fnet_launch_k(ctx->pdl, k_sock_epilogue, dim3(1), dim3(32), 0, ctx->stream, (const int *)ctx->d_flags, (int *)(hp + nRaw + nFrc))));
        return 0;
      };
      struct PdlScope { fnetgpu_ctx *c; PdlScope(fnetgpu_ctx *c_, bool on) : c(c_) { c->pdl = on; } ~PdlScope() { c->pdl = false; } } pdlScope(ctx, ctx->usePdl && !ctx->profiling);
      // the sequence depends on the launch plans and the buffers only: captured once as a CUDA graph, replayed per
      // MD step (one graph launch instead of 4-6 kernel launches, 3 memsets and 5 copies); the first step with a new
      // plan runs eagerly (it may allocate), profiling runs eagerly (per-kernel events)
      std::vector<long long> key = {
          (long long)s.N, s.nStruct, T.F, n.nOut, Lv.cap, Lv.capC, Lv.G, (long long)Lv.lean, (long long)Lv.smem, Lv.nSplit,
          Lf.cap, Lf.capC, Lf.G, (long long)Lf.lean, (long long)Lf.smem, Lf.nSplit, (long long)Lf.local, (long long)s.zscored,
          (long long)(size_t)s.d_coords, (long long)(size_t)s.d_lat, (long long)(size_t)s.d_feat, (long long)(size_t)s.d_raw,
          (long long)(size_t)s.d_dEdG, (long long)(size_t)s.d_forces, (long long)(size_t)hp, (long long)(size_t)hin,
          (long long)(size_t)ctx->d_wb, (long long)(size_t)ctx->stream, ctx->mlpLegacy, ctx->mlpNoFuse, ctx->acsfGeneric,
          (long long)(size_t)ctx->d_fpart, (long long)ctx->netEpoch, (long long)ctx->acsfEpoch, (long long)warpMlp,
          (long long)(size_t)s.d_tilesW, (long long)(ctx->fpartZeroN >= ctx->fpartN)};
      const bool graphOk = !ctx->profiling && ctx->useGraphs;
      memcpy(hin, coords, (size_t)3 * s.N * sizeof(double));
      memcpy(hin + (size_t)3 * s.N, s.h_lat.data(), (size_t)9 * s.nStruct * sizeof(double));
      if (graphOk && s.sockGraph && key == s.sockKey) {
        CUDA_TRY(ctx, cudaGraphLaunch(s.sockGraph, ctx->stream));
        ctx->launches += s.sockGraphLaunches;
        s.featValid = true; s.lastPath = FNET_PATH_STRUCT;
      } else if (graphOk && s.sockKeyWanted == key) {
        // second step with this plan: capture (everything is allocated by now), then replay
        if (s.sockGraph) { cudaGraphExecDestroy(s.sockGraph); s.sockGraph = nullptr; }
        const long long l0 = ctx->launches;
        cudaGraph_t g = nullptr;
        CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int erc = enqueue(hin, hin + (size_t)3 * s.N);
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        if (erc != 0 || ce != cudaSuccess || !g) {
          if (g) cudaGraphDestroy(g);
          cudaGetLastError();
          ctx->useGraphs = false;                       // no graphs on this context any more: eager steps
          s.sockKeyWanted.clear();
          if (enqueue(hin, hin + (size_t)3 * s.N)) return 1;
        } else {
          s.sockGraphLaunches = ctx->launches - l0;
          ctx->launches = l0;
          const cudaError_t ie = cudaGraphInstantiate(&s.sockGraph, g, 0);
          cudaGraphDestroy(g);
          if (ie != cudaSuccess) { s.sockGraph = nullptr; ctx->useGraphs = false; if (enqueue(hin, hin + (size_t)3 * s.N)) return 1; }
          else {
            s.sockKey = key;
            CUDA_TRY(ctx, cudaGraphLaunch(s.sockGraph, ctx->stream));
            ctx->launches += s.sockGraphLaunches;
          }
        }
      } else {
        if (enqueue(hin, hin + (size_t)3 * s.N)) return 1;
        s.sockKeyWanted = key;
      }
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      int h[8];
      memcpy(h, hp + nRaw + nFrc, sizeof(h));
      if (h[4] != 0) { s.structPath = 0; s.maxNeigh = -1; s.featValid = false; break; }   // lattice too small: general path below
      if (h[7] != 0) { s.featValid = false; break; }
      if (h[1] != 0) { s.maxNeigh = h[1]; s.featValid = false; continue; }               // neighbour buffers too small: retry
      if (Lv.lean && h[0] > 0 && h[0] != s.maxNeigh) s.maxNeigh = std::max(s.maxNeigh, h[0]);   // capacity hint only grows here: a stable plan keeps the graph
      if (atomicPred) memcpy(atomicPred, hp, nRaw * sizeof(double));
      if (forces) memcpy(forces, hp + nRaw, nFrc * sizeof(double));
      if (globalPred)
        for (int st = 0; st < s.nStruct; st++)
          for (int k = 0; k < n.nOut; k++) {
            double e = 0.0;
            for (int i = s.h_offsets[st]; i < s.h_offsets[st + 1]; i++) e += hp[(size_t)n.nOut * i + k];
            globalPred[(size_t)n.nOut * st + k] = e;
          }
      done = true;
    }
  }
  if (done) return 0;
  // general path: the blocking entry points, any precision, any structure size
  if (fnetgpu_coords_update(ctx, slot, coords, latvecs)) return 1;
  const int zs = s.zscored ? 1 : 0;
  if (ctx->precision == 64) { if (acsf_calculate_t<double>(ctx, s, zs, nullptr, zs ? 2 : 0)) return 1; }
  else { if (acsf_calculate_t<float>(ctx, s, zs, nullptr, zs ? 2 : 0)) return 1; }
  std::vector<double> raw(nRaw);
  if (fnetgpu_predict(ctx, slot, raw.data())) return 1;
  if (forces) { if (fnetgpu_forces(ctx, slot, forces)) return 1; }
  if (atomicPred) memcpy(atomicPred, raw.data(), nRaw * sizeof(double));
  if (globalPred)
    for (int st = 0; st < s.nStruct; st++)
      for (int k = 0; k < n.nOut; k++) {
        double e = 0.0;
        for (int i = s.h_offsets[st]; i < s.h_offsets[st + 1]; i++) e += raw[(size_t)n.nOut * i + k];
        globalPred[(size_t)n.nOut * st + k] = e;
      }
  return 0;
}

// ------------------------------------------------------------------------------------------
// live roofline denominators of THIS device (the pool's B200 boxes differ by up to 30 % in FP64-bound kernel time at the
// same reported clocks): FP64 FMA, FP64 tensor (DMMA m8n8k4) and FP32 FMA issue rates in T FMA/s.  ~30 ms.
// ------------------------------------------------------------------------------------------
template <typename T, int CH>
__global__ void k_peak_fma(int iters, T *out, T a, T b) {
  T acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = (T)(threadIdx.x + c);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = acc[c] * a + b;
  }
  T sres = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) sres += acc[c];
  if (sres == (T)123456789) out[0] = sres;
}
template <int CH>
__global__ void k_peak_dmma(int iters, double *out, double a, double b) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { c0[c] = threadIdx.x; c1[c] = c; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
  }
  double sres = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) sres += c0[c] + c1[c];
  if (sres == 123456789.0) out[0] = sres;
}
extern "C" int fnetgpu_measure_peaks(fnetgpu_ctx *ctx, double *out /* [3]: DFMA, DMMA, FFMA in T FMA/s */) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  const int iters = 2048, threads = 256, ctas = ctx->nSM * 8;
  double *scratch = (double *)ctx->d_flags;    // never written (the guard value cannot occur)
  auto timeit = [&](auto launch, double fmaPerLaunch, double &res) -> int {
    launch();
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int r = 0; r < 3; r++) launch();
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res = fmaPerLaunch * 3.0 / (ms * 1e-3) / 1e12;
    return 0;
  };
  const double lanes = (double)ctas * threads;
  if (timeit([&] { k_peak_fma<double, 8><<<ctas, threads, 0, ctx->stream>>>(iters, scratch, 1.0000001, 1e-9); }, lanes * iters * 8.0, out[0])) return 1;
  if (timeit([&] { k_peak_dmma<8><<<ctas, threads, 0, ctx->stream>>>(iters, scratch, 1.0000001, 1e-9); }, (double)ctas * (threads / 32) * iters * 8.0 * 256.0, out[1])) return 1;
  if (timeit([&] { k_peak_fma<float, 8><<<ctas, threads, 0, ctx->stream>>>(iters, (float *)scratch, 1.0000001f, 1e-9f); }, lanes * iters * 8.0, out[2])) return 1;
  return 0;
}

#include "multi.cuh"
