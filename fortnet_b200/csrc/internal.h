// internal.h -- shared declarations of libfnetgpu (context, device tables, launch macros).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>

#include "../../include/fnetgpu.h"

struct CRec;
#define FNET_WARP 32
#define FNET_MAX_LAYERS 16
#define FNET_MAX_CODES 14          // distinct atomic numbers referenced by species-resolved functions
#define FNET_RCHUNK 8              // radial functions per register chunk
#define FNET_LADDER 8              // angular functions per ladder slot
#define FNET_SLOTS 4               // ladder slots per angular pass (=> 32 accumulators)

// ------------------------------------------------------------------------------------------
// kernel ids for launch counting / per-kernel event timing
// ------------------------------------------------------------------------------------------
enum KernelId {
  K_BIN_COUNT = 0, K_BIN_SCAN, K_BIN_FILL, K_BIN_SORT, K_NEIGH_COUNT, K_ACSF, K_ZSTAT, K_ZSTAT_FINAL,
  K_ZAPPLY, K_EXT_CONCAT, K_MLP_FWD, K_STRUCT_LOSS, K_LOSS_FINAL, K_MLP_GRAD, K_GRAD_REDUCE,
  K_MLP_INGRAD, K_ACSF_FORCE, K_MISC, K_MLP_WARP, K_NUM_KERNELS
};

struct StructInfo {
  double lat[9];     // lat[3*k+c]: component c of lattice (or bounding-box) vector k
  double inv[9];     // fractional s_k = sum_c inv[3*k+c] * (r_c - lo_c)
  double lo[3];      // origin of the bounding box (0 for periodic cells)
  int nb[3];         // bins per direction
  int D[3];          // bin reach per direction (1 unless the cell is thinner than rc)
  int binBase;       // first global bin of this structure
  int periodic;
  int atomBeg, atomEnd;
};

// Radial group: functions sharing (type, rc, atomId, species list); evaluated in chunks of
// FNET_RCHUNK functions, nChunksP2 = chunk count padded to a power of two <= 32.
struct RadialGroup {
  int type;          // 1..3
  int code;          // species code of the neighbour list, -1 = all neighbours
  int atomId;        // 0 or 1-based ext row
  int fBeg, fCnt;    // slice of the radial function tables
  int nChunksP2;     // 1, 2 or 4 chunks (<= 32 functions per group)
  double rc;
  // G2 on an arithmetic rs-ladder with one eta (auto scheme): Gaussian recurrence constants
  int ladder;
  double eta, rs0, drs;
  double kk[FNET_RCHUNK - 1];   // kk[m] = exp(-eta drs^2 (2m+1))
};

struct LadderSlot {
  double lam, xi0, dxi;
  int count;
  int cont;                    // continues the previous slot's ladder (same lam, dxi; xi0 = prev.xi0 + 8 dxi)
  int feat[FNET_LADDER];       // output feature index
  double pref[FNET_LADDER];    // 2^(1-xi)
  double xi[FNET_LADDER];      // exponents (used by the derivative kernel)
};

// Angular pass: up to FNET_SLOTS ladders evaluated in one sweep over the pairs of
// (list code1) x (list code2) with shared pair geometry (same type, rc, eta, atomId).
struct AngularPass {
  int type;          // 4 or 5
  int code1, code2;  // -1 = all neighbours
  int same;          // lists identical (unordered pairs incl. diagonal)
  int atomId;
  int nSlots;
  int fast;          // G5, nSlots == the kernel's NS, every slot a fresh ladder starting at xi = 1 (the auto scheme): straight-line pair loop
  int keepFc;        // same (rc, eta, atomId) as the previous pass: its per-neighbour factors fc * exp(-eta r^2) are still valid
  double rc, eta;
  LadderSlot slot[FNET_SLOTS];
};

struct AcsfTables {           // device pointers + sizes, passed by value to kernels
  int F;                      // number of ACSF functions
  int nRadialGroups, nAngularPasses;
  const RadialGroup *rgroups;
  const int *rfeat;           // radial function tables
  const double *rp1, *rp2;    // (eta, rs) for G2, (kappa, -) for G3
  const AngularPass *apasses;
  int nCodes;                 // number of species codes
  int zcodes[FNET_MAX_CODES]; // atomic number of each code
  int anyAtomId;
  int redRows;                // rows of the per-warp reduction scratch: 8 * max(NS, radial chunks)
  double rcMax;
};

// ---- lean ACSF kernel (acsf_lean.cuh): tables of the automatic parameter scheme ----
#define FNET_POW_TAB_N 256                 // mantissa intervals of the power table
#define FNET_POW_KMIN (-62)                // smallest binary exponent with an own 2^(k delta) entry (b <= 2 -> k <= 1)
#define FNET_POW_NK (2 - FNET_POW_KMIN)    // 64 entries
#define FNET_POW_DOUBLES (2 * FNET_POW_TAB_N + FNET_POW_NK)
#define FNET_PAIR_TAB_MAXN 64              // strict-triangle pair table covers lists of <= 64 neighbours
#define FNET_PAIR_TAB_N (FNET_PAIR_TAB_MAXN * (FNET_PAIR_TAB_MAXN - 1) / 2 + 1)
#define FNET_LEAN_MAXACC 32                // accumulators per lane and pass

struct LeanRadial {            // G2 group on an arithmetic rs-ladder with one eta (auto scheme, acsf.F90:320-336)
  int code;                    // species code of the neighbour list, -1 = all neighbours
  int fBeg, fCnt;              // slice of AcsfTables::rfeat
  int nch, lgn;                // chunks of 8 functions padded to a power of two, and its log2
  int sharedFc;                // rc equals LeanTables::rcShared: the per-neighbour cutoff values are reused
  double rc, invrc, eta, rs0, drs;
  double kk[FNET_RCHUNK - 1];  // kk[m] = exp(-eta drs^2 (2m+1))
  double kk7, c16;             // exp(-15 eta drs^2), exp(-16 eta drs^2): the value kernel chains two chunks (functions 8..15 from 0..7)
};

struct LeanPass {              // NL lambda-groups x NC chained 8-function slots of one xi-ladder each
  int code1, code2, same;
  int recomp;                  // (rc, eta) differ from the previous pass: per-neighbour factors are recomputed
  int m0;                      // first ladder index of this pass: xi = 1 + (m0 + f) delta
  double rc, invrc, eta;
  double lam[2];
  int feat[FNET_LEAN_MAXACC];      // output feature of accumulator e = (l * NC + chunk) * 8 + f, -1 = unused
  double pref[FNET_LEAN_MAXACC];   // 2^(1 - xi) (x 2 for identical lists: unordered pairs)
  double dA[FNET_LEAN_MAXACC];     // diagonal of identical lists: value += dA * sum fcE^2 + dB * sum fcE^2 eps
  double dB[FNET_LEAN_MAXACC];
};

struct LeanTables {            // passed by value (kernel parameter space = constant bank)
  int nRadial, nPasses;
  const LeanRadial *rad;       // device
  const LeanPass *pass;        // device
  const double *powtab;        // device: (1/c_i, c_i^delta) x 256, then 2^(k delta), k = KMIN .. 1
  const unsigned short *pairtab;   // device: p -> j | k << 8 of the strict triangle, entry p = k (k-1)/2 + j
  double powC[5];              // binom(delta, 1..5)
  int redRows;                 // rows of the per-warp reduction scratch: max(8 NL NC, 8)
  int stageBytes;              // pass + radial tables are copied to shared memory per CTA when they are this small (else 0)
  double rcShared, invrcShared, etaShared;   // (rc, eta) of the per-neighbour factors formed right after the gather
};

struct NetTables {
  int nSpecies, L, act, nTot, nOut;
  int dims[FNET_MAX_LAYERS];
  int woff[FNET_MAX_LAYERS], boff[FNET_MAX_LAYERS];
  int aoff[FNET_MAX_LAYERS];  // row offset of layer l activations in the smem tile
  int rowsA;                  // sum of dims
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void free_() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Slot {
  bool used = false;
  int nStruct = 0, N = 0, nG = 0, nA = 0, nExt = 0;
  // host copies needed later
  std::vector<int> h_offsets;
  std::vector<int> h_globalsp;      // 0-based
  std::vector<double> h_lat;
  std::vector<int> h_periodic;
  // device data
  int *d_offsets = nullptr, *d_structOf = nullptr, *d_atnum = nullptr, *d_sp = nullptr, *d_periodic = nullptr;
  double *d_coords = nullptr;       // [N][3] as given
  double *d_lat = nullptr;          // [nStruct][9] lattice vectors (whole-structure ACSF path)
  int maxAtoms = 0;                 // largest structure
  int lastPath = -1;                // path of the last ACSF launch (fnetgpu_acsf_path_get)
  int lastLaunch[6] = {0, 0, 0, 0, -1, 0};   // fnetgpu_acsf_launch_info
  // fnetgpu_socket_step as a CUDA graph: instantiated for one (launch plans, buffers) key
  cudaGraphExec_t sockGraph = nullptr;
  std::vector<long long> sockKey, sockKeyWanted;
  long long sockGraphLaunches = 0;
  int structPath = 1;               // whole-structure path allowed for the current lattices (reset by coords_update)
  double *d_fpos = nullptr;         // [N][3] folded positions (atom order)
  CRec *d_crec = nullptr;    // [N] 32-byte records in cell order (position, atom index, atomic number)
  int *d_binStruct = nullptr;       // [totalBins] structure of each bin
  StructInfo *d_sinfo = nullptr;
  int *d_atomCell = nullptr, *d_cellStart = nullptr, *d_cellCount = nullptr, *d_cellAtoms = nullptr;
  int totalBins = 0;
  size_t capCoords = 0, capFpos = 0, capCrec = 0, capBinStruct = 0, capAtomCell = 0, capCellAtoms = 0, capCellStart = 0, capCellCount = 0, capSinfo = 0;
  bool neighStale = true;           // maxNeigh is only a hint for the current geometry
  // a value launch whose overflow / lattice flags were read back and found clean: the same plan on the same geometry
  // cannot trip them again, so the repeat launches of a resident slot need no flag read-back and no host synchronisation
  unsigned long long geomEpoch = 1, okEpoch = 0;
  int okCap = 0, okCapC = 0, okPath = -1, okG = 0, okLean = -1;
  double cellRc = -1.0;             // cutoff the cell list was built for
  double *d_dsw = nullptr;          // [nStruct] dataset weights as double
  double *d_aw = nullptr;           // [N]
  double *d_gt = nullptr, *d_at = nullptr, *d_ext = nullptr;
  void *d_feat = nullptr;           // [N][nFeat] real
  int nFeat = 0;
  bool featValid = false;
  bool zscored = false;             // features were standardised with ctx->d_zprec
  int maxNeigh = -1; double meanNeigh = 0.0;
  int maxCand = -1;                 // max candidates (atoms in the neighbour cells) of a bin
  int maxBinPop = 0;                // largest bin population (hint, from the last cell-list build)
  int maxCells = 0;                 // max neighbour cells per bin: (2D0+1)(2D1+1)(2D2+1)
  // species-sorted processing order
  int *d_perm = nullptr;            // [N] atom ids sorted by species (stable)
  std::vector<int> spBeg;           // [nSpecies+1]
  int *d_tiles = nullptr; int nTiles = 0, tileT = 0; // (start, count, species) triples
  int *d_tilesS = nullptr; int nTilesS = 0;          // structure-aligned rounds (fused per-structure sums), 0 = not applicable
  int *d_tiles16 = nullptr; int nTiles16 = 0;        // rounds (64 atoms of one species) of the FP64 tensor-core path (mlp_mma.cuh)
  // cluster-fused per-structure sums (k_bpnn_mma<0,..,2>): super-rounds of clusterCS rounds holding whole structures
  int *d_tilesC = nullptr, *d_permC = nullptr, *d_segBE = nullptr;
  int nTilesC = -1;                 // super-rounds; -1 = not planned yet, 0 = not applicable
  int clusterCS = 0, clusterGrid = 0;
  int lastGrad[4] = {0, 0, 0, 0};   // fnetgpu_grad_launch_info
  int *d_tilesW = nullptr; int nTilesW = 0;          // <= 4 atoms of one species per entry: the one-warp-per-atom latency kernel (mlp_warp.cuh)
  // work buffers
  void *d_raw = nullptr;            // [N][nOut] real
  double *d_gS = nullptr;           // [nStruct][nG] loss gradients of the global targets
  double *d_Es = nullptr;           // [nStruct][nG]
  double *d_lossPart = nullptr;     // [nStruct][2]
  void *d_dEdG = nullptr;           // [N][nG][F] real
  size_t capDEdGBytes = 0;
  double *d_forces = nullptr;       // [N][3*nOut]
  size_t capForces = 0;
};

struct fnetgpu_ctx {
  int device = 0;
  int precision = 64;
  int deterministic = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;   // chunked host->device copies overlapped with the ACSF kernel (fnetgpu_acsf_update_calculate)
  cudaStream_t arStream = nullptr;     // gradient all-reduce: overlaps whatever the caller enqueues next (the next step's ACSF kernel)
  cudaEvent_t evGrad = nullptr, evAR = nullptr;
  bool arPending = false;              // an all-reduce on arStream still owns d_dd
  cudaEvent_t evChunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ownStream = true;
  std::string err;
  int nSM = 148;
  long long launches = 0;
  bool profiling = false;
  double kms[K_NUM_KERNELS];
  long long klaunch[K_NUM_KERNELS];
  long long kprof[K_NUM_KERNELS];
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  struct ProfEvent *prof = nullptr;   // ring of event pairs (allocated on first use)
  int profUsed = 0;
  // ACSF config
  bool acsfSet = false;
  AcsfTables acsf;                  // device pointers inside
  std::vector<RadialGroup> h_rgroups;
  std::vector<AngularPass> h_apasses;
  int maxSlots = 1;                 // largest nSlots of any angular pass (selects the kernel instantiation)
  RadialGroup *d_rgroups = nullptr; int *d_rfeat = nullptr; double *d_rp1 = nullptr, *d_rp2 = nullptr;
  AngularPass *d_apasses = nullptr;
  // lean kernel (acsf_lean.cuh): the configuration is an automatic-scheme one
  bool leanOK = false;
  int leanNL = 2, leanNC = 1;       // accumulator shape of the instantiation: lambda-groups x chained 8-function slots
  bool leanSorted = false;          // species-resolved functions present
  LeanTables lean;                  // device pointers inside
  LeanRadial *d_lrad = nullptr; LeanPass *d_lpass = nullptr; double *d_powtab = nullptr; unsigned short *d_pairtab = nullptr;
  int acsfGeneric = 0;              // FNETGPU_ACSF_KERNEL=generic / acsf_kernel_set(1): always k_acsf (tests, A/B)
  std::vector<int> extIdx;          // 0-based rows of ext appended to the features
  int *d_extIdx = nullptr;
  double *d_zprec = nullptr;        // [2F] means, sigmas
  bool haveZ = false;
  // network
  bool netSet = false;
  NetTables net;
  void *d_wb = nullptr;             // [nSpecies][nTot] real
  double *d_wb64 = nullptr;
  double *d_fpart = nullptr; size_t fpartN = 0;   // partial forces of structures split over several CTAs (k_force_reduce)
  size_t fpartZeroN = 0;            // leading elements of d_fpart known to be zero (k_force_reduce clears what it has summed)
  double *d_conv = nullptr; size_t convN = 0;   // FP32 mode: device-side float -> double staging of downloads
  bool paramsSet = false;
  double reguLambda = 0.0, reguAlpha = 0.0, reguDiv = 0.0;   // fnetgpu_regularization_set
  // gradient work
  double *d_partials = nullptr; size_t partialsN = 0;
  double *d_dd = nullptr;           // [nSpecies*nTot + 2] reduced gradient + loss numerator/denominator
  double *h_pinned = nullptr; size_t pinnedN = 0;
  double *h_pinIn = nullptr; size_t pinInN = 0;   // socket step: geometry staged for the graph's host-to-device copies
  bool useGraphs = true;            // FNETGPU_GRAPHS=0: eager socket steps
  bool pdl = false;                 // set while the socket step enqueues: kernels are launched with programmatic stream
                                    // serialization (the next kernel's launch and table staging overlap the previous kernel)
  bool usePdl = true;               // FNETGPU_PDL=0 switches it off (A/B)
  long long acsfEpoch = 0, netEpoch = 0;
  int *d_flags = nullptr;           // [8] statistics / overflow flags (cells.cuh, acsf.cuh)
  int mlpNoFuse = 0;                // FNETGPU_MLP=nofuse / mlp_path_set(2): DMMA kernels without the fused per-structure sums (tests, A/B)
  int mlpLegacy = 0;                // FNETGPU_MLP=legacy: register-tiled DFMA kernels of mlp.cuh also in precision 64 (tests, A/B)
  int acsfPathCells = 0;            // FNETGPU_ACSF_PATH=cells: never use the whole-structure path (tests, A/B)
  // comm
  void *nccl = nullptr;             // dlopen handle
  void *comm = nullptr;             // ncclComm_t
  int nRanks = 1, rank = 0;
  Slot slots[FNETGPU_MAX_SLOTS];
};

#define CUDA_TRY(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                       \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define FNET_FAIL(ctx, msg)                                                                  \
  do { (ctx)->err = (msg); return 1; } while (0)

// Launch wrapper: counts launches and (in profile mode) brackets the kernel with a pair of CUDA
// events on the launching stream.  Events are resolved lazily (fnetgpu_profile_get), so profiling
// adds no host synchronisation to the timed region.
#define FNET_PROF_RING 8192
struct ProfEvent { cudaEvent_t a, b; int kernel; };

#define LAUNCH(ctx, kid, ...)                                                                \
  do {                                                                                       \
    ProfEvent *_pe = nullptr;                                                                \
    if ((ctx)->profiling && (ctx)->profUsed < FNET_PROF_RING) {                              \
      _pe = &(ctx)->prof[(ctx)->profUsed++];                                                 \
      _pe->kernel = (kid);                                                                      \
      cudaEventRecord(_pe->a, (ctx)->stream);                                                \
    }                                                                                        \
    __VA_ARGS__;                                                                             \
    (ctx)->launches++; (ctx)->klaunch[kid]++;                                                \
    if (_pe) cudaEventRecord(_pe->b, (ctx)->stream);                                         \
    cudaError_t _le = cudaGetLastError();                                                    \
    if (_le != cudaSuccess) {                                                                \
      (ctx)->err = std::string("kernel launch failed (") + fnetgpu_kernel_name(kid) + "): " + \
                   cudaGetErrorString(_le);                                                  \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// Programmatic dependent launch (griddepcontrol, sm_90+): a kernel launched with the attribute may start while its
// predecessor in the stream is still running; it must not touch anything the predecessor produces before FNET_PDL_WAIT()
// (which returns when the predecessor has completed and its writes are visible).  Every kernel of the socket step
// triggers its dependents at once (FNET_PDL_TRIGGER) -- the step is a chain of small kernels on a mostly idle GPU, so
// the next kernel's launch latency, CTA scheduling and constant-table staging disappear behind the current kernel.
// Without the launch attribute both instructions are no-ops.
#define FNET_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define FNET_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
template <typename... KA, typename... A>
static inline void fnet_launch_k(bool pdl, void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t st, A &&...a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, std::forward<A>(a)...);
}

template <typename T>
static inline int dev_alloc(fnetgpu_ctx *ctx, T **p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  CUDA_TRY(ctx, cudaMalloc((void **)p, n * sizeof(T)));
  return 0;
}
// grow-only allocation: keeps the buffer when it is already large enough (hot e2e path)
template <typename T>
static inline int dev_reserve(fnetgpu_ctx *ctx, T **p, size_t *cap, size_t n) {
  if (*p && *cap >= n) return 0;
  if (*p) { cudaFree(*p); *p = nullptr; }
  *cap = 0;
  if (n == 0) n = 1;
  CUDA_TRY(ctx, cudaMalloc((void **)p, n * sizeof(T)));
  *cap = n;
  return 0;
}
template <typename T>
static inline int dev_upload(fnetgpu_ctx *ctx, T **p, const T *h, size_t n) {
  if (dev_alloc(ctx, p, n)) return 1;
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(*p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
