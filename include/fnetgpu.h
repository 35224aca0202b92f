/*
 * fnetgpu.h -- C ABI of libfnetgpu.so, the B200 (sm_100a) implementation of Fortnet's
 * data-parallel hot path: ACSF featurisation (+ analytic Cartesian derivatives) and the
 * per-species Behler-Parrinello subnetwork forward/backward.
 *
 * The reference (vanderhe/fortnet v0.7.4, Fortran 2008) has no FFI for this path; its seams
 * are type-bound procedures called from prg_fnet/fortnet.F90 and lib_nn/bpnn.F90.  Every
 * entry point below names the reference call site it replaces (paths relative to
 * /root/reference/prog/fortnet/).  The ISO_C_BINDING shim a maintainer adds on the Fortran
 * side is shown in INTEGRATION.md (fortnet_b200/host/fnet_gpu.F90).
 *
 * Conventions
 *   - plain C: ints are 32-bit, reals are FP64 regardless of the compute precision, all
 *     arrays are caller-owned HOST buffers laid out exactly as the Fortran driver holds them
 *     (column-major, so `array(F, nAtom)` is `feature fastest`); the library copies in/out
 *     during the call and owns all device memory behind the opaque context;
 *   - atom/species/feature indices that cross the boundary are 1-based like the reference's;
 *   - lengths are Bohr; coordinates Cartesian; latvecs[9*s + 3*k + c] = latVecs(c,k) of
 *     structure s (lib_dftbp/typegeometry.F90:24-59);
 *   - every function returns 0 on success, non-zero on error (message via
 *     fnetgpu_last_error); nothing throws, exits or falls back to the CPU.  The shim maps a
 *     non-zero status to `call error(msg)` (lib_dftbp/message.F90:73-103);
 *   - one host thread per context, calls are blocking unless stated otherwise.
 */
#ifndef FNETGPU_H
#define FNETGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fnetgpu_ctx fnetgpu_ctx;

/* activation ids (lib_nn/transfer.F90:54-342, lib_nn/layer.F90:88-150) */
enum {
  FNETGPU_ACT_GAUSSIAN = 0, FNETGPU_ACT_RELU = 1, FNETGPU_ACT_LRELU = 2, FNETGPU_ACT_SOFTPLUS = 3,
  FNETGPU_ACT_BENT = 4, FNETGPU_ACT_ATAN = 5, FNETGPU_ACT_SIGMOID = 6, FNETGPU_ACT_HEAVISIDE = 7,
  FNETGPU_ACT_TANH = 8, FNETGPU_ACT_LINEAR = 9
};
/* loss ids (lib_fortnet/initprogram.F90:1195-1226, lib_common/loss.F90:217-281,370-721) */
enum { FNETGPU_LOSS_MSE = 0, FNETGPU_LOSS_RMS = 1, FNETGPU_LOSS_MAE = 2, FNETGPU_LOSS_MAPE = 3 };
/* G-function types (lib_descriptors/acsf.F90:225-270) */
enum { FNETGPU_G1 = 1, FNETGPU_G2 = 2, FNETGPU_G3 = 3, FNETGPU_G4 = 4, FNETGPU_G5 = 5 };

#define FNETGPU_MAX_SLOTS 8        /* dataset slots: 0 = training set, 1 = validation set, ... */
#define FNETGPU_UNIQUE_ID_BYTES 128

/* ---- lifecycle: replaces TEnv_init / destructGlobalEnv (lib_dftbp/globalenv.F90:88-172) ---- */
/* device < 0: use $LOCAL_RANK if set, else device 0.  precision: 64 or 32 (compute/storage
 * type of features and activations; parity guarantees are stated for 64).  deterministic is
 * accepted for the reference's sake: features, predictions, the training gradient and the loss
 * are always reduced in a fixed order (bit-reproducible run to run); only the force scatter uses
 * FP64 atomics, i.e. forces are reproducible to the last bits, not bit for bit. */
int fnetgpu_init(fnetgpu_ctx **ctx, int device, int precision, int deterministic);
int fnetgpu_finalize(fnetgpu_ctx *ctx);
const char *fnetgpu_last_error(const fnetgpu_ctx *ctx);
int fnetgpu_synchronize(fnetgpu_ctx *ctx);

/* ---- dataset: what calculateMappings / nTrain receive as TDataset
 *      (prg_fnet/fortnet.F90:883-889, lib_nn/bpnn.F90:438-441) ---- */
int fnetgpu_dataset_upload(fnetgpu_ctx *ctx, int slot, int nStruct,
                           const int *natomOffsets /* [nStruct+1], 0-based prefix sums */,
                           const double *coords /* [3*N] */, const int *periodic /* [nStruct] */,
                           const double *latvecs /* [9*nStruct] */, const int *atnum /* [N] localAtToAtNum */,
                           const int *globalsp /* [N] localAtToGlobalSp, 1-based */,
                           const int *dsWeights /* [nStruct] or NULL (=1) */,
                           const double *atomicWeights /* [N] or NULL (=1) */,
                           int nG, const double *gTargets /* [nG*nStruct] */,
                           int nA, const double *aTargets /* [nA*N] */,
                           int nExt, const double *ext /* [nExt*N] = extFeatures(e,i) */);
/* new geometry for the same atoms (socket/MD path, prg_fnet/fortnet.F90:138-165) */
int fnetgpu_coords_update(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs);

/* ---- ACSF: TAcsf_init / TAcsf%calculate (lib_descriptors/acsf.F90:148-167, 540-639) ---- */
int fnetgpu_acsf_set(fnetgpu_ctx *ctx, int F, const int *type, const double *rcut, const double *kappa,
                     const double *rs, const double *eta, const double *lambda, const double *xi,
                     const int *atomid, const int *atomicnumbers /* [2*F] */);
/* external features appended after the ACSF block: TFeatures_collect
 * (lib_types/features.F90:200-265); extIndices are 1-based rows of ext */
int fnetgpu_features_config(fnetgpu_ctx *ctx, int nExtSel, const int *extIndices);
/* standardize != 0: z-score (acsf.F90:626-631).  have_zprec != 0: zprec[0..F) = means,
 * zprec[F..2F) = "variances" (population sigma) are inputs; else they are computed from this
 * slot with the dataset weights (acsf.F90:445-486) and returned. */
int fnetgpu_acsf_calculate(fnetgpu_ctx *ctx, int slot, int standardize, double *zprec, int have_zprec);
/* parity/debug: the assembled feature matrix array(nFeat, N) of the slot */
int fnetgpu_features_get(fnetgpu_ctx *ctx, int slot, double *out /* [nFeat*N] */);
/* bypass ACSF: load precomputed features (TFeatures%trainFeatures) */
int fnetgpu_features_set(fnetgpu_ctx *ctx, int slot, int nFeat, const double *in /* [nFeat*N] */);

/* ---- network: TBpnn_init + serializedWeightsAndBiases / serialWeightsAndBiasesFillup
 *      (lib_nn/bpnn.F90:96-143, 782-822; layout lib_nn/network.F90:397-459) ---- */
int fnetgpu_net_set(fnetgpu_ctx *ctx, int nSpecies, int nLayers, const int *dims, int activationId);
int fnetgpu_ntot(const fnetgpu_ctx *ctx);                       /* nWeights + nBiases per species */
int fnetgpu_params_set(fnetgpu_ctx *ctx, const double *wb /* [nTot*nSpecies] */);

/* ---- training gradient: TBpnn_updateGradients + loss (bpnn.F90:277-283, 317-323, 394-481)
 * ddSerial = TDerivs_serialized(resDd) (lib_common/nestedtypes.F90:428-470): un-normalised,
 * before regularisation -- exactly what TBpnn_update (bpnn.F90:708-778) consumes.
 * loss = loss(resPredicts, ...) with dataset weights.  globalPred (optional) = per-structure
 * sums of the first nG outputs.  shuffle is accepted for interface parity; it only permutes
 * the summation order (bpnn.F90:436-437) and is ignored.  With an initialised communicator
 * (fnetgpu_comm_init) ddSerial and loss are the sums over all ranks' shards.
 * ddSerial == NULL && loss == NULL: compute only, results stay on the device. */
int fnetgpu_grad(fnetgpu_ctx *ctx, int slot, int lossId, const int *shuffle,
                 double *ddSerial /* [nTot*nSpecies] */, double *loss, double *globalPred /* [nG*nStruct] or NULL */);

/* ---- prediction: TBpnn_predictBatch (bpnn.F90:1001-1058) and validation loss ---- */
/* TBpnn_update's host work on the gradient moved to the device (lib_nn/bpnn.F90:750-767): with strength > 0 the
 * elastic-net term of TWeightDerivs_elasticNetRegularization (lib_common/nestedtypes.F90:336-370; alpha = 0 ridge,
 * 1 lasso) is added to the weight entries of ddSerial, and with nDatapoints > 0 (sum(trainDataset%weights),
 * bpnn.F90:298-299) the gradient is divided by it -- both once, after the all-reduce.  strength = 0, nDatapoints <= 0
 * (the default) leave fnetgpu_grad's plain summed gradient.  fnetgpu_regularization_loss returns reguLoss of every
 * species (lib_common/loss.F90:119-196) for the current parameters, out[nSpecies]. */
int fnetgpu_regularization_set(fnetgpu_ctx *ctx, double strength, double alpha, double nDatapoints);
int fnetgpu_regularization_loss(fnetgpu_ctx *ctx, double *out);
int fnetgpu_predict(fnetgpu_ctx *ctx, int slot, double *raw /* [nOut*N] = predicts(t,i), dataset atom order */);
int fnetgpu_loss(fnetgpu_ctx *ctx, int slot, int lossId, double *loss);

/* ---- forces: TAcsf%calculatePrime + TBpnn%nJacobian + forceAnalysis_analytical fused
 *      (acsf.F90:643-717, bpnn.F90:937-997, lib_analysis/forces.F90:317-425).
 * forces[(3*nOut)*f + 3*t + c] = forces%geos(s)%array(c+3(t-1), f).  The dense
 * [3,F,N,N] tensor and the Jacobians never exist. */
int fnetgpu_forces(fnetgpu_ctx *ctx, int slot, double *forces /* [3*nOut*N] */);

/* ---- multi-GPU (one process per GPU): replaces the mpifx_allreduce of dw/db
 *      (bpnn.F90:460-467) by ONE NCCL all-reduce of [ddSerial | loss terms] on the
 *      library's stream, and the z-score statistics by two small all-reduces. ---- */
int fnetgpu_comm_unique_id(char *id /* [FNETGPU_UNIQUE_ID_BYTES] */);
int fnetgpu_comm_init(fnetgpu_ctx *ctx, int nRanks, int rank, const char *id);

/* fnetgpu_coords_update followed by fnetgpu_acsf_calculate as ONE blocking call (same arguments):
 * for datasets of small structures the host->device copy is cut into chunks of structures and
 * overlapped with the ACSF kernel of the chunks already on the device.  coords should be pinned
 * host memory for the overlap to happen (pageable memory works, the copies are then staged). */
int fnetgpu_acsf_update_calculate(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs_or_null,
                                  int standardize, double *zprec_inout, int have_zprec);

/* One MD / i-PI step for a resident slot -- replaces calculateMappingsForSocketComm +
 * predictForSocketComm (prg_fnet/fortnet.F90:430-609): new coordinates (and cell, or NULL when it is
 * unchanged) in; per-structure summed outputs globalPred[nOut*nStruct], per-atom outputs
 * atomicPred[nOut*N] and forces[3*nOut*N] out (any of them may be NULL).  The slot keeps the
 * topology, species and standardisation of the last fnetgpu_dataset_upload / fnetgpu_acsf_calculate
 * (call the latter once before the first step).  Precision 64 + small structures: one host
 * synchronisation per step.  Not defined with external features (fortnet.F90:560-561). */
int fnetgpu_socket_step(fnetgpu_ctx *ctx, int slot, const double *coords, const double *latvecs_or_null,
                        double *globalPred, double *atomicPred, double *forces);

/* ---- plumbing for benchmarks / profiling ---- */
int fnetgpu_set_stream(fnetgpu_ctx *ctx, void *cudaStream);   /* run on the caller's stream */
long long fnetgpu_launch_count(const fnetgpu_ctx *ctx);       /* kernels launched so far */
/* per-kernel CUDA-event timing: enable, run, then read (kernel ids: see fnetgpu_kernel_name) */
int fnetgpu_profile(fnetgpu_ctx *ctx, int enable);
int fnetgpu_profile_get(fnetgpu_ctx *ctx, int kernelId, double *ms_total, long long *launches);
const char *fnetgpu_kernel_name(int kernelId);                /* NULL past the last id */
int fnetgpu_max_neighbors(fnetgpu_ctx *ctx, int slot, int *maxNeigh, double *meanNeigh);
/* live roofline denominators of this device: FP64 FMA, FP64 tensor (DMMA) and FP32 FMA issue rates, T FMA/s (~30 ms) */
int fnetgpu_measure_peaks(fnetgpu_ctx *ctx, double *out /* [3] */);
/* Neighbour search of the ACSF kernels.  mode 0 (default): automatic -- datasets whose structures
 * all have <= 256 atoms and lattice-plane spacings >= 2 rc (or are clusters) take the
 * whole-structure / minimum-image path, everything else the cell list (what replaces
 * dynneighlist.F90:233-320 either way); mode 1: always the cell list (tests, A/B).  The
 * environment variable FNETGPU_ACSF_PATH=cells selects mode 1 at fnetgpu_init. */
int fnetgpu_acsf_path_set(fnetgpu_ctx *ctx, int mode);
/* path the last ACSF value / force launch of this slot took: 0 cell list (direct), 1 cell list
 * (candidates staged per bin), 2 whole structure; -1 before the first launch */
int fnetgpu_acsf_path_get(const fnetgpu_ctx *ctx, int slot);
/* ACSF value kernel.  mode 0 (default): configurations of the automatic parameter scheme
 * (TGFunctions_fromAutoScheme, acsf.F90:276-363: G2 on an arithmetic rs-ladder, G5 on xi-ladders from
 * xi = 1 with one common step, also after the species-resolved expansion, no atom-id scaling) run
 * k_acsf_lean (table-driven powers, closed-form diagonal), everything else k_acsf; mode 1: always
 * k_acsf (tests, A/B).  FNETGPU_ACSF_KERNEL=generic selects mode 1 at fnetgpu_init. */
int fnetgpu_acsf_kernel_set(fnetgpu_ctx *ctx, int mode);
/* 1: the configured functions run through k_acsf_lean, 0: k_acsf, -1: no configuration */
int fnetgpu_acsf_kernel_get(const fnetgpu_ctx *ctx);
/* what the last value launch of the slot actually used: info[0] = 1 k_acsf_lean / 0 k_acsf (the lean kernel needs the
 * candidates staged in shared memory; bins with too many candidates fall back), [1] central atoms per warp,
 * [2] neighbour capacity, [3] staged-candidate capacity, [4] path (as fnetgpu_acsf_path_get), [5] shared memory bytes */
int fnetgpu_acsf_launch_info(const fnetgpu_ctx *ctx, int slot, int *info /* [6] */);
/* what the last fnetgpu_grad of the slot launched: info[0] = 0 two passes (forward kernel, per-structure sums, gradient
 * kernel), 1 per-structure sums fused into the gradient kernel (every structure inside one 64-atom round), 2 the same
 * across a thread-block cluster (multi-species data / structures of up to 8 x 64 atoms: partial sums exchanged through
 * distributed shared memory); [1] = cluster size, [2] = grid (CTAs), [3] = rounds (super-rounds for 2).
 * FNETGPU_MLP_CLUSTER=0 disables mode 2, =2..8 pins its cluster size (tests, A/B). */
int fnetgpu_grad_launch_info(const fnetgpu_ctx *ctx, int slot, int *info /* [4] */);
/* Subnetwork kernels in precision 64.  mode 0 (default): FP64 tensor-core (DMMA) kernels when the
 * network fits their limits (sum of layer widths <= 128, <= 72 8x8 weight-gradient tiles, shared
 * memory), else the register-tiled DFMA kernels -- and, for single-species datasets with <= 64
 * atoms per structure and global targets only, the per-structure sums / loss gradients fused into
 * the gradient kernel (no separate forward pass); mode 1: always the register-tiled kernels;
 * mode 2: DMMA kernels without that fusion (tests, A/B; also FNETGPU_MLP=legacy|nofuse at
 * fnetgpu_init).  Precision 32 always uses the FFMA kernels. */
int fnetgpu_mlp_path_set(fnetgpu_ctx *ctx, int mode);
/* 1 if fnetgpu_grad / fnetgpu_predict would take the DMMA kernels for the current network */
int fnetgpu_mlp_path_get(const fnetgpu_ctx *ctx);

/* ---- single-process multi-GPU layer: replaces the MPI structure-level parallelism for a driver that
 *      stays ONE process -- getStartAndEndIndex (lib_common/parallel.F90:23-56) becomes a contiguous split of
 *      the structures over the devices balanced by atom count, the per-structure mpifx_allreduce of
 *      lib_nn/bpnn.F90:455-467 ONE ncclAllReduce of [ddSerial | loss terms] per iteration (communicators from
 *      ncclCommInitAll), the z-score statistics (lib_descriptors/acsf.F90:618-636) two small all-reduces.
 *      nDevicesRequested <= 0: all visible devices.  Arguments as for the per-device entry points, with the
 *      arrays of the WHOLE dataset: the layer shards on upload and gathers predictions, forces and features;
 *      gradient, loss and statistics come back replicated.  One host thread per device inside every call. ---- */
typedef struct fnetgpu_mg fnetgpu_mg;
int fnetgpu_mg_init(fnetgpu_mg **mg, int nDevicesRequested, int precision, int deterministic);
int fnetgpu_mg_finalize(fnetgpu_mg *mg);
const char *fnetgpu_mg_last_error(const fnetgpu_mg *mg);
int fnetgpu_mg_device_count(const fnetgpu_mg *mg);
fnetgpu_ctx *fnetgpu_mg_context(fnetgpu_mg *mg, int device);        /* per-device context (queries, tuning switches) */
int fnetgpu_mg_shard(const fnetgpu_mg *mg, int slot, int device, int *st0, int *st1, int *a0, int *a1);
int fnetgpu_mg_dataset_upload(fnetgpu_mg *mg, int slot, int nStruct, const int *offsets, const double *coords,
                              const int *periodic, const double *latvecs, const int *atnum, const int *globalsp,
                              const int *dsWeights, const double *atomicWeights, int nG, const double *gTargets, int nA,
                              const double *aTargets, int nExt, const double *ext);
int fnetgpu_mg_coords_update(fnetgpu_mg *mg, int slot, const double *coords, const double *latvecs);
int fnetgpu_mg_acsf_set(fnetgpu_mg *mg, int F, const int *type, const double *rcut, const double *kappa, const double *rs,
                        const double *eta, const double *lambda, const double *xi, const int *atomid, const int *atomicnumbers);
int fnetgpu_mg_features_config(fnetgpu_mg *mg, int nExtSel, const int *extIndices);
int fnetgpu_mg_acsf_calculate(fnetgpu_mg *mg, int slot, int standardize, double *zprec, int have_zprec);
int fnetgpu_mg_features_get(fnetgpu_mg *mg, int slot, double *out);
int fnetgpu_mg_net_set(fnetgpu_mg *mg, int nSpecies, int nLayers, const int *dims, int activationId);
int fnetgpu_mg_params_set(fnetgpu_mg *mg, const double *wb);
int fnetgpu_mg_grad(fnetgpu_mg *mg, int slot, int lossId, const int *shuffle, double *ddSerial, double *loss, double *globalPred);
int fnetgpu_mg_loss(fnetgpu_mg *mg, int slot, int lossId, double *loss);
int fnetgpu_mg_predict(fnetgpu_mg *mg, int slot, double *raw);
int fnetgpu_mg_forces(fnetgpu_mg *mg, int slot, double *forces);

#ifdef __cplusplus
}
#endif
#endif /* FNETGPU_H */
