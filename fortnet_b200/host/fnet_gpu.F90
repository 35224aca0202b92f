!> ISO_C_BINDING shim between Fortnet's Fortran driver and libfnetgpu.so (include/fnetgpu.h).
!!
!! SOURCE ONLY: this image has no Fortran compiler, so the module is shipped as the binding a
!! Fortnet maintainer adds (see INTEGRATION.md for the call sites it replaces); it contains no
!! arithmetic.  The tested mirror of the same call order is fortnet_b200/host/fnetgpu.hpp (C++)
!! and fortnet_b200/context.py (ctypes).
module fnet_gpu

  use, intrinsic :: iso_c_binding
  use dftbp_accuracy, only: dp
  use dftbp_message, only : error
  implicit none
  private

  public :: TGpuEnv, TGpuEnv_init, TGpuEnv_final
  public :: gpuUploadDataset, gpuCoordsUpdate, gpuAcsfSet, gpuFeaturesConfig, gpuAcsfCalculate, gpuNetSet, gpuParamsSet
  public :: gpuSetRegularization, gpuUpdateGradients, gpuLoss, gpuPredictBatch, gpuForces, gpuSocketStep
  public :: gpuUnserialize, TGpuEnv_initRank, gpuCommUniqueId

  !> nDevices = 1: one context (ctx); nDevices > 1: the single-process multi-GPU layer (mg, fnetgpu_mg_*), which
  !! shards the structures over the devices and all-reduces the gradient -- the driver stays ONE process and is
  !! built with WITH_MPI = FALSE.  (An MPI build instead keeps one rank per GPU: TGpuEnv_initRank.)
  type :: TGpuEnv
    type(c_ptr) :: ctx = c_null_ptr
    type(c_ptr) :: mg = c_null_ptr
    integer :: nDevices = 1
  end type TGpuEnv

  interface
    integer(c_int) function fnetgpu_init(ctx, device, precision, deterministic) bind(C, name='fnetgpu_init')
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device, precision, deterministic
    end function
    integer(c_int) function fnetgpu_finalize(ctx) bind(C, name='fnetgpu_finalize')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    type(c_ptr) function fnetgpu_last_error(ctx) bind(C, name='fnetgpu_last_error')
      import :: c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fnetgpu_dataset_upload(ctx, slot, nStruct, offsets, coords, periodic, latvecs,&
        & atnum, globalsp, dsWeights, atomicWeights, nG, gTargets, nA, aTargets, nExt, ext)&
        & bind(C, name='fnetgpu_dataset_upload')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot, nStruct, nG, nA, nExt
      integer(c_int), intent(in) :: offsets(*), periodic(*), atnum(*), globalsp(*), dsWeights(*)
      real(c_double), intent(in) :: coords(*), latvecs(*), atomicWeights(*), gTargets(*), aTargets(*), ext(*)
    end function
    integer(c_int) function fnetgpu_acsf_set(ctx, nFunc, ftype, rcut, kappa, rs, eta, lambda, xi, atomid,&
        & atomicnumbers) bind(C, name='fnetgpu_acsf_set')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: nFunc
      integer(c_int), intent(in) :: ftype(*), atomid(*), atomicnumbers(*)
      real(c_double), intent(in) :: rcut(*), kappa(*), rs(*), eta(*), lambda(*), xi(*)
    end function
    integer(c_int) function fnetgpu_acsf_calculate(ctx, slot, standardize, zprec, have_zprec)&
        & bind(C, name='fnetgpu_acsf_calculate')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot, standardize, have_zprec
      real(c_double), intent(inout) :: zprec(*)
    end function
    integer(c_int) function fnetgpu_net_set(ctx, nSpecies, nLayers, dims, activationId) bind(C, name='fnetgpu_net_set')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nSpecies, nLayers, activationId
      integer(c_int), intent(in) :: dims(*)
    end function
    integer(c_int) function fnetgpu_params_set(ctx, wb) bind(C, name='fnetgpu_params_set')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: wb(*)
    end function
    integer(c_int) function fnetgpu_grad(ctx, slot, lossId, shuffle, ddSerial, loss, globalPred) bind(C, name='fnetgpu_grad')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx, shuffle, globalPred
      integer(c_int), value :: slot, lossId
      real(c_double), intent(out) :: ddSerial(*), loss
    end function
    integer(c_int) function fnetgpu_predict(ctx, slot, raw) bind(C, name='fnetgpu_predict')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot
      real(c_double), intent(out) :: raw(*)
    end function
    integer(c_int) function fnetgpu_forces(ctx, slot, forces) bind(C, name='fnetgpu_forces')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot
      real(c_double), intent(out) :: forces(*)
    end function
    integer(c_int) function fnetgpu_socket_step(ctx, slot, coords, latvecs, globalPred, atomicPred, forces)&
        & bind(C, name='fnetgpu_socket_step')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot
      real(c_double), intent(in) :: coords(*)
      type(c_ptr), value :: latvecs                     ! c_null_ptr: cell unchanged
      real(c_double), intent(out) :: globalPred(*), atomicPred(*), forces(*)
    end function
    integer(c_int) function fnetgpu_coords_update(ctx, slot, coords, latvecs) bind(C, name='fnetgpu_coords_update')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx, latvecs
      integer(c_int), value :: slot
      real(c_double), intent(in) :: coords(*)
    end function
    integer(c_int) function fnetgpu_features_config(ctx, nExtSel, extIndices) bind(C, name='fnetgpu_features_config')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nExtSel
      integer(c_int), intent(in) :: extIndices(*)
    end function
    integer(c_int) function fnetgpu_loss(ctx, slot, lossId, loss) bind(C, name='fnetgpu_loss')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: slot, lossId
      real(c_double), intent(out) :: loss
    end function
    integer(c_int) function fnetgpu_regularization_set(ctx, strength, alpha, nDatapoints)&
        & bind(C, name='fnetgpu_regularization_set')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), value :: strength, alpha, nDatapoints
    end function
    integer(c_int) function fnetgpu_comm_unique_id(id) bind(C, name='fnetgpu_comm_unique_id')
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
    end function
    integer(c_int) function fnetgpu_comm_init(ctx, nRanks, rank, id) bind(C, name='fnetgpu_comm_init')
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: ctx
      integer(c_int), value :: nRanks, rank
      character(kind=c_char), intent(in) :: id(128)
    end function
    ! ---- single-process multi-GPU layer (same argument lists with the arrays of the WHOLE dataset) ----
    integer(c_int) function fnetgpu_mg_init(mg, nDevices, precision, deterministic) bind(C, name='fnetgpu_mg_init')
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: mg
      integer(c_int), value :: nDevices, precision, deterministic
    end function
    integer(c_int) function fnetgpu_mg_finalize(mg) bind(C, name='fnetgpu_mg_finalize')
      import :: c_ptr, c_int
      type(c_ptr), value :: mg
    end function
    type(c_ptr) function fnetgpu_mg_last_error(mg) bind(C, name='fnetgpu_mg_last_error')
      import :: c_ptr
      type(c_ptr), value :: mg
    end function
    integer(c_int) function fnetgpu_mg_dataset_upload(mg, slot, nStruct, offsets, coords, periodic, latvecs,&
        & atnum, globalsp, dsWeights, atomicWeights, nG, gTargets, nA, aTargets, nExt, ext)&
        & bind(C, name='fnetgpu_mg_dataset_upload')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: slot, nStruct, nG, nA, nExt
      integer(c_int), intent(in) :: offsets(*), periodic(*), atnum(*), globalsp(*), dsWeights(*)
      real(c_double), intent(in) :: coords(*), latvecs(*), atomicWeights(*), gTargets(*), aTargets(*), ext(*)
    end function
    integer(c_int) function fnetgpu_mg_acsf_set(mg, nFunc, ftype, rcut, kappa, rs, eta, lambda, xi, atomid,&
        & atomicnumbers) bind(C, name='fnetgpu_mg_acsf_set')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: nFunc
      integer(c_int), intent(in) :: ftype(*), atomid(*), atomicnumbers(*)
      real(c_double), intent(in) :: rcut(*), kappa(*), rs(*), eta(*), lambda(*), xi(*)
    end function
    integer(c_int) function fnetgpu_mg_features_config(mg, nExtSel, extIndices) bind(C, name='fnetgpu_mg_features_config')
      import :: c_ptr, c_int
      type(c_ptr), value :: mg
      integer(c_int), value :: nExtSel
      integer(c_int), intent(in) :: extIndices(*)
    end function
    integer(c_int) function fnetgpu_mg_acsf_calculate(mg, slot, standardize, zprec, have_zprec)&
        & bind(C, name='fnetgpu_mg_acsf_calculate')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: slot, standardize, have_zprec
      real(c_double), intent(inout) :: zprec(*)
    end function
    integer(c_int) function fnetgpu_mg_net_set(mg, nSpecies, nLayers, dims, activationId) bind(C, name='fnetgpu_mg_net_set')
      import :: c_ptr, c_int
      type(c_ptr), value :: mg
      integer(c_int), value :: nSpecies, nLayers, activationId
      integer(c_int), intent(in) :: dims(*)
    end function
    integer(c_int) function fnetgpu_mg_params_set(mg, wb) bind(C, name='fnetgpu_mg_params_set')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      real(c_double), intent(in) :: wb(*)
    end function
    integer(c_int) function fnetgpu_mg_grad(mg, slot, lossId, shuffle, ddSerial, loss, globalPred) bind(C, name='fnetgpu_mg_grad')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg, shuffle, globalPred
      integer(c_int), value :: slot, lossId
      real(c_double), intent(out) :: ddSerial(*), loss
    end function
    integer(c_int) function fnetgpu_mg_loss(mg, slot, lossId, loss) bind(C, name='fnetgpu_mg_loss')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: slot, lossId
      real(c_double), intent(out) :: loss
    end function
    integer(c_int) function fnetgpu_mg_predict(mg, slot, raw) bind(C, name='fnetgpu_mg_predict')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: slot
      real(c_double), intent(out) :: raw(*)
    end function
    integer(c_int) function fnetgpu_mg_forces(mg, slot, forces) bind(C, name='fnetgpu_mg_forces')
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: mg
      integer(c_int), value :: slot
      real(c_double), intent(out) :: forces(*)
    end function
    type(c_ptr) function fnetgpu_mg_context(mg, device) bind(C, name='fnetgpu_mg_context')
      import :: c_ptr, c_int
      type(c_ptr), value :: mg
      integer(c_int), value :: device
    end function
  end interface

contains

  subroutine check(env, iErr)
    type(TGpuEnv), intent(in) :: env
    integer(c_int), intent(in) :: iErr
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: buf
    integer :: ii
    if (iErr == 0) return
    if (c_associated(env%mg)) then
      call c_f_pointer(fnetgpu_mg_last_error(env%mg), msg, [512])
    else
      call c_f_pointer(fnetgpu_last_error(env%ctx), msg, [512])
    end if
    buf = ''
    do ii = 1, 512
      if (msg(ii) == c_null_char) exit
      buf(ii:ii) = msg(ii)
    end do
    call error('fnetgpu: ' // trim(buf))   ! abort-on-error like every other failure in Fortnet
  end subroutine check

  !> replaces TEnv_init for the hot path (device < 0: $LOCAL_RANK or 0; precision 64 | 32)
  !! nDevices (Options { NDevices }, default 1; 0 = all visible): > 1 selects the single-process multi-GPU layer
  subroutine TGpuEnv_init(env, device, precision, nDevices)
    type(TGpuEnv), intent(out) :: env
    integer, intent(in) :: device, precision
    integer, intent(in), optional :: nDevices
    env%nDevices = 1
    if (present(nDevices)) env%nDevices = nDevices
    if (env%nDevices == 1) then
      call check(env, fnetgpu_init(env%ctx, int(device, c_int), int(precision, c_int), 1_c_int))
    else
      call check(env, fnetgpu_mg_init(env%mg, int(env%nDevices, c_int), int(precision, c_int), 1_c_int))
      env%ctx = fnetgpu_mg_context(env%mg, 0_c_int)     ! queries only (fnetgpu_ntot, ...)
    end if
  end subroutine TGpuEnv_init

  !> MPI build (WITH_MPI = TRUE): one rank per GPU.  The lead rank creates the NCCL id, mpifx_bcast ships its
  !! 128 bytes, every rank joins; fnetgpu_grad / fnetgpu_loss / the statistics pass then return global sums and
  !! the driver's own mpifx_allreduce calls of the hot path (lib_nn/bpnn.F90:455-467) are dropped.
  !!   call TGpuEnv_initRank(gpu, env%globalMpiComm%rank, env%globalMpiComm%size, precision, id)
  !! with  if (lead) call gpuCommUniqueId(id);  call mpifx_bcast(env%globalMpiComm, id)  before it.
  subroutine TGpuEnv_initRank(env, rank, nRanks, precision, id)
    type(TGpuEnv), intent(out) :: env
    integer, intent(in) :: rank, nRanks, precision
    character(kind=c_char), intent(in) :: id(128)
    env%nDevices = 1
    call check(env, fnetgpu_init(env%ctx, -1_c_int, int(precision, c_int), 1_c_int))   ! device = $LOCAL_RANK
    call check(env, fnetgpu_comm_init(env%ctx, int(nRanks, c_int), int(rank, c_int), id))
  end subroutine TGpuEnv_initRank

  subroutine gpuCommUniqueId(id)
    character(kind=c_char), intent(out) :: id(128)
    if (fnetgpu_comm_unique_id(id) /= 0) call error('fnetgpu: cannot create the NCCL unique id')
  end subroutine gpuCommUniqueId

  subroutine TGpuEnv_final(env)
    type(TGpuEnv), intent(inout) :: env
    integer(c_int) :: iErr
    if (c_associated(env%mg)) then
      iErr = fnetgpu_mg_finalize(env%mg)
    else
      iErr = fnetgpu_finalize(env%ctx)
    end if
    env%ctx = c_null_ptr
    env%mg = c_null_ptr
  end subroutine TGpuEnv_final

  !> Flattens a TDataset (ragged per-structure arrays) and ships it to the GPU once.
  !! Called where calculateMappings receives the dataset (prg_fnet/fortnet.F90:883).
  subroutine gpuUploadDataset(env, slot, offsets, coords, periodic, latVecs, atNum, globalSp, weights,&
      & atomicWeights, nGlobal, globalTargets, nAtomic, atomicTargets, nExt, extFeatures)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot, offsets(:), periodic(:), atNum(:), globalSp(:), weights(:)
    real(dp), intent(in) :: coords(:,:), latVecs(:,:,:), atomicWeights(:)
    integer, intent(in) :: nGlobal, nAtomic, nExt
    real(dp), intent(in) :: globalTargets(:,:), atomicTargets(:,:), extFeatures(:,:)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_dataset_upload(env%mg, int(slot, c_int), int(size(offsets) - 1, c_int), offsets, coords,&
          & periodic, latVecs, atNum, globalSp, weights, atomicWeights, int(nGlobal, c_int), globalTargets,&
          & int(nAtomic, c_int), atomicTargets, int(nExt, c_int), extFeatures))
    else
      call check(env, fnetgpu_dataset_upload(env%ctx, int(slot, c_int), int(size(offsets) - 1, c_int), offsets, coords,&
          & periodic, latVecs, atNum, globalSp, weights, atomicWeights, int(nGlobal, c_int), globalTargets,&
          & int(nAtomic, c_int), atomicTargets, int(nExt, c_int), extFeatures))
    end if
  end subroutine gpuUploadDataset

  !> TAcsf_init: passes this%gFunctions%func(:) as flat tables (type 'g1'..'g5' -> 1..5)
  subroutine gpuAcsfSet(env, ftype, rcut, kappa, rs, eta, lambda, xi, atomId, atomicNumbers)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: ftype(:), atomId(:), atomicNumbers(:,:)
    real(dp), intent(in) :: rcut(:), kappa(:), rs(:), eta(:), lambda(:), xi(:)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_acsf_set(env%mg, int(size(ftype), c_int), ftype, rcut, kappa, rs, eta, lambda, xi,&
          & atomId, atomicNumbers))
    else
      call check(env, fnetgpu_acsf_set(env%ctx, int(size(ftype), c_int), ftype, rcut, kappa, rs, eta, lambda, xi,&
          & atomId, atomicNumbers))
    end if
  end subroutine gpuAcsfSet

  !> TAcsf%calculate: zPrec(F,2) is this%zPrec; tHave = allocated(this%zPrec) on entry
  subroutine gpuAcsfCalculate(env, slot, tZscore, zPrec, tHave)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    logical, intent(in) :: tZscore, tHave
    real(dp), intent(inout) :: zPrec(:,:)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_acsf_calculate(env%mg, int(slot, c_int), merge(1_c_int, 0_c_int, tZscore), zPrec,&
          & merge(1_c_int, 0_c_int, tHave)))
    else
      call check(env, fnetgpu_acsf_calculate(env%ctx, int(slot, c_int), merge(1_c_int, 0_c_int, tZscore), zPrec,&
          & merge(1_c_int, 0_c_int, tHave)))
    end if
  end subroutine gpuAcsfCalculate

  subroutine gpuNetSet(env, nSpecies, dims, activationId)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: nSpecies, dims(:), activationId
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_net_set(env%mg, int(nSpecies, c_int), int(size(dims), c_int), dims, int(activationId, c_int)))
    else
      call check(env, fnetgpu_net_set(env%ctx, int(nSpecies, c_int), int(size(dims), c_int), dims, int(activationId, c_int)))
    end if
  end subroutine gpuNetSet

  !> after TBpnn%serializedWeightsAndBiases(weightsAndBiases) (bpnn.F90:801-822)
  subroutine gpuParamsSet(env, weightsAndBiases)
    type(TGpuEnv), intent(in) :: env
    real(dp), intent(in) :: weightsAndBiases(:,:)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_params_set(env%mg, weightsAndBiases))
    else
      call check(env, fnetgpu_params_set(env%ctx, weightsAndBiases))
    end if
  end subroutine gpuParamsSet

  !> replaces the body of TBpnn_updateGradients + the loss(...) call (bpnn.F90:277-283, 317-323);
  !! ddSerial(nTot, nSpecies) is then un-serialised into resDd (inverse of TDerivs_serialized).
  subroutine gpuUpdateGradients(env, slot, lossId, ddSerial, loss)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot, lossId
    real(dp), intent(out) :: ddSerial(:,:), loss
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_grad(env%mg, int(slot, c_int), int(lossId, c_int), c_null_ptr, ddSerial, loss, c_null_ptr))
    else
      call check(env, fnetgpu_grad(env%ctx, int(slot, c_int), int(lossId, c_int), c_null_ptr, ddSerial, loss, c_null_ptr))
    end if
  end subroutine gpuUpdateGradients

  subroutine gpuPredictBatch(env, slot, predicts)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(out) :: predicts(:,:)   ! (nOut, nTotAtoms), split per structure by the caller
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_predict(env%mg, int(slot, c_int), predicts))
    else
      call check(env, fnetgpu_predict(env%ctx, int(slot, c_int), predicts))
    end if
  end subroutine gpuPredictBatch

  subroutine gpuForces(env, slot, forces)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(out) :: forces(:,:)     ! (3*nOut, nTotAtoms)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_forces(env%mg, int(slot, c_int), forces))
    else
      call check(env, fnetgpu_forces(env%ctx, int(slot, c_int), forces))
    end if
  end subroutine gpuForces

  !> One i-PI / MD step: replaces calculateMappingsForSocketComm + predictForSocketComm
  !! (prg_fnet/fortnet.F90:430-609) for the geometry resident in slot.
  subroutine gpuSocketStep(env, slot, coords, latVecs, globalPrediction, atomicPredictions, atomicForces)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(in) :: coords(:,:)                 ! (3, nAtom)
    real(dp), intent(in), target :: latVecs(:,:)        ! (3, 3)
    real(dp), intent(out) :: globalPrediction(:)        ! (nOut)
    real(dp), intent(out) :: atomicPredictions(:,:)     ! (nOut, nAtom)
    real(dp), intent(out) :: atomicForces(:,:)          ! (3*nOut, nAtom)
    call check(env, fnetgpu_socket_step(env%ctx, int(slot, c_int), coords, c_loc(latVecs), globalPrediction,&
        & atomicPredictions, atomicForces))
  end subroutine gpuSocketStep

  !> new geometry for a resident dataset (finite-difference forces, active-learning loops): latVecs optional
  subroutine gpuCoordsUpdate(env, slot, coords, latVecs)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(in) :: coords(:,:)
    real(dp), intent(in), target, optional :: latVecs(:,:,:)
    type(c_ptr) :: pLat
    pLat = c_null_ptr
    if (present(latVecs)) pLat = c_loc(latVecs)
    call check(env, fnetgpu_coords_update(env%ctx, int(slot, c_int), coords, pLat))
  end subroutine gpuCoordsUpdate

  !> TFeatures_collect (lib_types/features.F90:200-265): 1-based rows of extFeatures appended behind the ACSF
  subroutine gpuFeaturesConfig(env, extIndices)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: extIndices(:)
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_features_config(env%mg, int(size(extIndices), c_int), extIndices))
    else
      call check(env, fnetgpu_features_config(env%ctx, int(size(extIndices), c_int), extIndices))
    end if
  end subroutine gpuFeaturesConfig

  !> TBpnn_update's gradient post-processing on the device (bpnn.F90:750-767): elastic-net term + / sum(weights).
  !! With it, TBpnn_update skips elasticNetRegularization and the division and goes straight to norm2 / next().
  subroutine gpuSetRegularization(env, strength, alpha, nDatapoints)
    type(TGpuEnv), intent(in) :: env
    real(dp), intent(in) :: strength, alpha, nDatapoints
    integer :: iDev
    if (c_associated(env%mg)) then
      do iDev = 0, env%nDevices - 1
        call check(env, fnetgpu_regularization_set(fnetgpu_mg_context(env%mg, int(iDev, c_int)), strength, alpha, nDatapoints))
      end do
    else
      call check(env, fnetgpu_regularization_set(env%ctx, strength, alpha, nDatapoints))
    end if
  end subroutine gpuSetRegularization

  !> validation-loss monitoring (bpnn.F90:284-290, 324-330): predictBatch + loss without the per-atom download
  subroutine gpuLoss(env, slot, lossId, loss)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot, lossId
    real(dp), intent(out) :: loss
    if (c_associated(env%mg)) then
      call check(env, fnetgpu_mg_loss(env%mg, int(slot, c_int), int(lossId, c_int), loss))
    else
      call check(env, fnetgpu_loss(env%ctx, int(slot, c_int), int(lossId, c_int), loss))
    end if
  end subroutine gpuLoss

  !> inverse of TDerivs_serialized (lib_common/nestedtypes.F90:428-470): per species all weight arrays
  !! dw(1..nArrays) column by column (including the dummy dw(nArrays) of shape (d_L, 1)), then all bias arrays
  !! db(1..nArrays) (including the unused db(1)).  dd must be initialised (TDerivs_init).
  subroutine gpuUnserialize(ddSerial, dd)
    use fnet_nestedtypes, only : TDerivs
    real(dp), intent(in) :: ddSerial(:,:)
    type(TDerivs), intent(inout) :: dd
    integer :: iStruc, iLayer, ii, jj, ind
    do iStruc = 1, size(dd%db)
      ind = 1
      do iLayer = 1, size(dd%db(iStruc)%db)
        do ii = 1, size(dd%dw(iStruc)%dw(iLayer)%array, dim=2)
          do jj = 1, size(dd%dw(iStruc)%dw(iLayer)%array, dim=1)
            dd%dw(iStruc)%dw(iLayer)%array(jj, ii) = ddSerial(ind, iStruc)
            ind = ind + 1
          end do
        end do
      end do
      do iLayer = 1, size(dd%db(iStruc)%db)
        do ii = 1, size(dd%db(iStruc)%db(iLayer)%array)
          dd%db(iStruc)%db(iLayer)%array(ii) = ddSerial(ind, iStruc)
          ind = ind + 1
        end do
      end do
    end do
  end subroutine gpuUnserialize

end module fnet_gpu
