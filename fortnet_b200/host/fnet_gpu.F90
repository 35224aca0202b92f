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
  public :: gpuUploadDataset, gpuAcsfSet, gpuAcsfCalculate, gpuNetSet, gpuParamsSet
  public :: gpuUpdateGradients, gpuPredictBatch, gpuForces

  type :: TGpuEnv
    type(c_ptr) :: ctx = c_null_ptr
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
  end interface

contains

  subroutine check(env, iErr)
    type(TGpuEnv), intent(in) :: env
    integer(c_int), intent(in) :: iErr
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: buf
    integer :: ii
    if (iErr == 0) return
    call c_f_pointer(fnetgpu_last_error(env%ctx), msg, [512])
    buf = ''
    do ii = 1, 512
      if (msg(ii) == c_null_char) exit
      buf(ii:ii) = msg(ii)
    end do
    call error('fnetgpu: ' // trim(buf))   ! abort-on-error like every other failure in Fortnet
  end subroutine check

  !> replaces TEnv_init for the hot path (device < 0: $LOCAL_RANK or 0; precision 64 | 32)
  subroutine TGpuEnv_init(env, device, precision)
    type(TGpuEnv), intent(out) :: env
    integer, intent(in) :: device, precision
    call check(env, fnetgpu_init(env%ctx, int(device, c_int), int(precision, c_int), 1_c_int))
  end subroutine TGpuEnv_init

  subroutine TGpuEnv_final(env)
    type(TGpuEnv), intent(inout) :: env
    integer(c_int) :: iErr
    iErr = fnetgpu_finalize(env%ctx)
    env%ctx = c_null_ptr
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
    call check(env, fnetgpu_dataset_upload(env%ctx, int(slot, c_int), int(size(offsets) - 1, c_int), offsets, coords,&
        & periodic, latVecs, atNum, globalSp, weights, atomicWeights, int(nGlobal, c_int), globalTargets,&
        & int(nAtomic, c_int), atomicTargets, int(nExt, c_int), extFeatures))
  end subroutine gpuUploadDataset

  !> TAcsf_init: passes this%gFunctions%func(:) as flat tables (type 'g1'..'g5' -> 1..5)
  subroutine gpuAcsfSet(env, ftype, rcut, kappa, rs, eta, lambda, xi, atomId, atomicNumbers)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: ftype(:), atomId(:), atomicNumbers(:,:)
    real(dp), intent(in) :: rcut(:), kappa(:), rs(:), eta(:), lambda(:), xi(:)
    call check(env, fnetgpu_acsf_set(env%ctx, int(size(ftype), c_int), ftype, rcut, kappa, rs, eta, lambda, xi,&
        & atomId, atomicNumbers))
  end subroutine gpuAcsfSet

  !> TAcsf%calculate: zPrec(F,2) is this%zPrec; tHave = allocated(this%zPrec) on entry
  subroutine gpuAcsfCalculate(env, slot, tZscore, zPrec, tHave)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    logical, intent(in) :: tZscore, tHave
    real(dp), intent(inout) :: zPrec(:,:)
    call check(env, fnetgpu_acsf_calculate(env%ctx, int(slot, c_int), merge(1_c_int, 0_c_int, tZscore), zPrec,&
        & merge(1_c_int, 0_c_int, tHave)))
  end subroutine gpuAcsfCalculate

  subroutine gpuNetSet(env, nSpecies, dims, activationId)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: nSpecies, dims(:), activationId
    call check(env, fnetgpu_net_set(env%ctx, int(nSpecies, c_int), int(size(dims), c_int), dims, int(activationId, c_int)))
  end subroutine gpuNetSet

  !> after TBpnn%serializedWeightsAndBiases(weightsAndBiases) (bpnn.F90:801-822)
  subroutine gpuParamsSet(env, weightsAndBiases)
    type(TGpuEnv), intent(in) :: env
    real(dp), intent(in) :: weightsAndBiases(:,:)
    call check(env, fnetgpu_params_set(env%ctx, weightsAndBiases))
  end subroutine gpuParamsSet

  !> replaces the body of TBpnn_updateGradients + the loss(...) call (bpnn.F90:277-283, 317-323);
  !! ddSerial(nTot, nSpecies) is then un-serialised into resDd (inverse of TDerivs_serialized).
  subroutine gpuUpdateGradients(env, slot, lossId, ddSerial, loss)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot, lossId
    real(dp), intent(out) :: ddSerial(:,:), loss
    call check(env, fnetgpu_grad(env%ctx, int(slot, c_int), int(lossId, c_int), c_null_ptr, ddSerial, loss, c_null_ptr))
  end subroutine gpuUpdateGradients

  subroutine gpuPredictBatch(env, slot, predicts)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(out) :: predicts(:,:)   ! (nOut, nTotAtoms), split per structure by the caller
    call check(env, fnetgpu_predict(env%ctx, int(slot, c_int), predicts))
  end subroutine gpuPredictBatch

  subroutine gpuForces(env, slot, forces)
    type(TGpuEnv), intent(in) :: env
    integer, intent(in) :: slot
    real(dp), intent(out) :: forces(:,:)     ! (3*nOut, nTotAtoms)
    call check(env, fnetgpu_forces(env%ctx, int(slot, c_int), forces))
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

end module fnet_gpu
