! example_emgmm_b200.f90 -- what switching the reference's EM test program (tests/test_emgmm.f90, testGMM) to the B200
! library looks like: the step-object calls are unchanged, only the type name and the data generation differ.
!   gfortran ... mlf_emgmm_b200.f90 example_emgmm_b200.f90 -lmlfortran -lmlfortran_b200
! (Not compiled in the authoring image: no Fortran compiler.  Usage: example_emgmm_b200 ND NC NX Nstep sigmaC sigmaX)
Program example_emgmm_b200
  Use iso_c_binding
  Use mlf_intf
  Use mlf_hdf5
  Use mlf_emgmm_b200
  IMPLICIT NONE
  integer :: ND, NC, NX, Nstep, info, k
  integer(kind=8) :: niter
  real(c_double) :: sigmaC, sigmaX, dt
  real(c_double), allocatable :: X(:,:), XC(:,:), C12(:,:,:)
  type(mlf_algo_emgmm_b200) :: em
  type(mlf_hdf5_file) :: h5
  character(len=32) :: arg

  CALL GET_COMMAND_ARGUMENT(1, arg); READ(arg, *) ND
  CALL GET_COMMAND_ARGUMENT(2, arg); READ(arg, *) NC
  CALL GET_COMMAND_ARGUMENT(3, arg); READ(arg, *) NX
  CALL GET_COMMAND_ARGUMENT(4, arg); READ(arg, *) Nstep
  CALL GET_COMMAND_ARGUMENT(5, arg); READ(arg, *) sigmaC
  CALL GET_COMMAND_ARGUMENT(6, arg); READ(arg, *) sigmaX
  info = mlf_init()
  ALLOCATE(X(ND, NX*NC), XC(ND, NC), C12(ND, ND, NC))

  ! the mixture of the test program, generated on the GPU (seeded)
  info = mlf_b200_synth_gmm_f(X, XC, C12, sigmaC, sigmaX, 20240601_c_int64_t)
  If(info < 0) STOP "generator failed"

  mlf_b200_use_all_gpus = .TRUE.          ! shard the points over every GPU of the box
  info = em%init(X, NC, 1d0)              ! as tests/test_emgmm.f90:125 (minProba = 1)
  If(info < 0) STOP "init failed"
  niter = Nstep
  info = em%step(dt, niter)               ! first iteration: k-means initialiser; then Nstep-1 EM iterations on the GPUs
  Print *, "Time: ", dt, " sumLL: ", em%sumLL
  Do k = 1, NC
    Print *, "component", k, " weight ", em%lambda(k), " mean ", em%Mu(:,k)
  End Do

  info = h5%createFile("emgmm.h5")        ! the reference's checkpoint container, unchanged
  info = h5%pushState(em)
  CALL h5%finalize()
  CALL em%finalize()
  info = mlf_quit()
End Program example_emgmm_b200
