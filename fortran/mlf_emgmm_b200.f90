! mlf_emgmm_b200.f90 -- Fortran side of the B200 drop-in for MlFortran's Gaussian-mixture EM step object.
!
! Deliverable form (A) of INTEGRATION.md: compile this file next to the reference's sources and link
! libmlfortran_b200.so.  `mlf_algo_emgmm_b200` extends the reference's own `mlf_step_obj`, registers the same seven
! resource slots (iPar rPar iVar rVar Mu Cov lambda) through the reference's resource runtime, and forwards the
! arithmetic of `stepF` (ExpStep + MaxStep, src/unsupervised/mlf_emgmm.f90:125-207) to the CUDA engine through the
! iso_c_binding interfaces below (include/mlf_b200.h, "engine-level entry points").  Because the object *is* an
! mlf_step_obj, everything generic in the reference keeps working on it unchanged: this%step(dt, niter), mlf_step,
! mlf_getrsc / mlf_updatersc, mlf_hdf5_file%pushState and restore through data_handler, constraints_step, finalize.
!
! Interface parity with mlf_algo_emgmm (src/unsupervised/mlf_emgmm.f90:46-56, 76-153):
!   info = obj%init(X, nC, minProba, data_handler)    same arguments, same meaning; X is borrowed AND copied to HBM
!   info = obj%reinit()                               zeroes Mu/Cov, initialized = .FALSE.
!   info = obj%stepF(niter)                           first iteration of a fresh object = k-means initialiser
!   public components Mu, Cov, lambda, sumLL, minProba, initialized; ProbaC through obj%getProbaC()
!
! This file cannot be compiled in the authoring image (no Fortran compiler); it is kept syntactically plain
! Fortran 2008 and is exercised structurally by tests/test_fortran_shim_cpu.py (every bind(C) name below must be
! exported by the library with a matching argument count).
Module mlf_emgmm_b200
  Use iso_c_binding
  Use mlf_intf
  Use mlf_errors
  Use mlf_rsc_array
  Use mlf_step_algo
  Use mlf_models
  Use mlf_utils
  IMPLICIT NONE
  PRIVATE

  Integer(c_int), Parameter :: nkm_default = 16   ! hard-coded in the reference (src/unsupervised/mlf_emgmm.f90:131)

  ! .TRUE.: objects initialised from now on shard their points over every visible GPU inside this process (peer-memory
  ! all-reduce, mlf_b200_engine_create_multi); .FALSE.: the current CUDA device only.
  Logical, Public, Save :: mlf_b200_use_all_gpus = .FALSE.
  Public :: mlf_b200_synth_gmm_f, mlf_b200_nccl_unique_id_f

  Type, Public, Extends(mlf_step_obj) :: mlf_algo_emgmm_b200
    real(c_double), pointer :: X(:,:), Mu(:,:), Cov(:,:,:), lambda(:)
    real(c_double), pointer :: sumLL, minProba
    logical :: initialized = .FALSE.
    integer(c_int) :: device = -1                                   ! -1: current CUDA device
    integer(c_int64_t) :: seed = int(z'1E3779B97F4A7C15', c_int64_t) ! k-means stream (reference: unseeded RANDOM_NUMBER)
    type(c_ptr) :: engine = C_NULL_PTR
  Contains
    procedure :: init => emgmm_b200_init
    procedure :: reinit => emgmm_b200_reinit
    procedure :: stepF => emgmm_b200_stepF
    procedure :: getProbaC => emgmm_b200_getProbaC
    procedure :: joinNccl => emgmm_b200_joinNccl
    procedure :: finalize => emgmm_b200_finalize
  End Type mlf_algo_emgmm_b200

  Type, Public, Extends(mlf_class_proba_model) :: mlf_model_emgmm_b200
    class(mlf_algo_emgmm_b200), pointer :: top
  Contains
    procedure :: getProba => emgmm_b200_getLikelihood
    procedure :: getNumClasses => emgmm_b200_getNumClasses
  End Type mlf_model_emgmm_b200

  Interface
    Function c_engine_create(X, nX, nY, nC, device, flags) Result(eng) bind(C, name="mlf_b200_engine_create")
      Import :: c_ptr, c_double, c_int64_t, c_int
      real(c_double), intent(in) :: X(*)
      integer(c_int64_t), value :: nX
      integer(c_int), value :: nY, nC, device, flags
      type(c_ptr) :: eng
    End Function c_engine_create
    Function c_engine_create_multi(X, nX, nY, nC, devices, ndevices, flags) Result(eng) bind(C, name="mlf_b200_engine_create_multi")
      Import :: c_ptr, c_double, c_int64_t, c_int
      real(c_double), intent(in) :: X(*)
      integer(c_int64_t), value :: nX
      integer(c_int), value :: nY, nC, ndevices, flags
      type(c_ptr), value :: devices        ! C_NULL_PTR with ndevices = 0: every visible device
      type(c_ptr) :: eng
    End Function c_engine_create_multi
    Subroutine c_engine_destroy(eng) bind(C, name="mlf_b200_engine_destroy")
      Import :: c_ptr
      type(c_ptr), value :: eng
    End Subroutine c_engine_destroy
    Integer(c_int) Function c_engine_step(eng, niter, minProba, Mu, Cov, lambda, sumLL) bind(C, name="mlf_b200_engine_step")
      Import :: c_ptr, c_double, c_int64_t, c_int
      type(c_ptr), value :: eng
      integer(c_int64_t), value :: niter
      real(c_double), value :: minProba
      real(c_double), intent(inout) :: Mu(*), Cov(*), lambda(*)
      real(c_double), intent(out) :: sumLL
    End Function c_engine_step
    Integer(c_int) Function c_engine_kmeans_init(eng, nkm, seed, Mu, Cov, lambda) bind(C, name="mlf_b200_engine_kmeans_init")
      Import :: c_ptr, c_double, c_int64_t, c_int
      type(c_ptr), value :: eng
      integer(c_int), value :: nkm
      integer(c_int64_t), value :: seed
      real(c_double), intent(out) :: Mu(*), Cov(*), lambda(*)
    End Function c_engine_kmeans_init
    Integer(c_int) Function c_engine_expectation(eng, Mu, Cov, lambda, Xq, nIn, P, Cl) bind(C, name="mlf_b200_engine_expectation")
      Import :: c_ptr, c_double, c_int64_t, c_int
      type(c_ptr), value :: eng
      real(c_double), intent(in) :: Mu(*), Cov(*), lambda(*), Xq(*)
      integer(c_int64_t), value :: nIn
      type(c_ptr), value :: P, Cl        ! nullable outputs: c_loc(array) or C_NULL_PTR
    End Function c_engine_expectation
    Integer(c_int) Function c_engine_proba(eng, ProbaC) bind(C, name="mlf_b200_engine_proba")
      Import :: c_ptr, c_double, c_int
      type(c_ptr), value :: eng
      real(c_double), intent(out) :: ProbaC(*)
    End Function c_engine_proba
    ! One process per GPU (MPI): rank 0 draws the 128-byte ncclUniqueId, the caller broadcasts it (MPI_Bcast), every rank joins
    ! before its first step.  The statistics all-reduce of every iteration is then an ncclAllReduce issued by the library.
    Integer(c_int) Function c_nccl_unique_id(id) bind(C, name="mlf_b200_nccl_unique_id")
      Import :: c_int, c_signed_char
      integer(c_signed_char), intent(out) :: id(128)
    End Function c_nccl_unique_id
    Integer(c_int) Function c_engine_nccl_init(eng, id, rank, world) bind(C, name="mlf_b200_engine_nccl_init")
      Import :: c_ptr, c_int, c_signed_char
      type(c_ptr), value :: eng
      integer(c_signed_char), intent(in) :: id(128)
      integer(c_int), value :: rank, world
    End Function c_engine_nccl_init
    ! Workload generator of tests/test_emgmm.f90:105-124 (RandN centres, randOrthogonal(C12, weight), RandN(X, X0, C12), randPerm)
    ! on the device: points n0+1 .. n0+nX of an nTotal-point job into X(nY, nX) (host array with flags = 0); XC(nY, nC) and
    ! C12(nY, nY, nC) come back as the test program's own arrays would.  Replaces the generation loop of testGMM at scale.
    Integer(c_int) Function c_synth_gmm(X, n0, nX, nTotal, nY, nC, sigmaC, sigmaX, seed, device, flags, XC, C12) &
        bind(C, name="mlf_b200_synth_gmm")
      Import :: c_double, c_int64_t, c_int
      real(c_double), intent(out) :: X(*), XC(*), C12(*)
      integer(c_int64_t), value :: n0, nX, nTotal, seed
      integer(c_int), value :: nY, nC, device, flags
      real(c_double), value :: sigmaC, sigmaX
    End Function c_synth_gmm
  End Interface

Contains

  ! Same slot layout as mlf_algo_emgmm_init (src/unsupervised/mlf_emgmm.f90:76-112); the only additions are the
  ! engine handle and the host->HBM copy of X made by mlf_b200_engine_create.
  Integer Function emgmm_b200_init(this, X, nC, minProba, data_handler) Result(info)
    class(mlf_algo_emgmm_b200), intent(out), target :: this
    class(mlf_data_handler), intent(inout), optional :: data_handler
    real(c_double), optional :: minProba
    real(c_double), target :: X(:,:)
    integer, intent(inout) :: nC
    type(mlf_step_numFields) :: nf
    integer(c_int64_t) :: shapeMu(2), shapeCov(3), nComp
    CALL nf%initFields(nRVar = 1, nRsc = 3, nIVar = 1)
    info = mlf_step_obj_init(this, nf, data_handler = data_handler)
    this%X => X
    If(info < 0) RETURN
    shapeMu = [INT(SIZE(X,1), c_int64_t), INT(nC, c_int64_t)]
    info = this%add_rmatrix(nf, shapeMu, this%Mu, C_CHAR_"Mu", data_handler = data_handler, &
      fixed_dims = [.TRUE., .FALSE.])                        ! K may come from the checkpoint
    If(CheckF(info, "emgmm_b200: error creating Mu")) RETURN
    shapeCov = [shapeMu(1), shapeMu(1), shapeMu(2)]
    info = this%add_rmatrix3d(nf, shapeCov, this%Cov, C_CHAR_"Cov", data_handler = data_handler, &
      fixed_dims = [.TRUE., .TRUE., .TRUE.])
    If(CheckF(info, "emgmm_b200: error creating Cov")) RETURN
    nComp = shapeMu(2)
    info = this%add_rarray(nf, nComp, this%lambda, C_CHAR_"lambda", data_handler = data_handler, &
      fixed_dims = [.TRUE.])
    If(CheckF(info, "emgmm_b200: error creating lambda")) RETURN
    nC = INT(nComp, KIND=4)
    CALL this%addRVar(nf, this%sumLL, "sumLL")
    CALL this%addRPar(nf, this%minProba, "minProba")
    CALL InitOrDefault(this%minProba, 0.0d0, minProba)
    If(mlf_b200_use_all_gpus) Then
      this%engine = c_engine_create_multi(X, INT(SIZE(X,2), c_int64_t), INT(SIZE(X,1), c_int), INT(nC, c_int), &
        C_NULL_PTR, 0_c_int, 0_c_int)
    Else
      this%engine = c_engine_create(X, INT(SIZE(X,2), c_int64_t), INT(SIZE(X,1), c_int), INT(nC, c_int), this%device, 0_c_int)
    Endif
    If(.NOT. C_ASSOCIATED(this%engine)) Then
      info = mlf_OTHERERROR                                   ! no GPU / no library: there is no CPU fallback
      RETURN
    Endif
    If(PRESENT(data_handler)) Then
      this%initialized = .TRUE.
    Else
      info = this%reinit()
    Endif
    CALL emgmm_b200_model_init(this%model, this)
  End Function emgmm_b200_init

  Subroutine emgmm_b200_model_init(model, top)
    class(mlf_model), intent(out), allocatable :: model
    class(mlf_algo_emgmm_b200), intent(in), target :: top
    ALLOCATE(mlf_model_emgmm_b200 :: model)
    Select Type(model)
    Class is (mlf_model_emgmm_b200)
      model%top => top
    End Select
  End Subroutine emgmm_b200_model_init

  Integer Function emgmm_b200_reinit(this) Result(info)
    class(mlf_algo_emgmm_b200), intent(inout), target :: this
    info = mlf_step_obj_reinit(this)
    If(ASSOCIATED(this%Mu)) this%Mu = 0
    If(ASSOCIATED(this%Cov)) this%Cov = 0
    this%initialized = .FALSE.
  End Function emgmm_b200_reinit

  ! mlf_algo_emgmm_stepF (src/unsupervised/mlf_emgmm.f90:125-153): the first iteration of a fresh object is the
  ! k-means initialiser; every other iteration is ExpStep + MaxStep, all of them run by one engine call.
  Integer Function emgmm_b200_stepF(this, niter) Result(info)
    class(mlf_algo_emgmm_b200), intent(inout), target :: this
    integer(kind=8), intent(inout), optional :: niter
    integer(c_int64_t) :: nEM
    nEM = 1
    If(PRESENT(niter)) nEM = niter
    info = 0
    If(nEM <= 0) RETURN
    If(.NOT. this%initialized) Then
      info = c_engine_kmeans_init(this%engine, nkm_default, this%seed, this%Mu, this%Cov, this%lambda)
      If(info < 0) RETURN
      this%initialized = .TRUE.
      nEM = nEM - 1
    Endif
    If(nEM > 0) info = c_engine_step(this%engine, nEM, this%minProba, this%Mu, this%Cov, this%lambda, this%sumLL)
  End Function emgmm_b200_stepF

  ! The reference exposes ProbaC(nX,nC) as a public allocatable; here it lives in HBM and is copied out on demand.
  Integer Function emgmm_b200_getProbaC(this, ProbaC) Result(info)
    class(mlf_algo_emgmm_b200), intent(inout), target :: this
    real(c_double), intent(out) :: ProbaC(:,:)
    info = c_engine_proba(this%engine, ProbaC)
  End Function emgmm_b200_getProbaC

  ! Sharded use, one MPI rank per GPU: X passed to init is this rank's contiguous shard of the points.
  !   If(rank == 0) info = mlf_b200_nccl_unique_id_f(id);  CALL MPI_Bcast(id, 128, MPI_BYTE, 0, comm, ierr);  info = em%joinNccl(id, rank, world)
  Integer Function mlf_b200_nccl_unique_id_f(id) Result(info)
    integer(c_signed_char), intent(out) :: id(128)
    info = c_nccl_unique_id(id)
  End Function mlf_b200_nccl_unique_id_f

  Integer Function emgmm_b200_joinNccl(this, id, rank, world) Result(info)
    class(mlf_algo_emgmm_b200), intent(inout), target :: this
    integer(c_signed_char), intent(in) :: id(128)
    integer, intent(in) :: rank, world
    info = c_engine_nccl_init(this%engine, id, INT(rank, c_int), INT(world, c_int))
  End Function emgmm_b200_joinNccl

  Subroutine emgmm_b200_finalize(this)
    class(mlf_algo_emgmm_b200), intent(inout) :: this
    If(C_ASSOCIATED(this%engine)) CALL c_engine_destroy(this%engine)
    this%engine = C_NULL_PTR
    CALL mlf_obj_finalize(this)
  End Subroutine emgmm_b200_finalize

  Integer Function emgmm_b200_getNumClasses(this) Result(nC)
    class(mlf_model_emgmm_b200), intent(in), target :: this
    nC = SIZE(this%top%Mu, 2)
  End Function emgmm_b200_getNumClasses

  ! mlf_algo_emgmm_getLikelihood (src/unsupervised/mlf_emgmm.f90:161-171): ExpStep on caller data, Proba(nIn,nC).
  Integer Function emgmm_b200_getLikelihood(this, X, Proba) Result(info)
    class(mlf_model_emgmm_b200), intent(in), target :: this
    real(c_double), intent(in) :: X(:,:)
    real(c_double), intent(out) :: Proba(:,:)
    If(.NOT. this%top%initialized) Then
      info = mlf_UNINIT
      RETURN
    Endif
    info = run_expectation(this%top, X, SIZE(X,1), SIZE(X,2), Proba, SIZE(this%top%Mu,2))
  End Function emgmm_b200_getLikelihood

  ! Explicit-shape dummies make the arrays contiguous for C (copy-in/copy-out if the caller passed a section).
  Integer Function run_expectation(top, X, nY, nIn, Proba, nC) Result(info)
    class(mlf_algo_emgmm_b200), intent(in) :: top
    integer, intent(in) :: nY, nIn, nC
    real(c_double), intent(in) :: X(nY, nIn)
    real(c_double), intent(out), target :: Proba(nIn, nC)
    info = c_engine_expectation(top%engine, top%Mu, top%Cov, top%lambda, X, INT(nIn, c_int64_t), &
      C_LOC(Proba), C_NULL_PTR)
  End Function run_expectation

  ! The data of the reference's own EM test program (tests/test_emgmm.f90:105-124: RandN centres, randOrthogonal(C12, weight),
  ! RandN(X, X0 = XC(:,k), C12 = C12(:,:,k)), randPerm) generated by one CUDA kernel: X(ND, N), XC(ND, NC), C12(ND, ND, NC).
  ! Seeded and counter-based (the reference's RANDOM_NUMBER stream is not): same seed, same data on any number of GPUs.
  Integer Function mlf_b200_synth_gmm_f(X, XC, C12, sigmaC, sigmaX, seed) Result(info)
    real(c_double), intent(out), contiguous :: X(:,:), XC(:,:), C12(:,:,:)
    real(c_double), intent(in) :: sigmaC, sigmaX
    integer(c_int64_t), intent(in) :: seed
    integer(c_int64_t) :: nPts
    nPts = INT(SIZE(X,2), c_int64_t)
    info = c_synth_gmm(X, 0_c_int64_t, nPts, nPts, INT(SIZE(X,1), c_int), INT(SIZE(XC,2), c_int), sigmaC, sigmaX, seed, &
      0_c_int, 0_c_int, XC, C12)
  End Function mlf_b200_synth_gmm_f

End Module mlf_emgmm_b200
