/*
 * mlf_b200.h -- C-ABI of the B200-native drop-in for MlFortran's Gaussian-mixture EM path.
 *
 * The first block re-declares, with the *Fortran-true* signatures, the subset of the reference's C handle
 * interface (include/mlf_cintf.h) that touches the EM path.  Where mlf_cintf.h and the Fortran bind(C)
 * definitions disagree (mlf_getClass/mlf_getProba, mlf_evalGaussian/mlf_maxGaussian have swapped prototypes
 * in the header, include/mlf_cintf.h:109-111,122-123), the Fortran definitions are the real ABI and are
 * what is declared here.  Handles returned by this library are only valid with this library.
 *
 * Array layouts are the reference's (Fortran column-major):
 *   X(d,N) -> X[i*d + a];  Mu(d,K) -> Mu[k*d + a];  Cov(d,d,K) -> Cov[k*d*d + b*d + a];  Proba(N,K) -> P[k*N + i].
 */
#ifndef MLF_B200_H
#define MLF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLF_MAXRANK 15

typedef void MLF_OBJ;

/* include/mlf_cintf.h:42-45 */
typedef enum { mlf_UNK = 0, mlf_BOOL = 1, mlf_INT = 2, mlf_INT64 = 3, mlf_SIZE = 4, mlf_FLOAT = 5, mlf_DOUBLE = 6, mlf_SIZEPARAM = 7, mlf_RAW = 8 } MLF_DATATYPE;
typedef enum { mlf_DIRECT = 1, mlf_INDIRECT = 2, mlf_READONLY = 3, mlf_COPYONLY = 4, mlf_WRITEONLY = 5 } MLF_ACCESSTYPE;
typedef enum { mlf_NAME = 1, mlf_DESC = 2, mlf_FIELDS = 3, mlf_FUNINFO } MLF_INFOTYPE;
typedef enum { mlf_OK = 0, mlf_UNINIT = -1, mlf_FUNERROR = -2, mlf_WRONGTYPE = -3, mlf_WRONGRANK = -4, mlf_FILENOTFOUND = -5, mlf_FILEERROR = -6, mlf_OTHERERROR = -7 } MLF_ERRORTYPE;
#define mlf_STEP_OK 0
#define mlf_STEP_STOP 1 /* src/mlf_step_algo.f90:40 */

typedef struct {
  short dt;     /* MLF_DATATYPE */
  short access; /* MLF_ACCESSTYPE */
} MLF_DT;      /* include/mlf_cintf.h:47-50, src/mlf_intf.f90:66-69 */

/* ---- reference interface (same names, argument meaning and error behaviour) ------------------------------- */

int mlf_init(void);  /* src/mlf_intf.f90:439-443 (h5open_f) -- here: checks that a CUDA device exists */
int mlf_quit(void);  /* src/mlf_intf.f90:445-449 */

int mlf_dealloc(MLF_OBJ *obj);   /* src/mlf_intf.f90:257-278; -1 on NULL */
int mlf_getnumrsc(MLF_OBJ *obj); /* src/mlf_intf.f90:323-331; 7 for an EM object, -1 on NULL/foreign */

/* src/mlf_intf.f90:335-351 + src/mlf_rsc_array.f90:617-635,718-792.  id is 1-based:
 * 1 iPar(int64[1]) 2 rPar(f64[2]) 3 iVar(int64[2]) 4 rVar(f64[3]) 5 Mu(f64 d x K) 6 Cov(f64 d x d x K) 7 lambda(f64 K).
 * Fills rank and Fortran-order dims, dt = {mlf_DOUBLE|mlf_INT64, mlf_DIRECT}; returns a direct pointer into the
 * object's host array when data == NULL, otherwise memcpy's into data and returns data.  NULL on error. */
void *mlf_getrsc(MLF_OBJ *obj, int id, MLF_DT *dt, int *rank, int *dims, void *data);
char *mlf_getinfo(MLF_OBJ *obj, int id, int type);      /* src/mlf_intf.f90:354-366; NUL-terminated, object-owned */
int mlf_updatersc(MLF_OBJ *obj, int id, void *data);    /* src/mlf_intf.f90:369-382 */
int mlf_pushState(MLF_OBJ *handler, MLF_OBJ *obj, int override_); /* src/mlf_intf.f90:384-396, src/mlf_hdf5.f90:671-737 */

/* src/mlf_step_algo.f90:204-219: niter iterations; on a fresh object the first one is the initialiser
 * (src/unsupervised/mlf_emgmm.f90:135-145).  Returns 0, mlf_STEP_STOP (1) when cpuTime>cpuMax or realTime>realMax,
 * -1 on NULL/foreign handles, other negatives on errors.  *dt receives the wall time of the call (the reference
 * never writes it: "dt0 = dt", src/mlf_step_algo.f90:195). */
int mlf_step(MLF_OBJ *obj, double *dt, int64_t *niter);

/* Model interface, Fortran-true argument order (src/mlf_models.f90:277-289,302-313,324-331): nX = dimension d,
 * nIn = number of points.  Cl is 1-based (MAXLOC).  P is written as Proba(nIn, nCl): P[k*nIn + i]
 * (the reference's c_getProba views it as [nCl,nIn], which is inconsistent with ExpStep). */
int mlf_getClass(MLF_OBJ *obj, const double *X, int *Cl, int nX, int nIn);
int mlf_getProba(MLF_OBJ *obj, const double *X, double *P, int nX, int nIn, int nCl);
int mlf_getNumClasses(MLF_OBJ *obj);

MLF_OBJ *mlf_hdf5_createFile(const char *fname, int trunk); /* src/mlf_hdf5.f90:168-188 */
MLF_OBJ *mlf_hdf5_openFile(const char *fname, int rw);      /* src/mlf_hdf5.f90:205-225 */

/* Single-component entry points (src/unsupervised/mlf_gaussian.f90:53-59,81-88). */
int mlf_evalGaussian(int ND, int NX, const double *X, double *P, const double *mu, const double *C, double lambda, double *sumLL);
double mlf_maxGaussian(int ND, int NX, const double *X, const double *Proba, double *mu, double *C);

/* ---- extension the reference lacks: an EM constructor for C callers ------------------------------------------
 * Modelled on mlf_kmeans_c (src/unsupervised/mlf_kmeans_naive.f90:99-116) and on mlf_getoptimobj's optional
 * checkpoint handle (src/mlf_optim_intf.f90:51-88).  X is nY x nX (nX points of dimension nY), borrowed like the
 * reference's `this%X => X` but copied once to HBM.  handler != NULL restores Mu/Cov/lambda (and K) from the
 * checkpoint and marks the object initialised (src/unsupervised/mlf_emgmm.f90:90-110).  NULL on failure. */
MLF_OBJ *mlf_emgmm_c(double *X, int nX, int nY, int nC, double minProba, MLF_OBJ *handler);

/* ---- B200-specific extensions ----------------------------------------------------------------------------------- */
#define MLF_B200_X_ON_DEVICE 1u /* X is a device pointer (kept alive by the caller) */
#define MLF_B200_COV_DIAG 2u    /* EXTENSION: diagonal covariances (as mlf_b200_set_cov_type(obj, 1), but known at construction, which
                                   lets K = 256 at d = 8 run through the single-sweep kernel: BASELINE config 5) */

/* As mlf_emgmm_c with 64-bit point count, device selection (-1 = current) and flags. */
MLF_OBJ *mlf_b200_emgmm_create(const double *X, int64_t nX, int nY, int nC, double minProba, MLF_OBJ *handler, int device, unsigned flags);

/* Single-process multi-GPU object: the points are sharded in contiguous blocks over `devices` (ndevices <= 0: every visible
 * device), one engine and one host thread per shard, and the statistics all-reduce is done by the library itself over peer
 * memory (NVLink P2P loads, summed in rank order so that all shards hold bit-identical parameters).  This is how a plain
 * one-process caller of the reference's API (Fortran `init(X, nC)` / `step`) gets every GPU of the box.  X must be a host
 * array.  The handle behaves like any other EM handle (mlf_step, mlf_getrsc, mlf_pushState, mlf_getClass, ...). */
MLF_OBJ *mlf_b200_emgmm_create_multi(const double *X, int64_t nX, int nY, int nC, double minProba, MLF_OBJ *handler,
                                     const int *devices, int ndevices, unsigned flags);

/* Sharded operation (one process per GPU): install an in-place sum all-reduce over `count` doubles at device
 * pointer buf, enqueued on CUDA stream `stream`.  The object then treats its X as one shard of the global data. */
typedef int (*mlf_b200_allreduce_fn)(void *ctx, double *buf, int64_t count, void *stream);
int mlf_b200_set_allreduce(MLF_OBJ *obj, mlf_b200_allreduce_fn fn, void *ctx, int rank, int world);
/* Native NCCL instead of a callback (one process per GPU; libnccl.so.2 is dlopen'ed on first use, the library does not link
 * it): the per-iteration statistics all-reduce becomes one ncclAllReduce(ncclDouble, ncclSum) on the object's stream.
 *   mlf_b200_set_nccl_comm   borrow a communicator the caller built (ncclComm_t passed as void*); rank / world come from it;
 *   mlf_b200_nccl_unique_id  rank 0 fills 128 bytes (ncclUniqueId) that the caller distributes (MPI_Bcast, a file, ...);
 *   mlf_b200_nccl_init       every rank: ncclCommInitRank on the object's device with that id; the communicator belongs to the
 *                            object and is destroyed by mlf_dealloc.
 * All ranks must construct their objects with shards of one data set and call mlf_step the same number of times. */
int mlf_b200_set_nccl_comm(MLF_OBJ *obj, void *nccl_comm);
int mlf_b200_nccl_unique_id(void *id128);
int mlf_b200_nccl_init(MLF_OBJ *obj, const void *id128, int rank, int world);

int mlf_b200_set_initialized(MLF_OBJ *obj, int initialized); /* the Fortran API exposes %initialized directly */
int mlf_b200_get_initialized(MLF_OBJ *obj);
/* Seed of the k-means initialiser's random stream (the reference uses an unseeded RANDOM_NUMBER, src/mlf_rand.f90:173-182). */
int mlf_b200_set_seed(MLF_OBJ *obj, uint64_t seed);
/* EXTENSION (the reference has full covariances only): cov_type 0 = full (default, reference behaviour), 1 = diagonal --
 * the E-step reads diag(C_k) only and the M-step keeps the diagonal of the reference's C_k (BASELINE config 5).
 * Returns -7 (mlf_OTHERERROR) when asked for full covariances on an object constructed with MLF_B200_COV_DIAG whose
 * kernels were planned around the diagonal-compressed parameter pack (d <= 8 with K >= 128): construct a new object. */
int mlf_b200_set_cov_type(MLF_OBJ *obj, int cov_type);
/* Re-copy X after the caller changed it.  For a host array on the fused path (d <= 16) the upload is DEFERRED: it runs chunk
 * by chunk underneath the sweep of the next mlf_step, so X must stay valid and unchanged until that step has returned (the
 * reference's own contract is stronger: `this%X => X` borrows the array for the object's lifetime,
 * src/unsupervised/mlf_emgmm.f90:86).  Other paths copy before returning. */
int mlf_b200_refresh_x(MLF_OBJ *obj, const double *X);
/* Page-locked host memory for arrays that are handed to mlf_emgmm_c / mlf_b200_refresh_x / mlf_getClass / mlf_getProba
 * (pageable arrays work too, through a staging copy at roughly half the PCIe rate).  write_combined != 0 asks for
 * write-combined memory: not snooped by the CPU caches during the GPU's reads, for upload-only buffers that the host
 * fills once and never reads back.  NULL on failure. */
void *mlf_b200_host_alloc(size_t bytes, int write_combined);
void mlf_b200_host_free(void *p);
const char *mlf_b200_last_error(MLF_OBJ *obj);
/* Which kernel generation this object's (d, K) runs through (object-owned string). */
const char *mlf_b200_kernel_plan(MLF_OBJ *obj);
/* sumLL of each EM iteration of the most recent mlf_step call (at most n values); returns how many exist. */
int64_t mlf_b200_sumll_trace(MLF_OBJ *obj, double *out, int64_t n);
/* Responsibilities (ProbaC(N,K), P[k*N+i]) and labels of the resident shard, as of the last E-pass. */
int mlf_b200_resident_proba(MLF_OBJ *obj, double *P);
int mlf_b200_resident_labels(MLF_OBJ *obj, int *Cl);
/* Device pointer of the resident responsibilities (component k starts at k * *ld doubles; columns >= N are zero). */
const double *mlf_b200_resident_proba_device(MLF_OBJ *obj, int64_t *ld);
/* Profiling counters: out[0..5] = ms in refresh, sweep / E-pass, mid, M-pass / side-list resolution, finalize, and iterations
 * counted; out[6] = sweeps over X (the fused path needs 2 in an iteration whose threshold band missed), out[7] = pairs settled
 * through the side list. */
int mlf_b200_set_profile(MLF_OBJ *obj, int on);
int mlf_b200_get_times(MLF_OBJ *obj, double out[8]);
/* Integer counters since the last mlf_b200_set_profile / reset: out[0] = EM iterations counted (profiling on), out[1] = sweeps
 * over X, out[2] = pairs settled through the side list, out[3] = (point, component) pairs whose covariance term was accumulated
 * inside the sweep ("kept" by the reference's W >= minW test, src/mlf_matrix.f90:135; counted by the 16-warp sweep only). */
int mlf_b200_get_counters(MLF_OBJ *obj, int64_t out[4]);
/* The 16-warp sweep (d <= 16, K <= 64) evaluates the log-densities either by whitening the points or, when that is safe, as the
 * expanded quadratic form c + b.x' - 1/2 x'^T invC x' about the data's mean (fewer tensor-core tiles, no squares / reduction).
 * "Safe" is decided per refresh from a bound S_k on the magnitude of the form's terms over the data (rounding error ~ 4 u S_k,
 * held against MLF_B200_QF_TOL = 5e-11 by default; MLF_B200_QF=0 always whitens, =2 never does).  *qf_sweeps = sweeps since the
 * last counter reset that took the quadratic form, *bound = max_k S_k of the most recent sweep.  Either pointer may be NULL. */
int mlf_b200_get_qf_stats(MLF_OBJ *obj, int64_t *qf_sweeps, double *bound);
/* Iterations (since the last counter reset) whose covariance thresholds could not be predicted and that therefore ran
 * "statistics first": one sweep for the E-step and the pass-A statistics with the responsibilities kept in HBM (8 N K bytes,
 * allocated at first need; MLF_B200_GSTORE=0 forbids it), then a scatter pass over the stored responsibilities with the exact
 * thresholds -- instead of a second full sweep. */
int64_t mlf_b200_scatter_passes(MLF_OBJ *obj);
int64_t mlf_b200_launch_count(MLF_OBJ *obj);
void *mlf_b200_stream(MLF_OBJ *obj);
int64_t mlf_b200_global_points(MLF_OBJ *obj);
const char *mlf_b200_version(void);
/* Self-test hook: the device exp of the softmax (exp_tab.cuh: table-driven, 256 entries, arguments <= 0; k_em16) evaluated on n host values on
 * device `device`; out[i] = exp(in[i]).  Exists so that the tests can pin the kernel's transcendental against libm. */
int mlf_b200_selftest_exp(const double *in, double *out, int64_t n, int device);
/* The same for the 128-entry form of the table-driven exp that k_emd (diagonal covariances) keeps. */
int mlf_b200_selftest_exp128(const double *in, double *out, int64_t n, int device);

/* Workload generator of the reference's own EM test program (tests/test_emgmm.f90:105-124, testGMM) run on the device:
 * centres = sigmaC N(0,I) (RandN, src/mlf_rand.f90:241-257), C12_k = O_k diag(w), O_k Haar orthogonal (randOrthogonal,
 * src/mlf_matrix.f90:153-180), w_i = sqrt(sigmaX (1/i) / mean_j(1/j)), x = c_k + C12_k N(0,I), points shuffled (randPerm).
 * Points [n0, n0 + nX) of a job of nTotal points are written to X (nY x nX, i.e. [nX][nY] in memory; a device pointer with
 * MLF_B200_X_ON_DEVICE in flags, else a host array).  Every value is a pure function of (seed, global point index) --
 * counter-based Philox4x32-10 + Box-Muller, keyed Feistel permutation -- so shards generated on different GPUs form the same
 * data set as one call with n0 = 0, nX = nTotal.  The component of output point i is pi(i) mod nC.  The reference's RNG is
 * gfortran's unseeded RANDOM_NUMBER: the distribution is reproduced, not its bits.  centres (nY x nC) and C12 (nY x nY x nC,
 * Fortran order) are host outputs, nullable.  nX = 0 only computes those (no device needed).  0, or a negative MLF_ERRORTYPE. */
int mlf_b200_synth_gmm(double *X, int64_t n0, int64_t nX, int64_t nTotal, int nY, int nC, double sigmaC, double sigmaX, uint64_t seed,
                       int device, unsigned flags, double *centres, double *C12);

/* ---- engine-level entry points: what a Fortran extension type binds with iso_c_binding --------------------------
 * Deliverable form (A) of INTEGRATION.md: fortran/mlf_emgmm_b200.f90 declares a type extending the reference's
 * mlf_step_obj whose init/reinit/stepF forward here, so that the reference's generic handle functions, HDF5
 * checkpoint and resource slots keep working unchanged on Fortran-owned host arrays.  The engine owns only device
 * state; Mu/Cov/lambda are caller arrays in the reference layout (uploaded at the start of a call, downloaded at
 * its end), replacing ExpStep/MaxStep of src/unsupervised/mlf_emgmm.f90:174-207. */
typedef void MLF_B200_ENGINE;
MLF_B200_ENGINE *mlf_b200_engine_create(const double *X, int64_t nX, int nY, int nC, int device, unsigned flags);
/* As above, sharded over `devices` inside this process (ndevices <= 0: every visible device); see mlf_b200_emgmm_create_multi. */
MLF_B200_ENGINE *mlf_b200_engine_create_multi(const double *X, int64_t nX, int nY, int nC, const int *devices, int ndevices, unsigned flags);
void mlf_b200_engine_destroy(MLF_B200_ENGINE *eng);
/* niter x (ExpStep + MaxStep) from the caller's current Mu/Cov/lambda; sumLL of the last iteration. 0 or negative. */
int mlf_b200_engine_step(MLF_B200_ENGINE *eng, int64_t niter, double minProba, double *Mu, double *Cov, double *lambda, double *sumLL);
/* km_algo%init + stepF(nkm) + Cov = I + lambda = 1/nC  (src/unsupervised/mlf_emgmm.f90:135-145). */
int mlf_b200_engine_kmeans_init(MLF_B200_ENGINE *eng, int nkm, uint64_t seed, double *Mu, double *Cov, double *lambda);
/* ExpStep on caller data with the given parameters: P (nullable) as Proba(nIn,nC), Cl (nullable) 1-based MAXLOC. */
int mlf_b200_engine_expectation(MLF_B200_ENGINE *eng, const double *Mu, const double *Cov, const double *lambda,
                                const double *Xq, int64_t nIn, double *P, int *Cl);
/* ProbaC(nX,nC) of the resident points as of the last ExpStep (the Fortran type's public ProbaC component). */
int mlf_b200_engine_proba(MLF_B200_ENGINE *eng, double *ProbaC);
int mlf_b200_engine_set_allreduce(MLF_B200_ENGINE *eng, mlf_b200_allreduce_fn fn, void *ctx, int rank, int world);
/* One process per GPU from Fortran (MPI): rank 0 calls mlf_b200_nccl_unique_id, the 128 bytes are broadcast (MPI_Bcast), every
 * rank joins with this call before its first step; the communicator is destroyed with the engine. */
int mlf_b200_engine_nccl_init(MLF_B200_ENGINE *eng, const void *id128, int rank, int world);
const char *mlf_b200_engine_last_error(MLF_B200_ENGINE *eng);

/* Generic dataset access on a checkpoint handle (file or group handle; int64 / float64; dims in file (C) order, i.e. the
 * reverse of the Fortran shape; `name` may be a path "group/sub/name" below the handle).  mlf_b200_h5_get with
 * data == NULL only reports type/rank/dims.  mlf_b200_h5_count / _name enumerate the datasets below the handle. */
int mlf_b200_h5_put(MLF_OBJ *handler, const char *name, int dtype, int rank, const int64_t *dims, const void *data, int override_);
int mlf_b200_h5_get(MLF_OBJ *handler, const char *name, int *dtype, int *rank, int64_t *dims, void *data, int64_t capacity_bytes);
int mlf_b200_h5_count(MLF_OBJ *handler);
const char *mlf_b200_h5_name(MLF_OBJ *handler, int idx);
/* Objects / features of the opened file that lie outside the checkpoint layout (other types, chunked or compressed
 * layouts, attributes, a user block): readable files that hold any are opened read-only; mlf_hdf5_openFile(.., rw=1)
 * refuses them, because the built-in writer rewrites the file as a whole and would drop that content (the reference's
 * H5Fopen_f(H5F_ACC_RDWR_F), src/mlf_hdf5.f90:204-207, keeps it). */
int mlf_b200_h5_unsupported(MLF_OBJ *handler);
/* Group handler below a file or group handler: mlf_hdf5_openGroup / getSubGroup (src/mlf_hdf5.f90:106-151).
 * mode 0 = open an existing group, 1 = create (NULL if it exists), 2 = open, else create.  mlf_pushState on the returned
 * handle writes the object's slots inside the group, as Hdf5PushSubObject does for sub-objects (:657-669), and
 * mlf_emgmm_c(..., handler = group) restores from it.  Free with mlf_dealloc; the file is written when something was put. */
MLF_OBJ *mlf_b200_h5_open_group(MLF_OBJ *handler, const char *name, int mode);

#ifdef __cplusplus
}
#endif
#endif /* MLF_B200_H */
