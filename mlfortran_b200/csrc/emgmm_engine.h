// Device engine of the EM-GMM step object: owns the HBM-resident shard of X, the responsibilities and the
// per-iteration kernel sequence.  Host code (mlf_capi.cu) keeps the reference-visible host mirrors.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "emgmm_kernels.cuh"

namespace mlfb200 {

// In-place sum all-reduce of `count` doubles that live at device pointer `buf`, enqueued on `stream`.
// Returns 0 on success.  Installed by the multi-GPU host layer (torch.distributed / NCCL); absent => 1 rank.
typedef int (*allreduce_fn)(void* ctx, double* buf, int64_t count, void* stream);

struct EngineTimes {  // CUDA-event milliseconds accumulated over iterate() calls (profiling counters)
  double refresh = 0, estep = 0, mid = 0, mstep = 0, fin = 0;
  int64_t iters = 0;
  int64_t sweeps = 0;     // fused path: sweeps over X (1 per iteration when the threshold band held, 2 otherwise)
  int64_t amb_pairs = 0;  // fused path: pairs settled through the side list
  int64_t kept_pairs = 0; // fused path (k_em16): (point, component) pairs whose scatter term was accumulated inside the sweep
  int64_t scatter_passes = 0;  // statistics-first iterations whose second pass scattered from stored responsibilities
  int64_t qf_sweeps = 0;  // sweeps whose logits came from the expanded quadratic form (the others whitened the points)
  double qf_bound = 0;    // max_k S_k of the last sweep (magnitude bound of the quadratic form's terms); inf: radius unknown
};

class EmgmmEngine {
 public:
  EmgmmEngine() = default;
  ~EmgmmEngine();
  EmgmmEngine(const EmgmmEngine&) = delete;
  EmgmmEngine& operator=(const EmgmmEngine&) = delete;

  // X: N_local x d, point-major (Fortran X(d,N)).  x_on_device: X is a device pointer that the caller keeps
  // alive (borrowed, like the reference's `this%X => X`); otherwise it is copied host->device here.
  // cov_diag: diagonal-covariance extension, fixed at construction so that the kernel plan can use the compressed pack
  int init(const double* X, bool x_on_device, int64_t n_local, int d, int K, int device, bool cov_diag = false);
  void set_allreduce(allreduce_fn fn, void* ctx, int rank, int world) {
    ar_ = fn; ar_ctx_ = ctx; rank_ = rank; world_ = world < 1 ? 1 : world; moments_ready_ = false; shift_ready_ = false; rad_ready_ = false;
  }
  void set_stream(cudaStream_t st);
  int refresh_x(const double* X, bool x_on_device);
  // Global moments + shift of the resident data.  COLLECTIVE when an all-reduce is installed: every rank / shard must
  // call it (empty shards too).  step / kmeans_init / expectation / resident_* call it themselves when the moments are
  // stale (after init, refresh_x, set_allreduce), so the first of those calls is collective as well; callers that want
  // to run inference on one rank only call prepare() on all ranks first.
  int prepare();

  int upload_params(const double* Mu, const double* Cov, const double* lambda);
  int download_params(double* Mu, double* Cov, double* lambda);

  // niter EM iterations (ExpStep + MaxStep, src/unsupervised/mlf_emgmm.f90:147-149) on resident data.
  // sumLL_last receives sumLL of the last iteration; sumLL_trace (optional, host, niter entries) all of them.
  int iterate(int64_t niter, double minProba, double* sumLL_last, double* sumLL_trace);

  // ExpStep on caller data with the current device parameters (getProba / getClass).  Xq host or device.
  // P (optional) is written component-major P[k*nq + i] to host memory; labels (optional) 1-based.
  int expectation(const double* Xq, bool on_device, int64_t nq, double* P_host, int32_t* labels_host);
  // Labels / responsibilities of the resident shard from the most recent E-pass parameters.
  int resident_labels(int32_t* labels_host);
  int resident_proba(double* P_host);  // ProbaC(N,K): P[k*N + i]
  const double* gamma_device();   // materialises ProbaC on demand
  int materialize_gamma();
  int64_t gamma_ld() const { return ldg_; }

  // Single-component entry points (K = 1 engines): see emgmm_engine.cu.
  int eval_gaussian_raw(double* P_host, double* sumLL);
  int max_gaussian(const double* Proba_host, double* mu_host, double* C_host, double* sumPi);

  // k-means initialiser (src/unsupervised/mlf_emgmm.f90:135-145): seeding + nkm-1 Lloyd steps, Cov=I, lambda=1/K.
  int kmeans_init(int nkm, uint64_t seed);

  int64_t n_local() const { return n_; }
  int64_t n_global() const { return nglob_; }
  int d() const { return geom_.d; }
  int K() const { return geom_.K; }
  const Geom& geom() const { return geom_; }
  cudaStream_t stream() const { return st_; }
  int device() const { return device_; }
  const std::string& last_error() const { return err_; }
  // which kernel generation this (d, K) runs through (for bench output and logs)
  std::string plan() const;
  EngineTimes& times() { return times_; }
  void reset_times() { times_ = EngineTimes(); }
  void set_profile(bool on) { profile_ = on; }
  // EXTENSION (no reference semantics): diagonal covariances -- the E-step reads diag(C_k) only and the
  // M-step keeps the diagonal of the reference's C_k.  Default: full covariances (reference behaviour).
  // false when the object was planned around the diagonal-compressed pack (diagonal covariances at construction, K >= 128
  // at d <= 8): its kernels cannot represent full covariances, a new object is needed for that
  bool set_cov_diag(bool on) {
    if (!on && fgeom_.diag) { err_ = "object planned for diagonal covariances: construct a new one for full covariances"; return false; }
    cov_diag_ = on;
    return true;
  }
  bool cov_diag() const { return cov_diag_; }
  int64_t launches() const { return launches_; }

 private:
  int fail(const char* what, cudaError_t e);
  int ensure_moments();
  int run_estep(const double* X, int64_t n, double* gamma, int64_t ldg, int32_t* labels, bool want_stats, bool raw = false,
                bool hard = false);
  int allreduce(double* buf, int64_t count);
  int iterate_fused(int64_t niter, double minProba, bool prof);
  int iterate_two_pass(int64_t niter, double minProba, bool prof);
  int ensure_gamma();
  int do_refresh(bool with_sumll = true);
  int run_mstep(const double* gamma);
  int flush_pending();
  int finish_moments(int grid);
  int streamed_sweep(const EmArgs& base);
  int draw_far_point(double u, int grid, int64_t chunk, double* dMinDist, double* dPart, double* dOut);

  Geom geom_{};
  FusedGeom fgeom_{};
  bool fused_ = false;        // single-sweep path available for this (d, K)
  bool used_ = false;         // diagonal fused sweep on the CUDA cores (emgmm_diag.cu; MLF_B200_EMD=0 disables)
  int spl_ = 1;               // k_em16: warps that share one component block's statistics (em16_split; 1 at K > 32)
  bool use16_ = false;        // 16-warp fused sweep with the scatter accumulators in tensor memory (MLF_B200_EM16=0 disables)
  bool legacy_estep_ = false; // MLF_B200_LEGACY=1: first-generation kernels (A/B comparisons)
  bool big_ = false;          // large-shape kernels (d > 32 or pack too large for shared memory; MLF_B200_BIG=1 forces them)
  int64_t mchunk_ = 0;
  int64_t refresh_count_ = 0;  // warm-started Jacobi refresh at large d (see k_refresh)
  int device_ = 0, sms_ = 148;
  size_t smem_limit_ = 0;
  cudaStream_t st_ = nullptr;
  bool own_stream_ = false;
  int64_t n_ = 0, nglob_ = 0, ldg_ = 0;
  const double* dX_ = nullptr;
  double* dXown_ = nullptr;
  double* dGamma_ = nullptr;
  double *dMu_ = nullptr, *dCov_ = nullptr, *dLambda_ = nullptr;
  double *dMupad_ = nullptr, *dThr_ = nullptr, *dPack_ = nullptr, *dScratch_ = nullptr, *dSumLLk_ = nullptr;
  double *dSumLL_ = nullptr;   // trace buffer on device
  int64_t trace_cap_ = 0;
  double *dStatsA_ = nullptr, *dStatsB_ = nullptr, *dPartA_ = nullptr, *dPartB_ = nullptr;
  double *dSc_ = nullptr, *dMean_ = nullptr, *dNglob_ = nullptr;
  int* dInfo_ = nullptr;
  int* dVValid_ = nullptr;   // per component: the eigenvectors in scratch are valid (the eigen branch of k_refresh ran)
  // fused path
  double *dPack2_ = nullptr, *dMuOld_ = nullptr, *dTlo_ = nullptr, *dThi_ = nullptr, *dAB_ = nullptr, *dPartB2_ = nullptr, *dStatsC_ = nullptr;
  double *dPrevMu_ = nullptr, *dPrevCov_ = nullptr, *dPrevLambda_ = nullptr;  // parameters at the start of the last iteration (ProbaC)
  bool prev_valid_ = false;
  // streamed refresh of X: host pointer waiting to be uploaded underneath the next sweep
  int kStreamChunks = 16;     // pieces of a streamed upload (MLF_B200_STREAM_CHUNKS)
  const double* pending_host_ = nullptr;
  double *dShift_ = nullptr, *dMomPart_ = nullptr, *dMomRaw_ = nullptr;
  bool shift_ready_ = false;
  cudaStream_t cst_ = nullptr;
  std::vector<cudaEvent_t> cev_;
  bool gamma_current_ = false;  // dGamma_ holds the responsibilities of the last E-step
  AmbEntry* dAmb_ = nullptr;
  int* dAmbCount_ = nullptr;
  int amb_cap_ = 2048;          // entries per list; grown (x4, up to kAmbCapMax) when a sweep overflows a list
  static constexpr int kAmbCapMax = 262144;
  int amb_lists_ = 0;           // lists allocated (16 per block for the 16-warp sweep, 8 otherwise)
  int* dAmbComp_ = nullptr;     // k_amb_partial scratch: components present per list
  double* dAmbPart_ = nullptr;  //                       and their partial [scatter | sum g z | sum g]
  int amb_cmax_ = 8;
  bool debug_ = false;          // MLF_B200_DEBUG=1: one stderr line per sweep (threshold band / side list / repeat decisions)
  int ensure_amb_scratch();
  int grow_amb_lists(int64_t longest);
  std::vector<double> s_prev_, s_prev2_, s_prev3_;   // s_k of the last three iterations (threshold-band predictor)
  int band_misses_ = 0;                             // > 0: recent prediction miss -> statistics-first iterations
  std::vector<double> band_widen_;                  // per-component widening factor: x4 after a miss, halved after a hit
  // quadratic-form variant of the 16-warp sweep (emgmm_kernels.cuh, qf_*): coefficient pack, per-component magnitude bounds,
  // per-rank radius of the data about the shift (gathered), and the limit the bounds are held against
  double *dPackQ_ = nullptr, *dQfS_ = nullptr, *dRad_ = nullptr, *dRadPart_ = nullptr;
  int rad_world_ = 0;           // rank slots allocated in dRad_
  bool qf16_ = false, qfd_ = false;   // which sweep has an expanded-form variant: k_em16 (full covariances), k_emd (diagonal)
  bool qf_ = false;             // variant available for this shape (MLF_B200_QF=0 disables, =2 forces it whatever the bounds say)
  bool rad_ready_ = false;
  double qf_limit_ = 0;         // S_k <= qf_limit_  <=>  estimated logit error 4 u S_k <= MLF_B200_QF_TOL (default 5e-11)
  int ensure_radius();
  double* dGTiles_ = nullptr;   // responsibilities of statistics-first iterations, [tile][64][K8*8]
  bool gtiles_failed_ = false;
  bool ensure_gtiles();
  int32_t* dLabels_ = nullptr;
  int egrid_ = 0, mgx_ = 0, mgy_ = 0;
  bool moments_ready_ = false;
  allreduce_fn ar_ = nullptr;
  int rank_ = 0, world_ = 1;
  void* ar_ctx_ = nullptr;
  std::string err_;
  EngineTimes times_;
  bool profile_ = false;
  bool cov_diag_ = false;
  int64_t launches_ = 0;
  std::vector<cudaEvent_t> pev_;  // 6 profiling events per iteration (only when set_profile(true))
  // streamed inference on caller data (expectation): persistent double buffers, three streams (H2D copy, compute, D2H copy)
  struct InferBuf {
    double *dX = nullptr, *dP = nullptr, *hX = nullptr, *hP = nullptr;   // device chunk buffers; pinned staging (pageable callers)
    int32_t *dL = nullptr, *hL = nullptr;
    cudaEvent_t in = nullptr, comp = nullptr, out = nullptr;
    int64_t s = 0, n = 0;   // chunk in flight whose staged output still has to be handed to the caller
    bool pending = false;
  };
  InferBuf inf_[2];
  int64_t inf_chunk_ = 0;
  bool inf_hasP_ = false, inf_hasL_ = false, inf_hasHX_ = false;
  cudaStream_t dst_ = nullptr;   // device -> host copies of inference results
  int ensure_infer_buffers(int64_t chunk, bool wantP, bool wantL, bool stageX);
  void free_infer_buffers();
};

}  // namespace mlfb200
