// EM-GMM device engine: see emgmm_engine.h.  Reference call sequence being replaced:
// mlf_algo_emgmm_stepF -> ExpStep -> mlf_EvalGaussian / MaxStep -> mlf_MaxGaussian -> GECovMatrix
// (src/unsupervised/mlf_emgmm.f90:125-207).
#include "emgmm_engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace mlfb200 {

#define CKE(call, what)                         \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) return fail(what, e__); \
  } while (0)

int EmgmmEngine::fail(const char* what, cudaError_t e) {
  err_ = std::string(what) + ": " + cudaGetErrorString(e);
  fprintf(stderr, "mlfortran_b200: %s\n", err_.c_str());
  return -7;  // mlf_OTHERERROR
}

EmgmmEngine::~EmgmmEngine() {
  if (st_ || dGamma_ || dXown_) {
    cudaSetDevice(device_);
    if (st_) cudaStreamSynchronize(st_);
  }
  double* bufs[] = {dXown_, dGamma_, dMu_, dCov_, dLambda_, dMupad_, dThr_, dPack_, dScratch_, dSumLLk_, dSumLL_,
                    dStatsA_, dStatsB_, dPartA_, dPartB_, dSc_, dMean_, dNglob_};
  for (double* b : bufs)
    if (b) cudaFree(b);
  double* fb[] = {dPack2_, dMuOld_, dTlo_, dThi_, dAB_, dPartB2_, dStatsC_, dPrevMu_, dPrevCov_, dPrevLambda_, dShift_, dMomPart_, dMomRaw_,
                  dPackQ_, dQfS_, dRad_, dRadPart_, dGTiles_};
  for (auto& e : cev_)
    if (e) cudaEventDestroy(e);
  if (cst_) cudaStreamDestroy(cst_);
  for (double* b : fb)
    if (b) cudaFree(b);
  if (dAmb_) cudaFree(dAmb_);
  if (dAmbCount_) cudaFree(dAmbCount_);
  if (dAmbComp_) cudaFree(dAmbComp_);
  if (dAmbPart_) cudaFree(dAmbPart_);
  if (dInfo_) cudaFree(dInfo_);
  if (dVValid_) cudaFree(dVValid_);
  if (dLabels_) cudaFree(dLabels_);
  for (auto& e : pev_)
    if (e) cudaEventDestroy(e);
  free_infer_buffers();
  for (InferBuf& b : inf_) {
    if (b.in) cudaEventDestroy(b.in);
    if (b.comp) cudaEventDestroy(b.comp);
    if (b.out) cudaEventDestroy(b.out);
  }
  if (dst_) cudaStreamDestroy(dst_);
  if (own_stream_ && st_) cudaStreamDestroy(st_);
}

void EmgmmEngine::set_stream(cudaStream_t st) {
  if (own_stream_ && st_) {
    cudaStreamSynchronize(st_);
    cudaStreamDestroy(st_);
  }
  st_ = st;
  own_stream_ = false;
}

int EmgmmEngine::init(const double* X, bool x_on_device, int64_t n_local, int d, int K, int device, bool cov_diag) {
  cov_diag_ = cov_diag;
  if (n_local < 0 || d < 1 || K < 1) { err_ = "invalid shape"; return -7; }
  if (d > 128) { err_ = "d > 128 is not supported by this build"; return -7; }
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) {
    err_ = "no CUDA device: mlfortran_b200 has no CPU fallback";
    fprintf(stderr, "mlfortran_b200: %s\n", err_.c_str());
    return -7;
  }
  device_ = device < 0 ? 0 : device;
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  cudaDeviceProp prop;
  CKE(cudaGetDeviceProperties(&prop, device_), "cudaGetDeviceProperties");
  sms_ = prop.multiProcessorCount;
  smem_limit_ = prop.sharedMemPerBlockOptin;
  geom_ = make_geom(d, K);
  fgeom_ = make_fused_geom(geom_, cov_diag_, smem_limit_);
  {
    const char* lg = getenv("MLF_B200_LEGACY");
    legacy_estep_ = lg && lg[0] == '1';
    const char* bg = getenv("MLF_B200_BIG");
    big_ = bg && bg[0] == '1';
    const char* dbg = getenv("MLF_B200_DEBUG");
    debug_ = dbg && dbg[0] == '1';
    if (const char* sc = getenv("MLF_B200_STREAM_CHUNKS")) kStreamChunks = std::max(1, std::min(atoi(sc), 256));
    if (const char* ac = getenv("MLF_B200_AMB_CAP")) amb_cap_ = std::max(8, std::min(atoi(ac), kAmbCapMax));   // tests: force overflows
  }
  // kernel generation for this shape: fused single sweep (pack + tile in shared memory, scatter accumulators in
  // registers) > resident-pack E-pass + stored responsibilities > large-shape kernels (pack streamed through L1)
  if (d > 32 || (fgeom_.smem > smem_limit_ && estep_smem_bytes(geom_) > smem_limit_)) big_ = true;
  if (big_) {
    geom_ = make_geom(d, K, true);
    fgeom_ = make_fused_geom(geom_, false, smem_limit_);
    if (!estep_big_ok(geom_, smem_limit_)) {
      err_ = "this (d, K) is outside the shapes the large-shape kernels are built for (d <= 128, K <= 256)";
      return -7;
    }
    legacy_estep_ = false;
  } else if (!legacy_estep_ && fgeom_.smem > smem_limit_) {
    legacy_estep_ = true;  // fused pack does not fit: first-generation E-pass
  }
  fused_ = !big_ && !legacy_estep_ && fgeom_.scatter_ok;
  {
    const char* e16 = getenv("MLF_B200_EM16");
    use16_ = fused_ && em16_ok(geom_, fgeom_, smem_limit_) && !(e16 && e16[0] == '0');
    const char* ed = getenv("MLF_B200_EMD");
    used_ = fused_ && cov_diag_ && emd_ok(geom_, fgeom_, smem_limit_) && !(ed && ed[0] == '0');
    // fewer than 8 component blocks: the statistics of a block are shared by several warps of a group (MLF_B200_EM16_SPLIT=0: one warp)
    const char* es = getenv("MLF_B200_EM16_SPLIT");
    spl_ = (use16_ && !(es && es[0] == '0')) ? em16_split(geom_.K8) : 1;
    // quadratic-form logits inside the 16-warp sweep, guarded by the magnitude bounds of k_refresh
    const char* q = getenv("MLF_B200_QF");
    const int qmode = q ? atoi(q) : 1;
    qf16_ = use16_ && !used_ && !cov_diag_ && qmode != 0 && em16_qf_ok(geom_, fgeom_, smem_limit_);
    qfd_ = used_ && qmode != 0;   // expanded-form P1 of the diagonal sweep (k_emd QP1), same guard
    qf_ = qf16_ || qfd_;
    double tol = 5e-11;
    if (const char* qt = getenv("MLF_B200_QF_TOL")) tol = atof(qt);
    qf_limit_ = qmode == 2 ? INFINITY : tol / (4.0 * 1.1102230246251565e-16);
  }
  n_ = n_local;
  nglob_ = n_local;
  ldg_ = ((n_ + 31) / 32) * 32;
  if (ldg_ == 0) ldg_ = 32;
  CKE(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking), "cudaStreamCreate");
  own_stream_ = true;
  const Geom& g = geom_;
  const int KP = g.K8 * 8;
  const int tp = legacy_estep_ ? kEstepTilePts : fgeom_.TP;
  const int64_t etiles = (n_ + tp - 1) / tp;
  egrid_ = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, etiles));
  mstep_grid(g, sms_, n_, &mgx_, &mgy_);
  if (big_) {
    egrid_ = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, (n_ + estep_big_tile(geom_, smem_limit_) - 1) / estep_big_tile(geom_, smem_limit_)));
    mstep_big_grid(g, sms_, n_, &mgx_, &mchunk_);
  }
  auto A = [&](double** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(double)); };
  if (!fused_) {
    if (int r = ensure_gamma()) return r;  // two-pass path keeps ProbaC(N,K) resident; the fused sweep never stores it
  }
  CKE(A(&dPack2_, (size_t)KP * fgeom_.PSF), "cudaMalloc");
  CKE(A(&dMuOld_, (size_t)KP * 8 * g.DN), "cudaMalloc");
  CKE(A(&dTlo_, KP), "cudaMalloc");
  CKE(A(&dThi_, KP), "cudaMalloc");
  CKE(A(&dAB_, (size_t)KP * g.DS + (size_t)KP * fgeom_.MSF + 8), "cudaMalloc");
  CKE(A(&dStatsC_, (size_t)K * (d * d + d + 1)), "cudaMalloc");
  CKE(A(&dShift_, (size_t)8 * g.DN + 8), "cudaMalloc");
  CKE(A(&dMomPart_, (size_t)sms_ * std::max(g.MB * 64 + 8 * g.DN, d + 1)), "cudaMalloc");
  CKE(A(&dMomRaw_, (size_t)g.MB * 64 + 8 * g.DN), "cudaMalloc");
  CKE(cudaStreamCreateWithFlags(&cst_, cudaStreamNonBlocking), "cudaStreamCreate");
  if (qf_) {
    if (qf16_) CKE(A(&dPackQ_, (size_t)g.K8 * qf_stride(g.DM)), "cudaMalloc");
    CKE(A(&dQfS_, KP), "cudaMalloc");
    CKE(A(&dRadPart_, (size_t)sms_ * 16), "cudaMalloc");
  }
  CKE(A(&dPrevMu_, (size_t)K * d), "cudaMalloc");
  CKE(A(&dPrevCov_, (size_t)K * d * d), "cudaMalloc");
  CKE(A(&dPrevLambda_, K), "cudaMalloc");
  if (fused_) {
    // the 16-warp variant writes partials per (block, half) and keeps one side list per warp
    CKE(A(&dPartB2_, (size_t)2 * sms_ * spl_ * KP * fgeom_.MSF), "cudaMalloc");
    amb_lists_ = sms_ * 16;
    // components a single list can hold: the warp that writes it owns one 8-component block (16-warp sweep), CBMAX blocks
    // (8-warp sweep) or 32 lanes = 32 components (diagonal CUDA-core sweep)
    amb_cmax_ = use16_ ? 8 : (used_ ? 32 : 8 * fgeom_.CBMAX);
    CKE(cudaMalloc((void**)&dAmb_, (size_t)amb_lists_ * amb_cap_ * sizeof(AmbEntry)), "cudaMalloc side list");
    CKE(cudaMalloc((void**)&dAmbCount_, (size_t)amb_lists_ * sizeof(int)), "cudaMalloc");
    CKE(cudaMemsetAsync(dAmbCount_, 0, (size_t)amb_lists_ * sizeof(int), st_), "memset");
  }
  CKE(A(&dMu_, (size_t)K * d), "cudaMalloc");
  CKE(A(&dCov_, (size_t)K * d * d), "cudaMalloc");
  CKE(A(&dLambda_, K), "cudaMalloc");
  CKE(A(&dMupad_, (size_t)K * 8 * g.DN), "cudaMalloc");
  CKE(A(&dThr_, K), "cudaMalloc");
  CKE(A(&dPack_, (size_t)KP * g.PS), "cudaMalloc");
  CKE(A(&dScratch_, (size_t)K * 3 * d * d), "cudaMalloc");
  CKE(A(&dSumLLk_, KP), "cudaMalloc");
  CKE(A(&dStatsA_, (size_t)KP * g.DS), "cudaMalloc");
  CKE(A(&dStatsB_, (size_t)K * g.MS), "cudaMalloc");
  CKE(A(&dPartA_, (size_t)2 * sms_ * spl_ * KP * g.DS), "cudaMalloc");
  // pass-B partials: [mgx_][K][MS] (two-pass kernels, mgx_ <= sms_; large-shape kernels, mgx_*K <= 2*sms_+K), and the
  // unit-weight moments run of the large-shape M-pass ([<= 2*sms_][1][MS])
  CKE(A(&dPartB_, (size_t)std::max<int64_t>((int64_t)mgx_ * K, 2 * (int64_t)sms_ + K + 2) * g.MS), "cudaMalloc");
  CKE(A(&dSc_, (size_t)d * d), "cudaMalloc");
  CKE(A(&dMean_, (size_t)8 * g.DN + 8), "cudaMalloc");
  CKE(A(&dNglob_, 8), "cudaMalloc");
  CKE(cudaMalloc((void**)&dInfo_, sizeof(int) * KP), "cudaMalloc");
  CKE(cudaMalloc((void**)&dVValid_, sizeof(int) * KP), "cudaMalloc");
  CKE(cudaMemsetAsync(dVValid_, 0, sizeof(int) * KP, st_), "memset");
  return refresh_x(X, x_on_device);
}

std::string EmgmmEngine::plan() const {
  char buf[384];
  const Geom& g = geom_;
  if (use16_)
    snprintf(buf, sizeof buf, "k_em16<%d>: fused single sweep, 16 warps/SM, log-densities as the expanded quadratic form (%s) or by whitening, split tile barrier, "
             "covariance-scatter accumulators in tensor memory%s", g.DM, qf16_ ? "guarded by the magnitude bound of its terms" : "disabled",
             spl_ > 1 ? ", statistics of a component block shared by several warps" : "");
  else if (used_)
    snprintf(buf, sizeof buf, "k_emd: fused single sweep for diagonal covariances on the fp64 CUDA cores (lane = component)");
  else if (fused_)
    snprintf(buf, sizeof buf, "k_em<%d,%d,true,%d,%s>: fused single sweep, 8 warps/SM, %d-point tiles", g.DM, fgeom_.CBMAX, fgeom_.GPW,
             fgeom_.diag ? "diag" : "full", fgeom_.TP);
  else if (big_)
    snprintf(buf, sizeof buf, "k_estep_big<%d,%d> + k_mstep_big<%d>: two-pass, Cholesky fragments %s", g.DM, (g.K8 + 7) / 8, g.DN,
             estep_big_stage(g, smem_limit_) ? "staged in shared memory one component ahead (cp.async, two buffers)" : "streamed through L1");
  else if (legacy_estep_)
    snprintf(buf, sizeof buf, "k_estep<%d> + k_mstep<%d>: first-generation two-pass kernels", g.DM, g.DN);
  else
    snprintf(buf, sizeof buf, "k_em<%d,%d,false> + k_mstep<%d>: two-pass (scatter accumulators exceed the register file)", g.DM, fgeom_.CBMAX, g.DN);
  return buf;
}

int EmgmmEngine::ensure_gamma() {
  if (dGamma_) return 0;
  CKE(cudaMalloc((void**)&dGamma_, std::max<size_t>((size_t)geom_.K * ldg_, 1) * sizeof(double)), "cudaMalloc gamma (N*K responsibilities)");
  CKE(cudaMemsetAsync(dGamma_, 0, (size_t)geom_.K * ldg_ * sizeof(double), st_), "memset gamma");
  return 0;
}

int EmgmmEngine::ensure_amb_scratch() {
  if (dAmbPart_) return 0;
  const int d = geom_.d;
  CKE(cudaMalloc((void**)&dAmbComp_, (size_t)amb_lists_ * (amb_cmax_ + 1) * sizeof(int)), "cudaMalloc side-list scratch");
  CKE(cudaMalloc((void**)&dAmbPart_, (size_t)amb_lists_ * amb_cmax_ * (d * d + d + 1) * sizeof(double)), "cudaMalloc side-list scratch");
  return 0;
}

// A sweep overflowed a side list: four times the capacity for the sweeps to come (the current iteration is repeated with
// the exact threshold, which lists nothing).
int EmgmmEngine::grow_amb_lists(int64_t longest) {
  if (amb_cap_ >= kAmbCapMax) return 0;
  // room for the longest list this rank just saw (+25 %), at least four times the old capacity; never more than a quarter
  // of the free device memory or 8 GB
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = 0; }
  const int64_t budget = (int64_t)std::min<size_t>(fr / 4, (size_t)8 << 30) / ((int64_t)amb_lists_ * (int64_t)sizeof(AmbEntry));
  int64_t want = std::max<int64_t>((int64_t)amb_cap_ * 4, longest + longest / 4 + 64);
  want = std::min<int64_t>(std::min<int64_t>(want, kAmbCapMax), budget);
  if (want <= amb_cap_) return 0;
  const int cap = (int)want;
  AmbEntry* p = nullptr;
  if (cudaMalloc((void**)&p, (size_t)amb_lists_ * cap * sizeof(AmbEntry)) != cudaSuccess) {
    cudaGetLastError();   // not fatal: the protocol falls back to the repeated sweep
    return 0;
  }
  CKE(cudaStreamSynchronize(st_), "sync");
  cudaFree(dAmb_);
  dAmb_ = p;
  amb_cap_ = cap;
  return 0;
}

// InverseSymMatrix eigen branch + packing for both kernel generations (k_refresh).
int EmgmmEngine::do_refresh(bool with_sumll) {
  QfRefresh q{};
  if (qf_) { q.packq = dPackQ_; q.qfS = dQfS_; q.rad = rad_ready_ ? dRad_ : nullptr; q.nrad = rad_world_; }
  CKE(launch_refresh(geom_, dMu_, dCov_, dLambda_, dSc_, dMean_, (double)nglob_, dScratch_, dPack_, dSumLLk_, dInfo_, st_, dPack2_, dMuOld_,
                     dShift_, with_sumll ? 1 : 0, cov_diag_ ? 1 : 0, &fgeom_, (geom_.d > 32 && refresh_count_ % 8 != 0) ? 1 : 0, qf_ ? &q : nullptr, dVValid_), "refresh");
  // eigen branch at d > 32: warm-started from the previous eigenvectors, every 8th refresh restarts them from the identity (bounds
  // their orthogonality drift).  At d <= 32 the sweeps are cheap and always start cold, so that a restored object continues bit for
  // bit like the uninterrupted one whichever branch its covariances take.
  ++refresh_count_;
  ++launches_;
  return 0;
}

int EmgmmEngine::refresh_x(const double* X, bool x_on_device) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  const size_t bytes = (size_t)n_ * geom_.d * sizeof(double);
  pending_host_ = nullptr;
  if (x_on_device) {
    dX_ = X;
  } else {
    if (!dXown_) CKE(cudaMalloc((void**)&dXown_, std::max<size_t>(bytes, 8)), "cudaMalloc X");
    dX_ = dXown_;
    if (fused_ && shift_ready_ && n_ >= (int64_t)kStreamChunks * 4096) {
      pending_host_ = X;   // uploaded chunk by chunk underneath the next sweep (iterate_fused), not here
    } else if (bytes) {
      CKE(cudaMemcpyAsync(dXown_, X, bytes, cudaMemcpyHostToDevice, st_), "H2D X");
    }
  }
  moments_ready_ = false;
  rad_ready_ = false;
  if (!fused_) shift_ready_ = false;  // the shift is re-taken as the exact mean of the new data
  return 0;
}

int EmgmmEngine::flush_pending() {
  if (!pending_host_) return 0;
  const size_t bytes = (size_t)n_ * geom_.d * sizeof(double);
  CKE(cudaMemcpyAsync(dXown_, pending_host_, bytes, cudaMemcpyHostToDevice, st_), "H2D X");
  pending_host_ = nullptr;
  return 0;
}

int EmgmmEngine::allreduce(double* buf, int64_t count) {
  if (!ar_) return 0;
  const int r = ar_(ar_ctx_, buf, count, (void*)st_);
  if (r != 0) { err_ = "all-reduce callback failed"; return -7; }
  return 0;
}

// Global moments of the data: N, exact mean m and centred scatter Sc = sum_i (x_i-m)(x_i-m)^T, used by the closed form
// of sum_i maha_ik (sumLL, src/unsupervised/mlf_gaussian.f90:73-76).  The first call fixes the shift vector (the global
// mean at that time): the sweep whitens x - shift, and later moments are taken about the same shift in ONE pass.
int EmgmmEngine::ensure_moments() {
  if (moments_ready_) return 0;
  if (int r = flush_pending()) return r;
  const Geom& g = geom_;
  const int d = g.d;
  if (!shift_ready_) {
    const int grid = std::max(1, std::min<int>(sms_, (int)((n_ + 255) / 256)));
    CKE(launch_colsum(dX_, n_, d, dMomPart_, grid, st_), "colsum");
    CKE(launch_reduce_partials(dMomPart_, grid, d, dShift_, st_), "reduce colsum");
    launches_ += 2;
    const double nl = (double)n_;
    CKE(cudaMemcpyAsync(dShift_ + d, &nl, sizeof(double), cudaMemcpyHostToDevice, st_), "H2D n");
    CKE(cudaStreamSynchronize(st_), "sync");
    if (int r = allreduce(dShift_, d + 1)) return r;
    std::vector<double> h(8 * g.DN + 8, 0.0);
    CKE(cudaMemcpyAsync(h.data(), dShift_, (d + 1) * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H mean");
    CKE(cudaStreamSynchronize(st_), "sync");
    nglob_ = (int64_t)llround(h[d]);
    const double ng = h[d];
    for (int a = 0; a < d; ++a) h[a] /= ng;
    for (size_t a = d; a < h.size(); ++a) h[a] = 0.0;
    CKE(cudaMemcpyAsync(dShift_, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D shift");
    CKE(cudaStreamSynchronize(st_), "sync");
    shift_ready_ = true;
  }
  if (g.DN > 4) {
    // large d: scatter about the shift through the large-shape M-pass kernel with unit weights.  Without the fused
    // path the shift is re-taken at every refresh of X, so sum (x - shift) is zero by construction.
    Geom g1 = make_geom(d, 1, true);
    int gx;
    int64_t chunk;
    mstep_big_grid(g1, sms_, n_, &gx, &chunk);
    const double zero = 0.0;
    CKE(cudaMemcpyAsync(dThr_, &zero, sizeof(double), cudaMemcpyHostToDevice, st_), "H2D thr");
    CKE(launch_mstep_big(g1, dX_, n_, nullptr, 0, dShift_, dThr_, dPartB_, gx, chunk, st_), "moments (large shape)");
    CKE(launch_reduce_partials(dPartB_, gx, (int64_t)g1.MS, dMomRaw_, st_), "reduce moments");  // MS = MB*64 + 2 <= PM
    CKE(cudaMemsetAsync(dMomRaw_ + (size_t)g.MB * 64, 0, (size_t)8 * g.DN * sizeof(double), st_), "memset");
    launches_ += 2;
    const int PM = g.MB * 64 + 8 * g.DN;
    if (int r = allreduce(dMomRaw_, PM)) return r;
    CKE(launch_moments_finish(g, dMomRaw_, dShift_, (double)nglob_, dMean_, dSc_, st_), "moments finish");
    ++launches_;
    moments_ready_ = true;
    return 0;
  }
  const int grid = std::max(1, std::min<int>(sms_, (int)((n_ + 127) / 128)));
  CKE(launch_moments(g, dX_, n_, dShift_, dMomPart_, grid, 0, st_), "moments");
  ++launches_;
  return finish_moments(grid);
}

// HBM home of the responsibilities in statistics-first iterations ([tile][64][KP]); false (and the protocol keeps its second
// full sweep) when disabled (MLF_B200_GSTORE=0) or when the allocation does not fit.
bool EmgmmEngine::ensure_gtiles() {
  if (dGTiles_) return true;
  if (gtiles_failed_) return false;
  const char* e = getenv("MLF_B200_GSTORE");
  if (e && e[0] == '0') { gtiles_failed_ = true; return false; }
  const size_t ntiles = (size_t)((n_ + 63) / 64);
  const size_t bytes = std::max<size_t>(ntiles, 1) * 64 * (size_t)geom_.K8 * 8 * sizeof(double);
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = 0; }
  if (bytes + ((size_t)2 << 30) > fr || cudaMalloc((void**)&dGTiles_, bytes) != cudaSuccess) {
    cudaGetLastError();
    dGTiles_ = nullptr;
    gtiles_failed_ = true;   // keep at least 2 GB of head room for the caller; fall back to the repeated sweep
    return false;
  }
  return true;
}

// Radius of the resident data about the shift, per coordinate and rank (guard of the quadratic-form sweep).  COLLECTIVE like
// ensure_moments: the per-rank maxima are gathered through the sum all-reduce (every rank fills its own slot of a zeroed
// buffer), so that all ranks hold the same bounds and take the same decision.
int EmgmmEngine::ensure_radius() {
  if (!qf_ || rad_ready_) return 0;
  if (rad_world_ < world_) {
    if (dRad_) cudaFree(dRad_);
    dRad_ = nullptr;
    CKE(cudaMalloc((void**)&dRad_, (size_t)world_ * 16 * sizeof(double)), "cudaMalloc");
    rad_world_ = world_;
  }
  CKE(cudaMemsetAsync(dRad_, 0, (size_t)rad_world_ * 16 * sizeof(double), st_), "memset");
  const int grid = std::max(1, std::min<int>(sms_, (int)((n_ + 255) / 256)));
  CKE(launch_absmax(dX_, n_, geom_.d, dShift_, dRadPart_, grid, dRad_ + (size_t)rank_ * 16, st_), "absmax");
  launches_ += 2;
  if (int r = allreduce(dRad_, (int64_t)rad_world_ * 16)) return r;
  rad_ready_ = true;
  return 0;
}

int EmgmmEngine::prepare() {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  return ensure_moments();
}

// M-pass dispatch: thresholded centred scatter of every component into dStatsB_ (fixed-order reduction over chunks).
int EmgmmEngine::run_mstep(const double* gamma) {
  const Geom& g = geom_;
  if (big_) {
    CKE(launch_mstep_big(g, dX_, n_, gamma, ldg_, dMupad_, dThr_, dPartB_, mgx_, mchunk_, st_), "mstep (large shape)");
  } else {
    CKE(launch_mstep(g, dX_, n_, gamma, ldg_, dMupad_, dThr_, dPartB_, mgx_, st_), "mstep");
  }
  CKE(launch_reduce_partials(dPartB_, mgx_, (int64_t)g.K * g.MS, dStatsB_, st_), "reduce statsB");
  launches_ += 2;
  return 0;
}

int EmgmmEngine::finish_moments(int grid) {
  const Geom& g = geom_;
  const int PM = g.MB * 64 + 8 * g.DN;
  CKE(launch_reduce_partials(dMomPart_, grid, PM, dMomRaw_, st_), "reduce moments");
  if (int r = allreduce(dMomRaw_, PM)) return r;
  CKE(launch_moments_finish(g, dMomRaw_, dShift_, (double)nglob_, dMean_, dSc_, st_), "moments finish");
  launches_ += 2;
  moments_ready_ = true;
  return 0;
}

int EmgmmEngine::upload_params(const double* Mu, const double* Cov, const double* lambda) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  const int d = geom_.d, K = geom_.K;
  CKE(cudaMemcpyAsync(dMu_, Mu, (size_t)K * d * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D Mu");
  CKE(cudaMemcpyAsync(dCov_, Cov, (size_t)K * d * d * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D Cov");
  CKE(cudaMemcpyAsync(dLambda_, lambda, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D lambda");
  CKE(cudaStreamSynchronize(st_), "sync");  // host mirrors may be pageable and change after we return
  return 0;
}

int EmgmmEngine::download_params(double* Mu, double* Cov, double* lambda) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  const int d = geom_.d, K = geom_.K;
  CKE(cudaMemcpyAsync(Mu, dMu_, (size_t)K * d * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H Mu");
  CKE(cudaMemcpyAsync(Cov, dCov_, (size_t)K * d * d * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H Cov");
  CKE(cudaMemcpyAsync(lambda, dLambda_, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H lambda");
  CKE(cudaStreamSynchronize(st_), "sync");
  return 0;
}

int EmgmmEngine::run_estep(const double* X, int64_t n, double* gamma, int64_t ldg, int32_t* labels, bool want_stats, bool raw, bool hard) {
  const Geom& g = geom_;
  const int tp = legacy_estep_ ? kEstepTilePts : fgeom_.TP;
  const int64_t etiles = (n + tp - 1) / tp;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, etiles));
  if (big_) {
    EmArgs a{};
    a.X = X; a.N = n; a.d = g.d; a.K = g.K; a.K8 = g.K8; a.S = g.S;
    a.pack2 = dPack2_; a.shift = dShift_; a.gamma = gamma; a.ldg = ldg; a.labels = labels; a.raw = raw ? 1 : 0; a.partA = dPartA_;
    a.hard = hard ? 1 : 0;
    const int bgrid = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, (n + estep_big_tile(geom_, smem_limit_) - 1) / estep_big_tile(geom_, smem_limit_)));
    CKE(launch_estep_big(g, a, bgrid, smem_limit_, st_), "estep (large shape)");
    ++launches_;
    if (want_stats) {
      CKE(launch_reduce_partials(dPartA_, bgrid, (int64_t)g.K8 * 8 * g.DS, dStatsA_, st_), "reduce statsA");
      ++launches_;
    }
    return 0;
  }
  if (legacy_estep_) {
    CKE(launch_estep(g, X, n, dPack_, gamma, ldg, labels, dPartA_, grid, smem_limit_, raw ? 1 : 0, st_), "estep");
  } else {
    EmArgs a{};
    a.X = X; a.N = n; a.d = g.d; a.K = g.K; a.K8 = g.K8; a.S = g.S;
    a.pack2 = dPack2_; a.shift = dShift_; a.muold = dMuOld_; a.tlo = nullptr; a.thi = nullptr;
    a.gamma = gamma; a.ldg = ldg; a.labels = labels; a.raw = raw ? 1 : 0;
    a.partA = dPartA_; a.partB = nullptr; a.amb = nullptr; a.amb_cap = 0; a.amb_count = nullptr;
    a.accumulate = 0; a.p_offset = 0; a.hard = hard ? 1 : 0;
    CKE(launch_em(g, fgeom_, a, false, grid, smem_limit_, st_), "em sweep (E-pass)");
  }
  ++launches_;
  if (want_stats) {
    CKE(launch_reduce_partials(dPartA_, grid, (int64_t)g.K8 * 8 * g.DS, dStatsA_, st_), "reduce statsA");
    ++launches_;
  }
  return 0;
}

int EmgmmEngine::iterate(int64_t niter, double minProba, double* sumLL_last, double* sumLL_trace) {
  if (niter <= 0) return 0;
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (!fused_) {
    if (int r = flush_pending()) return r;
  }
  if (!(fused_ && pending_host_ && shift_ready_))
    if (int r = ensure_moments()) return r;
  if (trace_cap_ < niter) {
    if (dSumLL_) cudaFree(dSumLL_);
    trace_cap_ = std::max<int64_t>(niter, 64);
    CKE(cudaMalloc((void**)&dSumLL_, trace_cap_ * sizeof(double)), "cudaMalloc trace");
  }
  const bool prof = profile_ && niter <= 4096;
  if (prof) {
    while ((int64_t)pev_.size() < 6 * niter) {
      cudaEvent_t e;
      CKE(cudaEventCreate(&e), "cudaEventCreate");
      pev_.push_back(e);
    }
  }
  if (int r = fused_ ? iterate_fused(niter, minProba, prof) : iterate_two_pass(niter, minProba, prof)) return r;
  std::vector<double> tr((size_t)niter);
  CKE(cudaMemcpyAsync(tr.data(), dSumLL_, niter * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H sumLL");
  CKE(cudaStreamSynchronize(st_), "sync");
  if (prof) {  // read the per-iteration events only after the final sync
    for (int64_t it = 0; it < niter; ++it) {
      cudaEvent_t* e = &pev_[(size_t)6 * it];
      float ms;
      cudaEventElapsedTime(&ms, e[0], e[1]); times_.refresh += ms;
      cudaEventElapsedTime(&ms, e[1], e[2]); times_.estep += ms;
      cudaEventElapsedTime(&ms, e[2], e[3]); times_.mid += ms;
      cudaEventElapsedTime(&ms, e[3], e[4]); times_.mstep += ms;
      cudaEventElapsedTime(&ms, e[4], e[5]); times_.fin += ms;
      times_.iters += 1;
    }
  }
  if (sumLL_last) *sumLL_last = tr[(size_t)niter - 1];
  if (sumLL_trace) std::memcpy(sumLL_trace, tr.data(), niter * sizeof(double));
  return 0;
}

// Two-pass iteration with the responsibilities stored in HBM: E-pass -> all-reduce -> thresholds -> M-pass -> all-reduce.
// Used where the fused sweep's scatter accumulators do not fit the register file (d > 16, or d > 8 with K > 64).
int EmgmmEngine::iterate_two_pass(int64_t niter, double minProba, bool prof) {
  const Geom& g = geom_;
  if (int r = ensure_gamma()) return r;
  const double minW = 1e-3 / (double)(float)nglob_;  // 1d-3/real(NX)  (src/unsupervised/mlf_gaussian.f90:50)
  for (int64_t it = 0; it < niter; ++it) {
    cudaEvent_t* ev_ = prof ? &pev_[(size_t)6 * it] : nullptr;
    if (prof) cudaEventRecord(ev_[0], st_);
    if (int r = do_refresh()) return r;
    CKE(launch_sum_k(dSumLLk_, g.K, dSumLL_ + it, st_), "sum_k");
    ++launches_;
    if (prof) cudaEventRecord(ev_[1], st_);
    if (int r = run_estep(dX_, n_, dGamma_, ldg_, nullptr, true)) return r;
    if (int r = allreduce(dStatsA_, (int64_t)g.K8 * 8 * g.DS)) return r;
    if (prof) cudaEventRecord(ev_[2], st_);
    CKE(launch_mid(g, dStatsA_, minW, dMu_, dMupad_, dThr_, st_), "mid");
    ++launches_;
    if (prof) cudaEventRecord(ev_[3], st_);
    if (int r = run_mstep(dGamma_)) return r;
    if (int r = allreduce(dStatsB_, (int64_t)g.K * g.MS)) return r;
    if (prof) cudaEventRecord(ev_[4], st_);
    CKE(launch_finalize(g, dStatsA_, dStatsB_, minProba, dCov_, dLambda_, st_, cov_diag_ ? 1 : 0), "finalize");
    ++launches_;
    if (prof) cudaEventRecord(ev_[5], st_);
  }
  gamma_current_ = true;
  return 0;
}

// First sweep after refresh_x(host pointer): the shard is uploaded in kStreamChunks pieces on a copy stream while the
// compute stream sweeps the pieces already there; every piece adds to the same per-block partials (fixed order, so the
// result is deterministic) and to the global moments about the fixed shift.
int EmgmmEngine::streamed_sweep(const EmArgs& base) {
  const Geom& g = geom_;
  const FusedGeom& f = fgeom_;
  const int d = g.d, KP = g.K8 * 8;
  const int PM = g.MB * 64 + 8 * g.DN;
  const int nlists = sms_ * 16;
  while ((int)cev_.size() < kStreamChunks + 1) {
    cudaEvent_t e;
    CKE(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    cev_.push_back(e);
  }
  CKE(cudaMemsetAsync(dPartA_, 0, (size_t)2 * sms_ * spl_ * KP * g.DS * sizeof(double), st_), "memset");
  CKE(cudaMemsetAsync(dPartB2_, 0, (size_t)2 * sms_ * spl_ * KP * f.MSF * sizeof(double), st_), "memset");
  CKE(cudaMemsetAsync(dAmbCount_, 0, (size_t)nlists * sizeof(int), st_), "memset");
  CKE(cudaMemsetAsync(dMomPart_, 0, (size_t)sms_ * PM * sizeof(double), st_), "memset");
  // the copy may not overwrite X before everything queued so far (earlier sweeps) has read it
  CKE(cudaEventRecord(cev_[kStreamChunks], st_), "event");
  CKE(cudaStreamWaitEvent(cst_, cev_[kStreamChunks], 0), "wait");
  int64_t chunk = (n_ + kStreamChunks - 1) / kStreamChunks;
  chunk = ((chunk + kEstepTilePts - 1) / kEstepTilePts) * kEstepTilePts;
  int c = 0;
  for (int64_t off = 0; off < n_; off += chunk, ++c) {
    const int64_t n = std::min(chunk, n_ - off);
    CKE(cudaMemcpyAsync(dXown_ + (size_t)off * d, pending_host_ + (size_t)off * d, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, cst_),
        "H2D X chunk");
    CKE(cudaEventRecord(cev_[c], cst_), "event");
    CKE(cudaStreamWaitEvent(st_, cev_[c], 0), "wait");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, (n + f.TP - 1) / f.TP));
    CKE(launch_moments(g, dXown_ + (size_t)off * d, n, dShift_, dMomPart_, grid, 1, st_), "moments chunk");
    EmArgs a = base;
    a.X = dXown_ + (size_t)off * d;
    a.N = n;
    a.accumulate = 1;
    a.p_offset = off;
    if (use16_) CKE(launch_em16(g, f, a, grid, smem_limit_, st_), "em sweep (streamed chunk, 16 warps)");
    else if (used_) CKE(launch_emd(g, f, a, grid, smem_limit_, st_), "em sweep (streamed chunk, diagonal)");
    else CKE(launch_em(g, f, a, true, grid, smem_limit_, st_), "em sweep (streamed chunk)");
    launches_ += 2;
  }
  pending_host_ = nullptr;
  return finish_moments(sms_);
}

// Single-sweep iteration (emgmm_fused.cu).  Per iteration: refresh -> one sweep over X that produces the pass-A
// statistics AND the covariance scatter of the certainly-kept pairs -> ONE all-reduce of both -> host check that the
// finished s_k lies inside the predicted threshold band (repeat the sweep with the exact threshold otherwise) ->
// side-list resolution (+ a second, tiny all-reduce only if any pair was listed) -> finalize.
int EmgmmEngine::iterate_fused(int64_t niter, double minProba, bool prof) {
  const Geom& g = geom_;
  const FusedGeom& f = fgeom_;
  const int K = g.K, KP = g.K8 * 8, d = g.d;
  const int nlists = egrid_ * (use16_ ? 16 : kEstepThreads / 32);
  const int nslots = use16_ ? 2 * egrid_ * spl_ : egrid_;
  const int64_t nA = (int64_t)KP * g.DS, nB = (int64_t)KP * f.MSF;
  const double minW = 1e-3 / (double)(float)nglob_;  // 1d-3/real(NX)  (src/unsupervised/mlf_gaussian.f90:50)
  double* dFlags = dAB_ + nA + nB;                    // [listed pairs, overflowing lists] ride in the same all-reduce
  std::vector<double> hQ((size_t)KP);
  std::vector<double> hA((size_t)nA + 0), hflag(5), tlo((size_t)KP), thi((size_t)KP), plo((size_t)KP), phi((size_t)KP), sk((size_t)K);
  gamma_current_ = false;
  for (int64_t it = 0; it < niter; ++it) {
    cudaEvent_t* ev_ = prof ? &pev_[(size_t)6 * it] : nullptr;
    if (prof) cudaEventRecord(ev_[0], st_);
    // parameters this E-step is evaluated with (what ProbaC refers to after the step)
    CKE(cudaMemcpyAsync(dPrevMu_, dMu_, (size_t)K * d * sizeof(double), cudaMemcpyDeviceToDevice, st_), "copy");
    CKE(cudaMemcpyAsync(dPrevCov_, dCov_, (size_t)K * d * d * sizeof(double), cudaMemcpyDeviceToDevice, st_), "copy");
    CKE(cudaMemcpyAsync(dPrevLambda_, dLambda_, (size_t)K * sizeof(double), cudaMemcpyDeviceToDevice, st_), "copy");
    prev_valid_ = true;
    // X was refreshed from host memory: its upload runs chunk by chunk underneath this iteration's first sweep, and the
    // global moments (hence sumLL) become final only after it
    const bool streamed = pending_host_ != nullptr;
    if (!streamed)
      if (int r = ensure_radius()) return r;
    if (int r = do_refresh(!streamed)) return r;
    if (!streamed) {
      CKE(launch_sum_k(dSumLLk_, K, dSumLL_ + it, st_), "sum_k");
      ++launches_;
    }
    if (prof) cudaEventRecord(ev_[1], st_);
    // ---- threshold band [plo, phi] = minW * (c -+ hw): s_k = sum_i gamma_ik predicted from the previous iterations ------
    // In its smooth phase EM moves s_k steadily, so the last step is extrapolated (centre s + ds) and the half-width
    // follows the CHANGE of the step (second difference): the band must hold the finished s_k, and every pair whose
    // responsibility falls inside it costs a side-list entry.  In the violent phase (overlapping clusters under the
    // reference's determinant sign, SURVEY fact 0.2: components halve or double per iteration) nothing predicts s_k; then
    // the iteration is run as "statistics first": pass 0 with infinite thresholds (E-step + pass-A statistics only: no
    // scatter, no list), pass 1 with the exact thresholds.  The prediction is still scored against the finished s_k, and
    // the single-sweep mode returns after it would have held.
    bool smooth = !s_prev2_.empty();
    for (int k = 0; k < KP; ++k) {
      double lo, hi;
      if (k >= K) { lo = hi = INFINITY; }
      else {
        double c, hw;
        if (s_prev_.empty()) { c = (double)nglob_ / K; hw = 0.5 * c; }
        else {
          const double s1 = s_prev_[(size_t)k];
          if (s_prev2_.empty()) { c = s1; hw = 0.25 * std::fabs(s1); }
          else {
            const double ds = s1 - s_prev2_[(size_t)k];
            if (!(std::fabs(ds) <= 0.05 * std::fabs(s1))) smooth = false;
            c = s1 + ds;
            if (s_prev3_.empty()) hw = 2.0 * std::fabs(ds);
            else {
              const double dds = ds - (s_prev2_[(size_t)k] - s_prev3_[(size_t)k]);
              hw = std::max(3.0 * std::fabs(dds), 0.125 * std::fabs(ds));
            }
          }
          hw = std::max(hw, 1e-7 * std::fabs(c));
          if (band_widen_.size() == (size_t)K) hw *= band_widen_[(size_t)k];
          hw = std::min(hw, 0.6 * std::fabs(c));
        }
        lo = minW * std::max(c - hw, 0.0);
        hi = minW * (c + hw);
      }
      plo[(size_t)k] = lo; phi[(size_t)k] = hi;
    }
    const bool stats_first = !s_prev2_.empty() && (!smooth || band_misses_ > 0);
    // statistics-first iterations of the 16-warp sweep keep the responsibilities of pass 0 in HBM (8 N KP bytes, allocated at
    // first need; 51 GB at N=1e8, K=64) so that pass 1 is a scatter from stored tiles instead of a second full sweep
    const bool gstore = stats_first && !streamed && use16_ && ensure_gtiles();
    bool gvote = false;   // every rank of the job stored its responsibilities in pass 0 (all-reduced vote)
    bool exact = false;
    for (int pass = 0; pass < 2; ++pass) {
      if (gstore && pass == 1 && gvote) {
        for (int k = 0; k < KP; ++k) thi[(size_t)k] = k < K ? minW * sk[(size_t)k] : INFINITY;
        CKE(cudaMemcpyAsync(dThi_, thi.data(), KP * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D thi");
        EmArgs a{};
        a.X = dX_; a.N = n_; a.d = d; a.K = K; a.K8 = g.K8; a.S = g.S;
        a.muold = dMuOld_; a.thi = dThi_; a.partB = dPartB2_; a.gtiles = dGTiles_; a.qf_limit = -1.0; a.spl = spl_;
        CKE(launch_scatter16(g, f, a, egrid_, smem_limit_, st_), "scatter pass (stored responsibilities)");
        CKE(launch_reduce_partials(dPartB2_, nslots, nB, dAB_ + nA, st_), "reduce statsB");
        CKE(launch_amb_summary(dAmbCount_, nlists, amb_cap_, dFlags, st_, dAB_ + nA, KP, f.MSF, g.MB * 64 + 8 * g.DN + 1), "amb summary");
        launches_ += 3;
        times_.scatter_passes += 1;
        if (int r = allreduce(dAB_ + nA, nB + 4)) return r;
        CKE(cudaMemcpyAsync(hflag.data(), dFlags, 5 * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H flags");
        CKE(cudaStreamSynchronize(st_), "sync");
        if (debug_ && rank_ == 0)
          fprintf(stderr, "mlfortran_b200[debug]: iteration %lld pass 1 (scatter from stored responsibilities, exact thresholds): kept pairs %.0f -> accepted\n",
                  (long long)it, hflag[2]);
        break;
      }
      for (int k = 0; k < KP; ++k) {
        if (k >= K || (pass == 0 && stats_first)) tlo[(size_t)k] = thi[(size_t)k] = INFINITY;
        else if (exact) tlo[(size_t)k] = thi[(size_t)k] = minW * sk[(size_t)k];
        else { tlo[(size_t)k] = plo[(size_t)k]; thi[(size_t)k] = phi[(size_t)k]; }
      }
      CKE(cudaMemcpyAsync(dTlo_, tlo.data(), KP * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D tlo");
      CKE(cudaMemcpyAsync(dThi_, thi.data(), KP * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D thi");
      EmArgs a{};
      a.X = dX_; a.N = n_; a.d = d; a.K = K; a.K8 = g.K8; a.S = g.S;
      a.pack2 = dPack2_; a.shift = dShift_; a.muold = dMuOld_; a.tlo = dTlo_; a.thi = dThi_;
      a.gamma = nullptr; a.ldg = 0; a.labels = nullptr; a.raw = 0;
      a.partA = dPartA_; a.partB = dPartB2_; a.amb = dAmb_; a.amb_cap = amb_cap_; a.amb_count = dAmbCount_;
      a.accumulate = 0; a.p_offset = 0;
      // quadratic-form logits when the bounds of this refresh allow (decided by the kernels); never while X is still arriving
      a.packq = dPackQ_; a.qfS = dQfS_; a.qf_limit = (qf_ && rad_ready_ && !(streamed && pass == 0)) ? qf_limit_ : -1.0;
      if (streamed && pass == 0) {
        if (int r = streamed_sweep(a)) return r;
        CKE(launch_sumll(g, dMu_, dSc_, dMean_, (double)nglob_, dScratch_, dPack_, dSumLLk_, st_), "sumll");
        CKE(launch_sum_k(dSumLLk_, K, dSumLL_ + it, st_), "sum_k");
        launches_ += 2;
      } else {
        a.gtiles = dGTiles_; a.spl = spl_;
        if (use16_) CKE(launch_em16(g, f, a, egrid_, smem_limit_, st_, gstore && pass == 0), "em sweep (fused, 16 warps)");
        else if (used_) CKE(launch_emd(g, f, a, egrid_, smem_limit_, st_), "em sweep (fused, diagonal)");
        else CKE(launch_em(g, f, a, true, egrid_, smem_limit_, st_), "em sweep (fused)");
        // with an expanded-form variant both variants of the sweep are launched and one of them returns at once
        launches_ += ((use16_ || used_) && a.qf_limit >= 0.0) ? 2 : 1;
      }
      CKE(launch_reduce_partials(dPartA_, nslots, nA, dAB_, st_), "reduce statsA");
      CKE(launch_reduce_partials(dPartB2_, nslots, nB, dAB_ + nA, st_), "reduce statsB");
      // the vote rides in the all-reduce: the scatter pass replaces the second sweep only if EVERY rank stored its responsibilities
      CKE(launch_amb_summary(dAmbCount_, nlists, amb_cap_, dFlags, st_, dAB_ + nA, KP, f.MSF, g.MB * 64 + 8 * g.DN + 1, (gstore && pass == 0) ? 1.0 : 0.0),
          "amb summary");
      launches_ += 3;
      times_.sweeps += 1;
      if (int r = allreduce(dAB_, nA + nB + 4)) return r;
      // verification needs s_k and the list flags on the host: one small D2H + sync per sweep
      CKE(cudaMemcpyAsync(hA.data(), dAB_, nA * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H statsA");
      CKE(cudaMemcpyAsync(hflag.data(), dFlags, 5 * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H flags");
      if (a.qf_limit >= 0.0) CKE(cudaMemcpyAsync(hQ.data(), dQfS_, KP * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H bounds");
      CKE(cudaStreamSynchronize(st_), "sync");
      if (a.qf_limit >= 0.0) {   // the decision the blocks took (same rule), for the counters only
        bool bad = false;
        double mx = 0;
        for (int k = 0; k < KP; ++k) { bad |= !(hQ[(size_t)k] <= a.qf_limit); mx = std::fmax(mx, hQ[(size_t)k]); if (hQ[(size_t)k] != hQ[(size_t)k]) mx = NAN; }
        times_.qf_bound = mx;
        if (!bad) times_.qf_sweeps += 1;
      }
      if (pass == 0) gvote = hflag[3] == (double)world_;
      bool ok = hflag[1] == 0.0 && !(pass == 0 && stats_first);
      int missed = 0;
      if (!exact && band_widen_.size() != (size_t)K) band_widen_.assign((size_t)K, 1.0);
      for (int k = 0; k < K; ++k) {
        sk[(size_t)k] = hA[(size_t)k * g.DS];
        if (exact) continue;
        // the prediction is scored in both modes (in statistics-first mode in hindsight)
        const double tau = minW * sk[(size_t)k];
        const bool hit = plo[(size_t)k] <= tau && tau <= phi[(size_t)k];   // false for NaN as well
        if (!hit) { ok = false; ++missed; }
        band_widen_[(size_t)k] = hit ? std::max(1.0, 0.5 * band_widen_[(size_t)k]) : std::min(256.0, 4.0 * band_widen_[(size_t)k]);
      }
      if (!exact) band_misses_ = missed > 0 ? 2 : std::max(0, band_misses_ - 1);   // two clean iterations re-arm the single sweep
      if (debug_ && rank_ == 0)
        fprintf(stderr, "mlfortran_b200[debug]: iteration %lld pass %d (%s): listed pairs %.0f, overflowed lists %.0f (cap %d), kept pairs %.0f, "
                "components outside the predicted band %d%s -> %s\n", (long long)it, pass,
                exact ? "exact thresholds" : (stats_first ? "statistics only" : "predicted band"), hflag[0], hflag[1], amb_cap_, hflag[2], missed,
                smooth ? "" : " (not smooth)", (ok || exact) ? "accepted" : "second pass with exact thresholds");
      if (!ok && !exact && hflag[1] != 0.0)
        if (int r = grow_amb_lists((int64_t)hflag[4])) return r;
      if (ok || exact) break;
      exact = true;  // s_k is known now: repeat the sweep with tlo == thi == minW*s_k (the side list stays empty)
    }
    if (prof) cudaEventRecord(ev_[2], st_);
    if (prof) cudaEventRecord(ev_[3], st_);
    const double* statsC = nullptr;
    if (hflag[0] > 0.0) {
      if (int r = ensure_amb_scratch()) return r;
      CKE(launch_amb_resolve(g, dX_, dAmb_, amb_cap_, dAmbCount_, nlists, dAB_, minW, dMuOld_, dStatsC_, st_, amb_cmax_, dAmbComp_, dAmbPart_),
          "amb resolve");
      launches_ += 2;
      if (int r = allreduce(dStatsC_, (int64_t)K * (d * d + d + 1))) return r;
      statsC = dStatsC_;
      times_.amb_pairs += (int64_t)hflag[0];
    }
    if (prof) cudaEventRecord(ev_[4], st_);
    CKE(launch_finalize_fused(g, f, dAB_, dAB_ + nA, statsC, dMuOld_, minProba, dMu_, dCov_, dLambda_, st_, cov_diag_ ? 1 : 0), "finalize");
    ++launches_;
    if (prof) cudaEventRecord(ev_[5], st_);
    times_.kept_pairs += (int64_t)hflag[2];
    s_prev3_ = s_prev2_;
    s_prev2_ = s_prev_;
    s_prev_ = sk;
  }
  return 0;
}

namespace {
// host copy on a few threads: a single thread moves ~10 GB/s, the PCIe link behind it 55
void parallel_rows_copy(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t width, size_t rows) {
  const size_t total = width * rows;
  const int nt = total < (8u << 20) ? 1 : (total < (64u << 20) ? 4 : 8);
  auto work = [&](int t) {
    if (rows == 1) {
      const size_t lo = total * t / nt, hi = total * (t + 1) / nt;
      std::memcpy(dst + lo, src + lo, hi - lo);
      return;
    }
    for (size_t r = rows * t / nt; r < rows * (t + 1) / nt; ++r) std::memcpy(dst + r * dst_pitch, src + r * src_pitch, width);
  };
  if (nt == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}
bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
}  // namespace

void EmgmmEngine::free_infer_buffers() {
  for (InferBuf& b : inf_) {
    if (b.dX) cudaFree(b.dX);
    if (b.dP) cudaFree(b.dP);
    if (b.dL) cudaFree(b.dL);
    if (b.hX) cudaFreeHost(b.hX);
    if (b.hP) cudaFreeHost(b.hP);
    if (b.hL) cudaFreeHost(b.hL);
    b.dX = b.dP = b.hX = b.hP = nullptr;
    b.dL = b.hL = nullptr;
  }
  inf_chunk_ = 0;
  inf_hasP_ = inf_hasL_ = inf_hasHX_ = false;
}

// Persistent buffers of the streamed inference path: allocated at first use, kept for the object's lifetime, regrown only
// when a call needs something the previous ones did not (responsibilities after labels-only calls, a larger chunk).
int EmgmmEngine::ensure_infer_buffers(int64_t chunk, bool wantP, bool wantL, bool stageX) {
  const int d = geom_.d, K = geom_.K;
  if (chunk <= inf_chunk_ && (!wantP || inf_hasP_) && (!wantL || inf_hasL_) && (!stageX || inf_hasHX_)) return 0;
  wantP |= inf_hasP_; wantL |= inf_hasL_; stageX |= inf_hasHX_;
  chunk = std::max(chunk, inf_chunk_);
  CKE(cudaStreamSynchronize(st_), "sync");
  free_infer_buffers();
  if (!dst_) CKE(cudaStreamCreateWithFlags(&dst_, cudaStreamNonBlocking), "cudaStreamCreate");
  const int64_t ldq = ((chunk + 31) / 32) * 32;
  for (InferBuf& b : inf_) {
    CKE(cudaMalloc((void**)&b.dX, (size_t)chunk * d * sizeof(double)), "cudaMalloc inference X chunk");
    if (stageX) CKE(cudaMallocHost((void**)&b.hX, (size_t)chunk * d * sizeof(double)), "cudaMallocHost inference X staging");
    if (wantP) {
      CKE(cudaMalloc((void**)&b.dP, (size_t)K * ldq * sizeof(double)), "cudaMalloc inference P chunk");
      CKE(cudaMallocHost((void**)&b.hP, (size_t)K * chunk * sizeof(double)), "cudaMallocHost inference P staging");
    }
    if (wantL) {
      CKE(cudaMalloc((void**)&b.dL, (size_t)chunk * sizeof(int32_t)), "cudaMalloc inference labels chunk");
      CKE(cudaMallocHost((void**)&b.hL, (size_t)chunk * sizeof(int32_t)), "cudaMallocHost inference labels staging");
    }
    if (!b.in) {
      CKE(cudaEventCreateWithFlags(&b.in, cudaEventDisableTiming), "cudaEventCreate");
      CKE(cudaEventCreateWithFlags(&b.comp, cudaEventDisableTiming), "cudaEventCreate");
      CKE(cudaEventCreateWithFlags(&b.out, cudaEventDisableTiming), "cudaEventCreate");
    }
  }
  inf_chunk_ = chunk; inf_hasP_ = wantP; inf_hasL_ = wantL; inf_hasHX_ = stageX;
  return 0;
}

// ExpStep on caller data (getProba / getClass, src/mlf_models.f90:239-249,302-313; src/unsupervised/mlf_emgmm.f90:161-171).
// The query is streamed through two persistent chunk buffers on three streams: while chunk c is evaluated, chunk c+1 is on
// its way to the device and the results of chunk c-1 on their way back.  Pinned caller memory is copied directly; pageable
// memory goes through the object's own pinned staging buffers (host copies on four threads), so the PCIe link, not a
// synchronous pageable copy, bounds the call.
int EmgmmEngine::expectation(const double* Xq, bool on_device, int64_t nq, double* P_host, int32_t* labels_host) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (nq <= 0) return 0;
  if (int r = ensure_moments()) return r;
  const Geom& g = geom_;
  const int d = g.d, K = g.K;
  if (int r = do_refresh()) return r;
  // chunk: ~128 MB of whichever of input / output is larger per point, at most 2^20 points
  const int64_t per_pt = 8 * (int64_t)std::max(d, P_host ? K : 1);
  int64_t chunk = std::min<int64_t>((int64_t)1 << 20, std::max<int64_t>(4096, ((int64_t)1 << 27) / per_pt));
  if (const char* ic = getenv("MLF_B200_INFER_CHUNK")) chunk = std::max<int64_t>(32, atoll(ic));   // tests: many small chunks
  chunk = std::min<int64_t>(((nq + 31) / 32) * 32, (chunk / 32) * 32);
  const bool x_direct = on_device || is_pinned(Xq);
  const bool p_direct = P_host && is_pinned(P_host);
  const bool l_direct = labels_host && is_pinned(labels_host);
  if (int r = ensure_infer_buffers(chunk, P_host != nullptr, labels_host != nullptr, !x_direct)) return r;
  chunk = std::min(chunk, inf_chunk_);
  const int64_t ldq = ((inf_chunk_ + 31) / 32) * 32;
  // hand the staged results of the chunk that used buffer b to the caller
  auto drain = [&](InferBuf& b) -> int {
    if (!b.pending) return 0;
    CKE(cudaEventSynchronize(b.out), "sync inference D2H");
    if (P_host && !p_direct)
      parallel_rows_copy((char*)(P_host + b.s), (size_t)nq * sizeof(double), (const char*)b.hP, (size_t)b.n * sizeof(double), (size_t)b.n * sizeof(double), K);
    if (labels_host && !l_direct) std::memcpy(labels_host + b.s, b.hL, (size_t)b.n * sizeof(int32_t));
    b.pending = false;
    return 0;
  };
  // everything queued on the compute stream so far (refresh) precedes the first chunk's kernel by stream order
  int rc = 0;
  int64_t c = 0;
  for (int64_t s0 = 0; s0 < nq && rc == 0; s0 += chunk, ++c) {
    InferBuf& b = inf_[c & 1];
    const int64_t n = std::min(chunk, nq - s0);
    if ((rc = drain(b))) break;   // also guarantees that the buffer's previous kernel and copies are complete
    const double* xs = Xq + (size_t)s0 * d;
    if (!on_device) {
      const double* src = xs;
      if (!x_direct) {
        parallel_rows_copy((char*)b.hX, 0, (const char*)xs, 0, (size_t)n * d * sizeof(double), 1);
        src = b.hX;
      }
      CKE(cudaMemcpyAsync(b.dX, src, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, cst_), "H2D Xq chunk");
      CKE(cudaEventRecord(b.in, cst_), "event");
      CKE(cudaStreamWaitEvent(st_, b.in, 0), "wait");
      xs = b.dX;
    }
    if ((rc = run_estep(xs, n, P_host ? b.dP : nullptr, ldq, labels_host ? b.dL : nullptr, false))) break;
    CKE(cudaEventRecord(b.comp, st_), "event");
    CKE(cudaStreamWaitEvent(dst_, b.comp, 0), "wait");
    if (P_host) {
      if (p_direct)
        CKE(cudaMemcpy2DAsync(P_host + s0, (size_t)nq * sizeof(double), b.dP, (size_t)ldq * sizeof(double), (size_t)n * sizeof(double), K,
                              cudaMemcpyDeviceToHost, dst_), "D2H P chunk");
      else
        CKE(cudaMemcpy2DAsync(b.hP, (size_t)n * sizeof(double), b.dP, (size_t)ldq * sizeof(double), (size_t)n * sizeof(double), K,
                              cudaMemcpyDeviceToHost, dst_), "D2H P chunk");
    }
    if (labels_host)
      CKE(cudaMemcpyAsync(l_direct ? labels_host + s0 : b.hL, b.dL, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, dst_), "D2H labels chunk");
    CKE(cudaEventRecord(b.out, dst_), "event");
    // (the next use of this buffer -- H2D of chunk c+2 -- is enqueued only after drain() has waited for b.out on the host, so
    // no stream-side wait is needed; one on the copy stream would also hold back chunk c+1's upload into the OTHER buffer)
    b.s = s0; b.n = n; b.pending = true;
  }
  for (InferBuf& b : inf_) {
    const int r = drain(b);
    if (rc == 0) rc = r;
  }
  if (rc != 0) {   // leave no copy in flight into buffers the caller owns
    cudaStreamSynchronize(cst_); cudaStreamSynchronize(st_);
    if (dst_) cudaStreamSynchronize(dst_);
    for (InferBuf& b : inf_) b.pending = false;
  }
  return rc;
}

// The fused sweep never stores ProbaC(N,K); it is re-evaluated on demand with the parameters the last E-step used.
int EmgmmEngine::materialize_gamma() {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (gamma_current_) return 0;
  if (int r = ensure_gamma()) return r;
  if (!prev_valid_) return 0;  // no E-step yet: ProbaC = 0 (src/unsupervised/mlf_emgmm.f90:119)
  if (int r = ensure_moments()) return r;  // collective: empty shards join it before they leave
  if (n_ == 0) return 0;
  CKE(launch_refresh(geom_, dPrevMu_, dPrevCov_, dPrevLambda_, dSc_, dMean_, (double)nglob_, dScratch_, dPack_, dSumLLk_, dInfo_, st_,
                     dPack2_, dMuOld_, dShift_, 1, cov_diag_ ? 1 : 0, &fgeom_, 0), "refresh");
  refresh_count_ = 0;  // scratch now holds the eigenvectors of the previous parameters: next refresh starts cold
  ++launches_;
  if (int r = run_estep(dX_, n_, dGamma_, ldg_, nullptr, false)) return r;
  gamma_current_ = true;
  return 0;
}

const double* EmgmmEngine::gamma_device() {
  if (materialize_gamma() != 0) return nullptr;
  cudaStreamSynchronize(st_);
  return dGamma_;
}

int EmgmmEngine::resident_proba(double* P_host) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (int r = materialize_gamma()) return r;
  if (n_ == 0) return 0;
  CKE(cudaMemcpy2DAsync(P_host, (size_t)n_ * sizeof(double), dGamma_, (size_t)ldg_ * sizeof(double), (size_t)n_ * sizeof(double),
                        geom_.K, cudaMemcpyDeviceToHost, st_), "D2H ProbaC");
  CKE(cudaStreamSynchronize(st_), "sync");
  return 0;
}

int EmgmmEngine::resident_labels(int32_t* labels_host) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (int r = ensure_moments()) return r;  // collective: empty shards join it before they leave
  if (n_ == 0) return 0;
  if (!dLabels_) CKE(cudaMalloc((void**)&dLabels_, (size_t)n_ * sizeof(int32_t)), "cudaMalloc labels");
  if (int r = do_refresh()) return r;
  if (legacy_estep_) {
    if (int r = ensure_gamma()) return r;
    gamma_current_ = false;
  }
  if (int r = run_estep(dX_, n_, legacy_estep_ ? dGamma_ : nullptr, ldg_, dLabels_, false)) return r;
  CKE(cudaMemcpyAsync(labels_host, dLabels_, (size_t)n_ * sizeof(int32_t), cudaMemcpyDeviceToHost, st_), "D2H labels");
  CKE(cudaStreamSynchronize(st_), "sync");
  return 0;
}

// mlf_EvalGaussian for the single component of a K = 1 engine (src/unsupervised/mlf_gaussian.f90:63-79):
// P_i = lambda * exp(-maha_i/2 - lnDet) (not normalised) and LL = N (log lambda - lnDet) + sum_i (-maha_i/2).
int EmgmmEngine::eval_gaussian_raw(double* P_host, double* sumLL) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (geom_.K != 1) { err_ = "eval_gaussian_raw needs a single-component engine"; return -7; }
  if (int r = ensure_moments()) return r;
  if (int r = do_refresh()) return r;
  if (int r = ensure_gamma()) return r;
  gamma_current_ = false;
  if (n_ > 0)
    if (int r = run_estep(dX_, n_, dGamma_, ldg_, nullptr, false, true)) return r;
  if (P_host && n_ > 0) CKE(cudaMemcpyAsync(P_host, dGamma_, (size_t)n_ * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H P");
  if (sumLL) CKE(cudaMemcpyAsync(sumLL, dSumLLk_, sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H LL");
  CKE(cudaStreamSynchronize(st_), "sync");
  return 0;
}

// mlf_MaxGaussian for given weights (src/unsupervised/mlf_gaussian.f90:41-51): W = P/sum(P); GECovMatrix with
// minW = 1d-3/real(NX).  K = 1 engine; returns sum(P).
int EmgmmEngine::max_gaussian(const double* Proba_host, double* mu_host, double* C_host, double* sumPi) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (geom_.K != 1) { err_ = "max_gaussian needs a single-component engine"; return -7; }
  const Geom& g = geom_;
  if (int r = ensure_gamma()) return r;
  gamma_current_ = false;
  if (n_ > 0) CKE(cudaMemcpyAsync(dGamma_, Proba_host, (size_t)n_ * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D Proba");
  const int grid = std::max(1, std::min<int>(sms_, (int)((n_ + 255) / 256)));
  CKE(launch_stats_from_gamma(g, dX_, n_, dGamma_, dPartA_, grid, st_), "stats_from_gamma");
  CKE(launch_reduce_partials(dPartA_, grid, g.DS, dStatsA_, st_), "reduce");
  const double minW = 1e-3 / (double)(float)n_;
  CKE(launch_mid(g, dStatsA_, minW, dMu_, dMupad_, dThr_, st_), "mid");
  if (int r = run_mstep(dGamma_)) return r;
  CKE(launch_finalize(g, dStatsA_, dStatsB_, 0.0, dCov_, dLambda_, st_), "finalize");
  launches_ += 4;
  CKE(cudaMemcpyAsync(mu_host, dMu_, g.d * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H mu");
  CKE(cudaMemcpyAsync(C_host, dCov_, (size_t)g.d * g.d * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H C");
  if (sumPi) CKE(cudaMemcpyAsync(sumPi, dStatsA_, sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H sum");
  CKE(cudaStreamSynchronize(st_), "sync");
  return 0;
}

namespace {
// splitmix64: the host-side stream that replaces the reference's unseeded RANDOM_NUMBER (src/mlf_rand.f90:173-182).
// Every rank draws the same numbers from the same seed, so all ranks agree on the sampled point.
inline double next_uniform(uint64_t& st) {
  uint64_t z = (st += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
}  // namespace

// mlf_GetFarPoint (src/unsupervised/mlf_kmeans.f90:140-149): draw a data point with probability proportional to
// minDist^2 over the *global* data set.  dPart holds this rank's per-block sums; the owning rank copies the
// point into dOut (d doubles), the others contribute zeros, and a sum all-reduce broadcasts it.
int EmgmmEngine::draw_far_point(double u, int grid, int64_t chunk, double* dMinDist, double* dPart, double* dOut) {
  const int d = geom_.d;
  std::vector<double> part((size_t)grid);
  CKE(cudaMemcpyAsync(part.data(), dPart, grid * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H seed partials");
  CKE(cudaStreamSynchronize(st_), "sync");
  double local = 0;
  for (double v : part) local += v;
  std::vector<double> tot((size_t)world_, 0.0);
  tot[(size_t)rank_] = local;
  if (world_ > 1) {
    CKE(cudaMemcpyAsync(dOut, tot.data(), world_ * sizeof(double), cudaMemcpyHostToDevice, st_), "H2D totals");
    if (int r = allreduce(dOut, world_)) return r;
    CKE(cudaMemcpyAsync(tot.data(), dOut, world_ * sizeof(double), cudaMemcpyDeviceToHost, st_), "D2H totals");
    CKE(cudaStreamSynchronize(st_), "sync");
  }
  double all = 0;
  for (double v : tot) all += v;
  double target = u * all;  // r*W(size(W)) (src/mlf_rand.f90:180)
  int owner = world_ - 1;
  while (owner > 0 && tot[(size_t)owner] <= 0) --owner;   // last rank with any weight
  for (int r = 0; r < world_; ++r) {
    if (tot[(size_t)r] > 0 && target <= tot[(size_t)r]) { owner = r; break; }
    target -= tot[(size_t)r];
  }
  if (all <= 0) { owner = 0; target = 0; }  // degenerate: every point coincides with a centre -> first point
  CKE(cudaMemsetAsync(dOut, 0, (size_t)std::max(d, world_) * sizeof(double), st_), "memset");
  if (owner == rank_ && n_ > 0) {
    int b = grid - 1;
    while (b > 0 && part[(size_t)b] <= 0) --b;
    double t = target;
    for (int i = 0; i < grid; ++i) {
      if (part[(size_t)i] > 0 && t <= part[(size_t)i]) { b = i; break; }
      t -= part[(size_t)i];
    }
    if (t > part[(size_t)b]) t = part[(size_t)b];
    const int64_t lo = (int64_t)b * chunk, hi = std::min<int64_t>(n_, lo + chunk);
    CKE(launch_seed_pick(dX_, d, dMinDist, lo, hi, t, dOut, st_), "seed_pick");
    ++launches_;
  }
  if (world_ > 1)
    if (int r = allreduce(dOut, d)) return r;
  return 0;
}

// km_algo%init + km_algo%stepF(nkm) + Cov = I + lambda = 1/K  (src/unsupervised/mlf_emgmm.f90:135-145):
// one seeding pass (mlf_InitKMeans) followed by nkm-1 Lloyd steps (src/unsupervised/mlf_kmeans_naive.f90:55-96).
// The reference draws from an unseeded RANDOM_NUMBER stream; here the stream is splitmix64(seed), so parity with the
// reference is distributional only.  Empty clusters are re-seeded with mlf_GetFarPoint
// against the finished centres (the reference mixes finished centres and raw sums at that point, :86-92).
int EmgmmEngine::kmeans_init(int nkm, uint64_t seed) {
  CKE(cudaSetDevice(device_), "cudaSetDevice");
  if (int r = ensure_moments()) return r;
  const Geom& g = geom_;
  const int d = g.d, K = g.K;
  if (nglob_ < 1) { err_ = "k-means initialiser needs at least one point"; return -7; }
  uint64_t rng = seed;
  const int grid = sms_;
  const int64_t chunk = std::max<int64_t>(1, (n_ + grid - 1) / grid);
  // second-generation kernels (emgmm_kmeans.cu; MLF_B200_KMEANS_V1=1 keeps the first generation for A/B runs)
  const char* v1 = getenv("MLF_B200_KMEANS_V1");
  const bool gen2 = !(v1 && v1[0] == '1');
  const bool fast_seed = gen2 && d <= 64;
  const bool fast_lloyd = gen2 && !legacy_estep_ && !big_ && lloyd16_ok(g);
  const int sgrid = (int)std::max<int64_t>(1, std::min<int64_t>(4 * (int64_t)sms_, (n_ + 255) / 256));
  const int64_t schunk = std::max<int64_t>(256, (((n_ + sgrid - 1) / sgrid + 255) / 256) * 256);
  double *dMinDist = nullptr, *dPart = nullptr, *dOut = nullptr, *dCnt = nullptr, *dGather = nullptr;
  CKE(cudaMalloc((void**)&dMinDist, std::max<int64_t>(n_, 1) * sizeof(double)), "cudaMalloc minDist");
  CKE(cudaMalloc((void**)&dPart, (size_t)std::max(grid, sgrid) * sizeof(double)), "cudaMalloc");
  CKE(cudaMalloc((void**)&dOut, (size_t)std::max(d, world_) * sizeof(double)), "cudaMalloc");
  CKE(cudaMalloc((void**)&dCnt, K * sizeof(double)), "cudaMalloc");
  CKE(cudaMalloc((void**)&dGather, (size_t)world_ * sizeof(double)), "cudaMalloc");
  cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
  if (debug_)
    for (auto& e : tev) cudaEventCreate(&e);
  if (debug_) cudaEventRecord(tev[0], st_);
  int rc = 0;
  auto seed_pass = [&](const double* centre, int first) -> int {
    CKE(launch_seed_update(dX_, n_, d, centre, dMinDist, first, chunk, grid, dPart, st_), "seed_update");
    ++launches_;
    return 0;
  };
  // Mu(:,1) = mean(X,2); Mu(:,1) = GetFarPoint(X, Mu(:,1:1)); Mu(:,i) = GetFarPoint(X, Mu(:,1:i-1))
  for (int i = 0; i < K && rc == 0; ++i) {
    const double* centre = (i == 0) ? dMean_ : dMu_ + (size_t)(i - 1) * d;
    if (fast_seed) {   // everything stays on the stream: no host round trip per centre
      const double u = next_uniform(rng);
      if (launch_seed_update2(dX_, n_, d, centre, dMinDist, i <= 1 ? 1 : 0, schunk, sgrid, dPart, st_) != cudaSuccess ||
          launch_seed_total(dPart, sgrid, dGather, world_, rank_, st_) != cudaSuccess) { rc = -7; break; }
      if (world_ > 1 && (rc = allreduce(dGather, world_))) break;
      if (launch_seed_draw(dX_, d, n_, dMinDist, dPart, sgrid, schunk, dGather, world_, rank_, u, dOut, st_) != cudaSuccess) { rc = -7; break; }
      if (world_ > 1 && (rc = allreduce(dOut, d))) break;
      launches_ += 3;
    } else {
      rc = seed_pass(centre, i <= 1 ? 1 : 0);
      if (rc == 0) rc = draw_far_point(next_uniform(rng), grid, chunk, dMinDist, dPart, dOut);
    }
    if (rc == 0 && cudaMemcpyAsync(dMu_ + (size_t)i * d, dOut, d * sizeof(double), cudaMemcpyDeviceToDevice, st_) != cudaSuccess) rc = -7;
  }
  if (debug_) cudaEventRecord(tev[1], st_);
  const int64_t ltiles = (n_ + 63) / 64;
  const int lgrid = (int)std::max<int64_t>(1, std::min<int64_t>(sms_, (ltiles + 1) / 2));
  std::vector<double> cnt((size_t)K);
  for (int it = 1; it < nkm && rc == 0; ++it) {
    // mlf_EvaluateClass + EvaluateCentres
    if (legacy_estep_) {
      if (launch_kmeans_assign(dX_, n_, d, K, dMu_, nullptr, nullptr, dPartA_, grid, st_) != cudaSuccess) { rc = -7; break; }
      if (launch_reduce_partials(dPartA_, grid, (int64_t)K * (d + 1), dStatsA_, st_) != cudaSuccess) { rc = -7; break; }
      launches_ += 2;
      if ((rc = allreduce(dStatsA_, (int64_t)K * (d + 1)))) break;
      if (launch_kmeans_centres(dStatsA_, K, d, dMu_, dCnt, st_) != cudaSuccess) { rc = -7; break; }
    } else if (fast_lloyd) {
      // one sweep: scores from the linear monomials only (DM DMMA k-steps per 8 points x 8 centres), first maximum, one-hot
      // DMMA statistics in the layout of EM's pass A
      if (launch_lloyd16(g, dX_, n_, dMu_, dShift_, dPartA_, lgrid, st_) != cudaSuccess) { rc = -7; break; }
      if (launch_reduce_partials(dPartA_, 2 * lgrid, (int64_t)g.K8 * 8 * g.DS, dStatsA_, st_) != cudaSuccess) { rc = -7; break; }
      launches_ += 2;
      if ((rc = allreduce(dStatsA_, (int64_t)g.K8 * 8 * g.DS))) break;
      if (launch_kmeans_centres_a(g, dStatsA_, dMu_, dCnt, st_) != cudaSuccess) { rc = -7; break; }
    } else {
      // assignment through the E-pass kernel in hard mode: distances as DMMA tiles, one-hot labels, and the cluster
      // counts / coordinate sums from the same fixed-order pass-A statistics as EM (deterministic, no atomics)
      if (launch_pack_kmeans(g, fgeom_, dMu_, dShift_, dPack2_, st_) != cudaSuccess) { rc = -7; break; }
      ++launches_;
      if ((rc = run_estep(dX_, n_, nullptr, ldg_, nullptr, true, false, true))) break;
      if ((rc = allreduce(dStatsA_, (int64_t)g.K8 * 8 * g.DS))) break;
      if (launch_kmeans_centres_a(g, dStatsA_, dMu_, dCnt, st_) != cudaSuccess) { rc = -7; break; }
    }
    ++launches_;
    if (cudaMemcpyAsync(cnt.data(), dCnt, K * sizeof(double), cudaMemcpyDeviceToHost, st_) != cudaSuccess ||
        cudaStreamSynchronize(st_) != cudaSuccess) { rc = -7; break; }
    bool any_empty = false;
    for (double c : cnt) any_empty |= !(c > 0);
    if (!any_empty) continue;
    // distances to the finished centres (empty ones sit at +inf and attract nothing)
    if (launch_kmeans_assign(dX_, n_, d, K, dMu_, nullptr, dMinDist, nullptr, grid, st_) != cudaSuccess) { rc = -7; break; }
    ++launches_;
    bool have_part = false;
    for (int k = 0; k < K && rc == 0; ++k) {
      if (cnt[(size_t)k] > 0) continue;
      if (!have_part) {  // per-block sums of the current minDist^2: min with an infinitely far "centre" changes nothing
        rc = seed_pass(dMu_ + (size_t)k * d, 0);
        have_part = true;
      }
      if (rc == 0) rc = draw_far_point(next_uniform(rng), grid, chunk, dMinDist, dPart, dOut);
      if (rc == 0 && cudaMemcpyAsync(dMu_ + (size_t)k * d, dOut, d * sizeof(double), cudaMemcpyDeviceToDevice, st_) != cudaSuccess) rc = -7;
      if (rc == 0) rc = seed_pass(dMu_ + (size_t)k * d, 0);
    }
  }
  if (rc == 0 && launch_identity_cov(K, d, dCov_, dLambda_, st_) != cudaSuccess) rc = -7;
  ++launches_;
  if (debug_) cudaEventRecord(tev[2], st_);
  cudaStreamSynchronize(st_);
  if (debug_) {
    float ms_seed = 0, ms_lloyd = 0;
    cudaEventElapsedTime(&ms_seed, tev[0], tev[1]);
    cudaEventElapsedTime(&ms_lloyd, tev[1], tev[2]);
    if (rank_ == 0)
      fprintf(stderr, "mlfortran_b200[debug]: k-means initialiser: seeding of %d centres %.1f ms (%s), %d Lloyd steps %.1f ms (%s)\n", K, ms_seed,
              fast_seed ? "device-side draw" : "host-side draw", nkm - 1, ms_lloyd, fast_lloyd ? "k_lloyd16" : "E-pass kernel in hard mode");
    for (auto& e : tev) cudaEventDestroy(e);
  }
  cudaFree(dMinDist); cudaFree(dPart); cudaFree(dOut); cudaFree(dCnt); cudaFree(dGather);
  if (rc != 0 && err_.empty()) err_ = "k-means initialiser failed";
  return rc;
}

}  // namespace mlfb200
