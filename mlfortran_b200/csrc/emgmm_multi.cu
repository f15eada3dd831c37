// See emgmm_multi.h.  Reference context: the reference parallelises ExpStep/MaxStep with OpenMP threads over the components
// inside one process (src/unsupervised/mlf_emgmm.f90:179,199); here the one process drives every GPU of the box, sharded
// over the points, and the per-iteration exchange is a sum of K*(3+d+d^2)-sized statistics.
#include "emgmm_multi.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>

namespace mlfb200 {

namespace {

constexpr int kMaxRanks = 16;
struct PeerTable { const double* p[kMaxRanks]; };

// out[j] = sum_r peers[r][j], r ascending: every rank computes the same sum in the same order (bit-identical results).
// Peer buffers are read directly over NVLink / NVSwitch (P2P); the buffers are small (<= a few MB), so one pass suffices.
__global__ void __launch_bounds__(256) k_peer_sum(PeerTable t, int world, int64_t count, double* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (int64_t)gridDim.x * blockDim.x) {
    double s = 0;
    for (int r = 0; r < world; ++r) s += t.p[r][j];
    out[j] = s;
  }
}

}  // namespace

MultiEngine::~MultiEngine() {
  for (auto& w : w_) {
    if (w.tmp) {
      cudaSetDevice(w.device);
      cudaFree(w.tmp);
    }
  }
}

bool MultiEngine::barrier() {
  std::unique_lock<std::mutex> lk(mu_);
  if (abort_) return false;
  const uint64_t gen = generation_;
  if (++waiting_ == (int)w_.size()) {
    waiting_ = 0;
    ++generation_;
    cv_.notify_all();
    return true;
  }
  cv_.wait(lk, [&] { return generation_ != gen || abort_; });
  return !abort_;
}

int MultiEngine::run_all(const std::function<int(Worker&)>& f) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    abort_ = false;
    waiting_ = 0;
  }
  std::vector<int> rc(w_.size(), 0);
  std::vector<std::thread> th;
  th.reserve(w_.size());
  for (size_t i = 0; i < w_.size(); ++i)
    th.emplace_back([&, i] {
      cudaSetDevice(w_[i].device);
      rc[i] = f(w_[i]);
      if (rc[i] != 0) {  // release peers that may be waiting for this rank at a barrier
        std::lock_guard<std::mutex> lk(mu_);
        abort_ = true;
        cv_.notify_all();
      }
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < w_.size(); ++i)
    if (rc[i] != 0) {
      if (w_[i].eng && !w_[i].eng->last_error().empty()) err_ = w_[i].eng->last_error();
      else if (err_.empty()) err_ = "a shard failed (peer aborted)";
      return rc[i];
    }
  return 0;
}

int MultiEngine::ar_trampoline(void* ctx, double* buf, int64_t count, void* stream) {
  ArCtx* c = static_cast<ArCtx*>(ctx);
  return c->self->peer_allreduce(c->rank, buf, count, (cudaStream_t)stream);
}

int MultiEngine::peer_allreduce(int rank, double* buf, int64_t count, cudaStream_t st) {
  Worker& w = w_[(size_t)rank];
  if (count <= 0) return 0;
  if (w.tmp_cap < count) {
    if (w.tmp) cudaFree(w.tmp);
    w.tmp_cap = std::max<int64_t>(count, 1 << 16);
    if (cudaMalloc((void**)&w.tmp, (size_t)w.tmp_cap * sizeof(double)) != cudaSuccess) return -1;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;   // this rank's partial statistics are complete
  table_[(size_t)rank] = buf;
  if (!barrier()) return -1;                                  // every rank's buffer is ready and published
  PeerTable t{};
  for (size_t r = 0; r < w_.size(); ++r) t.p[r] = table_[r];
  const int grid = (int)std::min<int64_t>((count + 255) / 256, 296);
  k_peer_sum<<<grid, 256, 0, st>>>(t, (int)w_.size(), count, w.tmp);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return -1;
  if (!barrier()) return -1;                                  // every rank has finished reading every buffer
  if (cudaMemcpyAsync(buf, w.tmp, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
  return 0;
}

int MultiEngine::init(const double* X, int64_t n, int d, int K, const std::vector<int>& devices, bool cov_diag) {
  if (devices.empty() || (int)devices.size() > kMaxRanks) { err_ = "between 1 and 16 shards"; return -7; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    err_ = "no CUDA device: mlfortran_b200 has no CPU fallback";
    fprintf(stderr, "mlfortran_b200: %s\n", err_.c_str());
    return -7;
  }
  for (int dv : devices)
    if (dv < 0 || dv >= ndev) { err_ = "device index out of range"; return -7; }
  n_ = n; d_ = d; K_ = K;
  const int G = (int)devices.size();
  w_.resize((size_t)G);
  ctx_.resize((size_t)G);
  table_.assign((size_t)G, nullptr);
  const int64_t base = n / G, rem = n % G;
  for (int r = 0; r < G; ++r) {
    Worker& w = w_[(size_t)r];
    w.rank = r;
    w.device = devices[(size_t)r];
    w.lo = r * base + std::min<int64_t>(r, rem);
    w.hi = w.lo + base + (r < rem ? 1 : 0);
    ctx_[(size_t)r] = ArCtx{this, r};
  }
  // peer access between every pair of distinct devices (the all-reduce kernel loads peers' buffers directly)
  for (int a = 0; a < G; ++a)
    for (int b = 0; b < G; ++b) {
      if (devices[(size_t)a] == devices[(size_t)b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[(size_t)a], devices[(size_t)b]);
      if (!can) { err_ = "devices without peer access cannot share one EM object"; return -7; }
      cudaSetDevice(devices[(size_t)a]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[(size_t)b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err_ = "cudaDeviceEnablePeerAccess failed"; return -7; }
      cudaGetLastError();
    }
  return run_all([&](Worker& w) {
    w.eng.reset(new EmgmmEngine());
    int rc = w.eng->init(X + (size_t)w.lo * d, false, w.hi - w.lo, d, K, w.device, cov_diag);
    if (rc != 0) return rc;
    w.eng->set_allreduce(&MultiEngine::ar_trampoline, &ctx_[(size_t)w.rank], w.rank, (int)w_.size());
    return 0;
  });
}

int MultiEngine::refresh_x(const double* X) {
  return run_all([&](Worker& w) { return w.eng->refresh_x(X + (size_t)w.lo * d_, false); });
}
int MultiEngine::upload_params(const double* Mu, const double* Cov, const double* lambda) {
  return run_all([&](Worker& w) { return w.eng->upload_params(Mu, Cov, lambda); });
}
int MultiEngine::download_params(double* Mu, double* Cov, double* lambda) {
  cudaSetDevice(w_[0].device);
  return w_[0].eng->download_params(Mu, Cov, lambda);   // every rank holds identical parameters
}
int MultiEngine::iterate(int64_t niter, double minProba, double* sumLL_last, double* sumLL_trace) {
  return run_all([&](Worker& w) {
    return w.eng->iterate(niter, minProba, w.rank == 0 ? sumLL_last : nullptr, w.rank == 0 ? sumLL_trace : nullptr);
  });
}
int MultiEngine::kmeans_init(int nkm, uint64_t seed) {
  return run_all([&](Worker& w) { return w.eng->kmeans_init(nkm, seed); });
}
int MultiEngine::expectation(const double* Xq, int64_t nq, double* P_host, int32_t* labels_host) {
  // the moments / shift behind the parameter pack are all-reduced over the shards: every shard takes part in that step
  // (right after init, a checkpoint restore, inject or refresh_x nothing has run yet), then shard 0 evaluates the query
  if (int r0 = run_all([&](Worker& w) { return w.eng->prepare(); })) return r0;
  cudaSetDevice(w_[0].device);
  int r = w_[0].eng->expectation(Xq, false, nq, P_host, labels_host);
  if (r != 0) err_ = w_[0].eng->last_error();
  return r;
}
int MultiEngine::resident_proba(double* P_host) {
  // each shard fills its column range of ProbaC(N,K): P[k*N + lo .. hi)
  return run_all([&](Worker& w) {
    const int64_t nl = w.hi - w.lo;
    std::vector<double> tmp((size_t)K_ * std::max<int64_t>(nl, 1));
    int r = w.eng->resident_proba(tmp.data());  // empty shards still join the collective moments step inside
    if (r != 0 || nl == 0) return r;
    for (int k = 0; k < K_; ++k) std::memcpy(P_host + (size_t)k * n_ + w.lo, tmp.data() + (size_t)k * nl, (size_t)nl * sizeof(double));
    return 0;
  });
}
int MultiEngine::resident_labels(int32_t* labels_host) {
  return run_all([&](Worker& w) { return w.eng->resident_labels(labels_host + w.lo); });  // empty shards join the collectives
}
void MultiEngine::set_profile(bool on) {
  for (auto& w : w_) w.eng->set_profile(on);
}
void MultiEngine::reset_times() {
  for (auto& w : w_) w.eng->reset_times();
}
EngineTimes MultiEngine::times() { return w_[0].eng->times(); }
int64_t MultiEngine::launches() {
  int64_t s = 0;
  for (auto& w : w_) s += w.eng->launches();
  return s;
}
std::string MultiEngine::plan() const {
  char buf[64];
  snprintf(buf, sizeof buf, "%d shards in one process (peer-memory all-reduce) x ", (int)w_.size());
  return std::string(buf) + w_[0].eng->plan();
}
bool MultiEngine::set_cov_diag(bool on) {
  bool ok = true;
  for (auto& w : w_)
    if (!w.eng->set_cov_diag(on)) { ok = false; err_ = w.eng->last_error(); }
  return ok;
}

}  // namespace mlfb200
