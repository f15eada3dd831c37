// Single-process, multi-GPU EM-GMM: one EmgmmEngine per device, each driven by its own host thread, the points sharded in
// contiguous blocks.  The statistics all-reduce is done by this library itself over peer memory (NVLink / NVSwitch P2P loads):
// every rank's kernel sums all ranks' partial buffers in rank order, so all ranks end with bit-identical results.
// This is what lets a plain Fortran / C caller of the reference's API (one process, `init(X, nC)` + `step`) use every GPU of
// the box; the one-process-per-GPU mode (mlf_b200_set_allreduce + NCCL) is the other way to shard.
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "emgmm_engine.h"

namespace mlfb200 {

class MultiEngine {
 public:
  MultiEngine() = default;
  ~MultiEngine();
  MultiEngine(const MultiEngine&) = delete;
  MultiEngine& operator=(const MultiEngine&) = delete;

  // X: host array N x d (point-major); devices: one entry per shard (the same device may appear twice: used by the tests).
  int init(const double* X, int64_t n, int d, int K, const std::vector<int>& devices, bool cov_diag);
  int refresh_x(const double* X);
  int upload_params(const double* Mu, const double* Cov, const double* lambda);
  int download_params(double* Mu, double* Cov, double* lambda);
  int iterate(int64_t niter, double minProba, double* sumLL_last, double* sumLL_trace);
  int kmeans_init(int nkm, uint64_t seed);
  int expectation(const double* Xq, int64_t nq, double* P_host, int32_t* labels_host);  // on shard 0's device
  int resident_proba(double* P_host);      // ProbaC(N,K) of all shards: P[k*N + i]
  int resident_labels(int32_t* labels_host);
  void set_profile(bool on);
  void reset_times();
  EngineTimes times();                      // rank 0's counters
  int64_t launches();
  int64_t n_global() const { return n_; }
  int world() const { return (int)w_.size(); }
  std::string plan() const;
  const std::string& last_error() const { return err_; }
  bool set_cov_diag(bool on);

 private:
  struct Worker {
    std::unique_ptr<EmgmmEngine> eng;
    int rank = 0, device = 0;
    int64_t lo = 0, hi = 0;
    double* tmp = nullptr;   // all-reduce staging on this worker's device
    int64_t tmp_cap = 0;
  };
  struct ArCtx { MultiEngine* self; int rank; };
  static int ar_trampoline(void* ctx, double* buf, int64_t count, void* stream);
  int peer_allreduce(int rank, double* buf, int64_t count, cudaStream_t st);
  int run_all(const std::function<int(Worker&)>& f);   // f on every worker concurrently (one host thread each)
  bool barrier();                                        // false if some worker aborted

  std::vector<Worker> w_;
  std::vector<ArCtx> ctx_;
  std::vector<double*> table_;
  int64_t n_ = 0;
  int d_ = 0, K_ = 0;
  std::string err_;
  std::mutex mu_;
  std::condition_variable cv_;
  int waiting_ = 0;
  uint64_t generation_ = 0;
  bool abort_ = false;
};

}  // namespace mlfb200
