// Device-side interface of the EM-GMM hot path (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mlfb200 {

// Geometry shared by host and device code.
struct Geom {
  int d;        // point dimension
  int K;        // mixture components
  int DM;       // ceil(d/4): k-blocks of the whitening GEMM
  int DN;       // ceil(d/8): n-blocks
  int NB;       // active (mb,nb) fragment blocks of the lower-triangular factor: sum_mb (mb/2+1)
  int PS;       // doubles per component in the E-pass parameter pack (multiple of 4)
  int K4;       // ceil(K/4)   (quads of components, dummy components have c=-inf)
  int K8;       // ceil(K/8)
  int S;        // row stride (doubles) of the per-block logit/responsibility tile, == 4 (mod 16)
  int DS;       // doubles per component of the pass-A statistics: [sum g, sum g^2, sum g*x (8*DN)]
  int MB;       // lower-triangular 8x8 blocks of the scatter matrix: DN*(DN+1)/2
  int MS;       // doubles per component of the pass-B statistics: [MB*64 scatter blocks, sum g^2] padded
};

Geom make_geom(int d, int K, bool big = false);

// Geometry of the fused single-pass kernel (emgmm_fused.cu).
struct FusedGeom {
  int OFFB;       // offset of the accumulator bias -L^T(mu - m) inside a component's pack (= NB*32)
  int OFFC;       // offset of log(lambda) - lnDet
  int PSF;        // doubles per component in the fused pack
  int XS;         // row stride of the shared-memory X tile (8*DN + 4, conflict-free for both fragment shapes)
  int MSF;        // doubles per component of the fused pass-B statistics: [MB*64 scatter | 8*DN sum_kept g x | sum_kept g, pad]
  int CBMAX;      // component blocks (of 8) per warp
  int GPW;        // 8-point groups per warp: 2 (128-point tiles) or 1 (64-point tiles, large K)
  int TP;         // points per block tile = 64 * GPW
  bool diag;      // diagonal-compressed pack [8*DN diag(L) | 8*DN bias | const] (diagonal-covariance extension)
  size_t smem;    // dynamic shared memory of k_em
  bool scatter_ok;  // the scatter accumulators fit the register file for this (d, K)
};
FusedGeom make_fused_geom(const Geom& g, bool cov_diag = false, size_t smem_limit = 232448);

struct AmbEntry {  // pair whose keep/drop decision waits for the all-reduced s_k
  double g;
  long long key;   // point index * 1024 + component
};

struct EmArgs {
  const double* X; long long N; int d, K, K8, S;
  const double* pack2;   // [K8*8][PSF]
  const double* shift;   // [>= 4*DM] global mean of the data, zero padded
  const double* muold;   // [K8*8][8*DN] current means, zero padded
  const double* tlo; const double* thi;   // [K8*8] band around minW*s_k (dummy components: +inf)
  double* gamma; long long ldg; int* labels; int raw;
  double* partA;         // [grid][K8*8][DS]
  double* partB;         // [grid][K8*8][MSF]
  AmbEntry* amb; int amb_cap; int* amb_count;   // [grid*8][amb_cap], [grid*8]
  int accumulate;        // start from the partials / list counts already stored (chunked sweep of a streamed X)
  long long p_offset;    // index of X[0] inside the rank's shard (side-list keys)
  int hard;              // k-means assignment: one-hot responsibilities of the largest logit (E-pass variants only)
  // expanded-quadratic-form logits of the 16-warp sweep (see qf_* below); qf_limit < 0: whitening only
  const double* packq;   // [K8][qf_stride(DM)] coefficient fragments + the 8 constants of each component block
  const double* qfS;     // [K8*8] magnitude bound S_k of the terms of component k's quadratic form over the data
  double qf_limit;       // the sweep uses the quadratic form iff every S_k <= qf_limit (decided on the device, no host sync)
  // "statistics first" iterations: responsibilities of the sweep's 64-point tiles, [tile][64][K8*8], written by the statistics
  // pass (launch_em16 with store = true) and read back by the scatter pass (launch_scatter16)
  double* gtiles;
  // k_em16 / k_scatter16 with fewer than 8 component blocks: the statistics of a component block are split over spl = 8 / K8'
  // warps (K8' = K8 rounded up to a power of two), each taking 16 / spl of the tile's 4-point groups and writing its own slot of
  // partials -- so that small K does not leave 7 of a group's 8 warps waiting for the one that owns the only block
  int spl;               // 0 or 1: no split
};
inline int em16_split(int K8) { return K8 <= 1 ? 8 : (K8 <= 2 ? 4 : (K8 <= 4 ? 2 : 1)); }

// ---- expanded quadratic form of the log-density (k_em16<.., QF = true>) ------------------------------------------------------------
// logit_k(x) = c_k - 1/2 (x'-mu')^T invC_k (x'-mu')  with x' = x - shift, mu' = mu_k - shift
//            = [c_k - 1/2 mu'^T invC mu'] + (invC mu') . x' - 1/2 x'^T invC x'
// is ONE dot product of the component's coefficient vector with the point's monomials: D4 linear ones (x'_i) and the
// D4 (D4+1)/2 products x'_i x'_j (i <= j), D4 = 4 DM padded coordinates.  At d = 16 that is 152 features = 38 DMMA k-steps
// for 8 points x 8 components (4.75 DMMA per component instead of the 6 of the triangular whitening, and no squares or
// quad reduction afterwards: the accumulator is the logit).  Its rounding error is u * S_k, S_k = sum of the magnitudes of
// the terms (it grows with (distance from the shift / cluster width)^2, where whitening grows linearly), so the engine
// bounds S_k per component at every refresh and the sweep falls back to whitening when the bound exceeds qf_limit.
__host__ __device__ constexpr int qf_nfeat(int DM) { return 4 * DM + (4 * DM) * (4 * DM + 1) / 2; }
__host__ __device__ constexpr int qf_nsteps(int DM) { return (qf_nfeat(DM) + 3) / 4; }
__host__ __device__ constexpr int qf_stride(int DM) { return qf_nsteps(DM) * 32 + 8; }
// feature f -> factors (i, j): coordinate indices into the padded point [x'_0 .. x'_{D4-1}, 1, 0]; f >= qf_nfeat: (D4+1, D4+1) = 0
__host__ __device__ constexpr void qf_feature(int D4, int f, int& i, int& j) {
  if (f < D4) { i = f; j = D4; return; }
  int q = f - D4;
  for (int r = 0; r < D4; ++r) {
    if (q < D4 - r) { i = r; j = r + q; return; }
    q -= D4 - r;
  }
  i = D4 + 1; j = D4 + 1;
}
struct QfRefresh {       // extra outputs of k_refresh for the quadratic-form sweep
  double* packq;         // [K8][qf_stride(DM)]
  double* qfS;           // [K8*8]
  const double* rad;     // [nrad][16]: per-rank maxima of |x_a - shift_a| (gathered through the sum all-reduce); nullptr: unknown
  int nrad;
};

constexpr int kEstepThreads = 256;      // 8 warps per persistent block
constexpr int kEstepTilePts = 128;      // 16 points per warp
constexpr int kMstepThreads = 256;
constexpr int kMstepTilePts = 2048;     // points per block tile of the M-pass sweep

// Parameter refresh (reference InverseSymMatrix eigen branch + E-pass packing + closed-form sum of maha).
//   scratch: K * 3*d*d doubles.  Sc (d*d, column-major) / mean (d) / Nglob describe the whole data set.
cudaError_t launch_refresh(const Geom& g, const double* Mu, const double* Cov, const double* lambda, const double* Sc,
                           const double* mean, double Nglob, double* scratch, double* pack, double* sumLLk, int* info,
                           cudaStream_t st, double* pack2 = nullptr, double* muold = nullptr, const double* shift = nullptr,
                           int do_sumll = 1, int cov_diag = 0, const FusedGeom* fg = nullptr, int warm = 0, const QfRefresh* qf = nullptr,
                           int* vvalid = nullptr);   // vvalid: [K8*8] ints, zeroed by the caller once: enables the Cholesky fast path
// max_i |x_ia - shift_a| per coordinate (d <= 16): partial[grid][16], then out[a] = max over blocks (16 values)
cudaError_t launch_absmax(const double* X, int64_t N, int d, const double* shift, double* partial, int grid, double* out, cudaStream_t st);
cudaError_t launch_sumll(const Geom& g, const double* Mu, const double* Sc, const double* mean, double Nglob, const double* scratch,
                         const double* pack, double* sumLLk, cudaStream_t st);
// Global moments about `shift`: partial[grid][MB*64 + 8*DN]; finish turns the all-reduced sums into mean and centred scatter.
cudaError_t launch_moments(const Geom& g, const double* X, int64_t N, const double* shift, double* partial, int grid, int accumulate,
                           cudaStream_t st);
cudaError_t launch_moments_finish(const Geom& g, const double* raw, const double* shift, double Nglob, double* mean, double* Sc,
                                  cudaStream_t st);

// Fused sweep (scatter = true) or E-pass + pass-A statistics only (scatter = false); see emgmm_fused.cu.
cudaError_t launch_em(const Geom& g, const FusedGeom& f, const EmArgs& a, bool scatter, int grid, size_t smem_limit, cudaStream_t st);
// 16-warp variant with the scatter accumulators in tensor memory (emgmm_fused16.cu): partials per (block, half) -> 2*grid slots,
// side lists per warp -> 16*grid lists.
bool em16_ok(const Geom& g, const FusedGeom& f, size_t smem_limit);
bool em16_qf_ok(const Geom& g, const FusedGeom& f, size_t smem_limit);   // the quadratic-form variant fits shared memory
cudaError_t launch_em16(const Geom& g, const FusedGeom& f, const EmArgs& a, int grid, size_t smem_limit, cudaStream_t st, bool store = false);
// Scatter pass of a "statistics first" iteration: thresholded covariance scatter + kept sums from the stored responsibilities
// (a.gtiles) with the exact thresholds a.thi; same tile -> (block, group) assignment and the same arithmetic as the fused sweep,
// so the pass-B partials are bit-identical to a repeated full sweep.  Writes a.partB only.
cudaError_t launch_scatter16(const Geom& g, const FusedGeom& f, const EmArgs& a, int grid, size_t smem_limit, cudaStream_t st);
// Diagonal-covariance extension at small d / large K on the fp64 CUDA cores (emgmm_diag.cu): lane = component.
bool emd_ok(const Geom& g, const FusedGeom& f, size_t smem_limit);
cudaError_t launch_selftest_exp(const double* in, double* out, int64_t n, int which, cudaStream_t st);   // exp_tab (which = 0) / exp_tab128 (1) on n device values
cudaError_t launch_emd(const Geom& g, const FusedGeom& f, EmArgs a, int grid, size_t smem_limit, cudaStream_t st);
// cmax: upper bound of distinct components per list (a warp's lists only hold the components that warp owns);
// pcomp: [nlists][cmax+1] ints, partial: [nlists][cmax][d*d+d+1] doubles of scratch
cudaError_t launch_amb_resolve(const Geom& g, const double* X, const AmbEntry* amb, int cap, const int* counts, int nlists,
                               const double* statsA, double minW, const double* muold, double* outC, cudaStream_t st, int cmax, int* pcomp,
                               double* partial);
cudaError_t launch_amb_summary(const int* counts, int nlists, int cap, double* flags, cudaStream_t st, const double* statsB = nullptr,
                               int KP = 0, int MSF = 0, int off = 0, double vote = 0.0);   // flags[0..3] are all-reduced, flags[4] is local
cudaError_t launch_finalize_fused(const Geom& g, const FusedGeom& f, const double* statsA, const double* statsB, const double* statsC,
                                  const double* muold, double minProba, double* Mu, double* Cov, double* lambda, cudaStream_t st,
                                  int cov_diag = 0);

// Large-shape path (emgmm_big.cu): 32 < d <= 128, or packs that do not fit shared memory.
size_t estep_big_smem(const Geom& g);
bool estep_big_ok(const Geom& g, size_t smem_limit);
int estep_big_tile(const Geom& g, size_t smem_limit);
bool estep_big_stage(const Geom& g, size_t smem_limit);   // pack staged in shared memory with cp.async (MLF_B200_BIG_STAGE=0 disables)
cudaError_t launch_estep_big(const Geom& g, EmArgs a, int grid, size_t smem_limit, cudaStream_t st);
void mstep_big_grid(const Geom& g, int sms, int64_t N, int* gx, int64_t* chunk);
cudaError_t launch_mstep_big(const Geom& g, const double* X, int64_t N, const double* gamma, int64_t ldg, const double* mupad,
                             const double* thr, double* partial, int gx, int64_t chunk, cudaStream_t st);

// E-pass: responsibilities (component-major gamma[k*ldg + i]), optional 1-based labels, and per-block
// partial pass-A statistics partial[block][K8*8][DS].  Returns the grid size used through *grid_out.
cudaError_t launch_estep(const Geom& g, const double* X, int64_t N, const double* pack, double* gamma, int64_t ldg,
                         int32_t* labels, double* partial, int grid, size_t smem_limit, int raw, cudaStream_t st);
size_t estep_smem_bytes(const Geom& g);
int estep_cbmax(const Geom& g);

// Pass-A statistics of a single component from given weights gamma[0..N): partial[grid][DS].
cudaError_t launch_stats_from_gamma(const Geom& g, const double* X, int64_t N, const double* gamma, double* partial, int grid,
                                    cudaStream_t st);

// Fixed-order reduction over blocks: out[j] = sum_b partial[b*n + j].
cudaError_t launch_reduce_partials(const double* partial, int nblocks, int64_t n, double* out, cudaStream_t st);

// After pass A (and its all-reduce): s_k, mu_k = (sum g x)/s_k, thr_k = minW*s_k; writes Mu (K*d) and
// mupad (K*8*DN, zero padded), thr (K).
cudaError_t launch_mid(const Geom& g, const double* statsA, double minW, double* Mu, double* mupad, double* thr,
                       cudaStream_t st);

// M-pass: thresholded centred scatter sum_{g>=thr} g (x-mu)(x-mu)^T per component via DMMA, plus sum g^2.
// partial[block][K][MS].  gamma may be nullptr (weight 1 for every point; used for the global moments).
cudaError_t launch_mstep(const Geom& g, const double* X, int64_t N, const double* gamma, int64_t ldg, const double* mupad,
                         const double* thr, double* partial, int grid, cudaStream_t st);
int mstep_cpw(const Geom& g);
void mstep_grid(const Geom& g, int sms, int64_t N, int* gx, int* gy);

// After pass B (and its all-reduce): Cov_k = ms^2/s_k * S_k (lower triangle mirrored), lambda floor + normalise.
cudaError_t launch_finalize(const Geom& g, const double* statsA, const double* statsB, double minProba, double* Cov,
                            double* lambda, cudaStream_t st, int cov_diag = 0);

// sum over points of x (per block partials [grid][d]) for the one-time global mean.
cudaError_t launch_colsum(const double* X, int64_t N, int d, double* partial, int grid, cudaStream_t st);

// sumLL = sum_k sumLLk[k] in k order (single thread; deterministic).
cudaError_t launch_sum_k(const double* v, int K, double* out, cudaStream_t st);

// k-means (initialiser): assignment + per-block centre sums.  Lloyd step of the reference
// (src/unsupervised/mlf_kmeans.f90:151-165, mlf_kmeans_naive.f90:74-96).
cudaError_t launch_kmeans_assign(const double* X, int64_t N, int d, int K, const double* Mu, int32_t* cl, double* minDist,
                                 double* partial /*[grid][K][d+1]*/, int grid, cudaStream_t st);


// k-means seeding: minDist[p] = first ? |x_p - c| : min(minDist[p], |x_p - c|); partial[b] = sum over block b's
// contiguous chunk of minDist^2.  k_seed_pick copies the point at which the running sum reaches `target`.
cudaError_t launch_seed_update(const double* X, int64_t N, int d, const double* centre, double* minDist, int first,
                               int64_t chunk, int grid, double* partial, cudaStream_t st);
cudaError_t launch_seed_pick(const double* X, int d, const double* minDist, int64_t lo, int64_t hi, double target,
                             double* out, cudaStream_t st);
cudaError_t launch_kmeans_centres(const double* sums, int K, int d, double* Mu, double* counts, cudaStream_t st);
// Second-generation initialiser kernels (emgmm_kmeans.cu): seeding without host round trips, Lloyd step in one sweep.
cudaError_t launch_seed_update2(const double* X, int64_t N, int d, const double* centre, double* minDist, int first, int64_t chunk, int grid,
                                double* partial, cudaStream_t st);   // d <= 64
cudaError_t launch_seed_total(const double* partial, int grid, double* gather, int world, int rank, cudaStream_t st);
cudaError_t launch_seed_draw(const double* X, int d, int64_t n, const double* minDist, const double* partial, int grid, int64_t chunk,
                             const double* gather, int world, int rank, double u, double* out, cudaStream_t st);
bool lloyd16_ok(const Geom& g);
cudaError_t launch_lloyd16(const Geom& g, const double* X, int64_t N, const double* Mu, const double* shift, double* partA, int grid,
                           cudaStream_t st);   // partA[2 grid][K8*8][DS]
cudaError_t launch_kmeans_centres_a(const Geom& g, const double* statsA, double* Mu, double* counts, cudaStream_t st);
cudaError_t launch_identity_cov(int K, int d, double* Cov, double* lambda, cudaStream_t st);
// Pack for the k-means assignment through the E-pass kernels: L = I, bias = -(centre - shift), constant 0.
cudaError_t launch_pack_kmeans(const Geom& g, const FusedGeom& f, const double* Mu, const double* shift, double* pack2, cudaStream_t st);

}  // namespace mlfb200
