// k-means initialiser of the EM step object on the GPU (src/unsupervised/mlf_emgmm.f90:135-145 -> mlf_InitKMeans +
// Lloyd steps, src/unsupervised/mlf_kmeans.f90:140-182, mlf_kmeans_naive.f90:55-96), second generation:
//   * D^2 seeding with no host round trip per centre: the per-block sums of minDist^2, the (gathered) per-rank totals and the
//     search for the drawn point all stay on the device (k_seed_update2 -> k_seed_total -> [all-reduce] -> k_seed_draw ->
//     [all-reduce]); one pass over X per centre, coalesced through a shared-memory transpose;
//   * Lloyd step as ONE sweep (k_lloyd16, d <= 16, K <= 64): the assignment needs only argmax_k (c_k.x - |c_k|^2/2), i.e. the
//     linear monomials of the quadratic-form sweep -- DM DMMA k-steps per 8 points x 8 centres instead of a whitening with L = I
//     (32 instead of 384 DMMA per 8 points at d = 16, K = 64) -- and the cluster counts / coordinate sums come from the same
//     fixed-order one-hot DMMA statistics as EM's pass A (deterministic, no atomics).
#include "emgmm_kernels.cuh"

#include <cmath>

namespace mlfb200 {

namespace {
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
constexpr unsigned kFull = 0xffffffffu;
constexpr int kSeedPts = 256;   // points per tile of k_seed_update2 (= threads per block)
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// seeding pass: minDist[p] = first ? |x_p - c| : min(minDist[p], |x_p - c|); partial[b] = sum over block b's contiguous chunk
// of minDist^2.  The tile of 256 points is loaded with fully coalesced 8-byte loads and transposed through shared memory
// (row stride d + 1 (d even) keeps the thread-per-point reads conflict-free).
// ---------------------------------------------------------------------------------------------------------------
template <int DC>   // DC = d when d is one of the instantiated sizes (the tile's loads are then held in registers one tile ahead), else 0
__global__ void __launch_bounds__(kSeedPts) k_seed_update2(const double* __restrict__ X, int64_t N, int d, const double* __restrict__ centre,
                                                           double* __restrict__ minDist, int first, int64_t chunk, double* __restrict__ partial) {
  extern __shared__ double sm[];
  const int RS = d | 1;   // odd row stride
  double* sc = sm;        // [d]
  double* tile = sm + ((d + 1) & ~1);   // [256][RS]
  __shared__ double red[kSeedPts];
  const int tid = threadIdx.x;
  for (int a = tid; a < d; a += kSeedPts) sc[a] = centre[a];
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = (lo + chunk < N) ? lo + chunk : N;
  double s = 0;
  if constexpr (DC > 0) {
    // software pipeline: the next tile's coordinates (DC coalesced 8-byte loads per thread) and its old minDist are in flight
    // while the current tile is transposed through shared memory and reduced -- one exposed memory latency per tile instead of two
    double nx[DC], nm = 0.0;
    auto fetch = [&](int64_t p0) {
      const int np = (int)((hi - p0 < kSeedPts) ? hi - p0 : kSeedPts);
      const double* src = X + p0 * DC;
#pragma unroll
      for (int i = 0; i < DC; ++i) nx[i] = (i * kSeedPts + tid < np * DC) ? src[i * kSeedPts + tid] : 0.0;
      nm = (!first && tid < np) ? minDist[p0 + tid] : 0.0;
    };
    if (lo < hi) fetch(lo);
    // element e = i * 256 + tid of a tile belongs to point e / DC, coordinate e % DC: advanced incrementally (no division per element)
    const int r0 = tid / DC, a0 = tid - r0 * DC, dr = kSeedPts / DC, da = kSeedPts - dr * DC;
    for (int64_t p0 = lo; p0 < hi; p0 += kSeedPts) {
      const int np = (int)((hi - p0 < kSeedPts) ? hi - p0 : kSeedPts);
      __syncthreads();   // previous tile consumed (and sc visible)
      {
        int r = r0, a = a0;
#pragma unroll
        for (int i = 0; i < DC; ++i) {
          tile[r * RS + a] = nx[i];
          r += dr; a += da;
          if (a >= DC) { a -= DC; ++r; }
        }
      }
      const double mold = nm;
      if (p0 + kSeedPts < hi) fetch(p0 + kSeedPts);
      __syncthreads();
      if (tid < np) {
        const double* x = tile + tid * RS;
        double q = 0;
#pragma unroll
        for (int a = 0; a < DC; ++a) { const double t = x[a] - sc[a]; q = fma(t, t, q); }
        double m = sqrt(q);  // norm2 (src/unsupervised/mlf_kmeans.f90:160)
        if (!first) m = fmin(mold, m);
        minDist[p0 + tid] = m;
        s = fma(m, m, s);
      }
    }
  } else {
    for (int64_t p0 = lo; p0 < hi; p0 += kSeedPts) {
      const int np = (int)((hi - p0 < kSeedPts) ? hi - p0 : kSeedPts);
      __syncthreads();   // previous tile consumed (and sc visible)
      const double* src = X + p0 * d;
      const int ne = np * d;
      for (int e = tid; e < ne; e += kSeedPts) {
        const int r = e / d, a = e - r * d;
        tile[r * RS + a] = src[e];
      }
      __syncthreads();
      if (tid < np) {
        const double* x = tile + tid * RS;
        double q = 0;
        for (int a = 0; a < d; ++a) { const double t = x[a] - sc[a]; q = fma(t, t, q); }
        double m = sqrt(q);  // norm2 (src/unsupervised/mlf_kmeans.f90:160)
        if (!first) m = fmin(minDist[p0 + tid], m);
        minDist[p0 + tid] = m;
        s = fma(m, m, s);
      }
    }
  }
  red[tid] = s;
  __syncthreads();
  for (int st = kSeedPts / 2; st > 0; st >>= 1) {
    if (tid < st) red[tid] += red[tid + st];
    __syncthreads();
  }
  if (tid == 0) partial[blockIdx.x] = red[0];
}
cudaError_t launch_seed_update2(const double* X, int64_t N, int d, const double* centre, double* minDist, int first, int64_t chunk, int grid,
                                double* partial, cudaStream_t st) {
  const size_t smem = ((size_t)((d + 1) & ~1) + (size_t)kSeedPts * (d | 1)) * sizeof(double);
  auto go = [&](auto kern) -> cudaError_t {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<grid, kSeedPts, smem, st>>>(X, N, d, centre, minDist, first, chunk, partial);
    return cudaGetLastError();
  };
  switch (d) {
    case 2: return go(k_seed_update2<2>);
    case 4: return go(k_seed_update2<4>);
    case 8: return go(k_seed_update2<8>);
    case 16: return go(k_seed_update2<16>);
    default: return go(k_seed_update2<0>);
  }
}

// gather[r] = (r == rank) ? sum_b partial[b] : 0   (fixed order; all-reduced by the caller when world > 1)
__global__ void k_seed_total(const double* __restrict__ partial, int grid, double* __restrict__ gather, int world, int rank) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double local = 0;
  for (int b = 0; b < grid; ++b) local += partial[b];
  for (int r = 0; r < world; ++r) gather[r] = (r == rank) ? local : 0.0;
}
cudaError_t launch_seed_total(const double* partial, int grid, double* gather, int world, int rank, cudaStream_t st) {
  k_seed_total<<<1, 32, 0, st>>>(partial, grid, gather, world, rank);
  return cudaGetLastError();
}

// mlf_rand_class on the cumulative vector of minDist^2 (src/mlf_rand.f90:173-182), for a data set spread over `world` ranks and
// `grid` contiguous chunks per rank: target = u * total; the owner rank is the first whose running total reaches it, inside the
// owner the first chunk, inside the chunk the first point.  The owner writes the drawn point to out[0..d), every other rank
// zeros (the caller's sum all-reduce then broadcasts it).  One block.
__global__ void __launch_bounds__(256) k_seed_draw(const double* __restrict__ X, int d, int64_t n, const double* __restrict__ minDist,
                                                   const double* __restrict__ partial, int grid, int64_t chunk, const double* __restrict__ gather,
                                                   int world, int rank, double u, double* __restrict__ out) {
  __shared__ double loc[256], pre[256];
  __shared__ long long s_idx, s_lo, s_hi;
  __shared__ double s_t;
  __shared__ int s_mine;
  const int tid = threadIdx.x;
  if (tid == 0) {
    double all = 0;
    for (int r = 0; r < world; ++r) all += gather[r];
    double target = u * all;  // r*W(size(W)) (src/mlf_rand.f90:180)
    int owner = world - 1;
    while (owner > 0 && gather[owner] <= 0) --owner;   // last rank with any weight
    for (int r = 0; r < world; ++r) {
      if (gather[r] > 0 && target <= gather[r]) { owner = r; break; }
      target -= gather[r];
    }
    if (all <= 0) { owner = 0; target = 0; }  // degenerate: every point coincides with a centre -> first point
    s_mine = (owner == rank && n > 0) ? 1 : 0;
    if (s_mine) {
      int b = grid - 1;
      while (b > 0 && partial[b] <= 0) --b;
      double t = target;
      for (int i = 0; i < grid; ++i) {
        if (partial[i] > 0 && t <= partial[i]) { b = i; break; }
        t -= partial[i];
      }
      if (t > partial[b]) t = partial[b];
      s_lo = (long long)b * chunk;
      s_hi = (s_lo + chunk < n) ? s_lo + chunk : n;
      s_t = t;
    }
  }
  __syncthreads();
  if (!s_mine) {
    for (int a = tid; a < d; a += 256) out[a] = 0.0;
    return;
  }
  const int64_t lo = s_lo, hi = s_hi;
  const double target = s_t;
  const int64_t cn = hi - lo, sub = (cn + 255) / 256;
  const int64_t a0 = lo + tid * sub, a1 = (a0 + sub < hi) ? a0 + sub : hi;
  double s = 0;
  for (int64_t p = a0; p < a1; ++p) { const double m = minDist[p]; s = fma(m, m, s); }
  loc[tid] = s;
  __syncthreads();
  if (tid == 0) {
    double run = 0;
    int owner = -1, lastnz = -1;
    for (int t = 0; t < 256; ++t) {
      pre[t] = run;
      run += loc[t];
      if (loc[t] > 0) lastnz = t;
      if (owner < 0 && loc[t] > 0 && run >= target) owner = t;
    }
    if (owner < 0) owner = lastnz;
    s_idx = (owner < 0) ? (long long)lo : -(long long)owner - 1;  // negative: thread -(v+1) scans its slice
  }
  __syncthreads();
  const long long code = s_idx;
  __syncthreads();
  if (code < 0 && tid == (int)(-code - 1)) {
    double run = pre[tid];
    long long idx = -1, lastnz = a0;
    for (int64_t p = a0; p < a1; ++p) {
      const double m = minDist[p];
      if (m > 0) lastnz = p;
      run = fma(m, m, run);
      if (m > 0 && run >= target) { idx = p; break; }
    }
    s_idx = idx >= 0 ? idx : lastnz;
  }
  __syncthreads();
  const int64_t idx = s_idx;
  for (int a = tid; a < d; a += 256) out[a] = X[idx * d + a];
}
cudaError_t launch_seed_draw(const double* X, int d, int64_t n, const double* minDist, const double* partial, int grid, int64_t chunk,
                             const double* gather, int world, int rank, double u, double* out, cudaStream_t st) {
  k_seed_draw<<<1, 256, 0, st>>>(X, d, n, minDist, partial, grid, chunk, gather, world, rank, u, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Lloyd step in one sweep (d <= 16, K <= 64): 148 persistent blocks x 16 warps in two independent groups of 8 (as k_em16),
// 64-point tiles per group.  Warp w: scores of its 8 points against every centre (DMMA, accumulator starts at -|c'|^2/2),
// first maximum = MINLOC of the distances (mlf_EvaluateClass, src/unsupervised/mlf_kmeans.f90:151-165); then the one-hot
// statistics of centre block w over the tile (counts and coordinate sums, EvaluateCentres, mlf_kmeans_naive.f90:74-96).
// partA[2 grid][KP][DS]: [count, count, sum x (8 DN)] -- the layout of EM's pass-A statistics.
// ---------------------------------------------------------------------------------------------------------------
template <int DM>
__global__ void __launch_bounds__(512, 1) k_lloyd16(const double* __restrict__ X, int64_t N, int d, int K, int K8, const double* __restrict__ Mu,
                                                    const double* __restrict__ shift, double* __restrict__ partA) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int XS = 8 * DN + 4;
  constexpr int DS = 8 * DN + 2;
  constexpr int TP = 64;
  extern __shared__ __align__(16) double smem[];
  const int KP = K8 * 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int cb = warp & 7, half = warp >> 3, wg = warp & 7;
  double* cpk = smem;                                   // [K8][DM][32] B fragments: lane (g, t) holds c'[8 cb + g][4 mb + t]
  double* cc = cpk + (size_t)K8 * DM * 32;              // [KP] -|c'|^2 / 2 (dead / dummy centres: -inf)
  double* xs = cc + KP + (size_t)half * TP * XS;        // [2][64][XS] raw points
  int* lab = reinterpret_cast<int*>(cc + KP + (size_t)2 * TP * XS) + half * TP;   // [2][64]
  for (int i = tid; i < K8 * DM * 32; i += 512) {
    const int l = i & 31, mb = (i >> 5) % DM, cq = (i >> 5) / DM;
    const int k = 8 * cq + (l >> 2), a = 4 * mb + (l & 3);
    const bool live = k < K && isfinite(Mu[(size_t)(k < K ? k : 0) * d]);
    cpk[i] = (live && a < d) ? Mu[(size_t)k * d + a] - shift[a] : 0.0;
  }
  for (int k = tid; k < KP; k += 512) {
    const bool live = k < K && isfinite(Mu[(size_t)(k < K ? k : 0) * d]);
    double q = 0;
    if (live)
      for (int a = 0; a < d; ++a) { const double c = Mu[(size_t)k * d + a] - shift[a]; q = fma(c, c, q); }
    cc[k] = live ? -0.5 * q : -INFINITY;
  }
  double sh[DM];
#pragma unroll
  for (int mb = 0; mb < DM; ++mb) sh[mb] = (4 * mb + t < d) ? shift[4 * mb + t] : 0.0;
  __syncthreads();
  const bool own = cb < K8;
  double accA[DN][2], sg = 0.0;
#pragma unroll
  for (int b = 0; b < DN; ++b) { accA[b][0] = 0.0; accA[b][1] = 0.0; }
  const int64_t ntiles = (N + TP - 1) / TP;
  double xraw[DM];
  bool first = true;
  for (int64_t bt = 2 * (int64_t)blockIdx.x + half; bt < ntiles; bt += 2 * (int64_t)gridDim.x) {
    const int64_t p = bt * TP + wg * 8 + g;
    if (first) {
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) {
        const int a = 4 * mb + t;
        xraw[mb] = (p < N && a < d) ? X[p * d + a] : 0.0;
      }
      first = false;
    }
    double xa[DM];
    double* xr = xs + (size_t)(wg * 8 + g) * XS;
#pragma unroll
    for (int mb = 0; mb < DM; ++mb) { xr[4 * mb + t] = xraw[mb]; xa[mb] = xraw[mb] - sh[mb]; }
    if (8 * DN > 4 * DM) xr[4 * DM + t] = 0.0;
    double best = -INFINITY;
    int bi = 0;
    for (int cq = 0; cq < K8; ++cq) {
      const double2 v = *reinterpret_cast<const double2*>(cc + 8 * cq + 2 * t);
      double c[2] = {v.x, v.y};
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) dmma(c, xa[mb], cpk[(size_t)(cq * DM + mb) * 32 + lane]);
      if (c[0] > best) { best = c[0]; bi = 8 * cq + 2 * t; }
      if (c[1] > best) { best = c[1]; bi = 8 * cq + 2 * t + 1; }
    }
#pragma unroll
    for (int m = 1; m <= 2; m <<= 1) {   // first maximum over the quad: larger score, or the same score at a lower index
      const double ob = __shfl_xor_sync(kFull, best, m);
      const int oi = __shfl_xor_sync(kFull, bi, m);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (t == 0) lab[wg * 8 + g] = (p < N) ? bi : -1;
    asm volatile("bar.sync %0, 256;\n" ::"r"(1 + half) : "memory");
    {  // next tile's points under the statistics phase
      const int64_t pn = p + 2 * (int64_t)gridDim.x * TP;
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) {
        const int a = 4 * mb + t;
        xraw[mb] = (pn < N && a < d) ? X[pn * d + a] : 0.0;
      }
    }
    if (own) {
#pragma unroll 4
      for (int pb = 0; pb < 16; ++pb) {
        const double a = (lab[4 * pb + t] == 8 * cb + g) ? 1.0 : 0.0;
        sg += a;
#pragma unroll
        for (int db = 0; db < DN; ++db) dmma(accA[db], a, xs[(size_t)(4 * pb + t) * XS + 8 * db + g]);
      }
    }
    asm volatile("bar.sync %0, 256;\n" ::"r"(1 + half) : "memory");
  }
  if (own) {
    double a = sg;
    a += __shfl_xor_sync(kFull, a, 1);
    a += __shfl_xor_sync(kFull, a, 2);
    double* o = partA + ((size_t)(2 * blockIdx.x + half) * KP + 8 * cb + g) * DS;
    if (t == 0) { o[0] = a; o[1] = a; }
#pragma unroll
    for (int db = 0; db < DN; ++db) {
      o[2 + 8 * db + 2 * t] = accA[db][0];
      o[2 + 8 * db + 2 * t + 1] = accA[db][1];
    }
  }
}

bool lloyd16_ok(const Geom& g) { return g.DM <= 4 && g.K8 <= 8; }
cudaError_t launch_lloyd16(const Geom& g, const double* X, int64_t N, const double* Mu, const double* shift, double* partA, int grid,
                           cudaStream_t st) {
  if (!lloyd16_ok(g)) return cudaErrorInvalidConfiguration;
  const int XS = 8 * g.DN + 4;
  const size_t smem = ((size_t)g.K8 * g.DM * 32 + g.K8 * 8 + 2 * 64 * (size_t)XS + 64) * sizeof(double);
#define MLF_L16(DMv)                                                                                         \
  case DMv: {                                                                                                \
    cudaError_t e = cudaFuncSetAttribute(k_lloyd16<DMv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                          \
    k_lloyd16<DMv><<<grid, 512, smem, st>>>(X, N, g.d, g.K, g.K8, Mu, shift, partA);                         \
  } break;
  switch (g.DM) {
    MLF_L16(1) MLF_L16(2) MLF_L16(3) MLF_L16(4)
    default: return cudaErrorInvalidConfiguration;
  }
#undef MLF_L16
  return cudaGetLastError();
}

}  // namespace mlfb200
