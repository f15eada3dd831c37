// Workload generator of the reference's own EM test program (tests/test_emgmm.f90:105-124, testGMM) on the device.
//   reference recipe                                              here
//   RandN(XC, sigmaC)                  (src/mlf_rand.f90:241-257)  centres c_k = sigmaC * N(0,I)          [host, stream 1]
//   weight_i = sqrt(sigmaX (1/i) / mean_j(1/j))  (test :116-117)   the same spectrum
//   randOrthogonal(C12(:,:,k), weight) (src/mlf_matrix.f90:153-180) Haar orthogonal by n-1 Householder
//                                                                   reflectors, columns scaled by D(k) w_k   [host, stream 2]
//   RandN(X(:, block k), X0 = XC(:,k), C12 = C12(:,:,k))           x = c_k + C12_k z,  z ~ N(0,I)         [device, stream 0]
//   randPerm(idx); X = X(:,idx)        (test :122-124)             output position i holds source point pi(i), pi a keyed
//                                                                   bijection of [0, N) (Feistel network + cycle walking)
// The reference draws from gfortran's unseeded RANDOM_NUMBER stream (xorshift1024*, Marsaglia polar method,
// src/mlf_rand.f90:198-209), which cannot be reproduced bit for bit; the distribution is what is kept.  Every draw here is a
// pure function of (seed, global source index): counter-based Philox4x32-10 + Box-Muller, so any shard [n0, n0 + nX) of the
// job can be generated independently on any GPU and the union is the same data set whatever the number of GPUs.
// Component of source point s is s mod nC (the reference's equal blocks of NX points per component up to relabelling).
// (The tests hold an independent CPU restatement of this file.)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mlf_b200.h"

namespace mlfb200 {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = mulhi32(M0, c0), l0 = M0 * c0, h1 = mulhi32(M1, c2), l1 = M1 * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += W0; k1 += W1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// 53 random bits -> a double strictly inside (0, 1)
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
  const uint64_t a = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)a + 0.5) * 0x1.0p-53;
}

enum : uint32_t { STREAM_POINTS = 0, STREAM_CENTRES = 1, STREAM_ORTHO = 2, STREAM_PERM = 3 };

// pair `pair` of normals of item `idx` in stream `stream`
__host__ __device__ inline void normal_pair(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t pair, double& z0, double& z1) {
  const Philox4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), pair, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u1 = u01(r.v[0], r.v[1]), u2 = u01(r.v[2], r.v[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
#ifdef __CUDA_ARCH__
  sincospi(2.0 * u2, &s, &c);
#else
  s = sin(6.283185307179586476925286766559 * u2);
  c = cos(6.283185307179586476925286766559 * u2);
#endif
  z0 = rad * c;
  z1 = rad * s;
}

struct PermKey {
  uint32_t rk[4];
  int half;         // bits per Feistel half
  uint64_t n;       // domain [0, n)
};

__host__ __device__ inline uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

// keyed bijection of [0, n): 4-round balanced Feistel network on 2*half bits, cycle-walked back into the domain
__host__ __device__ inline uint64_t permute(const PermKey& k, uint64_t i) {
  const uint64_t mask = (1ull << k.half) - 1ull;
  do {
    uint64_t L = i >> k.half, R = i & mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint64_t F = fmix32((uint32_t)R * 0x9E3779B1u + k.rk[r]) & mask;
      const uint64_t t = L ^ F;
      L = R;
      R = t;
    }
    i = (L << k.half) | R;
  } while (i >= k.n);
  return i;
}

static PermKey make_perm_key(uint64_t seed, uint64_t n) {
  PermKey k;
  int bits = 2;
  while (bits < 62 && (1ull << bits) < n) bits += 2;
  k.half = bits / 2;
  k.n = n;
  const Philox4 r = philox4x32_10(0, 0, 0, STREAM_PERM, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = 0; i < 4; ++i) k.rk[i] = r.v[i];
  return k;
}

constexpr int kSynthPB = 32;   // points per block iteration

// X[i - n0][:] = c_k + C12_k z(s),  s = pi(i), k = s mod K;  C12 in the Fortran layout C12(a,b,k) -> C12[(k d + b) d + a]
__global__ void __launch_bounds__(256) k_synth_points(double* __restrict__ X, int64_t n0, int64_t nX, int d, int K, uint64_t seed, PermKey pk,
                                                      const double* __restrict__ centres, const double* __restrict__ C12) {
  extern __shared__ double zs[];                 // [kSynthPB][dz], dz = 2 ceil(d/2) + 1
  __shared__ unsigned long long src[kSynthPB];
  const int hd = (d + 1) / 2, dz = 2 * hd + 1, tid = threadIdx.x;
  const int64_t nblk = (nX + kSynthPB - 1) / kSynthPB;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t p0 = blk * kSynthPB;
    const int np = (int)((nX - p0 < kSynthPB) ? (nX - p0) : kSynthPB);
    if (tid < np) src[tid] = permute(pk, (uint64_t)(n0 + p0 + tid));
    __syncthreads();
    for (int idx = tid; idx < np * hd; idx += blockDim.x) {
      const int p = idx / hd, j = idx - p * hd;
      double z0, z1;
      normal_pair(seed, STREAM_POINTS, src[p], (uint32_t)j, z0, z1);
      zs[p * dz + 2 * j] = z0;
      zs[p * dz + 2 * j + 1] = z1;       // the odd tail (2j+1 == d) is generated and ignored, as mlf_randNVect does
    }
    __syncthreads();
    for (int idx = tid; idx < np * d; idx += blockDim.x) {
      const int p = idx / d, a = idx - p * d;
      const int k = (int)(src[p] % (unsigned long long)K);
      const double* Ck = C12 + (size_t)k * d * d + a;
      const double* zp = zs + p * dz;
      double acc = 0.0;
      for (int b = 0; b < d; ++b) acc = fma(__ldg(Ck + (size_t)b * d), zp[b], acc);
      X[(p0 + p) * d + a] = centres[(size_t)k * d + a] + acc;
    }
    __syncthreads();
  }
}

// randOrthogonal(H, Sigma) of src/mlf_matrix.f90:153-180 (Stewart 1980, as SciPy's ortho_group): H <- H_{n-1} ... H_1 with H_k the
// Householder reflector of a Gaussian vector of length n-k+1, columns then scaled by D(k) Sigma(k), D(n) chosen for det = +1.
// Hm is column-major (Fortran) n x n.
static void rand_orthogonal(uint64_t seed, uint64_t item, int n, const double* sigma, double* Hm) {
  std::vector<double> v(n + 1), D(n, 1.0), w(n);
  for (int i = 0; i < n * n; ++i) Hm[i] = 0.0;
  for (int i = 0; i < n; ++i) Hm[i * n + i] = 1.0;
  uint32_t pair = 0;
  for (int k = 0; k + 1 < n; ++k) {
    const int m = n - k;
    for (int i = 0; i < m; i += 2) normal_pair(seed, STREAM_ORTHO, item, pair++, v[i], v[i + 1]);
    double nrm = 0.0;
    for (int i = 0; i < m; ++i) nrm += v[i] * v[i];
    D[k] = std::signbit(v[0]) ? -1.0 : 1.0;
    v[0] -= D[k] * std::sqrt(nrm);
    double vv = 0.0;
    for (int i = 0; i < m; ++i) vv += v[i] * v[i];
    // rows k.. of H  <-  (I - 2 v v^T / v.v) rows k.. of H
    for (int c = 0; c < n; ++c) {
      double dot = 0.0;
      for (int i = 0; i < m; ++i) dot += v[i] * Hm[c * n + k + i];
      w[c] = 2.0 * dot / vv;
    }
    for (int c = 0; c < n; ++c)
      for (int i = 0; i < m; ++i) Hm[c * n + k + i] -= v[i] * w[c];
  }
  double prod = 1.0;
  for (int k = 0; k + 1 < n; ++k) prod *= D[k];
  D[n - 1] = prod * ((n & 1) ? 1.0 : -1.0);
  for (int c = 0; c < n; ++c) {
    const double s = D[c] * (sigma ? sigma[c] : 1.0);
    for (int i = 0; i < n; ++i) Hm[c * n + i] *= s;
  }
}

}  // namespace mlfb200

extern "C" int mlf_b200_synth_gmm(double* X, int64_t n0, int64_t nX, int64_t nTotal, int nY, int nC, double sigmaC, double sigmaX, uint64_t seed,
                                  int device, unsigned flags, double* centres, double* C12) {
  using namespace mlfb200;
  if (nY < 1 || nC < 1 || nX < 0 || n0 < 0 || nTotal < 1 || n0 + nX > nTotal || (nX > 0 && !X)) return mlf_WRONGRANK;   // shape arguments out of range
  const int d = nY, K = nC;
  std::vector<double> hc((size_t)K * d), hC((size_t)K * d * d), w(d);
  double mean = 0.0;
  for (int i = 0; i < d; ++i) mean += 1.0 / (double)(i + 1);
  mean /= d;
  for (int i = 0; i < d; ++i) w[i] = std::sqrt(sigmaX * (1.0 / (double)(i + 1)) / mean);
  for (int k = 0; k < K; ++k) {
    for (int j = 0; j < d; j += 2) {
      double z0, z1;
      normal_pair(seed, STREAM_CENTRES, (uint64_t)k, (uint32_t)(j / 2), z0, z1);
      hc[(size_t)k * d + j] = sigmaC * z0;
      if (j + 1 < d) hc[(size_t)k * d + j + 1] = sigmaC * z1;
    }
    rand_orthogonal(seed, (uint64_t)k, d, w.data(), hC.data() + (size_t)k * d * d);
  }
  if (centres) std::copy(hc.begin(), hc.end(), centres);
  if (C12) std::copy(hC.begin(), hC.end(), C12);
  if (nX == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    fprintf(stderr, "mlf_b200_synth_gmm: no CUDA device %d; mlfortran_b200 has no CPU fallback\n", device);
    return -7;
  }
  if (cudaSetDevice(device) != cudaSuccess) return -7;
  const bool on_dev = (flags & MLF_B200_X_ON_DEVICE) != 0;
  double *dX = X, *dc = nullptr, *dC = nullptr;
  int rc = -7;
  const size_t xbytes = (size_t)nX * d * sizeof(double);
  do {
    if (!on_dev && cudaMalloc((void**)&dX, xbytes) != cudaSuccess) { dX = nullptr; rc = -2; break; }
    if (cudaMalloc((void**)&dc, hc.size() * sizeof(double)) != cudaSuccess || cudaMalloc((void**)&dC, hC.size() * sizeof(double)) != cudaSuccess) { rc = -2; break; }
    if (cudaMemcpy(dc, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (cudaMemcpy(dC, hC.data(), hC.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) break;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t nblk = (nX + kSynthPB - 1) / kSynthPB;
    const int grid = (int)std::min<int64_t>(nblk, (int64_t)sms * 8);
    const size_t smem = (size_t)kSynthPB * (2 * ((d + 1) / 2) + 1) * sizeof(double);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_synth_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) break;
    k_synth_points<<<grid, 256, smem>>>(dX, n0, nX, d, K, seed, make_perm_key(seed, (uint64_t)nTotal), dc, dC);
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) break;
    if (!on_dev && cudaMemcpy(X, dX, xbytes, cudaMemcpyDeviceToHost) != cudaSuccess) break;
    rc = 0;
  } while (false);
  if (!on_dev && dX) cudaFree(dX);
  cudaFree(dc);
  cudaFree(dC);
  return rc;
}
