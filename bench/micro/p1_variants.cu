// Micro-benchmark of the whitening loop (P1 of the fused sweep, d = 16, K = 64): instruction-scheduling variants timed in
// isolation.  Nothing here is on the product path; it exists to choose the P1 structure of emgmm_fused.cu by measurement.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o p1_variants p1_variants.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int DM = 4, DN = 2, NB = 6, OFFB = NB * 32, OFFC = OFFB + 8 * DN, PSF = OFFC + 4, K = 64, K4 = 16, S = 68;
__host__ __device__ constexpr int blk_off(int mb) { int s = 0; for (int m = 0; m < mb; ++m) s += m / 2 + 1; return s; }

__device__ __forceinline__ void dmma_v(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_n(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <bool VOL>
__device__ __forceinline__ void dm(double (&c)[2], double a, double b) { if (VOL) dmma_v(c, a, b); else dmma_n(c, a, b); }

// quad-transposed reduction of part[4] -> maha of component 4*k4+t, then logit store + running max
__device__ __forceinline__ void reduce_store(const double (&part)[4], double ck, double* row, int k4, int t, double& rmax, int& rarg) {
  const bool hi2 = (t & 2) != 0, hi1 = (t & 1) != 0;
  const double send0 = hi2 ? part[0] : part[2], send1 = hi2 ? part[1] : part[3];
  const double keep0 = hi2 ? part[2] : part[0], keep1 = hi2 ? part[3] : part[1];
  const double a0 = keep0 + shx(send0, 2), a1 = keep1 + shx(send1, 2);
  const double maha = (hi1 ? a1 : a0) + shx(hi1 ? a0 : a1, 1);
  const double logit = fma(-0.5, maha, ck);
  row[4 * k4 + t] = logit;
  if (logit > rmax) { rmax = logit; rarg = 4 * k4 + t; }
}

template <bool VOL>
__device__ __forceinline__ void comp_pair(const double* pk, int lane, int t, const double (&xa)[2][DM], double& p0, double& p1) {
  double Lf[NB];
  double2 nb2[DN];
#pragma unroll
  for (int b = 0; b < NB; ++b) Lf[b] = pk[b * 32 + lane];
#pragma unroll
  for (int nb = 0; nb < DN; ++nb) nb2[nb] = *reinterpret_cast<const double2*>(pk + OFFB + 8 * nb + 2 * t);
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    double c[DN][2];
#pragma unroll
    for (int nb = 0; nb < DN; ++nb) { c[nb][0] = nb2[nb].x; c[nb][1] = nb2[nb].y; }
#pragma unroll
    for (int mb = 0; mb < DM; ++mb)
#pragma unroll
      for (int nb = 0; nb <= mb / 2; ++nb) dm<VOL>(c[nb], xa[grp][mb], Lf[blk_off(mb) + nb]);
    double s = c[0][0] * c[0][0];
    s = fma(c[0][1], c[0][1], s);
#pragma unroll
    for (int nb = 1; nb < DN; ++nb) { s = fma(c[nb][0], c[nb][0], s); s = fma(c[nb][1], c[nb][1], s); }
    if (grp == 0) p0 = s; else p1 = s;
  }
}

// VAR 0: as in emgmm_fused.cu.  1: k4 loop unrolled by 2.  2: tail of iteration k4-1 reduced after the DMMAs of k4 are issued.
// 3: non-volatile asm.  4: variant 2 + non-volatile.
template <int VAR>
__global__ void __launch_bounds__(256, 1) k_p1(const double* __restrict__ pack, const double* __restrict__ X, int tiles, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  double* sp = smem;
  double* tile = smem + K * PSF;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < K * PSF; i += 256) sp[i] = pack[i];
  __syncthreads();
  constexpr bool VOL = (VAR != 3 && VAR != 4);
  double acc = 0;
  for (int bt = 0; bt < tiles; ++bt) {
    double xa[2][DM];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp)
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) xa[grp][mb] = X[((size_t)(blockIdx.x * tiles + bt) * 128 + warp * 16 + 8 * grp + g) * 16 + 4 * mb + t];
    double* row[2] = {tile + (size_t)(warp * 16 + g) * S, tile + (size_t)(warp * 16 + 8 + g) * S};
    double rmax[2] = {-1e300, -1e300};
    int rarg[2] = {0, 0};
    if (VAR == 0 || VAR == 3) {
      for (int k4 = 0; k4 < K4; ++k4) {
        double part[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) comp_pair<VOL>(sp + (size_t)(4 * k4 + j) * PSF, lane, t, xa, part[0][j], part[1][j]);
        const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
        reduce_store(part[0], ck, row[0], k4, t, rmax[0], rarg[0]);
        reduce_store(part[1], ck, row[1], k4, t, rmax[1], rarg[1]);
      }
    } else if (VAR == 1) {
#pragma unroll 2
      for (int k4 = 0; k4 < K4; ++k4) {
        double part[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) comp_pair<VOL>(sp + (size_t)(4 * k4 + j) * PSF, lane, t, xa, part[0][j], part[1][j]);
        const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
        reduce_store(part[0], ck, row[0], k4, t, rmax[0], rarg[0]);
        reduce_store(part[1], ck, row[1], k4, t, rmax[1], rarg[1]);
      }
    } else {
      double prev[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) comp_pair<VOL>(sp + (size_t)j * PSF, lane, t, xa, prev[0][j], prev[1][j]);
      for (int k4 = 1; k4 < K4; ++k4) {
        double part[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) comp_pair<VOL>(sp + (size_t)(4 * k4 + j) * PSF, lane, t, xa, part[0][j], part[1][j]);
        const double ck = sp[(size_t)(4 * (k4 - 1) + t) * PSF + OFFC];
        reduce_store(prev[0], ck, row[0], k4 - 1, t, rmax[0], rarg[0]);
        reduce_store(prev[1], ck, row[1], k4 - 1, t, rmax[1], rarg[1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { prev[0][j] = part[0][j]; prev[1][j] = part[1][j]; }
      }
      const double ck = sp[(size_t)(4 * (K4 - 1) + t) * PSF + OFFC];
      reduce_store(prev[0], ck, row[0], K4 - 1, t, rmax[0], rarg[0]);
      reduce_store(prev[1], ck, row[1], K4 - 1, t, rmax[1], rarg[1]);
    }
    acc += rmax[0] + rmax[1] + rarg[0] + rarg[1];
  }
  if (acc == 1.2345e-300) out[0] = acc + tile[tid];
}

// pure DMMA stream with the same ILP structure (2 chains) for reference
__global__ void __launch_bounds__(256, 1) k_dmma_ref(int iters, double* out) {
  double c0[2] = {threadIdx.x * 1e-9, 0}, c1[2] = {1, 2}, c2[2] = {3, 4}, c3[2] = {5, 6};
  const double a = 1.0000001, b = 0.9999999;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 12; ++u) { dmma_v(c0, a, b); dmma_v(c1, a, b); dmma_v(c2, a, b); dmma_v(c3, a, b); }
  }
  if (c0[0] + c1[0] + c2[0] + c3[0] == 1.2345e-300) out[0] = c0[1] + c1[1] + c2[1] + c3[1];
}

template <int VAR>
static void run(const double* dpack, const double* dX, int tiles, double* dout, size_t smem, const char* name) {
  CK(cudaFuncSetAttribute(k_p1<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_p1<VAR><<<148, 256, smem>>>(dpack, dX, tiles, dout);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k_p1<VAR><<<148, 256, smem>>>(dpack, dX, tiles, dout);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pairs = 148.0 * tiles * 128 * K;
  const double dmma = pairs * 6 / 8;
  const double tf = dmma * 512 / (best * 1e-3) * 1e-12;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"ns_per_pair\": %.5f, \"dmma_tflops\": %.3f, \"frac_of_37.16\": %.3f}\n", name, best, best * 1e6 / pairs, tf, tf / 37.16);
}

// NW warps per block, GPW 8-point groups per warp (tile = NW*8*GPW points): does more warps per scheduler fill the gaps?
template <int NW, int GPW>
__global__ void __launch_bounds__(NW * 32, 1) k_p1w(const double* __restrict__ pack, const double* __restrict__ X, int tiles, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  double* sp = smem;
  double* tile = smem + K * PSF;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < K * PSF; i += NW * 32) sp[i] = pack[i];
  __syncthreads();
  constexpr int TP = NW * 8 * GPW;
  double acc = 0;
  for (int bt = 0; bt < tiles; ++bt) {
    double xa[GPW][DM];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp)
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) xa[grp][mb] = X[(((size_t)(blockIdx.x * tiles + bt) * TP + warp * 8 * GPW + 8 * grp + g) % (148u * 200u * 128u)) * 16 + 4 * mb + t];
    double* row[GPW];
    double rmax[GPW];
    int rarg[GPW];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) { row[grp] = tile + (size_t)(warp * 8 * GPW + 8 * grp + g) * S; rmax[grp] = -1e300; rarg[grp] = 0; }
    for (int k4 = 0; k4 < K4; ++k4) {
      double part[GPW][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* pk = sp + (size_t)(4 * k4 + j) * PSF;
        double Lf[NB];
        double2 nb2[DN];
#pragma unroll
        for (int b = 0; b < NB; ++b) Lf[b] = pk[b * 32 + lane];
#pragma unroll
        for (int nb = 0; nb < DN; ++nb) nb2[nb] = *reinterpret_cast<const double2*>(pk + OFFB + 8 * nb + 2 * t);
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp) {
          double c[DN][2];
#pragma unroll
          for (int nb = 0; nb < DN; ++nb) { c[nb][0] = nb2[nb].x; c[nb][1] = nb2[nb].y; }
#pragma unroll
          for (int mb = 0; mb < DM; ++mb)
#pragma unroll
            for (int nb = 0; nb <= mb / 2; ++nb) dmma_v(c[nb], xa[grp][mb], Lf[blk_off(mb) + nb]);
          double s2 = c[0][0] * c[0][0];
          s2 = fma(c[0][1], c[0][1], s2);
#pragma unroll
          for (int nb = 1; nb < DN; ++nb) { s2 = fma(c[nb][0], c[nb][0], s2); s2 = fma(c[nb][1], c[nb][1], s2); }
          part[grp][j] = s2;
        }
      }
      const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
#pragma unroll
      for (int grp = 0; grp < GPW; ++grp) reduce_store(part[grp], ck, row[grp], k4, t, rmax[grp], rarg[grp]);
    }
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) acc += rmax[grp] + rarg[grp];
  }
  if (acc == 1.2345e-300) out[0] = acc + tile[tid];
}

template <int NW, int GPW>
static void runw(const double* dpack, const double* dX, int tiles, double* dout, const char* name) {
  const size_t smem = ((size_t)K * PSF + (size_t)NW * 8 * GPW * S) * 8;
  CK(cudaFuncSetAttribute(k_p1w<NW, GPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_p1w<NW, GPW><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k_p1w<NW, GPW><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pairs = 148.0 * tiles * NW * 8 * GPW * K;
  const double tf = pairs * 6 / 8 * 512 / (best * 1e-3) * 1e-12;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"ns_per_pair\": %.5f, \"dmma_tflops\": %.3f, \"frac_of_37.16\": %.3f}\n", name, best, best * 1e6 / pairs, tf, tf / 37.16);
}

// Warp pairs: NW warps, warp (pi, ch) takes the 16 points of pair pi against half ch of the components (tile = NW*8 points):
// the work per warp and the shared-memory footprint of "NW warps x 8 points", with each fragment load feeding two DMMAs.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_p1h(const double* __restrict__ pack, const double* __restrict__ X, int tiles, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  double* sp = smem;
  double* tile = smem + K * PSF;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int pi = warp >> 1, ch = warp & 1;
  for (int i = tid; i < K * PSF; i += NW * 32) sp[i] = pack[i];
  __syncthreads();
  constexpr int TP = NW * 8;
  double acc = 0;
  for (int bt = 0; bt < tiles; ++bt) {
    double xa[2][DM];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp)
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) xa[grp][mb] = X[(((size_t)(blockIdx.x * tiles + bt) * TP + pi * 16 + 8 * grp + g) % (148u * 200u * 128u)) * 16 + 4 * mb + t];
    double* row[2];
    double rmax[2];
    int rarg[2];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) { row[grp] = tile + (size_t)(pi * 16 + 8 * grp + g) * S; rmax[grp] = -1e300; rarg[grp] = 0; }
    for (int k4 = ch * (K4 / 2); k4 < (ch + 1) * (K4 / 2); ++k4) {
      double part[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) comp_pair<true>(sp + (size_t)(4 * k4 + j) * PSF, lane, t, xa, part[0][j], part[1][j]);
      const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
      reduce_store(part[0], ck, row[0], k4, t, rmax[0], rarg[0]);
      reduce_store(part[1], ck, row[1], k4, t, rmax[1], rarg[1]);
    }
    acc += rmax[0] + rmax[1] + rarg[0] + rarg[1];
  }
  if (acc == 1.2345e-300) out[0] = acc + tile[tid];
}
template <int NW>
static void runh(const double* dpack, const double* dX, int tiles, double* dout, const char* name) {
  const size_t smem = ((size_t)K * PSF + (size_t)NW * 8 * S) * 8;
  CK(cudaFuncSetAttribute(k_p1h<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_p1h<NW><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k_p1h<NW><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pairs = 148.0 * tiles * NW * 8 * K;
  const double tf = pairs * 6 / 8 * 512 / (best * 1e-3) * 1e-12;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"ns_per_pair\": %.5f, \"dmma_tflops\": %.3f, \"frac_of_37.16\": %.3f}\n", name, best, best * 1e6 / pairs, tf, tf / 37.16);
}

// Ablations of k_p1w: NOLDS keeps one component's fragments in registers (no shared-memory loads in the loop); NOTAIL drops
// the quad reduction / logit store.  Where do the 23% between the whitening loop and the pure DMMA stream go?
template <int NW, int GPW, bool NOLDS, bool NOTAIL>
__global__ void __launch_bounds__(NW * 32, 1) k_p1x(const double* __restrict__ pack, const double* __restrict__ X, int tiles, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  double* sp = smem;
  double* tile = smem + K * PSF;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < K * PSF; i += NW * 32) sp[i] = pack[i];
  __syncthreads();
  constexpr int TP = NW * 8 * GPW;
  double acc = 0;
  for (int bt = 0; bt < tiles; ++bt) {
    double xa[GPW][DM];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp)
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) xa[grp][mb] = X[(((size_t)(blockIdx.x * tiles + bt) * TP + warp * 8 * GPW + 8 * grp + g) % (148u * 200u * 128u)) * 16 + 4 * mb + t];
    double* row[GPW];
    double rmax[GPW];
    int rarg[GPW];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) { row[grp] = tile + (size_t)(warp * 8 * GPW + 8 * grp + g) * S; rmax[grp] = -1e300; rarg[grp] = 0; }
    double Lr[NB];
    double2 nr2[DN];
#pragma unroll
    for (int b = 0; b < NB; ++b) Lr[b] = sp[b * 32 + lane];
#pragma unroll
    for (int nb = 0; nb < DN; ++nb) nr2[nb] = *reinterpret_cast<const double2*>(sp + OFFB + 8 * nb + 2 * t);
    for (int k4 = 0; k4 < K4; ++k4) {
      double part[GPW][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* pk = sp + (size_t)(4 * k4 + j) * PSF;
        double Lf[NB];
        double2 nb2[DN];
#pragma unroll
        for (int b = 0; b < NB; ++b) Lf[b] = NOLDS ? __hiloint2double(__double2hiint(Lr[b]), __double2loint(Lr[b]) ^ (4 * k4 + j)) : pk[b * 32 + lane];
#pragma unroll
        for (int nb = 0; nb < DN; ++nb) nb2[nb] = NOLDS ? nr2[nb] : *reinterpret_cast<const double2*>(pk + OFFB + 8 * nb + 2 * t);
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp) {
          double c[DN][2];
#pragma unroll
          for (int nb = 0; nb < DN; ++nb) { c[nb][0] = nb2[nb].x; c[nb][1] = nb2[nb].y; }
#pragma unroll
          for (int mb = 0; mb < DM; ++mb)
#pragma unroll
            for (int nb = 0; nb <= mb / 2; ++nb) dmma_v(c[nb], xa[grp][mb], Lf[blk_off(mb) + nb]);
          double s2 = c[0][0] * c[0][0];
          s2 = fma(c[0][1], c[0][1], s2);
#pragma unroll
          for (int nb = 1; nb < DN; ++nb) { s2 = fma(c[nb][0], c[nb][0], s2); s2 = fma(c[nb][1], c[nb][1], s2); }
          part[grp][j] = s2;
        }
      }
      if (NOTAIL) {
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp) rmax[grp] += (part[grp][0] + part[grp][1]) + (part[grp][2] + part[grp][3]);
      } else {
        const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp) reduce_store(part[grp], ck, row[grp], k4, t, rmax[grp], rarg[grp]);
      }
    }
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) acc += rmax[grp] + rarg[grp];
  }
  if (acc == 1.2345e-300) out[0] = acc + tile[tid];
}

template <int NW, int GPW, bool NOLDS, bool NOTAIL>
static void runx(const double* dpack, const double* dX, int tiles, double* dout, const char* name) {
  const size_t smem = ((size_t)K * PSF + (size_t)NW * 8 * GPW * S) * 8;
  CK(cudaFuncSetAttribute(k_p1x<NW, GPW, NOLDS, NOTAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_p1x<NW, GPW, NOLDS, NOTAIL><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k_p1x<NW, GPW, NOLDS, NOTAIL><<<148, NW * 32, smem>>>(dpack, dX, tiles, dout);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pairs = 148.0 * tiles * NW * 8 * GPW * K;
  const double tf = pairs * 6 / 8 * 512 / (best * 1e-3) * 1e-12;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"ns_per_pair\": %.5f, \"dmma_tflops\": %.3f, \"frac_of_37.16\": %.3f}\n", name, best, best * 1e6 / pairs, tf, tf / 37.16);
}

// DMMA stream whose operands differ per instruction (NA distinct A registers, NBR distinct B registers, 8 accumulator chains):
// does operand variety (register-file bandwidth / bank conflicts) lower the DMMA rate below the same-operand peak?
template <int NA, int NBR>
__global__ void __launch_bounds__(256, 1) k_dmma_var(int iters, const double* __restrict__ src, double* out) {
  double a[NA], b[NBR], c[8][2];
#pragma unroll
  for (int i = 0; i < NA; ++i) a[i] = src[threadIdx.x + 256 * i];
#pragma unroll
  for (int i = 0; i < NBR; ++i) b[i] = src[threadIdx.x + 256 * (NA + i)];
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j][0] = j; c[j][1] = threadIdx.x * 1e-9; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 6; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma_v(c[j], a[(u * 8 + j) % NA], b[(u * 3 + j) % NBR]);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 1.2345e-300) out[0] = s;
}
template <int NA, int NBR>
static void run_var(const double* dsrc, double* dout, const char* name) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000;
  k_dmma_var<NA, NBR><<<148, 256>>>(iters, dsrc, dout); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); k_dmma_var<NA, NBR><<<148, 256>>>(iters, dsrc, dout); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double tf = 148.0 * 8 * iters * 48 * 512 / (ms * 1e-3) * 1e-12;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"dmma_tflops\": %.3f}\n", name, ms, tf);
}

int main() {
  const int tiles = 200;
  std::vector<double> hp((size_t)K * PSF), hx((size_t)148 * tiles * 128 * 16);
  for (size_t i = 0; i < hp.size(); ++i) hp[i] = 0.01 * ((i * 2654435761u) % 1000) / 1000.0;
  for (size_t i = 0; i < hx.size(); ++i) hx[i] = ((i * 40503u) % 2000) / 1000.0 - 1.0;
  double *dp, *dx, *dout;
  CK(cudaMalloc(&dp, hp.size() * 8)); CK(cudaMalloc(&dx, hx.size() * 8)); CK(cudaMalloc(&dout, 8));
  CK(cudaMemcpy(dp, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 8, cudaMemcpyHostToDevice));
  const size_t smem = ((size_t)K * PSF + 128 * S) * 8;
  run<0>(dp, dx, tiles, dout, smem, "0 current");
  run<1>(dp, dx, tiles, dout, smem, "1 k4 unroll 2");
  run<2>(dp, dx, tiles, dout, smem, "2 pipelined tail");
  run<3>(dp, dx, tiles, dout, smem, "3 non-volatile asm");
  run<4>(dp, dx, tiles, dout, smem, "4 pipelined tail + non-volatile");
  runw<8, 2>(dp, dx, tiles, dout, "8 warps x 16 points (= variant 0 structure)");
  runw<16, 1>(dp, dx, tiles, dout, "16 warps x 8 points");
  runw<12, 2>(dp, dx, tiles, dout, "12 warps x 16 points");
  runw<12, 1>(dp, dx, tiles, dout, "12 warps x 8 points");
  runw<8, 1>(dp, dx, tiles, dout, "8 warps x 8 points");
  runh<16>(dp, dx, tiles, dout, "16 warps as 8 pairs: 16 points x half the components per warp");
  runh<12>(dp, dx, tiles, dout, "12 warps as 6 pairs: 16 points x half the components per warp");
  runx<16, 1, true, false>(dp, dx, tiles, dout, "16 warps x 8 points, fragments in registers (no LDS)");
  runx<16, 1, false, true>(dp, dx, tiles, dout, "16 warps x 8 points, no quad reduction / store");
  runx<16, 1, true, true>(dp, dx, tiles, dout, "16 warps x 8 points, no LDS, no reduction");
  runx<12, 2, true, false>(dp, dx, tiles, dout, "12 warps x 16 points, fragments in registers (no LDS)");
  runx<12, 2, false, true>(dp, dx, tiles, dout, "12 warps x 16 points, no quad reduction / store");
  runx<12, 2, true, true>(dp, dx, tiles, dout, "12 warps x 16 points, no LDS, no reduction");
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    k_dmma_ref<<<148, 256>>>(iters, dout); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_dmma_ref<<<148, 256>>>(iters, dout); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 148.0 * 8 * iters * 48 * 512 / (ms * 1e-3) * 1e-12;
    printf("{\"variant\": \"pure DMMA, 8 warps/SM, 4 chains\", \"ms\": %.4f, \"dmma_tflops\": %.3f}\n", ms, tf);
  }
  run_var<1, 1>(dx, dout, "DMMA 8 chains, 1 A x 1 B operand register");
  run_var<8, 6>(dx, dout, "DMMA 8 chains, 8 A x 6 B operand registers (as the whitening loop)");
  run_var<16, 12>(dx, dout, "DMMA 8 chains, 16 A x 12 B operand registers");
  return 0;
}
