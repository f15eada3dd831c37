// Step-0 microbenchmark: measured fp64 ceilings of one B200 (sm_100a).
// Nothing here is on the product path; it produces the roofline denominators
// (profiles/fp64_peaks_r01.json) that bench.py and DESIGN.md cite.
//
//   dfma      : register-resident DFMA chains, ILP 1..8
//   dmma      : mma.sync.aligned.m8n8k4.f64 chains, ILP 1..8
//   mixed     : DFMA and DMMA interleaved in the same warp (separate pipes?)
//   exp       : fp64 exp() throughput
//   lds_dfma  : DFMA whose second operand is a warp-broadcast LDS.128 (2 doubles / load),
//               P points per thread share each load (P = 1, 2, 4)
//   copy      : HBM copy bandwidth (cross-check of MEASURED_PEAKS.json)
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) acc[j] = threadIdx.x * 1e-9 + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) acc[j] = fma(acc[j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += acc[j];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) { c0[j] = threadIdx.x * 1e-9 + j; c1[j] = j; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < ILP; ++j) dmma884(c0[j], c1[j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += c0[j] + c1[j];
  if (s == 12345.678) out[0] = s;
}

// NF DFMA chains + NM DMMA chains interleaved in every warp.
template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
  double acc[NF], c0[NM], c1[NM];
#pragma unroll
  for (int j = 0; j < NF; ++j) acc[j] = threadIdx.x * 1e-9 + j;
#pragma unroll
  for (int j = 0; j < NM; ++j) { c0[j] = threadIdx.x * 1e-9 + j; c1[j] = j; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < NM; ++j) dmma884(c0[j], c1[j], a, b);
#pragma unroll
      for (int j = 0; j < NF; ++j) acc[j] = fma(acc[j], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NF; ++j) s += acc[j];
#pragma unroll
  for (int j = 0; j < NM; ++j) s += c0[j] + c1[j];
  if (s == 12345.678) out[0] = s;
}

// half the warps of a block run DFMA, the other half DMMA
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double a, double b) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = threadIdx.x * 1e-9 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(acc[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
  } else {
    double c0[8], c1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c0[j] = threadIdx.x * 1e-9 + j; c1[j] = j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c0[j], c1[j], a, b);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c0[j] + c1[j];
  }
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double x0) {
  double v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) v[j] = -x0 - 1e-3 * (threadIdx.x + 32 * j);
  double s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += exp(v[j]); v[j] += 1e-7; }
  }
  if (s == 12345.678) out[0] = s;
}

// DFMA fed by warp-broadcast LDS.128: coefficient table in smem, P points per thread.
template <int P>
__global__ void __launch_bounds__(256) k_lds_dfma(double* out, int iters, double z0) {
  __shared__ double2 coef[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) coef[i] = make_double2(1e-3 * i, -1e-3 * i);
  __syncthreads();
  double z[P], acc[P][2];
#pragma unroll
  for (int p = 0; p < P; ++p) { z[p] = z0 + threadIdx.x + p; acc[p][0] = 0; acc[p][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 64; ++u) {
      const double2 c = coef[(u * 8 + it) & 511];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        acc[p][0] = fma(c.x, z[p], acc[p][0]);
        acc[p][1] = fma(c.y, z[p], acc[p][1]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int p = 0; p < P; ++p) s += acc[p][0] + acc[p][1];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

static cudaEvent_t e0, e1;
template <class F>
static double time_ms(F&& launch, int reps = 5) {
  launch();  // warm-up
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const char* json_path = argc > 1 ? argv[1] : nullptr;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device: %s  SMs=%d  cc=%d.%d\n", prop.name, sms, prop.major, prop.minor);
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* out; CK(cudaMalloc(&out, 64));
  std::string js = "{";
  js += "\"gpu\": \"" + std::string(prop.name) + "\", \"sms\": " + std::to_string(sms);
  auto add = [&](const char* k, double v) { char b[128]; snprintf(b, sizeof b, ", \"%s\": %.4f", k, v); js += b; };

  const int threads = 256, iters = 4096;
  // --- DFMA ---
  for (int bps : {1, 2, 4, 8}) {
    const int grid = sms * bps;
    const double n = (double)grid * threads * iters * 8.0;
    double t1 = time_ms([&] { k_dfma<1><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t2 = time_ms([&] { k_dfma<2><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t4 = time_ms([&] { k_dfma<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t8 = time_ms([&] { k_dfma<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("dfma  blocks/SM=%d (warps/SM=%d): ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f TFLOP/s\n", bps, bps * 8,
           2 * n * 1 / t1 * 1e-9, 2 * n * 2 / t2 * 1e-9, 2 * n * 4 / t4 * 1e-9, 2 * n * 8 / t8 * 1e-9);
    if (bps == 8) add("dfma_tflops", 2 * n * 8 / t8 * 1e-9);
    if (bps == 1) { add("dfma_tflops_8warps_ilp1", 2 * n * 1 / t1 * 1e-9); add("dfma_tflops_8warps_ilp8", 2 * n * 8 / t8 * 1e-9); }
  }
  // single warp latency: 1 block of 32 threads, ILP1: cycles per dependent DFMA
  {
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    double t = time_ms([&] { k_dfma<1><<<1, 32>>>(out, 1 << 16, 1.0000001, 1e-9); });
    double cyc = t * 1e-3 * clk * 1e3 / ((1 << 16) * 8.0);
    printf("dfma dependent-issue latency ~ %.2f cycles (at nominal %d kHz)\n", cyc, clk);
    add("dfma_latency_cyc_nominal_clk", cyc);
  }
  // --- DMMA ---
  for (int bps : {1, 2, 4, 8}) {
    const int grid = sms * bps;
    const double nw = (double)grid * (threads / 32) * iters * 8.0;  // warp-level mma count per ILP
    const double f = 2.0 * 8 * 8 * 4;
    double t1 = time_ms([&] { k_dmma<1><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t2 = time_ms([&] { k_dmma<2><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t4 = time_ms([&] { k_dmma<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    double t8 = time_ms([&] { k_dmma<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("dmma  blocks/SM=%d (warps/SM=%d): ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f TFLOP/s\n", bps, bps * 8,
           f * nw * 1 / t1 * 1e-9, f * nw * 2 / t2 * 1e-9, f * nw * 4 / t4 * 1e-9, f * nw * 8 / t8 * 1e-9);
    if (bps == 8) add("dmma_tflops", f * nw * 8 / t8 * 1e-9);
  }
  {
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    double t = time_ms([&] { k_dmma<1><<<1, 32>>>(out, 1 << 14, 1.0000001, 1e-9); });
    double cyc = t * 1e-3 * clk * 1e3 / ((1 << 14) * 8.0);
    printf("dmma dependent-issue latency ~ %.2f cycles\n", cyc);
    add("dmma_latency_cyc_nominal_clk", cyc);
  }
  // --- mixed ---
  {
    const int grid = sms * 4;
    const double per = (double)grid * iters * 8.0;
    auto rep = [&](const char* name, int nf, int nm, double t) {
      const double ff = 2.0 * per * threads * nf, fm = 2.0 * 256 * per * (threads / 32) * nm;
      printf("mixed %-10s: DFMA %.2f + DMMA %.2f = %.2f TFLOP/s\n", name, ff / t * 1e-9, fm / t * 1e-9, (ff + fm) / t * 1e-9);
      char k[64]; snprintf(k, sizeof k, "mixed_f%d_m%d_tflops", nf, nm); add(k, (ff + fm) / t * 1e-9);
    };
    rep("F4+M4", 4, 4, time_ms([&] { k_mixed<4, 4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }));
    rep("F8+M1", 8, 1, time_ms([&] { k_mixed<8, 1><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }));
    rep("F8+M2", 8, 2, time_ms([&] { k_mixed<8, 2><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }));
    rep("F4+M1", 4, 1, time_ms([&] { k_mixed<4, 1><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }));
    double t = time_ms([&] { k_split<<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
    const double ff = 2.0 * per * (threads / 2) * 8, fm = 2.0 * 256 * per * (threads / 64) * 8;
    printf("split warps    : DFMA %.2f + DMMA %.2f = %.2f TFLOP/s\n", ff / t * 1e-9, fm / t * 1e-9, (ff + fm) / t * 1e-9);
    add("split_warps_tflops", (ff + fm) / t * 1e-9);
  }
  // --- exp ---
  {
    const int grid = sms * 8, it2 = 2048;
    const double n = (double)grid * threads * it2;
    double t1 = time_ms([&] { k_exp<1><<<grid, threads>>>(out, it2, 3.0); });
    double t4 = time_ms([&] { k_exp<4><<<grid, threads>>>(out, it2, 3.0); });
    printf("exp(fp64): ILP1 %.1f  ILP4 %.1f Gexp/s  (=> %.1f DFMA-equivalents each at dfma peak)\n", n / t1 * 1e-6, 4 * n / t4 * 1e-6, 0.0);
    add("exp_gps", 4 * n / t4 * 1e-6);
  }
  // --- LDS-fed DFMA ---
  {
    const int it3 = 1024;
    for (int bps : {1, 2, 4}) {
      const int grid = sms * bps;
      const double n = (double)grid * threads * it3 * 64.0 * 2.0;
      double t1 = time_ms([&] { k_lds_dfma<1><<<grid, threads>>>(out, it3, 0.5); });
      double t2 = time_ms([&] { k_lds_dfma<2><<<grid, threads>>>(out, it3, 0.5); });
      double t4 = time_ms([&] { k_lds_dfma<4><<<grid, threads>>>(out, it3, 0.5); });
      printf("lds-fed dfma blocks/SM=%d: P1 %.2f  P2 %.2f  P4 %.2f TFLOP/s\n", bps, 2 * n * 1 / t1 * 1e-9, 2 * n * 2 / t2 * 1e-9, 2 * n * 4 / t4 * 1e-9);
      if (bps == 2) { add("lds_dfma_p1_tflops", 2 * n / t1 * 1e-9); add("lds_dfma_p2_tflops", 4 * n / t2 * 1e-9); add("lds_dfma_p4_tflops", 8 * n / t4 * 1e-9); }
    }
  }
  // --- HBM copy ---
  {
    const size_t bytes = (size_t)4 << 30;
    double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    const size_t n = bytes / sizeof(double2);
    double t = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, n); }, 10);
    printf("hbm copy: %.1f GB/s (read+write)\n", 2.0 * bytes / t * 1e-6);
    add("hbm_copy_gbs", 2.0 * bytes / t * 1e-6);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  js += "}\n";
  if (json_path) { FILE* f = fopen(json_path, "w"); if (f) { fputs(js.c_str(), f); fclose(f); } }
  printf("%s", js.c_str());
  return 0;
}
