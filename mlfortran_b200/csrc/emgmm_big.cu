// Large-shape kernels of the EM-GMM path for sm_100a: 32 < d <= 128 (BASELINE config 4: d = 128, K = 32), or any
// (d, K) whose parameter pack does not fit shared memory next to the logit tile (e.g. d = 8, K = 256).
// Reference: mlf_EvalGaussian / mlf_MaxGaussian (src/unsupervised/mlf_gaussian.f90:41-79), GECovMatrix
// (src/mlf_matrix.f90:99-147).
//
//   k_estep_big : same arithmetic as the fused kernel's P1-P3 (whitening u = L^T(x-m) - L^T(mu-m) as DMMA tiles whose
//                 accumulators start at the bias, log-sum-exp, pass-A statistics), but the Cholesky fragments of one
//                 component (69.6 KB at d = 128) are streamed from L2 through L1 instead of living in shared memory,
//                 8 points per warp, 64 points per block tile.
//   k_mstep_big : thresholded centred scatter sum g (x-mu)(x-mu)^T with the lower-triangular 8x8 output blocks
//                 (136 at d = 128) distributed over the 8 warps of a block; a 32-point strip of (x - mu) is staged in
//                 shared memory, strips without a kept point are skipped.
// The responsibilities are stored (two-pass iteration); the fused single sweep covers d <= 16.
#include "emgmm_kernels.cuh"

#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace mlfb200 {

namespace {

__host__ __device__ constexpr int b_blk_off(int mb) {
  int s = 0;
  for (int m = 0; m < mb; ++m) s += m / 2 + 1;
  return s;
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
constexpr unsigned kFull = 0xffffffffu;

// same branch-free exp as the fused kernel (see emgmm_fused.cu)
__device__ __forceinline__ double exp_bf(double a) {
  a = (a < -750.0) ? -750.0 : a;
  a = (a > 710.0) ? 710.0 : a;
  const double magic = 6755399441055744.0;
  double t = fma(a, 1.4426950408889634074, magic);
  const int n = __double2loint(t);
  t -= magic;
  double r = fma(t, -6.93147180369123816490e-01, a);
  r = fma(t, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821614599e-10;
  p = fma(p, r, 2.0876756987868098979e-09);
  p = fma(p, r, 2.5052108385441718775e-08);
  p = fma(p, r, 2.7557319223985890653e-07);
  p = fma(p, r, 2.7557319223985892510e-06);
  p = fma(p, r, 2.4801587301587301566e-05);
  p = fma(p, r, 1.9841269841269841253e-04);
  p = fma(p, r, 1.3888888888888889419e-03);
  p = fma(p, r, 8.3333333333333332177e-03);
  p = fma(p, r, 4.1666666666666664354e-02);
  p = fma(p, r, 1.6666666666666665741e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int n1 = n >> 1, n2 = n - n1;
  return (p * __hiloint2double((n1 + 1023) << 20, 0)) * __hiloint2double((n2 + 1023) << 20, 0);
}

constexpr int kBigTile = 64;  // points per block tile (8 per warp)

}  // namespace

size_t estep_big_smem(const Geom& g) {
  const int XS = 8 * g.DN + 4;
  const int S = ((g.K8 * 8 + 15) / 16) * 16 + 4;
  return ((size_t)kBigTile * XS + (size_t)kBigTile * S) * sizeof(double);
}
int estep_big_cb(const Geom& g) { return (g.K8 + 7) / 8; }
bool estep_big_ok(const Geom& g, size_t smem_limit) {
  const bool inst = (g.DM == 1 || g.DM == 2 || g.DM == 4 || g.DM == 8 || g.DM == 12 || g.DM == 16 || g.DM == 24 || g.DM == 32);
  return inst && g.K8 * 8 <= 256 && estep_big_smem(g) <= smem_limit;
}
// component blocks per warp whose pass-A accumulators stay in registers for the whole sweep (1, 2 or 4), or 0: the generic
// variant, which accumulates one component block at a time and adds it to the block's partial slot after every tile
static int estep_big_cbt(const Geom& g) {
  const int cb = estep_big_cb(g);
  return (cb * g.DN <= 16 && (cb == 1 || cb == 2 || cb == 4)) ? cb : 0;
}

// pack2 layout per component (FusedGeom): [NB*32 L fragments | 8*DN bias -L^T(mu-m) | log(lambda)-lnDet, pad]
// GPW: 8-point groups per warp.  GPW = 2 (128-point tiles, large d) halves the fragment traffic per DMMA; its output
// coordinates are then processed in H = 2 halves so that the accumulators (GPW x DN/H x 2 doubles) stay in registers.
// STAGE: the component's pack is copied into shared memory with cp.async one component ahead of the DMMAs (two buffers, one
// block barrier per component) instead of being read through L1 by every warp; the raw points are then not kept in shared
// memory (the pass-A statistics re-read them from L2), which is what makes room for the two buffers at d = 128.
template <int DM, int CB, int GPW, bool STAGE>
__global__ void __launch_bounds__(256, 1) k_estep_big(EmArgs A) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int NB = b_blk_off(DM);
  constexpr int OFFB = NB * 32, OFFC = OFFB + 8 * DN, PSF = OFFC + 4;
  constexpr int XS = 8 * DN + 4;
  constexpr int DS = 8 * DN + 2;
  constexpr int TP = 64 * GPW;
  constexpr int H = (GPW * DN > 16) ? 2 : 1;
  constexpr int DH = (DN + H - 1) / H;   // output blocks per half
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, S = A.S, d = A.d, K = A.K;
  const int64_t N = A.N;
  double* xs = smem;                                           // [TP][XS] raw points of the tile (not with STAGE)
  double* tile = STAGE ? smem : xs + (size_t)TP * XS;          // [TP][S]  logits -> responsibilities
  double* fbuf = tile + (size_t)TP * S;                        // STAGE: [2][PSF] pack of the current / next component
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  auto stage = [&](int k) {   // all threads: 16-byte chunks of component k's pack into buffer k & 1
    const char* src = reinterpret_cast<const char*>(A.pack2 + (size_t)k * PSF);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(fbuf + (size_t)(k & 1) * PSF);
    for (int i = tid; i < PSF / 2; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * (uint32_t)i), "l"(src + 16 * (size_t)i) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  constexpr int CBX = CB > 0 ? CB : 1;
  double accA[CBX][DN][2], sg[CBX], sg2[CBX];
#pragma unroll
  for (int c = 0; c < CBX; ++c) {
    sg[c] = 0; sg2[c] = 0;
#pragma unroll
    for (int b = 0; b < DN; ++b) { accA[c][b][0] = 0; accA[c][b][1] = 0; }
  }
  double* outA = A.partA + (size_t)blockIdx.x * KP * DS;
  if (CB == 0) {   // generic variant: this warp's component blocks start from zero in the block's partial slot
    for (int cb = warp; cb < K8; cb += 8) {
      double* o = outA + (size_t)(8 * cb + g) * DS;
      if (t == 0) { o[0] = 0.0; o[1] = 0.0; }
#pragma unroll
      for (int db = 0; db < DN; ++db) { o[2 + 8 * db + 2 * t] = 0.0; o[2 + 8 * db + 2 * t + 1] = 0.0; }
    }
  }
  const int64_t ntiles = (N + TP - 1) / TP;
  for (int64_t bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {
    const int64_t p0 = bt * TP + warp * 8 * GPW + g;
    // ---- P1: logits of this warp's points against every component ---------------------------------------------
    double xa[GPW][DM];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) {
      const int64_t p = p0 + 8 * grp;
      double* xr = xs + (size_t)(warp * 8 * GPW + 8 * grp + g) * XS;
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) {
        const int a = 4 * mb + t;
        const double x = (p < N && a < d) ? A.X[p * d + a] : 0.0;
        if (!STAGE) xr[a] = x;
        xa[grp][mb] = x - A.shift[a];
      }
      if (!STAGE && 8 * DN > 4 * DM) xr[4 * DM + t] = 0.0;
    }
    double* row[GPW];
    double rmax[GPW];
    int rarg[GPW];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) { row[grp] = tile + (size_t)(warp * 8 * GPW + 8 * grp + g) * S; rmax[grp] = -INFINITY; rarg[grp] = 0; }
    if (STAGE) stage(0);   // the previous tile ended with a block barrier: both buffers are free
    for (int k = 0; k < KP; ++k) {
      const double* pk = A.pack2 + (size_t)k * PSF;
      if (STAGE) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();   // component k has landed for every thread, and every warp is done with component k - 1
        if (k + 1 < KP) stage(k + 1);
        pk = fbuf + (size_t)(k & 1) * PSF;
      }
      double s[GPW];
#pragma unroll
      for (int grp = 0; grp < GPW; ++grp) s[grp] = 0;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const int lo = h * DH;
        double c[GPW][DH][2];
#pragma unroll
        for (int nb = 0; nb < DH; ++nb) {
          double2 v = make_double2(0.0, 0.0);
          if (lo + nb < DN) v = *reinterpret_cast<const double2*>(pk + OFFB + 8 * (lo + nb) + 2 * t);
#pragma unroll
          for (int grp = 0; grp < GPW; ++grp) { c[grp][nb][0] = v.x; c[grp][nb][1] = v.y; }
        }
#pragma unroll
        for (int mb = 0; mb < DM; ++mb)
#pragma unroll
          for (int nb = 0; nb < DH; ++nb)
            if (lo + nb < DN && lo + nb <= mb / 2) {
              const double Lf = STAGE ? pk[(b_blk_off(mb) + lo + nb) * 32 + lane] : __ldg(pk + (b_blk_off(mb) + lo + nb) * 32 + lane);
#pragma unroll
              for (int grp = 0; grp < GPW; ++grp) dmma(c[grp][nb], xa[grp][mb], Lf);
            }
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp)
#pragma unroll
          for (int nb = 0; nb < DH; ++nb) { s[grp] = fma(c[grp][nb][0], c[grp][nb][0], s[grp]); s[grp] = fma(c[grp][nb][1], c[grp][nb][1], s[grp]); }
      }
      const double ck = pk[OFFC];
#pragma unroll
      for (int grp = 0; grp < GPW; ++grp) {
        double v = s[grp];
        v += shx(v, 1);
        v += shx(v, 2);
        const double logit = fma(-0.5, v, ck);
        if (t == 0) row[grp][k] = logit;
        if (logit > rmax[grp]) { rmax[grp] = logit; rarg[grp] = k; }  // k ascending: first maximum wins (MAXLOC)
      }
    }
    __syncwarp();
    // ---- P2: log-sum-exp over the components; lane t handles components t, t+4, ... ----------------------------
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) {
      const int64_t p = p0 + 8 * grp;
      const bool valid = p < N;
      double* rw = row[grp];
      if (A.hard) {  // k-means assignment: one-hot of the nearest centre (see emgmm_fused.cu)
        for (int k = t; k < KP; k += 4) rw[k] = (valid && k == rarg[grp]) ? 1.0 : 0.0;
      } else {
        const double mx = A.raw ? 0.0 : rmax[grp];
        double sum = 0;
#pragma unroll 4
        for (int k = t; k < KP; k += 4) {
          const double e = exp_bf(rw[k] - mx);
          rw[k] = e;
          sum += e;
        }
        sum += shx(sum, 1);
        sum += shx(sum, 2);
        const double inv = valid ? (A.raw ? 1.0 : 1.0 / sum) : 0.0;
        for (int k = t; k < KP; k += 4) {
          const double gk = rw[k] * inv;
          rw[k] = gk;
          if (A.gamma != nullptr && valid && k < K) A.gamma[(size_t)k * A.ldg + p] = gk;
        }
      }
      if (A.labels != nullptr && valid && t == 0) A.labels[p] = rarg[grp] + 1;
    }
    __syncthreads();
    // ---- P3: pass-A statistics; warp w owns component blocks w, w+8, ... -----------------------------------------
    if (CB == 0) {
      // any number of component blocks: one block's accumulators at a time, added to the block's own partial slot after
      // the tile (plain read-modify-write by the owning warp; at these shapes one tile's P1 takes milliseconds)
      for (int cb = warp; cb < K8; cb += 8) {
        double a1 = 0, a2 = 0, acc[DN][2];
#pragma unroll
        for (int db = 0; db < DN; ++db) { acc[db][0] = 0; acc[db][1] = 0; }
#pragma unroll 2
        for (int pb = 0; pb < TP / 4; ++pb) {
          const double a = tile[(size_t)(4 * pb + t) * S + 8 * cb + g];
          a1 += a;
          a2 = fma(a, a, a2);
#pragma unroll
          for (int db = 0; db < DN; ++db) dmma(acc[db], a, xs[(size_t)(4 * pb + t) * XS + 8 * db + g]);
        }
        a1 += shx(a1, 1); a1 += shx(a1, 2);
        a2 += shx(a2, 1); a2 += shx(a2, 2);
        double* o = outA + (size_t)(8 * cb + g) * DS;
        if (t == 0) { o[0] += a1; o[1] += a2; }
#pragma unroll
        for (int db = 0; db < DN; ++db) {
          o[2 + 8 * db + 2 * t] += acc[db][0];
          o[2 + 8 * db + 2 * t + 1] += acc[db][1];
        }
      }
    }
#pragma unroll
    for (int ci = 0; ci < CB; ++ci) {
      const int cb = warp + 8 * ci;
      if (cb < K8) {
#pragma unroll 2
        for (int pb = 0; pb < TP / 4; ++pb) {
          const double a = tile[(size_t)(4 * pb + t) * S + 8 * cb + g];
          sg[ci] += a;
          sg2[ci] = fma(a, a, sg2[ci]);
#pragma unroll
          for (int db = 0; db < DN; ++db) {
            double xv;
            if (STAGE) {   // rows of the tile that lie beyond N carry a = 0
              const int64_t pp = bt * TP + 4 * pb + t;
              xv = (pp < N && 8 * db + g < d) ? A.X[pp * d + 8 * db + g] : 0.0;
            } else {
              xv = xs[(size_t)(4 * pb + t) * XS + 8 * db + g];
            }
            dmma(accA[ci][db], a, xv);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ci = 0; ci < CB; ++ci) {
    const int cb = warp + 8 * ci;
    if (cb < K8) {
      double a = sg[ci], b = sg2[ci];
      a += shx(a, 1); a += shx(a, 2);
      b += shx(b, 1); b += shx(b, 2);
      double* o = outA + (size_t)(8 * cb + g) * DS;
      if (t == 0) { o[0] = a; o[1] = b; }
#pragma unroll
      for (int db = 0; db < DN; ++db) {
        o[2 + 8 * db + 2 * t] = accA[ci][db][0];
        o[2 + 8 * db + 2 * t + 1] = accA[ci][db][1];
      }
    }
  }
}

// 128-point tiles (two groups per warp) where d is large and the tiles fit shared memory; 64-point tiles otherwise
int estep_big_gpw(const Geom& g, size_t smem_limit) {
  const int XS = 8 * g.DN + 4;
  const int S = ((g.K8 * 8 + 15) / 16) * 16 + 4;
  const size_t need2 = ((size_t)128 * XS + (size_t)128 * S) * sizeof(double);
  return (g.DM >= 16 && estep_big_cbt(g) == 1 && need2 <= smem_limit) ? 2 : 1;
}

template <int DM, int CB, int GPW, bool STAGE = false>
static cudaError_t launch_big_t(const EmArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_estep_big<DM, CB, GPW, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_estep_big<DM, CB, GPW, STAGE><<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}
// staged variant: 128-point logit tile + two pack buffers
size_t estep_big_stage_smem(const Geom& g) {
  const int S = ((g.K8 * 8 + 15) / 16) * 16 + 4;
  const int PSF = g.NB * 32 + 8 * g.DN + 4;
  return ((size_t)128 * S + (size_t)2 * PSF) * sizeof(double);
}
bool estep_big_stage(const Geom& g, size_t smem_limit) {
  const char* e = getenv("MLF_B200_BIG_STAGE");   // read per call: the tests flip it inside one process
  const bool off = e && e[0] == '0';
  return !off && estep_big_gpw(g, smem_limit) == 2 && estep_big_stage_smem(g) <= smem_limit;
}
template <int DM>
static cudaError_t launch_big_dm(const EmArgs& a, int cb, int gpw, bool stage, int grid, size_t smem, cudaStream_t st) {
  constexpr int DN = (DM + 1) / 2;
  if constexpr (DM >= 16) {
    if (gpw == 2 && cb == 1 && stage) return launch_big_t<DM, 1, 2, true>(a, grid, smem, st);
    if (gpw == 2 && cb == 1) return launch_big_t<DM, 1, 2>(a, grid, smem, st);
  }
  if constexpr (1 * DN <= 16) { if (cb == 1) return launch_big_t<DM, 1, 1>(a, grid, smem, st); }
  if constexpr (2 * DN <= 16) { if (cb == 2) return launch_big_t<DM, 2, 1>(a, grid, smem, st); }
  if constexpr (4 * DN <= 16) { if (cb == 4) return launch_big_t<DM, 4, 1>(a, grid, smem, st); }
  if (cb == 0) return launch_big_t<DM, 0, 1>(a, grid, smem, st);
  return cudaErrorInvalidConfiguration;
}
cudaError_t launch_estep_big(const Geom& g, EmArgs a, int grid, size_t smem_limit, cudaStream_t st) {
  if (!estep_big_ok(g, smem_limit)) return cudaErrorInvalidConfiguration;
  const int gpw = estep_big_gpw(g, smem_limit);
  a.S = ((g.K8 * 8 + 15) / 16) * 16 + 4;
  const bool stage = estep_big_stage(g, smem_limit);
  const size_t smem = stage ? estep_big_stage_smem(g) : ((size_t)64 * gpw * (8 * g.DN + 4) + (size_t)64 * gpw * a.S) * sizeof(double);
  const int cb = estep_big_cbt(g);
  switch (g.DM) {
    case 1: return launch_big_dm<1>(a, cb, gpw, stage, grid, smem, st);
    case 2: return launch_big_dm<2>(a, cb, gpw, stage, grid, smem, st);
    case 4: return launch_big_dm<4>(a, cb, gpw, stage, grid, smem, st);
    case 8: return launch_big_dm<8>(a, cb, gpw, stage, grid, smem, st);
    case 12: return launch_big_dm<12>(a, cb, gpw, stage, grid, smem, st);
    case 16: return launch_big_dm<16>(a, cb, gpw, stage, grid, smem, st);
    case 24: return launch_big_dm<24>(a, cb, gpw, stage, grid, smem, st);
    case 32: return launch_big_dm<32>(a, cb, gpw, stage, grid, smem, st);
    default: return cudaErrorInvalidConfiguration;
  }
}
int estep_big_tile(const Geom& g, size_t smem_limit) { return 64 * estep_big_gpw(g, smem_limit); }

// Row / column of lower-triangular block number blk (row-major over the triangle: 0 -> (0,0), 1 -> (1,0), 2 -> (1,1), ...).
__host__ __device__ constexpr int tri_row(int blk) {
  int ab = 0;
  while ((ab + 1) * (ab + 2) / 2 <= blk) ++ab;
  return ab;
}
// Scatter of `ng` staged 4-point groups into warp W's output blocks W, W + 8, ...: the 8-coordinate blocks of (x - mu) that
// those output blocks touch are loaded once per group (compile-time set, <= DN loads) and reused by every DMMA; the
// weighted copy w * z feeds the row operand (same product as the reference's W(i) * (X - Mu) outer product).
template <int DN, int W>
__device__ __forceinline__ void mb_scatter(double (&acc)[(DN * (DN + 1) / 2 + 7) / 8][2], const double* __restrict__ zb,
                                           const double* __restrict__ wsb, int ng, int g, int t) {
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int OWN = (MB + 7) / 8;
  constexpr int ZS = 8 * DN + 4;
  for (int q = 0; q < ng; ++q) {
    const double w = wsb[4 * q + t];
    const double* zr = zb + (size_t)(4 * q + t) * ZS + g;
    double z[DN], wz[DN];
#pragma unroll
    for (int b = 0; b < DN; ++b) {
      bool need = false;   // does any of this warp's blocks use coordinate block b?
#pragma unroll
      for (int o = 0; o < OWN; ++o) {
        const int blk = W + 8 * o;
        if (blk < MB) { const int ab = tri_row(blk), bb = blk - ab * (ab + 1) / 2; need = need || ab == b || bb == b; }
      }
      z[b] = need ? zr[8 * b] : 0.0;
      wz[b] = w * z[b];
    }
#pragma unroll
    for (int o = 0; o < OWN; ++o) {
      const int blk = W + 8 * o;
      if (blk < MB) {
        const int ab = tri_row(blk), bb = blk - ab * (ab + 1) / 2;
        dmma(acc[o], wz[ab], z[bb]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// M-pass for large d.  grid = (point chunks, components); block = 8 warps; warp w owns the lower-triangular 8x8 output
// blocks w, w+8, ... of the scatter matrix (<= 17 at d = 128).  partial[chunk][K][MS] as k_mstep.
// gamma == nullptr: unit weights (global moments).
// ---------------------------------------------------------------------------------------------------------------
template <int DN>
__global__ void __launch_bounds__(256) k_mstep_big(const double* __restrict__ X, int64_t N, int d, int K,
                                                   const double* __restrict__ gamma, int64_t ldg, const double* __restrict__ mupad,
                                                   const double* __restrict__ thr, double* __restrict__ partial, int MS,
                                                   int64_t chunk) {
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int OWN = (MB + 7) / 8;
  constexpr int ZS = 8 * DN + 4;
  extern __shared__ __align__(16) double zdyn[];       // [2][32][ZS]: (x - mu) of up to 32 kept points, double-buffered
  __shared__ double ws[2][32];         // kept weight of each staged point (0 if dropped)
  __shared__ long long pidx[2][32];    // global index of each staged point
  __shared__ double smu[8 * DN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k = blockIdx.y;
  for (int a = tid; a < 8 * DN; a += 256) smu[a] = mupad[(size_t)k * 8 * DN + a];
  const double th = thr[k];
  const double* gk = gamma ? gamma + (size_t)k * ldg : nullptr;
  double acc[OWN][2];
#pragma unroll
  for (int o = 0; o < OWN; ++o) { acc[o][0] = 0; acc[o][1] = 0; }
  double sg2 = 0;
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = (lo + chunk < N) ? lo + chunk : N;
  // scan state.  The responsibilities of the chunk are looked at in rounds of 64 strips of 32 points: warp w fetches strips
  // 8w..8w+7 (eight loads in flight per lane, 2048 points per round trip to memory), leaves the kept weights and the
  // kept-point masks in shared memory; warp 0 turns the masks into prefix counts -- everything that follows is uniform over
  // the block, and the order of the kept points is the point order (deterministic).  One warp looking at one strip at a time waited a full
  // memory latency per ~1 kept point at K = 32 (profiles/ncu_r01h_big_d128.md: the scan, not the scatter, bounded the pass).
  constexpr int SR = 64;
  __shared__ double wr[SR][32];   // kept weight of every point of the round (0 if dropped)
  __shared__ unsigned mr[SR];     // kept points of strip s
  __shared__ int pre[SR + 1];     // kept points of the round before strip s (pre[SR]: all of them)
  int64_t ib = lo - 32 * SR;      // first point of the current round
  int used = 0, total = 0;        // kept points of the round already listed / in all
  // list up to 32 kept points (compacted across strips and rounds, in point order) into buffer `buf`; returns how many.
  // The r-th kept point of a round is found by a binary search of the per-strip prefix counts and a find-n-th-set-bit on
  // the strip's mask: 32 lanes place 32 points at once, nothing walks the strips one by one.
  auto fill = [&](int buf) -> int {
    int c = 0;
    while (c < 32) {
      if (used == total) {
        ib += 32 * SR;
        if (ib >= hi) break;
        __syncthreads();   // the previous round's weights have been copied out by every warp that needed them
        double gv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t pl = ib + 32 * (8 * warp + u) + lane;
          gv[u] = (pl < hi) ? (gk ? gk[pl] : 1.0) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t pl = ib + 32 * (8 * warp + u) + lane;
          sg2 = fma(gv[u], gv[u], sg2);                                  // SUM(W*W) runs over all points (src/mlf_matrix.f90:119)
          const bool kept = !(gv[u] < th) && pl < hi && gv[u] != 0.0;    // "If(W(i) < minW0) CYCLE" (:135)
          wr[8 * warp + u][lane] = kept ? gv[u] : 0.0;
          const unsigned m = __ballot_sync(kFull, kept);
          if (lane == 0) mr[8 * warp + u] = m;
        }
        __syncthreads();
        if (warp == 0) {   // exclusive prefix counts of the 64 strips
          const int c0 = __popc(mr[2 * lane]), c1 = __popc(mr[2 * lane + 1]);
          int inc = c0 + c1;
#pragma unroll
          for (int sh = 1; sh < 32; sh <<= 1) {
            const int o = __shfl_up_sync(kFull, inc, sh);
            if (lane >= sh) inc += o;
          }
          pre[2 * lane] = inc - c0 - c1;
          pre[2 * lane + 1] = inc - c1;
          if (lane == 31) pre[SR] = inc;
        }
        __syncthreads();
        used = 0;
        total = pre[SR];
        continue;
      }
      const int take = (total - used < 32 - c) ? total - used : 32 - c;
      if (warp == 0 && lane < take) {
        const int r = used + lane;
        int s0 = 0;   // last strip with pre[s0] <= r
#pragma unroll
        for (int step = SR / 2; step > 0; step >>= 1)
          if (pre[s0 + step] <= r) s0 += step;
        const int bit = __fns(mr[s0], 0, r - pre[s0] + 1);
        ws[buf][c + lane] = wr[s0][bit];
        pidx[buf][c + lane] = ib + 32 * s0 + bit;
      }
      used += take;
      c += take;
    }
    if (warp == 0 && lane >= c) ws[buf][lane] = 0.0;   // unused slots of the last group weigh nothing
    return c;
  };
  // gather the listed points' rows into registers (DN loads per thread in flight), store them centred
  auto gather = [&](int buf, int cnt, double (&xv)[DN]) {
#pragma unroll
    for (int u = 0; u < DN; ++u) {
      const int i = tid + 256 * u;
      const int pt = i / (8 * DN), a = i % (8 * DN);
      const long long pp = pt < cnt ? pidx[buf][pt] : -1;
      xv[u] = (pp >= 0 && pp < hi && a < d) ? X[pp * d + a] - smu[a] : 0.0;
    }
  };
  auto store = [&](int buf, const double (&xv)[DN]) {
    double* zb = zdyn + (size_t)buf * 32 * ZS;
#pragma unroll
    for (int u = 0; u < DN; ++u) {
      const int i = tid + 256 * u;
      zb[(size_t)(i / (8 * DN)) * ZS + i % (8 * DN)] = xv[u];
    }
  };
  __syncthreads();
  int cur = 0;
  int cc = fill(0);
  double xv[DN];
  __syncthreads();
  gather(0, cc, xv);
  store(0, xv);
  __syncthreads();
  while (cc > 0) {
    const int cn = fill(cur ^ 1);   // list the next batch ...
    __syncthreads();
    gather(cur ^ 1, cn, xv);        // ... and start fetching its rows: they land while this batch is multiplied
    const double* zb = zdyn + (size_t)cur * 32 * ZS;
    switch (warp) {   // compile-time block map per warp: every shared-memory offset is an immediate
      case 0: mb_scatter<DN, 0>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 1: mb_scatter<DN, 1>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 2: mb_scatter<DN, 2>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 3: mb_scatter<DN, 3>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 4: mb_scatter<DN, 4>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 5: mb_scatter<DN, 5>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      case 6: mb_scatter<DN, 6>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
      default: mb_scatter<DN, 7>(acc, zb, ws[cur], (cc + 3) / 4, g, t); break;
    }
    store(cur ^ 1, xv);
    __syncthreads();                // next batch staged; this batch's lists and rows are free again
    cur ^= 1;
    cc = cn;
  }
  double* out = partial + ((size_t)blockIdx.x * K + k) * MS;
#pragma unroll
  for (int o = 0; o < OWN; ++o) {
    const int blk = warp + 8 * o;
    if (blk < MB) {
      out[(size_t)blk * 64 + g * 8 + 2 * t] = acc[o][0];
      out[(size_t)blk * 64 + g * 8 + 2 * t + 1] = acc[o][1];
    }
  }
  {
    __shared__ double sg2w[8];
    double s = sg2;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += shx(s, sh);
    if (lane == 0) sg2w[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0;
      for (int w = 0; w < 8; ++w) tot += sg2w[w];   // fixed order
      out[(size_t)MB * 64] = tot;
      out[(size_t)MB * 64 + 1] = 0.0;
    }
  }
}

void mstep_big_grid(const Geom& g, int sms, int64_t N, int* gx, int64_t* chunk) {
  int x = sms * 2 / (g.K > 0 ? g.K : 1);
  if (x < 1) x = 1;
  int64_t c = (N + x - 1) / x;
  c = ((c + 31) / 32) * 32;
  if (c < 32) c = 32;
  x = (int)((N + c - 1) / c);
  if (x < 1) x = 1;
  *gx = x;
  *chunk = c;
}

cudaError_t launch_mstep_big(const Geom& g, const double* X, int64_t N, const double* gamma, int64_t ldg, const double* mupad,
                             const double* thr, double* partial, int gx, int64_t chunk, cudaStream_t st) {
  const dim3 grid(gx, g.K);
#define MLF_MB(DNv)                                                                                                        \
  case DNv: {                                                                                                              \
    const size_t dyn = (size_t)2 * 32 * (8 * DNv + 4) * sizeof(double);                                                    \
    cudaError_t e = cudaFuncSetAttribute(k_mstep_big<DNv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);         \
    if (e != cudaSuccess) return e;                                                                                        \
    k_mstep_big<DNv><<<grid, 256, dyn, st>>>(X, N, g.d, g.K, gamma, ldg, mupad, thr, partial, g.MS, chunk);               \
  } break;
  switch (g.DN) {
    MLF_MB(1) MLF_MB(2) MLF_MB(4) MLF_MB(6) MLF_MB(8) MLF_MB(12) MLF_MB(16)
    default: return cudaErrorInvalidConfiguration;
  }
#undef MLF_MB
  return cudaGetLastError();
}

}  // namespace mlfb200
