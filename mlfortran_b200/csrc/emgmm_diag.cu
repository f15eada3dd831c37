// Fused single-sweep EM kernel for DIAGONAL covariances at small d and large K (BASELINE config 5: d = 8, K = 256; an
// extension -- the reference only has full covariances) on the fp64 CUDA cores.
//
// With a diagonal covariance the whitening is d multiply-adds per (point, component), not a d x d contraction: tensor-core
// tiles would be 7/8 zeros.  So here (north_star: "fp64 FMA on the CUDA cores for small d") every lane OWNS ONE COMPONENT --
// its diagonal factor, bias and constant live in registers for the whole sweep, as do its statistics -- and the points of a
// 64-point tile are broadcast to the warp from shared memory:
//   P1  lane = component (warp w owns components 32w..32w+31): logit = c_k - 1/2 sum_a (l_ka x'_a + b_ka)^2 for every point of
//       the tile (2 DFMA per dimension), written to the shared logit tile;
//   P2  warp = 8 points: log-sum-exp over all components (branch-free exp), responsibilities in place;
//   P3  lane = component again: sum g, sum g^2, sum g x, and -- for the pairs that pass the reference's W >= minW test --
//       sum g (x-mu)^2, sum g (x-mu), sum g, all in registers; undecided pairs go to the per-warp side list.
// Partial layouts, threshold band and side list are those of emgmm_fused.cu, so the engine's protocol is unchanged.
#include "emgmm_kernels.cuh"

#include <cmath>

namespace mlfb200 {

namespace {

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double exp_bf(double a) {  // see emgmm_fused.cu
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

constexpr int kTP = 64;   // points per block tile
constexpr int kD = 8;     // dimensions held in registers (d <= 8; padding dimensions are zero)

}  // namespace

size_t emd_smem_bytes(const Geom& g) { return ((size_t)kTP * 12 + (size_t)kTP * g.S) * sizeof(double); }
bool emd_ok(const Geom& g, const FusedGeom& f, size_t smem_limit) {
  return f.diag && g.DN == 1 && g.K8 * 8 <= 256 && emd_smem_bytes(g) <= smem_limit;
}

// pack2: diagonal-compressed layout of FusedGeom (DN = 1): [8 diag(L) | 8 bias | const, pad3] = 20 doubles per component
__global__ void __launch_bounds__(256, 1) k_emd(EmArgs A) {
  constexpr int PSF = 20, XS = 12, DS = 10, MSF = 74;
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, K4 = 2 * K8, S = A.S, d = A.d;
  const int64_t N = A.N;
  double* xs = smem;                       // [64][12]: x' = x - shift in columns 0..7 (zero padded), raw x not needed separately
  double* tile = xs + (size_t)kTP * XS;    // [64][S] logits -> responsibilities
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int comp = 32 * warp + lane;       // the component this lane owns in P1 / P3
  const bool own = comp < KP;
  const bool wact = 32 * warp < KP;        // warp-uniform: this warp owns at least one component

  // component parameters and statistics live in registers for the whole sweep
  double l[kD], b[kD], mu[kD], sh[kD], ck = -INFINITY, thi = INFINITY, tlo = INFINITY;
#pragma unroll
  for (int a = 0; a < kD; ++a) {
    sh[a] = A.shift[a];
    l[a] = own ? A.pack2[(size_t)comp * PSF + a] : 0.0;
    b[a] = own ? A.pack2[(size_t)comp * PSF + 8 + a] : 0.0;
    mu[a] = own ? A.muold[(size_t)comp * 8 + a] - sh[a] : 0.0;   // scatter centre in shifted coordinates
  }
  if (own) { ck = A.pack2[(size_t)comp * PSF + 16]; thi = A.thi[comp]; tlo = A.tlo[comp]; }
  const bool accu = A.accumulate != 0;
  double* oA = A.partA + ((size_t)blockIdx.x * KP + (own ? comp : 0)) * DS;
  double* oB = A.partB + ((size_t)blockIdx.x * KP + (own ? comp : 0)) * MSF;
  double sg = 0, sg2 = 0, sk = 0, sgx[kD], skx[kD], skz2[kD];
#pragma unroll
  for (int a = 0; a < kD; ++a) {
    const bool ld = accu && own;
    // the partial slots hold raw sums (see the end of the kernel): back to the shifted / centred accumulators
    sgx[a] = (ld && a < d) ? oA[2 + a] - sh[a] * oA[0] : 0.0;
    skx[a] = (ld && a < d) ? oB[64 + a] - (mu[a] + sh[a]) * oB[72] : 0.0;
    skz2[a] = ld ? oB[a * 9] : 0.0;   // diagonal entry (a, a) of the 8x8 scatter block
  }
  if (accu && own) { sg = oA[0]; sg2 = oA[1]; sk = oB[72]; }
  const int nlist = blockIdx.x * 8 + warp;
  int nAmb = accu ? A.amb_count[nlist] : 0;
  AmbEntry* ambw = A.amb + (size_t)nlist * A.amb_cap;

  const int64_t ntiles = (N + kTP - 1) / kTP;
  for (int64_t bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {
    const int64_t q0 = bt * kTP;
    // ---- stage the tile: x' = x - shift, zero padded to 8 dimensions; rows beyond N are zero ---------------------
    for (int i = tid; i < kTP * 8; i += 256) {
      const int pt = i >> 3, a = i & 7;
      const int64_t p = q0 + pt;
      xs[pt * XS + a] = (p < N && a < d) ? A.X[p * d + a] - A.shift[a] : 0.0;
    }
    __syncthreads();
    // ---- P1: lane = component; logits of every point of the tile -----------------------------------------------------
    if (own) {
#pragma unroll 4
      for (int pt = 0; pt < kTP; ++pt) {
        const double2* xr = reinterpret_cast<const double2*>(xs + pt * XS);
        double s = 0;
#pragma unroll
        for (int a2 = 0; a2 < kD / 2; ++a2) {
          const double2 x = xr[a2];   // broadcast load: every lane reads the same point
          const double u0 = fma(l[2 * a2], x.x, b[2 * a2]);
          const double u1 = fma(l[2 * a2 + 1], x.y, b[2 * a2 + 1]);
          s = fma(u0, u0, s);
          s = fma(u1, u1, s);
        }
        tile[(size_t)pt * S + comp] = fma(-0.5, s, ck);
      }
    }
    __syncthreads();
    // ---- P2: warp = 8 points, lane (g, t) = point g, components == t (mod 4): log-sum-exp over the components ----------
    {
      const int64_t p = q0 + warp * 8 + g;
      double* row = tile + (size_t)(warp * 8 + g) * S;
      double m = -INFINITY;
      int am = 0;
      for (int k4 = 0; k4 < K4; ++k4) {
        const double v = row[4 * k4 + t];
        if (v > m) { m = v; am = 4 * k4 + t; }
      }
#pragma unroll
      for (int s2 = 1; s2 <= 2; s2 <<= 1) {
        const double om = shx(m, s2);
        const int oa = __shfl_xor_sync(kFull, am, s2);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
      }
      double sum = 0;
#pragma unroll 4
      for (int k4 = 0; k4 < K4; k4 += 2) {
        const double e0 = exp_bf(row[4 * k4 + t] - m), e1 = exp_bf(row[4 * k4 + 4 + t] - m);
        row[4 * k4 + t] = e0; row[4 * k4 + 4 + t] = e1;
        sum += e0; sum += e1;
      }
      sum += shx(sum, 1);
      sum += shx(sum, 2);
      const double inv = (p < N) ? 1.0 / sum : 0.0;
      for (int k4 = 0; k4 < K4; ++k4) row[4 * k4 + t] *= inv;
      if (A.labels != nullptr && p < N && t == 0) A.labels[p] = am + 1;
    }
    __syncthreads();
    // ---- P3: lane = component; statistics of the tile ---------------------------------------------------------------------
    if (wact) {   // warp-uniform branch (ballots inside); lanes without a component carry g = 0
#pragma unroll 2
      for (int pt = 0; pt < kTP; ++pt) {
        const double a = own ? tile[(size_t)pt * S + comp] : 0.0;
        const double2* xr = reinterpret_cast<const double2*>(xs + pt * XS);
        double x[kD];
#pragma unroll
        for (int a2 = 0; a2 < kD / 2; ++a2) { const double2 v = xr[a2]; x[2 * a2] = v.x; x[2 * a2 + 1] = v.y; }
        sg += a;
        sg2 = fma(a, a, sg2);
#pragma unroll
        for (int c = 0; c < kD; ++c) sgx[c] = fma(a, x[c], sgx[c]);   // about the shift; the shift is added back at the end
        // "If(W(i) < minW0) CYCLE" (src/mlf_matrix.f90:135); g == 0 adds nothing and is never listed
        const bool hi = !(a < thi) && a != 0.0;
        const bool mid = !hi && !(a < tlo) && a != 0.0;
        if (__any_sync(kFull, hi)) {
          const double ah = hi ? a : 0.0;
          sk += ah;
#pragma unroll
          for (int c = 0; c < kD; ++c) {
            const double z = x[c] - mu[c];
            const double az = ah * z;
            skx[c] += az;
            skz2[c] = fma(az, z, skz2[c]);
          }
        }
        const unsigned mm = __ballot_sync(kFull, mid);
        if (mm) {  // undecided until s_k is final: side list, settled by k_amb_resolve
          const int sl = nAmb + __popc(mm & ((1u << lane) - 1u));
          if (mid && sl < A.amb_cap) {
            AmbEntry e;
            e.g = a;
            e.key = (long long)(A.p_offset + q0 + pt) * 1024 + comp;
            ambw[sl] = e;
          }
          nAmb += __popc(mm);
        }
      }
    }
    __syncthreads();
  }
  // ---- per-block partials in the layouts of the fused kernels ---------------------------------------------------------------
  if (own) {
    oA[0] = sg;
    oA[1] = sg2;
    // sum g x = sum g x' + shift * sum g;  sum_kept g x = sum_kept g (z + mu') + shift * sum_kept g
#pragma unroll
    for (int a = 0; a < kD; ++a) {
      oA[2 + a] = (a < d) ? fma(sh[a], sg, sgx[a]) : 0.0;
      oB[64 + a] = (a < d) ? skx[a] + (mu[a] + sh[a]) * sk : 0.0;
    }
    for (int i = 0; i < 64; ++i) oB[i] = 0.0;
#pragma unroll
    for (int a = 0; a < kD; ++a) oB[a * 9] = skz2[a];
    oB[72] = sk;
    oB[73] = 0.0;
  }
  if (lane == 0) A.amb_count[nlist] = nAmb;
}

cudaError_t launch_emd(const Geom& g, const FusedGeom& f, EmArgs a, int grid, size_t smem_limit, cudaStream_t st) {
  if (!emd_ok(g, f, smem_limit)) return cudaErrorInvalidConfiguration;
  const size_t smem = emd_smem_bytes(g);
  cudaError_t e = cudaFuncSetAttribute(k_emd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  a.S = g.S;
  k_emd<<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace mlfb200
