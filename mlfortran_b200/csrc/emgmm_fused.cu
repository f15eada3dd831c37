// Fused single-pass EM kernel for sm_100a (reference: ExpStep + MaxStep, src/unsupervised/mlf_emgmm.f90:174-207;
// mlf_EvalGaussian / mlf_MaxGaussian, src/unsupervised/mlf_gaussian.f90:41-79; GECovMatrix, src/mlf_matrix.f90:99-147).
//
// One sweep over X per iteration, responsibilities never leave the SM:
//   P1  whitening  u = L_k^T (x - m) - L_k^T (mu_k - m)  as [8 points x 4 dims] x [4 dims x 8 coords] DMMA tiles whose
//       accumulators start at the per-component bias (m = global mean of the data, so nothing is subtracted per
//       component); maha = |u|^2; logits log(lambda_k) - lnDet_k - maha/2 into a shared-memory tile;
//   P2  log-sum-exp over the components, responsibilities in place (optionally stored / labelled);
//   P3  pass-A statistics sum g, sum g^2, sum g x as [8 comps x 4 points] x [4 points x 8 dims] DMMA, and -- fused --
//       the covariance scatter sum g (x - mu_k)(x - mu_k)^T of the points that pass the reference's
//       "W(i) < minW -> CYCLE" test (src/mlf_matrix.f90:135), centred at the *current* mean and re-centred exactly
//       at the new mean by k_finalize_fused.
//
// The reference's threshold needs s_k = sum_i g_ik of the *finished* E-step.  The kernel therefore classifies
// against a band [tlo_k, thi_k] that brackets minW*s_k: g >= thi is certainly kept (accumulated here), g < tlo is
// certainly dropped, and the rare pairs in between go to a per-warp side list that k_amb_resolve settles once the
// all-reduced s_k is known.  With tlo == thi (s_k known exactly) the list stays empty.  The host checks that the
// finished s_k lies inside the band it predicted and repeats the sweep with the exact threshold otherwise, so the
// result is always the reference's.
#include "emgmm_kernels.cuh"

#include <cmath>

namespace mlfb200 {

namespace {

__host__ __device__ constexpr int f_blk_off(int mb) {
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

// Branch-free fp64 exp for the softmax (arguments are logit - max <= 0, or raw logits): Cody-Waite reduction
// r = a - n ln2 with |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18 relative), and a two-factor power of
// two so that results in the denormal range round correctly.  Branch-free, so the compiler interleaves the eight
// evaluations a lane has in flight; libm-style exp() carries three data-dependent branches per call.
__device__ __forceinline__ double exp_bf(double a) {
  a = (a < -750.0) ? -750.0 : a;   // exp(-750) underflows to 0 below; NaN stays NaN, +inf handled by the caller's data
  a = (a > 710.0) ? 710.0 : a;     // 2^1024 overflows to +inf through the scaling
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  double t = fma(a, 1.4426950408889634074, magic);
  const int n = __double2loint(t);
  t -= magic;
  double r = fma(t, -6.93147180369123816490e-01, a);
  r = fma(t, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821614599e-10;            // 1/13!
  p = fma(p, r, 2.0876756987868098979e-09);        // 1/12!
  p = fma(p, r, 2.5052108385441718775e-08);        // 1/11!
  p = fma(p, r, 2.7557319223985890653e-07);        // 1/10!
  p = fma(p, r, 2.7557319223985892510e-06);        // 1/9!
  p = fma(p, r, 2.4801587301587301566e-05);        // 1/8!
  p = fma(p, r, 1.9841269841269841253e-04);        // 1/7!
  p = fma(p, r, 1.3888888888888889419e-03);        // 1/6!
  p = fma(p, r, 8.3333333333333332177e-03);        // 1/5!
  p = fma(p, r, 4.1666666666666664354e-02);        // 1/4!
  p = fma(p, r, 1.6666666666666665741e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int n1 = n >> 1, n2 = n - n1;
  const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

}  // namespace

FusedGeom make_fused_geom(const Geom& g, bool cov_diag, size_t smem_limit) {
  FusedGeom f{};
  f.XS = 8 * g.DN + 4;
  f.MSF = g.MB * 64 + 8 * g.DN + 2;
  f.CBMAX = g.K8 <= 8 ? 1 : (g.K8 <= 16 ? 2 : (g.K8 <= 32 ? 4 : 8));
  // scatter accumulators: 8 comps x CBMAX x MB blocks x 2 doubles per lane must stay in registers
  f.scatter_ok = (16 * f.CBMAX * g.MB <= 64) && g.K8 <= 64;
  auto plan = [&](int gpw, bool diag) {
    f.GPW = gpw;
    f.diag = diag;
    f.TP = 64 * gpw;
    f.OFFB = diag ? 8 * g.DN : g.NB * 32;
    f.OFFC = f.OFFB + 8 * g.DN;
    f.PSF = f.OFFC + 4;
    f.smem = ((size_t)g.K8 * 8 * f.PSF + (size_t)g.K8 * 8 * 8 * g.DN + (size_t)f.TP * f.XS + (size_t)f.TP * g.S) * sizeof(double);
    return f.smem <= smem_limit;
  };
  // preference: 128-point tiles with the full pack; 64-point tiles for large K at d <= 8; the diagonal-compressed
  // pack (extension) when the caller asked for diagonal covariances and nothing else fits
  // diagonal covariances at d <= 8 with >= 128 components: the lane-per-component CUDA-core kernel (emgmm_diag.cu) wants
  // the compressed pack
  if (cov_diag && g.DN == 1 && g.K8 * 8 >= 128 && plan(1, true)) return f;
  if (plan(2, false)) return f;
  const bool small = (g.DM <= 2) && (f.CBMAX == 2 || f.CBMAX == 4);
  if (small && plan(1, false)) return f;
  if (small && cov_diag && plan(1, true)) return f;
  plan(2, false);
  return f;
}

// GPW: 8-point groups per warp (2: 128-point tiles; 1: 64-point tiles for large K, whose logit tile would not fit).
// DIAG: diagonal-covariance extension with the compressed pack [8*DN diag(L) | 8*DN bias | const]: the B fragments are
// synthesised from the diagonal, which is what lets K = 256 at d = 8 (BASELINE config 5) fit shared memory.
template <int DM, int CBMAX, bool SCATTER, int GPW, bool DIAG>
__global__ void __launch_bounds__(kEstepThreads, 1) k_em(EmArgs A) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int NB = f_blk_off(DM);
  constexpr int OFFB = DIAG ? 8 * DN : NB * 32, OFFC = OFFB + 8 * DN, PSF = OFFC + 4;
  constexpr int TP = 64 * GPW;   // points per block tile
  constexpr int WP = 8 * GPW;    // points per warp
  constexpr int XS = 8 * DN + 4;
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int DS = 8 * DN + 2;
  constexpr int MSF = MB * 64 + 8 * DN + 2;
  constexpr int SC = SCATTER ? CBMAX : 1;
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, K4 = 2 * K8, S = A.S, d = A.d, K = A.K;
  const int64_t N = A.N;
  double* sp = smem;                                  // [KP][PSF]  L fragments | -L^T(mu-m) | log(lambda)-lnDet
  double* smu = sp + (size_t)KP * PSF;                // [KP][8 DN] current means (scatter centre)
  double* xs = smu + (size_t)KP * 8 * DN;             // [128][XS]  raw points of the tile
  double* tile = xs + (size_t)TP * XS;                // [TP][S]    logits -> responsibilities
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < KP * PSF; i += kEstepThreads) sp[i] = A.pack2[i];
  for (int i = tid; i < KP * 8 * DN; i += kEstepThreads) smu[i] = A.muold[i];
  double sh[DM];
#pragma unroll
  for (int mb = 0; mb < DM; ++mb) sh[mb] = A.shift[4 * mb + t];
  __syncthreads();

  double accA[CBMAX][DN][2], sg[CBMAX], sg2[CBMAX];
  double accK[SC][DN][2], sk[SC], thi_r[SC], tlo_r[SC];
  double accS[SC][8][MB][2];
  // accumulators start at zero, or -- chunked sweep of a streamed X -- at what this block's partial slots hold
  const bool accu = A.accumulate != 0;
  double* outA = A.partA + (size_t)blockIdx.x * KP * DS;
  double* outB = SCATTER ? A.partB + (size_t)blockIdx.x * KP * MSF : nullptr;
#pragma unroll
  for (int c = 0; c < CBMAX; ++c) {
    const int cb = warp + 8 * c;
    const bool ld = accu && cb < K8;
    const double* o = outA + (size_t)(8 * cb + g) * DS;
    sg[c] = (ld && t == 0) ? o[0] : 0.0;
    sg2[c] = (ld && t == 0) ? o[1] : 0.0;
#pragma unroll
    for (int b = 0; b < DN; ++b) {
      accA[c][b][0] = ld ? o[2 + 8 * b + 2 * t] : 0.0;
      accA[c][b][1] = ld ? o[2 + 8 * b + 2 * t + 1] : 0.0;
    }
  }
#pragma unroll
  for (int c = 0; c < SC; ++c) {
    const int cb = warp + 8 * c;
    const bool ld = SCATTER && accu && cb < K8;
    thi_r[c] = (SCATTER && cb < K8) ? A.thi[8 * cb + g] : INFINITY;
    tlo_r[c] = (SCATTER && cb < K8) ? A.tlo[8 * cb + g] : INFINITY;
    const double* ok = ld ? outB + (size_t)(8 * cb + g) * MSF + MB * 64 : nullptr;
    sk[c] = (ld && t == 0) ? ok[8 * DN] : 0.0;
#pragma unroll
    for (int b = 0; b < DN; ++b) {
      accK[c][b][0] = ld ? ok[8 * b + 2 * t] : 0.0;
      accK[c][b][1] = ld ? ok[8 * b + 2 * t + 1] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double* oq = ld ? outB + (size_t)(8 * cb + q) * MSF : nullptr;
#pragma unroll
      for (int b = 0; b < MB; ++b) {
        accS[c][q][b][0] = ld ? oq[b * 64 + g * 8 + 2 * t] : 0.0;
        accS[c][q][b][1] = ld ? oq[b * 64 + g * 8 + 2 * t + 1] : 0.0;
      }
    }
  }
  int nAmb = (SCATTER && accu) ? A.amb_count[blockIdx.x * (kEstepThreads / 32) + warp] : 0;
  AmbEntry* ambw = SCATTER ? A.amb + ((size_t)blockIdx.x * (kEstepThreads / 32) + warp) * A.amb_cap : nullptr;

  const int64_t ntiles = (N + TP - 1) / TP;
  for (int64_t bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {
    const int64_t p0 = bt * TP + warp * WP;
    // ---- P1: logits of this warp's 16 points against every component ---------------------------------------------
    double xa[GPW][DM];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) {
      const int64_t p = p0 + 8 * grp + g;
      double* xr = xs + (size_t)(warp * WP + 8 * grp + g) * XS;
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) {
        const int a = 4 * mb + t;
        const double x = (p < N && a < d) ? A.X[p * d + a] : 0.0;
        xr[a] = x;
        xa[grp][mb] = x - sh[mb];
      }
      if (8 * DN > 4 * DM) xr[4 * DM + t] = 0.0;  // zero padding up to 8*DN
    }
    double* row[GPW];
    double rmax[GPW];
    int rarg[GPW];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) { row[grp] = tile + (size_t)(warp * WP + 8 * grp + g) * S; rmax[grp] = -INFINITY; rarg[grp] = 0; }
    for (int k4 = 0; k4 < K4; ++k4) {
      double part[GPW][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* pk = sp + (size_t)(4 * k4 + j) * PSF;
        double Lf[NB];
        double2 nb2[DN];
        if (DIAG) {  // lane (k = t, n = g) of block (mb, nb) holds L[4mb+t][8nb+g]: only the diagonal is non-zero
#pragma unroll
          for (int mb = 0; mb < DM; ++mb)
#pragma unroll
            for (int nb = 0; nb <= mb / 2; ++nb) Lf[f_blk_off(mb) + nb] = (4 * mb + t == 8 * nb + g) ? pk[8 * nb + g] : 0.0;
        } else {
#pragma unroll
          for (int b = 0; b < NB; ++b) Lf[b] = pk[b * 32 + lane];
        }
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
            for (int nb = 0; nb <= mb / 2; ++nb) dmma(c[nb], xa[grp][mb], Lf[f_blk_off(mb) + nb]);
          double s = c[0][0] * c[0][0];
          s = fma(c[0][1], c[0][1], s);
#pragma unroll
          for (int nb = 1; nb < DN; ++nb) { s = fma(c[nb][0], c[nb][0], s); s = fma(c[nb][1], c[nb][1], s); }
          part[grp][j] = s;
        }
      }
      const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
#pragma unroll
      for (int grp = 0; grp < GPW; ++grp) {
        // transposed reduction over the quad: lane t ends with the full |u|^2 of component 4*k4+t
        const bool hi2 = (t & 2) != 0, hi1 = (t & 1) != 0;
        const double send0 = hi2 ? part[grp][0] : part[grp][2];
        const double send1 = hi2 ? part[grp][1] : part[grp][3];
        const double keep0 = hi2 ? part[grp][2] : part[grp][0];
        const double keep1 = hi2 ? part[grp][3] : part[grp][1];
        const double a0 = keep0 + shx(send0, 2);
        const double a1 = keep1 + shx(send1, 2);
        const double maha = (hi1 ? a1 : a0) + shx(hi1 ? a0 : a1, 1);
        const double logit = fma(-0.5, maha, ck);
        row[grp][4 * k4 + t] = logit;
        if (logit > rmax[grp]) { rmax[grp] = logit; rarg[grp] = 4 * k4 + t; }
      }
    }
    // ---- P2: log-sum-exp over the components (4 independent exp chains per lane) --------------------------------
    double mx[GPW], inv[GPW], sum[GPW];
    int am[GPW];
#pragma unroll
    for (int grp = 0; grp < GPW; ++grp) {
      double m = rmax[grp];
      int a = rarg[grp];
#pragma unroll
      for (int s2 = 1; s2 <= 2; s2 <<= 1) {
        const double om = shx(m, s2);
        const int oa = __shfl_xor_sync(kFull, a, s2);
        if (om > m || (om == m && oa < a)) { m = om; a = oa; }
      }
      mx[grp] = A.raw ? 0.0 : m;  // raw: P = lambda*exp(v - lnDet) un-normalised (src/unsupervised/mlf_gaussian.f90:77)
      am[grp] = a;
      sum[grp] = 0;
    }
    bool hard_done = false;
    if constexpr (!SCATTER) {
      if (A.hard) {
        // k-means assignment (mlf_EvaluateClass, src/unsupervised/mlf_kmeans.f90:151-165): with L = I and equal constants
        // the largest logit is the nearest centre (first minimum wins); responsibilities are the one-hot labels, so the
        // pass-A statistics below are the cluster counts and coordinate sums of EvaluateCentres (mlf_kmeans_naive.f90:74-96)
#pragma unroll
        for (int grp = 0; grp < GPW; ++grp) {
          const int64_t p = p0 + 8 * grp + g;
          const bool valid = p < N;
          for (int k4 = 0; k4 < K4; ++k4) row[grp][4 * k4 + t] = (valid && 4 * k4 + t == am[grp]) ? 1.0 : 0.0;
          if (A.labels != nullptr && valid && t == 0) A.labels[p] = am[grp] + 1;
        }
        hard_done = true;
      }
    }
    if (!hard_done) {
      if constexpr (GPW == 2) {
  #pragma unroll 2
        for (int k4 = 0; k4 < K4; k4 += 2) {   // K4 is even; 4 independent exp chains per trip, 2 trips unrolled
          const double a00 = row[0][4 * k4 + t] - mx[0], a01 = row[0][4 * k4 + 4 + t] - mx[0];
          const double a10 = row[1][4 * k4 + t] - mx[1], a11 = row[1][4 * k4 + 4 + t] - mx[1];
          const double e00 = exp_bf(a00), e01 = exp_bf(a01), e10 = exp_bf(a10), e11 = exp_bf(a11);
          row[0][4 * k4 + t] = e00; row[0][4 * k4 + 4 + t] = e01;
          row[1][4 * k4 + t] = e10; row[1][4 * k4 + 4 + t] = e11;
          sum[0] += e00; sum[0] += e01;
          sum[1] += e10; sum[1] += e11;
        }
      } else {
  #pragma unroll 2
        for (int k4 = 0; k4 < K4; k4 += 2) {   // one group per warp: 2 chains per trip, 2 trips unrolled
          const double a0 = row[0][4 * k4 + t] - mx[0], a1 = row[0][4 * k4 + 4 + t] - mx[0];
          const double e0 = exp_bf(a0), e1 = exp_bf(a1);
          row[0][4 * k4 + t] = e0; row[0][4 * k4 + 4 + t] = e1;
          sum[0] += e0; sum[0] += e1;
        }
      }
  #pragma unroll
      for (int grp = 0; grp < GPW; ++grp) {
        sum[grp] += shx(sum[grp], 1);
        sum[grp] += shx(sum[grp], 2);
        const bool v = (p0 + 8 * grp + g) < N;
        inv[grp] = v ? (A.raw ? 1.0 : 1.0 / sum[grp]) : 0.0;
      }
  #pragma unroll
      for (int grp = 0; grp < GPW; ++grp) {
        const int64_t p = p0 + 8 * grp + g;
        const bool valid = p < N;
        for (int k4 = 0; k4 < K4; ++k4) {
          const int k = 4 * k4 + t;
          const double gk = row[grp][k] * inv[grp];
          row[grp][k] = gk;
          if (A.gamma != nullptr && valid && k < K) A.gamma[(size_t)k * A.ldg + p] = gk;
        }
        if (A.labels != nullptr && valid && t == 0) A.labels[p] = am[grp] + 1;  // MAXLOC: first maximum, 1-based
      }
    }
    __syncthreads();
    // ---- P3: statistics of the block tile; warp w owns component blocks w, w+8, ... ------------------------------
    // Loop A (branch-light, pipelined): pass-A DMMA for every 4-point group pb, kept/undecided classification; lane pb
    // remembers which of the warp's 8 components keep a point of group pb.  Loop B: per component (static index, so
    // its accumulators stay in named registers) walk only the groups that component keeps.
#pragma unroll
    for (int ci = 0; ci < CBMAX; ++ci) {
      const int cb = warp + 8 * ci;
      if (cb < K8) {
        constexpr int cs = SCATTER ? 1 : 0;
        const int c2 = ci * cs;
        const int64_t q0 = bt * TP;
        const double* tcol = tile + 8 * cb + g;
        unsigned act = 0;  // lane pb: bit q set <=> component 8*cb+q keeps some point of group pb
#pragma unroll 4
        for (int pb = 0; pb < TP / 4; ++pb) {
          const double a = tcol[(size_t)(4 * pb + t) * S];
          double b[DN];
#pragma unroll
          for (int db = 0; db < DN; ++db) b[db] = xs[(size_t)(4 * pb + t) * XS + 8 * db + g];
          sg[ci] += a;
          sg2[ci] = fma(a, a, sg2[ci]);
#pragma unroll
          for (int db = 0; db < DN; ++db) dmma(accA[ci][db], a, b[db]);
          if (SCATTER) {
            // "If(W(i) < minW0) CYCLE" (src/mlf_matrix.f90:135); g == 0 adds nothing and is never listed
            const bool hi = !(a < thi_r[c2]) && a != 0.0;
            const bool mid = !hi && !(a < tlo_r[c2]) && a != 0.0;
            const double ahi = hi ? a : 0.0;
            sk[c2] += ahi;
            const unsigned mh = __ballot_sync(kFull, hi);
            const unsigned mm = __ballot_sync(kFull, mid);
            if (mh) {
#pragma unroll
              for (int db = 0; db < DN; ++db) dmma(accK[c2][db], ahi, b[db]);
              unsigned cm = mh | (mh >> 1);
              cm |= cm >> 2;
              cm &= 0x11111111u;                       // bit 4q: component q keeps a point of this group
              cm = (cm | (cm >> 3)) & 0x03030303u;     // compress the 8 nibble flags into 8 contiguous bits
              cm = (cm | (cm >> 6)) & 0x000f000fu;
              cm = (cm | (cm >> 12)) & 0xffu;
              if (lane == pb) act = cm;
            }
            if (mm) {  // undecided until s_k is final: side list, settled by k_amb_resolve
              const int slot = nAmb + __popc(mm & ((1u << lane) - 1u));
              if (mid && slot < A.amb_cap) {
                AmbEntry e;
                e.g = a;
                e.key = (long long)(A.p_offset + q0 + 4 * pb + t) * 1024 + (8 * cb + g);
                ambw[slot] = e;
              }
              nAmb += __popc(mm);
            }
          }
        }
        if (SCATTER) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            unsigned pm = __ballot_sync(kFull, (act >> q) & 1u);   // groups kept by component 8*cb+q
            if (pm == 0) continue;
            const double thq = __shfl_sync(kFull, thi_r[c2], 4 * q);
            const double* mq = smu + (size_t)(8 * cb + q) * 8 * DN;
            double mu_q[DN];
#pragma unroll
            for (int ab = 0; ab < DN; ++ab) mu_q[ab] = mq[8 * ab + g];
            const double* tq = tile + 8 * cb + q;
            while (pm) {
              const int pb = __ffs(pm) - 1;
              pm &= pm - 1;
              const double aq = tq[(size_t)(4 * pb + t) * S];
              const double w = (!(aq < thq) && aq != 0.0) ? aq : 0.0;
              double z[DN], wz[DN];
#pragma unroll
              for (int ab = 0; ab < DN; ++ab) {
                z[ab] = xs[(size_t)(4 * pb + t) * XS + 8 * ab + g] - mu_q[ab];
                wz[ab] = w * z[ab];
              }
#pragma unroll
              for (int ab = 0; ab < DN; ++ab)
#pragma unroll
                for (int bb = 0; bb <= ab; ++bb) dmma(accS[c2][q][ab * (ab + 1) / 2 + bb], wz[ab], z[bb]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- per-block partials ----------------------------------------------------------------------------------------
#pragma unroll
  for (int ci = 0; ci < CBMAX; ++ci) {
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
  if (SCATTER) {
#pragma unroll
    for (int ci = 0; ci < SC; ++ci) {
      const int cb = warp + 8 * ci;
      if (cb < K8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          double* o = outB + (size_t)(8 * cb + q) * MSF;
#pragma unroll
          for (int b = 0; b < MB; ++b) {
            o[b * 64 + g * 8 + 2 * t] = accS[ci][q][b][0];
            o[b * 64 + g * 8 + 2 * t + 1] = accS[ci][q][b][1];
          }
        }
        double* o = outB + (size_t)(8 * cb + g) * MSF + MB * 64;
#pragma unroll
        for (int db = 0; db < DN; ++db) {
          o[8 * db + 2 * t] = accK[ci][db][0];
          o[8 * db + 2 * t + 1] = accK[ci][db][1];
        }
        double s = sk[ci];
        s += shx(s, 1); s += shx(s, 2);
        if (t == 0) { o[8 * DN] = s; o[8 * DN + 1] = 0.0; }
      }
    }
    if (lane == 0) A.amb_count[blockIdx.x * (kEstepThreads / 32) + warp] = nAmb;
  }
}

template <int DM, int CBMAX, bool SCATTER, int GPW, bool DIAG>
static cudaError_t launch_em_t(const EmArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_em<DM, CBMAX, SCATTER, GPW, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_em<DM, CBMAX, SCATTER, GPW, DIAG><<<grid, kEstepThreads, smem, st>>>(a);
  return cudaGetLastError();
}

// 128-point tiles, full pack: every (DM, CBMAX) the register file allows
template <int DM>
static cudaError_t launch_em_dm(const EmArgs& a, int cbmax, bool scatter, int grid, size_t smem, cudaStream_t st) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int MB = DN * (DN + 1) / 2;
  if (scatter) {
    // instantiate the fused variant only where its accumulators fit the register file
    if constexpr (16 * 1 * MB <= 64) { if (cbmax == 1) return launch_em_t<DM, 1, true, 2, false>(a, grid, smem, st); }
    if constexpr (16 * 2 * MB <= 64) { if (cbmax == 2) return launch_em_t<DM, 2, true, 2, false>(a, grid, smem, st); }
    if constexpr (16 * 4 * MB <= 64) { if (cbmax == 4) return launch_em_t<DM, 4, true, 2, false>(a, grid, smem, st); }
    return cudaErrorInvalidConfiguration;
  }
  switch (cbmax) {
    case 1: return launch_em_t<DM, 1, false, 2, false>(a, grid, smem, st);
    case 2: return launch_em_t<DM, 2, false, 2, false>(a, grid, smem, st);
    case 4: return launch_em_t<DM, 4, false, 2, false>(a, grid, smem, st);
    default: return launch_em_t<DM, 8, false, 2, false>(a, grid, smem, st);
  }
}

// 64-point tiles (large K at d <= 8), full or diagonal-compressed pack
template <int DM, bool DIAG>
static cudaError_t launch_em_small(const EmArgs& a, int cbmax, bool scatter, int grid, size_t smem, cudaStream_t st) {
  if (scatter) {
    if (cbmax == 2) return launch_em_t<DM, 2, true, 1, DIAG>(a, grid, smem, st);
    if (cbmax == 4) return launch_em_t<DM, 4, true, 1, DIAG>(a, grid, smem, st);
  } else {
    if (cbmax == 2) return launch_em_t<DM, 2, false, 1, DIAG>(a, grid, smem, st);
    if (cbmax == 4) return launch_em_t<DM, 4, false, 1, DIAG>(a, grid, smem, st);
  }
  return cudaErrorInvalidConfiguration;
}

cudaError_t launch_em(const Geom& g, const FusedGeom& f, const EmArgs& a, bool scatter, int grid, size_t smem_limit, cudaStream_t st) {
  if (f.smem > smem_limit || g.K8 > 64) return cudaErrorInvalidConfiguration;
  if (scatter && !f.scatter_ok) return cudaErrorInvalidConfiguration;
  if (f.GPW == 1) {
    if (g.DM == 1) return f.diag ? launch_em_small<1, true>(a, f.CBMAX, scatter, grid, f.smem, st) : launch_em_small<1, false>(a, f.CBMAX, scatter, grid, f.smem, st);
    if (g.DM == 2) return f.diag ? launch_em_small<2, true>(a, f.CBMAX, scatter, grid, f.smem, st) : launch_em_small<2, false>(a, f.CBMAX, scatter, grid, f.smem, st);
    return cudaErrorInvalidConfiguration;
  }
  switch (g.DM) {
    case 1: return launch_em_dm<1>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 2: return launch_em_dm<2>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 3: return launch_em_dm<3>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 4: return launch_em_dm<4>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 5: return launch_em_dm<5>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 6: return launch_em_dm<6>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 7: return launch_em_dm<7>(a, f.CBMAX, scatter, grid, f.smem, st);
    case 8: return launch_em_dm<8>(a, f.CBMAX, scatter, grid, f.smem, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// side list of undecided pairs: settle against the exact threshold minW * s_k and accumulate their contribution
// relative to the current mean.  outC[k] = [d*d scatter (column-major) | d sum g z | sum g].
// Two kernels, both with a fixed summation order (entries in list order, lists in index order), so the result is
// deterministic:
//   k_amb_partial  one block per list: for every component that occurs in the list (<= cmax), thread e accumulates
//                  entry e of that component's [scatter | sum g z | sum g] over the list's entries;
//   k_amb_reduce   one block per component: sums the lists' partials in list order.
// The first version walked every entry of every list serially in every thread of every component's block and capped
// the protocol at 200 000 listed pairs; this one settles millions of pairs in well under a millisecond.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAmbStage = 1024;   // entries staged in shared memory at a time

__global__ void __launch_bounds__(288) k_amb_partial(int d, int DS, const double* __restrict__ X, const AmbEntry* __restrict__ amb, int cap,
                                                     const int* __restrict__ counts, const double* __restrict__ statsA, double minW,
                                                     const double* __restrict__ muold, int ldmu, int cmax, int* __restrict__ pcomp,
                                                     double* __restrict__ partial) {
  __shared__ AmbEntry s_e[kAmbStage];
  __shared__ unsigned s_mask[32];      // components 0..1023 present in the list
  __shared__ int s_comp[64];
  __shared__ int s_ncomp;
  const int l = blockIdx.x;
  const int n = counts[l] < cap ? counts[l] : cap;
  const AmbEntry* list = amb + (size_t)l * cap;
  const int ne = d * d + d + 1;
  int* pc = pcomp + (size_t)l * (cmax + 1);
  if (n <= 0) {
    if (threadIdx.x == 0) pc[0] = 0;
    return;
  }
  if (threadIdx.x < 32) s_mask[threadIdx.x] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int k = (int)(list[j].key & 1023);
    atomicOr(&s_mask[k >> 5], 1u << (k & 31));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < 32; ++w)
      for (unsigned m = s_mask[w]; m; m &= m - 1)
        if (c < cmax && c < 64) s_comp[c++] = 32 * w + __ffs(m) - 1;
    s_ncomp = c;
    pc[0] = c;
    for (int i = 0; i < c; ++i) pc[1 + i] = s_comp[i];
  }
  __syncthreads();
  const int nc = s_ncomp;
  for (int ci = 0; ci < nc; ++ci) {
    const int k = s_comp[ci];
    const double thr = minW * statsA[(size_t)k * DS];
    const double* mu = muold + (size_t)k * ldmu;
    for (int e0 = 0; e0 < ne; e0 += blockDim.x) {
      const int e = e0 + threadIdx.x;
      const int a = (e < d * d) ? e % d : (e - d * d);
      const int b = (e < d * d) ? e / d : -1;
      double acc = 0;
      for (int j0 = 0; j0 < n; j0 += kAmbStage) {
        const int m = min(kAmbStage, n - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) s_e[j] = list[j0 + j];
        __syncthreads();
        if (e < ne) {
          for (int j = 0; j < m; ++j) {
            const AmbEntry en = s_e[j];
            if ((int)(en.key & 1023) != k || en.g < thr) continue;
            const long long p = en.key >> 10;
            if (e == d * d + d) acc += en.g;
            else if (b < 0) acc = fma(en.g, X[p * d + a] - mu[a], acc);
            else acc = fma(en.g * (X[p * d + a] - mu[a]), X[p * d + b] - mu[b], acc);
          }
        }
      }
      if (e < ne) partial[((size_t)l * cmax + ci) * ne + e] = acc;
    }
  }
}

__global__ void __launch_bounds__(288) k_amb_reduce(int d, int nlists, int cmax, const int* __restrict__ pcomp, const double* __restrict__ partial,
                                                    double* __restrict__ outC) {
  extern __shared__ short s_slot[];   // [nlists]: slot of component k inside list l's partials, or -1
  const int k = blockIdx.x;
  const int ne = d * d + d + 1;
  for (int l = threadIdx.x; l < nlists; l += blockDim.x) {
    const int* pc = pcomp + (size_t)l * (cmax + 1);
    short slot = -1;
    for (int i = 0; i < pc[0]; ++i)
      if (pc[1 + i] == k) slot = (short)i;
    s_slot[l] = slot;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    double acc = 0;
    for (int l = 0; l < nlists; ++l) {
      const int sl = s_slot[l];
      if (sl >= 0) acc += partial[((size_t)l * cmax + sl) * ne + e];
    }
    outC[(size_t)k * ne + e] = acc;
  }
}

cudaError_t launch_amb_resolve(const Geom& g, const double* X, const AmbEntry* amb, int cap, const int* counts, int nlists,
                               const double* statsA, double minW, const double* muold, double* outC, cudaStream_t st, int cmax, int* pcomp,
                               double* partial) {
  k_amb_partial<<<nlists, 288, 0, st>>>(g.d, g.DS, X, amb, cap, counts, statsA, minW, muold, 8 * g.DN, cmax, pcomp, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_amb_reduce<<<g.K, 288, (size_t)nlists * sizeof(short), st>>>(g.d, nlists, cmax, pcomp, partial, outC);
  return cudaGetLastError();
}

// flags[0] = listed pairs on this rank, flags[1] = lists that overflowed their capacity, flags[2] = pairs that passed the
// certainly-kept test in the sweep (pad slot of the pass-B statistics; profiling counter) -- all-reduced as doubles
__global__ void __launch_bounds__(256) k_amb_summary(const int* __restrict__ counts, int nlists, int cap, double* __restrict__ flags,
                                                     const double* __restrict__ statsB, int KP, int MSF, int off, double vote) {
  __shared__ long long s_tot[256];
  __shared__ int s_over[256], s_max[256];
  long long tot = 0;
  int over = 0, mx = 0;
  for (int l = threadIdx.x; l < nlists; l += 256) {
    tot += counts[l];
    over += counts[l] > cap;
    mx = max(mx, counts[l]);
  }
  s_tot[threadIdx.x] = tot;
  s_over[threadIdx.x] = over;
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_tot[threadIdx.x] += s_tot[threadIdx.x + s]; s_over[threadIdx.x] += s_over[threadIdx.x + s];
      s_max[threadIdx.x] = max(s_max[threadIdx.x], s_max[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    flags[0] = (double)s_tot[0];
    flags[1] = (double)s_over[0];
    double kept = 0;
    if (statsB)
      for (int k = 0; k < KP; ++k) kept += statsB[(size_t)k * MSF + off];
    flags[2] = kept;
    flags[3] = vote;               // all-reduced: ranks that can run the scatter pass from stored responsibilities
    flags[4] = (double)s_max[0];   // longest list of THIS rank (outside the all-reduced range)
  }
}
cudaError_t launch_amb_summary(const int* counts, int nlists, int cap, double* flags, cudaStream_t st, const double* statsB, int KP, int MSF,
                               int off, double vote) {
  k_amb_summary<<<1, 256, 0, st>>>(counts, nlists, cap, flags, statsB, KP, MSF, off, vote);
  return cudaGetLastError();
}

// GECovMatrix epilogue for the fused sweep: Mu_k = sum g x / s_k; the scatter was centred at the current mean mu0,
// so with delta = Mu_k - mu0, m = sum_kept g (x - mu0), sK = sum_kept g:
//   sum_kept g (x-Mu)(x-Mu)^T = S - delta m^T - m delta^T + sK delta delta^T            (exact algebra)
// then C_k = ms^2/s_k * that (src/mlf_matrix.f90:119,137,145-146), lambda floor + normalise (mlf_emgmm.f90:184-187).
__global__ void k_finalize_fused(Geom g, int MSF, const double* __restrict__ statsA, const double* __restrict__ statsB,
                                 const double* __restrict__ statsC /*nullable*/, const double* __restrict__ muold,
                                 double minProba, double* __restrict__ Mu, double* __restrict__ Cov, double* __restrict__ lambda,
                                 int cov_diag) {
  const int d = g.d;
  __shared__ double s_dl[128], s_m[128];
  for (int k = blockIdx.x; k < g.K; k += gridDim.x) {
    const double* sa = statsA + (size_t)k * g.DS;
    const double* sb = statsB + (size_t)k * MSF;
    const double* sc = statsC ? statsC + (size_t)k * (d * d + d + 1) : nullptr;
    const double* m0 = muold + (size_t)k * 8 * g.DN;
    const double s = sa[0];
    const double sK = sb[g.MB * 64 + 8 * g.DN] + (sc ? sc[d * d + d] : 0.0);
    for (int a = threadIdx.x; a < d; a += blockDim.x) {
      const double mu = sa[2 + a] / s;
      s_dl[a] = mu - m0[a];
      s_m[a] = (sb[g.MB * 64 + a] - sb[g.MB * 64 + 8 * g.DN] * m0[a]) + (sc ? sc[d * d + a] : 0.0);
      Mu[(size_t)k * d + a] = mu;
    }
    __syncthreads();
    const double sw2 = sa[1] / (s * s);
    const double ms = 1.0 / (1.0 - sw2);
    const double scale = ms * ms / s;
    double* Ck = Cov + (size_t)k * d * d;
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
      const int a = i % d, b = i / d;
      if (a < b) continue;
      const int ab = a >> 3, bb = b >> 3;
      double S = sb[(ab * (ab + 1) / 2 + bb) * 64 + (a & 7) * 8 + (b & 7)];
      if (sc) S += sc[(size_t)b * d + a];
      const double v = (cov_diag && a != b) ? 0.0 : scale * (S - s_dl[a] * s_m[b] - s_m[a] * s_dl[b] + sK * s_dl[a] * s_dl[b]);
      Ck[(size_t)b * d + a] = v;
      Ck[(size_t)a * d + b] = v;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double mx = -INFINITY;
    for (int k = 0; k < g.K; ++k) mx = fmax(mx, statsA[(size_t)k * g.DS]);
    double sum = 0;
    for (int k = 0; k < g.K; ++k) {
      double l = statsA[(size_t)k * g.DS];
      if (minProba > 0 && !(l >= minProba * mx)) l = minProba * mx;
      lambda[k] = l;
      sum += l;
    }
    for (int k = 0; k < g.K; ++k) lambda[k] /= sum;
  }
}
cudaError_t launch_finalize_fused(const Geom& g, const FusedGeom& f, const double* statsA, const double* statsB, const double* statsC,
                                  const double* muold, double minProba, double* Mu, double* Cov, double* lambda, cudaStream_t st,
                                  int cov_diag) {
  k_finalize_fused<<<g.K < 148 ? g.K : 148, 128, 0, st>>>(g, f.MSF, statsA, statsB, statsC, muold, minProba, Mu, Cov, lambda, cov_diag);
  return cudaGetLastError();
}

}  // namespace mlfb200
