// Hand-written sm_100a kernels of the EM-GMM hot path (reference: src/unsupervised/mlf_emgmm.f90:174-207,
// src/unsupervised/mlf_gaussian.f90:41-79, src/mlf_matrix.f90:99-147,229-300).
//
// All arithmetic is IEEE fp64.  The two N*K*d^2 contractions run on the fp64 tensor pipe
// (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4): on B200 DMMA and DFMA share one pipe, DMMA peaks at
// 37.2 TFLOP/s vs 34.2 for register-resident DFMA and needs 8x fewer issue slots (profiles/fp64_peaks_r01.json).
//
//  k_refresh : per component: Jacobi eigendecomposition -> reference pseudo-inverse and (sign-quirked) lnDet
//              -> Cholesky factor L of invC packed in DMMA B-fragment order; closed-form sum_i maha_ik.
//  k_estep   : u = L^T (x - mu_k) as [8 points x 4 dims] x [4 dims x 8 coords] DMMA tiles, maha = |u|^2,
//              log-sum-exp over k, responsibilities, and the pass-A statistics (sum g, sum g^2, sum g x).
//  k_mstep   : thresholded centred scatter sum g (x-mu)(x-mu)^T as [8 x 4 points] x [4 points x 8] DMMA tiles,
//              skipping 4-point groups in which no point passes the reference's W >= minW test.
#include "emgmm_kernels.cuh"

#include <cstdlib>

#include <cmath>
#include <cstdio>

namespace mlfb200 {

// ---------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int blk_off(int mb) {  // fragment blocks before k-block mb: sum_{m<mb} (m/2+1)
  int s = 0;
  for (int m = 0; m < mb; ++m) s += m / 2 + 1;
  return s;
}

Geom make_geom(int d, int K, bool big) {
  Geom g{};
  g.d = d;
  g.K = K;
  g.DM = (d + 3) / 4;
  if (big) {  // large-shape kernels are instantiated for these k-block counts only; extra dimensions are zero padding
    const int sizes[] = {1, 2, 4, 8, 12, 16, 24, 32};
    for (int s : sizes)
      if (s >= g.DM) { g.DM = s; break; }
  }
  g.DN = (g.DM + 1) / 2;
  g.NB = blk_off(g.DM);
  g.PS = g.NB * 32 + 4 * g.DM + 4;
  g.K8 = (K + 7) / 8;
  g.K4 = 2 * g.K8;
  const int Kt = ((g.K8 * 8 + 15) / 16) * 16;
  g.S = Kt + 4;
  g.DS = 8 * g.DN + 2;
  g.MB = g.DN * (g.DN + 1) / 2;
  g.MS = g.MB * 64 + 2;
  return g;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// ---------------------------------------------------------------------------------------------------------------
// parameter refresh
// ---------------------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// One block (128 threads) per component.  Matrices are column-major M(a,b) = M[b*d + a] in global scratch.
// Follows InverseSymMatrix(C, invC, lnDet, .TRUE.)  (src/mlf_matrix.f90:256-298) with a cyclic Jacobi
// eigensolver in place of dsytrd/dorgtr/dsteqr (only invC and lnDet enter the result, not the eigenvectors).
__global__ void __launch_bounds__(1024) k_refresh(Geom g, const double* __restrict__ Mu, const double* __restrict__ Cov,
                                                 const double* __restrict__ lambda, const double* __restrict__ Sc,
                                                 const double* __restrict__ mean, double Nglob, double* __restrict__ scratch,
                                                 double* __restrict__ pack, double* __restrict__ sumLLk,
                                                 int* __restrict__ info, double* __restrict__ pack2, double* __restrict__ muold,
                                                 int PSF, int OFFB2, int OFFC2, const double* __restrict__ shift, int do_sumll, int use_smem,
                                                 int cov_diag, int diag_pack, int warm, QfRefresh qf, int fastpath, int* __restrict__ vvalid) {
  const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, d = g.d;
  // quadratic-form pack (k_em16 QF variant): component k is column k % 8 of block k / 8, B-fragment order (lane = 4 column + k-index)
  const int qNF = qf_nsteps(g.DM), qD4 = 4 * g.DM;
  double* pq = qf.packq ? qf.packq + (size_t)(k >> 3) * qf_stride(g.DM) : nullptr;
  const int OFFMU = g.NB * 32, OFFC = OFFMU + 4 * g.DM;
  double* pk = pack + (size_t)k * g.PS;
  if (k >= g.K) {  // dummy component: contributes gamma = 0
    for (int i = tid; i < g.PS; i += nt) pk[i] = 0.0;
    __syncthreads();
    if (tid == 0) pk[OFFC] = -INFINITY;
    if (pack2) {
      double* p2 = pack2 + (size_t)k * PSF;
      for (int i = tid; i < PSF; i += nt) p2[i] = (i == OFFC2) ? -INFINITY : 0.0;
      for (int i = tid; i < 8 * g.DN; i += nt) muold[(size_t)k * 8 * g.DN + i] = 0.0;
    }
    if (pq) {
      for (int f = tid; f < 4 * qNF; f += nt) pq[(f >> 2) * 32 + (k & 7) * 4 + (f & 3)] = 0.0;
      if (tid == 0) { pq[qNF * 32 + (k & 7)] = -INFINITY; qf.qfS[k] = 0.0; }
    } else if (qf.qfS && tid == 0) {
      qf.qfS[k] = 0.0;
    }
    return;
  }
  extern __shared__ __align__(16) double s_dyn[];
  // the working copy A of the Jacobi sweeps lives in shared memory for large d (global scratch round trips dominated
  // the refresh at d = 128); V and W stay in global scratch
  double* A = use_smem ? s_dyn : scratch + (size_t)k * 3 * d * d;
  double* V = scratch + (size_t)k * 3 * d * d + (size_t)d * d;
  double* W = V + (size_t)d * d;
  const double* Ck = Cov + (size_t)k * d * d;
  __shared__ double s_c[64], s_s[64], s_red[1024], s_f[128];
  __shared__ int s_p[64], s_q[64];

  // dsytrd 'U' reads the upper triangle only (src/mlf_matrix.f90:269)
  auto csym = [&](int a, int b) -> double {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return (cov_diag && a != b) ? 0.0 : Ck[(size_t)hi * d + lo];  // cov_diag: extension, diagonal of C only
  };
  // ---- fast path for well-conditioned covariances ---------------------------------------------------------------------------
  // InverseSymMatrix floors the spectrum at valM = 1e-8 max(lambda): when every eigenvalue is above that (cond(C) < 1e8) its
  // result IS the plain inverse and lnDet = (d log 2pi - log det C)/2.  Both come out of one Cholesky factorisation -- d^3
  // flops in ~4d block steps instead of Jacobi sweeps of d rounds each -- and a bound on the condition number comes with them:
  // cond_2(C) <= |C|_F |C^-1|_F.  Taken only when that bound is below 1e6 (two decades inside the floor); otherwise, or when a
  // pivot is not positive, the eigen branch below runs as before.  (Rounding: both routes carry cond(C) u.)
  __shared__ int s_fast;
  double lnDet = 0.0;
  bool fast = false;
  if (fastpath) {
    double cn = 0;
    for (int i = tid; i < d * d; i += nt) {
      const int a = i % d, b = i / d;
      const double v = csym(a, b);
      A[i] = v;
      cn = fma(v, v, cn);
    }
    cn = block_sum(cn, s_red);   // |C|_F^2 (also the barrier after the copy)
    if (tid == 0) s_fast = 1;
    __syncthreads();
    for (int j = 0; j < d; ++j) {   // in-place Cholesky A = L L^T (lower), right-looking
      const double pj = A[(size_t)j * d + j];
      if (!(pj > 0.0) || !(pj < INFINITY)) {   // uniform: every thread reads the same pivot
        if (tid == 0) s_fast = 0;
        break;
      }
      const double ljj = sqrt(pj);
      __syncthreads();
      for (int i = j + tid; i < d; i += nt) A[(size_t)j * d + i] = (i == j) ? ljj : A[(size_t)j * d + i] / ljj;
      __syncthreads();
      const int rem = d - j - 1;
      for (int i = tid; i < rem * rem; i += nt) {
        const int a = j + 1 + i % rem, b = j + 1 + i / rem;
        if (a >= b) A[(size_t)b * d + a] -= A[(size_t)j * d + a] * A[(size_t)j * d + b];
      }
      __syncthreads();
    }
    __syncthreads();
    if (s_fast) {
      // W = L^-1 (lower, column-major): thread j solves L x = e_j by forward substitution
      for (int j = tid; j < d; j += nt) {
        double* x = W + (size_t)j * d;
        for (int i = 0; i < j; ++i) x[i] = 0.0;
        x[j] = 1.0 / A[(size_t)j * d + j];
        for (int i = j + 1; i < d; ++i) {
          double acc = 0;
          for (int e = j; e < i; ++e) acc = fma(A[(size_t)e * d + i], x[e], acc);
          x[i] = -acc / A[(size_t)i * d + i];
        }
      }
      double ld = 0;
      for (int a = 0; a < d; ++a) ld += log(A[(size_t)a * d + a]);   // every thread, same order
      __syncthreads();
      // invC = L^-T L^-1 -> A (the factor is not needed any more), then its norm
      double cin = 0;
      for (int i = tid; i < d * d; i += nt) {
        const int a = i % d, b = i / d;
        if (a < b) continue;
        double acc = 0;
        const double* wa = W + (size_t)a * d;
        const double* wb = W + (size_t)b * d;
        for (int e = a; e < d; ++e) acc = fma(wa[e], wb[e], acc);
        cin += (a == b ? 1.0 : 2.0) * acc * acc;
        // the strict upper part of A still holds nothing we need: write both triangles after the barrier below
        A[(size_t)b * d + a] = acc;   // lower (a >= b)
      }
      cin = block_sum(cin, s_red);
      for (int i = tid; i < d * d; i += nt) {
        const int a = i % d, b = i / d;
        if (a < b) A[(size_t)b * d + a] = A[(size_t)a * d + b];
      }
      __syncthreads();
      fast = cn * cin < 1e12;   // (|C|_F |C^-1|_F)^2 < (1e6)^2; false for NaN
      if (fast) {
        for (int i = tid; i < d * d; i += nt) W[i] = A[i];
        lnDet = 0.5 * d * log(2.0 * acos(-1.0)) - ld;
        if (tid == 0 && info) info[k] = 0;
      }
      __syncthreads();
    }
  }
  if (!fast) {
  if (warm && vvalid && !vvalid[k]) warm = 0;   // the eigenvectors in scratch are only there if the eigen branch ran before
  if (warm) {
    // warm start: the eigenvectors V of the previous EM iteration are still in scratch and nearly diagonalise the
    // new covariance, so the sweeps start from A = V^T C V (two d^3 products) and converge in 2-3 sweeps instead of ~9
    for (int i = tid; i < d * d; i += nt) {   // W = C V
      const int a = i % d, b = i / d;
      double acc = 0;
      for (int e = 0; e < d; ++e) acc = fma(csym(a, e), V[(size_t)b * d + e], acc);
      W[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < d * d; i += nt) {   // A = V^T W, symmetric by construction: computed for a <= b and mirrored
      const int a = i % d, b = i / d;
      if (a > b) continue;
      double acc = 0;
      for (int e = 0; e < d; ++e) acc = fma(V[(size_t)a * d + e], W[(size_t)b * d + e], acc);
      A[(size_t)b * d + a] = acc;
      A[(size_t)a * d + b] = acc;
    }
  } else {
    for (int i = tid; i < d * d; i += nt) {
      const int a = i % d, b = i / d;
      A[i] = csym(a, b);
      V[i] = (a == b) ? 1.0 : 0.0;
    }
  }
  __syncthreads();

  const int ne = d + (d & 1);
  int sweeps = 0;
  for (; sweeps < 60 && d > 1; ++sweeps) {
    double off = 0, dg = 0;
    for (int i = tid; i < d * d; i += nt) {
      const int a = i % d, b = i / d;
      const double v = A[i];
      if (a < b) off += v * v;
      else if (a == b) dg += v * v;
    }
    off = block_sum(off, s_red);
    dg = block_sum(dg, s_red);
    if (!(off > 1e-31 * (dg + off))) break;  // also exits on NaN
    for (int rnd = 0; rnd < ne - 1; ++rnd) {
      for (int j = tid; j < ne / 2; j += nt) {  // round-robin tournament pairing
        int a, b;
        if (j == 0) { a = rnd; b = ne - 1; }
        else { a = (rnd + j) % (ne - 1); b = (rnd - j + ne - 1) % (ne - 1); }
        const int p = a < b ? a : b, q = a < b ? b : a;
        double c = 1.0, s = 0.0;
        if (q < d) {
          const double apq = A[(size_t)q * d + p];
          if (apq != 0.0) {
            const double theta = (A[(size_t)q * d + q] - A[(size_t)p * d + p]) / (2.0 * apq);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(tt * tt + 1.0);
            s = tt * c;
          }
        }
        s_p[j] = p; s_q[j] = q; s_c[j] = c; s_s[j] = s;
      }
      __syncthreads();
      for (int i = tid; i < (ne / 2) * d; i += nt) {  // A <- A J, V <- V J   (columns p,q)
        const int j = i / d, r = i % d, p = s_p[j], q = s_q[j];
        if (q >= d) continue;
        const double c = s_c[j], s = s_s[j];
        const double arp = A[(size_t)p * d + r], arq = A[(size_t)q * d + r];
        A[(size_t)p * d + r] = c * arp - s * arq;
        A[(size_t)q * d + r] = s * arp + c * arq;
        const double vrp = V[(size_t)p * d + r], vrq = V[(size_t)q * d + r];
        V[(size_t)p * d + r] = c * vrp - s * vrq;
        V[(size_t)q * d + r] = s * vrp + c * vrq;
      }
      __syncthreads();
      for (int i = tid; i < (ne / 2) * d; i += nt) {  // A <- J^T A   (rows p,q)
        const int j = i / d, r = i % d, p = s_p[j], q = s_q[j];
        if (q >= d) continue;
        const double c = s_c[j], s = s_s[j];
        const double apr = A[(size_t)r * d + p], aqr = A[(size_t)r * d + q];
        A[(size_t)r * d + p] = c * apr - s * aqr;
        A[(size_t)r * d + q] = s * apr + c * aqr;
      }
      __syncthreads();
    }
  }
  if (tid == 0 && info) info[k] = (sweeps >= 60) ? 1 : 0;

  // eigenvalues -> pseudo-inverse spectrum with penalty, lnDet  (src/mlf_matrix.f90:283-290)
  double mx = -INFINITY;
  for (int a = tid; a < d; a += nt) mx = fmax(mx, A[(size_t)a * d + a]);
  s_red[tid] = mx;
  __syncthreads();
  for (int s = nt / 2; s > 0; s >>= 1) {
    if (tid < s) s_red[tid] = fmax(s_red[tid], s_red[tid + s]);
    __syncthreads();
  }
  mx = s_red[0];
  __syncthreads();
  const double valM = (double)10e-9f * mx;
  const double penality = 1.0 / valM;
  for (int a = tid; a < d; a += nt) {
    const double ld = A[(size_t)a * d + a];
    s_f[a] = (ld <= valM) ? penality : 1.0 / ld;
  }
  __syncthreads();
  {
    double s = 0;
    for (int a = 0; a < d; ++a) s += log(s_f[a]);  // every thread, same order: deterministic
    lnDet = 0.5 * (d * log(2.0 * acos(-1.0)) + s);
  }
  // invC = (V diag f) V^T  -> W
  for (int i = tid; i < d * d; i += nt) {
    const int a = i % d, b = i / d;
    double acc = 0;
    for (int e = 0; e < d; ++e) acc += V[(size_t)e * d + a] * s_f[e] * V[(size_t)e * d + b];
    W[i] = acc;
  }
  if (tid == 0 && vvalid) vvalid[k] = 1;
  __syncthreads();
  }   // !fast
  if (pq) {
    // coefficients of the expanded quadratic form about the shift, straight from invC (both triangles, as the reference's
    // z^T (invC z) reads them): linear invC mu', quadratic -1/2 invC_ii and -(invC_ij + invC_ji)/2, constant c_k - 1/2 mu'^T invC mu'
    __shared__ double s_mp[16], s_b[16], s_r[16];
    if (tid < 16) {
      s_mp[tid] = (tid < d) ? Mu[(size_t)k * d + tid] - shift[tid] : 0.0;
      double r = INFINITY;   // radius of the data about the shift, per coordinate; unknown => the guard rejects
      if (qf.rad) {
        r = 0.0;
        for (int w = 0; w < qf.nrad; ++w) r = fmax(r, qf.rad[(size_t)w * 16 + tid]);
      }
      s_r[tid] = (tid < d) ? r : 0.0;
    }
    __syncthreads();
    if (tid < 16) {
      double acc = 0;
      if (tid < d)
        for (int e = 0; e < d; ++e) acc = fma(0.5 * (W[(size_t)e * d + tid] + W[(size_t)tid * d + e]), s_mp[e], acc);
      s_b[tid] = acc;
    }
    __syncthreads();
    double sabs = 0;
    for (int f = tid; f < 4 * qNF; f += nt) {
      int i, j;
      qf_feature(qD4, f, i, j);
      double th = 0.0, mag = 0.0;
      if (i < d && j == qD4) { th = s_b[i]; mag = fabs(th) * s_r[i]; }
      else if (i < d && j < d) {
        th = (i == j) ? -0.5 * W[(size_t)i * d + i] : -0.5 * (W[(size_t)j * d + i] + W[(size_t)i * d + j]);
        mag = fabs(th) * s_r[i] * s_r[j];
      }
      pq[(f >> 2) * 32 + (k & 7) * 4 + (f & 3)] = th;
      sabs += mag;
    }
    sabs = block_sum(sabs, s_red);
    if (tid == 0) {
      double c2 = 0;
      for (int a = 0; a < d; ++a) c2 = fma(s_b[a], s_mp[a], c2);
      const double ckq = log(lambda[k]) - lnDet;
      pq[qNF * 32 + (k & 7)] = ckq - 0.5 * c2;
      qf.qfS[k] = sabs + fabs(0.5 * c2) + fabs(ckq);
    }
    __syncthreads();
  }
  // Cholesky invC = L L^T (lower), right-looking, in the shared-memory working matrix (3 d barriers with global-memory round
  // trips in between were most of the refresh at small d); L goes back to W for the packing below and for k_sumll
  for (int i = tid; i < d * d; i += nt) A[i] = W[i];
  __syncthreads();
  for (int j = 0; j < d; ++j) {
    if (tid == 0) A[(size_t)j * d + j] = sqrt(A[(size_t)j * d + j]);
    __syncthreads();
    const double ljj = A[(size_t)j * d + j];
    for (int i = j + 1 + tid; i < d; i += nt) A[(size_t)j * d + i] /= ljj;
    __syncthreads();
    const int rem = d - j - 1;
    for (int i = tid; i < rem * rem; i += nt) {
      const int a = j + 1 + i % rem, b = j + 1 + i / rem;
      if (a >= b) A[(size_t)b * d + a] -= A[(size_t)j * d + a] * A[(size_t)j * d + b];
    }
    __syncthreads();
  }
  for (int i = tid; i < d * d; i += nt) W[i] = A[i];
  __syncthreads();
  // pack: B fragments of u = L^T z, block (mb, nb): lane l holds L[4mb + l%4][8nb + l/4]
  for (int i = tid; i < g.NB * 32; i += nt) {
    const int blk = i >> 5, l = i & 31;
    int mb = 0;
    while (blk_off(mb + 1) <= blk) ++mb;
    const int nb = blk - blk_off(mb);
    const int m = 4 * mb + (l & 3), n = 8 * nb + (l >> 2);
    pk[i] = (m < d && n < d && m >= n) ? W[(size_t)n * d + m] : 0.0;
  }
  const double* muk = Mu + (size_t)k * d;
  for (int a = tid; a < 4 * g.DM + 4; a += nt) pk[OFFMU + a] = (a < d) ? muk[a] : 0.0;
  __syncthreads();
  const double ck = log(lambda[k]) - lnDet;
  if (tid == 0) pk[OFFC] = ck;
  if (pack2) {  // fused-kernel pack: same fragments, accumulator bias -L^T(mu - m) with m the global mean, constant
    double* p2 = pack2 + (size_t)k * PSF;
    if (diag_pack) {  // diagonal-compressed pack: [8*DN diag(L) | bias | const]
      for (int n = tid; n < 8 * g.DN; n += nt) p2[n] = (n < d) ? W[(size_t)n * d + n] : 0.0;
    } else {
      for (int i = tid; i < g.NB * 32; i += nt) p2[i] = pk[i];
    }
    for (int n = tid; n < 8 * g.DN; n += nt) {
      double acc = 0;
      if (n < d)
        for (int m = n; m < d; ++m) acc = fma(W[(size_t)n * d + m], muk[m] - shift[m], acc);
      p2[OFFB2 + n] = -acc;
      muold[(size_t)k * 8 * g.DN + n] = (n < d) ? muk[n] : 0.0;
    }
    if (tid == 0) { p2[OFFC2] = ck; p2[OFFC2 + 1] = 0; p2[OFFC2 + 2] = 0; p2[OFFC2 + 3] = 0; }
    if (diag_pack && qf.qfS && !pq) {
      // magnitude bound of the expanded diagonal logit (k_emd QP1): sum_a (l_a^2 R_a^2 / 2 + |l_a b_a| R_a + b_a^2 / 2) + |c|,
      // R_a = radius of the data about the shift in coordinate a (unknown: +inf, the guard rejects)
      double part = 0;
      for (int n = tid; n < d; n += nt) {
        double r = INFINITY;
        if (qf.rad) {
          r = 0.0;
          for (int w = 0; w < qf.nrad; ++w) r = fmax(r, qf.rad[(size_t)w * 16 + n]);
        }
        const double l = W[(size_t)n * d + n], bb = l * (muk[n] - shift[n]);
        part += 0.5 * l * l * r * r + fabs(l * bb) * r + 0.5 * bb * bb;
      }
      part = block_sum(part, s_red);
      if (tid == 0) qf.qfS[k] = part + fabs(ck);
    }
  }

  if (!do_sumll) return;   // streamed refresh of X: the moments are not final yet, k_sumll runs after the sweep
  // closed form of sum_i (x_i-mu)^T invC (x_i-mu) = tr(L^T Sc L) + N |L^T (m - mu)|^2   (Sc = centred scatter)
  double part = 0;
  for (int n = tid; n < d; n += nt) {
    double y = 0, q = 0;
    for (int a = n; a < d; ++a) {
      const double lan = W[(size_t)n * d + a];
      y += lan * (mean[a] - muk[a]);
      double tsum = 0;
      for (int b = n; b < d; ++b) tsum += Sc[(size_t)b * d + a] * W[(size_t)n * d + b];
      q += lan * tsum;
    }
    part += q + Nglob * y * y;
  }
  part = block_sum(part, s_red);
  // sumLL_k = N (log lambda_k - lnDet_k) + sum_i (-1/2 maha_ik)   (src/unsupervised/mlf_gaussian.f90:73-76)
  if (tid == 0) sumLLk[k] = Nglob * ck - 0.5 * part;
}

cudaError_t launch_refresh(const Geom& g, const double* Mu, const double* Cov, const double* lambda, const double* Sc,
                           const double* mean, double Nglob, double* scratch, double* pack, double* sumLLk, int* info,
                           cudaStream_t st, double* pack2, double* muold, const double* shift, int do_sumll, int cov_diag,
                           const FusedGeom* fg, int warm, const QfRefresh* qf, int* vvalid) {
  const FusedGeom f = fg ? *fg : make_fused_geom(g);
  QfRefresh q{};
  if (qf && g.DM <= 4) q = *qf;
  const bool large = g.d > 32;
  // the working matrix (Jacobi's A, the Cholesky factor of the fast path) lives in shared memory for every d
  const size_t dyn = (size_t)g.d * g.d * sizeof(double);
  if (dyn > 32 * 1024) {   // static shared memory (reduction scratch, rotation parameters) comes on top of the 48 KB default limit
    cudaError_t e = cudaFuncSetAttribute(k_refresh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  const char* fe = getenv("MLF_B200_REFRESH_FAST");   // A/B switch, read per launch
  const bool fastpath = !(fe && fe[0] == '0');
  // d > 64: the eigenvector updates go to global scratch (A and V do not both fit shared memory at d = 128) and are bound by
  // loads in flight, so the block is as large as the register file allows
  k_refresh<<<g.K8 * 8, g.d > 64 ? 1024 : (large ? 256 : 128), dyn, st>>>(g, Mu, Cov, lambda, Sc, mean, Nglob, scratch, pack, sumLLk, info, pack2, muold,
                                                     f.PSF, f.OFFB, f.OFFC, shift ? shift : mean, do_sumll, 1, cov_diag, f.diag ? 1 : 0,
                                                     warm ? 1 : 0, q, (fastpath && vvalid) ? 1 : 0, vvalid);
  return cudaGetLastError();
}

// sumLL_k from the global moments, for the case where they become final only after the sweep (streamed X):
// same closed form as the tail of k_refresh, reading L (scratch W) and log(lambda_k) - lnDet_k (pack) it left behind.
__global__ void __launch_bounds__(128) k_sumll(Geom g, const double* __restrict__ Mu, const double* __restrict__ Sc,
                                               const double* __restrict__ mean, double Nglob, const double* __restrict__ scratch,
                                               const double* __restrict__ pack, double* __restrict__ sumLLk) {
  const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, d = g.d;
  __shared__ double s_red[128];
  const double* W = scratch + (size_t)k * 3 * d * d + 2 * (size_t)d * d;
  const double* muk = Mu + (size_t)k * d;
  const double ck = pack[(size_t)k * g.PS + g.NB * 32 + 4 * g.DM];
  double part = 0;
  for (int n = tid; n < d; n += nt) {
    double y = 0, q = 0;
    for (int a = n; a < d; ++a) {
      const double lan = W[(size_t)n * d + a];
      y += lan * (mean[a] - muk[a]);
      double tsum = 0;
      for (int b = n; b < d; ++b) tsum += Sc[(size_t)b * d + a] * W[(size_t)n * d + b];
      q += lan * tsum;
    }
    part += q + Nglob * y * y;
  }
  part = block_sum(part, s_red);
  if (tid == 0) sumLLk[k] = Nglob * ck - 0.5 * part;
}
cudaError_t launch_sumll(const Geom& g, const double* Mu, const double* Sc, const double* mean, double Nglob, const double* scratch,
                         const double* pack, double* sumLLk, cudaStream_t st) {
  k_sumll<<<g.K, 128, 0, st>>>(g, Mu, Sc, mean, Nglob, scratch, pack, sumLLk);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// per-coordinate radius of the data about the shift: max_i |x_ia - shift_a| (d <= 16).  Bounds the magnitude of the terms
// of the expanded quadratic form (guard of the k_em16 QF variant).  One HBM pass, once per resident data set.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_absmax(const double* __restrict__ X, int64_t N, int d, const double* __restrict__ shift,
                                                double* __restrict__ partial) {
  __shared__ double red[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ppw = 32 / d;                        // points per warp and load
  const int pl = lane / d, a = lane - pl * d;    // lanes >= ppw*d idle
  const bool on = pl < ppw;
  const double sh = on ? shift[a] : 0.0;
  double m = 0.0;
  const int64_t stride = (int64_t)gridDim.x * 8 * ppw;
  for (int64_t p = ((int64_t)blockIdx.x * 8 + warp) * ppw + pl; p < N; p += stride)
    if (on) m = fmax(m, fabs(X[p * d + a] - sh));   // fmax drops NaN: a NaN coordinate poisons the sweep either way
  if (lane < 16) red[warp][lane] = 0.0;
  __syncwarp();
  // lanes of the same coordinate: combine through shared memory (ppw <= 32 values per coordinate and warp)
  for (int q = 0; q < ppw; ++q) {
    if (on && pl == q) red[warp][a] = fmax(red[warp][a], m);
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v = fmax(v, red[w][threadIdx.x]);
    partial[(size_t)blockIdx.x * 16 + threadIdx.x] = v;
  }
}
__global__ void k_absmax_finish(const double* __restrict__ partial, int grid, double* __restrict__ out) {
  const int a = threadIdx.x;
  if (a >= 16) return;
  double v = 0.0;
  for (int b = 0; b < grid; ++b) v = fmax(v, partial[(size_t)b * 16 + a]);
  out[a] = v;
}
cudaError_t launch_absmax(const double* X, int64_t N, int d, const double* shift, double* partial, int grid, double* out, cudaStream_t st) {
  if (d > 16) return cudaErrorInvalidConfiguration;
  k_absmax<<<grid, 256, 0, st>>>(X, N, d, shift, partial);
  k_absmax_finish<<<1, 32, 0, st>>>(partial, grid, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// global moments of the data about a fixed shift vector: sum (x - sh), sum (x - sh)(x - sh)^T (lower 8x8 blocks) as
// [8 dims x 4 points] x [4 points x 8 dims] DMMA tiles.  partial[block][PM], PM = MB*64 + 8*DN; `accumulate` adds to
// what the slot already holds (chunked upload of X).  Used for the closed form of sum_i maha_ik (sumLL).
// ---------------------------------------------------------------------------------------------------------------
template <int DN>
__global__ void __launch_bounds__(256) k_moments(const double* __restrict__ X, int64_t N, int d, const double* __restrict__ shift,
                                                 double* __restrict__ partial, int accumulate) {
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int PM = MB * 64 + 8 * DN;
  __shared__ double red[8][PM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  double sh[DN], zs[DN], acc[MB][2];
#pragma unroll
  for (int ab = 0; ab < DN; ++ab) { sh[ab] = (8 * ab + g < d) ? shift[8 * ab + g] : 0.0; zs[ab] = 0; }
#pragma unroll
  for (int b = 0; b < MB; ++b) { acc[b][0] = 0; acc[b][1] = 0; }
  const int64_t ngroups = (N + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * 8;
#pragma unroll 4
  for (int64_t gi = (int64_t)blockIdx.x * 8 + warp; gi < ngroups; gi += stride) {
    const int64_t p = 4 * gi + t;
    double z[DN];
#pragma unroll
    for (int ab = 0; ab < DN; ++ab) {
      const int dim = 8 * ab + g;
      z[ab] = (p < N && dim < d) ? X[p * d + dim] - sh[ab] : 0.0;
      zs[ab] += z[ab];
    }
#pragma unroll
    for (int ab = 0; ab < DN; ++ab)
#pragma unroll
      for (int bb = 0; bb <= ab; ++bb) dmma884(acc[ab * (ab + 1) / 2 + bb], z[ab], z[bb]);
  }
#pragma unroll
  for (int b = 0; b < MB; ++b) {
    red[warp][b * 64 + g * 8 + 2 * t] = acc[b][0];
    red[warp][b * 64 + g * 8 + 2 * t + 1] = acc[b][1];
  }
#pragma unroll
  for (int ab = 0; ab < DN; ++ab) {
    double v = zs[ab];
    v += shfl_xor_d(v, 1);
    v += shfl_xor_d(v, 2);
    if (t == 0) red[warp][MB * 64 + 8 * ab + g] = v;
  }
  __syncthreads();
  double* o = partial + (size_t)blockIdx.x * PM;
  for (int i = threadIdx.x; i < PM; i += 256) {
    double v = accumulate ? o[i] : 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][i];
    o[i] = v;
  }
}
cudaError_t launch_moments(const Geom& g, const double* X, int64_t N, const double* shift, double* partial, int grid, int accumulate,
                           cudaStream_t st) {
  switch (g.DN) {
    case 1: k_moments<1><<<grid, 256, 0, st>>>(X, N, g.d, shift, partial, accumulate); break;
    case 2: k_moments<2><<<grid, 256, 0, st>>>(X, N, g.d, shift, partial, accumulate); break;
    case 3: k_moments<3><<<grid, 256, 0, st>>>(X, N, g.d, shift, partial, accumulate); break;
    case 4: k_moments<4><<<grid, 256, 0, st>>>(X, N, g.d, shift, partial, accumulate); break;
    default: return cudaErrorInvalidConfiguration;
  }
  return cudaGetLastError();
}

// raw = [MB*64 scatter blocks about shift | 8*DN sum (x - shift)] (all-reduced); writes the exact mean (zero padded,
// 8*DN + 8 doubles) and the centred scatter Sc (d x d, column-major): Sc = S - N delta delta^T, delta = mean - shift.
__global__ void k_moments_finish(Geom g, const double* __restrict__ raw, const double* __restrict__ shift, double Nglob,
                                 double* __restrict__ mean, double* __restrict__ Sc) {
  const int d = g.d;
  __shared__ double dl[128];
  for (int a = threadIdx.x; a < 8 * g.DN + 8; a += blockDim.x) {
    const double del = (a < d) ? raw[g.MB * 64 + a] / Nglob : 0.0;
    if (a < 128) dl[a] = del;
    mean[a] = (a < d) ? shift[a] + del : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
    const int a = i % d, b = i / d;
    if (a < b) continue;
    const int ab = a >> 3, bb = b >> 3;
    const double v = raw[(ab * (ab + 1) / 2 + bb) * 64 + (a & 7) * 8 + (b & 7)] - Nglob * dl[a] * dl[b];
    Sc[(size_t)b * d + a] = v;
    Sc[(size_t)a * d + b] = v;
  }
}
cudaError_t launch_moments_finish(const Geom& g, const double* raw, const double* shift, double Nglob, double* mean, double* Sc,
                                  cudaStream_t st) {
  k_moments_finish<<<1, 256, 0, st>>>(g, raw, shift, Nglob, mean, Sc);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// E-pass
// ---------------------------------------------------------------------------------------------------------------
template <int DM, int CBMAX>
__global__ void __launch_bounds__(kEstepThreads, 1)
k_estep(const double* __restrict__ X, int64_t N, int d, int K, int K8, int S, const double* __restrict__ pack,
        double* __restrict__ gamma, int64_t ldg, int32_t* __restrict__ labels, double* __restrict__ partial, int raw) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int NB = blk_off(DM);
  constexpr int PS = NB * 32 + 4 * DM + 4;
  constexpr int OFFMU = NB * 32, OFFC = OFFMU + 4 * DM;
  constexpr int DS = 8 * DN + 2;
  extern __shared__ __align__(16) double smem[];
  const int KP = K8 * 8, K4 = 2 * K8;
  double* sp = smem;                       // [KP][PS] parameter pack
  double* tile = smem + (size_t)KP * PS;   // [128][S] logits -> exp -> responsibilities
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < KP * PS; i += kEstepThreads) sp[i] = pack[i];
  __syncthreads();

  double acc[CBMAX][DN][2];
  double sg[CBMAX], sg2[CBMAX];
#pragma unroll
  for (int c = 0; c < CBMAX; ++c) {
    sg[c] = 0; sg2[c] = 0;
#pragma unroll
    for (int b = 0; b < DN; ++b) { acc[c][b][0] = 0; acc[c][b][1] = 0; }
  }

  const int64_t ntiles = (N + kEstepTilePts - 1) / kEstepTilePts;
  for (int64_t bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {
    const int64_t p0 = bt * kEstepTilePts + warp * 16;
    // ---- phase 1: logits of 16 points x all components -------------------------------------------------------
    double xf[2][DM];
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      const int64_t p = p0 + 8 * grp + g;
#pragma unroll
      for (int mb = 0; mb < DM; ++mb) {
        const int a = 4 * mb + t;
        xf[grp][mb] = (p < N && a < d) ? X[p * d + a] : 0.0;
      }
    }
    double* row[2] = {tile + (size_t)(warp * 16 + g) * S, tile + (size_t)(warp * 16 + 8 + g) * S};
    double rmax[2] = {-INFINITY, -INFINITY};
    int rarg[2] = {0, 0};
    for (int k4 = 0; k4 < K4; ++k4) {
      double part[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* pk = sp + (size_t)(4 * k4 + j) * PS;
        double Lf[NB], mu[DM];
#pragma unroll
        for (int b = 0; b < NB; ++b) Lf[b] = pk[b * 32 + lane];
#pragma unroll
        for (int mb = 0; mb < DM; ++mb) mu[mb] = pk[OFFMU + 4 * mb + t];
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
          double c[DN][2];
#pragma unroll
          for (int nb = 0; nb < DN; ++nb) { c[nb][0] = 0; c[nb][1] = 0; }
#pragma unroll
          for (int mb = 0; mb < DM; ++mb) {
            const double z = xf[grp][mb] - mu[mb];
#pragma unroll
            for (int nb = 0; nb <= mb / 2; ++nb) dmma884(c[nb], z, Lf[blk_off(mb) + nb]);
          }
          double s = c[0][0] * c[0][0];
          s = fma(c[0][1], c[0][1], s);
#pragma unroll
          for (int nb = 1; nb < DN; ++nb) { s = fma(c[nb][0], c[nb][0], s); s = fma(c[nb][1], c[nb][1], s); }
          part[grp][j] = s;
        }
      }
      const double ck = sp[(size_t)(4 * k4 + t) * PS + OFFC];
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        // transposed reduction over the 4 lanes of a row: lane t ends with the full |u|^2 of component 4*k4+t
        const bool hi2 = (t & 2) != 0, hi1 = (t & 1) != 0;
        const double send0 = hi2 ? part[grp][0] : part[grp][2];
        const double send1 = hi2 ? part[grp][1] : part[grp][3];
        const double keep0 = hi2 ? part[grp][2] : part[grp][0];
        const double keep1 = hi2 ? part[grp][3] : part[grp][1];
        const double a0 = keep0 + shfl_xor_d(send0, 2);
        const double a1 = keep1 + shfl_xor_d(send1, 2);
        const double maha = (hi1 ? a1 : a0) + shfl_xor_d(hi1 ? a0 : a1, 1);
        const double logit = fma(-0.5, maha, ck);  // log lambda_k - lnDet_k - maha/2
        row[grp][4 * k4 + t] = logit;
        if (logit > rmax[grp]) { rmax[grp] = logit; rarg[grp] = 4 * k4 + t; }
      }
    }
    // ---- phase 2: log-sum-exp over k, responsibilities ---------------------------------------------------------
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      const int64_t p = p0 + 8 * grp + g;
      const bool valid = p < N;
      double m = rmax[grp];
      int am = rarg[grp];
#pragma unroll
      for (int sh = 1; sh <= 2; sh <<= 1) {
        const double om = shfl_xor_d(m, sh);
        const int oa = __shfl_xor_sync(0xffffffffu, am, sh);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
      }
      if (raw) m = 0.0;  // mlf_evalGaussian: P = lambda*exp(v - lnDet), not normalised (src/unsupervised/mlf_gaussian.f90:77)
      double sum = 0;
      for (int k4 = 0; k4 < K4; ++k4) {
        const double e = exp(row[grp][4 * k4 + t] - m);
        row[grp][4 * k4 + t] = e;
        sum += e;
      }
      sum += shfl_xor_d(sum, 1);
      sum += shfl_xor_d(sum, 2);
      const double inv = valid ? (raw ? 1.0 : 1.0 / sum) : 0.0;
      for (int k4 = 0; k4 < K4; ++k4) {
        const int k = 4 * k4 + t;
        const double gk = row[grp][k] * inv;
        row[grp][k] = gk;
        if (gamma != nullptr && valid && k < K) gamma[(size_t)k * ldg + p] = gk;   // labels-only calls (getClass) pass no buffer
      }
      if (labels != nullptr && valid && t == 0) labels[p] = am + 1;  // MAXLOC: first maximum, 1-based
    }
    __syncthreads();
    // ---- phase 3: pass-A statistics of this block tile; warp w owns component blocks w, w+8, ... ---------------
#pragma unroll
    for (int ci = 0; ci < CBMAX; ++ci) {
      const int cb = warp + 8 * ci;
      if (cb < K8) {
        const int64_t q0 = bt * kEstepTilePts;
#pragma unroll 4
        for (int pb = 0; pb < kEstepTilePts / 4; ++pb) {
          const double a = tile[(size_t)(4 * pb + t) * S + 8 * cb + g];
          sg[ci] += a;
          sg2[ci] = fma(a, a, sg2[ci]);
          const int64_t p = q0 + 4 * pb + t;
#pragma unroll
          for (int db = 0; db < DN; ++db) {
            const int dim = 8 * db + g;
            const double b = (p < N && dim < d) ? X[p * d + dim] : 0.0;
            dmma884(acc[ci][db], a, b);
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- per-block partial statistics [KP][DS]: [sum g, sum g^2, sum g x] ------------------------------------------
  double* out = partial + (size_t)blockIdx.x * KP * DS;
#pragma unroll
  for (int ci = 0; ci < CBMAX; ++ci) {
    const int cb = warp + 8 * ci;
    if (cb < K8) {
      double a = sg[ci], b = sg2[ci];
      a += shfl_xor_d(a, 1); a += shfl_xor_d(a, 2);
      b += shfl_xor_d(b, 1); b += shfl_xor_d(b, 2);
      double* o = out + (size_t)(8 * cb + g) * DS;
      if (t == 0) { o[0] = a; o[1] = b; }
#pragma unroll
      for (int db = 0; db < DN; ++db) {
        o[2 + 8 * db + 2 * t] = acc[ci][db][0];
        o[2 + 8 * db + 2 * t + 1] = acc[ci][db][1];
      }
    }
  }
}

size_t estep_smem_bytes(const Geom& g) {
  return ((size_t)g.K8 * 8 * g.PS + (size_t)kEstepTilePts * g.S) * sizeof(double);
}
int estep_cbmax(const Geom& g) { return g.K8 <= 8 ? 1 : (g.K8 <= 16 ? 2 : (g.K8 <= 32 ? 4 : 8)); }

template <int DM, int CBMAX>
static cudaError_t launch_estep_t(const Geom& g, const double* X, int64_t N, const double* pack, double* gamma, int64_t ldg,
                                  int32_t* labels, double* partial, int grid, size_t smem, int raw, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_estep<DM, CBMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_estep<DM, CBMAX><<<grid, kEstepThreads, smem, st>>>(X, N, g.d, g.K, g.K8, g.S, pack, gamma, ldg, labels, partial, raw);
  return cudaGetLastError();
}

template <int DM>
static cudaError_t launch_estep_dm(const Geom& g, const double* X, int64_t N, const double* pack, double* gamma, int64_t ldg,
                                   int32_t* labels, double* partial, int grid, size_t smem, int raw, cudaStream_t st) {
  switch (estep_cbmax(g)) {
    case 1: return launch_estep_t<DM, 1>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 2: return launch_estep_t<DM, 2>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 4: return launch_estep_t<DM, 4>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    default: return launch_estep_t<DM, 8>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
  }
}

cudaError_t launch_estep(const Geom& g, const double* X, int64_t N, const double* pack, double* gamma, int64_t ldg,
                         int32_t* labels, double* partial, int grid, size_t smem_limit, int raw, cudaStream_t st) {
  const size_t smem = estep_smem_bytes(g);
  if (smem > smem_limit || g.K8 > 64) return cudaErrorInvalidConfiguration;
  switch (g.DM) {
    case 1: return launch_estep_dm<1>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 2: return launch_estep_dm<2>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 3: return launch_estep_dm<3>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 4: return launch_estep_dm<4>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 5: return launch_estep_dm<5>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 6: return launch_estep_dm<6>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 7: return launch_estep_dm<7>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    case 8: return launch_estep_dm<8>(g, X, N, pack, gamma, ldg, labels, partial, grid, smem, raw, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// reductions / small per-component kernels
// ---------------------------------------------------------------------------------------------------------------
// Four lanes per element: lane q sums the q-th quarter of the slots in slot order, then the four partial sums are combined in
// a fixed tree -- a fixed association (run-to-run deterministic), with a dependent chain a quarter as long as one thread per
// element (these reductions sit on the critical path of every iteration: 296 slots at the headline shape).
__global__ void __launch_bounds__(128) k_reduce_partials(const double* __restrict__ partial, int nblocks, int64_t n, double* __restrict__ out) {
  const int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 2;
  const int q = threadIdx.x & 3;
  const int per = (nblocks + 3) >> 2;
  const int b0 = q * per, b1 = min(nblocks, b0 + per);
  double s = 0;
  if (j < n)
    for (int b = b0; b < b1; ++b) s += partial[(size_t)b * n + j];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (j < n && q == 0) out[j] = s;
}
cudaError_t launch_reduce_partials(const double* partial, int nblocks, int64_t n, double* out, cudaStream_t st) {
  k_reduce_partials<<<(unsigned)((4 * n + 127) / 128), 128, 0, st>>>(partial, nblocks, n, out);
  return cudaGetLastError();
}

// Pass-A statistics of ONE component from given weights (mlf_maxGaussian entry point, src/unsupervised/mlf_gaussian.f90:41-51):
// partial[block][DS] = [sum g, sum g^2, sum g x].  Thread = (point lane, dimension).
__global__ void __launch_bounds__(256) k_stats_from_gamma(const double* __restrict__ X, int64_t N, int d, int DS,
                                                          const double* __restrict__ gamma, double* __restrict__ partial) {
  __shared__ double red[3][256];
  const int tid = threadIdx.x;
  const int lanes = 256 / d;
  const int a = tid % d, pl = tid / d;
  double sx = 0, sg = 0, sg2 = 0;
  if (pl < lanes)
    for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < N; p += (int64_t)gridDim.x * lanes) {
      const double gv = gamma[p];
      sx = fma(gv, X[p * d + a], sx);
      if (a == 0) { sg += gv; sg2 = fma(gv, gv, sg2); }
    }
  red[0][tid] = (pl < lanes) ? sx : 0.0;
  red[1][tid] = sg;
  red[2][tid] = sg2;
  __syncthreads();
  double* o = partial + (size_t)blockIdx.x * DS;
  if (tid < d) {
    double r = 0;
    for (int j = 0; j < lanes; ++j) r += red[0][j * d + tid];
    o[2 + tid] = r;
  } else if (tid < DS - 2) {
    o[2 + tid] = 0.0;
  }
  if (tid == 0) {
    double r1 = 0, r2 = 0;
    for (int j = 0; j < lanes; ++j) { r1 += red[1][j * d]; r2 += red[2][j * d]; }
    o[0] = r1;
    o[1] = r2;
  }
}
cudaError_t launch_stats_from_gamma(const Geom& g, const double* X, int64_t N, const double* gamma, double* partial, int grid,
                                    cudaStream_t st) {
  k_stats_from_gamma<<<grid, 256, 0, st>>>(X, N, g.d, g.DS, gamma, partial);
  return cudaGetLastError();
}

// mlf_MaxGaussian / GECovMatrix prologue: W = P/s, Mu = A.W, minW  (src/unsupervised/mlf_gaussian.f90:47-50,
// src/mlf_matrix.f90:132).  "W(i) < minW -> CYCLE" becomes gamma < minW*s.
__global__ void k_mid(Geom g, const double* __restrict__ statsA, double minW, double* __restrict__ Mu,
                      double* __restrict__ mupad, double* __restrict__ thr) {
  const int k = blockIdx.x;
  const double* sa = statsA + (size_t)k * g.DS;
  const double s = sa[0];
  for (int a = threadIdx.x; a < 8 * g.DN; a += blockDim.x) {
    const double m = (a < g.d) ? sa[2 + a] / s : 0.0;
    mupad[(size_t)k * 8 * g.DN + a] = m;
    if (a < g.d) Mu[(size_t)k * g.d + a] = m;
  }
  if (threadIdx.x == 0) thr[k] = minW * s;
}
cudaError_t launch_mid(const Geom& g, const double* statsA, double minW, double* Mu, double* mupad, double* thr,
                       cudaStream_t st) {
  k_mid<<<g.K, 32, 0, st>>>(g, statsA, minW, Mu, mupad, thr);
  return cudaGetLastError();
}

// GECovMatrix epilogue + MaxStep tail (src/mlf_matrix.f90:119,137,145-146; src/unsupervised/mlf_emgmm.f90:184-187).
__global__ void k_finalize(Geom g, const double* __restrict__ statsA, const double* __restrict__ statsB, double minProba,
                           double* __restrict__ Cov, double* __restrict__ lambda, int cov_diag) {
  const int d = g.d;
  for (int k = blockIdx.x; k < g.K; k += gridDim.x) {
    const double s = statsA[(size_t)k * g.DS];
    const double* sb = statsB + (size_t)k * g.MS;
    const double sw2 = sb[g.MB * 64] / (s * s);
    const double ms = 1.0 / (1.0 - sw2);
    const double scale = ms * ms / s;
    double* Ck = Cov + (size_t)k * d * d;
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
      const int a = i % d, b = i / d;
      if (a < b) continue;  // lower triangle a >= b, mirrored (SymmetrizeMatrix, src/mlf_matrix.f90:56-63)
      const int ab = a >> 3, bb = b >> 3;
      const double v = (cov_diag && a != b) ? 0.0 : scale * sb[(ab * (ab + 1) / 2 + bb) * 64 + (a & 7) * 8 + (b & 7)];
      Ck[(size_t)b * d + a] = v;
      Ck[(size_t)a * d + b] = v;
    }
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
cudaError_t launch_finalize(const Geom& g, const double* statsA, const double* statsB, double minProba, double* Cov,
                            double* lambda, cudaStream_t st, int cov_diag) {
  k_finalize<<<g.K < 148 ? g.K : 148, 128, 0, st>>>(g, statsA, statsB, minProba, Cov, lambda, cov_diag);
  return cudaGetLastError();
}

__global__ void k_sum_k(const double* __restrict__ v, int K, double* __restrict__ out) {
  double s = 0;
  for (int k = 0; k < K; ++k) s += v[k];
  *out = s;
}
cudaError_t launch_sum_k(const double* v, int K, double* out, cudaStream_t st) {
  k_sum_k<<<1, 1, 0, st>>>(v, K, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ X, int64_t N, int d, double* __restrict__ partial) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const int lanes = 256 / d;  // point lanes per block (d <= 128)
  const int a = tid % d, pl = tid / d;
  double s = 0;
  if (pl < lanes)
    for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < N; p += (int64_t)gridDim.x * lanes) s += X[p * d + a];
  red[tid] = (pl < lanes) ? s : 0.0;
  __syncthreads();
  if (tid < d) {
    double r = 0;
    for (int j = 0; j < lanes; ++j) r += red[j * d + tid];
    partial[(size_t)blockIdx.x * d + tid] = r;
  }
}
cudaError_t launch_colsum(const double* X, int64_t N, int d, double* partial, int grid, cudaStream_t st) {
  k_colsum<<<grid, 256, 0, st>>>(X, N, d, partial);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// M-pass
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMstepWarps = 16;
constexpr int kMstepBatch = 8;  // gamma strips in flight per warp

template <int DN, int CPW>
__global__ void __launch_bounds__(kMstepWarps * 32, 1)
k_mstep(const double* __restrict__ X, int64_t N, int d, int K, const double* __restrict__ gamma, int64_t ldg,
        const double* __restrict__ mupad, const double* __restrict__ thr, double* __restrict__ partial, int MS) {
  constexpr int MB = DN * (DN + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int cbase = (blockIdx.y * kMstepWarps + warp) * CPW;
  double acc[CPW][MB][2];
  double sg2[CPW], muf[CPW][DN], th[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int k = cbase + c;
    sg2[c] = 0;
    th[c] = (k < K) ? thr[k] : INFINITY;
#pragma unroll
    for (int b = 0; b < MB; ++b) { acc[c][b][0] = 0; acc[c][b][1] = 0; }
#pragma unroll
    for (int ab = 0; ab < DN; ++ab) muf[c][ab] = (k < K) ? mupad[(size_t)k * 8 * DN + 8 * ab + g] : 0.0;
  }
  const int64_t ntiles = (N + kMstepTilePts - 1) / kMstepTilePts;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t ibeg = tile * kMstepTilePts;
    const int64_t iend = (ibeg + kMstepTilePts < N) ? ibeg + kMstepTilePts : N;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const int k = cbase + c;
      if (k >= K) continue;
      const double* gk = gamma ? gamma + (size_t)k * ldg : nullptr;
      for (int64_t i0 = ibeg; i0 < iend; i0 += 32 * kMstepBatch) {
        double gv[kMstepBatch];
#pragma unroll
        for (int u = 0; u < kMstepBatch; ++u) {
          const int64_t p = i0 + 32 * u + lane;
          gv[u] = (p < iend) ? (gk ? gk[p] : 1.0) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kMstepBatch; ++u) {
          const double gval = gv[u];
          sg2[c] = fma(gval, gval, sg2[c]);  // SUM(W*W) runs over all points (src/mlf_matrix.f90:119)
          const bool kept = !(gval < th[c]) && (i0 + 32 * u + lane < iend);  // "if(W(i)<minW0) CYCLE" (:135)
          const unsigned mask = __ballot_sync(0xffffffffu, kept);
          if (mask == 0) continue;
          unsigned gm = mask | (mask >> 1);
          gm |= gm >> 2;
          gm &= 0x11111111u;  // bit 4q set <=> some point of 4-group q is kept
          while (gm) {
            const int q = (__ffs(gm) - 1) >> 2;
            gm &= gm - 1;
            const double wq = __shfl_sync(0xffffffffu, gval, 4 * q + t);
            const double w = ((mask >> (4 * q + t)) & 1u) ? wq : 0.0;
            const int64_t pp = i0 + 32 * u + 4 * q + t;
            double z[DN], wz[DN];
#pragma unroll
            for (int ab = 0; ab < DN; ++ab) {
              const int dim = 8 * ab + g;
              const double xa = (pp < N && dim < d) ? X[pp * d + dim] : 0.0;
              z[ab] = (dim < d) ? xa - muf[c][ab] : 0.0;
              wz[ab] = w * z[ab];
            }
#pragma unroll
            for (int ab = 0; ab < DN; ++ab)
#pragma unroll
              for (int bb = 0; bb <= ab; ++bb) dmma884(acc[c][ab * (ab + 1) / 2 + bb], wz[ab], z[bb]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int k = cbase + c;
    if (k >= K) continue;
    double* o = partial + ((size_t)blockIdx.x * K + k) * MS;
#pragma unroll
    for (int b = 0; b < MB; ++b) {
      o[b * 64 + g * 8 + 2 * t] = acc[c][b][0];
      o[b * 64 + g * 8 + 2 * t + 1] = acc[c][b][1];
    }
    double s = sg2[c];
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += shfl_xor_d(s, sh);
    if (lane == 0) { o[MB * 64] = s; o[MB * 64 + 1] = 0.0; }
  }
}

int mstep_cpw(const Geom& g) {
  const int cmax = g.DN == 1 ? 4 : (g.DN == 2 ? 4 : (g.DN == 3 ? 2 : 1));
  int c = (g.K + kMstepWarps - 1) / kMstepWarps;
  if (c > cmax) c = cmax;
  if (c == 3) c = (cmax >= 4) ? 4 : 2;
  return c < 1 ? 1 : c;
}
void mstep_grid(const Geom& g, int sms, int64_t N, int* gx, int* gy) {
  const int cpw = mstep_cpw(g);
  *gy = (g.K + kMstepWarps * cpw - 1) / (kMstepWarps * cpw);
  const int64_t ntiles = (N + kMstepTilePts - 1) / kMstepTilePts;
  int64_t x = ntiles < sms ? ntiles : sms;
  *gx = (int)(x < 1 ? 1 : x);
}

template <int DN, int CPW>
static cudaError_t launch_mstep_t(const Geom& g, const double* X, int64_t N, const double* gamma, int64_t ldg,
                                  const double* mupad, const double* thr, double* partial, dim3 grid, cudaStream_t st) {
  k_mstep<DN, CPW><<<grid, kMstepWarps * 32, 0, st>>>(X, N, g.d, g.K, gamma, ldg, mupad, thr, partial, g.MS);
  return cudaGetLastError();
}

cudaError_t launch_mstep(const Geom& g, const double* X, int64_t N, const double* gamma, int64_t ldg, const double* mupad,
                         const double* thr, double* partial, int gridx, cudaStream_t st) {
  const int cpw = mstep_cpw(g);
  const int gy = (g.K + kMstepWarps * cpw - 1) / (kMstepWarps * cpw);
  const dim3 grid(gridx, gy);
#define MLF_M(DNv, CPWv) \
  if (g.DN == DNv && cpw == CPWv) return launch_mstep_t<DNv, CPWv>(g, X, N, gamma, ldg, mupad, thr, partial, grid, st);
  MLF_M(1, 1) MLF_M(1, 2) MLF_M(1, 4)
  MLF_M(2, 1) MLF_M(2, 2) MLF_M(2, 4)
  MLF_M(3, 1) MLF_M(3, 2)
  MLF_M(4, 1)
#undef MLF_M
  return cudaErrorInvalidConfiguration;
}

// ---------------------------------------------------------------------------------------------------------------
// k-means Lloyd step (initialiser; src/unsupervised/mlf_kmeans.f90:151-165, mlf_kmeans_naive.f90:74-96)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_kmeans_assign(const double* __restrict__ X, int64_t N, int d, int K,
                                                       const double* __restrict__ Mu, int32_t* __restrict__ cl,
                                                       double* __restrict__ minDist, double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* smu = sm;                     // [K][d]
  double* ssum = sm + (size_t)K * d;    // [K][d+1]
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) smu[i] = Mu[i];
  for (int i = threadIdx.x; i < K * (d + 1); i += blockDim.x) ssum[i] = 0.0;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const double* x = X + p * d;
    int best = 0;
    double bd = INFINITY;
    for (int k = 0; k < K; ++k) {
      double s = 0;
      for (int a = 0; a < d; ++a) { const double tdiff = x[a] - smu[(size_t)k * d + a]; s = fma(tdiff, tdiff, s); }
      const double dist = sqrt(s);  // norm2
      if (dist < bd) { bd = dist; best = k; }  // MINLOC: first minimum
    }
    if (cl) cl[p] = best + 1;
    if (minDist) minDist[p] = bd;
    if (partial) {
      for (int a = 0; a < d; ++a) atomicAdd(&ssum[(size_t)best * (d + 1) + a], x[a]);
      atomicAdd(&ssum[(size_t)best * (d + 1) + d], 1.0);
    }
  }
  __syncthreads();
  if (partial)
    for (int i = threadIdx.x; i < K * (d + 1); i += blockDim.x) partial[(size_t)blockIdx.x * K * (d + 1) + i] = ssum[i];
}
cudaError_t launch_kmeans_assign(const double* X, int64_t N, int d, int K, const double* Mu, int32_t* cl, double* minDist,
                                 double* partial, int grid, cudaStream_t st) {
  const size_t smem = ((size_t)K * d + (size_t)K * (d + 1)) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_kmeans_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_kmeans_assign<<<grid, 256, smem, st>>>(X, N, d, K, Mu, cl, minDist, partial);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// k-means seeding (mlf_InitKMeans / mlf_GetFarPoint, src/unsupervised/mlf_kmeans.f90:140-149,168-182): the next
// centre is a data point drawn with probability proportional to minDist^2.  Block b owns the contiguous
// chunk [b*chunk, (b+1)*chunk) so that the host can locate the drawn point from the per-block sums.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_seed_update(const double* __restrict__ X, int64_t N, int d,
                                                     const double* __restrict__ centre, double* __restrict__ minDist,
                                                     int first, int64_t chunk, double* __restrict__ partial) {
  __shared__ double sc[128];
  __shared__ double red[256];
  const int tid = threadIdx.x;
  for (int a = tid; a < d; a += 256) sc[a] = centre[a];
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * chunk;
  const int64_t hi = (lo + chunk < N) ? lo + chunk : N;
  double s = 0;
  for (int64_t p = lo + tid; p < hi; p += 256) {
    const double* x = X + p * d;
    double q = 0;
    for (int a = 0; a < d; ++a) { const double t = x[a] - sc[a]; q = fma(t, t, q); }
    double m = sqrt(q);  // norm2 (src/unsupervised/mlf_kmeans.f90:160)
    if (!first) m = fmin(minDist[p], m);
    minDist[p] = m;
    s = fma(m, m, s);
  }
  red[tid] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) red[tid] += red[tid + st];
    __syncthreads();
  }
  if (tid == 0) partial[blockIdx.x] = red[0];
}
cudaError_t launch_seed_update(const double* X, int64_t N, int d, const double* centre, double* minDist, int first,
                               int64_t chunk, int grid, double* partial, cudaStream_t st) {
  k_seed_update<<<grid, 256, 0, st>>>(X, N, d, centre, minDist, first, chunk, partial);
  return cudaGetLastError();
}

// First index i in [lo, hi) whose running sum of minDist^2 reaches `target` (mlf_rand_class / mlf_di_search on the
// cumulative vector, src/mlf_rand.f90:173-182); copies that point into out[0..d).  One block.
__global__ void __launch_bounds__(256) k_seed_pick(const double* __restrict__ X, int d, const double* __restrict__ minDist,
                                                   int64_t lo, int64_t hi, double target, double* __restrict__ out) {
  __shared__ double loc[256], pre[256];
  __shared__ long long s_idx;
  const int tid = threadIdx.x;
  const int64_t n = hi - lo, sub = (n + 255) / 256;
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
cudaError_t launch_seed_pick(const double* X, int d, const double* minDist, int64_t lo, int64_t hi, double target,
                             double* out, cudaStream_t st) {
  k_seed_pick<<<1, 256, 0, st>>>(X, d, minDist, lo, hi, target, out);
  return cudaGetLastError();
}

// Pack of the k-means assignment for the E-pass kernels (hard mode): identity factor in fragment order (or its diagonal for
// the compressed layout), accumulator bias -(centre - shift), constant 0; empty / dummy components get -inf.
__global__ void k_pack_kmeans(Geom g, int PSF, int OFFB, int OFFC, int diag, const double* __restrict__ Mu,
                              const double* __restrict__ shift, double* __restrict__ pack2) {
  const int k = blockIdx.x, d = g.d;
  double* p2 = pack2 + (size_t)k * PSF;
  const bool live = k < g.K && isfinite(Mu[(size_t)(k < g.K ? k : 0) * d]);
  if (diag) {
    for (int n = threadIdx.x; n < 8 * g.DN; n += blockDim.x) p2[n] = (live && n < d) ? 1.0 : 0.0;
  } else {
    for (int i = threadIdx.x; i < g.NB * 32; i += blockDim.x) {
      const int blk = i >> 5, l = i & 31;
      int mb = 0;
      while (blk_off(mb + 1) <= blk) ++mb;
      const int nb = blk - blk_off(mb);
      const int m = 4 * mb + (l & 3), n = 8 * nb + (l >> 2);
      p2[i] = (live && m == n && m < d) ? 1.0 : 0.0;
    }
  }
  for (int n = threadIdx.x; n < 8 * g.DN; n += blockDim.x) p2[OFFB + n] = (live && n < d) ? -(Mu[(size_t)k * d + n] - shift[n]) : 0.0;
  if (threadIdx.x == 0) { p2[OFFC] = live ? 0.0 : -INFINITY; p2[OFFC + 1] = 0; p2[OFFC + 2] = 0; p2[OFFC + 3] = 0; }
}
cudaError_t launch_pack_kmeans(const Geom& g, const FusedGeom& f, const double* Mu, const double* shift, double* pack2, cudaStream_t st) {
  k_pack_kmeans<<<g.K8 * 8, 128, 0, st>>>(g, f.PSF, f.OFFB, f.OFFC, f.diag ? 1 : 0, Mu, shift, pack2);
  return cudaGetLastError();
}

// EvaluateCentres (src/unsupervised/mlf_kmeans_naive.f90:74-96): Mu_k = sum_k / n_k; empty clusters are flagged
// (count 0) and given an infinite centre so that they attract no point until re-seeded by the host logic.
__global__ void k_kmeans_centres(const double* __restrict__ sums, int K, int d, double* __restrict__ Mu,
                                 double* __restrict__ counts) {
  const int k = blockIdx.x;
  const double n = sums[(size_t)k * (d + 1) + d];
  for (int a = threadIdx.x; a < d; a += blockDim.x) Mu[(size_t)k * d + a] = n > 0 ? sums[(size_t)k * (d + 1) + a] / n : INFINITY;
  if (threadIdx.x == 0) counts[k] = n;
}
// Same, reading the pass-A statistics layout of the E-pass kernels ([count, sum g^2, sums...] per component, stride DS).
__global__ void k_kmeans_centres_a(const double* __restrict__ statsA, int DS, int K, int d, double* __restrict__ Mu,
                                   double* __restrict__ counts) {
  const int k = blockIdx.x;
  const double n = statsA[(size_t)k * DS];
  for (int a = threadIdx.x; a < d; a += blockDim.x) Mu[(size_t)k * d + a] = n > 0 ? statsA[(size_t)k * DS + 2 + a] / n : INFINITY;
  if (threadIdx.x == 0) counts[k] = n;
}
cudaError_t launch_kmeans_centres_a(const Geom& g, const double* statsA, double* Mu, double* counts, cudaStream_t st) {
  k_kmeans_centres_a<<<g.K, 32, 0, st>>>(statsA, g.DS, g.K, g.d, Mu, counts);
  return cudaGetLastError();
}
cudaError_t launch_kmeans_centres(const double* sums, int K, int d, double* Mu, double* counts, cudaStream_t st) {
  k_kmeans_centres<<<K, 32, 0, st>>>(sums, K, d, Mu, counts);
  return cudaGetLastError();
}

// Cov_k = I (IdentityMatrix, src/mlf_matrix.f90:90-95), lambda_k = 1/K  (src/unsupervised/mlf_emgmm.f90:141-145)
__global__ void k_identity_cov(int K, int d, double* __restrict__ Cov, double* __restrict__ lambda) {
  const int k = blockIdx.x;
  for (int i = threadIdx.x; i < d * d; i += blockDim.x) Cov[(size_t)k * d * d + i] = (i % d == i / d) ? 1.0 : 0.0;
  if (threadIdx.x == 0) lambda[k] = 1.0 / (double)K;
}
cudaError_t launch_identity_cov(int K, int d, double* Cov, double* lambda, cudaStream_t st) {
  k_identity_cov<<<K, 128, 0, st>>>(K, d, Cov, lambda);
  return cudaGetLastError();
}

}  // namespace mlfb200
