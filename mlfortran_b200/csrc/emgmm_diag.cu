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
#include "exp_tab.cuh"

#include <climits>
#include <cmath>
#include <cstdlib>

namespace mlfb200 {

namespace {

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

constexpr int kD = 8;     // dimensions held in registers (d <= 8; padding dimensions are zero)
constexpr int kXS = 12;   // row stride of the staged points: 8 coordinates, 1/sum, padding

// row stride of the logit tile: the P2 access (TP/8 rows per warp, 32/(TP/8) lanes of 8 bytes each) must spread over the banks
int emd_stride(int KP, int tp) { return ((KP + 15) / 16) * 16 + (tp == 64 ? 4 : 8); }
size_t emd_smem(int KP, int tp) {
  return ((size_t)tp * kXS + (size_t)tp * emd_stride(KP, tp) + 128 + 8 * (size_t)KP + (size_t)KP) * sizeof(double) + 8 * (size_t)tp * sizeof(int);
}

}  // namespace

bool emd_ok(const Geom& g, const FusedGeom& f, size_t smem_limit) {
  return f.diag && g.DN == 1 && g.K8 * 8 <= 256 && emd_smem(g.K8 * 8, 64) <= smem_limit;
}

// pack2: diagonal-compressed layout of FusedGeom (DN = 1): [8 diag(L) | 8 bias | const, pad3] = 20 doubles per component.
// 8 warps, one lane per component, 238 registers per thread: two warps per scheduler hide the fp64 latencies through the
// 8-point / 4-point blocking below.  Measured alternatives (config-5 shard, 185 ms with this kernel): TP = 32 with two blocks
// per SM spills at 128 registers (457 ms); 16 warps with a pair of lanes per component (half the dimensions each, 128
// registers, no spills) doubles the shuffle / REDUX / shared-memory instructions per pair and lands at 210 ms.
// The statistics loop is bound by instruction issue (shared-memory loads first), not by the fp64 pipe: threshold tests as one
// comparison and strength-reduced addressing took 185 -> 174 ms, sum g x' as DMMA tiles 174 -> 166, the whole main path of P3 in
// the DMMA fragment layout 166 -> 160 (MLF_B200_EMD_XMMA=0 keeps the lane = component form).  No gain: driving the kept /
// listed branches from one warp-wide OR of per-point flags, computed in the main path (177 ms) or inside the rare path (161 ms).
// QP1 (round 2): P1 on the tensor datapath after all -- not as tiles of the diagonal factor (7/8 zeros) but as the expanded form
// logit = [c - 1/2 sum b_a^2] - sum_a (l_a b_a) x'_a - 1/2 sum_a l_a^2 x'_a^2: a dense dot product of 16 coefficients with the 16
// monomials (x'_a, x'_a^2) of the point = 4 DMMA k-steps for 8 points x 8 components, with the warp's coefficient fragments
// (4 blocks x 4 k-steps) in registers for the whole sweep.  Same datapath time as the 2 DFMA per dimension of the lane = component
// form, but 16 DMMA + 2 loads per 8 points x 32 components instead of 128 DFMA + 32 broadcast loads: P1 was bound by instruction
// issue, not by the fp64 pipe.  The expanded form cancels (terms ~ (distance from the shift / width)^2), so it is taken only
// when the bound S_k that k_refresh leaves for every component keeps the estimated logit error under the engine's limit
// (A.qfS, A.qf_limit; decided by every block, as in k_em16); the other variant returns at once.
// Measured on the config-5 shard (N=1.25e8, d=8, K=256): 203.6 -> 200.2 ms per 1.4-sweep iteration only -- the lane = component P1 already ran
// near the datapath rate (136 DFMA against 16 DMMA per 8 points x 32 components hold the fp64 pipe for the same 270 clocks);
// what keeps this kernel at ~45 % of the pipe are the latency-bound softmax and statistics phases at two warps per scheduler.
// Also measured with QP1 and not kept: the kept-set sums in the fragment layout as raw moments about the shift (sum_kept g x',
// sum_kept g x'^2 as two DMMA per block and candidate group, re-centred at mu_old in the epilogue) in place of the lane = component
// path: 233 ms per iteration against 200 -- four votes, two comparisons and a ballot per 4-point group and block cost more issue
// slots than the one vote per group of the existing path saves.
template <int TP, bool XMMA, bool QP1>
__global__ void __launch_bounds__(256, 64 / TP) k_emd(EmArgs A) {
  constexpr int PSF = 20, XS = kXS, DS = 10, MSF = 74;
  constexpr int GP = TP / 8;       // points per warp in P2
  constexpr int TT = 32 / GP;      // lanes per point in P2
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, S = A.S, d = A.d;
  const int64_t N = A.N;
  double* xs = smem;                       // [TP][12]: x' = x - shift in columns 0..7 (zero padded), 1/sum of the point in column 8
  double* tile = xs + (size_t)TP * XS;     // [TP][S] logits -> softmax numerators
  double* etab = tile + (size_t)TP * S;    // [128] 2^(j/128)
  double* mus = etab + 128;                // [8][KP] scatter centres mu_old - shift, dimension-major
  int* wmax = reinterpret_cast<int*>(mus + 8 * (size_t)KP);   // [8 warps][TP] ordered key of the warp's largest logit, per point
  double* cst = reinterpret_cast<double*>(wmax + 8 * TP);     // [KP] QP1: constant of the expanded logit
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane / TT, t = lane % TT;
  {  // which P1 runs: every S_k inside the limit -> expanded form on the tensor datapath, otherwise (or NaN, or no limit) lane = component
    bool bad = !(A.qf_limit >= 0.0) || A.qfS == nullptr;
    if (!bad)
      for (int i = tid; i < KP; i += 256) bad |= !(A.qfS[i] <= A.qf_limit);
    const bool use_q = __syncthreads_or(bad ? 1 : 0) == 0;
    if (use_q != QP1) return;
  }
  const int comp = 32 * warp + lane;       // the component this lane owns in P1 / P3
  const bool own = comp < KP;
  const bool wact = 32 * warp < KP;        // warp-uniform: this warp owns at least one component

  // component parameters and statistics live in registers for the whole sweep
  double l[kD], b[kD], ck = -INFINITY, thi = INFINITY, tlo = INFINITY;
#pragma unroll
  for (int a = 0; a < kD; ++a) {
    l[a] = own ? A.pack2[(size_t)comp * PSF + a] : 0.0;
    b[a] = own ? A.pack2[(size_t)comp * PSF + 8 + a] : 0.0;
    if (own) mus[a * KP + comp] = A.muold[(size_t)comp * 8 + a] - A.shift[a];
  }
  if (own) { ck = A.pack2[(size_t)comp * PSF + 16]; thi = A.thi[comp]; tlo = A.tlo[comp]; }
  const int gq = lane >> 2, tq = lane & 3;
  double thq[4][4];   // QP1: coefficient fragments (B operand): block b, k-step s -> coefficient of monomial 4 s + tq for component 32 warp + 8 b + gq
  if (QP1) {
    double bb = 0;
#pragma unroll
    for (int a = 0; a < kD; ++a) bb = fma(b[a], b[a], bb);
    if (own) cst[comp] = ck - 0.5 * bb;   // -inf for dummy components
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
      const int c = 32 * warp + 8 * bq + gq;
      const bool ok = c < KP;
      const double l0 = ok ? A.pack2[(size_t)c * PSF + tq] : 0.0, l1 = ok ? A.pack2[(size_t)c * PSF + 4 + tq] : 0.0;
      const double b0 = ok ? A.pack2[(size_t)c * PSF + 8 + tq] : 0.0, b1 = ok ? A.pack2[(size_t)c * PSF + 12 + tq] : 0.0;
      thq[bq][0] = -l0 * b0; thq[bq][1] = -l1 * b1;               // x'_tq, x'_{4+tq}
      thq[bq][2] = -0.5 * l0 * l0; thq[bq][3] = -0.5 * l1 * l1;   // x'_tq^2, x'_{4+tq}^2
    }
  }
  // "not below the threshold and not zero" as one comparison: responsibilities are >= 0 (or NaN), so
  // !(a < t) && a != 0  <=>  !(a < max(t, smallest subnormal)); NaN stays a candidate as before
  const double tlo_e = fmax(tlo, 4.9406564584124654e-324), thi_e = fmax(thi, 4.9406564584124654e-324);
  const bool accu = A.accumulate != 0;
  double* oA = A.partA + ((size_t)blockIdx.x * KP + (own ? comp : 0)) * DS;
  double* oB = A.partB + ((size_t)blockIdx.x * KP + (own ? comp : 0)) * MSF;
  double sg = 0, sg2 = 0, sk = 0, sgx[kD], skx[kD], skz2[kD];
  // XMMA: sum g x' as DMMA tiles [8 components x 4 points].[4 points x 8 dims], one per 8-component block of the warp; lane
  // (gq, tq) holds dims 2 tq, 2 tq + 1 of component 32 warp + 8 b + gq.  One shared-memory load per block instead of the four
  // broadcast LDS.128 per point of the lane = component form (the P3 loop is bound by shared-memory instruction issue).
  double accX[4][2], sgq[4], sg2q[4], tloq[4];   // sum g and sum g^2 ride along in the same layout (per lane: its point column)
#pragma unroll
  for (int a = 0; a < kD; ++a) {
    const bool ld = accu && own && a < d;
    // the partial slots hold raw sums (see the end of the kernel): back to the shifted / centred accumulators
    sgx[a] = (ld && !XMMA) ? oA[2 + a] - A.shift[a] * oA[0] : 0.0;
    skx[a] = ld ? oB[64 + a] - A.muold[(size_t)comp * 8 + a] * oB[72] : 0.0;
    skz2[a] = ld ? oB[a * 9] : 0.0;   // diagonal entry (a, a) of the 8x8 scatter block
  }
  if (accu && own) { sg = oA[0]; sg2 = oA[1]; sk = oB[72]; }   // (sg, sg2: lane = component form only)
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int c = 32 * warp + 8 * b + gq;
    const double* oc = A.partA + ((size_t)blockIdx.x * KP + (c < KP ? c : 0)) * DS;
    const bool ld = XMMA && accu && c < KP;
#pragma unroll
    for (int i = 0; i < 2; ++i) accX[b][i] = (ld && 2 * tq + i < d) ? oc[2 + 2 * tq + i] - A.shift[2 * tq + i] * oc[0] : 0.0;
    sgq[b] = (ld && tq == 0) ? oc[0] : 0.0;
    sg2q[b] = (ld && tq == 0) ? oc[1] : 0.0;
    tloq[b] = (c < KP) ? fmax(A.tlo[c], 4.9406564584124654e-324) : INFINITY;
  }
  const int nlist = blockIdx.x * 8 + warp;
  int nAmb = accu ? A.amb_count[nlist] : 0;
  AmbEntry* ambw = A.amb + (size_t)nlist * A.amb_cap;

  if (tid < 128) etab[tid] = kExpTab128[tid];
  if (!wact)
    for (int i = lane; i < TP; i += 32) wmax[warp * TP + i] = INT_MIN;   // never written again
  const int amb_cap = A.amb_cap;
  const long long key0 = (long long)A.p_offset * 1024 + comp;

  const int64_t ntiles = (N + TP - 1) / TP;
  const int spt = tid >> 3, sa = tid & 7;   // staging role: point spt (and spt + 32), dimension sa
  const double ssh = (sa < d) ? A.shift[sa] : 0.0;
  double nx0, nx1 = 0.0;   // raw prefetched coordinates (the shift itself where there is no point: x - shift = 0); the shift is
                           // subtracted when the tile is staged, so that nothing waits on the loads before then
  {
    const int64_t qn = (int64_t)blockIdx.x * TP;
    nx0 = (qn + spt < N && sa < d) ? A.X[(qn + spt) * d + sa] : ssh;
    if (TP == 64) nx1 = (qn + spt + 32 < N && sa < d) ? A.X[(qn + spt + 32) * d + sa] : ssh;
  }
  for (int64_t bt = blockIdx.x; bt < ntiles; bt += gridDim.x) {
    const int64_t q0 = bt * TP;
    // ---- stage the tile: x' = x - shift, zero padded to 8 dimensions; rows beyond N are zero.  The values were fetched
    //      into registers one tile ahead, so the global-memory latency hides under the previous tile's arithmetic ------
    xs[spt * XS + sa] = nx0 - ssh;
    if (TP == 64) xs[(spt + 32) * XS + sa] = nx1 - ssh;
    {
      const int64_t qn = (bt + gridDim.x) * TP;
      nx0 = (qn + spt < N && sa < d) ? A.X[(qn + spt) * d + sa] : ssh;
      if (TP == 64) nx1 = (qn + spt + 32 < N && sa < d) ? A.X[(qn + spt + 32) * d + sa] : ssh;
    }
    __syncthreads();
    // ---- P1: lane = component; logits of PB points at a time (PB independent chains), and the warp's maximum per point ----
    if (wact && QP1) {
#pragma unroll 1
      for (int pt0 = 0; pt0 < TP; pt0 += 8) {
        const double xlo = xs[(pt0 + gq) * XS + tq], xhi = xs[(pt0 + gq) * XS + 4 + tq];
        const double x2lo = xlo * xlo, x2hi = xhi * xhi;
        double c[4][2];
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
          const int cc = (32 * warp + 8 * bq < KP) ? 32 * warp + 8 * bq + 2 * tq : 0;
          const double2 v = *reinterpret_cast<const double2*>(cst + cc);
          c[bq][0] = v.x; c[bq][1] = v.y;
        }
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) dmma(c[bq], xlo, thq[bq][0]);
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) dmma(c[bq], xhi, thq[bq][1]);
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) dmma(c[bq], x2lo, thq[bq][2]);
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) dmma(c[bq], x2hi, thq[bq][3]);
        double mx = -INFINITY;
#pragma unroll
        for (int bq = 0; bq < 4; ++bq)
          if (32 * warp + 8 * bq < KP) {   // warp-uniform
            *reinterpret_cast<double2*>(tile + (size_t)(pt0 + gq) * S + 32 * warp + 8 * bq + 2 * tq) = make_double2(c[bq][0], c[bq][1]);
            mx = fmax(mx, fmax(c[bq][0], c[bq][1]));
          }
        const int hi = __double2hiint(mx);
        int mk = hi ^ ((hi >> 31) & 0x7fffffff);
        mk = max(mk, __shfl_xor_sync(kFull, mk, 1));
        mk = max(mk, __shfl_xor_sync(kFull, mk, 2));
        if (tq == 0) wmax[warp * TP + pt0 + gq] = mk;
      }
    }
    if (wact && !QP1) {
      constexpr int PB = (TP == 64) ? 8 : 4;
#pragma unroll 1
      for (int pt0 = 0; pt0 < TP; pt0 += PB) {
        double v[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j) v[j] = 0.0;
#pragma unroll
        for (int a2 = 0; a2 < kD / 2; ++a2) {
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const double2 x = *reinterpret_cast<const double2*>(xs + (pt0 + j) * XS + 2 * a2);   // broadcast load
            const double u0 = fma(l[2 * a2], x.x, b[2 * a2]);
            const double u1 = fma(l[2 * a2 + 1], x.y, b[2 * a2 + 1]);
            v[j] = fma(u0, u0, v[j]);
            v[j] = fma(u1, u1, v[j]);
          }
        }
        // the warp's largest logit per point, to 2^-20 relative: one integer REDUX on the order-preserving image of the
        // high word (the softmax only needs the offset to be within a few hundred of the true maximum; see P2)
        int mk = INT_MIN;
#pragma unroll
        for (int j = 0; j < PB; ++j) {
          v[j] = fma(-0.5, v[j], ck);                  // lanes without a component: -inf
          if (own) tile[(size_t)(pt0 + j) * S + comp] = v[j];
          const int hi = __double2hiint(v[j]);
          const int r = __reduce_max_sync(kFull, hi ^ ((hi >> 31) & 0x7fffffff));
          mk = (lane == j) ? r : mk;
        }
        if (lane < PB) wmax[warp * TP + pt0 + lane] = mk;
      }
    }
    __syncthreads();
    // ---- P2: warp = GP points, lane (g, t) = point g, components == t (mod TT): softmax numerators and 1/sum ------------
    {
      const int pl = warp * GP + g;
      const int64_t p = q0 + pl;
      double* row = tile + (size_t)pl * S;
      const int cnt = KP / TT;
      int mk = INT_MIN;
#pragma unroll
      for (int w = t; w < 8; w += TT) mk = max(mk, wmax[w * TP + pl]);
#pragma unroll
      for (int s2 = 1; s2 < TT; s2 <<= 1) mk = max(mk, __shfl_xor_sync(kFull, mk, s2));
      // softmax offset: the largest logit with its low word dropped (|m - max| <= 2^-20 |max|): below 2^27 the numerators
      // neither overflow nor lose their leading terms.  Beyond that, and for label queries, scan for the exact maximum.
      double m = __hiloint2double(mk ^ ((mk >> 31) & 0x7fffffff), 0);
      if (__any_sync(kFull, A.labels != nullptr || !(fabs(m) < 134217728.0))) {
        double mx = -INFINITY;
        int am = 0;
        for (int k = 0; k < cnt; ++k) {
          const double v = row[TT * k + t];
          if (v > mx) { mx = v; am = TT * k + t; }
        }
#pragma unroll
        for (int s2 = 1; s2 < TT; s2 <<= 1) {
          const double om = shx(mx, s2);
          const int oa = __shfl_xor_sync(kFull, am, s2);
          if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
        }
        m = mx;
        if (A.labels != nullptr && p < N && t == 0) A.labels[p] = am + 1;   // getClass: lowest index among the largest
      }
      double sum0 = 0, sum1 = 0;
      int k = 0;
#pragma unroll (TP == 64 ? 4 : 1)
      for (; k + 1 < cnt; k += 2) {
        const double e0 = exp_tab128(row[TT * k + t] - m, etab), e1 = exp_tab128(row[TT * k + TT + t] - m, etab);
        row[TT * k + t] = e0; row[TT * k + TT + t] = e1;
        sum0 += e0; sum1 += e1;
      }
      if (k < cnt) {
        const double e0 = exp_tab128(row[TT * k + t] - m, etab);
        row[TT * k + t] = e0;
        sum0 += e0;
      }
      double sum = sum0 + sum1;
#pragma unroll
      for (int s2 = 1; s2 < TT; s2 <<= 1) sum += shx(sum, s2);
      if (t == 0) xs[pl * XS + 8] = (p < N) ? 1.0 / sum : 0.0;
    }
    __syncthreads();
    // ---- P3: statistics of the tile, 4 points at a time ---------------------------------------------------------------------
    // kept-set sums and side list of one 4-point group, lane = component; a[j] = responsibility of (this lane's component, point j)
    auto kept_and_listed = [&](int pt0, const double (&a)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double aj = a[j];
        const bool hi = !(aj < thi_e);
        const bool mid = !hi && !(aj < tlo_e);
        if (__any_sync(kFull, hi)) {
          const double ah = hi ? aj : 0.0;
          sk += ah;
#pragma unroll
          for (int c = 0; c < kD; ++c) {
            const double z = xs[(pt0 + j) * XS + c] - (own ? mus[c * KP + comp] : 0.0);
            const double az = ah * z;
            skx[c] += az;
            skz2[c] = fma(az, z, skz2[c]);
          }
        }
        const unsigned mm = __ballot_sync(kFull, mid);
        if (mm) {  // undecided until s_k is final: side list, settled by k_amb_resolve
          const int sl = nAmb + __popc(mm & ((1u << lane) - 1u));
          if (mid && sl < amb_cap) {
            AmbEntry e;
            e.g = aj;
            e.key = key0 + (long long)(q0 + pt0 + j) * 1024;
            ambw[sl] = e;
          }
          nAmb += __popc(mm);
        }
      }
    };
    if (wact && XMMA) {
      // main path in DMMA fragment layout: lane (gq, tq) reads the responsibility of (component 32 warp + 8 b + gq, point pt0 + tq)
      // for the four 8-component blocks b of the warp: sum g, sum g^2, sum g x' and the candidate test need one shared-memory
      // load per block; the lane = component form is rebuilt only for groups that hold a candidate of the kept set
#pragma unroll 1
      for (int pt0 = 0; pt0 < TP; pt0 += 4) {
        const double invq = xs[(pt0 + tq) * XS + 8], xq = xs[(pt0 + tq) * XS + gq];
        const double* tq_p = tile + (size_t)(pt0 + tq) * S + 32 * warp + gq;
        bool cand = false;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (32 * warp + 8 * b < KP) {   // warp-uniform: whole blocks exist or not
            const double af = tq_p[8 * b] * invq;
            dmma(accX[b], af, xq);
            sgq[b] += af;
            sg2q[b] = fma(af, af, sg2q[b]);
            // "If(W(i) < minW0) CYCLE" (src/mlf_matrix.f90:135): candidates for the kept set reach the lower band edge
            cand = cand || !(af < tloq[b]);
          }
        if (__any_sync(kFull, cand)) {
          double a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = own ? tile[(size_t)(pt0 + j) * S + comp] * xs[(pt0 + j) * XS + 8] : 0.0;
          kept_and_listed(pt0, a);
        }
      }
    }
    if (wact && !XMMA) {   // lane = component throughout; lanes without a component carry g = 0
      double en[4];
      const double* tcol = tile + (own ? comp : 0);   // lanes without a component read a valid column and discard it
      const int S2 = 2 * S, S3 = 3 * S, S4 = 4 * S;
      en[0] = own ? tcol[0] : 0.0; en[1] = own ? tcol[S] : 0.0; en[2] = own ? tcol[S2] : 0.0; en[3] = own ? tcol[S3] : 0.0;
#pragma unroll 1
      for (int pt0 = 0; pt0 < TP; pt0 += 4) {
        double a[4];
        bool cand = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a[j] = en[j] * xs[(pt0 + j) * XS + 8];
          cand = cand || !(a[j] < tlo_e);   // g == 0 adds nothing and is never listed
        }
        if (pt0 + 4 < TP) {   // next group, ahead of the votes
          tcol += S4;
          en[0] = own ? tcol[0] : 0.0; en[1] = own ? tcol[S] : 0.0; en[2] = own ? tcol[S2] : 0.0; en[3] = own ? tcol[S3] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sg += a[j];
          sg2 = fma(a[j], a[j], sg2);
#pragma unroll
          for (int c2 = 0; c2 < kD / 2; ++c2) {
            const double2 x = *reinterpret_cast<const double2*>(xs + (pt0 + j) * XS + 2 * c2);
            sgx[2 * c2] = fma(a[j], x.x, sgx[2 * c2]);        // about the shift; the shift is added back at the end
            sgx[2 * c2 + 1] = fma(a[j], x.y, sgx[2 * c2 + 1]);
          }
        }
        if (__any_sync(kFull, cand)) kept_and_listed(pt0, a);
      }
    }
    __syncthreads();
  }
  // ---- per-block partials in the layouts of the fused kernels ---------------------------------------------------------------
  if (own) {
    if (!XMMA) { oA[0] = sg; oA[1] = sg2; }
    // sum g x = sum g x' + shift * sum g;  sum_kept g x = sum_kept g z + mu_old * sum_kept g
#pragma unroll
    for (int a = 0; a < kD; ++a) {
      if (!XMMA) oA[2 + a] = (a < d) ? fma(A.shift[a], sg, sgx[a]) : 0.0;
      oB[64 + a] = (a < d) ? skx[a] + A.muold[(size_t)comp * 8 + a] * sk : 0.0;
    }
    for (int i = 0; i < 64; ++i) oB[i] = 0.0;
#pragma unroll
    for (int a = 0; a < kD; ++a) oB[a * 9] = skz2[a];
    oB[72] = sk;
    oB[73] = 0.0;
  }
  if (XMMA && wact) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int c = 32 * warp + 8 * b + gq;
      double sgc = sgq[b], sg2c = sg2q[b];
      sgc += shx(sgc, 1); sgc += shx(sgc, 2);
      sg2c += shx(sg2c, 1); sg2c += shx(sg2c, 2);
      if (c < KP) {
        double* oc = A.partA + ((size_t)blockIdx.x * KP + c) * DS;
        if (tq == 0) { oc[0] = sgc; oc[1] = sg2c; }
#pragma unroll
        for (int i = 0; i < 2; ++i) oc[2 + 2 * tq + i] = (2 * tq + i < d) ? fma(A.shift[2 * tq + i], sgc, accX[b][i]) : 0.0;
      }
    }
  }
  if (lane == 0) A.amb_count[nlist] = nAmb;
}

template <int TP>
static cudaError_t launch_emd_t(const Geom& g, EmArgs a, int grid, cudaStream_t st) {
  const int KP = g.K8 * 8;
  const size_t smem = emd_smem(KP, TP);
  const char* ev = getenv("MLF_B200_EMD_XMMA");   // A/B switch: 0 = sum g x on the CUDA cores (lane = component)
  const bool xmma = !(ev && ev[0] == '0');
  a.S = emd_stride(KP, TP);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 256, smem, st>>>(a);
    return cudaSuccess;
  };
  // expanded-form P1 first (returns at once unless every bound S_k is inside the limit), then the lane = component P1
  const bool qp1 = xmma && a.qf_limit >= 0.0 && a.qfS;
  if (!qp1) a.qf_limit = -1.0;
  cudaError_t e = cudaSuccess;
  if (qp1) e = go(k_emd<TP, true, true>);
  if (e == cudaSuccess) e = xmma ? go(k_emd<TP, true, false>) : go(k_emd<TP, false, false>);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

namespace {
template <bool T256>
__global__ void k_selftest_exp(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
  __shared__ double etab[kExpTabN];
  for (int i = threadIdx.x; i < (T256 ? kExpTabN : 128); i += blockDim.x) etab[i] = T256 ? kExpTab[i] : kExpTab128[i];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = T256 ? exp_tab(in[i], etab) : exp_tab128(in[i], etab);
}
}  // namespace

// which: 0 = exp_tab (256 entries, k_em16), 1 = exp_tab128 (k_emd)
cudaError_t launch_selftest_exp(const double* in, double* out, int64_t n, int which, cudaStream_t st) {
  if (which == 0) k_selftest_exp<true><<<148, 256, 0, st>>>(in, out, n);
  else k_selftest_exp<false><<<148, 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_emd(const Geom& g, const FusedGeom& f, EmArgs a, int grid, size_t smem_limit, cudaStream_t st) {
  if (!emd_ok(g, f, smem_limit)) return cudaErrorInvalidConfiguration;
  // (32-point tiles with two blocks per SM -- 128 registers -- were measured again with the expanded-form P1: 184 bytes of
  // spills, 261 ms per iteration against 200; only the 64-point instantiation is built)
  return launch_emd_t<64>(g, a, grid, st);
}

}  // namespace mlfb200
