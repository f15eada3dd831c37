// 16-warp variant of the fused single-sweep EM kernel (see emgmm_fused.cu for the algorithm and the threshold protocol).
//
// Why a second variant: the 8-warp kernel keeps the covariance-scatter accumulators (8 components x 3 blocks x 2 doubles
// = 96 registers per thread) alive across the whole tile loop, which pins it at 255 registers and two warps per
// scheduler; its softmax and statistics phases are latency-bound at that occupancy (profiles/ncu_r01*).  Here the
// scatter accumulators live in TENSOR MEMORY (TMEM, 256 KB per SM, otherwise unused by an fp64 kernel -- tcgen05 has no
// f64 MMA): each warp parks its 4 components x MB blocks in its own TMEM lane quarter and pulls one component at a time
// into registers with tcgen05.ld while it walks that component's kept groups, then puts it back with tcgen05.st.
// That frees the registers for 16 warps per SM (<= 128 registers each, four warps per scheduler):
// The 16 warps form TWO INDEPENDENT GROUPS of 8 (group = w >> 3), each sweeping its own 64-point tiles with its own logit
// tile and point tile in shared memory and its own 256-thread named barrier; the groups share only the parameter pack.
// Nothing synchronises them inside the tile loop.  (Measured: forcing the groups into antiphase with a start offset changes
// nothing -- at four warps per scheduler the sweep is bound by the fp64 datapath's throughput, not by phase latency.)
//   warp w of a group:  P1/P2 on its own 8 points of the group's 64-point tile;
//                       P3a pass-A statistics of component block (w & 7) over the tile's sixteen 4-point groups;
//                       P3b covariance scatter of the 8 components of that block over the kept groups.
// Per-block partials are written per group: slot = 2*blockIdx.x + (w >> 3).
#include "emgmm_kernels.cuh"
#include "exp_tab.cuh"

#include <cmath>
#include <cstdlib>

namespace mlfb200 {

namespace {

__host__ __device__ constexpr int g_blk_off(int mb) {
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

// ---- tensor memory as per-warp scratch: lane quarter (warp & 3), 4 consecutive 32-bit columns = two doubles ----------
__device__ __forceinline__ void tmem_st2(uint32_t taddr, double v0, double v1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(__double2loint(v0)),
               "r"(__double2hiint(v0)), "r"(__double2loint(v1)), "r"(__double2hiint(v1))
               : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, double& v0, double& v1) {
  int a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  v0 = __hiloint2double(b, a);
  v1 = __hiloint2double(d, c);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

constexpr int kThreads16 = 512;

#ifdef MLF_TRACE
// Debug builds only (make EXTRA=-DMLF_TRACE=1; scripts/trace_phases.py): phase timeline of block 0's sixteen warps over 64 tiles
__device__ long long g_trace[64 * 16 * 8];
#define MLF_STAMP(ph)                                                                                         \
  do {                                                                                                        \
    if (blockIdx.x == 0 && lane == 0 && tcount >= 8 && tcount < 72) g_trace[((tcount - 8) * 16 + warp) * 8 + (ph)] = clock64(); \
  } while (0)
#else
#define MLF_STAMP(ph) do { } while (0)
#endif
constexpr int kTmemCols = 512;  // 4 warps per lane quarter x 8 components x 16 columns (MB <= 3 blocks x 4 columns, padded)

}  // namespace

size_t em16_smem_bytes(const Geom& g, const FusedGeom& f) { return f.smem + (kExpTabN + 128) * sizeof(double) + 32; }
// quadratic-form variant: the coefficient pack [K8][qf_stride] replaces the whitening pack; plus the feature table and the
// shifted points of every warp ([16 warps x 8 points][4 DM + 4])
size_t em16_qf_smem_bytes(const Geom& g, const FusedGeom& f) {
  const size_t KP = (size_t)g.K8 * 8;
  return em16_smem_bytes(g, f) - KP * f.PSF * sizeof(double) +
         ((size_t)g.K8 * qf_stride(g.DM) + 2 * (size_t)qf_nsteps(g.DM) + 128 * (size_t)(4 * g.DM + 4)) * sizeof(double);
}
bool em16_ok(const Geom& g, const FusedGeom& f, size_t smem_limit) {
  return f.GPW == 2 && !f.diag && g.DN <= 2 && g.K8 <= 8 && f.scatter_ok && em16_smem_bytes(g, f) <= smem_limit;
}
bool em16_qf_ok(const Geom& g, const FusedGeom& f, size_t smem_limit) {
  return em16_ok(g, f, smem_limit) && em16_qf_smem_bytes(g, f) <= smem_limit;
}

// VAR (MLF_B200_EM16_VAR, default 53; every variant produces the same bits):
//   bit 0      the 1/sum of a point is kept beside the tile and applied where the responsibilities are consumed (P3) instead of
//              in a third pass over the logit row (same product);
//   bits 1-2   0: quad reduction of a component group right after its DMMAs; 1: software-pipelined -- the reduction of group
//              k4-1 is issued after the DMMAs of group k4, so its shuffles run under them; 2: the same, unrolled twice
//              (ping-pong partial sums, no register moves);
//   bits 3-4   the barrier that protects the group's tiles from being overwritten by the next tile is split (mbarrier): a warp
//              arrives when its statistics are done and waits only where it first writes shared memory -- 1: after the next
//              tile's points have been requested, 2: after the first group of DMMAs of the next tile, 3: after the second;
//   bit 5      the next tile's points are fetched into registers before the statistics phase.
// Measured at N=1e8, d=16, K=64 (ms per sweep): 0: 123.5, 1: 121.9, 2: 119.7, 4: 121.0, 5: 119.4, 13: 118.2, 21: 117.6,
// 29: 117.4 (4-byte spill), 53: 117.0, 61: 116.9 (4-byte spill).  (Later in the round, on top of variant 53: one-comparison threshold
// tests 116.1, 256-entry exp table 115.7.)  Also measured, no gain: issuing the DMMA chains of 2 or 4
// components interleaved (119.5 / 120.5: the loop is not bound by dependent-issue latency), parking the pass-A accumulators in
// spare tensor-memory columns to free registers for fragment prefetch (+2 ms), unconditional kept-sum DMMAs in P3a (+3 ms),
// skipping 4-point groups whose responsibilities are all exactly 0 (+-0), a two-stage P3a (branch-free pass-A sweep that records
// which 4-point groups hold kept / undecided pairs, then those groups only: +1.4 ms), shift and thresholds re-read from shared
// memory per tile instead of living in registers (+0.5 ms), eight instead of four exps in flight in P2 (+-0),
// requesting the next component's scatter accumulators from tensor memory while the current one is updated (dynamic component
// index instead of the unrolled loop: +1.4 ms).
// QF: logits by the expanded quadratic form (emgmm_kernels.cuh, qf_*) instead of whitening: the point's monomials are the
// register-resident A operand (two passes of <= 19 k-steps at d = 16, half-way logits parked in the logit tile), the
// coefficient fragments of 8 components the B operand, the accumulator is the logit.  Both variants are launched for a sweep;
// each block decides from the bounds S_k left by k_refresh which of the two does the work (the other returns at once), so
// the choice needs no host round trip.  Measured in isolation (bench/micro/p1_quadform.cu): 2.24 ms against 3.27 ms for the
// whitening loop on the same points; in the sweep at N=1e8, d=16, K=64: 95.2 ms against 116.9.  Scheduling experiments on this
// variant, all measured at the headline shape and none kept (QFM >> 1 selects them, not instantiated): accumulator chains of two
// component blocks interleaved 96.5 ms; eight exps in flight in P2 95.1; P3a unrolled 8 times 95.4; the barrier between softmax and
// statistics replaced by one mbarrier per producer warp, consumers starting with their own points' groups, 97.1; the two groups
// forced to take turns in P1 (P1 of one under P2 / P3 of the other, two mbarriers) 97.3, with two chains 98.6; the two groups
// forced to start every tile together (block-wide barrier) 102.8; the exps of a component block skipped when all of its 64 pairs
// flush to zero (a third of the blocks on the bench mixture; bit-identical) 96.1 against 96.0 with the test compiled in and unused.
// Why it sits at ~95 ms whatever is moved (phase timeline of a block, scripts/trace_phases.py on a -DMLF_TRACE=1 build, and
// bench/micro/p1_quadform_occ.cu): a warp issues one DMMA per ~32 clocks at best, so the datapath (one DMMA per 16 clocks per
// scheduler) needs >= 3 warps of a scheduler in P1 to saturate -- 0.88 of the peak with 4, 0.86 with 3, 0.75-0.80 with 2, 0.44-0.50 with
// 1, prefetching or interleaving chains notwithstanding.  The two groups drift into near-synchrony: all four warps of a scheduler
// are in P1 35 % of the time, two 29 %, none 22 %.  Taking turns removes the "none" state but leaves two warps in P1 all the time
// (0.75) and doubles P3a's duration under the contention.  A third group would fix it; tensor memory (2 x 98 KB of scatter
// accumulators) and shared memory (200 KB) are both full with two.
template <int DM, int VAR, int QFM, bool STORE, bool SPLT>
__global__ void __launch_bounds__(kThreads16, 1) k_em16(EmArgs A) {
  // STORE: "statistics first" pass of an iteration whose covariance thresholds cannot be predicted: E-step + pass-A statistics
  // only (no classification, no scatter), and the tile of finished responsibilities goes to HBM ([tile][64][KP], A.gtiles) for
  // the scatter pass k_scatter16, which then knows the exact thresholds.
  constexpr bool QF = (QFM & 1) != 0;   // QFM: 0 whitening, odd: quadratic form; QFM >> 1 = experiment (MLF_B200_QF_EXP, A/B runs)
  constexpr int CH = 1;                // interleaved accumulator chains per warp in P1 (2 was measured: see above)
  constexpr bool FOLD = (VAR & 1) != 0;
  static_assert(!QF || FOLD, "the quadratic-form variant keeps 1/sum beside the tile");
  constexpr int PIPE = (VAR >> 1) & 3;
  constexpr int SPLIT = (VAR >> 3) & 3;   // trailing group barrier as a split mbarrier: 1 wait after the X loads are issued, 2 after the first DMMA group, 3 after the second
  constexpr bool PREF = (VAR & 32) != 0;  // the next tile's points are fetched into registers before the statistics phase
  constexpr int DN = (DM + 1) / 2;
  constexpr int NB = g_blk_off(DM);
  constexpr int OFFB = NB * 32, OFFC = OFFB + 8 * DN, PSF = OFFC + 4;
  constexpr int XS = 8 * DN + 4;
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int DS = 8 * DN + 2;
  constexpr int MSF = MB * 64 + 8 * DN + 2;
  constexpr int TP = 64;   // points per group tile
  constexpr int D4 = 4 * DM, NF = qf_nsteps(DM), QS = qf_stride(DM), XQ = D4 + 4;
  constexpr int NPASS = NF > 20 ? 2 : 1, NH = (NF + NPASS - 1) / NPASS;   // k-steps of the first pass (registers: 2 per k-step)
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, K4 = 2 * K8, S = A.S, d = A.d;
  const int64_t N = A.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  {  // which variant sweeps: every S_k inside the limit -> quadratic form, otherwise (or NaN, or no limit) whitening
    bool bad = !(A.qf_limit >= 0.0) || A.qfS == nullptr;
    if (!bad)
      for (int i = tid; i < KP; i += kThreads16) bad |= !(A.qfS[i] <= A.qf_limit);
    const bool use_qf = __syncthreads_or(bad ? 1 : 0) == 0;
    if (use_qf != QF) return;
  }
  double* sp = smem;                                  // [KP][PSF] whitening pack / [K8][QS] quadratic-form pack
  double* smu = sp + (QF ? (size_t)K8 * QS : (size_t)KP * PSF);   // [KP][8 DN]
  const int half = warp >> 3;                   // group index
  // SPLT: fewer than 8 component blocks, the statistics of a block are shared by A.spl = 2, 4 or 8 warps of the group (a template
  // flag, not a run-time value: the headline shape has no register to spare for it)
  const int SPL = SPLT ? A.spl : 1;
  const int cb = (warp & 7) % (8 / SPL), spw = (warp & 7) / (8 / SPL);   // component block and share of the tile's 4-point groups
  const int wg = warp & 7;                      // warp index inside the group
  double* xs = smu + (size_t)KP * 8 * DN + (size_t)half * TP * XS;                    // [2][64][XS] raw points, per group
  double* tile = smu + (size_t)KP * 8 * DN + (size_t)2 * TP * XS + (size_t)half * TP * S;  // [2][64][S] logits -> responsibilities
  double* etab = smu + (size_t)KP * 8 * DN + (size_t)2 * TP * XS + (size_t)2 * TP * S;   // [kExpTabN] 2^(j/256) (exp_tab.cuh)
  double* invs = etab + kExpTabN + half * TP;                                                 // [2][64] 1/sum per point (FOLD)
  uint32_t* s_taddr = reinterpret_cast<uint32_t*>(etab + kExpTabN + 2 * TP);
  const bool own = cb < K8;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(etab + kExpTabN + 2 * TP + 2 + half);   // one 8-byte mbarrier per group
  if (SPLIT && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;\n" ::"r"(mbar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;\n" ::"r"(mbar + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < kExpTabN) etab[tid] = kExpTab[tid];
  // QF: feature table (byte offsets of the two factors inside a staged point) and this warp's 8 staged points x' = x - shift
  uint32_t* ftab = reinterpret_cast<uint32_t*>(etab + kExpTabN + 2 * TP + 4);                   // [4 NF]
  double* xq = etab + kExpTabN + 2 * TP + 4 + 2 * NF + (size_t)(warp * 8 + g) * XQ;              // [16 x 8][XQ]
  if (QF) {
    for (int f = tid; f < 4 * NF; f += kThreads16) {
      int i = 0, j = 0;
      qf_feature(D4, f, i, j);
      ftab[f] = (uint32_t)(8 * i) | ((uint32_t)(8 * j) << 16);
    }
    if (t == 0) { xq[D4] = 1.0; xq[D4 + 1] = 0.0; }
    for (int i = tid; i < K8 * QS; i += kThreads16) sp[i] = A.packq[i];
  } else {
    for (int i = tid; i < KP * PSF; i += kThreads16) sp[i] = A.pack2[i];
  }
  for (int i = tid; i < KP * 8 * DN; i += kThreads16) smu[i] = A.muold[i];
  double sh[DM];
  // VEC (quadratic-form variant at d = 16): lane t of a point's quad loads coordinates 4t .. 4t+3 as two 16-byte loads instead of
  // 4t', 4t'+4, ... as four 8-byte ones -- the monomials are formed from shared memory, so any register layout will do; the
  // whitening variant needs x[4 mb + t] as its A fragments and keeps the scalar loads
  const bool vec = QF && DM == 4 && d == 16 && (reinterpret_cast<uintptr_t>(A.X) & 15) == 0;
#pragma unroll
  for (int mb = 0; mb < DM; ++mb) sh[mb] = vec ? A.shift[4 * t + mb] : A.shift[4 * mb + t];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(s_taddr)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // this warp's TMEM region: lane quarter (warp & 3), 128 columns starting at 128*(warp >> 2); component q at +16*q, block b at +4*b
  const uint32_t tbase = *s_taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));

  const bool accu = A.accumulate != 0;
  const int slot = (2 * blockIdx.x + half) * SPL + spw;
  double* outA = A.partA + (size_t)slot * KP * DS;
  double* outB = A.partB + (size_t)slot * KP * MSF;
  double accA[DN][2], accK[DN][2], sg, sg2, sk, thi_r, tlo_r, nk0;
  int nk = 0;   // pairs of this lane's component that passed the certainly-kept test (profiling counter, pad slot of MSF)
  {
    const bool ld = accu && own;
    const double* o = outA + (size_t)(8 * cb + g) * DS;
    const double* ok = outB + (size_t)(8 * cb + g) * MSF + MB * 64;
    sg = (ld && t == 0) ? o[0] : 0.0;
    sg2 = (ld && t == 0) ? o[1] : 0.0;
    sk = (ld && t == 0) ? ok[8 * DN] : 0.0;
    nk0 = (ld && t == 0) ? ok[8 * DN + 1] : 0.0;
    // "not below the threshold and not zero" as one comparison: responsibilities are >= 0 (or NaN), so
    // !(a < t) && a != 0  <=>  !(a < max(t, smallest subnormal))
    thi_r = own ? fmax(A.thi[8 * cb + g], 4.9406564584124654e-324) : INFINITY;
    tlo_r = own ? fmax(A.tlo[8 * cb + g], 4.9406564584124654e-324) : INFINITY;
#pragma unroll
    for (int b = 0; b < DN; ++b) {
      accA[b][0] = ld ? o[2 + 8 * b + 2 * t] : 0.0;
      accA[b][1] = ld ? o[2 + 8 * b + 2 * t + 1] : 0.0;
      accK[b][0] = ld ? ok[8 * b + 2 * t] : 0.0;
      accK[b][1] = ld ? ok[8 * b + 2 * t + 1] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double* oq = outB + (size_t)(8 * cb + q) * MSF;
#pragma unroll
      for (int b = 0; b < MB; ++b)
        tmem_st2(tbase + 16 * q + 4 * b, ld ? oq[b * 64 + g * 8 + 2 * t] : 0.0, ld ? oq[b * 64 + g * 8 + 2 * t + 1] : 0.0);
    }
    tmem_wait_st();
  }
  const int nlist = blockIdx.x * 16 + warp;
  int nAmb = accu ? A.amb_count[nlist] : 0;
  AmbEntry* ambw = A.amb + (size_t)nlist * A.amb_cap;

  const int64_t ntiles = (N + TP - 1) / TP;
  bool first = true;
  uint32_t phase = 0;
  double xraw[DM];
  int tcount = -1;
  for (int64_t bt = 2 * (int64_t)blockIdx.x + half; bt < ntiles; bt += 2 * (int64_t)gridDim.x) {
    const int64_t p = bt * TP + wg * 8 + g;
    ++tcount;
    MLF_STAMP(0);
    // ---- P1: logits of this warp's 8 points against every component ---------------------------------------------
    double xa[DM];
    double* xr = xs + (size_t)(wg * 8 + g) * XS;
    auto load_x = [&](const int64_t pp) {
      if (vec) {
        if constexpr (DM == 4) {
          double2 v0 = make_double2(0.0, 0.0), v1 = v0;
          if (pp < N) {
            const double2* src = reinterpret_cast<const double2*>(A.X + pp * 16 + 4 * t);
            v0 = src[0]; v1 = src[1];
          }
          xraw[0] = v0.x; xraw[1] = v0.y; xraw[2] = v1.x; xraw[3] = v1.y;
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < DM; ++mb) {
          const int a = 4 * mb + t;
          xraw[mb] = (pp < N && a < d) ? A.X[pp * d + a] : 0.0;
        }
      }
    };
    if (!PREF || first) load_x(p);
    // the previous tile's buffers may be overwritten only after every warp of the group is done with them
    auto wait_prev = [&]() {
      if (SPLIT && !first) {
        uint32_t done;
        do {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n"
                       : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
        } while (!done);
        phase ^= 1u;
      }
    };
    auto stage_x = [&]() {
      if (vec) {
        if constexpr (DM == 4) {
          *reinterpret_cast<double2*>(xr + 4 * t) = make_double2(xraw[0], xraw[1]);
          *reinterpret_cast<double2*>(xr + 4 * t + 2) = make_double2(xraw[2], xraw[3]);
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < DM; ++mb) xr[4 * mb + t] = xraw[mb];
      }
      if (8 * DN > 4 * DM) xr[4 * DM + t] = 0.0;
    };
    if (SPLIT < 2) { wait_prev(); stage_x(); }
    double* row = tile + (size_t)(wg * 8 + g) * S;
    double rmax = -INFINITY;
    if constexpr (QF) {
      // ---- P1 (quadratic form): monomials of the warp's 8 points x coefficient fragments of 8 components per DMMA ---------
      __syncwarp();   // every lane is done with the previous tile's staged points
      if (vec) {
        if constexpr (DM == 4) {
          *reinterpret_cast<double2*>(xq + 4 * t) = make_double2(xraw[0] - sh[0], xraw[1] - sh[1]);
          *reinterpret_cast<double2*>(xq + 4 * t + 2) = make_double2(xraw[2] - sh[2], xraw[3] - sh[3]);
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < DM; ++mb) xq[4 * mb + t] = xraw[mb] - sh[mb];
      }
      __syncwarp();
      const char* xqb = reinterpret_cast<const char*>(xq);
#pragma unroll
      for (int pass = 0; pass < NPASS; ++pass) {
        const int s0 = pass * NH, ns = pass == 0 ? NH : NF - NH;
        double f[NH];
#pragma unroll
        for (int s = 0; s < NH; ++s)
          if (s < ns) {
            const uint32_t w = ftab[4 * (s0 + s) + t];
            f[s] = *reinterpret_cast<const double*>(xqb + (w & 0xffffu)) * *reinterpret_cast<const double*>(xqb + (w >> 16));
          }
        for (int cq = 0; cq < K8; cq += CH) {
          const double* q[CH];
          double c[CH][2];
          bool live[CH];
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            live[u] = cq + u < K8;   // odd K8: the last pair's second chain recomputes the last block and is not stored
            const int cu = live[u] ? cq + u : K8 - 1;
            q[u] = sp + (size_t)cu * QS;
            const double2 v = pass == 0 ? *reinterpret_cast<const double2*>(q[u] + NF * 32 + 2 * t) : *reinterpret_cast<const double2*>(row + 8 * cu + 2 * t);
            c[u][0] = v.x; c[u][1] = v.y;
          }
#pragma unroll
          for (int s = 0; s < NH; ++s)
            if (s < ns) {
#pragma unroll
              for (int u = 0; u < CH; ++u) dmma(c[u], f[s], q[u][(s0 + s) * 32 + lane]);
            }
          if (SPLIT >= 2 && pass == 0 && cq == 0) { wait_prev(); stage_x(); }
#pragma unroll
          for (int u = 0; u < CH; ++u)
            if (live[u]) {
              *reinterpret_cast<double2*>(row + 8 * (cq + u) + 2 * t) = make_double2(c[u][0], c[u][1]);
              if (pass == NPASS - 1) rmax = fmax(rmax, fmax(c[u][0], c[u][1]));
            }
        }
      }
    } else {
#pragma unroll
    for (int mb = 0; mb < DM; ++mb) xa[mb] = xraw[mb] - sh[mb];
    // |u|^2 of the four components of group k4 for this thread's output columns
    auto comp4 = [&](int k4, double (&part)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double* pk = sp + (size_t)(4 * k4 + j) * PSF;
        double c[DN][2];
#pragma unroll
        for (int nb = 0; nb < DN; ++nb) {
          const double2 v = *reinterpret_cast<const double2*>(pk + OFFB + 8 * nb + 2 * t);
          c[nb][0] = v.x; c[nb][1] = v.y;
        }
#pragma unroll
        for (int mb = 0; mb < DM; ++mb)
#pragma unroll
          for (int nb = 0; nb <= mb / 2; ++nb) dmma(c[nb], xa[mb], pk[(g_blk_off(mb) + nb) * 32 + lane]);
        double s = c[0][0] * c[0][0];
        s = fma(c[0][1], c[0][1], s);
#pragma unroll
        for (int nb = 1; nb < DN; ++nb) { s = fma(c[nb][0], c[nb][0], s); s = fma(c[nb][1], c[nb][1], s); }
        part[j] = s;
      }
    };
    // transposed reduction over the quad: lane t ends with the full |u|^2 of component 4*k4+t
    auto reduce4 = [&](int k4, const double (&part)[4]) {
      const double ck = sp[(size_t)(4 * k4 + t) * PSF + OFFC];
      const bool hi2 = (t & 2) != 0, hi1 = (t & 1) != 0;
      const double send0 = hi2 ? part[0] : part[2], send1 = hi2 ? part[1] : part[3];
      const double keep0 = hi2 ? part[2] : part[0], keep1 = hi2 ? part[3] : part[1];
      const double a0 = keep0 + shx(send0, 2), a1 = keep1 + shx(send1, 2);
      const double maha = (hi1 ? a1 : a0) + shx(hi1 ? a0 : a1, 1);
      const double logit = fma(-0.5, maha, ck);
      row[4 * k4 + t] = logit;
      rmax = fmax(rmax, logit);
    };
    if (PIPE == 0) {
      for (int k4 = 0; k4 < K4; ++k4) {
        double part[4];
        comp4(k4, part);
        reduce4(k4, part);
      }
    } else if (PIPE == 1) {
      double prev[4];
      comp4(0, prev);
      for (int k4 = 1; k4 < K4; ++k4) {
        double part[4];
        comp4(k4, part);
        reduce4(k4 - 1, prev);
#pragma unroll
        for (int j = 0; j < 4; ++j) prev[j] = part[j];
      }
      reduce4(K4 - 1, prev);
    } else if (SPLIT < 3) {   // K4 = 2 K8 is even
      double pa[4], pb[4];
      comp4(0, pa);
      if (SPLIT == 2) { wait_prev(); stage_x(); }
      for (int k4 = 1; k4 + 1 < K4; k4 += 2) {
        comp4(k4, pb);
        reduce4(k4 - 1, pa);
        comp4(k4 + 1, pa);
        reduce4(k4, pb);
      }
      comp4(K4 - 1, pb);
      reduce4(K4 - 2, pa);
      reduce4(K4 - 1, pb);
    } else {
      double pa[4], pb[4];
      comp4(0, pa);
      comp4(1, pb);
      wait_prev();
      stage_x();
      reduce4(0, pa);
      for (int k4 = 2; k4 + 1 < K4; k4 += 2) {
        comp4(k4, pa);
        reduce4(k4 - 1, pb);
        comp4(k4 + 1, pb);
        reduce4(k4, pa);
      }
      reduce4(K4 - 1, pb);
    }
    }   // !QF
    MLF_STAMP(1);
    // ---- P2: log-sum-exp over the components ---------------------------------------------------------------------
    rmax = fmax(rmax, shx(rmax, 1));
    rmax = fmax(rmax, shx(rmax, 2));
    double sum = 0;
    if constexpr (QF) {   // each lane exponentiates the two logits per component block it wrote itself
#pragma unroll 2
      for (int cq = 0; cq < K8; ++cq) {
        const double2 v = *reinterpret_cast<const double2*>(row + 8 * cq + 2 * t);
        const double e0 = exp_tab(v.x - rmax, etab), e1 = exp_tab(v.y - rmax, etab);
        *reinterpret_cast<double2*>(row + 8 * cq + 2 * t) = make_double2(e0, e1);
        sum += e0; sum += e1;
      }
    } else {
#pragma unroll 2
    for (int k4 = 0; k4 < K4; k4 += 2) {
      const double e0 = exp_tab(row[4 * k4 + t] - rmax, etab), e1 = exp_tab(row[4 * k4 + 4 + t] - rmax, etab);
      row[4 * k4 + t] = e0; row[4 * k4 + 4 + t] = e1;
      sum += e0; sum += e1;
    }
    }
    sum += shx(sum, 1);
    sum += shx(sum, 2);
    const double inv = (p < N) ? 1.0 / sum : 0.0;
    if (FOLD) {
      if (t == 0) invs[wg * 8 + g] = inv;
    } else {
      for (int k4 = 0; k4 < K4; ++k4) row[4 * k4 + t] *= inv;
    }
    if constexpr (STORE) {   // this warp's 8 rows, normalised, as one coalesced 16-byte store per lane and row
      __syncwarp();
      double* go = A.gtiles + ((size_t)bt * TP + wg * 8) * KP;
      if (2 * lane < KP) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const double2 v = *reinterpret_cast<const double2*>(tile + (size_t)(wg * 8 + r) * S + 2 * lane);
          const double iv = invs[wg * 8 + r];
          *reinterpret_cast<double2*>(go + (size_t)r * KP + 2 * lane) = make_double2(v.x * iv, v.y * iv);
        }
      }
    }
    MLF_STAMP(2);
    // the group's responsibilities and points are complete: 256-thread named barrier of this group only
    asm volatile("bar.sync %0, 256;\n" ::"r"(1 + half) : "memory");
    MLF_STAMP(3);
    if (PREF) load_x(p + 2 * (int64_t)gridDim.x * TP);
    unsigned act = 0;   // lane pb: bit q set <=> component 8*cb+q keeps some point of 4-point group pb
    // ---- P3a: pass-A statistics of component block cb over the tile's 4-point groups; kept-group flags ---------------
    if (own) {
      const int64_t q0 = bt * TP;
      const double* tcol = tile + 8 * cb + g;
      // this warp's share of the tile's sixteen 4-point groups (all of them without the split: compile-time bounds)
      const int i0 = SPLT ? spw * (16 / SPL) : 0, i1 = SPLT ? i0 + 16 / SPL : 16;
#pragma unroll 4
      for (int i = i0; i < i1; ++i) {
        const int pb = i;
        const double a = FOLD ? tcol[(size_t)(4 * pb + t) * S] * invs[4 * pb + t] : tcol[(size_t)(4 * pb + t) * S];
        double b[DN];
#pragma unroll
        for (int db = 0; db < DN; ++db) b[db] = xs[(size_t)(4 * pb + t) * XS + 8 * db + g];
        sg += a;
        sg2 = fma(a, a, sg2);
#pragma unroll
        for (int db = 0; db < DN; ++db) dmma(accA[db], a, b[db]);
        if constexpr (STORE) continue;
        // "If(W(i) < minW0) CYCLE" (src/mlf_matrix.f90:135); g == 0 adds nothing and is never listed
        const bool hi = !(a < thi_r);
        const bool mid = !hi && !(a < tlo_r);
        const double ahi = hi ? a : 0.0;
        sk += ahi;
        nk += hi ? 1 : 0;
        const unsigned mh = __ballot_sync(kFull, hi);
        const unsigned mm = __ballot_sync(kFull, mid);
        if (mh) {
#pragma unroll
          for (int db = 0; db < DN; ++db) dmma(accK[db], ahi, b[db]);
          unsigned cm = mh | (mh >> 1);
          cm |= cm >> 2;
          cm &= 0x11111111u;
          cm = (cm | (cm >> 3)) & 0x03030303u;
          cm = (cm | (cm >> 6)) & 0x000f000fu;
          cm = (cm | (cm >> 12)) & 0xffu;
          if (lane == pb) act = cm;
        }
        if (mm) {
          const int sl = nAmb + __popc(mm & ((1u << lane) - 1u));
          if (mid && sl < A.amb_cap) {
            AmbEntry e;
            e.g = a;
            e.key = (long long)(A.p_offset + q0 + 4 * pb + t) * 1024 + (8 * cb + g);
            ambw[sl] = e;
          }
          nAmb += __popc(mm);
        }
      }
    }
    MLF_STAMP(4);
    // ---- P3b: covariance scatter of this block's 8 components over the kept groups of the tile ----------------------
    if (own && !STORE) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        unsigned pm = __ballot_sync(kFull, (act >> q) & 1u);
        if (pm == 0) continue;
        const double thq = __shfl_sync(kFull, thi_r, 4 * q);
        const double* mq = smu + (size_t)(8 * cb + q) * 8 * DN;
        double mu_q[DN];
#pragma unroll
        for (int ab = 0; ab < DN; ++ab) mu_q[ab] = mq[8 * ab + g];
        const double* tq = tile + 8 * cb + q;
        double acc[MB][2];
#pragma unroll
        for (int b = 0; b < MB; ++b) tmem_ld2(tbase + 16 * q + 4 * b, acc[b][0], acc[b][1]);
        while (pm) {
          const int pb = __ffs(pm) - 1;
          pm &= pm - 1;
          const double aq = FOLD ? tq[(size_t)(4 * pb + t) * S] * invs[4 * pb + t] : tq[(size_t)(4 * pb + t) * S];
          const double w = !(aq < thq) ? aq : 0.0;
          double z[DN], wz[DN];
#pragma unroll
          for (int ab = 0; ab < DN; ++ab) {
            z[ab] = xs[(size_t)(4 * pb + t) * XS + 8 * ab + g] - mu_q[ab];
            wz[ab] = w * z[ab];
          }
#pragma unroll
          for (int ab = 0; ab < DN; ++ab)
#pragma unroll
            for (int bb = 0; bb <= ab; ++bb) dmma(acc[ab * (ab + 1) / 2 + bb], wz[ab], z[bb]);
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) tmem_st2(tbase + 16 * q + 4 * b, acc[b][0], acc[b][1]);
      }
      tmem_wait_st();
    }
    MLF_STAMP(5);
    // the group's tiles may be overwritten only after every warp of the group is done with them
    if (SPLIT) {
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar) : "memory");
    } else {
      asm volatile("bar.sync %0, 256;\n" ::"r"(1 + half) : "memory");
    }
    first = false;
  }
  // ---- per-(block, half) partials -------------------------------------------------------------------------------------
  if (own) {
    double a = sg, b = sg2, c = sk;
    nk += __shfl_xor_sync(kFull, nk, 1); nk += __shfl_xor_sync(kFull, nk, 2);
    a += shx(a, 1); a += shx(a, 2);
    b += shx(b, 1); b += shx(b, 2);
    c += shx(c, 1); c += shx(c, 2);
    double* o = outA + (size_t)(8 * cb + g) * DS;
    double* ok = outB + (size_t)(8 * cb + g) * MSF + MB * 64;
    if (t == 0) { o[0] = a; o[1] = b; ok[8 * DN] = c; ok[8 * DN + 1] = nk0 + (double)nk; }
#pragma unroll
    for (int db = 0; db < DN; ++db) {
      o[2 + 8 * db + 2 * t] = accA[db][0];
      o[2 + 8 * db + 2 * t + 1] = accA[db][1];
      ok[8 * db + 2 * t] = accK[db][0];
      ok[8 * db + 2 * t + 1] = accK[db][1];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double* oq = outB + (size_t)(8 * cb + q) * MSF;
#pragma unroll
      for (int bq = 0; bq < MB; ++bq) {
        double v0, v1;
        tmem_ld2(tbase + 16 * q + 4 * bq, v0, v1);
        oq[bq * 64 + g * 8 + 2 * t] = v0;
        oq[bq * 64 + g * 8 + 2 * t + 1] = v1;
      }
    }
  }
  if (lane == 0) A.amb_count[nlist] = nAmb;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(*s_taddr), "n"(kTmemCols) : "memory");
  }
}

// Scatter pass of a "statistics first" iteration (see STORE above): the responsibilities of the sweep's tiles come back from HBM
// (cp.async, two tile buffers per group: tile i+1 lands while tile i is scattered), the thresholds are exact, nothing is listed.
// Same tile -> (block, group) assignment, warp -> component block assignment and instruction order of the accumulations as
// P3 of k_em16, so the partials equal those of a repeated full sweep bit for bit.  HBM-bound when few pairs are kept
// (8 (KP + d) bytes per point), DMMA-bound when most are (the dense d x d scatter of the reference's dsyrk).
template <int DM>
__global__ void __launch_bounds__(kThreads16, 1) k_scatter16(EmArgs A) {
  constexpr int DN = (DM + 1) / 2;
  constexpr int XS = 8 * DN + 4;
  constexpr int MB = DN * (DN + 1) / 2;
  constexpr int MSF = MB * 64 + 8 * DN + 2;
  constexpr int TP = 64;
  extern __shared__ __align__(16) double smem[];
  const int K8 = A.K8, KP = K8 * 8, S = A.S, d = A.d;
  const int64_t N = A.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int half = warp >> 3, wg = warp & 7;
  const int SPL = A.spl > 1 ? A.spl : 1;
  const int cb = wg % (8 / SPL), spw = wg / (8 / SPL);
  double* smu = smem;                                                     // [KP][8 DN]
  double* xsb = smu + (size_t)KP * 8 * DN + (size_t)half * 2 * TP * XS;   // [2 groups][2 buffers][64][XS]
  double* tlb = smu + (size_t)KP * 8 * DN + (size_t)4 * TP * XS + (size_t)half * 2 * TP * S;   // [2 groups][2 buffers][64][S]
  uint32_t* s_taddr = reinterpret_cast<uint32_t*>(smu + (size_t)KP * 8 * DN + (size_t)4 * TP * XS + (size_t)4 * TP * S);
  const bool own = cb < K8;
  for (int i = tid; i < KP * 8 * DN; i += kThreads16) smu[i] = A.muold[i];
  // padding of the point tiles (coordinates d .. 8 DN - 1) is never written by the copies: zero it once
  for (int i = tid; i < 4 * TP * XS; i += kThreads16) smu[(size_t)KP * 8 * DN + i] = 0.0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(s_taddr)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = *s_taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));
  const int slot = (2 * blockIdx.x + half) * SPL + spw;
  double* outB = A.partB + (size_t)slot * KP * MSF;
  double accK[DN][2], sk = 0.0;
  int nk = 0;
  const double thi_r = own ? fmax(A.thi[8 * cb + g], 4.9406564584124654e-324) : INFINITY;
#pragma unroll
  for (int b = 0; b < DN; ++b) { accK[b][0] = 0.0; accK[b][1] = 0.0; }
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int b = 0; b < MB; ++b) tmem_st2(tbase + 16 * q + 4 * b, 0.0, 0.0);
  tmem_wait_st();

  const int64_t ntiles = (N + TP - 1) / TP;
  const int64_t bt0 = 2 * (int64_t)blockIdx.x + half, btstep = 2 * (int64_t)gridDim.x;
  // this warp copies its own 8 rows of the tile: responsibilities (KP doubles, 16-byte pieces) and points (d doubles, 8-byte pieces;
  // rows past N are zero-filled by a zero source size)
  auto fetch = [&](int64_t bt, int buf) {
    const double* gsrc = A.gtiles + ((size_t)bt * TP + wg * 8) * KP;
    double* tdst = tlb + (size_t)buf * TP * S + (size_t)(wg * 8) * S;
    for (int i = lane; i < 8 * (KP / 2); i += 32) {
      const int r = i / (KP / 2), c = i - r * (KP / 2);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tdst + (size_t)r * S + 2 * c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc + (size_t)r * KP + 2 * c) : "memory");
    }
    double* xdst = xsb + (size_t)buf * TP * XS + (size_t)(wg * 8) * XS;
    const int64_t p0 = bt * TP + wg * 8;
    for (int i = lane; i < 8 * d; i += 32) {
      const int r = i / d, a = i - r * d;
      const bool in = p0 + r < N;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xdst + (size_t)r * XS + a);
      const double* src = A.X + (in ? (p0 + r) * d + a : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(in ? 8 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  if (bt0 < ntiles) fetch(bt0, 0);
  int buf = 0;
  for (int64_t bt = bt0; bt < ntiles; bt += btstep, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    // tile bt is complete for every warp of the group, and every warp is done with the other buffer (tile bt - step)
    asm volatile("bar.sync %0, 256;\n" ::"r"(1 + half) : "memory");
    if (bt + btstep < ntiles) fetch(bt + btstep, buf ^ 1);
    const double* tile = tlb + (size_t)buf * TP * S;
    const double* xs = xsb + (size_t)buf * TP * XS;
    unsigned act = 0;
    if (own) {
      const double* tcol = tile + 8 * cb + g;
      const int i0 = spw * (16 / SPL), i1 = i0 + 16 / SPL;
#pragma unroll 4
      for (int i = i0; i < i1; ++i) {
        const int pb = i;
        const double a = tcol[(size_t)(4 * pb + t) * S];
        const bool hi = !(a < thi_r);
        const double ahi = hi ? a : 0.0;
        sk += ahi;
        nk += hi ? 1 : 0;
        const unsigned mh = __ballot_sync(kFull, hi);
        if (mh) {
          double b[DN];
#pragma unroll
          for (int db = 0; db < DN; ++db) b[db] = xs[(size_t)(4 * pb + t) * XS + 8 * db + g];
#pragma unroll
          for (int db = 0; db < DN; ++db) dmma(accK[db], ahi, b[db]);
          unsigned cm = mh | (mh >> 1);
          cm |= cm >> 2;
          cm &= 0x11111111u;
          cm = (cm | (cm >> 3)) & 0x03030303u;
          cm = (cm | (cm >> 6)) & 0x000f000fu;
          cm = (cm | (cm >> 12)) & 0xffu;
          if (lane == i) act = cm;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        unsigned pm = __ballot_sync(kFull, (act >> q) & 1u);
        if (pm == 0) continue;
        const double thq = __shfl_sync(kFull, thi_r, 4 * q);
        const double* mq = smu + (size_t)(8 * cb + q) * 8 * DN;
        double mu_q[DN];
#pragma unroll
        for (int ab = 0; ab < DN; ++ab) mu_q[ab] = mq[8 * ab + g];
        const double* tq = tile + 8 * cb + q;
        double acc[MB][2];
#pragma unroll
        for (int b = 0; b < MB; ++b) tmem_ld2(tbase + 16 * q + 4 * b, acc[b][0], acc[b][1]);
        while (pm) {
          const int pb = __ffs(pm) - 1;
          pm &= pm - 1;
          const double aq = tq[(size_t)(4 * pb + t) * S];
          const double w = !(aq < thq) ? aq : 0.0;
          double z[DN], wz[DN];
#pragma unroll
          for (int ab = 0; ab < DN; ++ab) {
            z[ab] = xs[(size_t)(4 * pb + t) * XS + 8 * ab + g] - mu_q[ab];
            wz[ab] = w * z[ab];
          }
#pragma unroll
          for (int ab = 0; ab < DN; ++ab)
#pragma unroll
            for (int bb = 0; bb <= ab; ++bb) dmma(acc[ab * (ab + 1) / 2 + bb], wz[ab], z[bb]);
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) tmem_st2(tbase + 16 * q + 4 * b, acc[b][0], acc[b][1]);
      }
      tmem_wait_st();
    }
  }
  if (own) {
    double c = sk;
    nk += __shfl_xor_sync(kFull, nk, 1); nk += __shfl_xor_sync(kFull, nk, 2);
    c += shx(c, 1); c += shx(c, 2);
    double* ok = outB + (size_t)(8 * cb + g) * MSF + MB * 64;
    if (t == 0) { ok[8 * DN] = c; ok[8 * DN + 1] = (double)nk; }
#pragma unroll
    for (int db = 0; db < DN; ++db) {
      ok[8 * db + 2 * t] = accK[db][0];
      ok[8 * db + 2 * t + 1] = accK[db][1];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double* oq = outB + (size_t)(8 * cb + q) * MSF;
#pragma unroll
      for (int bq = 0; bq < MB; ++bq) {
        double v0, v1;
        tmem_ld2(tbase + 16 * q + 4 * bq, v0, v1);
        oq[bq * 64 + g * 8 + 2 * t] = v0;
        oq[bq * 64 + g * 8 + 2 * t + 1] = v1;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(*s_taddr), "n"(kTmemCols) : "memory");
  }
}

size_t scatter16_smem_bytes(const Geom& g, const FusedGeom& f) {
  return ((size_t)g.K8 * 8 * 8 * g.DN + 4 * 64 * (size_t)f.XS + 4 * 64 * (size_t)g.S + 2) * sizeof(double);
}
cudaError_t launch_scatter16(const Geom& g, const FusedGeom& f, const EmArgs& a, int grid, size_t smem_limit, cudaStream_t st) {
  if (!em16_ok(g, f, smem_limit) || !a.gtiles) return cudaErrorInvalidConfiguration;
  const size_t smem = scatter16_smem_bytes(g, f);
  if (smem > smem_limit) return cudaErrorInvalidConfiguration;
#define MLF_S16(DMv)                                                                                              \
  case DMv: {                                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(k_scatter16<DMv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                               \
    k_scatter16<DMv><<<grid, kThreads16, smem, st>>>(a);                                                          \
  } break;
  switch (g.DM) {
    MLF_S16(1) MLF_S16(2) MLF_S16(3) MLF_S16(4)
    default: return cudaErrorInvalidConfiguration;
  }
#undef MLF_S16
  return cudaGetLastError();
}

cudaError_t launch_em16(const Geom& g, const FusedGeom& f, const EmArgs& a, int grid, size_t smem_limit, cudaStream_t st, bool store) {
  if (!em16_ok(g, f, smem_limit)) return cudaErrorInvalidConfiguration;
  if (store && !a.gtiles) return cudaErrorInvalidConfiguration;
  const size_t smem = em16_smem_bytes(g, f);
  const char* ev = getenv("MLF_B200_EM16_VAR");   // A/B switch, read per launch (tests flip it inside one process)
  const int var = ev ? atoi(ev) : 53;
  // quadratic-form variant first (it returns at once unless every S_k is inside the limit), then the whitening variant
  // (which returns at once in exactly that case)
  const bool qf = a.qf_limit >= 0.0 && a.packq && a.qfS && em16_qf_ok(g, f, smem_limit);
  const size_t smem_q = em16_qf_smem_bytes(g, f);
  EmArgs aw = a;
  if (!qf) { aw.qf_limit = -1.0; }
  const bool split = a.spl > 1;   // (the whitening scheduling variants MLF_B200_EM16_VAR != 53 exist without the split only)
  auto go = [&](auto kern, size_t sm) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads16, sm, st>>>(aw);
    return cudaSuccess;
  };
  cudaError_t e = cudaSuccess;
#define MLF_E16V(DMv, V)                                                              \
  case 64 * DMv + V:                                                                   \
    if (store && split) {                                                              \
      if (qf) e = go(k_em16<DMv, 53, 1, true, true>, smem_q);                          \
      if (e == cudaSuccess) e = go(k_em16<DMv, 53, 0, true, true>, smem);              \
    } else if (store) {                                                                \
      if (qf) e = go(k_em16<DMv, 53, 1, true, false>, smem_q);                         \
      if (e == cudaSuccess) e = go(k_em16<DMv, 53, 0, true, false>, smem);             \
    } else if (split) {                                                                \
      if (qf) e = go(k_em16<DMv, 53, 1, false, true>, smem_q);                         \
      if (e == cudaSuccess) e = go(k_em16<DMv, 53, 0, false, true>, smem);             \
    } else {                                                                           \
      if (qf) e = go(k_em16<DMv, 53, 1, false, false>, smem_q);                        \
      if (e == cudaSuccess) e = go(k_em16<DMv, V, 0, false, false>, smem);             \
    }                                                                                  \
    break;
#define MLF_E16(DMv) MLF_E16V(DMv, 0) MLF_E16V(DMv, 5) MLF_E16V(DMv, 21) MLF_E16V(DMv, 53)
  switch (64 * g.DM + (var & 63)) {
    MLF_E16(1) MLF_E16(2) MLF_E16(3) MLF_E16(4)
    default: return cudaErrorInvalidConfiguration;
  }
  if (e != cudaSuccess) return e;
#undef MLF_E16V
#undef MLF_E16
  return cudaGetLastError();
}

}  // namespace mlfb200

#ifdef MLF_TRACE
extern "C" int mlf_b200_trace_read(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, mlfb200::g_trace, sizeof(long long) * (size_t)(n < 64 * 16 * 8 ? n : 64 * 16 * 8)) == cudaSuccess ? 0 : -1;
}
#endif
