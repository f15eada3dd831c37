/*
 * oracle/emgmm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of the Gaussian-mixture EM path of Etdescamps/MlFortran.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (mlfortran_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, no assertions and an unseeded RNG
 * for this path (tests/test_emgmm.f90:105-143 only prints), and it cannot be compiled in
 * this image (no Fortran compiler, no HDF5).  This restatement is therefore pinned to the
 * reference *formulas* through the analytic known-answer tests in tests/test_oracle_kat.py
 * and cross-checked against an independent NumPy restatement (oracle/emgmm_numpy.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Array layouts are the reference's (Fortran column-major):
 *   X(d,N)      -> X[i*d + a]          point-major
 *   Mu(d,K)     -> Mu[k*d + a]
 *   Cov(d,d,K)  -> Cov[k*d*d + b*d + a]   (symmetric)
 *   Proba(N,K)  -> P[k*N + i]          component-major
 *
 * LAPACK: the reference calls dsytrd/dorgtr/dsteqr (src/mlf_matrix.f90:269-275) and dsyrk
 * (:145).  orc_load_lapack() dlopens an LP64 LAPACK (the OpenBLAS bundled with SciPy exports
 * them with a "scipy_" prefix) so the same routines run in the same order; without it a
 * cyclic Jacobi eigensolver and a plain-loop SYRK are used (identical up to rounding:
 * only invC and lnDet -- not the eigenvectors -- enter the result).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*dsytrd_fn)(const char*, const int*, double*, const int*, double*, double*, double*, double*,
                          const int*, int*, size_t);
typedef void (*dorgtr_fn)(const char*, const int*, double*, const int*, const double*, double*, const int*,
                          int*, size_t);
typedef void (*dsteqr_fn)(const char*, const int*, double*, double*, double*, const int*, double*, int*, size_t);
typedef void (*dsyrk_fn)(const char*, const char*, const int*, const int*, const double*, const double*,
                         const int*, const double*, double*, const int*, size_t, size_t);

static dsytrd_fn p_dsytrd;
static dorgtr_fn p_dorgtr;
static dsteqr_fn p_dsteqr;
static dsyrk_fn p_dsyrk;
static int g_use_lapack = 0;
/* EXTENSION (not in the reference): diagonal covariances = Appendix A with every covariance restricted to
 * its diagonal: the E-step sees diag(C_k) only, the M-step keeps diag of the reference's C_k. */
static int g_cov_diag = 0;
void orc_set_cov_diag(int on) { g_cov_diag = on != 0; }
int orc_get_cov_diag(void) { return g_cov_diag; }

/* Returns 0 when all four routines were found (tries "scipy_<name>_" then "<name>_"). */
int orc_load_lapack(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  const char* pre[2] = {"scipy_", ""};
  for (int t = 0; t < 2; ++t) {
    char nm[64];
    snprintf(nm, sizeof nm, "%sdsytrd_", pre[t]); p_dsytrd = (dsytrd_fn)dlsym(h, nm);
    snprintf(nm, sizeof nm, "%sdorgtr_", pre[t]); p_dorgtr = (dorgtr_fn)dlsym(h, nm);
    snprintf(nm, sizeof nm, "%sdsteqr_", pre[t]); p_dsteqr = (dsteqr_fn)dlsym(h, nm);
    snprintf(nm, sizeof nm, "%sdsyrk_", pre[t]);  p_dsyrk = (dsyrk_fn)dlsym(h, nm);
    if (p_dsytrd && p_dorgtr && p_dsteqr && p_dsyrk) { g_use_lapack = 1; return 0; }
  }
  g_use_lapack = 0;
  return -2;
}
void orc_use_lapack(int on) { g_use_lapack = on && p_dsytrd && p_dorgtr && p_dsteqr && p_dsyrk; }
int orc_has_lapack(void) { return g_use_lapack; }
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Symmetric eigendecomposition A = V diag(w) V^T, ascending w, V column-major (V[i*n + a] is
 * component a of eigenvector i).  LAPACK path: the three calls of src/mlf_matrix.f90:269-275. */
static int sym_eig(int n, const double* A, double* w, double* V) {
  memcpy(V, A, sizeof(double) * n * n);
  if (g_use_lapack) {
    int info = 0, lwork = n * n > 1 ? n * n : 1;
    double* e = (double*)malloc(sizeof(double) * (n > 1 ? n - 1 : 1));
    double* tau = (double*)malloc(sizeof(double) * n);
    double* work = (double*)malloc(sizeof(double) * (lwork > 2 * n ? lwork : 2 * n + 2));
    p_dsytrd("U", &n, V, &n, w, e, tau, work, &lwork, &info, 1);
    if (info == 0) p_dorgtr("U", &n, V, &n, tau, work, &lwork, &info, 1);
    if (info == 0) p_dsteqr("V", &n, w, e, V, &n, work, &info, 1);
    free(e); free(tau); free(work);
    return info;
  }
  /* cyclic Jacobi (two-sided), then sort ascending */
  double* a = (double*)malloc(sizeof(double) * n * n);
  memcpy(a, A, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int p = 0; p < n; ++p) { diag += a[p * n + p] * a[p * n + p]; for (int q = p + 1; q < n; ++q) off += a[q * n + p] * a[q * n + p]; }
    if (off <= 1e-34 * (diag + off) || off == 0) break;
    for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
      const double apq = a[q * n + p];
      if (apq == 0) continue;
      const double theta = (a[q * n + q] - a[p * n + p]) / (2 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
      const double c = 1 / sqrt(t * t + 1), s = t * c;
      for (int r = 0; r < n; ++r) {  /* columns p,q */
        const double arp = a[p * n + r], arq = a[q * n + r];
        a[p * n + r] = c * arp - s * arq; a[q * n + r] = s * arp + c * arq;
      }
      for (int r = 0; r < n; ++r) {  /* rows p,q */
        const double apr = a[r * n + p], aqr = a[r * n + q];
        a[r * n + p] = c * apr - s * aqr; a[r * n + q] = s * apr + c * aqr;
      }
      for (int r = 0; r < n; ++r) {
        const double vrp = V[p * n + r], vrq = V[q * n + r];
        V[p * n + r] = c * vrp - s * vrq; V[q * n + r] = s * vrp + c * vrq;
      }
    }
  }
  for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
  for (int i = 0; i < n - 1; ++i) {  /* selection sort ascending */
    int m = i;
    for (int j = i + 1; j < n; ++j) if (w[j] < w[m]) m = j;
    if (m != i) {
      double t = w[i]; w[i] = w[m]; w[m] = t;
      for (int r = 0; r < n; ++r) { t = V[i * n + r]; V[i * n + r] = V[m * n + r]; V[m * n + r] = t; }
    }
  }
  free(a);
  return 0;
}

/* InverseSymMatrix(C, invC, lnDet, doEigen=.TRUE.)  -- src/mlf_matrix.f90:229-300, eigen branch :256-298.
 * Quirks kept: valM uses the default-real literal 10e-9 (:283); after LD <- 1/LD the
 * "lnDet" is 0.5*(d*log(2*pi) + sum(log(LD))) = d/2 log 2pi - 1/2 log det C (:285-290). */
int orc_inverse_sym_matrix(int d, const double* C, double* invC, double* lnDet) {
  double* LD = (double*)malloc(sizeof(double) * d);
  double* V = (double*)malloc(sizeof(double) * d * d);
  double* work = (double*)malloc(sizeof(double) * d * d);
  int info = sym_eig(d, C, LD, V);
  if (info != 0) { free(LD); free(V); free(work); return info; }
  double mx = LD[0];
  for (int i = 1; i < d; ++i) if (LD[i] > mx) mx = LD[i];
  const double valM = (double)10e-9f * mx; /* :283 */
  const double penality = 1.0 / valM;      /* :284 */
  for (int i = 0; i < d; ++i) LD[i] = (LD[i] <= valM) ? penality : 1.0 / LD[i]; /* :285-289 */
  double s = 0;
  for (int i = 0; i < d; ++i) s += log(LD[i]);
  const double pi = acos(-1.0); /* src/mlf_utils.f90:86 */
  *lnDet = 0.5 * (d * log(2.0 * pi) + s); /* :290 */
  for (int i = 0; i < d; ++i) for (int a = 0; a < d; ++a) work[i * d + a] = V[i * d + a] * LD[i]; /* :294-296 */
  for (int b = 0; b < d; ++b) for (int a = 0; a < d; ++a) { /* :297 invC = MATMUL(work, TRANSPOSE(V)) */
    double acc = 0;
    for (int i = 0; i < d; ++i) acc += work[i * d + a] * V[i * d + b];
    invC[b * d + a] = acc;
  }
  free(LD); free(V); free(work);
  return 0;
}

/* mlf_EvalGaussian -- src/unsupervised/mlf_gaussian.f90:63-79 */
int orc_eval_gaussian(int d, int64_t N, const double* X, double* P, const double* mu, const double* C,
                      double lambda, double* sumLL) {
  double* invC = (double*)malloc(sizeof(double) * d * d);
  double* z = (double*)malloc(sizeof(double) * 2 * d);
  double* t = z + d;
  double lnDet;
  double* Cd = NULL;
  if (g_cov_diag) { /* extension: off-diagonal entries are ignored */
    Cd = (double*)calloc((size_t)d * d, sizeof(double));
    for (int a = 0; a < d; ++a) Cd[a * d + a] = C[a * d + a];
    C = Cd;
  }
  int info = orc_inverse_sym_matrix(d, C, invC, &lnDet); /* :71 */
  free(Cd);
  if (info != 0) { free(invC); free(z); return info; }
  double ll = (double)N * (log(lambda) - lnDet); /* :73 */
  for (int64_t i = 0; i < N; ++i) {               /* :74-78 */
    const double* x = X + i * d;
    for (int a = 0; a < d; ++a) { z[a] = x[a] - mu[a]; t[a] = 0; }
    for (int b = 0; b < d; ++b) for (int a = 0; a < d; ++a) t[a] += invC[b * d + a] * z[b];
    double dp = 0;
    for (int a = 0; a < d; ++a) dp += z[a] * t[a];
    const double valProd = -0.5 * dp;
    ll += valProd;
    P[i] = lambda * exp(valProd - lnDet);
  }
  *sumLL = ll;
  free(invC); free(z);
  return 0;
}

/* GECovMatrix(C, A, Mu, W, minW) with W and minW present, Mu0 absent -- src/mlf_matrix.f90:99-147 */
void orc_ge_cov_matrix(int d, int64_t N, double* C, const double* A, double* Mu, const double* W, double minW) {
  double sw2 = 0;
  for (int64_t i = 0; i < N; ++i) sw2 += W[i] * W[i];
  const double ms = 1.0 / (1.0 - sw2); /* :119 */
  for (int a = 0; a < d; ++a) Mu[a] = 0;
  for (int64_t i = 0; i < N; ++i) for (int a = 0; a < d; ++a) Mu[a] += A[i * d + a] * W[i]; /* :132 */
  double* B = (double*)malloc(sizeof(double) * (size_t)d * (size_t)(N > 0 ? N : 1));
  int64_t j = 0;
  for (int64_t i = 0; i < N; ++i) { /* :134-138 */
    if (W[i] < minW) continue;
    const double f = ms * sqrt(W[i]);
    for (int a = 0; a < d; ++a) B[j * d + a] = f * (A[i * d + a] - Mu[a]);
    ++j;
  }
  if (g_use_lapack && j < 2147483647) { /* :145 dsyrk('L','N',d,j,1,B,d,0,C,d) */
    const int n = d, kk = (int)j; const double one = 1.0, zero = 0.0;
    p_dsyrk("L", "N", &n, &kk, &one, B, &n, &zero, C, &n, 1, 1);
  } else {
    for (int b = 0; b < d; ++b) for (int a = b; a < d; ++a) C[b * d + a] = 0;
    for (int64_t jj = 0; jj < j; ++jj)
      for (int b = 0; b < d; ++b) { const double bb = B[jj * d + b]; for (int a = b; a < d; ++a) C[b * d + a] += B[jj * d + a] * bb; }
  }
  for (int b = 0; b < d; ++b) for (int a = b + 1; a < d; ++a) C[a * d + b] = C[b * d + a]; /* :146 SymmetrizeMatrix, :56-63 */
  if (g_cov_diag) for (int b = 0; b < d; ++b) for (int a = 0; a < d; ++a) if (a != b) C[b * d + a] = 0.0; /* extension */
  free(B);
}

/* mlf_MaxGaussian -- src/unsupervised/mlf_gaussian.f90:41-51 */
double orc_max_gaussian(int d, int64_t N, const double* X, const double* Proba, double* mu, double* C) {
  double sumPi = 0;
  for (int64_t i = 0; i < N; ++i) sumPi += Proba[i]; /* :47 */
  double* W = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  for (int64_t i = 0; i < N; ++i) W[i] = Proba[i] / sumPi; /* :49 */
  const double minW = 1e-3 / (double)(float)N; /* :50  1d-3/real(NX): NX goes through default real */
  orc_ge_cov_matrix(d, N, C, X, mu, W, minW);
  free(W);
  return sumPi;
}

/* ExpStep -- src/unsupervised/mlf_emgmm.f90:191-207.  Returns 0 always (LAPACK status dropped, :198-202). */
int orc_exp_step(int d, int64_t N, int K, const double* X, double* P, const double* Mu, const double* Cov,
                 const double* Lambda, double* sumLL) {
  double* LL = (double*)calloc((size_t)K, sizeof(double));
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    double ll = 0;
    orc_eval_gaussian(d, N, X, P + (size_t)k * N, Mu + (size_t)k * d, Cov + (size_t)k * d * d, Lambda[k], &ll);
    LL[k] = ll;
  }
  double s = 0;
  for (int k = 0; k < K; ++k) s += LL[k]; /* :202 (reduction order is unspecified in the reference) */
  *sumLL = s;
  free(LL);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) { /* :206 */
    double r = 0;
    for (int k = 0; k < K; ++k) r += P[(size_t)k * N + i];
    for (int k = 0; k < K; ++k) P[(size_t)k * N + i] /= r;
  }
  return 0;
}

/* MaxStep -- src/unsupervised/mlf_emgmm.f90:174-188 */
void orc_max_step(int d, int64_t N, int K, const double* X, const double* P, double* Mu, double* Cov,
                  double* Lambda, double minProba) {
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k)
    Lambda[k] = orc_max_gaussian(d, N, X, P + (size_t)k * N, Mu + (size_t)k * d, Cov + (size_t)k * d * d);
  if (minProba > 0) { /* :184-186 */
    double mx = Lambda[0];
    for (int k = 1; k < K; ++k) if (Lambda[k] > mx) mx = Lambda[k];
    for (int k = 0; k < K; ++k) if (!(Lambda[k] >= minProba * mx)) Lambda[k] = minProba * mx; /* MAX(a,b) */
  }
  double s = 0;
  for (int k = 0; k < K; ++k) s += Lambda[k];
  for (int k = 0; k < K; ++k) Lambda[k] /= s; /* :187 */
}

/* niter EM iterations on an *initialised* object (stepF else-branch, src/unsupervised/mlf_emgmm.f90:146-150).
 * sumLL_trace[it] receives sumLL of iteration it (computed with the parameters at its start). P is N*K scratch. */
int orc_em_iterations(int d, int64_t N, int K, const double* X, double* P, double* Mu, double* Cov, double* Lambda,
                      double minProba, int64_t niter, double* sumLL_trace) {
  for (int64_t it = 0; it < niter; ++it) {
    double ll;
    int info = orc_exp_step(d, N, K, X, P, Mu, Cov, Lambda, &ll);
    if (info != 0) return info;
    if (sumLL_trace) sumLL_trace[it] = ll;
    orc_max_step(d, N, K, X, P, Mu, Cov, Lambda, minProba);
  }
  return 0;
}

/* mlf_model_getClass_proba -- src/mlf_models.f90:239-249: cl = MAXLOC(Proba, DIM=2); 1-based, first maximum.
 * gfortran's MAXLOC skips leading NaNs and yields 1 for an all-NaN row; reproduced here. */
void orc_get_class(int64_t N, int K, const double* P, int32_t* cl) {
  for (int64_t i = 0; i < N; ++i) {
    int best = -1; double bv = 0;
    for (int k = 0; k < K; ++k) {
      const double v = P[(size_t)k * N + i];
      if (v != v) continue;
      if (best < 0 || v > bv) { best = k; bv = v; }
    }
    cl[i] = best < 0 ? 1 : best + 1;
  }
}

/* mlf_EvaluateClass -- src/unsupervised/mlf_kmeans.f90:151-165: Euclidean distance, MINLOC first minimum, 1-based */
void orc_kmeans_evaluate_class(int d, int64_t N, int K, const double* X, const double* Mu, int32_t* cl, double* minDist) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    int best = 0; double bd = 0;
    for (int k = 0; k < K; ++k) {
      double s = 0;
      for (int a = 0; a < d; ++a) { const double t = X[i * d + a] - Mu[(size_t)k * d + a]; s += t * t; }
      const double dist = sqrt(s);
      if (k == 0 || dist < bd) { best = k; bd = dist; }
    }
    if (cl) cl[i] = best + 1;
    if (minDist) minDist[i] = bd;
  }
}

/* EvaluateCentres -- src/unsupervised/mlf_kmeans_naive.f90:74-96, without the RNG re-seeding of empty
 * clusters (:89-91); returns the number of empty clusters (their centre is left at 0). */
int orc_kmeans_evaluate_centres(int d, int64_t N, int K, const double* X, double* Mu, const int32_t* cl) {
  int64_t* n = (int64_t*)calloc((size_t)K, sizeof(int64_t));
  memset(Mu, 0, sizeof(double) * (size_t)K * d);
  for (int64_t i = 0; i < N; ++i) {
    const int k = cl[i] - 1; n[k]++;
    for (int a = 0; a < d; ++a) Mu[(size_t)k * d + a] += X[i * d + a];
  }
  int empty = 0;
  for (int k = 0; k < K; ++k) {
    if (n[k] == 0) { ++empty; continue; }
    for (int a = 0; a < d; ++a) Mu[(size_t)k * d + a] /= (double)n[k];
  }
  free(n);
  return empty;
}
