/* A plain C client of the drop-in boundary, in the shape of the reference's own C test for k-means
 * (tests/test_kmeans_c.c:38-67): create the object, step it, read the parameters through mlf_getrsc, checkpoint,
 * restore, continue.  Exit code 0 = every check passed.  Built and run by tests/test_c_client.py. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlf_b200.h"

#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, msg); return 1; } } while (0)

static uint64_t rng = 88172645463325252ull;
static double urand(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; }
static double nrand(void) { return sqrt(-2.0 * log(urand() + 1e-300)) * cos(6.283185307179586 * urand()); }

int main(int argc, char **argv) {
  const char *h5path = argc > 1 ? argv[1] : "emgmm_c_client.h5";
  const int nX = 6000, nY = 3, nC = 3;
  double *X = (double *)malloc(sizeof(double) * nX * nY);
  const double centres[3][3] = {{-8, 0, 0}, {8, 0, 0}, {0, 9, 3}};
  for (int i = 0; i < nX; ++i)
    for (int a = 0; a < nY; ++a) X[i * nY + a] = centres[i % 3][a] + nrand();

  CHECK(mlf_init() == 0, "mlf_init (no CUDA device?)");
  MLF_OBJ *em = mlf_emgmm_c(X, nX, nY, nC, 1.0, NULL);
  CHECK(em != NULL, "mlf_emgmm_c");
  CHECK(mlf_getnumrsc(em) == 7, "7 resource slots");
  CHECK(strcmp(mlf_getinfo(em, 5, mlf_NAME), "Mu") == 0, "slot 5 is Mu");
  CHECK(strcmp(mlf_getinfo(em, 4, mlf_FIELDS), "cpuTime;realTime;sumLL") == 0, "rVar fields");

  /* a fresh object initialises itself (k-means) in its first iteration, then runs niter-1 EM iterations */
  int64_t niter = 12;
  double dt = -1;
  CHECK(mlf_step(em, &dt, &niter) == 0, "mlf_step");
  CHECK(dt >= 0, "dt written");

  MLF_DT t;
  int rank, dims[MLF_MAXRANK];
  double *Mu = (double *)mlf_getrsc(em, 5, &t, &rank, dims, NULL); /* direct pointer, as tests/test_kmeans_c.c:58 */
  CHECK(Mu && rank == 2 && dims[0] == nY && dims[1] == nC && t.dt == mlf_DOUBLE && t.access == mlf_DIRECT, "Mu slot shape");
  for (int c = 0; c < 3; ++c) { /* every true centre is recovered by some component */
    double best = 1e300;
    for (int k = 0; k < nC; ++k) {
      double s = 0;
      for (int a = 0; a < nY; ++a) s += (Mu[k * nY + a] - centres[c][a]) * (Mu[k * nY + a] - centres[c][a]);
      if (s < best) best = s;
    }
    CHECK(sqrt(best) < 0.3, "centre recovered");
  }
  int64_t *iVar = (int64_t *)mlf_getrsc(em, 3, &t, &rank, dims, NULL);
  CHECK(iVar && t.dt == mlf_INT64 && iVar[0] == 12, "niter counter");
  double *lambda = (double *)mlf_getrsc(em, 7, &t, &rank, dims, NULL);
  CHECK(lambda && fabs(lambda[0] - 1.0 / 3) < 1e-15, "minProba = 1 forces uniform weights");

  /* labels of the first points through the model interface (Fortran-true argument order: nX = dimension) */
  int cl[9];
  CHECK(mlf_getClass(em, X, cl, nY, 9) == 0, "mlf_getClass");
  CHECK(cl[0] == cl[3] && cl[1] == cl[4] && cl[0] != cl[1] && cl[0] >= 1 && cl[0] <= nC, "labels follow the generating pattern");
  CHECK(mlf_getNumClasses(em) == nC, "mlf_getNumClasses");

  /* checkpoint, restore into a new object (nC comes from the file), continue both: identical parameters */
  MLF_OBJ *h5 = mlf_hdf5_createFile(h5path, 1);
  CHECK(h5 != NULL, "mlf_hdf5_createFile");
  CHECK(mlf_pushState(h5, em, 0) == 0, "mlf_pushState");
  CHECK(mlf_dealloc(h5) == 0, "close file");
  MLF_OBJ *h5r = mlf_hdf5_openFile(h5path, 0);
  CHECK(h5r != NULL, "mlf_hdf5_openFile");
  MLF_OBJ *em2 = mlf_emgmm_c(X, nX, nY, 1 /* overwritten by the checkpoint */, 1.0, h5r);
  CHECK(em2 != NULL && mlf_getNumClasses(em2) == nC, "restore");
  niter = 3;
  CHECK(mlf_step(em, &dt, &niter) == 0 && mlf_step(em2, &dt, &niter) == 0, "continue");
  double *Mu2 = (double *)mlf_getrsc(em2, 5, &t, &rank, dims, NULL);
  double *Cov = (double *)mlf_getrsc(em, 6, &t, &rank, dims, NULL), *Cov2 = (double *)mlf_getrsc(em2, 6, &t, &rank, dims, NULL);
  CHECK(memcmp(Mu, Mu2, sizeof(double) * nY * nC) == 0, "restored object continues bit-identically (Mu)");
  CHECK(memcmp(Cov, Cov2, sizeof(double) * nY * nY * nC) == 0, "restored object continues bit-identically (Cov)");

  /* error behaviour of the reference: -1 / NULL on foreign handles */
  CHECK(mlf_step(NULL, &dt, &niter) == -1 && mlf_getrsc(NULL, 5, &t, &rank, dims, NULL) == NULL, "NULL handles");
  CHECK(mlf_dealloc(em2) == 0 && mlf_dealloc(h5r) == 0 && mlf_dealloc(em) == 0 && mlf_quit() == 0, "teardown");
  free(X);
  printf("emgmm C client: all checks passed\n");
  return 0;
}
