// The reference's own C++ wrapper classes (include/mlf.hpp + src/cpp_intf/mlfcpp.cpp, compiled UNMODIFIED from
// /root/reference by tests/cpp_client/build.sh) driving an EM-GMM object of libmlfortran_b200.so: MlfStepObject::step /
// printFields / printLine, MlfObject::getRsc into MlfDataMatrix / MlfData3DMatrix / MlfDataVector (direct pointers),
// MlfHdf5::createFile / pushState / openFile ("must keep working").  Only the constructor is ours: the
// reference has no C constructor for EM, so mlf_emgmm_c is declared here by hand.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>

#include "mlf.hpp"

extern "C" MLF_OBJ *mlf_emgmm_c(double *X, int nX, int nY, int nC, double minProba, MLF_OBJ *handler);

using namespace MlFortran;

#define CHECK(c, msg) do { if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, msg); return 1; } } while (0)

int main(int argc, char **argv) {
  std::string path = argc > 1 ? argv[1] : "ref_wrapper_client.h5";
  const int nX = 4000, nY = 2, nC = 2;
  std::vector<double> X((size_t)nX * nY);
  unsigned long long s = 88172645463325252ull;
  auto u = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  for (int i = 0; i < nX; ++i)
    for (int a = 0; a < nY; ++a) X[(size_t)i * nY + a] = (i % 2 ? 6.0 : -6.0) * (a == 0) + std::sqrt(-2 * std::log(u() + 1e-300)) * std::cos(6.283185307179586 * u());
  CHECK(mlf_init() == 0, "mlf_init");
  MlfStepObject em(mlf_emgmm_c(X.data(), nX, nY, nC, 1.0, nullptr));
  CHECK(em.get() != nullptr, "mlf_emgmm_c");
  double dt = -1;
  int64_t n = 10;
  CHECK(em.step(dt, n) == 0, "MlfStepObject::step");
  std::ostringstream fields, line;
  em.printFields(fields);
  em.printLine(line);
  std::cout << fields.str() << line.str();
  CHECK(fields.str() == "niter;cpuTime;realTime;sumLL\n", "printFields");
  CHECK(line.str().rfind("10;0;", 0) == 0, "printLine starts with niter;<second iVar entry>");
  CHECK(em.getNumRsc() == 7 && em.getIdName("Mu") == 5 && em.getIdName("lambda") == 7, "slot names");
  MlfDataMatrix<double> Mu;
  MlfData3DMatrix<double> Cov;
  MlfDataVector<double> lambda;
  CHECK(em.getRsc(em.getIdName("Mu"), Mu) == MlfDataAccessType::Direct, "Mu is a direct pointer");
  CHECK(em.getRsc(em.getIdName("Cov"), Cov) == MlfDataAccessType::Direct, "Cov");
  CHECK(em.getRsc(em.getIdName("lambda"), lambda) == MlfDataAccessType::Direct, "lambda");
  CHECK(Mu.getSize() == (size_t)nY * nC && Cov.getSize() == (size_t)nY * nY * nC && lambda.getSize() == (size_t)nC, "shapes (Fortran order)");
  const double m1 = Mu(1, 1), m2 = Mu(1, 2);   // Fortran indexing: first coordinate of component 1 and 2
  CHECK(std::fabs(std::fabs(m1) - 6.0) < 0.2 && std::fabs(m1 + m2) < 0.3, "the two centres are found");
  CHECK(std::fabs(Cov(1, 1, 1) - 1.0) < 0.15 && std::fabs(lambda(1) - 0.5) < 1e-15, "unit covariance, uniform weights (minProba = 1)");
  bool threw = false;
  try { MlfDataVector<double> wrong; em.getRsc(em.getIdName("Mu"), wrong); } catch (MlfException &e) { threw = e.errorType == mlf_WRONGRANK; }
  CHECK(threw, "rank mismatch raises MlfException(mlf_WRONGRANK)");
  {
    MlfHdf5 h5;
    CHECK(h5.createFile(path) == 0 && h5.isInit() && !h5.hasData(), "MlfHdf5::createFile");
    CHECK(h5.pushState(em) == 0, "pushState (create)");
    n = 2;
    CHECK(em.step(dt, n) == 0, "continue");
    CHECK(h5.pushState(em) == 0, "pushState (overwrite = 1 after the first push)");
  }   // ~MlfHdf5 -> mlf_dealloc closes the file
  {
    MlfHdf5 h5;
    CHECK(h5.readWOrCreate(path) && h5.hasData(), "readWOrCreate opens the existing file");
    MlfStepObject em2(mlf_emgmm_c(X.data(), nX, nY, 1, 1.0, h5.get()));
    CHECK(em2.get() != nullptr, "restore from the checkpoint");
    MlfDataMatrix<double> Mu2;
    em2.getRsc(em2.getIdName("Mu"), Mu2);
    CHECK(Mu2.getSize() == Mu.getSize() && Mu2(1, 1) == Mu(1, 1) && Mu2(2, 2) == Mu(2, 2), "restored parameters are the pushed ones");
    std::ostringstream l2;
    em2.printLine(l2);
    CHECK(l2.str().rfind("12;", 0) == 0, "restored iteration counter");
  }
  em.finalize();
  CHECK(mlf_quit() == 0, "mlf_quit");
  std::cout << "reference C++ wrapper client: all checks passed" << std::endl;
  return 0;
}
