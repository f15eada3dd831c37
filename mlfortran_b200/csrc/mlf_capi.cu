// C-ABI of the drop-in (include/mlf_b200.h).  Host-side mirror of the reference's step-object runtime for the EM
// path: resource slots (src/mlf_rsc_array.f90), step wrapper and timers (src/mlf_step_algo.f90:144-219), EM object
// (src/unsupervised/mlf_emgmm.f90:76-153), model getters (src/mlf_models.f90:239-331), checkpoint
// (src/mlf_hdf5.f90:671-737).  All compute is delegated to EmgmmEngine (CUDA); there is no CPU fallback.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mlf_b200.h"
#include "emgmm_engine.h"
#include "emgmm_multi.h"
#include "mlf_hdf5_min.h"
#include "mlf_nccl.h"

using mlfb200::EmgmmEngine;
using mlfb200::MultiEngine;

namespace {

constexpr uint64_t kMagic = 0x4d4c4642323030ull;  // "MLFB200"
enum Kind { KIND_EMGMM = 1, KIND_HDF5 = 2 };

struct Handle {
  uint64_t magic = kMagic;
  Kind kind;
  explicit Handle(Kind k) : kind(k) {}
  virtual ~Handle() { magic = 0; }
};

struct Slot {
  std::string name, fields;  // NUL-terminated by std::string::c_str()
  short dt = mlf_DOUBLE;
  int rank = 1;
  int dims[3] = {0, 0, 0};  // Fortran order
  void* ptr = nullptr;
  size_t bytes = 0;
  bool has_fields = false;
};

// One open checkpoint file, shared by the file handle and the group handles opened from it (mlf_hdf5_file /
// mlf_hdf5_group, src/mlf_hdf5.f90:46-82).  The file is rewritten as a whole -- only when something was put since the
// last write (dirty), never on a plain open/close, and never for a file that holds anything this module cannot
// represent (mlf_hdf5_openFile refuses rw on such files: a rewrite would drop that content).
struct H5File {
  std::string path;
  bool writable = true;
  bool dirty = false;
  mlfh5::Contents c;
  mlfh5::Dataset* find(const std::string& n) {
    for (auto& d : c.ds)
      if (d.name == n) return &d;
    return nullptr;
  }
  bool has_group(const std::string& g) const {
    for (const auto& x : c.groups)
      if (x == g) return true;
    return false;
  }
  int flush() {
    if (!writable || !dirty) return 0;
    const int r = mlfh5::write_file(path, c);
    if (r == 0) dirty = false;
    return r;
  }
};

struct Hdf5Handle : Handle {
  std::shared_ptr<H5File> file;
  std::string prefix;   // "" for the file handle, "group/sub/" for a group handle
  Hdf5Handle() : Handle(KIND_HDF5) {}
  bool writable() const { return file->writable; }
  mlfh5::Dataset* find(const std::string& n) { return file->find(prefix + n); }
  // false when the path collides with the file's structure: a group of that name, or a dataset where a group is needed
  bool put(mlfh5::Dataset&& nd) {
    nd.name = prefix + nd.name;
    if (nd.name.empty() || nd.name.back() == '/' || file->has_group(nd.name)) return false;
    for (size_t p = nd.name.find('/'); p != std::string::npos; p = nd.name.find('/', p + 1)) {
      const std::string anc = nd.name.substr(0, p);
      if (anc.empty() || file->find(anc)) return false;
    }
    for (size_t p = nd.name.find('/'); p != std::string::npos; p = nd.name.find('/', p + 1))
      if (!file->has_group(nd.name.substr(0, p))) file->c.groups.push_back(nd.name.substr(0, p));
    mlfh5::Dataset* dst = file->find(nd.name);
    if (dst) *dst = std::move(nd);
    else file->c.ds.push_back(std::move(nd));
    file->dirty = true;
    return true;
  }
  int flush() { return file->flush(); }
};

struct EmObject : Handle {
  EmObject() : Handle(KIND_EMGMM) {}
  // slot storage.  rPar holds [cpuMax, realMax, minProba] but is exposed with extent 2: the reference binds
  // minProba to rPar(3) of a 2-element array (src/unsupervised/mlf_emgmm.f90:84,104; src/mlf_step_algo.f90:124),
  // so it is neither in the slot's dims nor in the checkpoint.  Same observable layout here, minus the overrun.
  int64_t iPar[1] = {0};
  double rPar[3] = {0, 0, 0};
  int64_t iVar[2] = {0, 0};
  double rVar[3] = {0, 0, 0};
  std::vector<double> Mu, Cov, lambda;
  Slot slots[7];
  int d = 0, K = 0;
  int64_t N = 0;
  bool initialized = false;
  const double* X = nullptr;
  bool x_on_device = false;
  EmgmmEngine eng;                     // single-GPU object (or one shard of a one-process-per-GPU job)
  std::unique_ptr<MultiEngine> multi;  // single-process multi-GPU object: one engine per device, peer-memory all-reduce
  // dispatch to whichever of the two this object is
  int e_upload() { return multi ? multi->upload_params(Mu.data(), Cov.data(), lambda.data()) : eng.upload_params(Mu.data(), Cov.data(), lambda.data()); }
  int e_download() { return multi ? multi->download_params(Mu.data(), Cov.data(), lambda.data()) : eng.download_params(Mu.data(), Cov.data(), lambda.data()); }
  int e_kmeans(int nkm, uint64_t sd) { return multi ? multi->kmeans_init(nkm, sd) : eng.kmeans_init(nkm, sd); }
  int e_iterate(int64_t n, double mp, double* ll, double* tr) { return multi ? multi->iterate(n, mp, ll, tr) : eng.iterate(n, mp, ll, tr); }
  int e_expect(const double* Xq, int64_t nq, double* P, int32_t* cl) { return multi ? multi->expectation(Xq, nq, P, cl) : eng.expectation(Xq, false, nq, P, cl); }
  std::string e_error() { return multi ? multi->last_error() : eng.last_error(); }
  std::vector<double> trace;
  std::string err, plan;
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;   // native NCCL all-reduce (mlf_b200_set_nccl_comm / mlf_b200_nccl_init)
  bool nccl_owned = false;
  ~EmObject() {
    if (nccl_comm && nccl_owned)
      if (const mlfb200::NcclApi* api = mlfb200::nccl_api()) api->CommDestroy(nccl_comm);
  }
  uint64_t seed = 0x9E3779B97F4A7C15ull;

  int64_t& niter() { return iVar[0]; }
  int64_t& niterMax() { return iPar[0]; }
  double& cpuTime() { return rVar[0]; }
  double& realTime() { return rVar[1]; }
  double& sumLL() { return rVar[2]; }
  double& cpuMax() { return rPar[0]; }
  double& realMax() { return rPar[1]; }
  double& minProba() { return rPar[2]; }

  void build_slots() {
    auto set = [&](int i, const char* nm, const char* fl, short dt, int rank, int d0, int d1, int d2, void* p, size_t b) {
      Slot& s = slots[i];
      s.name = nm; s.dt = dt; s.rank = rank; s.dims[0] = d0; s.dims[1] = d1; s.dims[2] = d2; s.ptr = p; s.bytes = b;
      s.has_fields = fl != nullptr;
      if (fl) s.fields = fl;
    };
    // names/fields: src/mlf_rsc_array.f90:507-510, src/mlf_step_algo.f90:127-141, src/unsupervised/mlf_emgmm.f90:90-104
    set(0, "iPar", "niterMax", mlf_INT64, 1, 1, 0, 0, iPar, sizeof(int64_t));
    set(1, "rPar", "cpuMax;realMax;minProba", mlf_DOUBLE, 1, 2, 0, 0, rPar, 2 * sizeof(double));
    set(2, "iVar", "niter", mlf_INT64, 1, 2, 0, 0, iVar, 2 * sizeof(int64_t));
    set(3, "rVar", "cpuTime;realTime;sumLL", mlf_DOUBLE, 1, 3, 0, 0, rVar, 3 * sizeof(double));
    set(4, "Mu", nullptr, mlf_DOUBLE, 2, d, K, 0, Mu.data(), Mu.size() * sizeof(double));
    set(5, "Cov", nullptr, mlf_DOUBLE, 3, d, d, K, Cov.data(), Cov.size() * sizeof(double));
    set(6, "lambda", nullptr, mlf_DOUBLE, 1, K, 0, 0, lambda.data(), lambda.size() * sizeof(double));
  }

  // mlf_step_obj_reinit + mlf_algo_emgmm_reinit (src/mlf_step_algo.f90:144-153, src/unsupervised/mlf_emgmm.f90:115-122)
  void reinit() {
    niter() = 0;
    niterMax() = std::numeric_limits<int64_t>::max();
    cpuTime() = 0;
    realTime() = 0;
    realMax() = std::numeric_limits<double>::max();
    cpuMax() = std::numeric_limits<double>::max();
    std::fill(Mu.begin(), Mu.end(), 0.0);
    std::fill(Cov.begin(), Cov.end(), 0.0);
    initialized = false;
  }
};

Handle* as_handle(MLF_OBJ* p) {
  if (!p) return nullptr;
  Handle* h = reinterpret_cast<Handle*>(p);
  return h->magic == kMagic ? h : nullptr;
}
EmObject* as_em(MLF_OBJ* p) {
  Handle* h = as_handle(p);
  return (h && h->kind == KIND_EMGMM) ? static_cast<EmObject*>(h) : nullptr;
}
Hdf5Handle* as_h5(MLF_OBJ* p) {
  Handle* h = as_handle(p);
  return (h && h->kind == KIND_HDF5) ? static_cast<Hdf5Handle*>(h) : nullptr;
}

// mlf_algo_emgmm_stepF (src/unsupervised/mlf_emgmm.f90:125-153)
int em_stepF(EmObject* o, int64_t niter) {
  if (niter <= 0) return 0;
  int64_t em_iters = niter;
  if (!o->initialized) {
    // first iteration of a fresh object: k-means (nkm = 16), Cov = I, lambda = 1/K  (:135-145)
    int r = o->e_kmeans(16, o->seed);
    if (r < 0) { o->err = o->e_error(); return r; }
    r = o->e_download();
    if (r < 0) return r;
    o->initialized = true;
    em_iters = niter - 1;
  } else {
    // host mirrors are the source of truth between steps: callers hold direct pointers (mlf_getrsc) and may
    // write them, and mlf_updatersc memcpy's into them
    int r = o->e_upload();
    if (r < 0) { o->err = o->e_error(); return r; }
  }
  if (em_iters > 0) {
    o->trace.assign((size_t)em_iters, 0.0);
    int r = o->e_iterate(em_iters, o->minProba(), &o->sumLL(), o->trace.data());
    if (r < 0) { o->err = o->e_error(); return r; }
    r = o->e_download();
    if (r < 0) { o->err = o->e_error(); return r; }
  } else {
    o->trace.clear();
  }
  return 0;
}

bool restore_from(Hdf5Handle* h, EmObject* o, int d, int& K) {
  // data_handler%getData for every slot (src/mlf_rsc_array.f90:336-338,366-368,396-398; src/mlf_hdf5.f90:303-410):
  // Mu is read with fixed_dims=[T,F] so K comes from the file (src/unsupervised/mlf_emgmm.f90:90-101)
  mlfh5::Dataset* mu = h->find("Mu");
  mlfh5::Dataset* cov = h->find("Cov");
  mlfh5::Dataset* lam = h->find("lambda");
  if (!mu || !cov || !lam) return false;
  if (mu->type != mlfh5::F64 || mu->dims.size() != 2 || (int)mu->dims[1] != d) return false;
  K = (int)mu->dims[0];
  if (cov->type != mlfh5::F64 || cov->dims.size() != 3 || (int)cov->dims[0] != K || (int)cov->dims[1] != d || (int)cov->dims[2] != d) return false;
  if (lam->type != mlfh5::F64 || lam->dims.size() != 1 || (int)lam->dims[0] != K) return false;
  o->Mu.assign((size_t)K * d, 0.0);
  o->Cov.assign((size_t)K * d * d, 0.0);
  o->lambda.assign((size_t)K, 0.0);
  std::memcpy(o->Mu.data(), mu->data.data(), o->Mu.size() * 8);
  std::memcpy(o->Cov.data(), cov->data.data(), o->Cov.size() * 8);
  std::memcpy(o->lambda.data(), lam->data.data(), o->lambda.size() * 8);
  auto rd = [&](const char* nm, mlfh5::DType t, void* dst, size_t n) {
    mlfh5::Dataset* s = h->find(nm);
    if (!s || s->type != t || s->dims.size() != 1 || s->dims[0] != n) return false;
    std::memcpy(dst, s->data.data(), n * 8);
    return true;
  };
  return rd("iPar", mlfh5::I64, o->iPar, 1) && rd("rPar", mlfh5::F64, o->rPar, 2) && rd("iVar", mlfh5::I64, o->iVar, 2) &&
         rd("rVar", mlfh5::F64, o->rVar, 3);
}

MLF_OBJ* em_create(const double* X, int64_t nX, int nY, int nC, double minProba, MLF_OBJ* handler, int device, unsigned flags,
                   const std::vector<int>* devices = nullptr) {
  if (!X && nX > 0) return nullptr;
  if (nX < 0 || nY < 1) return nullptr;
  std::unique_ptr<EmObject> o(new EmObject());
  o->d = nY;
  o->N = nX;
  o->X = X;
  o->x_on_device = (flags & MLF_B200_X_ON_DEVICE) != 0;
  int K = nC;
  Hdf5Handle* h5 = nullptr;
  if (handler) {
    h5 = as_h5(handler);
    if (!h5) return nullptr;
    if (!restore_from(h5, o.get(), nY, K)) {
      fprintf(stderr, "emgmm: error creating Mu\n");  // CheckF message of src/unsupervised/mlf_emgmm.f90:92
      return nullptr;
    }
  }
  if (K < 1) return nullptr;
  o->K = K;
  if (!h5) {
    o->Mu.assign((size_t)K * nY, 0.0);
    o->Cov.assign((size_t)K * nY * nY, 0.0);
    o->lambda.assign((size_t)K, 0.0);
  }
  o->build_slots();
  int cur = 0;
  if (device < 0) {
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    device = cur;
  }
  if (devices) {
    if (o->x_on_device) { fprintf(stderr, "mlf_b200_emgmm_create_multi: X must be a host array\n"); return nullptr; }
    o->multi.reset(new MultiEngine());
    if (o->multi->init(X, nX, nY, K, *devices, (flags & MLF_B200_COV_DIAG) != 0) < 0) {
      fprintf(stderr, "mlf_b200_emgmm_create_multi: %s\n", o->multi->last_error().c_str());
      return nullptr;
    }
  } else if (o->eng.init(X, o->x_on_device, nX, nY, K, device, (flags & MLF_B200_COV_DIAG) != 0) < 0) {
    fprintf(stderr, "mlf_emgmm_c: %s\n", o->eng.last_error().c_str());
    return nullptr;
  }
  if (h5) {
    o->minProba() = minProba;  // not part of the checkpoint (see rPar note above)
    o->initialized = true;     // src/unsupervised/mlf_emgmm.f90:106-107
    if (o->e_upload() < 0) return nullptr;
  } else {
    o->reinit();  // :109
    o->minProba() = minProba;  // InitOrDefault(this%minProba, 0d0, minProba) (:105)
  }
  return o.release();
}

}  // namespace

extern "C" {

const char* mlf_b200_version(void) { return "mlfortran_b200 0.1 (sm_100a, fp64 DMMA EM-GMM)"; }

static int selftest_exp_impl(const double* in, double* out, int64_t n, int device, int which) {
  if (!in || !out || n < 0) return -1;
  if (n == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -7;   // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return -7;
  double *di = nullptr, *dout = nullptr;
  int rc = -7;
  if (cudaMalloc((void**)&di, (size_t)n * sizeof(double)) == cudaSuccess && cudaMalloc((void**)&dout, (size_t)n * sizeof(double)) == cudaSuccess &&
      cudaMemcpy(di, in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
      mlfb200::launch_selftest_exp(di, dout, n, which, 0) == cudaSuccess &&
      cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess)
    rc = 0;
  cudaFree(di);
  cudaFree(dout);
  return rc;
}
int mlf_b200_selftest_exp(const double* in, double* out, int64_t n, int device) { return selftest_exp_impl(in, out, n, device, 0); }
int mlf_b200_selftest_exp128(const double* in, double* out, int64_t n, int device) { return selftest_exp_impl(in, out, n, device, 1); }

int mlf_init(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    fprintf(stderr, "mlf_init: no CUDA device visible; mlfortran_b200 has no CPU fallback\n");
    return -1;
  }
  return 0;
}
int mlf_quit(void) { return 0; }

int mlf_dealloc(MLF_OBJ* obj) {
  Handle* h = as_handle(obj);
  if (!h) return -1;
  if (h->kind == KIND_HDF5) static_cast<Hdf5Handle*>(h)->flush();
  delete h;
  return 0;
}

int mlf_getnumrsc(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return o ? 7 : -1;
}

void* mlf_getrsc(MLF_OBJ* obj, int id, MLF_DT* dt, int* rank, int* dims, void* data) {
  EmObject* o = as_em(obj);
  if (!o || id < 1 || id > 7) return nullptr;
  const Slot& s = o->slots[id - 1];
  if (rank) *rank = s.rank;
  if (dims)
    for (int i = 0; i < s.rank; ++i) dims[i] = s.dims[i];
  if (dt) { dt->dt = s.dt; dt->access = mlf_DIRECT; }
  if (data) {
    std::memcpy(data, s.ptr, s.bytes);
    return data;
  }
  return s.ptr;
}

char* mlf_getinfo(MLF_OBJ* obj, int id, int type) {
  EmObject* o = as_em(obj);
  if (!o || id < 1 || id > 7) return nullptr;
  Slot& s = o->slots[id - 1];
  if (type == mlf_NAME) return const_cast<char*>(s.name.c_str());
  if (type == mlf_FIELDS && s.has_fields) return const_cast<char*>(s.fields.c_str());
  return nullptr;  // r_desc is never set on this path
}

int mlf_updatersc(MLF_OBJ* obj, int id, void* data) {
  EmObject* o = as_em(obj);
  if (!o || id < 1 || id > 7) return -1;
  const Slot& s = o->slots[id - 1];
  if (data) std::memcpy(s.ptr, data, s.bytes);
  return 0;
}

int mlf_pushState(MLF_OBJ* handler, MLF_OBJ* obj, int override_) {
  Hdf5Handle* h = as_h5(handler);
  EmObject* o = as_em(obj);
  if (!h || !o) return -1;
  if (!h->writable()) return mlf_FILEERROR;
  // one dataset per allocated slot, named by r_name, dims reversed (src/mlf_hdf5.f90:696-736); with override the
  // datasets must exist already (H5Dopen_f) and are rewritten, without it they must not (H5Dcreate_f) -- :294-300
  for (const Slot& s : o->slots) {
    mlfh5::Dataset* dst = h->find(s.name);
    if (dst && !override_) {
      fprintf(stderr, "Error creating dataset: %s\n", s.name.c_str());
      return mlf_FILEERROR;
    }
    if (!dst && override_) {
      fprintf(stderr, "Error opening resource %s\n", s.name.c_str());
      return mlf_FILEERROR;
    }
  }
  for (const Slot& s : o->slots) {
    mlfh5::Dataset nd;
    nd.name = s.name;
    nd.type = (s.dt == mlf_INT64) ? mlfh5::I64 : mlfh5::F64;
    for (int i = s.rank - 1; i >= 0; --i) nd.dims.push_back((uint64_t)s.dims[i]);
    nd.data.assign((const uint8_t*)s.ptr, (const uint8_t*)s.ptr + s.bytes);
    if (!h->put(std::move(nd))) return mlf_FILEERROR;
  }
  return h->flush();
}

int mlf_step(MLF_OBJ* obj, double* dt, int64_t* niter) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  const int64_t n = niter ? *niter : 1;
  // mlf_step_f (src/mlf_step_algo.f90:187-201): timers around stepF, counters accumulate
  const auto t0 = std::chrono::steady_clock::now();
  const std::clock_t c0 = std::clock();
  int info = em_stepF(o, n);
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  o->realTime() += el;
  o->cpuTime() += (double)(std::clock() - c0) / CLOCKS_PER_SEC;
  o->niter() += n;
  if (dt) *dt = el;
  if (info != mlf_STEP_OK) return info;
  if (o->cpuTime() > o->cpuMax() || o->realTime() > o->realMax()) return mlf_STEP_STOP;
  return mlf_STEP_OK;
}

int mlf_getNumClasses(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return o ? o->K : -1;
}

static int em_expect(EmObject* o, const double* X, int nX, int nIn, double* P, int* Cl) {
  if (!o->initialized) return mlf_UNINIT;  // src/unsupervised/mlf_emgmm.f90:166-169
  if (nX != o->d || nIn < 0) return -1;
  int r = o->e_upload();
  if (r < 0) return r;
  return o->e_expect(X, nIn, P, Cl);
}

int mlf_getClass(MLF_OBJ* obj, const double* X, int* Cl, int nX, int nIn) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  return em_expect(o, X, nX, nIn, nullptr, Cl);
}

int mlf_getProba(MLF_OBJ* obj, const double* X, double* P, int nX, int nIn, int nCl) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  if (nCl != o->K) return -1;
  return em_expect(o, X, nX, nIn, P, nullptr);
}

MLF_OBJ* mlf_hdf5_createFile(const char* fname, int trunk) {
  if (!fname) return nullptr;
  if (trunk == 0) {  // H5F_ACC_EXCL_F: fail if the file exists (src/mlf_hdf5.f90:178-179)
    FILE* f = fopen(fname, "rb");
    if (f) { fclose(f); fprintf(stderr, "Error while creating file: %s\n", fname); return nullptr; }
  }
  std::unique_ptr<Hdf5Handle> h(new Hdf5Handle());
  h->file = std::make_shared<H5File>();
  h->file->path = fname;
  h->file->writable = true;
  h->file->dirty = true;   // H5Fcreate_f leaves an (empty) file behind at once
  if (h->flush() < 0) { fprintf(stderr, "Error while creating file: %s\n", fname); return nullptr; }
  return h.release();
}

MLF_OBJ* mlf_hdf5_openFile(const char* fname, int rw) {
  if (!fname) return nullptr;
  std::unique_ptr<Hdf5Handle> h(new Hdf5Handle());
  h->file = std::make_shared<H5File>();
  h->file->path = fname;
  h->file->writable = rw != 0;
  if (mlfh5::read_file(h->file->path, h->file->c) < 0) { fprintf(stderr, "Error while opening file: %s\n", fname); return nullptr; }
  if (rw != 0 && !h->file->c.lossy.empty()) {
    // the built-in writer rewrites the whole file from what the reader understood: content it cannot represent would be
    // lost (the reference's H5Fopen_f(H5F_ACC_RDWR_F) keeps it), so such files open read-only or not at all
    fprintf(stderr, "Error while opening file: %s (read-write refused: %zu object(s) outside the checkpoint layout, first: %s)\n", fname,
            h->file->c.lossy.size(), h->file->c.lossy[0].c_str());
    return nullptr;
  }
  return h.release();
}

MLF_OBJ* mlf_emgmm_c(double* X, int nX, int nY, int nC, double minProba, MLF_OBJ* handler) {
  return em_create(X, nX, nY, nC, minProba, handler, -1, 0);
}

MLF_OBJ* mlf_b200_emgmm_create(const double* X, int64_t nX, int nY, int nC, double minProba, MLF_OBJ* handler, int device,
                               unsigned flags) {
  return em_create(X, nX, nY, nC, minProba, handler, device, flags);
}

MLF_OBJ* mlf_b200_emgmm_create_multi(const double* X, int64_t nX, int nY, int nC, double minProba, MLF_OBJ* handler, const int* devices,
                                     int ndevices, unsigned flags) {
  std::vector<int> dv;
  if (ndevices <= 0 || !devices) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      fprintf(stderr, "mlf_b200_emgmm_create_multi: no CUDA device visible; mlfortran_b200 has no CPU fallback\n");
      return nullptr;
    }
    for (int i = 0; i < n; ++i) dv.push_back(i);
  } else {
    dv.assign(devices, devices + ndevices);
  }
  return em_create(X, nX, nY, nC, minProba, handler, -1, flags, &dv);
}

int mlf_b200_set_allreduce(MLF_OBJ* obj, mlf_b200_allreduce_fn fn, void* ctx, int rank, int world) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  if (o->multi) return -1;  // a single-process multi-GPU object does its own all-reduce
  o->eng.set_allreduce(fn, ctx, rank, world);
  o->rank = rank;
  o->world = world;
  return 0;
}

// ---- native NCCL: the statistics all-reduce as a plain ncclAllReduce on the engine's stream -----------------------
static int nccl_trampoline(void* ctx, double* buf, int64_t count, void* stream) {
  EmObject* o = static_cast<EmObject*>(ctx);
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!api || !o->nccl_comm) return -1;
  const int r = api->AllReduce(buf, buf, (size_t)count, 8 /* ncclFloat64 */, 0 /* ncclSum */, o->nccl_comm, stream);
  if (r != 0) fprintf(stderr, "mlfortran_b200: ncclAllReduce failed: %s\n", api->GetErrorString ? api->GetErrorString(r) : "?");
  return r == 0 ? 0 : -1;
}

int mlf_b200_nccl_unique_id(void* id128) {
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!api || !id128) return -1;
  return api->GetUniqueId(id128) == 0 ? 0 : mlf_OTHERERROR;
}

int mlf_b200_set_nccl_comm(MLF_OBJ* obj, void* comm) {
  EmObject* o = as_em(obj);
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!o || o->multi || !api || !comm) return -1;
  int rank = 0, world = 1;
  if (api->CommUserRank(comm, &rank) != 0 || api->CommCount(comm, &world) != 0) return mlf_OTHERERROR;
  if (o->nccl_comm && o->nccl_owned) api->CommDestroy(o->nccl_comm);
  o->nccl_comm = comm;
  o->nccl_owned = false;
  o->eng.set_allreduce(&nccl_trampoline, o, rank, world);
  o->rank = rank;
  o->world = world;
  return 0;
}

int mlf_b200_nccl_init(MLF_OBJ* obj, const void* id128, int rank, int world) {
  EmObject* o = as_em(obj);
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!o || o->multi || !api || !id128 || world < 1 || rank < 0 || rank >= world) return -1;
  mlfb200::NcclApi::UniqueId id;
  std::memcpy(id.b, id128, sizeof id.b);
  if (cudaSetDevice(o->eng.device()) != cudaSuccess) return mlf_OTHERERROR;
  void* comm = nullptr;
  const int r = api->CommInitRank(&comm, world, id, rank);
  if (r != 0) {
    fprintf(stderr, "mlfortran_b200: ncclCommInitRank failed: %s\n", api->GetErrorString ? api->GetErrorString(r) : "?");
    return mlf_OTHERERROR;
  }
  const int rc = mlf_b200_set_nccl_comm(obj, comm);
  if (rc != 0) { api->CommDestroy(comm); return rc; }
  o->nccl_owned = true;
  return 0;
}

int mlf_b200_set_initialized(MLF_OBJ* obj, int initialized) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  o->initialized = initialized != 0;
  return 0;
}
int mlf_b200_get_initialized(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return o ? (o->initialized ? 1 : 0) : -1;
}

int mlf_b200_set_cov_type(MLF_OBJ* obj, int cov_type) {
  EmObject* o = as_em(obj);
  if (!o || (cov_type != 0 && cov_type != 1)) return -1;
  const bool ok = o->multi ? o->multi->set_cov_diag(cov_type == 1) : o->eng.set_cov_diag(cov_type == 1);
  return ok ? 0 : -7;   // mlf_OTHERERROR: see mlf_b200_last_error
}

int mlf_b200_set_seed(MLF_OBJ* obj, uint64_t seed) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  o->seed = seed;
  return 0;
}

int mlf_b200_refresh_x(MLF_OBJ* obj, const double* X) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  o->X = X;
  return o->multi ? o->multi->refresh_x(X) : o->eng.refresh_x(X, o->x_on_device);
}

void* mlf_b200_host_alloc(size_t bytes, int write_combined) {
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (cudaHostAlloc(&p, bytes ? bytes : 1, flags) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void mlf_b200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

const char* mlf_b200_kernel_plan(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  if (!o) return "invalid handle";
  o->plan = o->multi ? o->multi->plan() : o->eng.plan();
  return o->plan.c_str();
}

const char* mlf_b200_last_error(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  if (!o) return "invalid handle";
  if (o->err.empty()) o->err = o->e_error();
  return o->err.c_str();
}

int64_t mlf_b200_sumll_trace(MLF_OBJ* obj, double* out, int64_t n) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  const int64_t m = (int64_t)o->trace.size();
  if (out)
    for (int64_t i = 0; i < m && i < n; ++i) out[i] = o->trace[(size_t)i];
  return m;
}

int mlf_b200_resident_proba(MLF_OBJ* obj, double* P) {
  EmObject* o = as_em(obj);
  if (!o || !P) return -1;
  return o->multi ? o->multi->resident_proba(P) : o->eng.resident_proba(P);
}

const double* mlf_b200_resident_proba_device(MLF_OBJ* obj, int64_t* ld) {
  EmObject* o = as_em(obj);
  if (!o) return nullptr;
  if (o->multi) return nullptr;  // sharded over several devices: use mlf_b200_resident_proba
  if (ld) *ld = o->eng.gamma_ld();
  return o->eng.gamma_device();
}

int mlf_b200_resident_labels(MLF_OBJ* obj, int* Cl) {
  EmObject* o = as_em(obj);
  if (!o || !Cl) return -1;
  if (!o->initialized) return mlf_UNINIT;
  int r = o->e_upload();
  if (r < 0) return r;
  return o->multi ? o->multi->resident_labels(Cl) : o->eng.resident_labels(Cl);
}

int mlf_b200_set_profile(MLF_OBJ* obj, int on) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  if (o->multi) { o->multi->set_profile(on != 0); o->multi->reset_times(); }
  else { o->eng.set_profile(on != 0); o->eng.reset_times(); }
  return 0;
}
int mlf_b200_get_times(MLF_OBJ* obj, double out[8]) {
  EmObject* o = as_em(obj);
  if (!o || !out) return -1;
  const mlfb200::EngineTimes t = o->multi ? o->multi->times() : o->eng.times();
  out[0] = t.refresh; out[1] = t.estep; out[2] = t.mid; out[3] = t.mstep; out[4] = t.fin; out[5] = (double)t.iters;
  out[6] = (double)t.sweeps; out[7] = (double)t.amb_pairs;
  return 0;
}
int mlf_b200_get_counters(MLF_OBJ* obj, int64_t out[4]) {
  EmObject* o = as_em(obj);
  if (!o || !out) return -1;
  const mlfb200::EngineTimes t = o->multi ? o->multi->times() : o->eng.times();
  out[0] = t.iters; out[1] = t.sweeps; out[2] = t.amb_pairs; out[3] = t.kept_pairs;
  return 0;
}
int mlf_b200_get_qf_stats(MLF_OBJ* obj, int64_t* qf_sweeps, double* bound) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  const mlfb200::EngineTimes t = o->multi ? o->multi->times() : o->eng.times();
  if (qf_sweeps) *qf_sweeps = t.qf_sweeps;
  if (bound) *bound = t.qf_bound;
  return 0;
}
int64_t mlf_b200_scatter_passes(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  if (!o) return -1;
  return (o->multi ? o->multi->times() : o->eng.times()).scatter_passes;
}
int64_t mlf_b200_launch_count(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return o ? (o->multi ? o->multi->launches() : o->eng.launches()) : -1;
}
void* mlf_b200_stream(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return (o && !o->multi) ? (void*)o->eng.stream() : nullptr;
}
int64_t mlf_b200_global_points(MLF_OBJ* obj) {
  EmObject* o = as_em(obj);
  return o ? (o->multi ? o->multi->n_global() : o->eng.n_global()) : -1;
}

int mlf_b200_h5_put(MLF_OBJ* handler, const char* name, int dtype, int rank, const int64_t* dims, const void* data, int override_) {
  Hdf5Handle* h = as_h5(handler);
  if (!h || !name || rank < 1 || rank > MLF_MAXRANK || !dims || !data) return -1;
  if (!h->writable()) return mlf_FILEERROR;
  if (dtype != mlf_INT64 && dtype != mlf_DOUBLE) return mlf_WRONGTYPE;
  mlfh5::Dataset nd;
  nd.name = name;
  nd.type = dtype == mlf_INT64 ? mlfh5::I64 : mlfh5::F64;
  for (int i = 0; i < rank; ++i) {
    if (dims[i] < 1) return mlf_WRONGRANK;
    nd.dims.push_back((uint64_t)dims[i]);
  }
  nd.data.assign((const uint8_t*)data, (const uint8_t*)data + nd.count() * 8);
  mlfh5::Dataset* dst = h->find(nd.name);
  if (dst && !override_) return mlf_FILEERROR;
  if (!h->put(std::move(nd))) return mlf_FILEERROR;
  return h->flush();
}

int mlf_b200_h5_get(MLF_OBJ* handler, const char* name, int* dtype, int* rank, int64_t* dims, void* data, int64_t capacity_bytes) {
  Hdf5Handle* h = as_h5(handler);
  if (!h || !name) return -1;
  mlfh5::Dataset* s = h->find(name);
  if (!s) return mlf_FILENOTFOUND;
  if (dtype) *dtype = s->type == mlfh5::I64 ? mlf_INT64 : mlf_DOUBLE;
  if (rank) *rank = (int)s->dims.size();
  if (dims)
    for (size_t i = 0; i < s->dims.size() && i < MLF_MAXRANK; ++i) dims[i] = (int64_t)s->dims[i];
  if (data) {
    if ((int64_t)s->data.size() > capacity_bytes) return mlf_WRONGRANK;
    std::memcpy(data, s->data.data(), s->data.size());
  }
  return 0;
}

// datasets below the handle (file: all of them; group handle: the ones inside the group), by path relative to it
int mlf_b200_h5_count(MLF_OBJ* handler) {
  Hdf5Handle* h = as_h5(handler);
  if (!h) return -1;
  int n = 0;
  for (const auto& d : h->file->c.ds) n += d.name.compare(0, h->prefix.size(), h->prefix) == 0;
  return n;
}
const char* mlf_b200_h5_name(MLF_OBJ* handler, int idx) {
  Hdf5Handle* h = as_h5(handler);
  if (!h || idx < 0) return nullptr;
  for (const auto& d : h->file->c.ds)
    if (d.name.compare(0, h->prefix.size(), h->prefix) == 0 && idx-- == 0) return d.name.c_str() + h->prefix.size();
  return nullptr;
}
int mlf_b200_h5_unsupported(MLF_OBJ* handler) {
  Hdf5Handle* h = as_h5(handler);
  return h ? (int)h->file->c.lossy.size() : -1;
}

// mlf_hdf5_openGroup / getSubGroup (src/mlf_hdf5.f90:106-151): a handler for the group `name` below this handler.
// mode 0 = open an existing group (H5Gopen_f), 1 = create it (H5Gcreate_f fails when it exists), 2 = open, else create
// (getSubGroup).  The group handle shares the file with its parent; mlf_pushState(group, obj, ...) writes the object's
// slots inside the group (Hdf5PushSubObject, :657-669); free it with mlf_dealloc.
MLF_OBJ* mlf_b200_h5_open_group(MLF_OBJ* handler, const char* name, int mode) {
  Hdf5Handle* h = as_h5(handler);
  if (!h || !name || !*name || std::strchr(name, '/') || !std::strcmp(name, ".") || !std::strcmp(name, "..")) return nullptr;
  const std::string path = h->prefix + name;
  const bool exists = h->file->has_group(path);
  if (h->file->find(path)) return nullptr;   // a dataset of that name
  if (mode == 0 && !exists) return nullptr;
  if (mode == 1 && exists) return nullptr;
  if (!exists) {
    if (!h->file->writable) return nullptr;
    h->file->c.groups.push_back(path);
    h->file->dirty = true;
    if (h->file->flush() < 0) return nullptr;
  }
  std::unique_ptr<Hdf5Handle> g(new Hdf5Handle());
  g->file = h->file;
  g->prefix = path + "/";
  return g.release();
}

// ---- engine-level entry points (bound by fortran/mlf_emgmm_b200.f90) ---------------------------------------------
namespace {
struct EngHandle {  // one device, or every listed device of the box inside this process
  std::unique_ptr<EmgmmEngine> single;
  std::unique_ptr<MultiEngine> multi;
  std::string err;
  void* nccl_comm = nullptr;   // mlf_b200_engine_nccl_init: owned
  ~EngHandle() {
    single.reset();            // the engine's stream is drained before the communicator goes
    if (nccl_comm)
      if (const mlfb200::NcclApi* api = mlfb200::nccl_api()) api->CommDestroy(nccl_comm);
  }
};
int eng_nccl_trampoline(void* ctx, double* buf, int64_t count, void* stream) {
  EngHandle* h = static_cast<EngHandle*>(ctx);
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!api || !h->nccl_comm) return -1;
  return api->AllReduce(buf, buf, (size_t)count, 8 /* ncclFloat64 */, 0 /* ncclSum */, h->nccl_comm, stream) == 0 ? 0 : -1;
}
EngHandle* as_eng(MLF_B200_ENGINE* p) { return static_cast<EngHandle*>(p); }
}  // namespace

MLF_B200_ENGINE* mlf_b200_engine_create(const double* X, int64_t nX, int nY, int nC, int device, unsigned flags) {
  if ((!X && nX > 0) || nX < 0 || nY < 1 || nC < 1) return nullptr;
  std::unique_ptr<EngHandle> h(new EngHandle());
  h->single.reset(new EmgmmEngine());
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
  if (h->single->init(X, (flags & MLF_B200_X_ON_DEVICE) != 0, nX, nY, nC, device, (flags & MLF_B200_COV_DIAG) != 0) < 0) {
    fprintf(stderr, "mlf_b200_engine_create: %s\n", h->single->last_error().c_str());
    return nullptr;
  }
  return h.release();
}
MLF_B200_ENGINE* mlf_b200_engine_create_multi(const double* X, int64_t nX, int nY, int nC, const int* devices, int ndevices, unsigned flags) {
  if ((!X && nX > 0) || nX < 0 || nY < 1 || nC < 1 || (flags & MLF_B200_X_ON_DEVICE)) return nullptr;
  std::vector<int> dv;
  if (ndevices <= 0 || !devices) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { fprintf(stderr, "mlf_b200_engine_create_multi: no CUDA device\n"); return nullptr; }
    for (int i = 0; i < n; ++i) dv.push_back(i);
  } else {
    dv.assign(devices, devices + ndevices);
  }
  std::unique_ptr<EngHandle> h(new EngHandle());
  h->multi.reset(new MultiEngine());
  if (h->multi->init(X, nX, nY, nC, dv, (flags & MLF_B200_COV_DIAG) != 0) < 0) {
    fprintf(stderr, "mlf_b200_engine_create_multi: %s\n", h->multi->last_error().c_str());
    return nullptr;
  }
  return h.release();
}
void mlf_b200_engine_destroy(MLF_B200_ENGINE* eng) { delete as_eng(eng); }
int mlf_b200_engine_step(MLF_B200_ENGINE* eng, int64_t niter, double minProba, double* Mu, double* Cov, double* lambda, double* sumLL) {
  EngHandle* e = as_eng(eng);
  if (!e || !Mu || !Cov || !lambda) return -1;
  if (niter <= 0) return 0;
  int r = e->multi ? e->multi->upload_params(Mu, Cov, lambda) : e->single->upload_params(Mu, Cov, lambda);
  if (r == 0) r = e->multi ? e->multi->iterate(niter, minProba, sumLL, nullptr) : e->single->iterate(niter, minProba, sumLL, nullptr);
  if (r == 0) r = e->multi ? e->multi->download_params(Mu, Cov, lambda) : e->single->download_params(Mu, Cov, lambda);
  return r;
}
int mlf_b200_engine_kmeans_init(MLF_B200_ENGINE* eng, int nkm, uint64_t seed, double* Mu, double* Cov, double* lambda) {
  EngHandle* e = as_eng(eng);
  if (!e || !Mu || !Cov || !lambda) return -1;
  int r = e->multi ? e->multi->kmeans_init(nkm, seed) : e->single->kmeans_init(nkm, seed);
  if (r == 0) r = e->multi ? e->multi->download_params(Mu, Cov, lambda) : e->single->download_params(Mu, Cov, lambda);
  return r;
}
int mlf_b200_engine_expectation(MLF_B200_ENGINE* eng, const double* Mu, const double* Cov, const double* lambda, const double* Xq,
                                int64_t nIn, double* P, int* Cl) {
  EngHandle* e = as_eng(eng);
  if (!e || !Mu || !Cov || !lambda || nIn < 0 || (!Xq && nIn > 0)) return -1;
  int r = e->multi ? e->multi->upload_params(Mu, Cov, lambda) : e->single->upload_params(Mu, Cov, lambda);
  if (r == 0) r = e->multi ? e->multi->expectation(Xq, nIn, P, Cl) : e->single->expectation(Xq, false, nIn, P, Cl);
  return r;
}
int mlf_b200_engine_proba(MLF_B200_ENGINE* eng, double* ProbaC) {
  EngHandle* e = as_eng(eng);
  if (!e || !ProbaC) return -1;
  return e->multi ? e->multi->resident_proba(ProbaC) : e->single->resident_proba(ProbaC);
}
int mlf_b200_engine_set_allreduce(MLF_B200_ENGINE* eng, mlf_b200_allreduce_fn fn, void* ctx, int rank, int world) {
  EngHandle* e = as_eng(eng);
  if (!e || e->multi) return -1;
  e->single->set_allreduce(fn, ctx, rank, world);
  return 0;
}
int mlf_b200_engine_nccl_init(MLF_B200_ENGINE* eng, const void* id128, int rank, int world) {
  EngHandle* e = as_eng(eng);
  const mlfb200::NcclApi* api = mlfb200::nccl_api();
  if (!e || e->multi || !api || !id128 || world < 1 || rank < 0 || rank >= world || e->nccl_comm) return -1;
  mlfb200::NcclApi::UniqueId id;
  std::memcpy(id.b, id128, sizeof id.b);
  if (cudaSetDevice(e->single->device()) != cudaSuccess) return mlf_OTHERERROR;
  void* comm = nullptr;
  if (api->CommInitRank(&comm, world, id, rank) != 0) return mlf_OTHERERROR;
  e->nccl_comm = comm;
  e->single->set_allreduce(&eng_nccl_trampoline, e, rank, world);
  return 0;
}
const char* mlf_b200_engine_last_error(MLF_B200_ENGINE* eng) {
  EngHandle* e = as_eng(eng);
  if (!e) return "invalid engine";
  e->err = e->multi ? e->multi->last_error() : e->single->last_error();
  return e->err.c_str();
}

// Single-component entry points (src/unsupervised/mlf_gaussian.f90:53-59,81-88): a temporary K = 1 engine over the
// caller's X runs the same kernels as the step object.
int mlf_evalGaussian(int ND, int NX, const double* X, double* P, const double* mu, const double* C, double lambda, double* sumLL) {
  if (ND < 1 || NX < 0 || !X || !mu || !C) return mlf_OTHERERROR;
  EmgmmEngine eng;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (eng.init(X, false, NX, ND, 1, dev) < 0) { fprintf(stderr, "mlf_evalGaussian: %s\n", eng.last_error().c_str()); return mlf_OTHERERROR; }
  if (eng.upload_params(mu, C, &lambda) < 0) return mlf_OTHERERROR;
  double ll = 0;
  if (eng.eval_gaussian_raw(P, &ll) < 0) return mlf_OTHERERROR;
  if (sumLL) *sumLL = ll;
  return 0;
}
double mlf_maxGaussian(int ND, int NX, const double* X, const double* Proba, double* mu, double* C) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  if (ND < 1 || NX < 0 || !X || !Proba || !mu || !C) return nan;
  EmgmmEngine eng;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (eng.init(X, false, NX, ND, 1, dev) < 0) { fprintf(stderr, "mlf_maxGaussian: %s\n", eng.last_error().c_str()); return nan; }
  double s = nan;
  if (eng.max_gaussian(Proba, mu, C, &s) < 0) return nan;
  return s;
}

}  // extern "C"
