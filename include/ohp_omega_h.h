/*
 * ohp_omega_h.h -- the slice of the Omega_h API that ex12's mesh adapter touches
 * (ex12.cpp:528-943), backed by libohp.so's `.osh` reader.  Enough for CreateQuadMesh's
 * read / ask_elem_verts / coords / get_array / HostRead calls to compile against the same names:
 *   Omega_h::Library (ex12.cpp:1494), Omega_h::Mesh (:797), binary::read (:813,:818),
 *   Mesh::dim/nverts/nelems/ask_elem_verts/coords/get_array/has_tag (:591-594,:627-634,:703-705,:784),
 *   Read<T>, HostRead<T> (:617,:670,:768-777), LO / GO / Real / I8.
 * Arrays live in host memory here (the reference copies them to the host anyway before handing
 * them to PETSc, ex12.cpp:860-861); device-side work goes through ohp_picpart_extract.
 */
#ifndef OHP_OMEGA_H_H
#define OHP_OMEGA_H_H

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ohp.h"

namespace Omega_h {
typedef int32_t LO;
typedef int64_t GO;
typedef double Real;
typedef int8_t I8;
enum { VERT = 0, EDGE = 1, FACE = 2, REGION = 3 };

template <class T> class Read {
 public:
  Read() : d_(std::make_shared<std::vector<T>>()) {}
  Read(const T *p, size_t n) : d_(std::make_shared<std::vector<T>>(p, p + n)) {}
  LO size() const { return (LO)d_->size(); }
  const T &operator[](LO i) const { return (*d_)[i]; }
  const T *data() const { return d_->data(); }
 private:
  std::shared_ptr<std::vector<T>> d_;
};
typedef Read<LO> LOs;
typedef Read<GO> GOs;
typedef Read<Real> Reals;
template <class T> class HostRead {
 public:
  HostRead() {}
  HostRead(Read<T> r) : r_(r) {}
  LO size() const { return r_.size(); }
  const T &operator[](LO i) const { return r_[i]; }
  const T *data() const { return r_.data(); }
 private:
  Read<T> r_;
};

class Comm {};
class Library {
 public:
  Library(int * = nullptr, char *** = nullptr) {}
  Comm *world() { return &c_; }
  Comm *self() { return &c_; }
 private:
  Comm c_;
};

class Mesh {
 public:
  explicit Mesh(Library * = nullptr) {}
  ~Mesh() { if (h_) ohp_osh_free(h_); }
  Mesh(const Mesh &) = delete;
  Mesh &operator=(const Mesh &) = delete;
  void load(const std::string &path, int part) {
    if (h_) { ohp_osh_free(h_); h_ = nullptr; }
    if (ohp_osh_read(path.c_str(), part, &h_)) throw std::runtime_error(ohp_last_error());
    ohp_osh_info(h_, &dim_, &nverts_, &nelems_, nullptr, nullptr);
  }
  int dim() const { return dim_; }
  LO nverts() const { return nverts_; }
  LO nelems() const { return nelems_; }
  LO nents(int d) const { int n = 0; ohp_osh_nents(h_, d, &n); return n; }
  LOs ask_elem_verts() const {
    const int32_t *ev = nullptr;
    ohp_osh_elem_verts(h_, &ev);
    return LOs(ev, (size_t)nelems_ * (dim_ + 1));
  }
  Reals coords() const {
    const double *x = nullptr;
    ohp_osh_coords(h_, &x);
    return Reals(x, (size_t)nverts_ * dim_);
  }
  bool has_tag(int d, const std::string &name) const { return ohp_osh_tag(h_, d, name.c_str(), nullptr, nullptr, nullptr, nullptr) == 0; }
  template <class T> Read<T> get_array(int d, const std::string &name) const {
    int nc = 0, type = 0; int64_t count = 0; const void *data = nullptr;
    if (ohp_osh_tag(h_, d, name.c_str(), &nc, &type, &count, &data)) throw std::runtime_error(ohp_last_error());
    const size_t want = sizeof(T);
    const size_t have = type == 0 ? 1 : type == 2 ? 4 : 8;
    if (want != have) throw std::runtime_error("Omega_h::Mesh::get_array: tag '" + name + "' has a different type");
    return Read<T>((const T *)data, (size_t)count);
  }
  void balance() {} /* partitions come from picpart files or the strip partitioner (north_star) */
 private:
  ohp_osh h_ = nullptr;
  int dim_ = 0, nverts_ = 0, nelems_ = 0;
};

namespace binary {
/* Omega_h::binary::read(path, comm, &mesh, strict) (ex12.cpp:813,818): a serial directory is read
 * as part 0; every picpart file is its own one-part directory (ex12.cpp:810-813) */
inline void read(const std::string &path, Comm *, Mesh *mesh, bool = true) { mesh->load(path, 0); }
}  // namespace binary
}  // namespace Omega_h
#endif
