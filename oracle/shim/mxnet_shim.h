// TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-ins for the MXNet / mshadow / dmlc-core / nnvm declarations that the reference's
// operator/multibox_{prior,target,detection}.cc need in order to compile *in place* from /root/reference
// (MXNet itself is not vendored in the reference tree and cannot be installed here).  The function templates
// mshadow::MultiBoxPriorForward / MultiBoxTargetForward / MultiBoxDetectionForward -- the reference's actual CPU
// algorithms -- are compiled verbatim from the reference sources; nothing of theirs is copied into this repo.
// What is NOT the reference's code in the resulting library: these type shims and the glue in ref_*.cc that plays
// the role of the operators' Forward() methods (output initialisation, the IoU plane, auto step, clip), which live
// in the -inl.h headers and depend on mshadow expression templates.
//
// Each ref_*.cc defines SHIM_PARAM / SHIM_OP / SHIM_PROP (the names its .cc refers to in the registration tail)
// and the include guard of the matching -inl.h so that the real header is skipped.
#ifndef ORACLE_SHIM_MXNET_SHIM_H_
#define ORACLE_SHIM_MXNET_SHIM_H_

#include <math.h>  // a CUDA-enabled MXNet build pulls <math.h> in (cuBLAS headers): unqualified exp() on float is expf

#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace shim {
struct Error : public std::runtime_error {
  explicit Error(const std::string &s) : std::runtime_error(s) {}
};
// dmlc-core's LogMessageFatal: collects the message and throws from its destructor.
struct Fatal {
  std::ostringstream os;
  Fatal(const char *file, int line, const char *what) { os << file << ":" << line << ": Check failed: " << what << " "; }
  std::ostringstream &stream() { return os; }
  ~Fatal() noexcept(false) { throw Error(os.str()); }
};
struct Registry {
  Registry &describe(const char *) { return *this; }
  Registry &add_argument(const char *, const char *, const char *) { return *this; }
  Registry &add_arguments(int) { return *this; }
};
}  // namespace shim

#define SHIM_CHECK_OP(a, b, op) \
  if (!((a)op(b))) ::shim::Fatal(__FILE__, __LINE__, #a " " #op " " #b).stream()
#define CHECK(x) \
  if (!(x)) ::shim::Fatal(__FILE__, __LINE__, #x).stream()
#define CHECK_EQ(a, b) SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) SHIM_CHECK_OP(a, b, !=)
#define CHECK_GE(a, b) SHIM_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) SHIM_CHECK_OP(a, b, >)
#define CHECK_LE(a, b) SHIM_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) SHIM_CHECK_OP(a, b, <)

namespace mshadow {
typedef unsigned index_t;
struct cpu {};
template <typename Device>
struct Stream {};

template <typename Device, int dim, typename DType>
struct Tensor {
  DType *dptr_;
  index_t shape_[dim];
  Stream<Device> *stream_;
  Tensor() : dptr_(nullptr), stream_(nullptr) {}
  Tensor(DType *p, std::initializer_list<index_t> s) : dptr_(p), stream_(nullptr) {
    int i = 0;
    for (index_t v : s) shape_[i++] = v;
  }
  index_t size(int i) const { return shape_[i]; }
  size_t stride0() const {
    size_t n = 1;
    for (int i = 1; i < dim; ++i) n *= shape_[i];
    return n;
  }
  Tensor<Device, dim - 1, DType> operator[](index_t i) const {
    Tensor<Device, dim - 1, DType> t;
    t.dptr_ = dptr_ + i * stride0();
    for (int k = 1; k < dim; ++k) t.shape_[k - 1] = shape_[k];
    t.stream_ = stream_;
    return t;
  }
};
template <typename Device, typename DType>
struct Tensor<Device, 1, DType> {
  DType *dptr_;
  index_t shape_[1];
  Stream<Device> *stream_;
  Tensor() : dptr_(nullptr), stream_(nullptr) {}
  index_t size(int) const { return shape_[0]; }
  DType &operator[](index_t i) const { return dptr_[i]; }
};
template <int dim, typename DType>
inline void Copy(Tensor<cpu, dim, DType> dst, const Tensor<cpu, dim, DType> &src, Stream<cpu> * = nullptr) {
  size_t n = 1;
  for (int i = 0; i < dim; ++i) n *= src.shape_[i];
  std::memcpy(dst.dptr_, src.dptr_, n * sizeof(DType));
}
}  // namespace mshadow

namespace nnvm {
template <typename T>
struct Tuple {
  std::vector<T> v;
  Tuple() {}
  Tuple(std::initializer_list<T> l) : v(l) {}
  unsigned ndim() const { return (unsigned)v.size(); }
  const T &operator[](size_t i) const { return v[i]; }
};
}  // namespace nnvm

// ---- just enough of the operator-registration surface for the tail of each .cc to compile ----
namespace mxnet {
using mshadow::cpu;
struct Context {};
struct TShape {};
class Operator {
 public:
  virtual ~Operator() {}
};
namespace op {
struct SHIM_PARAM {
  static int __FIELDS__() { return 0; }
};
template <typename xpu, typename DType>
class SHIM_OP : public Operator {
 public:
  explicit SHIM_OP(SHIM_PARAM) {}
};
template <typename xpu>
Operator *CreateOp(SHIM_PARAM param, int dtype);
class SHIM_PROP {
 public:
  bool InferShape(std::vector<TShape> *, std::vector<TShape> *, std::vector<TShape> *) const { return true; }
  bool InferType(std::vector<int> *, std::vector<int> *, std::vector<int> *) const { return true; }
  Operator *CreateOperatorEx(Context ctx, std::vector<TShape> *in_shape, std::vector<int> *in_type) const;
  SHIM_PARAM param_;
};
}  // namespace op
}  // namespace mxnet

#define MSHADOW_REAL_TYPE_SWITCH(type, DType, ...) \
  {                                                \
    typedef float DType;                           \
    (void)type;                                    \
    { __VA_ARGS__ }                                \
  }
#define DO_BIND_DISPATCH(Method, ...) return Method<cpu>(__VA_ARGS__)
#define SHIM_CAT_(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT_(a, b)
#define DMLC_REGISTER_PARAMETER(P) static int SHIM_CAT(shim_param_registered_, P) = 0
#define MXNET_REGISTER_OP_PROPERTY(name, Prop) static ::shim::Registry SHIM_CAT(shim_op_registered_, name) = ::shim::Registry()

#endif  // ORACLE_SHIM_MXNET_SHIM_H_
