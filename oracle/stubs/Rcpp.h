/* TEST INFRASTRUCTURE — mock of the small slice of Rcpp that the four hot-path glue
 * files use, so that they compile UNMODIFIED from /root/reference into oracle/_ref/:
 *   image.CornerDetectionHarris/src/rcpp_harris.cpp:19-59   (NumericVector, List::create, Named, SEXP)
 *   image.CannyEdges/src/rcpp_canny.cpp:122-245             (IntegerVector, NumericMatrix(Dimension), _["x"]=)
 *   image.dlib/src/rcpp_fhog.cpp:10-46, rcpp_surf.cpp:10-54 (NumericVector(n), NumericMatrix(n,64), op())
 * Values are held as doubles (as R would hold them); a List is an ordered name->Value map.
 * Not Rcpp, not product code.
 */
#ifndef ORACLE_STUB_RCPP_H
#define ORACLE_STUB_RCPP_H
#include <vector>
#include <string>
#include <cstddef>
#include <stdexcept>
#include <cstdlib>
#include <initializer_list>

struct SEXPREC { virtual ~SEXPREC() {} };
typedef SEXPREC *SEXP;

namespace Rcpp {

inline void stop(const std::string &msg) { throw std::runtime_error(msg); }

struct Dimension { long a, b; Dimension(long a_, long b_) : a(a_), b(b_) {} };

template <typename T> struct VecT {
  std::vector<T> d;
  VecT() {}
  explicit VecT(long n) : d((size_t)n, T(0)) {}
  VecT(const T *p, size_t n) : d(p, p + n) {}
  long size() const { return (long)d.size(); }
  T &operator[](long i) { return d[(size_t)i]; }
  const T &operator[](long i) const { return d[(size_t)i]; }
};
typedef VecT<double> NumericVector;
typedef VecT<int> IntegerVector;

struct NumericMatrix {
  std::vector<double> d; long nr, nc;
  NumericMatrix(long r, long c) : d((size_t)(r * c), 0.0), nr(r), nc(c) {}
  explicit NumericMatrix(const Dimension &dim) : d((size_t)(dim.a * dim.b), 0.0), nr(dim.a), nc(dim.b) {}
  double &operator[](long i) { return d[(size_t)i]; }
  double &operator()(long i, long j) { return d[(size_t)(i + j * nr)]; }   /* column-major, like R */
};

struct Value {
  std::string name; std::vector<double> data; long nr = -1, nc = -1;
};

struct NamedProxy {
  std::string name;
  explicit NamedProxy(const std::string &n) : name(n) {}
  template <typename T> Value scalar(T v) const { Value x; x.name = name; x.data.push_back((double)v); return x; }
  Value operator=(int v) const { return scalar(v); }
  Value operator=(long v) const { return scalar(v); }
  Value operator=(unsigned long v) const { return scalar(v); }
  Value operator=(double v) const { return scalar(v); }
  Value operator=(bool v) const { return scalar(v); }
  template <typename T> Value operator=(const std::vector<T> &v) const {
    Value x; x.name = name; x.data.assign(v.begin(), v.end()); return x; }
  template <typename T> Value operator=(const VecT<T> &v) const {
    Value x; x.name = name; x.data.assign(v.d.begin(), v.d.end()); return x; }
  Value operator=(const NumericMatrix &m) const {
    Value x; x.name = name; x.data = m.d; x.nr = m.nr; x.nc = m.nc; return x; }
};
inline NamedProxy Named(const std::string &n) { return NamedProxy(n); }
struct UnderscoreT { NamedProxy operator[](const char *n) const { return NamedProxy(n); } };
static const UnderscoreT _ = UnderscoreT();

struct List : public SEXPREC {
  std::vector<Value> items;
  template <typename... A> static List create(const A &... a) { List l; (void)std::initializer_list<int>{(l.items.push_back(a), 0)...}; return l; }
  const Value &get(const std::string &n) const {
    for (size_t i = 0; i < items.size(); i++) if (items[i].name == n) return items[i];
    throw std::runtime_error("List: no element " + n);
  }
  operator SEXP() const { return new List(*this); }
};

} // namespace Rcpp
#endif
