// oracle/stubs/Rcpp.h — TEST INFRASTRUCTURE ONLY.  This is NOT Rcpp.
//
// The handful of Rcpp names the MethyLasso reference sources use (List / DataFrame with named
// elements, _["name"] = value, as<>/wrap, Rcout, stop, checkUserInterrupt), so that the reference's
// unmodified sources compile in place into oracle/_ref/libmethylasso_ref.so without R.  Values are
// kept as R would keep them: integer, double and logical vectors, strings and lists.
#ifndef ORACLE_STUB_RCPP_H
#define ORACLE_STUB_RCPP_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include <Eigen/Core>

namespace Rcpp {

class exception : public std::runtime_error {
 public:
  explicit exception(const std::string& m) : std::runtime_error(m) {}
};
inline void stop(const std::string& m) { throw exception(m); }
inline void checkUserInterrupt() {}

// Rcout: silent unless ORACLE_REF_VERBOSE is set in the environment
class NullBuf : public std::streambuf { int overflow(int c) override { return c; } };
inline std::ostream& rcout_stream() {
  static NullBuf nb;
  static std::ostream null_os(&nb);
  static const bool verbose = std::getenv("ORACLE_REF_VERBOSE") != nullptr;
  return verbose ? std::cerr : null_os;
}
static std::ostream& Rcout = rcout_stream();

struct Value;
typedef std::shared_ptr<Value> RObject;
struct Value {
  enum Kind { NIL, INT, REAL, LGL, STR, LIST } kind = NIL;
  std::vector<int> iv;          // INT, LGL
  std::vector<double> dv;       // REAL
  std::string s;                // STR
  std::vector<std::pair<std::string, RObject> > items;   // LIST
};
inline RObject make(Value::Kind k) { RObject r = std::make_shared<Value>(); r->kind = k; return r; }

class List;
class Proxy;

// ---- wrap ---------------------------------------------------------------------------------------
inline RObject wrap(const RObject& r) { return r; }
inline RObject wrap(double v) { RObject r = make(Value::REAL); r->dv.push_back(v); return r; }
inline RObject wrap(int v) { RObject r = make(Value::INT); r->iv.push_back(v); return r; }
inline RObject wrap(unsigned v) { RObject r = make(Value::REAL); r->dv.push_back(v); return r; }   // Rcpp wraps unsigned as numeric
inline RObject wrap(bool v) { RObject r = make(Value::LGL); r->iv.push_back(v ? 1 : 0); return r; }
inline RObject wrap(const char* v) { RObject r = make(Value::STR); r->s = v; return r; }
inline RObject wrap(const std::string& v) { RObject r = make(Value::STR); r->s = v; return r; }
inline RObject wrap(const Eigen::ArrayXd& v) { RObject r = make(Value::REAL); r->dv.assign(v.data(), v.data() + v.size()); return r; }
inline RObject wrap(const Eigen::ArrayXi& v) { RObject r = make(Value::INT); r->iv.assign(v.data(), v.data() + v.size()); return r; }
inline RObject wrap(const std::vector<double>& v) { RObject r = make(Value::REAL); r->dv = v; return r; }
inline RObject wrap(const std::vector<int>& v) { RObject r = make(Value::INT); r->iv = v; return r; }
inline RObject wrap(const std::vector<unsigned>& v) { RObject r = make(Value::REAL); r->dv.assign(v.begin(), v.end()); return r; }
RObject wrap(const List& l);
RObject wrap(const Proxy& p);

// ---- as -----------------------------------------------------------------------------------------
template <class T> struct As;
template <> struct As<double> { static double get(const RObject& r) {
  if (r->kind == Value::REAL && r->dv.size() == 1) return r->dv[0];
  if ((r->kind == Value::INT || r->kind == Value::LGL) && r->iv.size() == 1) return r->iv[0];
  throw exception("stub Rcpp: expecting a single numeric value"); } };
template <> struct As<int> { static int get(const RObject& r) { return (int)As<double>::get(r); } };
template <> struct As<unsigned> { static unsigned get(const RObject& r) { return (unsigned)As<double>::get(r); } };
template <> struct As<long> { static long get(const RObject& r) { return (long)As<double>::get(r); } };
template <> struct As<bool> { static bool get(const RObject& r) { return As<double>::get(r) != 0; } };
template <> struct As<std::string> { static std::string get(const RObject& r) {
  if (r->kind != Value::STR) throw exception("stub Rcpp: expecting a string"); return r->s; } };
template <> struct As<Eigen::ArrayXd> { static Eigen::ArrayXd get(const RObject& r) {
  if (r->kind == Value::REAL) { Eigen::ArrayXd a(r->dv.size()); for (size_t i = 0; i < r->dv.size(); ++i) a(i) = r->dv[i]; return a; }
  if (r->kind == Value::INT || r->kind == Value::LGL) { Eigen::ArrayXd a(r->iv.size()); for (size_t i = 0; i < r->iv.size(); ++i) a(i) = r->iv[i]; return a; }
  throw exception("stub Rcpp: expecting a numeric vector"); } };
template <> struct As<Eigen::ArrayXi> { static Eigen::ArrayXi get(const RObject& r) {
  if (r->kind == Value::INT || r->kind == Value::LGL) { Eigen::ArrayXi a(r->iv.size()); for (size_t i = 0; i < r->iv.size(); ++i) a(i) = r->iv[i]; return a; }
  if (r->kind == Value::REAL) { Eigen::ArrayXi a(r->dv.size()); for (size_t i = 0; i < r->dv.size(); ++i) a(i) = (int)r->dv[i]; return a; }
  throw exception("stub Rcpp: expecting an integer vector"); } };
template <class T> T as(const RObject& r) { return As<typename std::remove_const<T>::type>::get(r); }

// ---- named arguments ------------------------------------------------------------------------------
struct NamedValue { std::string name; RObject value; };
class Named {
 public:
  explicit Named(const std::string& n) : name_(n) {}
  template <class T> NamedValue operator=(const T& v) const { return NamedValue{name_, wrap(v)}; }
 private:
  std::string name_;
};
struct NamedPlaceHolder { Named operator[](const std::string& n) const { return Named(n); } };
static NamedPlaceHolder _;

// ---- List / DataFrame -----------------------------------------------------------------------------
class Proxy {
 public:
  Proxy(const RObject& list, const std::string& name) : list_(list), name_(name) {}
  RObject get() const {
    for (const auto& kv : list_->items) if (kv.first == name_) return kv.second;
    throw exception("stub Rcpp: no element named '" + name_ + "'");
  }
  template <class T> operator T() const { return as<T>(get()); }
  template <class T> Proxy& operator=(const T& v) {
    for (auto& kv : list_->items) if (kv.first == name_) { kv.second = wrap(v); return *this; }
    list_->items.push_back(std::make_pair(name_, wrap(v)));
    return *this;
  }
 private:
  RObject list_; std::string name_;
};
inline RObject wrap(const Proxy& p) { return p.get(); }

class List {
 public:
  List() : v_(make(Value::LIST)) {}
  explicit List(const RObject& r) : v_(r) { if (r->kind != Value::LIST) throw exception("stub Rcpp: not a list"); }
  static List create() { return List(); }
  template <class... A> static List create(const NamedValue& a, const A&... rest) {
    List l = create(rest...);
    l.v_->items.insert(l.v_->items.begin(), std::make_pair(a.name, a.value));
    return l;
  }
  Proxy operator[](const std::string& n) const { return Proxy(v_, n); }
  bool containsElementNamed(const char* n) const {
    for (const auto& kv : v_->items) if (kv.first == n) return true;
    return false;
  }
  int size() const { return (int)v_->items.size(); }
  const RObject& robject() const { return v_; }
 protected:
  RObject v_;
};
inline RObject wrap(const List& l) { return l.robject(); }

class DataFrame : public List {
 public:
  DataFrame() {}
  DataFrame(const List& l) : List(l) {}
  template <class... A> static DataFrame create(const A&... a) { return DataFrame(List::create(a...)); }
  int nrows() const {
    if (v_->items.empty()) return 0;
    const RObject& c = v_->items[0].second;
    return (int)(c->kind == Value::REAL ? c->dv.size() : c->iv.size());
  }
};
template <> struct As<List> { static List get(const RObject& r) { return List(r); } };
template <> struct As<DataFrame> { static DataFrame get(const RObject& r) { return DataFrame(List(r)); } };

template <class T> T as(const Proxy& p) { return as<T>(p.get()); }

}  // namespace Rcpp
#endif
