// TEST INFRASTRUCTURE ONLY -- instrumented scalar for counting the floating-point operations of the oracle's rhs!.
#pragma once
#include <cmath>
#include <cstdint>
namespace orc {
struct OpCounters { uint64_t add = 0, mul = 0, div = 0, fma = 0, sqrt_ = 0, log_ = 0, exp_ = 0, trig = 0, pow_ = 0, cmp = 0, absneg = 0; };
inline OpCounters& opc() { static thread_local OpCounters c; return c; }
struct OpCountD {
  double v;
  OpCountD() : v(0) {}
  OpCountD(double x) : v(x) {}
  OpCountD(int x) : v(x) {}
  OpCountD(long x) : v((double)x) {}
  OpCountD(long long x) : v((double)x) {}
  OpCountD(unsigned long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0; }
  OpCountD& operator+=(OpCountD o) { ++opc().add; v += o.v; return *this; }
  OpCountD& operator-=(OpCountD o) { ++opc().add; v -= o.v; return *this; }
  OpCountD& operator*=(OpCountD o) { ++opc().mul; v *= o.v; return *this; }
  OpCountD& operator/=(OpCountD o) { ++opc().div; v /= o.v; return *this; }
  OpCountD operator-() const { ++opc().absneg; return OpCountD(-v); }
  OpCountD operator+() const { return *this; }
};
#define ORC_BIN(op, ctr) \
  inline OpCountD operator op(OpCountD a, OpCountD b) { ++opc().ctr; return OpCountD(a.v op b.v); } \
  inline OpCountD operator op(OpCountD a, double b) { ++opc().ctr; return OpCountD(a.v op b); } \
  inline OpCountD operator op(double a, OpCountD b) { ++opc().ctr; return OpCountD(a op b.v); } \
  inline OpCountD operator op(OpCountD a, int b) { ++opc().ctr; return OpCountD(a.v op b); } \
  inline OpCountD operator op(int a, OpCountD b) { ++opc().ctr; return OpCountD(a op b.v); }
ORC_BIN(+, add) ORC_BIN(-, add) ORC_BIN(*, mul) ORC_BIN(/, div)
#undef ORC_BIN
#define ORC_CMP(op) \
  inline bool operator op(OpCountD a, OpCountD b) { ++opc().cmp; return a.v op b.v; } \
  inline bool operator op(OpCountD a, double b) { ++opc().cmp; return a.v op b; } \
  inline bool operator op(double a, OpCountD b) { ++opc().cmp; return a op b.v; } \
  inline bool operator op(OpCountD a, int b) { ++opc().cmp; return a.v op b; } \
  inline bool operator op(int a, OpCountD b) { ++opc().cmp; return a op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP
inline OpCountD sqrt(OpCountD a) { ++opc().sqrt_; return OpCountD(std::sqrt(a.v)); }
inline OpCountD log(OpCountD a) { ++opc().log_; return OpCountD(std::log(a.v)); }
inline OpCountD exp(OpCountD a) { ++opc().exp_; return OpCountD(std::exp(a.v)); }
inline OpCountD sin(OpCountD a) { ++opc().trig; return OpCountD(std::sin(a.v)); }
inline OpCountD cos(OpCountD a) { ++opc().trig; return OpCountD(std::cos(a.v)); }
inline OpCountD acos(OpCountD a) { ++opc().trig; return OpCountD(std::acos(a.v)); }
inline OpCountD atan2(OpCountD a, OpCountD b) { ++opc().trig; return OpCountD(std::atan2(a.v, b.v)); }
inline OpCountD pow(OpCountD a, OpCountD b) { ++opc().pow_; return OpCountD(std::pow(a.v, b.v)); }
inline OpCountD pow(OpCountD a, double b) { ++opc().pow_; return OpCountD(std::pow(a.v, b)); }
inline OpCountD pow(OpCountD a, int b) { ++opc().pow_; return OpCountD(std::pow(a.v, b)); }
inline OpCountD pow(double a, OpCountD b) { ++opc().pow_; return OpCountD(std::pow(a, b.v)); }
inline OpCountD fabs(OpCountD a) { ++opc().absneg; return OpCountD(std::fabs(a.v)); }
inline OpCountD abs(OpCountD a) { ++opc().absneg; return OpCountD(std::fabs(a.v)); }
inline OpCountD fmax(OpCountD a, OpCountD b) { ++opc().cmp; return a.v > b.v ? a : b; }
inline OpCountD fmin(OpCountD a, OpCountD b) { ++opc().cmp; return a.v < b.v ? a : b; }
inline bool isnan(OpCountD a) { return std::isnan(a.v); }
inline bool isfinite(OpCountD a) { return std::isfinite(a.v); }
inline OpCountD nextafter(OpCountD a, OpCountD b) { return OpCountD(std::nextafter(a.v, b.v)); }
inline OpCountD floor(OpCountD a) { return OpCountD(std::floor(a.v)); }
}  // namespace orc
namespace std {
using orc::sqrt; using orc::log; using orc::exp; using orc::sin; using orc::cos; using orc::acos; using orc::atan2;
using orc::pow; using orc::fabs; using orc::abs; using orc::isnan; using orc::isfinite; using orc::nextafter; using orc::floor;
inline orc::OpCountD max(orc::OpCountD a, double b) { ++orc::opc().cmp; return a.v > b ? a : orc::OpCountD(b); }
inline orc::OpCountD max(double a, orc::OpCountD b) { ++orc::opc().cmp; return a > b.v ? orc::OpCountD(a) : b; }
inline orc::OpCountD min(orc::OpCountD a, double b) { ++orc::opc().cmp; return a.v < b ? a : orc::OpCountD(b); }
inline orc::OpCountD min(double a, orc::OpCountD b) { ++orc::opc().cmp; return a < b.v ? orc::OpCountD(a) : b; }
}
