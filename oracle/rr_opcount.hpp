/* rr_opcount.hpp -- operation-counting scalar for rr_oracle_count.cpp (measurement infrastructure only). */
#pragma once
#include <cmath>
struct creal;
extern "C" { extern long long rr_ops_total, rr_ops_useful; }
struct creal {
  double v;
  creal() = default;
  creal(double x) : v(x) {}
  creal(float x) : v(x) {}
  creal(int x) : v(x) {}
  creal(long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  explicit operator bool() const { return v != 0; }
};
static inline void rr_count(double a, double b) { rr_ops_total++; if (a != 0.0 && b != 0.0) rr_ops_useful++; }
static inline void rr_count_add(double a, double b) { rr_ops_total++; if (a != 0.0 || b != 0.0) rr_ops_useful++; }
#define RR_BINOP(op, cnt) \
  static inline creal operator op(creal a, creal b) { cnt(a.v, b.v); return creal(a.v op b.v); } \
  static inline creal operator op(creal a, double b) { cnt(a.v, b); return creal(a.v op b); } \
  static inline creal operator op(double a, creal b) { cnt(a, b.v); return creal(a op b.v); } \
  static inline creal operator op(creal a, int b) { cnt(a.v, b); return creal(a.v op b); } \
  static inline creal operator op(int a, creal b) { cnt(a, b.v); return creal(a op b.v); } \
  static inline creal operator op(creal a, float b) { cnt(a.v, b); return creal(a.v op b); } \
  static inline creal operator op(float a, creal b) { cnt(a, b.v); return creal(a op b.v); }
RR_BINOP(+, rr_count_add)
RR_BINOP(-, rr_count_add)
RR_BINOP(*, rr_count)
RR_BINOP(/, rr_count)
static inline creal operator-(creal a) { return creal(-a.v); }
static inline creal &operator+=(creal &a, creal b) { a = a + b; return a; }
static inline creal &operator-=(creal &a, creal b) { a = a - b; return a; }
static inline creal &operator*=(creal &a, creal b) { a = a * b; return a; }
static inline creal &operator/=(creal &a, creal b) { a = a / b; return a; }
#define RR_CMP(op) \
  static inline bool operator op(creal a, creal b) { return a.v op b.v; } \
  static inline bool operator op(creal a, double b) { return a.v op b; } \
  static inline bool operator op(double a, creal b) { return a op b.v; } \
  static inline bool operator op(creal a, int b) { return a.v op b; } \
  static inline bool operator op(int a, creal b) { return a op b.v; }
RR_CMP(<) RR_CMP(>) RR_CMP(<=) RR_CMP(>=) RR_CMP(==) RR_CMP(!=)
