// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/README.md).
// Nothing under genericschur.jl_b200/ may include, link or call this file.
//
// Scalar types the oracle is instantiated with:
//   double            <-> Julia Float64
//   DD  (hi,lo)       <-> a double-double type (DoubleFloats.Double64 / MultiFloats.Float64x2 layout:
//                         two consecutive Float64, high limb first)
//   MP  (MPFR 256)    <-> Julia BigFloat at setprecision(256)  (Julia's BigFloat *is* MPFR)
// and Cx<R> for Complex{R} with Julia's arithmetic conventions (4-multiplication product,
// Smith-style scaled division, hypot-based abs).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>

namespace gso {

// ---------------------------------------------------------------------------------------------
// double-double (host).  Algorithms: Dekker / Knuth error-free transformations, QD-style
// "accurate" addition, long division, Karp square root.
// ---------------------------------------------------------------------------------------------
struct DD {
    double hi, lo;
    DD() : hi(0.0), lo(0.0) {}
    DD(double h) : hi(h), lo(0.0) {}
    DD(int h) : hi((double)h), lo(0.0) {}
    DD(double h, double l) : hi(h), lo(l) {}
};

static inline void two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
static inline void quick_two_sum(double a, double b, double& s, double& e) {
    s = a + b;
    e = b - (s - a);
}
static inline void two_prod(double a, double b, double& p, double& e) {
    p = a * b;
    e = std::fma(a, b, -p);
}
static inline DD operator+(const DD& a, const DD& b) {
    double s1, s2, t1, t2;
    two_sum(a.hi, b.hi, s1, s2);
    if (!std::isfinite(s1)) return DD(s1, 0.0);
    two_sum(a.lo, b.lo, t1, t2);
    s2 += t1;
    quick_two_sum(s1, s2, s1, s2);
    s2 += t2;
    quick_two_sum(s1, s2, s1, s2);
    return DD(s1, s2);
}
static inline DD operator-(const DD& a) { return DD(-a.hi, -a.lo); }
static inline DD operator-(const DD& a, const DD& b) { return a + (-b); }
static inline DD operator*(const DD& a, const DD& b) {
    double p1, p2;
    two_prod(a.hi, b.hi, p1, p2);
    if (!std::isfinite(p1) || p1 == 0.0) return DD(p1, 0.0);
    p2 += a.hi * b.lo + a.lo * b.hi;
    quick_two_sum(p1, p2, p1, p2);
    return DD(p1, p2);
}
static inline DD operator/(const DD& a, const DD& b) {
    double q1 = a.hi / b.hi;
    if (!std::isfinite(q1) || q1 == 0.0) return DD(q1, 0.0);
    DD r = a - b * DD(q1);
    double q2 = r.hi / b.hi;
    r = r - b * DD(q2);
    double q3 = r.hi / b.hi;
    double s, e;
    quick_two_sum(q1, q2, s, e);
    DD q = DD(s, e) + DD(q3);
    return q;
}
static inline bool operator==(const DD& a, const DD& b) { return a.hi == b.hi && a.lo == b.lo; }
static inline bool operator!=(const DD& a, const DD& b) { return !(a == b); }
static inline bool operator<(const DD& a, const DD& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
static inline bool operator>(const DD& a, const DD& b) { return b < a; }
static inline bool operator<=(const DD& a, const DD& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
static inline bool operator>=(const DD& a, const DD& b) { return b <= a; }
static inline DD& operator+=(DD& a, const DD& b) { a = a + b; return a; }
static inline DD& operator-=(DD& a, const DD& b) { a = a - b; return a; }
static inline DD& operator*=(DD& a, const DD& b) { a = a * b; return a; }
static inline DD& operator/=(DD& a, const DD& b) { a = a / b; return a; }

static inline DD dd_sqrt(const DD& a) {
    if (a.hi <= 0.0) return (a.hi == 0.0) ? DD(0.0) : DD(std::numeric_limits<double>::quiet_NaN());
    if (!std::isfinite(a.hi)) return a;
    // Karp's trick: x = 1/sqrt(a) approx; sqrt(a) = a*x + (a - (a*x)^2) * x / 2
    double x = 1.0 / std::sqrt(a.hi);
    double ax = a.hi * x;
    DD axd(ax);
    DD r = a - axd * axd;
    double corr = r.hi * (x * 0.5);
    double s, e;
    two_sum(ax, corr, s, e);
    return DD(s, e);
}

// ---------------------------------------------------------------------------------------------
// MPFR 256-bit float, bound through hand-declared prototypes (the image ships libmpfr.so.6 but
// no headers).  Layout of __mpfr_struct for MPFR 4.x on LP64.
// ---------------------------------------------------------------------------------------------
extern "C" {
typedef struct {
    long _prec;
    int _sign;
    long _exp;
    unsigned long* _d;
} gso_mpfr_t;
void mpfr_init2(gso_mpfr_t*, long);
void mpfr_clear(gso_mpfr_t*);
int mpfr_set(gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_set_d(gso_mpfr_t*, double, int);
int mpfr_set_si(gso_mpfr_t*, long, int);
double mpfr_get_d(const gso_mpfr_t*, int);
int mpfr_add(gso_mpfr_t*, const gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_sub(gso_mpfr_t*, const gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_mul(gso_mpfr_t*, const gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_div(gso_mpfr_t*, const gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_sqrt(gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_neg(gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_abs(gso_mpfr_t*, const gso_mpfr_t*, int);
int mpfr_cmp(const gso_mpfr_t*, const gso_mpfr_t*);
int mpfr_nan_p(const gso_mpfr_t*);
int mpfr_inf_p(const gso_mpfr_t*);
int mpfr_zero_p(const gso_mpfr_t*);
int mpfr_sgn(const gso_mpfr_t*);
int mpfr_signbit(const gso_mpfr_t*);
int mpfr_add_d(gso_mpfr_t*, const gso_mpfr_t*, double, int);
int mpfr_set_ui_2exp(gso_mpfr_t*, unsigned long, long, int);
void mpfr_set_zero(gso_mpfr_t*, int);
}

#ifndef GSO_MP_PREC
#define GSO_MP_PREC 256
#endif

struct MP {
    gso_mpfr_t v[1];
    MP() { mpfr_init2(v, GSO_MP_PREC); mpfr_set_zero(v, 1); }
    MP(double d) { mpfr_init2(v, GSO_MP_PREC); mpfr_set_d(v, d, 0); }
    MP(int d) { mpfr_init2(v, GSO_MP_PREC); mpfr_set_si(v, d, 0); }
    MP(const MP& o) { mpfr_init2(v, GSO_MP_PREC); mpfr_set(v, o.v, 0); }
    MP(const DD& o) { mpfr_init2(v, GSO_MP_PREC); mpfr_set_d(v, o.hi, 0); mpfr_add_d(v, v, o.lo, 0); }
    MP& operator=(const MP& o) { if (this != &o) mpfr_set(v, o.v, 0); return *this; }
    ~MP() { mpfr_clear(v); }
    double to_double() const { return mpfr_get_d(v, 0); }
    DD to_dd() const {
        double h = mpfr_get_d(v, 0);
        if (!std::isfinite(h)) return DD(h, 0.0);
        MP r;
        MP hh(h);
        mpfr_sub(r.v, v, hh.v, 0);
        double l = mpfr_get_d(r.v, 0);
        double s, e;
        quick_two_sum(h, l, s, e);
        return DD(s, e);
    }
};
static inline MP operator+(const MP& a, const MP& b) { MP r; mpfr_add(r.v, a.v, b.v, 0); return r; }
static inline MP operator-(const MP& a, const MP& b) { MP r; mpfr_sub(r.v, a.v, b.v, 0); return r; }
static inline MP operator*(const MP& a, const MP& b) { MP r; mpfr_mul(r.v, a.v, b.v, 0); return r; }
static inline MP operator/(const MP& a, const MP& b) { MP r; mpfr_div(r.v, a.v, b.v, 0); return r; }
static inline MP operator-(const MP& a) { MP r; mpfr_neg(r.v, a.v, 0); return r; }
static inline bool mp_unordered(const MP& a, const MP& b) { return mpfr_nan_p(a.v) || mpfr_nan_p(b.v); }
static inline bool operator==(const MP& a, const MP& b) { return !mp_unordered(a, b) && mpfr_cmp(a.v, b.v) == 0; }
static inline bool operator!=(const MP& a, const MP& b) { return !(a == b); }
static inline bool operator<(const MP& a, const MP& b) { return !mp_unordered(a, b) && mpfr_cmp(a.v, b.v) < 0; }
static inline bool operator>(const MP& a, const MP& b) { return !mp_unordered(a, b) && mpfr_cmp(a.v, b.v) > 0; }
static inline bool operator<=(const MP& a, const MP& b) { return !mp_unordered(a, b) && mpfr_cmp(a.v, b.v) <= 0; }
static inline bool operator>=(const MP& a, const MP& b) { return !mp_unordered(a, b) && mpfr_cmp(a.v, b.v) >= 0; }
static inline MP& operator+=(MP& a, const MP& b) { mpfr_add(a.v, a.v, b.v, 0); return a; }
static inline MP& operator-=(MP& a, const MP& b) { mpfr_sub(a.v, a.v, b.v, 0); return a; }
static inline MP& operator*=(MP& a, const MP& b) { mpfr_mul(a.v, a.v, b.v, 0); return a; }
static inline MP& operator/=(MP& a, const MP& b) { mpfr_div(a.v, a.v, b.v, 0); return a; }

// ---------------------------------------------------------------------------------------------
// Real-scalar traits: the Julia functions eps / floatmin / floatmax / abs / sqrt / copysign /
// hypot / isnan for each instantiation.
// ---------------------------------------------------------------------------------------------
template <class R> struct RT;

template <> struct RT<double> {
    static double eps() { return 2.220446049250313e-16; }               // eps(Float64) = 2^-52
    static double floatmin() { return 2.2250738585072014e-308; }        // floatmin(Float64)
    static double floatmax() { return 1.7976931348623157e308; }         // floatmax(Float64)
    static double abs(double x) { return std::fabs(x); }
    static double sqrt(double x) { return std::sqrt(x); }
    static double copysign(double m, double s) { return std::copysign(m, s); }
    static double hypot(double a, double b) { return std::hypot(a, b); }
    static bool isnan(double x) { return std::isnan(x); }
    static double to_double(double x) { return x; }
};

template <> struct RT<DD> {
    // Conventions of the double-double type (stated, not verifiable here without Julia):
    //   eps      = 2^-104   (DoubleFloats: eps(Double64) = 4.93e-32)
    //   floatmin = 2^-969   (smallest value whose low limb is still a normal Float64)
    //   floatmax = floatmax(Float64)
    static DD eps() { return DD(std::ldexp(1.0, -104)); }
    static DD floatmin() { return DD(std::ldexp(1.0, -969)); }
    static DD floatmax() { return DD(1.7976931348623157e308, 0.0); }
    static DD abs(const DD& x) { return (x.hi < 0.0 || (x.hi == 0.0 && x.lo < 0.0)) ? -x : x; }
    static DD sqrt(const DD& x) { return dd_sqrt(x); }
    static DD copysign(const DD& m, const DD& s) {
        DD a = abs(m);
        return std::signbit(s.hi) ? -a : a;
    }
    static DD hypot(const DD& a, const DD& b) {
        DD aa = abs(a), ab = abs(b);
        DD w = (aa < ab) ? ab : aa, z = (aa < ab) ? aa : ab;
        if (w.hi == 0.0 || std::isnan(w.hi) || std::isnan(z.hi)) return w + z;
        DD q = z / w;
        return w * dd_sqrt(DD(1.0) + q * q);
    }
    static bool isnan(const DD& x) { return std::isnan(x.hi) || std::isnan(x.lo); }
    static double to_double(const DD& x) { return x.hi + x.lo; }
};

template <> struct RT<MP> {
    // eps(BigFloat) = 2^-(precision-1); floatmin(BigFloat) is "perverse" (test/real.jl:103) — a
    // fixed tiny power of two well inside MPFR's exponent range stands in for it.
    static MP eps() { MP r; mpfr_set_ui_2exp(r.v, 1, -(GSO_MP_PREC - 1), 0); return r; }
    static MP floatmin() { MP r; mpfr_set_ui_2exp(r.v, 1, -(1L << 40), 0); return r; }
    static MP floatmax() { MP r; mpfr_set_ui_2exp(r.v, 1, (1L << 40), 0); return r; }
    static MP abs(const MP& x) { MP r; mpfr_abs(r.v, x.v, 0); return r; }
    static MP sqrt(const MP& x) { MP r; mpfr_sqrt(r.v, x.v, 0); return r; }
    static MP copysign(const MP& m, const MP& s) {
        MP a = abs(m);
        return mpfr_signbit(s.v) ? -a : a;
    }
    static MP hypot(const MP& a, const MP& b) {
        MP aa = abs(a), ab = abs(b);
        MP w = (aa < ab) ? ab : aa, z = (aa < ab) ? aa : ab;
        if (mpfr_zero_p(w.v) || mpfr_nan_p(w.v) || mpfr_nan_p(z.v)) return w + z;
        MP q = z / w;
        return w * sqrt(MP(1.0) + q * q);
    }
    static bool isnan(const MP& x) { return mpfr_nan_p(x.v) != 0; }
    static double to_double(const MP& x) { return x.to_double(); }
};

template <class R> static inline R rmax(const R& a, const R& b) { return (a < b) ? b : a; }
template <class R> static inline R rmin(const R& a, const R& b) { return (b < a) ? b : a; }

// ---------------------------------------------------------------------------------------------
// Complex{R} with Julia's conventions.
// ---------------------------------------------------------------------------------------------
template <class R> struct Cx {
    R re, im;
    Cx() : re(0.0), im(0.0) {}
    Cx(const R& r) : re(r), im(0.0) {}
    Cx(const R& r, const R& i) : re(r), im(i) {}
};
template <class R> static inline Cx<R> operator+(const Cx<R>& a, const Cx<R>& b) { return Cx<R>(a.re + b.re, a.im + b.im); }
template <class R> static inline Cx<R> operator-(const Cx<R>& a, const Cx<R>& b) { return Cx<R>(a.re - b.re, a.im - b.im); }
template <class R> static inline Cx<R> operator-(const Cx<R>& a) { return Cx<R>(-a.re, -a.im); }
// Julia: *(z::Complex, w::Complex) = Complex(real(z)*real(w) - imag(z)*imag(w), real(z)*imag(w) + imag(z)*real(w))
template <class R> static inline Cx<R> operator*(const Cx<R>& a, const Cx<R>& b) {
    return Cx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> static inline Cx<R> operator*(const Cx<R>& a, const R& b) { return Cx<R>(a.re * b, a.im * b); }
template <class R> static inline Cx<R> operator*(const R& a, const Cx<R>& b) { return Cx<R>(a * b.re, a * b.im); }
template <class R> static inline Cx<R> operator/(const Cx<R>& a, const R& b) { return Cx<R>(a.re / b, a.im / b); }
// Smith's scaled complex division (what Julia's generic Complex `/` does).
template <class R> static inline Cx<R> operator/(const Cx<R>& a, const Cx<R>& b) {
    if (RT<R>::abs(b.re) >= RT<R>::abs(b.im)) {
        R r = b.im / b.re;
        R den = b.re + r * b.im;
        return Cx<R>((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    } else {
        R r = b.re / b.im;
        R den = b.im + r * b.re;
        return Cx<R>((a.re * r + a.im) / den, (a.im * r - a.re) / den);
    }
}
template <class R> static inline Cx<R>& operator+=(Cx<R>& a, const Cx<R>& b) { a = a + b; return a; }
template <class R> static inline Cx<R>& operator-=(Cx<R>& a, const Cx<R>& b) { a = a - b; return a; }
template <class R> static inline Cx<R>& operator*=(Cx<R>& a, const Cx<R>& b) { a = a * b; return a; }
template <class R> static inline Cx<R>& operator*=(Cx<R>& a, const R& b) { a = a * b; return a; }
template <class R> static inline Cx<R>& operator/=(Cx<R>& a, const R& b) { a = a / b; return a; }
template <class R> static inline bool operator==(const Cx<R>& a, const Cx<R>& b) { return a.re == b.re && a.im == b.im; }

template <class R> static inline Cx<R> conj_(const Cx<R>& a) { return Cx<R>(a.re, -a.im); }
template <class R> static inline R cabs(const Cx<R>& a) { return RT<R>::hypot(a.re, a.im); }
// sqrt(z::Complex): principal branch, computed as rho = sqrt((|z|+|x|)/2), eta = y/(2 rho).
template <class R> static inline Cx<R> csqrt(const Cx<R>& z) {
    R zero(0.0), two(2.0);
    if (z.re == zero && z.im == zero) return Cx<R>(zero, z.im);
    R m = rmax(RT<R>::abs(z.re), RT<R>::abs(z.im));
    R x = z.re / m, y = z.im / m;       // scaled copy: no over/underflow in the modulus
    R rho = RT<R>::sqrt((RT<R>::hypot(x, y) + RT<R>::abs(x)) / two);
    R sm = RT<R>::sqrt(m);
    R xi = rho, eta = (y / rho) / two;
    if (x < zero) {
        xi = RT<R>::abs(eta);
        eta = RT<R>::copysign(rho, y);
    }
    return Cx<R>(xi * sm, eta * sm);
}

// element-type helpers shared by real and complex code
static inline double conj_(double a) { return a; }
static inline DD conj_(const DD& a) { return a; }
static inline MP conj_(const MP& a) { return a; }

template <class T> struct ElemTraits;
template <> struct ElemTraits<double> { typedef double Real; static const bool is_complex = false; };
template <> struct ElemTraits<DD> { typedef DD Real; static const bool is_complex = false; };
template <> struct ElemTraits<MP> { typedef MP Real; static const bool is_complex = false; };
template <class R> struct ElemTraits<Cx<R>> { typedef R Real; static const bool is_complex = true; };

}  // namespace gso
