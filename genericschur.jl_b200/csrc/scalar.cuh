// Device scalar types of the Schur kernels: Float64, an in-kernel double-double ("Float64x2"), and
// Complex of either.  Arithmetic conventions follow what the reference's generic Julia code gets from
// its element types (src/GenericSchur.jl is generic over T<:AbstractFloat / Complex{T}): 4-multiply
// complex product, scaled complex division, hypot-based modulus.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gs {

#define GS_DEV __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// double-double.  Error-free transformations with explicit round-to-nearest intrinsics so that the
// compiler can neither contract nor reassociate them.
// ------------------------------------------------------------------------------------------------
struct __align__(16) dd_t {
    double hi, lo;
};

GS_DEV dd_t mk_dd(double h, double l = 0.0) {
    dd_t r;
    r.hi = h;
    r.lo = l;
    return r;
}
GS_DEV void two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
GS_DEV void quick_two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    e = __dsub_rn(b, __dsub_rn(s, a));
}
GS_DEV void two_prod(double a, double b, double& p, double& e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
// IEEE-style (accurate) addition: 20 flops
GS_DEV dd_t operator+(const dd_t& a, const dd_t& b) {
    double s1, s2, t1, t2;
    two_sum(a.hi, b.hi, s1, s2);
    two_sum(a.lo, b.lo, t1, t2);
    s2 = __dadd_rn(s2, t1);
    quick_two_sum(s1, s2, s1, s2);
    s2 = __dadd_rn(s2, t2);
    quick_two_sum(s1, s2, s1, s2);
    // inf/nan in the leading sum would otherwise poison lo with nan and hide an honest inf
    if (!isfinite(s1)) s2 = 0.0;
    return mk_dd(s1, s2);
}
GS_DEV dd_t operator-(const dd_t& a) { return mk_dd(-a.hi, -a.lo); }
GS_DEV dd_t operator-(const dd_t& a, const dd_t& b) { return a + (-b); }
GS_DEV dd_t operator*(const dd_t& a, const dd_t& b) {
    double p1, p2;
    two_prod(a.hi, b.hi, p1, p2);
    p2 = __fma_rn(a.hi, b.lo, p2);
    p2 = __fma_rn(a.lo, b.hi, p2);
    quick_two_sum(p1, p2, p1, p2);
    if (!isfinite(p1) || p1 == 0.0) p2 = 0.0;
    return mk_dd(p1, p2);
}
GS_DEV dd_t dd_mul_d(const dd_t& a, double b) {
    double p1, p2;
    two_prod(a.hi, b, p1, p2);
    p2 = __fma_rn(a.lo, b, p2);
    quick_two_sum(p1, p2, p1, p2);
    if (!isfinite(p1) || p1 == 0.0) p2 = 0.0;
    return mk_dd(p1, p2);
}
GS_DEV dd_t operator/(const dd_t& a, const dd_t& b) {
    double q1 = __ddiv_rn(a.hi, b.hi);
    if (!isfinite(q1) || q1 == 0.0) return mk_dd(q1, 0.0);
    dd_t r = a - dd_mul_d(b, q1);
    double q2 = __ddiv_rn(r.hi, b.hi);
    r = r - dd_mul_d(b, q2);
    double q3 = __ddiv_rn(r.hi, b.hi);
    double s, e;
    quick_two_sum(q1, q2, s, e);
    return mk_dd(s, e) + mk_dd(q3);
}
GS_DEV bool operator==(const dd_t& a, const dd_t& b) { return a.hi == b.hi && a.lo == b.lo; }
GS_DEV bool operator!=(const dd_t& a, const dd_t& b) { return !(a == b); }
GS_DEV bool operator<(const dd_t& a, const dd_t& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
GS_DEV bool operator>(const dd_t& a, const dd_t& b) { return b < a; }
GS_DEV bool operator<=(const dd_t& a, const dd_t& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
GS_DEV bool operator>=(const dd_t& a, const dd_t& b) { return b <= a; }
GS_DEV dd_t& operator+=(dd_t& a, const dd_t& b) { a = a + b; return a; }
GS_DEV dd_t& operator-=(dd_t& a, const dd_t& b) { a = a - b; return a; }
GS_DEV dd_t& operator*=(dd_t& a, const dd_t& b) { a = a * b; return a; }
GS_DEV dd_t& operator/=(dd_t& a, const dd_t& b) { a = a / b; return a; }

GS_DEV dd_t dd_sqrt(const dd_t& a) {
    if (!(a.hi > 0.0)) return (a.hi == 0.0) ? mk_dd(0.0) : mk_dd(nan(""));
    if (!isfinite(a.hi)) return a;
    // Karp: x ~ 1/sqrt(a);  sqrt(a) ~ a x + (a - (a x)^2) x / 2
    double x = __drcp_rn(__dsqrt_rn(a.hi));
    double ax = __dmul_rn(a.hi, x);
    dd_t axd = mk_dd(ax);
    dd_t r = a - axd * axd;
    double corr = __dmul_rn(r.hi, __dmul_rn(x, 0.5));
    double s, e;
    two_sum(ax, corr, s, e);
    return mk_dd(s, e);
}

// ------------------------------------------------------------------------------------------------
// real-scalar traits and the handful of real functions the algorithm uses
// ------------------------------------------------------------------------------------------------
template <class R> struct rtraits;
template <> struct rtraits<double> {
    GS_DEV static double eps() { return 2.220446049250313e-16; }
    GS_DEV static double floatmin() { return 2.2250738585072014e-308; }
    GS_DEV static double floatmax() { return 1.7976931348623157e308; }
    GS_DEV static double from(double x) { return x; }
    static constexpr int ndoubles = 1;
};
template <> struct rtraits<dd_t> {
    // eps = 2^-104, floatmin = 2^-969 (low limb stays a normal Float64), floatmax = floatmax(Float64)
    GS_DEV static dd_t eps() { return mk_dd(4.930380657631324e-32); }
    GS_DEV static dd_t floatmin() { return mk_dd(2.004168360008973e-292); }
    GS_DEV static dd_t floatmax() { return mk_dd(1.7976931348623157e308); }
    GS_DEV static dd_t from(double x) { return mk_dd(x); }
    static constexpr int ndoubles = 2;
};

GS_DEV double r_abs(double x) { return fabs(x); }
GS_DEV dd_t r_abs(const dd_t& x) { return (x.hi < 0.0 || (x.hi == 0.0 && x.lo < 0.0)) ? -x : x; }
GS_DEV double r_sqrt(double x) { return __dsqrt_rn(x); }
GS_DEV dd_t r_sqrt(const dd_t& x) { return dd_sqrt(x); }
GS_DEV bool r_signbit(double x) { return signbit(x); }
GS_DEV bool r_signbit(const dd_t& x) { return signbit(x.hi); }
GS_DEV double r_hi(double x) { return x; }
GS_DEV double r_hi(const dd_t& x) { return x.hi; }
GS_DEV bool r_isnan(double x) { return isnan(x); }
GS_DEV bool r_isnan(const dd_t& x) { return isnan(x.hi) || isnan(x.lo); }
template <class R> GS_DEV R r_max(const R& a, const R& b) { return (a < b) ? b : a; }
template <class R> GS_DEV R r_min(const R& a, const R& b) { return (b < a) ? b : a; }
template <class R> GS_DEV R r_copysign(const R& m, const R& s) {
    R a = r_abs(m);
    return r_signbit(s) ? -a : a;
}
template <class R> GS_DEV R r_const(double x) { return rtraits<R>::from(x); }
// safemin, src/util.jl:5-12 : for both types 1/floatmax < floatmin so safemin == floatmin
template <class R> GS_DEV R r_safemin() { return rtraits<R>::floatmin(); }

// sqrt(a^2 + b^2 + c^2 + d^2) without intermediate over/underflow (dlapy3-style, src/util.jl:562-570)
template <class R> GS_DEV R r_hypot4(const R& a, const R& b, const R& c, const R& d) {
    R aa = r_abs(a), ab = r_abs(b), ac = r_abs(c), ad = r_abs(d);
    R w = r_max(r_max(aa, ab), r_max(ac, ad));
    if (w == r_const<R>(0.0) || r_isnan(w)) return w + aa + ab + ac + ad;   // 0, or propagate nan
    R rw = r_const<R>(1.0) / w;
    R x = aa * rw, y = ab * rw, z = ac * rw, t = ad * rw;
    return w * r_sqrt(x * x + y * y + z * z + t * t);
}
template <class R> GS_DEV R r_hypot(const R& a, const R& b) {
    return r_hypot4(a, b, r_const<R>(0.0), r_const<R>(0.0));
}

// ------------------------------------------------------------------------------------------------
// Fast Float64 primitives for the reflector on the serial critical path of the QR sweeps.
// IEEE division / sqrt cost ~25-30 dependent instructions each (with slow-path branches); the
// reflector needs one sqrt and two reciprocals per bulge step.  Operands are first scaled by an exact
// power of two so that they are O(1); then MUFU seeds + Newton steps give results within ~1 ulp with
// no special-case branches.
// ------------------------------------------------------------------------------------------------
GS_DEV double fast_rcp(double a) {   // a finite, normal, non-zero
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    return y;
}
GS_DEV double fast_sqrt(double a) {  // a in a safe range (no denormals / overflow in a*y)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    // y <- y (1 + e/2 + 3 e^2 / 8),  e = 1 - a y^2   (one cubic step: 20 -> 60+ bits)
    double ay = a * y;
    double e = fma(-ay, y, 1.0);
    double c = fma(e, 0.375, 0.5);
    y = fma(y * e, c, y);
    double s = a * y;                 // sqrt(a) ~ a * rsqrt(a), then one correction
    double r = fma(-s, s, a);
    return fma(0.5 * y, r, s);
}
// exact powers of two 2^-e and 2^e with e = exponent(w); w finite and > 0 (subnormal w: e = -1022)
GS_DEV void pow2_scales(double w, double& down, double& up) {
    int ex = (__double2hiint(w) >> 20) & 0x7ff;      // biased exponent
    ex = ex < 1 ? 1 : (ex > 2045 ? 2045 : ex);
    down = __hiloint2double((2046 - ex) << 20, 0);   // 2^(1023 - ex)
    up = __hiloint2double(ex << 20, 0);              // 2^(ex - 1023)
}

// ------------------------------------------------------------------------------------------------
// Guarded fast versions for the per-sweep scalar bookkeeping (deflation test, shift, start-row search).  Inside
// a safe exponent window they use the MUFU + Newton primitives above (results within ~1 ulp of the IEEE ones, no
// slow-path branches); outside they fall back to IEEE division / sqrt.  The double-double overloads are the
// exact operations.
// ------------------------------------------------------------------------------------------------
// Double-double reciprocal and reciprocal square root for the reflector on the serial chain of the double-double sweeps:
// a Float64 seed (MUFU + Newton, ~1 ulp) and ONE Newton step carried out in double-double.  The generic operators cost
// three IEEE divisions (a / b) resp. an IEEE sqrt and a reciprocal (Karp) in sequence; the six of them in the generic
// complex reflector were ~10 000 cycles of dependent latency per bulge step.  Operands: finite, normal, in a range whose
// squares do not leave the Float64 range (the callers guard).  Relative error ~2^-104 (the step's own truncation error is
// e^3 resp. (3/8) e^2 with e ~ 2^-52).
GS_DEV dd_t dd_rcp_fast(const dd_t& b) {
    const double x0 = fast_rcp(b.hi);
    const dd_t e = mk_dd(1.0) - dd_mul_d(b, x0);          // 1 - b x0 ~ 2^-52
    const double corr = x0 * fma(e.hi, e.hi, e.hi);        // x0 (e + e^2)
    double s, t;
    quick_two_sum(x0, corr, s, t);
    return mk_dd(s, t);
}
GS_DEV dd_t dd_rsqrt_fast(const dd_t& q) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(q.hi));
    {   // two Float64 Newton steps: 20 -> 40 -> 53 bits
        double e = fma(-q.hi * y0, y0, 1.0);
        y0 = fma(y0 * e, fma(e, 0.375, 0.5), y0);
        e = fma(-q.hi * y0, y0, 1.0);
        y0 = fma(y0 * e, 0.5, y0);
    }
    double p, pe;
    two_prod(y0, y0, p, pe);                               // y0^2 exactly
    const dd_t e = mk_dd(1.0) - q * mk_dd(p, pe);          // 1 - q y0^2 ~ 2^-52
    const double corr = 0.5 * y0 * e.hi;
    double s, t;
    quick_two_sum(y0, corr, s, t);
    return mk_dd(s, t);
}

GS_DEV bool q_exp_in(double a, unsigned lo, unsigned hi) {   // biased exponent of a in [lo, hi]
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    return (e - lo) <= (hi - lo);
}
GS_DEV double q_rcp(double a) { return q_exp_in(a, 40u, 2000u) ? fast_rcp(a) : 1.0 / a; }
GS_DEV dd_t q_rcp(const dd_t& a) { return q_exp_in(a.hi, 1023u - 300u, 1023u + 300u) ? dd_rcp_fast(a) : mk_dd(1.0) / a; }
GS_DEV double q_sqrt(double a) { return (a > 0.0 && q_exp_in(a, 40u, 2000u)) ? fast_sqrt(a) : __dsqrt_rn(a); }
GS_DEV dd_t q_sqrt(const dd_t& a) { return dd_sqrt(a); }

// ------------------------------------------------------------------------------------------------
// complex
// ------------------------------------------------------------------------------------------------
template <class R> struct __align__(16) cx {
    R re, im;
};
template <class R> GS_DEV cx<R> mk_cx(const R& re, const R& im) {
    cx<R> r;
    r.re = re;
    r.im = im;
    return r;
}
template <class R> GS_DEV cx<R> operator+(const cx<R>& a, const cx<R>& b) { return mk_cx<R>(a.re + b.re, a.im + b.im); }
template <class R> GS_DEV cx<R> operator-(const cx<R>& a, const cx<R>& b) { return mk_cx<R>(a.re - b.re, a.im - b.im); }
template <class R> GS_DEV cx<R> operator-(const cx<R>& a) { return mk_cx<R>(-a.re, -a.im); }
template <class R> GS_DEV cx<R> operator*(const cx<R>& a, const cx<R>& b) {
    return mk_cx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> GS_DEV cx<R> operator*(const cx<R>& a, const R& b) { return mk_cx<R>(a.re * b, a.im * b); }
// Complex double-double composites as out-of-line functions: inlined at every use (~90 FP64 instructions per product,
// ~350 per reflector item) they made the complex double-double QR kernels 300+ KB of straight-line code whose hot loop
// does not fit the instruction caches (ncu: 2.5 "no instruction" stall cycles per issue).
//   e_axty(a, x, t, y) = a x + t y (t real): the reflector's inner product;  e_bsv(b, s, v) = b - s v: its update.
// The generic versions are the plain expressions (ComplexF64 code is unchanged by them).
// OUT = false keeps the plain expressions (whose complex product is still out of line for double-double): measured
// better for n <= 32 (18.5 k against 16.4 k matrices/s at 32x32), the composites for 96x96 (935 against 891; 709 inlined).
template <bool OUT, class R> GS_DEV cx<R> e_axty(const cx<R>& a, const cx<R>& x, const R& t, const cx<R>& y) { return a * x + t * y; }
template <bool OUT, class R> GS_DEV cx<R> e_bsv(const cx<R>& b, const cx<R>& s, const cx<R>& v) { return b - s * v; }
#ifndef GS_CDD_INLINE_MUL
GS_DEV cx<dd_t> cdd_mul_inl(const cx<dd_t>& a, const cx<dd_t>& b) {
    return mk_cx<dd_t>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static __device__ __noinline__ cx<dd_t> cdd_mul(cx<dd_t> a, cx<dd_t> b) { return cdd_mul_inl(a, b); }
static __device__ __noinline__ cx<dd_t> cdd_axty(cx<dd_t> a, cx<dd_t> x, dd_t t, cx<dd_t> y) {
    const cx<dd_t> p = cdd_mul_inl(a, x);
    return mk_cx<dd_t>(p.re + t * y.re, p.im + t * y.im);
}
static __device__ __noinline__ cx<dd_t> cdd_bsv(cx<dd_t> b, cx<dd_t> s, cx<dd_t> v) {
    const cx<dd_t> p = cdd_mul_inl(s, v);
    return mk_cx<dd_t>(b.re - p.re, b.im - p.im);
}
GS_DEV cx<dd_t> operator*(const cx<dd_t>& a, const cx<dd_t>& b) { return cdd_mul(a, b); }
template <> GS_DEV cx<dd_t> e_axty<true, dd_t>(const cx<dd_t>& a, const cx<dd_t>& x, const dd_t& t, const cx<dd_t>& y) { return cdd_axty(a, x, t, y); }
template <> GS_DEV cx<dd_t> e_bsv<true, dd_t>(const cx<dd_t>& b, const cx<dd_t>& s, const cx<dd_t>& v) { return cdd_bsv(b, s, v); }
#endif

template <class R> GS_DEV cx<R> operator*(const R& a, const cx<R>& b) { return mk_cx<R>(a * b.re, a * b.im); }
template <class R> GS_DEV cx<R> operator/(const cx<R>& a, const R& b) { return mk_cx<R>(a.re / b, a.im / b); }
template <class R> GS_DEV cx<R> operator/(const cx<R>& a, const cx<R>& b) {   // Smith
    if (r_abs(b.re) >= r_abs(b.im)) {
        R r = b.im / b.re;
        R den = b.re + r * b.im;
        return mk_cx<R>((a.re + a.im * r) / den, (a.im - a.re * r) / den);
    } else {
        R r = b.re / b.im;
        R den = b.im + r * b.re;
        return mk_cx<R>((a.re * r + a.im) / den, (a.im * r - a.re) / den);
    }
}
template <class R> GS_DEV cx<R>& operator+=(cx<R>& a, const cx<R>& b) { a = a + b; return a; }
template <class R> GS_DEV cx<R>& operator-=(cx<R>& a, const cx<R>& b) { a = a - b; return a; }
template <class R> GS_DEV cx<R>& operator*=(cx<R>& a, const cx<R>& b) { a = a * b; return a; }
template <class R> GS_DEV cx<R>& operator*=(cx<R>& a, const R& b) { a = a * b; return a; }
template <class R> GS_DEV cx<R> cconj(const cx<R>& a) { return mk_cx<R>(a.re, -a.im); }
GS_DEV double cconj(double a) { return a; }
GS_DEV dd_t cconj(const dd_t& a) { return a; }
template <class R> GS_DEV R c_abs(const cx<R>& a) { return r_hypot(a.re, a.im); }
template <class R> GS_DEV R abs1(const cx<R>& a) { return r_abs(a.re) + r_abs(a.im); }   // src/util.jl:31-32
GS_DEV double abs1(double a) { return fabs(a); }
GS_DEV dd_t abs1(const dd_t& a) { return r_abs(a); }
// principal square root: rho = sqrt((|z| + |x|)/2), eta = y / (2 rho), on a copy scaled by max(|x|,|y|)
template <class R> GS_DEV cx<R> c_sqrt(const cx<R>& z) {
    R zero = r_const<R>(0.0), half = r_const<R>(0.5);
    if (z.re == zero && z.im == zero) return mk_cx<R>(zero, z.im);
    R m = r_max(r_abs(z.re), r_abs(z.im));
    R x = z.re / m, y = z.im / m;
    R rho = r_sqrt((r_hypot(x, y) + r_abs(x)) * half);
    R sm = r_sqrt(m);
    R xi = rho, eta = (y / rho) * half;
    if (x < zero) {
        xi = r_abs(eta);
        eta = r_copysign(rho, y);
    }
    return mk_cx<R>(xi * sm, eta * sm);
}

// guarded fast complex square root / division / modulus (Float64); generic fallbacks otherwise
GS_DEV cx<double> c_sqrt_q(const cx<double>& z) {
    const double m = fmax(fabs(z.re), fabs(z.im));
    if (!q_exp_in(m, 1023u - 400u, 1023u + 400u)) return c_sqrt(z);   // also zero / inf / nan
    const double h = fast_sqrt(fma(z.re, z.re, z.im * z.im));
    const double rho = fast_sqrt((h + fabs(z.re)) * 0.5);
    double xi = rho, eta = (z.im * fast_rcp(rho)) * 0.5;
    if (z.re < 0.0) {
        xi = fabs(eta);
        eta = copysign(rho, z.im);
    }
    return mk_cx<double>(xi, eta);
}
// double-double: |z| and rho from reciprocal square roots (no division: eta = y / (2 rho) = y rsqrt(rho^2) / 2)
GS_DEV cx<dd_t> c_sqrt_q(const cx<dd_t>& z) {
    const double m = fmax(fabs(z.re.hi), fabs(z.im.hi));
    if (!q_exp_in(m, 1023u - 300u, 1023u + 300u)) return c_sqrt(z);   // also zero / inf / nan
    const dd_t q = z.re * z.re + z.im * z.im;
    const dd_t h = q * dd_rsqrt_fast(q);
    const dd_t q2 = dd_mul_d(h + r_abs(z.re), 0.5);
    const dd_t rs2 = dd_rsqrt_fast(q2);
    const dd_t rho = q2 * rs2;
    dd_t xi = rho, eta = dd_mul_d(z.im * rs2, 0.5);
    if (z.re.hi < 0.0) {
        xi = r_abs(eta);
        eta = r_copysign(rho, z.im);
    }
    return mk_cx<dd_t>(xi, eta);
}
GS_DEV cx<double> c_div_q(const cx<double>& a, const cx<double>& b) {
    const double m = fmax(fabs(b.re), fabs(b.im));
    if (!q_exp_in(m, 1023u - 400u, 1023u + 400u)) return a / b;
    const double rd = fast_rcp(fma(b.re, b.re, b.im * b.im));
    return mk_cx<double>(fma(a.re, b.re, a.im * b.im) * rd, fma(a.im, b.re, -a.re * b.im) * rd);
}
GS_DEV cx<dd_t> c_div_q(const cx<dd_t>& a, const cx<dd_t>& b) {
    const double m = fmax(fabs(b.re.hi), fabs(b.im.hi));
    if (!q_exp_in(m, 1023u - 300u, 1023u + 300u)) return a / b;
    const dd_t rd = dd_rcp_fast(b.re * b.re + b.im * b.im);
    return mk_cx<dd_t>((a.re * b.re + a.im * b.im) * rd, (a.im * b.re - a.re * b.im) * rd);
}
GS_DEV double c_abs_q(const cx<double>& a) {
    const double m = fmax(fabs(a.re), fabs(a.im));
    if (!q_exp_in(m, 1023u - 400u, 1023u + 400u)) return c_abs(a);
    return fast_sqrt(fma(a.re, a.re, a.im * a.im));
}
GS_DEV dd_t c_abs_q(const cx<dd_t>& a) {
    const double m = fmax(fabs(a.re.hi), fabs(a.im.hi));
    if (!q_exp_in(m, 1023u - 300u, 1023u + 300u)) return c_abs(a);
    const dd_t q = a.re * a.re + a.im * a.im;
    return q * dd_rsqrt_fast(q);
}

// element-type traits
template <class T> struct etraits;
template <> struct etraits<double> {
    typedef double real;
    static constexpr bool is_complex = false;
};
template <> struct etraits<dd_t> {
    typedef dd_t real;
    static constexpr bool is_complex = false;
};
template <class R> struct etraits<cx<R>> {
    typedef R real;
    static constexpr bool is_complex = true;
};
template <class T> GS_DEV T e_zero();
template <> GS_DEV double e_zero<double>() { return 0.0; }
template <> GS_DEV dd_t e_zero<dd_t>() { return mk_dd(0.0); }
template <> GS_DEV cx<double> e_zero<cx<double>>() { return mk_cx<double>(0.0, 0.0); }
template <> GS_DEV cx<dd_t> e_zero<cx<dd_t>>() { return mk_cx<dd_t>(mk_dd(0.0), mk_dd(0.0)); }
template <class T> GS_DEV T e_one();
template <> GS_DEV double e_one<double>() { return 1.0; }
template <> GS_DEV dd_t e_one<dd_t>() { return mk_dd(1.0); }
template <> GS_DEV cx<double> e_one<cx<double>>() { return mk_cx<double>(1.0, 0.0); }
template <> GS_DEV cx<dd_t> e_one<cx<dd_t>>() { return mk_cx<dd_t>(mk_dd(1.0), mk_dd(0.0)); }

// modulus of an element (norm(A, Inf) of a Matrix is max |a_ij|, src/util.jl:17)
GS_DEV double e_abs(double a) { return fabs(a); }
GS_DEV dd_t e_abs(const dd_t& a) { return r_abs(a); }
template <class R> GS_DEV R e_abs(const cx<R>& a) { return c_abs(a); }
// max over |re|, |im| parts (used by the scaled 2-norm)
GS_DEV double e_maxpart(double a) { return fabs(a); }
GS_DEV dd_t e_maxpart(const dd_t& a) { return r_abs(a); }
template <class R> GS_DEV R e_maxpart(const cx<R>& a) { return r_max(r_abs(a.re), r_abs(a.im)); }
// |a*s|^2 summed over parts
GS_DEV double e_sq_scaled(double a, double s) { double t = a * s; return t * t; }
GS_DEV dd_t e_sq_scaled(const dd_t& a, const dd_t& s) { dd_t t = a * s; return t * t; }
template <class R> GS_DEV R e_sq_scaled(const cx<R>& a, const R& s) {
    R x = a.re * s, y = a.im * s;
    return x * x + y * y;
}
// element * real
GS_DEV double e_scale(double a, double s) { return a * s; }
GS_DEV dd_t e_scale(const dd_t& a, const dd_t& s) { return a * s; }
template <class R> GS_DEV cx<R> e_scale(const cx<R>& a, const R& s) { return mk_cx<R>(a.re * s, a.im * s); }

// fused element multiply-adds: a b + c, c - a b, conj(a) b + c, c - a conj(b).  Float64 / ComplexF64: every product
// term is contracted into an FMA (2 / 4 DFMA instead of DMUL + DFMA + DADD per component, and one rounding fewer);
// the double-double kinds use their ordinary operators.
template <class T> GS_DEV T e_fma(const T& a, const T& b, const T& c) { return a * b + c; }
template <class T> GS_DEV T e_fnma(const T& a, const T& b, const T& c) { return c - a * b; }
template <class T> GS_DEV T e_fma_cja(const T& a, const T& b, const T& c) { return cconj(a) * b + c; }
template <class T> GS_DEV T e_fnma_cjb(const T& a, const T& b, const T& c) { return c - a * cconj(b); }
GS_DEV double e_fma(double a, double b, double c) { return fma(a, b, c); }                 // a b + c
GS_DEV double e_fnma(double a, double b, double c) { return fma(-a, b, c); }               // c - a b
GS_DEV double e_fma_cja(double a, double b, double c) { return fma(a, b, c); }             // conj(a) b + c
GS_DEV double e_fnma_cjb(double a, double b, double c) { return fma(-a, b, c); }           // c - a conj(b)
GS_DEV cx<double> e_fma(const cx<double>& a, const cx<double>& b, const cx<double>& c) {
    return mk_cx<double>(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im)));
}
GS_DEV cx<double> e_fnma(const cx<double>& a, const cx<double>& b, const cx<double>& c) {
    return mk_cx<double>(fma(-a.re, b.re, fma(a.im, b.im, c.re)), fma(-a.re, b.im, fma(-a.im, b.re, c.im)));
}
GS_DEV cx<double> e_fma_cja(const cx<double>& a, const cx<double>& b, const cx<double>& c) {
    return mk_cx<double>(fma(a.re, b.re, fma(a.im, b.im, c.re)), fma(a.re, b.im, fma(-a.im, b.re, c.im)));
}
GS_DEV cx<double> e_fnma_cjb(const cx<double>& a, const cx<double>& b, const cx<double>& c) {
    return mk_cx<double>(fma(-a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(-a.im, b.re, c.im)));
}

// warp shuffles for every scalar type
GS_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
GS_DEV dd_t shfl_xor(const dd_t& v, int m) {
    return mk_dd(__shfl_xor_sync(0xffffffffu, v.hi, m), __shfl_xor_sync(0xffffffffu, v.lo, m));
}

}  // namespace gs
