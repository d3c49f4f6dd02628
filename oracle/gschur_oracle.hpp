// TEST INFRASTRUCTURE ONLY — CPU oracle for the Schur hot path of RalphAS/GenericSchur.jl.
//
// This header is a plain C++ restatement of the reference's algorithm (pure Julia, cannot be run in
// this environment — no Julia toolchain).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it; the product (genericschur.jl_b200/) never does.
//
// PINNING STATUS: the reference holds no golden bit vectors for this path (its tests are property
// checks).  The oracle is pinned against (i) every known-answer fixture those tests
// hold — Godunov 7x7 eigenvalues (test/testfuncs.jl:127-142), the nearly-degenerate 2x2 of
// test/real.jl:311-317, Jordan blocks, the structural `==` post-conditions — and (ii) the LAPACK
// Fortran originals the reference says it was translated from (dlahqr/zlahqr/dlanv2/xLARFG/xGEHD2,
// reachable through scipy's OpenBLAS).  It could NOT be pinned against outputs of the Julia code
// itself: "parity unpinned against reference-executed outputs".
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Indices are 1-based through the accessor, to read like the Julia.
#pragma once
#include <vector>
#include "scalars.hpp"

namespace gso {

template <class T> struct Mat {
    T* p;
    long ld;
    int m, n;
    Mat(T* p_, int m_, int n_, long ld_) : p(p_), ld(ld_), m(m_), n(n_) {}
    T& operator()(int i, int j) const { return p[(i - 1) + (long)(j - 1) * ld]; }
};

struct Unconverged {   // UnconvergedException, src/GenericSchur.jl:39-45
    int maxiter;
};

struct Stats {
    long sweeps = 0;        // QR sweeps run
    long applications = 0;  // reflector applications (one per bulge step)
    long rowpairs = 0;      // updated row/column "units" (for the flop model)
    long exceptional = 0;
};

// src/util.jl:5-12
template <class R> R safemin() {
    R sfmin = RT<R>::floatmin();
    R small = R(1.0) / RT<R>::floatmax();
    if (small >= sfmin) sfmin = small * (R(1.0) + RT<R>::eps());
    return sfmin;
}

// src/util.jl:31-32
template <class R> R abs1(const R& z) { return RT<R>::abs(z); }
template <class R> R abs1(const Cx<R>& z) { return RT<R>::abs(z.re) + RT<R>::abs(z.im); }

template <class R> R real_(const R& z) { return z; }
template <class R> R real_(const Cx<R>& z) { return z.re; }
template <class R> R absval(const R& z) { return RT<R>::abs(z); }
template <class R> R absval(const Cx<R>& z) { return cabs(z); }

// src/util.jl:41-80  (xLASCL): multiply by cto/cfrom without over/underflow
template <class T, class R> void safescale(T* a, long count, R cfrom, R cto) {
    R smlnum = safemin<R>();
    R bignum = R(1.0) / smlnum;
    R cfromc = cfrom, ctoc = cto, mul(0.0);
    bool done = false;
    while (!done) {
        R cfrom1 = cfromc * smlnum;
        if (cfrom1 == cfromc) {
            mul = ctoc / cfromc;
            done = true;
        } else {
            R cto1 = ctoc / bignum;
            if (cto1 == ctoc) {
                mul = cto;
                done = true;
                cfromc = R(1.0);
            } else if (RT<R>::abs(cfrom1) > RT<R>::abs(ctoc) && ctoc != R(0.0)) {
                mul = smlnum;
                done = false;
                cfromc = R(1.0);      // as written at src/util.jl:66 (xLASCL has cfromc = cfrom1)
            } else if (RT<R>::abs(cto1) > RT<R>::abs(cfromc)) {
                mul = bignum;
                done = false;
                cfromc = cfrom1;      // as written at src/util.jl:70 (xLASCL has ctoc = cto1)
            } else {
                mul = ctoc / cfromc;
                done = true;
            }
        }
        for (long i = 0; i < count; ++i) a[i] = a[i] * mul;
    }
}
// NOTE: the two `done = false` branches differ from LAPACK xLASCL in the reference text; they are
// restated as written.  They are only reachable when cto/cfrom is not representable in one step,
// which cannot happen on the hot path: _scale! always calls with one argument equal to smlnum or
// bignum (sqrt(safemin)/eps or its inverse), so a single pass `mul = cto/cfrom` is taken.

// src/util.jl:14-29.  A is the full n x n column-major array with leading dimension ld.
template <class T> struct ScaleInfo {
    bool scaled;
    typename ElemTraits<T>::Real cscale, anrm;
};
template <class T> ScaleInfo<T> scale_matrix(Mat<T> A) {
    typedef typename ElemTraits<T>::Real R;
    R smlnum = RT<R>::sqrt(safemin<R>()) / RT<R>::eps();
    R bignum = R(1.0) / smlnum;
    R anrm(0.0);   // norm(A, Inf) of a Matrix = max |a_ij|  (modulus for complex)
    for (int j = 1; j <= A.n; ++j)
        for (int i = 1; i <= A.m; ++i) {
            R a = absval(A(i, j));
            if (RT<R>::isnan(a)) anrm = a;
            else if (a > anrm) anrm = a;
        }
    ScaleInfo<T> s;
    s.scaled = false;
    s.cscale = R(1.0);
    s.anrm = anrm;
    if (anrm > R(0.0) && anrm < smlnum) {
        s.scaled = true;
        s.cscale = smlnum;
    } else if (anrm > bignum) {
        s.scaled = true;
        s.cscale = bignum;
    }
    if (s.scaled)
        for (int j = 1; j <= A.n; ++j) safescale(&A(1, j), (long)A.m, anrm, s.cscale);
    return s;
}

// src/util.jl:506-527 (real) — scaled sum of squares over x[0..n-1] with stride 1
template <class R> R norm2(const R* x, int n) {
    if (n < 1) return R(0.0);
    if (n == 1) return RT<R>::abs(x[0]);
    R scale(0.0), ssq(0.0);
    for (int i = 0; i < n; ++i) {
        if (x[i] != R(0.0)) {
            R a = RT<R>::abs(x[i]);
            if (scale < a) {
                R q = scale / a;
                ssq = R(1.0) + ssq * (q * q);
                scale = a;
            } else {
                R q = a / scale;
                ssq += q * q;
            }
        }
    }
    return scale * RT<R>::sqrt(ssq);
}
// src/util.jl:529-557 (complex)
template <class R> R norm2(const Cx<R>* x, int n) {
    if (n < 1) return R(0.0);
    if (n == 1) return cabs(x[0]);
    R scale(0.0), ssq(0.0);
    for (int i = 0; i < n; ++i) {
        for (int part = 0; part < 2; ++part) {
            const R& xv = part == 0 ? x[i].re : x[i].im;
            if (xv != R(0.0)) {
                R a = RT<R>::abs(xv);
                if (scale < a) {
                    R q = scale / a;
                    ssq = R(1.0) + ssq * (q * q);
                    scale = a;
                } else {
                    R q = a / scale;
                    ssq += q * q;
                }
            }
        }
    }
    return scale * RT<R>::sqrt(ssq);
}

// src/util.jl:562-570 (dlapy3)
template <class R> R hypot3(const R& x, const R& y, const R& z) {
    R xa = RT<R>::abs(x), ya = RT<R>::abs(y), za = RT<R>::abs(z);
    R w = rmax(rmax(xa, ya), za);
    R rw = R(1.0) / w;
    R a = rw * xa, b = rw * ya, c = rw * za;
    return w * RT<R>::sqrt(a * a + b * b + c * c);
}

// src/householder.jl:12-54 (real xLARFG).  x has length n, stride 1; returns tau.
template <class R> R reflector(R* x, int n) {
    if (n <= 1) return R(0.0);
    R sfmin = R(2.0) * RT<R>::floatmin() / RT<R>::eps();
    R alpha = x[0];
    R xnorm = norm2(x + 1, n - 1);
    if (xnorm == R(0.0)) return R(0.0);
    R beta = -RT<R>::copysign(RT<R>::hypot(alpha, xnorm), alpha);
    int kount = 0;
    bool smallb = RT<R>::abs(beta) < sfmin;
    if (smallb) {
        R rsfmin = R(1.0) / sfmin;
        while (smallb) {
            kount += 1;
            for (int j = 1; j < n; ++j) x[j] *= rsfmin;
            beta *= rsfmin;
            alpha *= rsfmin;
            smallb = (RT<R>::abs(beta) < sfmin) && (kount < 20);
        }
        xnorm = norm2(x + 1, n - 1);
        beta = -RT<R>::copysign(RT<R>::hypot(alpha, xnorm), alpha);
    }
    R tau = (beta - alpha) / beta;
    R t = R(1.0) / (alpha - beta);
    for (int j = 1; j < n; ++j) x[j] *= t;
    for (int j = 1; j <= kount; ++j) beta *= sfmin;
    x[0] = beta;
    return tau;
}

// src/householder.jl:56-102 (complex xLARFG; beta real; n == 1 is a pure phase)
template <class R> Cx<R> reflector(Cx<R>* x, int n) {
    typedef Cx<R> C;
    if (n < 1) return C(R(0.0));
    R sfmin = RT<R>::floatmin() / RT<R>::eps();
    C alpha = x[0];
    R ar = alpha.re, ai = alpha.im;
    R xnorm = norm2(x + 1, n - 1);
    if (xnorm == R(0.0) && ai == R(0.0)) return C(R(0.0));
    R beta = -RT<R>::copysign(hypot3(ar, ai, xnorm), ar);
    int kount = 0;
    bool smallb = RT<R>::abs(beta) < sfmin;
    if (smallb) {
        R rsfmin = R(1.0) / sfmin;
        while (smallb) {
            kount += 1;
            for (int j = 1; j < n; ++j) x[j] *= rsfmin;
            beta *= rsfmin;
            ar *= rsfmin;
            ai *= rsfmin;
            smallb = (RT<R>::abs(beta) < sfmin) && (kount < 20);
        }
        xnorm = norm2(x + 1, n - 1);
        alpha = C(ar, ai);
        beta = -RT<R>::copysign(hypot3(ar, ai, xnorm), ar);
    }
    C tau((beta - ar) / beta, -ai / beta);
    C t = C(R(1.0)) / (alpha - C(beta));
    for (int j = 1; j < n; ++j) x[j] *= t;
    for (int j = 1; j <= kount; ++j) beta *= sfmin;
    x[0] = C(beta);
    return tau;
}

// src/hessenberg.jl:3-17 with the two Householder applications of src/householder.jl:157-172
// (lmul!(H', A): per column dot + axpy) and :140-155 (rmul!(A, H, x): gemv, 2 axpy, rank-1
// src/util.jl:103-114).  A n x n in place; tau[0..n-2].
template <class T> void hessenberg(Mat<T> A, T* tau) {
    int n = A.n;
    std::vector<T> xw(n);
    for (int i = 1; i <= n - 1; ++i) {
        T t = reflector(&A(i + 1, i), n - i);
        tau[i - 1] = t;
        int nv = n - i - 1;            // length of the stored tail v = A[i+2:n, i]
        T* v = nv > 0 ? &A(i + 2, i) : (T*)0;
        T tc = conj_(t);
        // lmul!(H', view(A, i+1:n, i+1:n))
        for (int j = i + 1; j <= n; ++j) {
            T va = A(i + 1, j);
            for (int r = 0; r < nv; ++r) va += conj_(v[r]) * A(i + 2 + r, j);   // dot(v, Aj) conjugates v
            va = tc * va;
            A(i + 1, j) -= va;
            for (int r = 0; r < nv; ++r) A(i + 2 + r, j) -= va * v[r];
        }
        // rmul!(view(A, :, i+1:n), H, xwrk)
        for (int r = 1; r <= n; ++r) {
            T x = A(r, i + 1);                                   // x = A1*v + a1
            for (int c = 0; c < nv; ++c) x += A(r, i + 2 + c) * v[c];
            xw[r - 1] = x;
        }
        for (int r = 1; r <= n; ++r) A(r, i + 1) -= t * xw[r - 1];   // a1 -= tau x
        for (int c = 0; c < nv; ++c) {                               // A1 += x * (-tau) * v'
            T yjc = conj_(v[c]);
            for (int r = 1; r <= n; ++r) A(r, i + 2 + c) += xw[r - 1] * (-t) * yjc;
        }
    }
}

// src/hessenberg.jl:150-166: explicit Q from the packed reflectors.  `Matrix(QRPackedQ(view(A,2:n,2:n), tau))`
// is the stdlib's backward accumulation Q1 = H1 H2 ... H_{n-1} applied to the identity
// (LinearAlgebra lmul!(::QRPackedQ, B): for k = last:-1:1, per column: vBj = B[k,j] + v_k' B[k+1:,j];
// vBj *= tau_k; B[k,j] -= vBj; B[k+1:,j] -= v_k vBj).  F holds the factors as left by hessenberg().
template <class T> void materializeQ(Mat<T> F, const T* tau, Mat<T> Q) {
    int n = F.n;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) Q(i, j) = (i == j) ? T(1.0) : T(0.0);
    // reflector k (1-based, k = 1..n-1) lives in F[k+2:n, k] and acts on rows k+1..n
    for (int k = n - 1; k >= 1; --k) {
        int nv = n - k - 1;
        const T* v = nv > 0 ? &F(k + 2, k) : (const T*)0;
        for (int j = 2; j <= n; ++j) {
            T vb = Q(k + 1, j);
            for (int r = 0; r < nv; ++r) vb += conj_(v[r]) * Q(k + 2 + r, j);
            vb = tau[k - 1] * vb;
            Q(k + 1, j) -= vb;
            for (int r = 0; r < nv; ++r) Q(k + 2 + r, j) -= v[r] * vb;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Complex path
// ---------------------------------------------------------------------------------------------

// src/GenericSchur.jl:374-504
template <class R>
void singleShiftQR(Mat<Cx<R>> HH, Mat<Cx<R>>* Z, const Cx<R>& shift, int istart, int iend, Stats* st) {
    typedef Cx<R> C;
    int n = HH.n;
    R ulp = RT<R>::eps();
    C v[2];
    int istart1 = -1;
    bool flag = false;
    for (int mm = iend - 1; mm >= istart + 1; --mm) {
        C h11 = HH(mm, mm), h22 = HH(mm + 1, mm + 1);
        C h11s = h11 - shift;
        R h21 = HH(mm + 1, mm).re;
        R s = abs1(h11s) + RT<R>::abs(h21);
        h11s /= s;
        h21 /= s;
        v[0] = h11s;
        v[1] = C(h21);
        R h10 = HH(mm, mm - 1).re;
        if (RT<R>::abs(h10) * RT<R>::abs(h21) <= ulp * (abs1(h11s) * (abs1(h11) + abs1(h22)))) {
            istart1 = mm;
            flag = true;
            break;
        }
    }
    if (!flag) {
        istart1 = istart;
        C h11 = HH(istart, istart);
        C h11s = h11 - shift;
        R h21 = HH(istart + 1, istart).re;
        R s = abs1(h11s) + RT<R>::abs(h21);
        h11s /= s;
        h21 /= s;
        v[0] = h11s;
        v[1] = C(h21);
    }
    if (st) st->sweeps += 1;
    for (int k = istart1; k <= iend - 1; ++k) {
        if (k > istart1) {
            v[0] = HH(k, k - 1);
            v[1] = HH(k + 1, k - 1);
        }
        C tau1 = reflector(v, 2);
        if (k > istart1) {
            HH(k, k - 1) = v[0];
            HH(k + 1, k - 1) = C(R(0.0));
        }
        C v2 = v[1];
        R tau2 = (tau1 * v2).re;
        C tau1c = conj_(tau1), v2c = conj_(v2);
        for (int j = k; j <= n; ++j) {
            C ss = tau1c * HH(k, j) + tau2 * HH(k + 1, j);
            HH(k, j) -= ss;
            HH(k + 1, j) -= ss * v2;
        }
        int jmax = (k + 2 < iend) ? k + 2 : iend;
        for (int j = 1; j <= jmax; ++j) {
            C ss = tau1 * HH(j, k) + tau2 * HH(j, k + 1);
            HH(j, k) -= ss;
            HH(j, k + 1) -= ss * v2c;
        }
        if (Z) {
            for (int j = 1; j <= Z->m; ++j) {
                C ss = tau1 * (*Z)(j, k) + tau2 * (*Z)(j, k + 1);
                (*Z)(j, k) -= ss;
                (*Z)(j, k + 1) -= ss * v2c;
            }
        }
        if (st) {
            st->applications += 1;
            st->rowpairs += (n - k + 1) + jmax + (Z ? Z->m : 0);
        }
        if (k == istart1 && istart1 > istart) {
            // src/GenericSchur.jl:461-482: keep HH[istart1, istart1-1] real after a late start
            C t = C(R(1.0)) - tau1;
            t /= cabs(t);
            C tc = conj_(t);
            HH(istart1 + 1, istart1) *= tc;
            if (istart1 + 2 <= iend) HH(istart1 + 2, istart1 + 1) *= t;
            for (int j = istart1; j <= iend; ++j) {
                if (j != istart1 + 1) {
                    for (int c = j + 1; c <= n; ++c) HH(j, c) *= t;
                    for (int r = 1; r <= j - 1; ++r) HH(r, j) *= tc;
                    if (Z)
                        for (int r = 1; r <= Z->m; ++r) (*Z)(r, j) *= tc;
                }
            }
        }
    }
    // src/GenericSchur.jl:486-500: make the tail sub-diagonal real
    C t = HH(iend, iend - 1);
    if (t.im != R(0.0)) {
        R rt = cabs(t);
        HH(iend, iend - 1) = C(rt);
        t /= rt;
        C tc = conj_(t);
        for (int c = iend + 1; c <= n; ++c) HH(iend, c) *= tc;
        for (int r = 1; r <= iend - 1; ++r) HH(r, iend) *= t;
        if (Z)
            for (int r = 1; r <= Z->m; ++r) (*Z)(r, iend) *= t;
    }
}

// src/GenericSchur.jl:194-335.  HH in place (ends upper triangular + whatever is below the
// sub-diagonal zeroed), Z accumulated if non-null, w = diag(HH).  Throws Unconverged.
// `checksd` failure is reported by returning -1 (ArgumentError in the reference), else 0.
template <class R>
int gschur_hess_complex(Mat<Cx<R>> HH, Mat<Cx<R>>* Z, Cx<R>* w, int maxiter, int maxinner, bool checksd,
                        Stats* st) {
    typedef Cx<R> C;
    int n = HH.n;
    if (maxiter <= 0) maxiter = 100 * n;
    if (maxinner <= 0) maxinner = 30 * n;
    int istart = 1, iend = n;
    if (checksd)
        for (int j = 1; j <= n - 1; ++j)
            if (HH(j + 1, j).im != R(0.0)) return -1;
    R ulp = RT<R>::eps();
    R smallnum = safemin<R>() * (R((double)n) / ulp);
    R rzero(0.0), half(0.5), threeq(0.75);
    for (int j = 1; j <= n - 1; ++j)
        for (int i = j + 2; i <= n; ++i) HH(i, j) = C(rzero);
    int it = 0;
    while (iend >= 1) {
        istart = 1;
        for (int its = 0; its <= maxinner; ++its) {
            it += 1;
            if (it > maxiter) throw Unconverged{maxiter};
            for (int is = iend - 1; is >= istart; --is) {
                if (abs1(HH(is + 1, is)) <= smallnum) {
                    istart = is + 1;
                    break;
                }
                R tst = abs1(HH(is, is)) + abs1(HH(is + 1, is + 1));
                if (tst == rzero) {
                    if (is - 1 >= 1) tst += RT<R>::abs(HH(is, is - 1).re);
                    if (is + 2 <= n) tst += RT<R>::abs(HH(is + 2, is + 1).re);
                }
                if (RT<R>::abs(HH(is + 1, is).re) <= ulp * tst) {
                    R a1 = abs1(HH(is + 1, is)), a2 = abs1(HH(is, is + 1));
                    R ab = rmax(a1, a2), ba = rmin(a1, a2);
                    R d1 = abs1(HH(is + 1, is + 1)), d2 = abs1(HH(is, is) - HH(is + 1, is + 1));
                    R aa = rmax(d1, d2), bb = rmin(d1, d2);
                    R s = aa + ab;
                    if (ba * (ab / s) <= rmax(smallnum, ulp * (bb * (aa / s)))) {
                        istart = is + 1;
                        break;
                    }
                }
            }
            if (istart > 1) HH(istart, istart - 1) = C(rzero);
            if (istart >= iend) {
                iend -= 1;
                break;
            }
            C t;
            if (its % 30 == 10) {
                R s = threeq * RT<R>::abs(HH(istart + 1, istart).re);
                t = C(s) + HH(istart, istart);
                if (st) st->exceptional += 1;
            } else if (its % 30 == 20) {
                R s = threeq * RT<R>::abs(HH(iend, iend - 1).re);
                t = C(s) + HH(iend, iend);
                if (st) st->exceptional += 1;
            } else {
                t = HH(iend, iend);
                C u = csqrt(HH(iend - 1, iend)) * csqrt(HH(iend, iend - 1));
                R s = abs1(u);
                if (s != rzero) {
                    C x = half * (HH(iend - 1, iend - 1) - t);
                    R sx = abs1(x);
                    s = rmax(s, abs1(x));
                    C xs = x / s, us = u / s;
                    C y = s * csqrt(xs * xs + us * us);
                    if (sx > rzero) {
                        C xx = x / sx;
                        if (xx.re * y.re + xx.im * y.im < rzero) y = -y;
                    }
                    t -= u * (u / (x + y));
                }
            }
            singleShiftQR(HH, Z, t, istart, iend, st);
        }
    }
    for (int j = 1; j <= n; ++j) w[j - 1] = HH(j, j);
    // Schur(triu(HH), Z, w): zero the strictly lower triangle
    for (int j = 1; j <= n; ++j)
        for (int i = j + 1; i <= n; ++i) HH(i, j) = C(rzero);
    return 0;
}

// src/GenericSchur.jl:350-372.  A -> T in place (the reference returns a fresh triu copy; the
// boundary writes T over A), Z n x n (nullable: wantZ=false), w[n].
template <class R>
int gschur_complex(Mat<Cx<R>> A, Mat<Cx<R>>* Z, Cx<R>* w, bool scale, int maxiter, Stats* st) {
    typedef Cx<R> C;
    int n = A.n;
    if (n == 0) return 0;
    ScaleInfo<C> si;
    si.scaled = false;
    if (scale) si = scale_matrix(A);
    std::vector<C> tau(n > 1 ? n - 1 : 1);
    hessenberg(A, tau.data());
    if (Z) materializeQ(A, tau.data(), *Z);
    int rc = gschur_hess_complex(A, Z, w, maxiter, 0, false, st);
    if (si.scaled) {
        for (int j = 1; j <= n; ++j) safescale(&A(1, j), (long)j, si.cscale, si.anrm);
        for (int j = 1; j <= n; ++j) w[j - 1] = A(j, j);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Real path
// ---------------------------------------------------------------------------------------------

// src/GenericSchur.jl:716-803 (dlanv2).  abcd in place; returns cs, sn and (w1, w2).
template <class R>
void gs2x2(R& a, R& b, R& c, R& d, R& cs, R& sn, Cx<R>& w1, Cx<R>& w2) {
    R zero(0.0), one(1.0), half(0.5);
    auto sgn = [&](const R& x) { return (x < zero) ? R(-1.0) : R(1.0); };
    R small = R(4.0) * RT<R>::eps();
    if (c == zero) {
        cs = one;
        sn = zero;
    } else if (b == zero) {
        cs = zero;
        sn = one;
        R a0 = a, c0 = c, d0 = d;
        a = d0;
        b = -c0;
        c = zero;
        d = a0;
    } else if ((a - d) == zero && (b * c < zero)) {
        cs = one;
        sn = zero;
    } else {
        R asubd = a - d;
        R p = half * asubd;
        R bcmax = rmax(RT<R>::abs(b), RT<R>::abs(c));
        R bcmis = rmin(RT<R>::abs(b), RT<R>::abs(c)) * sgn(b) * sgn(c);
        R scale = rmax(RT<R>::abs(p), bcmax);
        R z = (p / scale) * p + (bcmax / scale) * bcmis;
        if (z >= small) {
            z = p + RT<R>::sqrt(scale) * RT<R>::sqrt(z) * sgn(p);
            a = d + z;
            d -= (bcmax / z) * bcmis;
            R tau = RT<R>::hypot(c, z);
            cs = z / tau;
            sn = c / tau;
            b -= c;
            c = zero;
        } else {
            R sigma = b + c;
            R tau = RT<R>::hypot(sigma, asubd);
            cs = RT<R>::sqrt(half * (one + RT<R>::abs(sigma) / tau));
            sn = -(p / (tau * cs)) * sgn(sigma);
            R aa = a * cs + b * sn, bb = -a * sn + b * cs;
            R cc = c * cs + d * sn, dd = -c * sn + d * cs;
            a = aa * cs + cc * sn;
            b = bb * cs + dd * sn;
            c = -aa * sn + cc * cs;
            d = -bb * sn + dd * cs;
            R midad = half * (a + d);
            a = midad;
            d = a;
            if (c != zero) {
                if (b != zero) {
                    if (b * c >= zero) {
                        R sab = RT<R>::sqrt(RT<R>::abs(b)), sac = RT<R>::sqrt(RT<R>::abs(c));
                        p = sab * sac * sgn(c);
                        tau = one / RT<R>::sqrt(RT<R>::abs(b + c));
                        a = midad + p;
                        d = midad - p;
                        b -= c;
                        c = zero;
                        R cs1 = sab * tau, sn1 = sac * tau;
                        R csn = cs * cs1 - sn * sn1, snn = cs * sn1 + sn * cs1;
                        cs = csn;
                        sn = snn;
                    }
                } else {
                    b = -c;
                    c = zero;
                    R cs0 = cs;
                    cs = -sn;
                    sn = cs0;
                }
            }
        }
    }
    if (c == zero) {
        w1 = Cx<R>(a, zero);
        w2 = Cx<R>(d, zero);
    } else {
        R rti = RT<R>::sqrt(RT<R>::abs(b)) * RT<R>::sqrt(RT<R>::abs(c));
        w1 = Cx<R>(a, rti);
        w2 = Cx<R>(d, -rti);
    }
}

// src/GenericSchur.jl:837-952
template <class R>
void doubleShiftQR(Mat<R> H, Mat<R>* Z, const Cx<R>& shift1, const Cx<R>& shift2, int istart, int iend,
                   Stats* st) {
    int n = H.n;
    R zero(0.0), one(1.0);
    int i1 = 1, i2 = n;
    R r1r = shift1.re, r1i = shift1.im, r2r = shift2.re, r2i = shift2.im;
    R v[3] = {zero, zero, zero};
    int mx = istart;
    for (int m = iend - 2; m >= istart; --m) {
        R H21s = H(m + 1, m);
        R s = RT<R>::abs(H(m, m) - r2r) + RT<R>::abs(r2i) + RT<R>::abs(H21s);
        H21s /= s;
        v[0] = H21s * H(m, m + 1) + (H(m, m) - r1r) * ((H(m, m) - r2r) / s) - r1i * (r2i / s);
        v[1] = H21s * (H(m, m) + H(m + 1, m + 1) - r1r - r2r);
        v[2] = H21s * H(m + 2, m + 1);
        s = RT<R>::abs(v[0]) + RT<R>::abs(v[1]) + RT<R>::abs(v[2]);
        v[0] /= s;
        v[1] /= s;
        v[2] /= s;
        if (m > istart &&
            (RT<R>::abs(H(m, m - 1)) * (RT<R>::abs(v[1]) + RT<R>::abs(v[2])) <=
             RT<R>::eps() * RT<R>::abs(v[0]) *
                 (RT<R>::abs(H(m - 1, m - 1)) + RT<R>::abs(H(m, m)) + RT<R>::abs(H(m + 1, m + 1))))) {
            mx = m;
            break;
        }
    }
    if (st) st->sweeps += 1;
    for (int k = mx; k <= iend - 1; ++k) {
        int nr = (iend - k + 1 < 3) ? iend - k + 1 : 3;
        if (k > mx)
            for (int ii = 0; ii < nr; ++ii) v[ii] = H(k + ii, k - 1);
        R tau1 = reflector(v, nr);
        if (k > mx) {
            H(k, k - 1) = v[0];
            H(k + 1, k - 1) = zero;
            if (k < iend - 1) H(k + 2, k - 1) = zero;
        } else if (mx > istart) {
            H(k, k - 1) *= (one - tau1);
        }
        R v2 = v[1];
        R tau2 = tau1 * v2;
        if (nr == 3) {
            R v3 = v[2];
            R tau3 = tau1 * v3;
            for (int j = k; j <= i2; ++j) {
                R ss = H(k, j) + v2 * H(k + 1, j) + v3 * H(k + 2, j);
                H(k, j) -= ss * tau1;
                H(k + 1, j) -= ss * tau2;
                H(k + 2, j) -= ss * tau3;
            }
            int jmax = (k + 3 < iend) ? k + 3 : iend;
            for (int j = i1; j <= jmax; ++j) {
                R ss = H(j, k) + v2 * H(j, k + 1) + v3 * H(j, k + 2);
                H(j, k) -= ss * tau1;
                H(j, k + 1) -= ss * tau2;
                H(j, k + 2) -= ss * tau3;
            }
            if (Z) {
                for (int j = 1; j <= Z->m; ++j) {
                    R ss = (*Z)(j, k) + v2 * (*Z)(j, k + 1) + v3 * (*Z)(j, k + 2);
                    (*Z)(j, k) -= ss * tau1;
                    (*Z)(j, k + 1) -= ss * tau2;
                    (*Z)(j, k + 2) -= ss * tau3;
                }
            }
            if (st) {
                st->applications += 1;
                st->rowpairs += (i2 - k + 1) + jmax + (Z ? Z->m : 0);
            }
        } else if (nr == 2) {
            for (int j = k; j <= i2; ++j) {
                R ss = H(k, j) + v2 * H(k + 1, j);
                H(k, j) -= ss * tau1;
                H(k + 1, j) -= ss * tau2;
            }
            for (int j = i1; j <= iend; ++j) {
                R ss = H(j, k) + v2 * H(j, k + 1);
                H(j, k) -= ss * tau1;
                H(j, k + 1) -= ss * tau2;
            }
            if (Z) {
                for (int j = 1; j <= Z->m; ++j) {
                    R ss = (*Z)(j, k) + v2 * (*Z)(j, k + 1);
                    (*Z)(j, k) -= ss * tau1;
                    (*Z)(j, k + 1) -= ss * tau2;
                }
            }
            if (st) st->applications += 1;
        }
    }
}

// src/GenericSchur.jl:513-699.  HH in place -> quasi-triangular T; w[n] complex eigenvalues.
template <class R>
int gschur_hess_real(Mat<R> HH, Mat<R>* Z, Cx<R>* w, int maxiter, Stats* st) {
    int n = HH.n;
    if (maxiter <= 0) maxiter = 100 * n;
    R zero(0.0);
    int istart = 1, iend = n;
    for (int j = 1; j <= n; ++j)                // triu!(HH, -1)
        for (int i = j + 2; i <= n; ++i) HH(i, j) = zero;
    int iwcur = n;
    R eps = RT<R>::eps();
    R smallnum = RT<R>::floatmin() * (R((double)n) / eps);
    R threeq(0.75), m7_16(-0.4375);
    int iter = 0;
    while (iend >= 1) {
        istart = 1;
        int iterqr = 0;
        bool deflate = false;
        while (true) {
            iter += 1;
            if (iter > maxiter) throw Unconverged{maxiter};
            bool split = false;
            for (int k = iend; k >= istart + 1; --k) {
                if (RT<R>::abs(HH(k, k - 1)) < smallnum) {
                    split = true;
                } else {
                    R Hkk = HH(k, k), Hk1 = HH(k - 1, k - 1);
                    R t = RT<R>::abs(Hk1) + RT<R>::abs(Hkk);
                    if (t == zero) {
                        if (k > 2) t += RT<R>::abs(HH(k - 1, k - 2));
                        if (k + 1 <= n) t += RT<R>::abs(HH(k + 1, k));
                    }
                    R aHkk1 = RT<R>::abs(HH(k, k - 1));
                    if (aHkk1 <= t * eps) {
                        R aHk1k = RT<R>::abs(HH(k - 1, k));
                        R ab = rmax(aHkk1, aHk1k), ba = rmin(aHkk1, aHk1k);
                        R aa = rmax(RT<R>::abs(Hkk), RT<R>::abs(Hk1 - Hkk));
                        R bb = rmin(RT<R>::abs(Hkk), RT<R>::abs(Hk1 - Hkk));
                        R s = aa + bb;   // src/GenericSchur.jl:586 (LAPACK has aa + ab)
                        if (ba * (ab / s) <= rmax(smallnum, eps * (bb * (aa / s)))) split = true;
                    }
                }
                if (split) {
                    istart = k;
                    break;
                }
            }
            if (!split) istart = 1;
            if (istart > 1) HH(istart, istart - 1) = zero;
            if (istart >= iend - 1) {
                deflate = true;
                break;
            }
            iterqr += 1;
            R H11, H12, H21, H22;
            if (iterqr == 10) {
                R s = RT<R>::abs(HH(istart + 1, istart)) + RT<R>::abs(HH(istart + 2, istart + 1));
                H11 = threeq * s + HH(istart, istart);
                H12 = m7_16 * s;
                H21 = s;
                H22 = H11;
                if (st) st->exceptional += 1;
            } else if (iterqr == 20) {
                R s = RT<R>::abs(HH(iend, iend - 1)) + RT<R>::abs(HH(iend - 1, iend - 2));
                H11 = threeq * s + HH(iend, iend);
                H12 = m7_16 * s;
                H21 = s;
                H22 = H11;
                if (st) st->exceptional += 1;
            } else {
                H11 = HH(iend - 1, iend - 1);
                H21 = HH(iend, iend - 1);
                H12 = HH(iend - 1, iend);
                H22 = HH(iend, iend);
            }
            R s = RT<R>::abs(H11) + RT<R>::abs(H12) + RT<R>::abs(H21) + RT<R>::abs(H22);
            R r1r(0.0), r2r(0.0), r1i(0.0), r2i(0.0);
            if (!(s == zero)) {
                H11 /= s;
                H12 /= s;
                H21 /= s;
                H22 /= s;
                R tr = (H11 + H22) / R(2.0);
                R d = (H11 - tr) * (H22 - tr) - H12 * H21;
                R rtd = RT<R>::sqrt(RT<R>::abs(d));
                if (d >= zero) {
                    r1r = tr * s;
                    r2r = r1r;
                    r1i = rtd * s;
                    r2i = -r1i;
                } else {
                    r1r = tr + rtd;
                    r2r = tr - rtd;
                    if (RT<R>::abs(r1r - H22) <= RT<R>::abs(r2r - H22)) {
                        r1r *= s;
                        r2r = r1r;
                    } else {
                        r2r *= s;
                        r1r = r2r;
                    }
                    r1i = zero;
                    r2i = zero;
                }
            }
            doubleShiftQR(HH, Z, Cx<R>(r1r, r1i), Cx<R>(r2r, r2i), istart, iend, st);
        }
        if (deflate && istart >= iend) {
            w[iwcur - 1] = Cx<R>(HH(iend, iend), zero);
            iwcur -= 1;
            iend = istart - 1;
        } else if (deflate && istart + 1 == iend) {
            R a = HH(iend - 1, iend - 1), b = HH(iend - 1, iend), c = HH(iend, iend - 1), d = HH(iend, iend);
            R cs, sn;
            Cx<R> w1, w2;
            gs2x2(a, b, c, d, cs, sn, w1, w2);
            w[iwcur - 1] = w2;
            iwcur -= 1;
            w[iwcur - 1] = w1;
            iwcur -= 1;
            // lmul!(G2, view(HH, :, istart:n)):  rows (iend-1, iend) <- (c a1 + s a2, -s a1 + c a2)
            for (int j = istart; j <= n; ++j) {
                R a1 = HH(iend - 1, j), a2 = HH(iend, j);
                HH(iend - 1, j) = cs * a1 + sn * a2;
                HH(iend, j) = -sn * a1 + cs * a2;
            }
            // rmul!(view(HH, 1:iend, :), G2'): cols (iend-1, iend) <- (a1 c + a2 s, -a1 s + a2 c)
            for (int i = 1; i <= iend; ++i) {
                R a1 = HH(i, iend - 1), a2 = HH(i, iend);
                HH(i, iend - 1) = a1 * cs + a2 * sn;
                HH(i, iend) = -a1 * sn + a2 * cs;
            }
            HH(iend - 1, iend - 1) = a;
            HH(iend - 1, iend) = b;
            HH(iend, iend - 1) = c;
            HH(iend, iend) = d;
            if (iend > 2) HH(iend - 1, iend - 2) = zero;
            if (Z) {
                for (int i = 1; i <= Z->m; ++i) {
                    R a1 = (*Z)(i, iend - 1), a2 = (*Z)(i, iend);
                    (*Z)(i, iend - 1) = a1 * cs + a2 * sn;
                    (*Z)(i, iend) = -a1 * sn + a2 * cs;
                }
            }
        }
        iend = istart - 1;
    }
    return 0;
}

// src/GenericSchur.jl:805-835
template <class R>
int gschur_real(Mat<R> A, Mat<R>* Z, Cx<R>* w, bool scale, int maxiter, Stats* st) {
    int n = A.n;
    if (n == 0) return 0;
    ScaleInfo<R> si;
    si.scaled = false;
    if (scale) si = scale_matrix(A);
    std::vector<R> tau(n > 1 ? n - 1 : 1);
    hessenberg(A, tau.data());
    if (Z) materializeQ(A, tau.data(), *Z);
    int rc = gschur_hess_real(A, Z, w, maxiter, st);
    if (si.scaled) {
        for (int j = 1; j <= n; ++j) {
            int rows = (j + 1 < n) ? j + 1 : n;
            safescale(&A(1, j), (long)rows, si.cscale, si.anrm);
        }
        safescale(w, (long)n, si.cscale, si.anrm);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Checker side: residuals and eigenvalue condition numbers
// ---------------------------------------------------------------------------------------------

// reciprocal eigenvalue condition numbers s_i = |y_i' x_i| / (|x_i| |y_i|) from an upper triangular T,
// i.e. `rconde` of src/pirates.jl:109-113 with x, y from the triangular back-substitutions of
// src/vectors.jl:45-131 (right) and :372-460 (left); small pivots are perturbed as there.
template <class R> void eigvalscond_triu(Mat<Cx<R>> T, R* s) {
    typedef Cx<R> C;
    int n = T.n;
    R ulp = RT<R>::eps();
    R smallnum = safemin<R>() * (R((double)n) / ulp);
    std::vector<C> x(n + 1), y(n + 1);
    for (int ki = 1; ki <= n; ++ki) {
        C lam = T(ki, ki);
        R tnorm(0.0);
        for (int j = 1; j <= n; ++j)
            for (int i = 1; i <= j; ++i) tnorm = rmax(tnorm, abs1(T(i, j)));
        R smin = rmax(ulp * tnorm, smallnum);
        // right: (T[1:ki-1,1:ki-1] - lam I) x = -T[1:ki-1, ki], x_ki = 1
        x[ki] = C(R(1.0));
        for (int i = ki - 1; i >= 1; --i) {
            C acc = -T(i, ki);
            for (int j = i + 1; j <= ki - 1; ++j) acc -= T(i, j) * x[j];
            C piv = T(i, i) - lam;
            if (abs1(piv) < smin) piv = C(smin);
            x[i] = acc / piv;
        }
        // left: y' (T - lam I) = 0, y_ki = 1, entries ki+1..n
        y[ki] = C(R(1.0));
        for (int j = ki + 1; j <= n; ++j) {
            C acc(R(0.0));
            for (int i = ki; i <= j - 1; ++i) acc -= conj_(T(i, j)) * y[i];
            C piv = conj_(T(j, j) - lam);
            if (abs1(piv) < smin) piv = C(smin);
            y[j] = acc / piv;
        }
        R xn(0.0), yn(0.0);
        for (int i = 1; i <= ki; ++i) xn += x[i].re * x[i].re + x[i].im * x[i].im;
        for (int i = ki; i <= n; ++i) yn += y[i].re * y[i].re + y[i].im * y[i].im;
        // y' x: only index ki overlaps -> 1
        s[ki - 1] = R(1.0) / (RT<R>::sqrt(xn) * RT<R>::sqrt(yn));
    }
}

// ---------------------------------------------------------------------------------------------
// balance!(A; scale, permute) — src/balance.jl:33-199 (algo = :pr, p = 1) — and lmul!(B, V) / ldiv!(B, V) :203-260.
// Float64 / ComplexF64 only.  The two-norms are formed the way the CUDA kernel documents (per-"lane" partial sums
// over indices of stride 32, xor butterfly, unfused operations; amax = sqrt(max |x|^2) as _findamax for complex,
// src/norm1est.jl:76-100): norm(view, 2) of the reference differs from this by rounding only, and with the
// operation order fixed the power-of-two decisions of the two implementations can be compared exactly.
// ---------------------------------------------------------------------------------------------
inline double bal_abs2(double x) { return x * x; }
inline double bal_abs2(const Cx<double>& x) { return x.re * x.re + x.im * x.im; }
inline double bal_sabs2(double x, double s) { double t = x * s; return t * t; }
inline double bal_sabs2(const Cx<double>& x, double s) { double a = x.re * s, b = x.im * s; return a * a + b * b; }
inline double bal_mod(double x) { return std::fabs(x); }
inline double bal_mod(const Cx<double>& x) { return std::hypot(x.re, x.im); }
inline bool bal_nz(double x) { return x != 0.0; }
inline bool bal_nz(const Cx<double>& x) { return x.re != 0.0 || x.im != 0.0; }
inline double bal_times(double x, double f) { return x * f; }
inline Cx<double> bal_times(const Cx<double>& x, double f) { return Cx<double>(x.re * f, x.im * f); }

template <class E> void bal_norms(const E* x, long st, int len, double& amax, double& nrm) {
    double part[32];
    auto butterfly_sum = [&]() {
        for (int m = 16; m >= 1; m >>= 1) {
            double nw[32];
            for (int l = 0; l < 32; ++l) nw[l] = part[l] + part[l ^ m];
            for (int l = 0; l < 32; ++l) part[l] = nw[l];
        }
        return part[0];
    };
    double m2 = 0.0;
    for (int t = 0; t < len; ++t) m2 = std::fmax(m2, bal_abs2(x[(long)t * st]));
    amax = std::sqrt(m2);
    if (!(amax > 0.0) || !(amax < 1.7976931348623157e308)) {
        double mx = 0.0;
        bool nan = false;
        for (int t = 0; t < len; ++t) {
            double a = bal_mod(x[(long)t * st]);
            if (a != a) nan = true;
            mx = std::fmax(mx, a);
        }
        amax = nan ? std::nan("") : mx;
    }
    if (!(amax > 0.0)) {
        nrm = amax;
        return;
    }
    const double s = 1.0 / amax;
    for (int l = 0; l < 32; ++l) part[l] = 0.0;
    for (int t = 0; t < len; ++t) part[t & 31] = part[t & 31] + bal_sabs2(x[(long)t * st], s);
    nrm = amax * std::sqrt(butterfly_sum());
}

// returns 0, or -5 for the reference's error("NaN encountered while balancing")
template <class E> int balance(Mat<E> A, bool scale, bool permute, double* D, int* sp, int& ilo, int& ihi, bool& trivial) {
    const int n = A.n;
    for (int i = 0; i < n; ++i) { D[i] = 1.0; sp[i] = 0; }
    ilo = 1; ihi = n; trivial = true;
    auto swap_rc = [&](int js, int ms) {
        for (int i = 1; i <= ihi; ++i) std::swap(A(i, js), A(i, ms));
        for (int i = ilo; i <= n; ++i) std::swap(A(js, i), A(ms, i));
    };
    if (permute) {
        ihi = n + 1;
        while (ihi > 1) {
            ihi -= 1;
            bool exch = false;
            int js = 0, ms = 0;
            for (int j = ihi; j >= 1; --j) {
                exch = true;
                js = j;
                for (int i = 1; i <= ihi; ++i) {
                    if (i == j) continue;
                    if (bal_nz(A(j, i))) exch = false;
                }
                if (exch) { ms = ihi; break; }
            }
            if (exch) {
                sp[ms - 1] = js;
                if (js != ms) { trivial = false; swap_rc(js, ms); }
            } else break;
        }
        if (ihi > 1) {
            ilo = 0;
            while (ilo < n) {
                ilo += 1;
                bool exch = false;
                int js = 0, ms = 0;
                for (int j = ilo; j <= ihi; ++j) {
                    js = j;
                    exch = true;
                    for (int i = ilo; i <= ihi; ++i) {
                        if (i == j) continue;
                        if (bal_nz(A(i, j))) exch = false;
                    }
                    if (exch) { ms = ilo; break; }
                }
                if (exch) {
                    sp[ms - 1] = js;
                    if (ms != js) { trivial = false; swap_rc(js, ms); }
                } else break;
            }
        }
    }
    if (scale) {
        const double beta = 2.0, factor = 0.95;
        const double sfmin1 = 2.2250738585072014e-308 / 2.220446049250313e-16;
        const double sfmin2 = sfmin1 * beta, sfmax2 = 1.0 / sfmin2;
        bool converged = false;
        int guard = 0;
        while (!converged && guard++ < 10000) {
            converged = true;
            for (int i = ilo; i <= ihi; ++i) {
                double c, r, ca, ra;
                const int len = ihi - ilo + 1;
                bal_norms<E>(&A(ilo, i), 1, len, ca, c);
                bal_norms<E>(&A(i, ilo), A.ld, len, ra, r);
                if (c == 0.0 || r == 0.0) continue;
                double g = r / beta;
                const double s = c + r;
                double f = 1.0;
                while (c < r / beta) {
                    if (c >= g || (std::fmax(f, std::fmax(c, ca)) >= sfmax2) || (std::fmin(r, std::fmin(g, ra)) <= sfmin2)) break;
                    const double chk = c + f + ca + r + g + ra;
                    if (chk != chk) return -5;
                    f *= beta; c *= beta; ca *= beta; r /= beta; g /= beta; ra /= beta;
                }
                g = c / beta;
                while (r <= c / beta) {
                    if ((g < r) || (std::fmax(r, ra) >= sfmax2) || (std::fmin(std::fmin(f, c), std::fmin(g, ca)) <= sfmin2)) break;
                    f /= beta; c /= beta; g /= beta; ca /= beta; r *= beta; ra *= beta;
                }
                if (f != 1.0) trivial = false;
                if (c + r >= factor * s) continue;
                converged = false;
                D[i - 1] *= f;
                const double rf = 1.0 / f;
                for (int j = ilo; j <= n; ++j) A(i, j) = bal_times(A(i, j), rf);
                for (int j = 1; j <= ihi; ++j) A(j, i) = bal_times(A(j, i), f);
            }
        }
    }
    return 0;
}

// lmul!(B, V) (inv = false) / ldiv!(B, V) (inv = true), src/balance.jl:203-260
template <class E> void balance_apply(Mat<E> V, const double* D, const int* sp, int ilo, int ihi, bool trivial, bool inv) {
    const int n = V.n;
    if (trivial) return;
    if (ilo != ihi)
        for (int j = 1; j <= n; ++j)
            for (int i = 1; i <= n; ++i) V(i, j) = bal_times(V(i, j), inv ? 1.0 / D[i - 1] : D[i - 1]);
    for (int j = ilo - 1; j >= 1; --j) {
        const int m = sp[j - 1];
        if (m == j) continue;
        for (int i = 1; i <= n; ++i) std::swap(V(j, i), V(m, i));
    }
    for (int j = ihi + 1; j <= n; ++j) {
        const int m = sp[j - 1];
        if (m == j) continue;
        for (int i = 1; i <= n; ++i) std::swap(V(j, i), V(m, i));
    }
}

}  // namespace gso
